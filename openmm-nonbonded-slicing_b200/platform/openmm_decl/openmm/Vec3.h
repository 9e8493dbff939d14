#ifndef NBS_DECL_VEC3_H_
#define NBS_DECL_VEC3_H_
namespace OpenMM {
class Vec3 {
public:
    Vec3() : v{0, 0, 0} {}
    Vec3(double x, double y, double z) : v{x, y, z} {}
    double operator[](int i) const { return v[i]; }
    double& operator[](int i) { return v[i]; }
private:
    double v[3];
};
}
#endif
