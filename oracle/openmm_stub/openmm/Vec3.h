// Minimal stand-in for openmm/Vec3.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// Just enough of the OpenMM [external] API for the reference's Reference-platform
// arithmetic TUs to compile unmodified against it.
#ifndef NBS_STUB_OPENMM_VEC3_H_
#define NBS_STUB_OPENMM_VEC3_H_
#include <cassert>
#include <cmath>
namespace OpenMM {
class Vec3 {
public:
    Vec3() : v{0.0, 0.0, 0.0} {}
    Vec3(double x, double y, double z) : v{x, y, z} {}
    double operator[](int i) const { return v[i]; }
    double& operator[](int i) { return v[i]; }
    Vec3 operator+(const Vec3& o) const { return Vec3(v[0]+o.v[0], v[1]+o.v[1], v[2]+o.v[2]); }
    Vec3 operator-(const Vec3& o) const { return Vec3(v[0]-o.v[0], v[1]-o.v[1], v[2]-o.v[2]); }
    Vec3 operator-() const { return Vec3(-v[0], -v[1], -v[2]); }
    Vec3 operator*(double s) const { return Vec3(v[0]*s, v[1]*s, v[2]*s); }
    Vec3 operator/(double s) const { return Vec3(v[0]/s, v[1]/s, v[2]/s); }
    Vec3& operator+=(const Vec3& o) { v[0] += o.v[0]; v[1] += o.v[1]; v[2] += o.v[2]; return *this; }
    Vec3& operator-=(const Vec3& o) { v[0] -= o.v[0]; v[1] -= o.v[1]; v[2] -= o.v[2]; return *this; }
    Vec3& operator*=(double s) { v[0] *= s; v[1] *= s; v[2] *= s; return *this; }
    double dot(const Vec3& o) const { return v[0]*o.v[0] + v[1]*o.v[1] + v[2]*o.v[2]; }
private:
    double v[3];
};
static inline Vec3 operator*(double s, const Vec3& a) { return a*s; }
}
#endif
