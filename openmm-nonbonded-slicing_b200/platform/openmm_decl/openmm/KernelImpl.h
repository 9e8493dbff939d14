#ifndef NBS_DECL_KERNELIMPL_H_
#define NBS_DECL_KERNELIMPL_H_
#include "openmm/Platform.h"
#include <string>
namespace OpenMM {
class KernelImpl {
public:
    KernelImpl(std::string name, const Platform& platform) : name(name), platform(&platform) {}
    virtual ~KernelImpl() {}
    std::string getName() const { return name; }
    const Platform& getPlatform() { return *platform; }
private:
    std::string name;
    const Platform* platform;
};
}
#endif
