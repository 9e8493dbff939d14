#ifndef NBS_DECL_KERNEL_H_
#define NBS_DECL_KERNEL_H_
#include "openmm/KernelImpl.h"
namespace OpenMM {
class Kernel {
public:
    Kernel();
    ~Kernel();
    template <class T> T& getAs() { return dynamic_cast<T&>(*impl); }
private:
    KernelImpl* impl;
};
}
#endif
