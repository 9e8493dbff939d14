#ifndef NBS_DECL_PLATFORM_H_
#define NBS_DECL_PLATFORM_H_
#include <string>
namespace OpenMM {
class ContextImpl;
class KernelImpl;
class Platform;
class KernelFactory {
public:
    virtual ~KernelFactory() {}
    virtual KernelImpl* createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const = 0;
};
class Platform {
public:
    virtual ~Platform() {}
    virtual const std::string& getName() const = 0;
    static Platform& getPlatformByName(const std::string& name);
    void registerKernelFactory(const std::string& name, KernelFactory* factory);
};
}
#endif
