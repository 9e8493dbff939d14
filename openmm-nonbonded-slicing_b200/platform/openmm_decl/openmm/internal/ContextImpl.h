#ifndef NBS_DECL_CONTEXTIMPL_H_
#define NBS_DECL_CONTEXTIMPL_H_
#include "openmm/System.h"
#include <string>
namespace OpenMM {
class ContextImpl {
public:
    const System& getSystem() const;
    void* getPlatformData();
    double getParameter(std::string name);
};
}
#endif
