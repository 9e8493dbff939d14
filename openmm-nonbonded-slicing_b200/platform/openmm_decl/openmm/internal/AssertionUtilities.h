#ifndef NBS_DECL_ASSERTION_H_
#define NBS_DECL_ASSERTION_H_
#include <string>
namespace OpenMM { void throwException(const char* file, int line, const std::string& details); }
#endif
