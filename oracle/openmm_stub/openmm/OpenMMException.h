// Minimal stand-in for openmm/OpenMMException.h -- TEST INFRASTRUCTURE ONLY.
#ifndef NBS_STUB_OPENMM_EXCEPTION_H_
#define NBS_STUB_OPENMM_EXCEPTION_H_
#include <exception>
#include <string>
namespace OpenMM {
class OpenMMException : public std::exception {
public:
    explicit OpenMMException(const std::string& message) : message(message) {}
    ~OpenMMException() throw() {}
    const char* what() const throw() { return message.c_str(); }
private:
    std::string message;
};
}
#endif
