#ifndef NBS_DECL_EXCEPTION_H_
#define NBS_DECL_EXCEPTION_H_
#include <exception>
#include <string>
namespace OpenMM {
class OpenMMException : public std::exception {
public:
    explicit OpenMMException(const std::string& message) : message(message) {}
    const char* what() const throw() { return message.c_str(); }
private:
    std::string message;
};
}
#endif
