// Error type of the drop-in classes: the same name and interface as the reference's
// include/exception_handler.hpp:10-31 (a std::exception carrying a text set with set()), so code that
// catches StandardException or std::exception around CMatrix / CMatrixGenerator keeps working.
// The include guard is the reference's on purpose: whichever header is seen first wins.
#ifndef COSMO_PP_EXCEPTION_HANDLER_HPP
#define COSMO_PP_EXCEPTION_HANDLER_HPP

#include <exception>
#include <string>

class StandardException : public std::exception
{
public:
    StandardException() {}
    explicit StandardException(const std::string& text) : text_(text) {}
    ~StandardException() throw() {}

    void set(const std::string& text) { text_ = text; }
    virtual const char* what() const throw() { return text_.c_str(); }

private:
    std::string text_;
};

#endif
