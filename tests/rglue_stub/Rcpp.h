// Minimal stand-in for the parts of <Rcpp.h> / the R C API that rglue/singlet_cuda_glue.cpp touches, so that the glue can be
// type-checked (g++ -fsyntax-only) against include/singlet_cuda.h in an image without R. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <initializer_list>
#include <string>
#include <vector>

typedef enum { FALSE = 0, TRUE = 1 } Rboolean;
inline void Rprintf(const char*, ...) {}
inline void R_CheckUserInterrupt() {}
inline Rboolean R_ToplevelExec(void (*fn)(void*), void* data) { fn(data); return TRUE; }
#define ISNAN(x) (std::isnan(x))

namespace Rcpp {
[[noreturn]] inline void stop(const char*) { throw 1; }
[[noreturn]] inline void stop(const std::string&) { throw 1; }

template <typename T>
class Vector {
   public:
    std::vector<T> v;
    Vector() {}
    explicit Vector(int n) : v((size_t)n) {}
    template <typename It>
    Vector(It b, It e) : v(b, e) {}
    T* begin() { return v.data(); }
    const T* begin() const { return v.data(); }
    T& operator[](int i) { return v[(size_t)i]; }
    const T& operator[](int i) const { return v[(size_t)i]; }
    int size() const { return (int)v.size(); }
};
typedef Vector<double> NumericVector;
typedef Vector<int> IntegerVector;

class SparseMatrix {  // reference inst/include/singlet.h:36-102
   public:
    IntegerVector i, p, Dim;
    NumericVector x;
    int rows() const { return Dim[0]; }
    int cols() const { return Dim[1]; }
};

struct Object {  // any R object
    template <typename T>
    Object(const T&) {}
    Object() {}
};
struct NamedArg {
    Object value;
};
struct Named {
    explicit Named(const char*) {}
    template <typename T>
    NamedArg operator=(const T& t) const { return NamedArg{Object(t)}; }
};
class List {
   public:
    std::vector<Object> items;
    List() {}
    explicit List(int n) : items((size_t)n) {}
    template <typename... A>
    static List create(const A&...) { return List(); }
    int size() const { return (int)items.size(); }
    Object& operator[](int i) { return items[(size_t)i]; }
    std::vector<Object>::iterator begin() { return items.begin(); }
    std::vector<Object>::iterator end() { return items.end(); }
};
template <typename T>
T as(const Object&) { return T(); }
}  // namespace Rcpp
