// eigen_rcpp_shim.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/singlet_oracle.cpp header).
//
// The reference's kernels (src/singlet.cpp) are written against RcppEigen and Rcpp, neither of
// which exists in this image. This header provides the *minimum* surface of those two libraries
// that the hot-path functions touch, so that oracle/Makefile can compile the reference's own
// function bodies -- extracted at build time from /root/reference/src/singlet.cpp into
// oracle/_ref/, never committed -- into oracle/_ref/libsinglet_ref.so.
//
// Only the operations used at the call sites listed in SURVEY.md 8(c) are implemented:
//   Eigen::MatrixXd / VectorXd (column-major), Zero/Ones, (i,j), col(), row(), transpose(),
//   selfadjointView<Lower>().rankUpdate(), triangularView<Upper>() = X.transpose(),
//   diagonal().array() += s, rowwise().sum(), array() += s, vector axpy with scaled columns,
//   matrix difference, row * col dot product.
//   Rcpp::SparseMatrix (+InnerIterator) with the interface of inst/include/singlet.h:36-102,
//   Rcpp::List::create / Rcpp::Named, NumericVector / IntegerVector push_back, Rcpp::min,
//   Rprintf, Rcpp::checkUserInterrupt.
// Reductions are plain left-to-right sums (Eigen's SIMD summation order cannot be reproduced
// without Eigen).
#pragma once
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace Eigen {

enum { Lower = 1, Upper = 2 };

class MatrixXd;
class VectorXd;

struct ColRef {
    double* ptr;
    long n;
};
struct ConstColRef {
    const double* ptr;
    long n;
};
struct ScaledCol {
    const double* ptr;
    long n;
    double s;
};
struct ConstRowRef {
    const double* ptr;  // element (row, 0)
    long stride, n;
};

inline ScaledCol operator*(double s, const ConstColRef& c) { return {c.ptr, c.n, s}; }
inline ScaledCol operator*(const ConstColRef& c, double s) { return {c.ptr, c.n, s}; }
inline ScaledCol operator*(double s, const ColRef& c) { return {c.ptr, c.n, s}; }
inline ScaledCol operator*(const ColRef& c, double s) { return {c.ptr, c.n, s}; }
inline double operator*(const ConstRowRef& r, const ConstColRef& c) {
    double s = 0;
    for (long t = 0; t < r.n; ++t) s += r.ptr[t * r.stride] * c.ptr[t];
    return s;
}
inline double operator*(const ConstRowRef& r, const ColRef& c) { return r * ConstColRef{c.ptr, c.n}; }

struct ArrayProxy {
    double* ptr;
    long n, stride;
    void operator+=(double s) {
        for (long t = 0; t < n; ++t) ptr[t * stride] += s;
    }
    void operator*=(double s) {
        for (long t = 0; t < n; ++t) ptr[t * stride] *= s;
    }
    void operator-=(double s) {
        for (long t = 0; t < n; ++t) ptr[t * stride] -= s;
    }
};
struct DiagProxy {
    double* ptr;
    long n, stride;
    ArrayProxy array() { return {ptr, n, stride}; }
};

struct TransposeOf {
    const MatrixXd* m;
};

class VectorXd {
   public:
    std::vector<double> v;
    VectorXd() {}
    explicit VectorXd(long n) : v((size_t)n, 0.0) {}
    static VectorXd Zero(long n) { return VectorXd(n); }
    static VectorXd Ones(long n) {
        VectorXd r(n);
        for (auto& e : r.v) e = 1.0;
        return r;
    }
    long size() const { return (long)v.size(); }
    long rows() const { return (long)v.size(); }
    double& operator()(long i) { return v[(size_t)i]; }
    double operator()(long i) const { return v[(size_t)i]; }
    double& operator[](long i) { return v[(size_t)i]; }
    double operator[](long i) const { return v[(size_t)i]; }
    double* data() { return v.data(); }
    const double* data() const { return v.data(); }
    ArrayProxy array() { return {v.data(), (long)v.size(), 1}; }
    double sum() const {
        double s = 0;
        for (double e : v) s += e;
        return s;
    }
    VectorXd& setOnes() {
        for (auto& e : v) e = 1.0;
        return *this;
    }
    VectorXd& operator+=(const ScaledCol& c) {
        for (long t = 0; t < c.n; ++t) v[(size_t)t] += c.s * c.ptr[t];
        return *this;
    }
    VectorXd& operator-=(const ScaledCol& c) {
        for (long t = 0; t < c.n; ++t) v[(size_t)t] -= c.ptr[t] * c.s;
        return *this;
    }
};

struct RowwiseProxy {
    const MatrixXd* m;
    VectorXd sum() const;
};

struct ColAssign {
    double* ptr;
    long n;
    ColAssign& operator=(const ConstColRef& c) {
        for (long t = 0; t < n; ++t) ptr[t] = c.ptr[t];
        return *this;
    }
    ColAssign& operator=(const ColAssign& c) {
        for (long t = 0; t < n; ++t) ptr[t] = c.ptr[t];
        return *this;
    }
    operator ConstColRef() const { return {ptr, n}; }
};
inline ScaledCol operator*(double s, const ColAssign& c) { return {c.ptr, c.n, s}; }
inline ScaledCol operator*(const ColAssign& c, double s) { return {c.ptr, c.n, s}; }
inline double operator*(const ConstRowRef& r, const ColAssign& c) { return r * ConstColRef{c.ptr, c.n}; }

template <int UpLo>
struct SelfAdjointProxy {
    MatrixXd* m;
    void rankUpdate(const MatrixXd& A);
};
template <int UpLo>
struct TriangularProxy {
    MatrixXd* m;
    void operator=(const TransposeOf& t);
};

class MatrixXd {
   public:
    std::vector<double> v;
    long r_ = 0, c_ = 0;
    MatrixXd() {}
    MatrixXd(long r, long c) : v((size_t)r * (size_t)c, 0.0), r_(r), c_(c) {}
    MatrixXd(const TransposeOf& t) { *this = t; }
    static MatrixXd Zero(long r, long c) { return MatrixXd(r, c); }
    long rows() const { return r_; }
    long cols() const { return c_; }
    long size() const { return r_ * c_; }
    double* data() { return v.data(); }
    const double* data() const { return v.data(); }
    double& operator()(long i, long j) { return v[(size_t)j * (size_t)r_ + (size_t)i]; }
    double operator()(long i, long j) const { return v[(size_t)j * (size_t)r_ + (size_t)i]; }
    ColAssign col(long j) { return {v.data() + (size_t)j * (size_t)r_, r_}; }
    ConstColRef col(long j) const { return {v.data() + (size_t)j * (size_t)r_, r_}; }
    ConstRowRef row(long i) const { return {v.data() + i, r_, c_}; }
    TransposeOf transpose() const { return {this}; }
    MatrixXd& operator=(const TransposeOf& t) {
        const MatrixXd& s = *t.m;
        MatrixXd out(s.c_, s.r_);
        for (long j = 0; j < s.c_; ++j)
            for (long i = 0; i < s.r_; ++i) out(j, i) = s(i, j);
        v.swap(out.v);
        r_ = out.r_;
        c_ = out.c_;
        return *this;
    }
    MatrixXd& setZero() {
        for (auto& e : v) e = 0.0;
        return *this;
    }
    DiagProxy diagonal() { return {v.data(), r_ < c_ ? r_ : c_, r_ + 1}; }
    RowwiseProxy rowwise() const { return {this}; }
    template <int UpLo>
    SelfAdjointProxy<UpLo> selfadjointView() {
        return {this};
    }
    template <int UpLo>
    TriangularProxy<UpLo> triangularView() {
        return {this};
    }
};

inline VectorXd RowwiseProxy::sum() const {
    VectorXd out(m->rows());
    for (long i = 0; i < m->rows(); ++i) {
        double s = 0;
        for (long j = 0; j < m->cols(); ++j) s += (*m)(i, j);
        out(i) = s;
    }
    return out;
}

// lower triangle += A A^T
template <int UpLo>
inline void SelfAdjointProxy<UpLo>::rankUpdate(const MatrixXd& A) {
    static_assert(UpLo == Lower, "shim implements the Lower rank update only");
    for (long i = 0; i < A.rows(); ++i)
        for (long j = 0; j <= i; ++j) {
            double s = 0;
            for (long t = 0; t < A.cols(); ++t) s += A(i, t) * A(j, t);
            (*m)(i, j) += s;
        }
}
// strictly-upper + diagonal part = that of the transposed source
template <int UpLo>
inline void TriangularProxy<UpLo>::operator=(const TransposeOf& t) {
    static_assert(UpLo == Upper, "shim implements the Upper assignment only");
    const MatrixXd& s = *t.m;
    // the reference assigns AAt.transpose() to AAt's own upper view: read lower, write upper
    for (long j = 0; j < m->cols(); ++j)
        for (long i = 0; i <= j; ++i) (*m)(i, j) = s(j, i);
}

// dense matrix-vector product (predict on dense input: `b = w * A.col(i)`), accumulated column by column
inline VectorXd operator*(const MatrixXd& M, const ConstColRef& c) {
    VectorXd out(M.rows());
    for (long j = 0; j < M.cols(); ++j)
        for (long i = 0; i < M.rows(); ++i) out(i) += c.ptr[j] * M(i, j);
    return out;
}
inline VectorXd operator*(const MatrixXd& M, const ColAssign& c) { return M * ConstColRef{c.ptr, c.n}; }

inline MatrixXd operator-(const MatrixXd& a, const MatrixXd& b) {
    MatrixXd out(a.rows(), a.cols());
    for (size_t t = 0; t < out.v.size(); ++t) out.v[t] = a.v[t] - b.v[t];
    return out;
}

}  // namespace Eigen

// ---------------------------------------------------------------------------------------------
namespace Rcpp {

// same interface as the reference's zero-copy dgCMatrix view (inst/include/singlet.h:36-102)
class SparseMatrix {
   public:
    const double* x = nullptr;
    const int* i = nullptr;
    const int* p = nullptr;
    int Dim[2] = {0, 0};
    SparseMatrix() {}
    SparseMatrix(const double* x_, const int* i_, const int* p_, int nrow, int ncol) : x(x_), i(i_), p(p_) {
        Dim[0] = nrow;
        Dim[1] = ncol;
    }
    unsigned int rows() { return (unsigned int)Dim[0]; }
    unsigned int cols() { return (unsigned int)Dim[1]; }
    class InnerIterator {
       public:
        InnerIterator(SparseMatrix& ptr, int col) : ptr(ptr), col_(col), index(ptr.p[col]), max_index(ptr.p[col + 1]) {}
        operator bool() const { return (index < max_index); }
        InnerIterator& operator++() {
            ++index;
            return *this;
        }
        double value() const { return ptr.x[index]; }
        int row() const { return ptr.i[index]; }
        int col() const { return col_; }

       private:
        SparseMatrix& ptr;
        int col_, index, max_index;
    };
};

template <typename T>
class VecT {
   public:
    std::vector<T> v;
    void push_back(T e) { v.push_back(e); }
    long size() const { return (long)v.size(); }
    T& operator()(long i) { return v[(size_t)i]; }
    T& operator[](long i) { return v[(size_t)i]; }
    T operator()(long i) const { return v[(size_t)i]; }
    T operator[](long i) const { return v[(size_t)i]; }
};
using NumericVector = VecT<double>;
using IntegerVector = VecT<int>;

inline double min(const NumericVector& x) {
    double m = x.v[0];
    for (double e : x.v) m = e < m ? e : m;
    return m;
}

struct Entry {
    std::string name;
    std::vector<double> data;
    long rows = 0, cols = 0;
};
struct Named {
    std::string name;
    explicit Named(const char* n) : name(n) {}
    Entry operator=(const Eigen::MatrixXd& m) const { return {name, m.v, m.rows(), m.cols()}; }
    Entry operator=(const Eigen::VectorXd& m) const { return {name, m.v, m.size(), 1}; }
    Entry operator=(const NumericVector& m) const { return {name, m.v, m.size(), 1}; }
    Entry operator=(const IntegerVector& m) const {
        Entry e{name, {}, m.size(), 1};
        for (int q : m.v) e.data.push_back((double)q);
        return e;
    }
};

class List {
   public:
    std::vector<Entry> entries;
    std::vector<SparseMatrix> mats;  // a "list of dgCMatrix" argument
    template <typename... Es>
    static List create(Es... es) {
        List l;
        (l.entries.push_back(es), ...);
        return l;
    }
    std::vector<SparseMatrix>::iterator begin() { return mats.begin(); }
    std::vector<SparseMatrix>::iterator end() { return mats.end(); }
    const Entry* find(const char* n) const {
        for (const auto& e : entries)
            if (e.name == n) return &e;
        return nullptr;
    }
};

inline void checkUserInterrupt() {}

}  // namespace Rcpp

inline void Rprintf(const char*, ...) {}
