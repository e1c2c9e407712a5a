// Minimal stand-in for the Eigen types the glue uses (column-major dense matrix / vector with data()). TEST INFRASTRUCTURE ONLY.
#pragma once
#include <vector>
namespace Eigen {
class MatrixXd {
   public:
    std::vector<double> v;
    long r = 0, c = 0;
    MatrixXd() {}
    MatrixXd(long rows, long cols) : v((size_t)(rows * cols)), r(rows), c(cols) {}
    void resize(long rows, long cols) { v.resize((size_t)(rows * cols)); r = rows; c = cols; }
    long rows() const { return r; }
    long cols() const { return c; }
    double* data() { return v.data(); }
};
class VectorXd {
   public:
    std::vector<double> v;
    VectorXd() {}
    explicit VectorXd(long n) : v((size_t)n) {}
    void resize(long n) { v.resize((size_t)n); }
    double* data() { return v.data(); }
};
}  // namespace Eigen
