// eigen_stub_ivsparse.hpp -- TEST INFRASTRUCTURE ONLY. Just enough declarations of the Eigen names that appear in the
// SIGNATURES of the reference's vendored IVSparse headers for them to parse without Eigen: the codec paths compiled into
// oracle/_ref/libivsparse_ref.so (raw-CSC constructor, compressCSC, write, file constructor, InnerIterator) never touch them.
#pragma once
#include <cmath>
#include <limits>
#include <tuple>
#include <unordered_map>
namespace Eigen {
enum { ColMajor = 0, RowMajor = 1 };
template <typename T, int Options = 0, typename Index = int> class SparseMatrix;
template <typename T, int Options = 0, typename Index = int> class SparseVector;
template <typename T, int R, int C> class Matrix;
template <typename T> class Triplet;
template <typename T> class Map;
}
namespace Eigen {
class VectorXd {  // only what the (never instantiated) level-1 CSC BLAS code needs to parse
   public:
    VectorXd() {}
    template <typename X> VectorXd(const X&) {}
    long rows() const { return 0; }
    double operator()(long) const { return 0.0; }
};
class MatrixXd;
class VectorXf;
class MatrixXf;
}
