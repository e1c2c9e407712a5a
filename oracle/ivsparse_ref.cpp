// ivsparse_ref.cpp -- TEST INFRASTRUCTURE ONLY. C entry points around the REFERENCE's own IVSparse codec, compiled from the
// vendored headers where they lie (/root/reference/inst/include/IVSparse.h) with the same instantiations the reference uses
// (src/singlet.cpp:3-4): writes file images with its compressCSC + write, reads them back with its file constructor and
// InnerIterator. Pins csrc/ivsparse.cpp (tests/test_ivsparse.py); never part of the product.
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "shim/eigen_stub_ivsparse.hpp"
#include "IVSparse.h"

using IVCSC = IVSparse::SparseMatrix<float, uint64_t, 3, true>;
using VCSC = IVSparse::SparseMatrix<float, uint64_t, 2, true>;

extern "C" int ref_ivsparse_write(int level, float* vals, uint64_t* idx, uint64_t* ptr, uint32_t rows, uint32_t cols, uint32_t nnz, const char* path) {
    if (level == 3) {
        IVCSC A(vals, idx, ptr, rows, cols, nnz);
        A.write(path);
    } else {
        VCSC A(vals, idx, ptr, rows, cols, nnz);
        A.write(path);
    }
    return 0;
}
// coordinates in the order the reference's iterator yields them (column by column)
extern "C" int64_t ref_ivsparse_read(int level, const char* path, uint64_t* r, uint64_t* c, float* v) {
    int64_t n = 0;
    if (level == 3) {
        IVCSC A(path);
        for (uint32_t col = 0; col < A.cols(); ++col)
            for (IVCSC::InnerIterator it(A, col); it; ++it) { r[n] = it.row(); c[n] = col; v[n] = it.value(); ++n; }
    } else {
        VCSC A(path);
        for (uint32_t col = 0; col < A.cols(); ++col)
            for (VCSC::InnerIterator it(A, col); it; ++it) { r[n] = it.row(); c[n] = col; v[n] = it.value(); ++n; }
    }
    return n;
}
