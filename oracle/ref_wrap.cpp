// ref_wrap.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/singlet_oracle.cpp header).
//
// extern "C" face of oracle/_ref/libsinglet_ref.so: the reference's OWN hot-path function bodies
// (extracted by oracle/Makefile from /root/reference/src/singlet.cpp into
// oracle/_ref/singlet_extract.inc at build time; never committed) compiled against the minimal
// Eigen/Rcpp shim in oracle/shim/. Signatures mirror the orc_* functions of singlet_oracle.cpp so
// the tests can run both on the same buffers.
#include "shim/eigen_rcpp_shim.hpp"

#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "_ref/singlet_extract.inc"

extern "C" {

typedef struct {
    int64_t nrow, ncol;
    const int32_t* p;
    const int32_t* i;
    const double* x;
} orc_csc;

static Rcpp::SparseMatrix view(const orc_csc& c) {
    return Rcpp::SparseMatrix(c.x, c.i, c.p, (int)c.nrow, (int)c.ncol);
}
static std::vector<Rcpp::SparseMatrix> views(const orc_csc* c, int n) {
    std::vector<Rcpp::SparseMatrix> v;
    for (int q = 0; q < n; ++q) v.push_back(view(c[q]));
    return v;
}
static Eigen::MatrixXd mat(const double* src, long r, long c) {
    Eigen::MatrixXd m(r, c);
    if (src) std::memcpy(m.data(), src, sizeof(double) * (size_t)r * (size_t)c);
    return m;
}
static void put(const Rcpp::List& l, const char* name, double* dst) {
    const Rcpp::Entry* e = l.find(name);
    if (e && dst) std::memcpy(dst, e->data.data(), sizeof(double) * e->data.size());
}
static int threads_or_all(int threads) {
#ifdef _OPENMP
    return threads > 0 ? threads : omp_get_max_threads();
#else
    return 1;
#endif
}

uint64_t ref_rand1(uint64_t state, uint64_t i) { return rng(state).rand(i); }
uint64_t ref_rand2(uint64_t state, uint64_t i, uint64_t j) { return rng(state).rand(i, j); }
int ref_draw(uint64_t state, uint64_t i, uint64_t j, uint64_t inv_density) {
    return rng(state).draw(i, j, inv_density) ? 1 : 0;
}

double ref_cor(const double* x, const double* y, uint64_t n) {
    Eigen::MatrixXd a = mat(x, (long)n, 1), b = mat(y, (long)n, 1);
    return cor(a, b);
}
void ref_gram(const double* X, int k, int64_t cols, double* a) {
    Eigen::MatrixXd g = AAt(mat(X, k, (long)cols));
    std::memcpy(a, g.data(), sizeof(double) * (size_t)k * k);
}
void ref_scale(double* X, int k, int64_t cols, double* d) {
    Eigen::MatrixXd m = mat(X, k, (long)cols);
    Eigen::VectorXd dv = Eigen::VectorXd::Ones(k);
    scale(m, dv);
    std::memcpy(X, m.data(), sizeof(double) * (size_t)k * (size_t)cols);
    std::memcpy(d, dv.data(), sizeof(double) * (size_t)k);
}
int ref_nnls(const double* a, double* b, double* x, int k, double L1, double L2) {
    Eigen::MatrixXd am = mat(a, k, k), xm = mat(x, k, 1);
    Eigen::VectorXd bv(k);
    std::memcpy(bv.data(), b, sizeof(double) * (size_t)k);
    nnls(am, bv, xm, 0, L1, L2);
    std::memcpy(b, bv.data(), sizeof(double) * (size_t)k);
    std::memcpy(x, xm.data(), sizeof(double) * (size_t)k);
    return -1;  // the reference does not report its sweep count
}

void ref_predict(const orc_csc* A, int n_chunks, const double* w, int k, double* h, double L1, double L2,
                 int threads, int64_t*) {
    int64_t n = 0;
    for (int q = 0; q < n_chunks; ++q) n += A[q].ncol;
    Eigen::MatrixXd wm = mat(w, k, (long)A[0].nrow), hm = mat(h, k, (long)n);
    if (n_chunks == 1)
        predict(view(A[0]), wm, hm, L1, L2, threads_or_all(threads));
    else
        predict(views(A, n_chunks), wm, hm, L1, L2, threads_or_all(threads));
    std::memcpy(h, hm.data(), sizeof(double) * (size_t)k * (size_t)n);
}

void ref_predict_mask(const orc_csc* A, int n_chunks, uint64_t seed, uint64_t inv_density, const double* w, int k,
                      double* h, double L1, double L2, int threads, int mask_t) {
    int64_t n = 0;
    for (int q = 0; q < n_chunks; ++q) n += A[q].ncol;
    Eigen::MatrixXd wm = mat(w, k, (long)A[0].nrow), hm = mat(h, k, (long)n);
    if (n_chunks == 1) {
        predict_mask(view(A[0]), rng(seed), inv_density, wm, hm, L1, L2, threads_or_all(threads), mask_t != 0);
    } else {
        std::vector<Rcpp::SparseMatrix> v = views(A, n_chunks);
        predict_mask(v, rng(seed), inv_density, wm, hm, L1, L2, threads_or_all(threads), mask_t != 0);
    }
    std::memcpy(h, hm.data(), sizeof(double) * (size_t)k * (size_t)n);
}

double ref_mse_test(const orc_csc* A, int n_chunks, const double* w, const double* d, const double* h, int k,
                    uint64_t seed, uint64_t inv_density, int threads) {
    int64_t n = 0;
    for (int q = 0; q < n_chunks; ++q) n += A[q].ncol;
    Eigen::MatrixXd wm = mat(w, k, (long)A[0].nrow), hm = mat(h, k, (long)n);
    Eigen::VectorXd dv(k);
    std::memcpy(dv.data(), d, sizeof(double) * (size_t)k);
    if (n_chunks == 1) return mse_test(view(A[0]), wm, dv, hm, rng(seed), inv_density, (uint16_t)threads_or_all(threads));
    return mse_test(views(A, n_chunks), wm, dv, hm, rng(seed), inv_density, (uint16_t)threads_or_all(threads));
}

int ref_nmf(const orc_csc* A, int nA, const orc_csc* At, int nAt, double tol, uint16_t maxit, double L1_w,
            double L1_h, double L2_w, double L2_h, int threads, int k, double* w, double* d, double* h, double*) {
    Eigen::MatrixXd wm = mat(w, k, (long)A[0].nrow);
    Rcpp::List out;
    if (nA == 1 && nAt == 1) {
        Rcpp::SparseMatrix a = view(A[0]), at = view(At[0]);
        out = c_nmf(a, at, tol, maxit, false, L1_w, L1_h, L2_w, L2_h, (uint16_t)threads_or_all(threads), wm);
    } else {
        // the list entry point has a single L1/L2 (src/singlet.cpp:715)
        Rcpp::List la, lat;
        la.mats = views(A, nA);
        lat.mats = views(At, nAt);
        out = c_nmf_sparse_list(la, lat, tol, maxit, false, L1_w, L2_w, (uint16_t)threads_or_all(threads), wm);
    }
    put(out, "w", w);
    put(out, "d", d);
    put(out, "h", h);
    return -1;  // iteration count is not returned by the reference
}

int ref_ard_nmf(const orc_csc* A, int nA, const orc_csc* At, int nAt, double tol, uint16_t maxit, double L1,
                double L2, int threads, int k, double* w, double* d, double* h, uint64_t seed, uint64_t inv_density,
                double overfit_threshold, uint16_t trace_test_mse, double* test_mse, int32_t* iter_out,
                double* fit_tol, double* score_overfit, int trace_cap, int* n_trace) {
    Eigen::MatrixXd wm = mat(w, k, (long)A[0].nrow);
    Rcpp::List out;
    if (nA == 1 && nAt == 1) {
        Rcpp::SparseMatrix a = view(A[0]), at = view(At[0]);
        out = c_ard_nmf(a, at, tol, maxit, false, L1, L2, (uint16_t)threads_or_all(threads), wm, seed, inv_density,
                        overfit_threshold, trace_test_mse);
    } else {
        Rcpp::List la, lat;
        la.mats = views(A, nA);
        lat.mats = views(At, nAt);
        out = c_ard_nmf_sparse_list(la, lat, tol, maxit, false, L1, L2, (uint16_t)threads_or_all(threads), wm, seed,
                                    inv_density, overfit_threshold, trace_test_mse);
    }
    put(out, "w", w);
    put(out, "d", d);
    put(out, "h", h);
    const Rcpp::Entry* e = out.find("test_mse");
    int nt = e ? (int)e->data.size() : 0;
    if (nt > trace_cap) nt = trace_cap;
    const Rcpp::Entry* ei = out.find("iter");
    const Rcpp::Entry* et = out.find("tol");
    const Rcpp::Entry* es = out.find("score_overfit");
    for (int q = 0; q < nt; ++q) {
        test_mse[q] = e->data[(size_t)q];
        iter_out[q] = (int32_t)ei->data[(size_t)q];
        fit_tol[q] = et->data[(size_t)q];
        score_overfit[q] = es->data[(size_t)q];
    }
    *n_trace = nt;
    return nt ? iter_out[nt - 1] : 0;
}

void ref_project_model(const orc_csc* A, int, double* w, int k, double L1, double L2, int threads, double* h,
                       double* d) {
    Eigen::MatrixXd wm = mat(w, k, (long)A[0].nrow);
    Rcpp::List out = c_project_model(view(A[0]), wm, L1, L2, threads_or_all(threads));
    put(out, "h", h);
    put(out, "d", d);
}

// dense-input variants: A is m x n, At is n x m, both column-major
int ref_nmf_dense(const double* A, const double* At, int64_t m, int64_t n, double tol, uint16_t maxit, double L1_w, double L1_h,
                  double L2_w, double L2_h, int threads, int k, double* w, double* d, double* h) {
    Eigen::MatrixXd a = mat(A, (long)m, (long)n), at = mat(At, (long)n, (long)m), wm = mat(w, k, (long)m);
    Rcpp::List out = c_nmf_dense(a, at, tol, maxit, false, L1_w, L1_h, L2_w, L2_h, (uint16_t)threads_or_all(threads), wm);
    put(out, "w", w);
    put(out, "d", d);
    put(out, "h", h);
    return -1;
}
int ref_ard_nmf_dense(const double* A, const double* At, int64_t m, int64_t n, double tol, uint16_t maxit, double L1, double L2,
                      int threads, int k, double* w, double* d, double* h, uint64_t seed, uint64_t inv_density,
                      double overfit_threshold, uint16_t trace_test_mse, double* test_mse, int32_t* iter_out, int trace_cap,
                      int* n_trace) {
    Eigen::MatrixXd a = mat(A, (long)m, (long)n), at = mat(At, (long)n, (long)m), wm = mat(w, k, (long)m);
    Rcpp::List out = c_ard_nmf_dense(a, at, tol, maxit, false, L1, L2, (uint16_t)threads_or_all(threads), wm, seed, inv_density,
                                     overfit_threshold, trace_test_mse);
    put(out, "w", w);
    put(out, "d", d);
    put(out, "h", h);
    const Rcpp::Entry* e = out.find("test_mse");
    const Rcpp::Entry* ei = out.find("iter");
    int nt = e ? (int)e->data.size() : 0;
    if (nt > trace_cap) nt = trace_cap;
    for (int q = 0; q < nt; ++q) {
        test_mse[q] = e->data[(size_t)q];
        iter_out[q] = (int32_t)ei->data[(size_t)q];
    }
    *n_trace = nt;
    return nt;
}

int ref_linked_nmf(const orc_csc* A, const orc_csc* At, double tol, uint16_t maxit, double L1, double L2, int threads, int k,
                   double* w, double* d, double* h, const double* link_h, int lh_rows, int64_t lh_cols, const double* link_w,
                   int lw_rows, int64_t lw_cols) {
    Eigen::MatrixXd wm = mat(w, k, (long)A[0].nrow);
    Eigen::MatrixXd lh = mat(link_h, lh_rows, (long)lh_cols), lw = mat(link_w, lw_rows, (long)lw_cols);
    Rcpp::List out = c_linked_nmf(view(A[0]), view(At[0]), tol, maxit, false, L1, L2, (uint16_t)threads_or_all(threads), wm, lh, lw);
    put(out, "w", w);
    put(out, "d", d);
    put(out, "h", h);
    return -1;
}

int ref_max_threads(void) { return threads_or_all(0); }

}  // extern "C"
