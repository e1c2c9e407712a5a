// singlet_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
//
// FP64 CPU restatement of the ALS-NMF hot path of zdebruine/singlet (reference v0.99.8,
// src/singlet.cpp). It is the checker for the CUDA path: only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it. The product
// (singlet_b200/, include/) never links, imports or falls back to anything in oracle/.
//
// Parity status: the reference's own test-suite pins no numeric result for this path
// (tests/testthat/test-pbmc3k.R:1-7 asserts TRUE only), and the reference shared library
// cannot be built here (needs R, Rcpp, RcppEigen -- none installed). Two anchors exist:
//   (1) `class rng` (src/singlet.cpp:7-114) is dependency-free; oracle/Makefile extracts it
//       from the reference where it lies into oracle/_ref/ and the hash functions below are
//       pinned bit-for-bit against it (tests/test_oracle_rng.py);
//   (2) oracle/Makefile also compiles the reference's own kernels (cor, AAt, scale, nnls,
//       predict, predict_mask, mse_test, c_nmf_base, c_ard_nmf_base, c_project_model),
//       extracted from src/singlet.cpp at build time, against a minimal Eigen/Rcpp shim
//       (oracle/shim/) into oracle/_ref/libsinglet_ref.so; goldens in tests/golden/ are
//       generated from that build and this restatement is pinned against them.
// Everything else is "parity unpinned by the reference's own tests".
//
// Layout conventions follow the reference: factors are column-major k x cols
// (element (f, c) at c*k + f); sparse inputs are dgCMatrix-style CSC (int32 p, int32 i,
// double x), given as a list of column chunks (n_chunks == 1 for a plain matrix).
//
// Build: see oracle/Makefile (-O2 -ffp-contract=off, no fast-math: IEEE FP64 like R's default
// CXXFLAGS).

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {

typedef struct {
    int64_t nrow, ncol;
    const int32_t* p;  // ncol + 1
    const int32_t* i;  // nnz, 0-based, ascending within a column
    const double* x;   // nnz
} orc_csc;

// ---------------------------------------------------------------------------------------------
// speckled-mask hash. Follows reference src/singlet.cpp:30-45 (rand(i)), :47-64 (rand(i,j)),
// :91-95 (draw). All arithmetic is uint64 with wrap-around.
// ---------------------------------------------------------------------------------------------
uint64_t orc_rand1(uint64_t state, uint64_t i) {
    i ^= i << 19;
    i ^= i >> 7;
    i ^= i << 36;
    uint64_t x = state + i;
    x ^= x << 38;
    x ^= x >> 13;
    x ^= x << 23;
    return x;
}

uint64_t orc_rand2(uint64_t state, uint64_t i, uint64_t j) {
    uint64_t x = orc_rand1(state, i);
    j ^= j >> 7;
    j ^= j << 23;
    j ^= j >> 8;
    x += j;
    x ^= x >> 7;
    x ^= x << 53;
    x ^= x >> 4;
    return x;
}

int orc_draw(uint64_t state, uint64_t i, uint64_t j, uint64_t inv_density) {
    return (orc_rand2(state, i, j) % inv_density) == 0 ? 1 : 0;
}

// mask of one cell over all genes: out[g] = draw(cell, g). Used to build golden index lists.
void orc_mask_cell(uint64_t state, uint64_t cell, uint64_t n_genes, uint64_t inv_density, uint8_t* out) {
    for (uint64_t g = 0; g < n_genes; ++g) out[g] = (uint8_t)orc_draw(state, cell, g, inv_density);
}

// ---------------------------------------------------------------------------------------------
// small dense helpers
// ---------------------------------------------------------------------------------------------

// 1 - Pearson correlation of two flattened matrices. Reference src/singlet.cpp:184-197.
double orc_cor(const double* x, const double* y, uint64_t n) {
    double sx = 0, sy = 0, sxy = 0, sx2 = 0, sy2 = 0;
    for (uint64_t t = 0; t < n; ++t) {
        const double a = x[t], b = y[t];
        sx += a;
        sy += b;
        sxy += a * b;
        sx2 += a * a;
        sy2 += b * b;
    }
    const double nn = (double)n;
    return 1 - (nn * sxy - sx * sy) / std::sqrt((nn * sx2 - sx * sx) * (nn * sy2 - sy * sy));
}

// a = X X^T + 1e-15 I for column-major k x cols X. Reference src/singlet.cpp:200-206.
// (Eigen's rankUpdate summation order cannot be reproduced; plain left-to-right sums here.)
void orc_gram(const double* X, int k, int64_t cols, double* a) {
    for (int r = 0; r < k; ++r)
        for (int c = 0; c <= r; ++c) {
            double s = 0;
            for (int64_t t = 0; t < cols; ++t) s += X[t * k + r] * X[t * k + c];
            a[(size_t)c * k + r] = s;
            a[(size_t)r * k + c] = s;
        }
    for (int r = 0; r < k; ++r) a[(size_t)r * k + r] += 1e-15;
}

// d = rowSums(X) + 1e-15 ; X[r,:] /= d[r]. Reference src/singlet.cpp:219-225.
void orc_scale(double* X, int k, int64_t cols, double* d) {
    for (int r = 0; r < k; ++r) {
        double s = 0;
        for (int64_t t = 0; t < cols; ++t) s += X[t * k + r];
        d[r] = s + 1e-15;
    }
    for (int r = 0; r < k; ++r)
        for (int64_t t = 0; t < cols; ++t) X[t * k + r] /= d[r];
}

// Sequential coordinate-descent NNLS on one column. Reference src/singlet.cpp:229-250 and the
// distilled rules in SURVEY.md App. A-5: uint8 sweep counter, `tol` reset per sweep, L1 taken
// off after the division by the diagonal, L2 added with x, clamp branch sets tol = 1.
// a: k x k column-major (read), b: k (destroyed), x: k (warm start in, solution out).
// Returns the number of sweeps executed.
int orc_nnls(const double* a, double* b, double* x, int k, double L1, double L2) {
    double tol = 1;
    int sweeps = 0;
    for (uint8_t it = 0; it < 100 && (tol / k) > 1e-8; ++it) {
        tol = 0;
        ++sweeps;
        for (int c = 0; c < k; ++c) {
            double diff = b[c] / a[(size_t)c * k + c];
            if (L1 != 0) diff -= L1;
            if (L2 != 0) diff += L2 * x[c];
            if (-diff > x[c]) {
                if (x[c] != 0) {
                    const double m = -x[c];
                    for (int r = 0; r < k; ++r) b[r] -= a[(size_t)c * k + r] * m;
                    tol = 1;
                    x[c] = 0;
                }
            } else if (diff != 0) {
                x[c] += diff;
                for (int r = 0; r < k; ++r) b[r] -= a[(size_t)c * k + r] * diff;
                tol += std::fabs(diff / (x[c] + 1e-15));
            }
        }
    }
    return sweeps;
}

static inline int resolve_threads(int threads) {
#ifdef _OPENMP
    return threads > 0 ? threads : omp_get_max_threads();
#else
    (void)threads;
    return 1;
#endif
}

int orc_max_threads(void) { return resolve_threads(0); }

// ---------------------------------------------------------------------------------------------
// predict: update h (k x n_total) given w (k x rows) over a chunk list.
// Reference src/singlet.cpp:333-347 (single) and :384-402 (list; running column offset).
// sweeps_out (optional, may be NULL): total coordinate-descent sweeps, for statistics.
// ---------------------------------------------------------------------------------------------
void orc_predict(const orc_csc* A, int n_chunks, const double* w, int k, double* h, double L1, double L2,
                 int threads, int64_t* sweeps_out) {
    const int64_t rows = A[0].nrow;
    std::vector<double> a((size_t)k * k);
    orc_gram(w, k, rows, a.data());
    int64_t offset = 0, sweeps = 0;
    const int nt = resolve_threads(threads);
    for (int ch = 0; ch < n_chunks; ++ch) {
        const orc_csc& M = A[ch];
#pragma omp parallel for num_threads(nt) reduction(+ : sweeps)
        for (int64_t c = 0; c < M.ncol; ++c) {
            if (M.p[c] == M.p[c + 1]) continue;
            std::vector<double> b((size_t)k, 0.0);
            for (int32_t t = M.p[c]; t < M.p[c + 1]; ++t) {
                const double v = M.x[t];
                const double* wr = w + (size_t)M.i[t] * k;
                for (int f = 0; f < k; ++f) b[f] += v * wr[f];
            }
            sweeps += orc_nnls(a.data(), b.data(), h + (size_t)(c + offset) * k, k, L1, L2);
        }
        offset += M.ncol;
    }
    if (sweeps_out) *sweeps_out = sweeps;
}

// predict_link: as predict, with b multiplied element-wise by column c of `link` (link_rows x n_total,
// column-major) before the solve. Reference src/singlet.cpp:416-433.
void orc_predict_link(const orc_csc* A, int n_chunks, const double* w, int k, double* h, double L1, double L2,
                      int threads, const double* link, int link_rows) {
    const int64_t rows = A[0].nrow;
    std::vector<double> a((size_t)k * k);
    orc_gram(w, k, rows, a.data());
    int64_t offset = 0;
    const int nt = resolve_threads(threads);
    for (int ch = 0; ch < n_chunks; ++ch) {
        const orc_csc& M = A[ch];
#pragma omp parallel for num_threads(nt)
        for (int64_t c = 0; c < M.ncol; ++c) {
            if (M.p[c] == M.p[c + 1]) continue;
            std::vector<double> b((size_t)k, 0.0);
            for (int32_t t = M.p[c]; t < M.p[c + 1]; ++t) {
                const double v = M.x[t];
                const double* wr = w + (size_t)M.i[t] * k;
                for (int f = 0; f < k; ++f) b[f] += v * wr[f];
            }
            for (int j = 0; j < link_rows; ++j) b[j] *= link[(size_t)(c + offset) * link_rows + j];
            orc_nnls(a.data(), b.data(), h + (size_t)(c + offset) * k, k, L1, L2);
        }
        offset += M.ncol;
    }
}

// ---------------------------------------------------------------------------------------------
// predict_mask: as predict, with the speckled test set held out.
// Reference src/singlet.cpp:436-466 (single) and :469-503 (list). mask_t == 0: columns are
// cells -> draw(col, row); mask_t != 0: columns are genes -> draw(row, col).
// Chunk offsets are added to the column index (src/singlet.cpp:485).
// ---------------------------------------------------------------------------------------------
void orc_predict_mask(const orc_csc* A, int n_chunks, uint64_t seed, uint64_t inv_density, const double* w, int k,
                      double* h, double L1, double L2, int threads, int mask_t) {
    const int64_t rows = A[0].nrow;
    std::vector<double> a((size_t)k * k);
    orc_gram(w, k, rows, a.data());
    int64_t offset = 0;
    const int nt = resolve_threads(threads);
    for (int ch = 0; ch < n_chunks; ++ch) {
        const orc_csc& M = A[ch];
#pragma omp parallel for num_threads(nt) schedule(dynamic, 16)
        for (int64_t c = 0; c < M.ncol; ++c) {
            if (M.p[c] == M.p[c + 1]) continue;
            std::vector<double> b((size_t)k, 0.0);
            std::vector<uint64_t> held;
            held.reserve((size_t)(rows / (int64_t)inv_density) + 8);
            int32_t t = M.p[c];
            const int32_t t_end = M.p[c + 1];
            const uint64_t gc = (uint64_t)(c + offset);
            for (uint64_t r = 0; r < (uint64_t)rows; ++r) {
                const bool masked = mask_t ? orc_draw(seed, r, gc, inv_density) : orc_draw(seed, gc, r, inv_density);
                const bool at_nz = (t < t_end) && ((uint64_t)M.i[t] == r);
                if (masked) {
                    held.push_back(r);
                    if (at_nz) ++t;
                } else if (at_nz) {
                    const double v = M.x[t];
                    const double* wr = w + (size_t)r * k;
                    for (int f = 0; f < k; ++f) b[f] += v * wr[f];
                    ++t;
                }
            }
            // a_i = AAt(w) - AAt(w[:, held]); both carry the 1e-15 diagonal jitter (App. A-11)
            std::vector<double> asub((size_t)k * k, 0.0);
            for (int r_ = 0; r_ < k; ++r_)
                for (int c_ = 0; c_ <= r_; ++c_) {
                    double s = 0;
                    for (size_t q = 0; q < held.size(); ++q)
                        s += w[(size_t)held[q] * k + r_] * w[(size_t)held[q] * k + c_];
                    asub[(size_t)c_ * k + r_] = s;
                    asub[(size_t)r_ * k + c_] = s;
                }
            for (int r_ = 0; r_ < k; ++r_) asub[(size_t)r_ * k + r_] += 1e-15;
            std::vector<double> ai((size_t)k * k);
            for (size_t q = 0; q < ai.size(); ++q) ai[q] = a[q] - asub[q];
            orc_nnls(ai.data(), b.data(), h + (size_t)(c + offset) * k, k, L1, L2);
        }
        offset += M.ncol;
    }
}

// ---------------------------------------------------------------------------------------------
// mse_test: mean over cell-columns of the mean squared residual on held-out entries (including
// structural zeros). Reference src/singlet.cpp:536-568 (single) and :571-607 (list).
// A: genes x cells chunk list; w: k x m; d: k; h: k x n.
// ---------------------------------------------------------------------------------------------
double orc_mse_test(const orc_csc* A, int n_chunks, const double* w, const double* d, const double* h, int k,
                    uint64_t seed, uint64_t inv_density, int threads) {
    const int64_t m = A[0].nrow;
    int64_t n = 0;
    for (int ch = 0; ch < n_chunks; ++ch) n += A[ch].ncol;
    // w_ = t(w) with factor f scaled by d[f]  (m x k, stored row-major here: wd[g*k + f])
    std::vector<double> wd((size_t)m * k);
    for (int64_t g = 0; g < m; ++g)
        for (int f = 0; f < k; ++f) wd[(size_t)g * k + f] = w[(size_t)g * k + f] * d[f];
    std::vector<double> losses((size_t)n, 0.0);
    int64_t offset = 0;
    const int nt = resolve_threads(threads);
    for (int ch = 0; ch < n_chunks; ++ch) {
        const orc_csc& M = A[ch];
#pragma omp parallel for num_threads(nt) schedule(dynamic, 16)
        for (int64_t c = 0; c < M.ncol; ++c) {
            uint64_t cnt = 0;
            double s = 0;
            int32_t t = M.p[c];
            const int32_t t_end = M.p[c + 1];
            const uint64_t gc = (uint64_t)(c + offset);
            const double* hc = h + (size_t)gc * k;
            for (uint64_t g = 0; g < (uint64_t)m; ++g) {
                const bool at_nz = (t < t_end) && ((uint64_t)M.i[t] == g);
                if (orc_draw(seed, gc, g, inv_density)) {
                    ++cnt;
                    double pred = 0;
                    for (int f = 0; f < k; ++f) pred += wd[(size_t)g * k + f] * hc[f];
                    const double res = at_nz ? (pred - M.x[t]) : pred;
                    s += res * res;
                    if (at_nz) ++t;
                } else if (at_nz) {
                    ++t;
                }
            }
            losses[(size_t)gc] = cnt > 0 ? s / (double)cnt : 0.0;
        }
        offset += M.ncol;
    }
    double tot = 0;
    for (int64_t c = 0; c < n; ++c) tot += losses[(size_t)c];
    return tot / (double)n;
}

// Harness-defined train MSE (not in the reference): the same per-column-mean then
// mean-over-columns statistic over the entries that are NOT held out. inv_density == 0 means
// "no mask" (plain NMF: all m entries of every column).
double orc_mse_train(const orc_csc* A, int n_chunks, const double* w, const double* d, const double* h, int k,
                     uint64_t seed, uint64_t inv_density, int threads) {
    const int64_t m = A[0].nrow;
    int64_t n = 0;
    for (int ch = 0; ch < n_chunks; ++ch) n += A[ch].ncol;
    std::vector<double> wd((size_t)m * k);
    for (int64_t g = 0; g < m; ++g)
        for (int f = 0; f < k; ++f) wd[(size_t)g * k + f] = w[(size_t)g * k + f] * d[f];
    std::vector<double> losses((size_t)n, 0.0);
    int64_t offset = 0;
    const int nt = resolve_threads(threads);
    for (int ch = 0; ch < n_chunks; ++ch) {
        const orc_csc& M = A[ch];
#pragma omp parallel for num_threads(nt) schedule(dynamic, 16)
        for (int64_t c = 0; c < M.ncol; ++c) {
            uint64_t cnt = 0;
            double s = 0;
            int32_t t = M.p[c];
            const int32_t t_end = M.p[c + 1];
            const uint64_t gc = (uint64_t)(c + offset);
            const double* hc = h + (size_t)gc * k;
            for (uint64_t g = 0; g < (uint64_t)m; ++g) {
                const bool at_nz = (t < t_end) && ((uint64_t)M.i[t] == g);
                const bool masked = inv_density ? orc_draw(seed, gc, g, inv_density) : false;
                if (!masked) {
                    ++cnt;
                    double pred = 0;
                    for (int f = 0; f < k; ++f) pred += wd[(size_t)g * k + f] * hc[f];
                    const double res = at_nz ? (pred - M.x[t]) : pred;
                    s += res * res;
                }
                if (at_nz) ++t;
            }
            losses[(size_t)gc] = cnt > 0 ? s / (double)cnt : 0.0;
        }
        offset += M.ncol;
    }
    double tot = 0;
    for (int64_t c = 0; c < n; ++c) tot += losses[(size_t)c];
    return tot / (double)n;
}

// ---------------------------------------------------------------------------------------------
// drivers
// ---------------------------------------------------------------------------------------------

// Plain ALS NMF. Reference src/singlet.cpp:638-666 (c_nmf_base) and :715-743 (list variant).
// A: genes x cells chunks; At: cells x genes chunks (gene blocks). w: k x m in/out; d: k out;
// h: k x n out. tol_trace (optional, length >= maxit): per-iteration 1 - cor.
// Returns the number of iterations executed.
int orc_nmf(const orc_csc* A, int nA, const orc_csc* At, int nAt, double tol, uint16_t maxit, double L1_w,
            double L1_h, double L2_w, double L2_h, int threads, int k, double* w, double* d, double* h,
            double* tol_trace) {
    const int64_t m = A[0].nrow;
    int64_t n = 0;
    for (int ch = 0; ch < nA; ++ch) n += A[ch].ncol;
    std::memset(h, 0, sizeof(double) * (size_t)k * (size_t)n);
    for (int f = 0; f < k; ++f) d[f] = 1.0;
    double tol_ = 1;
    std::vector<double> w_it((size_t)k * (size_t)m);
    uint16_t iter_ = 0;
    for (; iter_ < maxit && tol_ > tol; ++iter_) {
        std::memcpy(w_it.data(), w, sizeof(double) * w_it.size());
        orc_predict(A, nA, w, k, h, L1_h, L2_h, threads, nullptr);
        orc_scale(h, k, n, d);
        orc_predict(At, nAt, h, k, w, L1_w, L2_w, threads, nullptr);
        orc_scale(w, k, m, d);
        tol_ = orc_cor(w, w_it.data(), (uint64_t)k * (uint64_t)m);
        if (tol_trace) tol_trace[iter_] = tol_;
    }
    return (int)iter_;
}

// Linked NMF. Reference src/singlet.cpp:1059-1086: linking is applied to the H update when link_h has one
// column per cell, to the W update when link_w has one column per gene; otherwise that side is plain.
int orc_linked_nmf(const orc_csc* A, const orc_csc* At, double tol, uint16_t maxit, double L1, double L2, int threads, int k,
                   double* w, double* d, double* h, const double* link_h, int lh_rows, int64_t lh_cols, const double* link_w,
                   int lw_rows, int64_t lw_cols) {
    const int64_t m = A[0].nrow, n = A[0].ncol;
    std::memset(h, 0, sizeof(double) * (size_t)k * (size_t)n);
    for (int f = 0; f < k; ++f) d[f] = 1.0;
    const bool linking_h = (lh_cols == n), linking_w = (lw_cols == m);
    double tol_ = 1;
    std::vector<double> w_it((size_t)k * (size_t)m);
    uint16_t iter_ = 0;
    for (; iter_ < maxit && tol_ > tol; ++iter_) {
        std::memcpy(w_it.data(), w, sizeof(double) * w_it.size());
        if (linking_h) orc_predict_link(A, 1, w, k, h, L1, L2, threads, link_h, lh_rows);
        else orc_predict(A, 1, w, k, h, L1, L2, threads, nullptr);
        orc_scale(h, k, n, d);
        if (linking_w) orc_predict_link(At, 1, h, k, w, L1, L2, threads, link_w, lw_rows);
        else orc_predict(At, 1, h, k, w, L1, L2, threads, nullptr);
        orc_scale(w, k, m, d);
        tol_ = orc_cor(w, w_it.data(), (uint64_t)k * (uint64_t)m);
    }
    return (int)iter_;
}

// weight_by_split. Reference src/singlet.cpp:119-144: every group's total is made equal to group 0's by
// dividing the values of the columns of group g != 0 by sums[g] / sums[0]. x is modified in place.
void orc_weight_by_split(const orc_csc* A, double* x, const int32_t* split_by, int n_groups) {
    std::vector<double> sums((size_t)n_groups, 0.0);
    for (int64_t j = 0; j < A->ncol; ++j)
        for (int32_t t = A->p[j]; t < A->p[j + 1]; ++t) sums[(size_t)split_by[j]] += x[t];
    for (int g = 1; g < n_groups; ++g) sums[(size_t)g] /= sums[0];
    for (int64_t j = 0; j < A->ncol; ++j)
        if (split_by[j] != 0)
            for (int32_t t = A->p[j]; t < A->p[j + 1]; ++t) x[t] /= sums[(size_t)split_by[j]];
}

// Cross-validated ("ard") ALS NMF with speckled mask. Reference src/singlet.cpp:1091-1152 and
// :1162-1234 (list variant); trace rules in SURVEY.md App. A-13.
// Trace outputs (capacity trace_cap each): test_mse, iter, fit_tol, score_overfit; *n_trace = count.
// Returns the final value of iter_.
int orc_ard_nmf(const orc_csc* A, int nA, const orc_csc* At, int nAt, double tol, uint16_t maxit, double L1,
                double L2, int threads, int k, double* w, double* d, double* h, uint64_t seed, uint64_t inv_density,
                double overfit_threshold, uint16_t trace_test_mse, double* test_mse, int32_t* iter_out,
                double* fit_tol, double* score_overfit, int trace_cap, int* n_trace) {
    const int64_t m = A[0].nrow;
    int64_t n = 0;
    for (int ch = 0; ch < nA; ++ch) n += A[ch].ncol;
    std::memset(h, 0, sizeof(double) * (size_t)k * (size_t)n);
    for (int f = 0; f < k; ++f) d[f] = 1.0;
    double tol_ = 1;
    std::vector<double> w_it((size_t)k * (size_t)m);
    int nt = 0;
    auto push = [&](double mse, int it, double ft) {
        if (nt >= trace_cap) return;
        test_mse[nt] = mse;
        iter_out[nt] = it;
        fit_tol[nt] = ft;
        double mn = test_mse[0];
        for (int q = 1; q <= nt; ++q) mn = test_mse[q] < mn ? test_mse[q] : mn;
        score_overfit[nt] = (mse - mn) / (mse + mn);
        ++nt;
    };
    uint16_t iter_ = 0;
    for (; iter_ < maxit && tol_ > tol; ++iter_) {
        std::memcpy(w_it.data(), w, sizeof(double) * w_it.size());
        orc_predict_mask(A, nA, seed, inv_density, w, k, h, L1, L2, threads, 0);
        orc_scale(h, k, n, d);
        orc_predict_mask(At, nAt, seed, inv_density, h, k, w, L1, L2, threads, 1);
        orc_scale(w, k, m, d);
        tol_ = orc_cor(w, w_it.data(), (uint64_t)k * (uint64_t)m);
        if (iter_ % trace_test_mse == 0) {
            push(orc_mse_test(A, nA, w, d, h, k, seed, inv_density, threads), (int)iter_, tol_);
            if (score_overfit[nt - 1] > overfit_threshold) break;
        }
    }
    if (iter_ % trace_test_mse != 0) push(orc_mse_test(A, nA, w, d, h, k, seed, inv_density, threads), (int)iter_, tol_);
    *n_trace = nt;
    return (int)iter_;
}

// project_model. Reference src/singlet.cpp:405-413. w must already be k x m (the caller performs
// the "transpose if m x k" step of :406); w is scaled in place like the reference's by-value copy.
void orc_project_model(const orc_csc* A, int nA, double* w, int k, double L1, double L2, int threads, double* h,
                       double* d) {
    const int64_t m = A[0].nrow;
    int64_t n = 0;
    for (int ch = 0; ch < nA; ++ch) n += A[ch].ncol;
    for (int f = 0; f < k; ++f) d[f] = 1.0;
    orc_scale(w, k, m, d);
    std::memset(h, 0, sizeof(double) * (size_t)k * (size_t)n);
    orc_predict(A, nA, w, k, h, L1, L2, threads, nullptr);
    orc_scale(h, k, n, d);
}

}  // extern "C"
