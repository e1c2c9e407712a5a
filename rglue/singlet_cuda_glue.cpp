// singlet_cuda_glue.cpp -- drop-in bodies for the reference's Rcpp entry points on the ALS-NMF path.
//
// This file REPLACES the bodies of c_nmf, c_nmf_sparse_list, c_ard_nmf, c_ard_nmf_sparse_list,
// c_project_model and Rcpp_predict in the reference's src/singlet.cpp (lines 350-367, 405-413, 669-672,
// 715-743, 1155-1234). Their signatures are unchanged, so the auto-generated shims in
// src/RcppExports.cpp (:67-116, :138-155, :283-327) and the R wrappers in R/RcppExports.R keep working
// byte for byte. Everything below only unpacks the R objects, calls the C ABI of libsinglet_cuda.so
// (include/singlet_cuda.h) and wraps the results. R is not installed in the build container, so this
// file is compiled where the package is built; the same calls are exercised
// from Python (singlet_b200/api.py) in the test-suite. Further down: the multi-GPU bodies of the two *_sparse_list entry
// points (sgl_multi_*: one R process, all GPUs of the box) and the IVSparse functions of src/singlet.cpp:783-995
// (save_IVSparse, read_IVSparse, run_nmf_on_sparsematrix_list) on top of sgl_ivsparse_*.
#include <Rcpp.h>
#include <RcppEigen.h>
#include <singlet.h>  // Rcpp::SparseMatrix (inst/include/singlet.h:36-102)

#include <vector>

#include "singlet_cuda.h"

namespace {

sgl_handle* handle() {  // one handle per R session: keeps A/At and the CV mask resident between calls
    static sgl_handle* h = nullptr;
    if (!h && sgl_create(0, nullptr, &h) != SGL_OK) Rcpp::stop(sgl_last_error());
    return h;
}

sgl_csc view(Rcpp::SparseMatrix& A) {  // zero-copy: points at the dgCMatrix slots
    sgl_csc c;
    c.nrow = A.rows();
    c.ncol = A.cols();
    c.p = A.p.begin();
    c.i = A.i.begin();
    c.x = A.x.begin();
    return c;
}
std::vector<sgl_csc> views(std::vector<Rcpp::SparseMatrix>& v) {
    std::vector<sgl_csc> out;
    for (auto& m : v) out.push_back(view(m));
    return out;
}
std::vector<Rcpp::SparseMatrix> as_list(Rcpp::List& L) {
    std::vector<Rcpp::SparseMatrix> v;
    for (auto&& e : L) v.push_back(Rcpp::as<Rcpp::SparseMatrix>(e));
    return v;
}

// Rprintf / checkUserInterrupt of src/singlet.cpp:643-663, 1102-1128, driven from the library's callbacks
struct Progress {
    bool verbose, masked;
};
void on_iter(void* u, int iter, double tol, double overfit) {
    Progress* p = static_cast<Progress*>(u);
    if (!p->verbose) return;
    if (!p->masked) Rprintf("%4d | %8.2e\n", iter, tol);
    else if (ISNAN(overfit)) Rprintf("%4d | %8.2e | %8s\n", iter, tol, "-");
    else Rprintf("%4d | %8.2e | %8.2e\n", iter, tol, overfit);
}
void check_interrupt_fn(void*) { R_CheckUserInterrupt(); }
int poll_interrupt(void*) {  // R_ToplevelExec returns FALSE when the user interrupted
    return R_ToplevelExec(check_interrupt_fn, nullptr) == FALSE ? 1 : 0;
}
sgl_callbacks callbacks(Progress& p) {
    sgl_callbacks cb;
    cb.user = &p;
    cb.poll_interrupt = poll_interrupt;
    cb.on_iter = on_iter;
    return cb;
}
void check(int rc) {
    if (rc == SGL_EINTERRUPT) Rcpp::stop("interrupted");  // the library has already unwound and freed
    if (rc != SGL_OK) Rcpp::stop(sgl_last_error());
}

Rcpp::List nmf_impl(std::vector<sgl_csc> A, std::vector<sgl_csc> At, double tol, uint16_t maxit, bool verbose, double L1_w,
                    double L1_h, double L2_w, double L2_h, Eigen::MatrixXd& w) {
    const int k = (int)w.rows();
    int64_t n = 0;
    for (auto& c : A) n += c.ncol;
    Eigen::MatrixXd h(k, n);
    Eigen::VectorXd d(k);
    Progress p{verbose, false};
    sgl_callbacks cb = callbacks(p);
    if (verbose) Rprintf("\n%4s | %8s \n---------------\n", "iter", "tol");
    check(sgl_nmf(handle(), A.data(), (int)A.size(), At.data(), (int)At.size(), tol, maxit, L1_w, L1_h, L2_w, L2_h, k, w.data(),
                  d.data(), h.data(), nullptr, nullptr, &cb));
    return Rcpp::List::create(Rcpp::Named("w") = w, Rcpp::Named("d") = d, Rcpp::Named("h") = h);
}

Rcpp::List ard_impl(std::vector<sgl_csc> A, std::vector<sgl_csc> At, double tol, uint16_t maxit, bool verbose, double L1, double L2,
                    Eigen::MatrixXd& w, uint64_t seed, uint64_t inv_density, double overfit_threshold, uint16_t trace_test_mse) {
    const int k = (int)w.rows();
    int64_t n = 0;
    for (auto& c : A) n += c.ncol;
    Eigen::MatrixXd h(k, n);
    Eigen::VectorXd d(k);
    const int cap = (int)maxit + 2;
    std::vector<double> mse(cap), ft(cap), so(cap);
    std::vector<int32_t> it(cap);
    sgl_trace tr{mse.data(), it.data(), ft.data(), so.data(), cap, 0};
    Progress p{verbose, true};
    sgl_callbacks cb = callbacks(p);
    if (verbose) Rprintf("\n%4s | %8s | %8s \n---------------------------\n", "iter", "tol", "overfit");
    check(sgl_ard_nmf(handle(), A.data(), (int)A.size(), At.data(), (int)At.size(), tol, maxit, L1, L2, k, w.data(), d.data(),
                      h.data(), seed, inv_density, overfit_threshold, trace_test_mse, &tr, &cb));
    Rcpp::NumericVector test_mse(mse.begin(), mse.begin() + tr.length), fit_tol(ft.begin(), ft.begin() + tr.length),
        score(so.begin(), so.begin() + tr.length);
    Rcpp::IntegerVector iter(it.begin(), it.begin() + tr.length);
    return Rcpp::List::create(Rcpp::Named("w") = w, Rcpp::Named("d") = d, Rcpp::Named("h") = h, Rcpp::Named("test_mse") = test_mse,
                              Rcpp::Named("iter") = iter, Rcpp::Named("tol") = fit_tol, Rcpp::Named("score_overfit") = score);
}

}  // namespace

//[[Rcpp::export]]
Rcpp::List c_nmf(Rcpp::SparseMatrix& A, Rcpp::SparseMatrix& At, const double tol, const uint16_t maxit, const bool verbose,
                 const double L1_w, const double L1_h, const double L2_w, const double L2_h, const uint16_t threads,
                 Eigen::MatrixXd w) {
    return nmf_impl({view(A)}, {view(At)}, tol, maxit, verbose, L1_w, L1_h, L2_w, L2_h, w);
}

//[[Rcpp::export]]
Rcpp::List c_nmf_sparse_list(Rcpp::List A_, Rcpp::List& At_, const double tol, const uint16_t maxit, const bool verbose,
                             const double L1, const double L2, const uint16_t threads, Eigen::MatrixXd w) {
    std::vector<Rcpp::SparseMatrix> A = as_list(A_), At = as_list(At_);
    return nmf_impl(views(A), views(At), tol, maxit, verbose, L1, L1, L2, L2, w);
}

// dense-input variants (reference src/singlet.cpp:1051-1054, 1357-1361): R matrices are column-major doubles
//[[Rcpp::export]]
Rcpp::List c_nmf_dense(Eigen::MatrixXd& A, Eigen::MatrixXd& At, const double tol, const uint16_t maxit, const bool verbose,
                       const double L1_w, const double L1_h, const double L2_w, const double L2_h, const uint16_t threads,
                       Eigen::MatrixXd w) {
    const int k = (int)w.rows();
    Eigen::MatrixXd h(k, A.cols());
    Eigen::VectorXd d(k);
    Progress p{verbose, false};
    sgl_callbacks cb = callbacks(p);
    if (verbose) Rprintf("\n%4s | %8s \n---------------\n", "iter", "tol");
    check(sgl_nmf_dense(handle(), A.data(), At.data(), A.rows(), A.cols(), tol, maxit, L1_w, L1_h, L2_w, L2_h, k, w.data(), d.data(),
                        h.data(), nullptr, nullptr, &cb));
    return Rcpp::List::create(Rcpp::Named("w") = w, Rcpp::Named("d") = d, Rcpp::Named("h") = h);
}

//[[Rcpp::export]]
Rcpp::List c_ard_nmf_dense(Eigen::MatrixXd& A, Eigen::MatrixXd& At, const double tol, const uint16_t maxit, const bool verbose,
                           const double L1, const double L2, const uint16_t threads, Eigen::MatrixXd w, const uint64_t seed,
                           const uint64_t inv_density, const double overfit_threshold, const uint16_t trace_test_mse) {
    const int k = (int)w.rows();
    Eigen::MatrixXd h(k, A.cols());
    Eigen::VectorXd d(k);
    const int cap = (int)maxit + 2;
    std::vector<double> mse(cap), ft(cap), so(cap);
    std::vector<int32_t> it(cap);
    sgl_trace tr{mse.data(), it.data(), ft.data(), so.data(), cap, 0};
    Progress p{verbose, true};
    sgl_callbacks cb = callbacks(p);
    if (verbose) Rprintf("\n%4s | %8s | %8s \n---------------------------\n", "iter", "tol", "overfit");
    check(sgl_ard_nmf_dense(handle(), A.data(), At.data(), A.rows(), A.cols(), tol, maxit, L1, L2, k, w.data(), d.data(), h.data(), seed,
                            inv_density, overfit_threshold, trace_test_mse, &tr, &cb));
    Rcpp::NumericVector test_mse(mse.begin(), mse.begin() + tr.length), fit_tol(ft.begin(), ft.begin() + tr.length),
        score(so.begin(), so.begin() + tr.length);
    Rcpp::IntegerVector iter(it.begin(), it.begin() + tr.length);
    return Rcpp::List::create(Rcpp::Named("w") = w, Rcpp::Named("d") = d, Rcpp::Named("h") = h, Rcpp::Named("test_mse") = test_mse,
                              Rcpp::Named("iter") = iter, Rcpp::Named("tol") = fit_tol, Rcpp::Named("score_overfit") = score);
}

// "next" row f2: linked NMF (reference src/singlet.cpp:1059-1086, called by R/RunLNMF.R:60)
//[[Rcpp::export]]
Rcpp::List c_linked_nmf(Rcpp::SparseMatrix A, Rcpp::SparseMatrix At, const double tol, const uint16_t maxit, const bool verbose,
                        const double L1, const double L2, const uint16_t threads, Eigen::MatrixXd w, Eigen::MatrixXd link_h,
                        Eigen::MatrixXd link_w) {
    const int k = (int)w.rows();
    Eigen::MatrixXd h(k, (int64_t)A.cols());
    Eigen::VectorXd d(k);
    Progress p{verbose, false};
    sgl_callbacks cb = callbacks(p);
    if (verbose) Rprintf("\n%4s | %8s \n---------------\n", "iter", "tol");
    sgl_csc a = view(A), at = view(At);
    check(sgl_linked_nmf(handle(), &a, &at, tol, maxit, L1, L2, k, w.data(), d.data(), h.data(), link_h.data(), (int)link_h.rows(),
                         link_h.cols(), link_w.data(), (int)link_w.rows(), link_w.cols(), nullptr, nullptr, &cb));
    return Rcpp::List::create(Rcpp::Named("w") = w, Rcpp::Named("d") = d, Rcpp::Named("h") = h);
}

//[[Rcpp::export]]
Rcpp::List c_ard_nmf(Rcpp::SparseMatrix& A, Rcpp::SparseMatrix& At, const double tol, const uint16_t maxit, const bool verbose,
                     const double L1, const double L2, const uint16_t threads, Eigen::MatrixXd w, const uint64_t seed,
                     const uint64_t inv_density, const double overfit_threshold, const uint16_t trace_test_mse) {
    return ard_impl({view(A)}, {view(At)}, tol, maxit, verbose, L1, L2, w, seed, inv_density, overfit_threshold, trace_test_mse);
}

//[[Rcpp::export]]
Rcpp::List c_ard_nmf_sparse_list(Rcpp::List A_, Rcpp::List At_, const double tol, const uint16_t maxit, const bool verbose,
                                 const double L1, const double L2, const uint16_t threads, Eigen::MatrixXd w,
                                 const uint64_t rng_seed, const uint64_t inv_density, const double overfit_threshold,
                                 const uint16_t trace_test_mse) {
    std::vector<Rcpp::SparseMatrix> A = as_list(A_), At = as_list(At_);
    return ard_impl(views(A), views(At), tol, maxit, verbose, L1, L2, w, rng_seed, inv_density, overfit_threshold, trace_test_mse);
}

//[[Rcpp::export]]
Rcpp::List c_project_model(Rcpp::SparseMatrix A, Eigen::MatrixXd w, const double L1, const double L2, const int threads) {
    const int64_t m = A.rows();
    const int k = (int)((w.rows() == m) ? w.cols() : w.rows());
    Eigen::MatrixXd h(k, (int64_t)A.cols());
    Eigen::VectorXd d(k);
    sgl_csc a = view(A);
    check(sgl_project_model(handle(), &a, 1, w.data(), w.rows(), w.cols(), L1, L2, h.data(), d.data()));
    return Rcpp::List::create(Rcpp::Named("h") = h, Rcpp::Named("d") = d);
}

//[[Rcpp::export]]
Eigen::MatrixXd Rcpp_predict(Rcpp::SparseMatrix A, Eigen::MatrixXd w, const double L1, const double L2, const int threads) {
    const int64_t m = A.rows();
    const int k = (int)((w.rows() == m && w.cols() != m) ? w.cols() : w.rows());
    Eigen::MatrixXd h(k, (int64_t)A.cols());
    sgl_csc a = view(A);
    check(sgl_predict(handle(), &a, 1, w.data(), w.rows(), w.cols(), L1, L2, h.data()));
    return h;
}

// Optional new export: the whole (rank, replicate) grid of R/cross_validate_nmf.R:69-97 in ONE call.
// `w_inits` is a list of k_j x m matrices (the `w_init[[rep]][1:k, ]` slices the R loop passes one at a time), `seeds` the
// matching `abs(.Random.seed[[3 + rep]])` values. Returns a list of the same model lists c_ard_nmf returns, in order.
// The R loop body becomes:  models <- c_ard_nmf_batch(A, At, tol, maxit, L1, L2, threads, w_inits, seeds, inv_density, ...)
//[[Rcpp::export]]
Rcpp::List c_ard_nmf_batch(Rcpp::SparseMatrix& A, Rcpp::SparseMatrix& At, const double tol, const uint16_t maxit, const double L1,
                           const double L2, const uint16_t threads, Rcpp::List w_inits, Rcpp::NumericVector seeds,
                           const uint64_t inv_density, const double overfit_threshold, const uint16_t trace_test_mse) {
    const int n_jobs = w_inits.size();
    const int64_t n = A.cols();
    const int cap = (int)maxit + 2;
    std::vector<Eigen::MatrixXd> w(n_jobs), h(n_jobs);
    std::vector<Eigen::VectorXd> d(n_jobs);
    std::vector<std::vector<double>> mse(n_jobs), ft(n_jobs), so(n_jobs);
    std::vector<std::vector<int32_t>> it(n_jobs);
    std::vector<sgl_trace> tr(n_jobs);
    std::vector<sgl_fit_job> jobs(n_jobs);
    for (int j = 0; j < n_jobs; ++j) {
        w[j] = Rcpp::as<Eigen::MatrixXd>(w_inits[j]);
        const int k = (int)w[j].rows();
        h[j].resize(k, n);
        d[j].resize(k);
        mse[j].resize(cap); ft[j].resize(cap); so[j].resize(cap); it[j].resize(cap);
        tr[j] = sgl_trace{mse[j].data(), it[j].data(), ft[j].data(), so[j].data(), cap, 0};
        jobs[j] = sgl_fit_job{k, 0, (uint64_t)seeds[j], w[j].data(), d[j].data(), h[j].data(), &tr[j]};
    }
    Progress p{false, true};
    sgl_callbacks cb = callbacks(p);  // only poll_interrupt is used by the batch entry point
    sgl_csc a = view(A), at = view(At);
    check(sgl_ard_nmf_batch(handle(), &a, 1, &at, 1, tol, maxit, L1, L2, inv_density, overfit_threshold, trace_test_mse, jobs.data(),
                            n_jobs, 0, &cb));
    Rcpp::List out(n_jobs);
    for (int j = 0; j < n_jobs; ++j) {
        const int q = tr[j].length;
        out[j] = Rcpp::List::create(Rcpp::Named("w") = w[j], Rcpp::Named("d") = d[j], Rcpp::Named("h") = h[j],
                                    Rcpp::Named("test_mse") = Rcpp::NumericVector(mse[j].begin(), mse[j].begin() + q),
                                    Rcpp::Named("iter") = Rcpp::IntegerVector(it[j].begin(), it[j].begin() + q),
                                    Rcpp::Named("tol") = Rcpp::NumericVector(ft[j].begin(), ft[j].begin() + q),
                                    Rcpp::Named("score_overfit") = Rcpp::NumericVector(so[j].begin(), so[j].begin() + q));
    }
    return out;
}

// ---- several GPUs from the one R process ------------------------------------------------------------------
// The chunk-list entry points are what shards: with options(singlet.gpus = G > 1) the two *_sparse_list functions above are
// replaced by these bodies. The plain fit ignores `At_` (the transposed blocks are built on the devices), so for run_nmf-style
// calls the "distributed transpose" of R/cross_validate_nmf.R:37-50 can be dropped; the masked fit uses it.
namespace {
sgl_multi* multi(int n_gpus) {  // one multi-GPU context per R session
    static sgl_multi* mg = nullptr;
    if (!mg && sgl_multi_create(n_gpus, nullptr, &mg) != SGL_OK) Rcpp::stop(sgl_last_error());
    return mg;
}
}  // namespace

Rcpp::List c_nmf_sparse_list_multi(int n_gpus, Rcpp::List A_, const double tol, const uint16_t maxit, const bool verbose,
                                   const double L1, const double L2, Eigen::MatrixXd w) {
    std::vector<Rcpp::SparseMatrix> A = as_list(A_);
    std::vector<sgl_csc> a = views(A);
    const int k = (int)w.rows();
    int64_t n = 0;
    for (auto& c : a) n += c.ncol;
    Eigen::MatrixXd h(k, n);
    Eigen::VectorXd d(k);
    Progress p{verbose, false};
    sgl_callbacks cb = callbacks(p);
    if (verbose) Rprintf("\n%4s | %8s \n---------------\n", "iter", "tol");
    check(sgl_multi_nmf(multi(n_gpus), a.data(), (int)a.size(), nullptr, 0, tol, maxit, L1, L1, L2, L2, k, w.data(), d.data(), h.data(),
                        nullptr, nullptr, &cb));
    return Rcpp::List::create(Rcpp::Named("w") = w, Rcpp::Named("d") = d, Rcpp::Named("h") = h);
}

// (the masked fit shards the genes for the W update too, so it takes the reference's "distributed transpose" list At_ as it is)
Rcpp::List c_ard_nmf_sparse_list_multi(int n_gpus, Rcpp::List A_, Rcpp::List At_, const double tol, const uint16_t maxit, const bool verbose,
                                       const double L1, const double L2, Eigen::MatrixXd w, const uint64_t rng_seed,
                                       const uint64_t inv_density, const double overfit_threshold, const uint16_t trace_test_mse) {
    std::vector<Rcpp::SparseMatrix> A = as_list(A_), At = as_list(At_);
    std::vector<sgl_csc> a = views(A), at = views(At);
    const int k = (int)w.rows();
    int64_t n = 0;
    for (auto& c : a) n += c.ncol;
    Eigen::MatrixXd h(k, n);
    Eigen::VectorXd d(k);
    const int cap = (int)maxit + 2;
    std::vector<double> mse(cap), ft(cap), so(cap);
    std::vector<int32_t> it(cap);
    sgl_trace tr{mse.data(), it.data(), ft.data(), so.data(), cap, 0};
    Progress p{verbose, true};
    sgl_callbacks cb = callbacks(p);
    check(sgl_multi_ard_nmf(multi(n_gpus), a.data(), (int)a.size(), at.data(), (int)at.size(), tol, maxit, L1, L2, k, w.data(), d.data(), h.data(),
                            rng_seed, inv_density, overfit_threshold, trace_test_mse, &tr, &cb));
    return Rcpp::List::create(Rcpp::Named("w") = w, Rcpp::Named("d") = d, Rcpp::Named("h") = h,
                              Rcpp::Named("test_mse") = Rcpp::NumericVector(mse.begin(), mse.begin() + tr.length),
                              Rcpp::Named("iter") = Rcpp::IntegerVector(it.begin(), it.begin() + tr.length),
                              Rcpp::Named("tol") = Rcpp::NumericVector(ft.begin(), ft.begin() + tr.length),
                              Rcpp::Named("score_overfit") = Rcpp::NumericVector(so.begin(), so.begin() + tr.length));
}

// ---- IVSparse wire formats (reference src/singlet.cpp:783-995) -----------------------------------------
// save_IVSparse / build_IVCSC2 / write_IVCSC / read_IVSparse keep their signatures and file names; the codec is the library's.
namespace {
std::vector<unsigned char> ivcsc_image(std::vector<sgl_csc>& a, int level) {
    const int64_t n = sgl_ivsparse_encode(a.data(), (int)a.size(), level, nullptr, 0);
    if (n < 0) Rcpp::stop(sgl_last_error());
    std::vector<unsigned char> img((size_t)n);
    if (sgl_ivsparse_encode(a.data(), (int)a.size(), level, img.data(), (uint64_t)img.size()) < 0) Rcpp::stop(sgl_last_error());
    return img;
}
void write_file(const char* path, const std::vector<unsigned char>& img) {
    FILE* fp = std::fopen(path, "wb");
    if (!fp) Rcpp::stop("cannot open the output file");
    std::fwrite(img.data(), 1, img.size(), fp);
    std::fclose(fp);
}
}  // namespace

//[[Rcpp::export]]
bool save_IVSparse(Rcpp::List A_, bool verbose = true) {
    std::vector<Rcpp::SparseMatrix> A = as_list(A_);
    std::vector<sgl_csc> a = views(A);
    if (verbose) Rprintf("writing to IVCSC_matrix.ivsparse\n");
    write_file("IVCSC_matrix.ivsparse", ivcsc_image(a, 3));
    return true;
}

//[[Rcpp::export]]
Rcpp::List read_IVSparse_slots() {  // the dgCMatrix slots of IVCSC_matrix.ivsparse (read_IVSparse wraps them into a dgCMatrix)
    FILE* fp = std::fopen("IVCSC_matrix.ivsparse", "rb");
    if (!fp) Rcpp::stop("cannot open IVCSC_matrix.ivsparse");
    std::fseek(fp, 0, SEEK_END);
    const long bytes = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    std::vector<unsigned char> img((size_t)bytes);
    if (std::fread(img.data(), 1, img.size(), fp) != img.size()) Rcpp::stop("short read");
    std::fclose(fp);
    int32_t level = 0, vb = 0;
    int64_t nrow = 0, ncol = 0, nnz = 0;
    check(sgl_ivsparse_info(img.data(), (uint64_t)img.size(), &level, &nrow, &ncol, &nnz, &vb));
    Rcpp::IntegerVector p((int)ncol + 1), i((int)nnz), dim(2);
    Rcpp::NumericVector x((int)nnz);
    if (sgl_ivsparse_decode(img.data(), (uint64_t)img.size(), 0, ncol, p.begin(), i.begin(), x.begin(), nnz) < 0) Rcpp::stop(sgl_last_error());
    dim[0] = (int)nrow;
    dim[1] = (int)ncol;
    return Rcpp::List::create(Rcpp::Named("p") = p, Rcpp::Named("i") = i, Rcpp::Named("x") = x, Rcpp::Named("Dim") = dim);
}

//[[Rcpp::export]]
Rcpp::List run_nmf_on_sparsematrix_list(Rcpp::List A_, const double tol, const uint16_t maxit, const bool verbose, const uint16_t threads,
                                        Eigen::MatrixXd w, bool use_vcsc = false, const double L1 = 0, const double L2 = 0) {
    // the reference packs the list into one IVCSC / VCSC matrix of FLOAT values (:790-822) and runs plain ALS on it (:946-995):
    // narrow the values through the codec, then fit the decoded chunks (t(A) is built on the device)
    std::vector<Rcpp::SparseMatrix> A = as_list(A_);
    std::vector<sgl_csc> a = views(A);
    std::vector<unsigned char> img = ivcsc_image(a, use_vcsc ? 2 : 3);
    std::vector<std::vector<int32_t>> P(a.size()), I(a.size());
    std::vector<std::vector<double>> X(a.size());
    std::vector<sgl_csc> chunks(a.size());
    int64_t col0 = 0;
    for (size_t q = 0; q < a.size(); ++q) {
        P[q].resize((size_t)a[q].ncol + 1);
        const int64_t nnz = sgl_ivsparse_decode(img.data(), (uint64_t)img.size(), col0, a[q].ncol, P[q].data(), nullptr, nullptr, 0);
        if (nnz < 0) Rcpp::stop(sgl_last_error());
        I[q].resize((size_t)nnz);
        X[q].resize((size_t)nnz);
        if (sgl_ivsparse_decode(img.data(), (uint64_t)img.size(), col0, a[q].ncol, P[q].data(), I[q].data(), X[q].data(), nnz) < 0)
            Rcpp::stop(sgl_last_error());
        chunks[q] = sgl_csc{a[q].nrow, a[q].ncol, P[q].data(), I[q].data(), X[q].data()};
        col0 += a[q].ncol;
    }
    return nmf_impl(chunks, {}, tol, maxit, verbose, L1, L1, L2, L2, w);
}
