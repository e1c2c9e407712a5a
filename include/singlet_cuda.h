/* singlet_cuda.h -- C ABI of libsinglet_cuda.so: the B200 (sm_100a) ALS-NMF engine that replaces the
 * C++ core of zdebruine/singlet behind its Rcpp entry points.
 *
 * Every function returns 0 on success or a negative SGL_E* code; the message is available from
 * sgl_last_error() (thread-local). No C++ exception and no longjmp ever crosses this boundary
 * (the reference relies on BEGIN_RCPP/END_RCPP, src/RcppExports.cpp:18,26). There is no CPU
 * fallback: without a CUDA device every compute entry point fails with SGL_ENODEVICE.
 *
 * Matrix conventions are the reference's: factors are column-major k x cols doubles
 * (w is k x m, h is k x n, element (f, c) at c*k + f); sparse inputs are dgCMatrix slot views
 * (inst/include/singlet.h:36-41) given as a list of column chunks (n_chunks = 1 for one matrix).
 *
 * Two layers:
 *   (1) host-facing entry points -- what src/RcppExports.cpp binds, one per reference routine;
 *   (2) device-level building blocks (sgl_dev_*) on caller-owned device buffers and the handle's
 *       stream -- used by the one-process-per-GPU driver (singlet_b200/sharded.py) which supplies
 *       the NCCL collectives between them, and by the parity tests.
 */
#ifndef SINGLET_CUDA_H
#define SINGLET_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGL_OK 0
#define SGL_EINVAL (-1)      /* bad argument (shape, NULL, k out of range ...) */
#define SGL_ENODEVICE (-2)   /* no CUDA device / wrong architecture */
#define SGL_ECUDA (-3)       /* CUDA runtime error (message in sgl_last_error) */
#define SGL_EINTERRUPT (-4)  /* poll_interrupt callback asked to stop */
#define SGL_ENOMEM (-5)

#define SGL_MAX_RANK 128

/* dgCMatrix view: replaces Rcpp::SparseMatrix (inst/include/singlet.h:36-41). Read-only, caller
 * owned, must stay valid for the duration of the call only. A view may also be a COLUMN RANGE of a larger
 * dgCMatrix, zero-copy: p points at the parent's p[c0] (so p[0] is the offset of the range's first non-zero)
 * while i and x still point at the parent's arrays. */
typedef struct sgl_csc {
    int64_t nrow, ncol;
    const int32_t* p; /* ncol + 1 */
    const int32_t* i; /* nnz, 0-based rows, ascending within a column */
    const double* x;  /* nnz */
} sgl_csc;

/* Replaces Rprintf / Rcpp::checkUserInterrupt (src/singlet.cpp:643-663, 1102-1128). Both callbacks
 * are invoked on the calling thread only, between half-iterations. Either may be NULL. */
typedef struct sgl_callbacks {
    void* user;
    int (*poll_interrupt)(void* user);                                  /* non-zero -> abort fit */
    void (*on_iter)(void* user, int iter, double tol, double overfit);  /* overfit = NaN if untraced */
} sgl_callbacks;

/* CV trace vectors of c_ard_nmf (src/singlet.cpp:1144-1151). Caller allocates `capacity` entries
 * (maxit + 2 is always enough); the library sets `length`. */
typedef struct sgl_trace {
    double* test_mse;
    int32_t* iter;
    double* tol;
    double* score_overfit;
    int32_t capacity;
    int32_t length;
} sgl_trace;

typedef struct sgl_handle sgl_handle; /* one device + one stream + cached device matrices */
typedef struct sgl_matrix sgl_matrix; /* device-resident sparse matrix in the engine's tiled layout */

/* ---- library / handle ------------------------------------------------------------------- */
int sgl_version(void);
const char* sgl_last_error(void);
int sgl_device_count(void);
/* stream: a cudaStream_t to run on (e.g. torch's current stream) or NULL to create one. */
int sgl_create(int device, void* stream, sgl_handle** out);
int sgl_destroy(sgl_handle* h);
/* Keep uploaded A/At (and the materialised mask) between calls when the same host buffers are
 * passed again (a CV sweep makes dozens of calls on one matrix). Default on. */
int sgl_set_cache(sgl_handle* h, int enabled);
/* Operand precision of the sparse product b = sum v * F[r, :] (src/singlet.cpp:341-343). Accumulation is FP32 in both
 * modes, the non-zero values v stay FP32, the solver, Gram, scale, cor and loss kernels are unaffected.
 *   SGL_PRECISION_MIXED16 (default): for padded ranks >= 32, on matrices with >= 256 non-zeros per row AND per column on
 *       average, the gathered factor F is staged as FP16 scaled by a power of two taken from max |F| and the non-zero
 *       values as stochastically rounded FP16 scaled by a power of two taken from max |v| (zero-mean relative rounding
 *       of 2^-12 / 2^-11 per element, averaged over the non-zeros of a column: ~1e-5 of a right-hand side at the BASELINE
 *       configs); every product is exact in FP32 and accumulated in FP32. This halves the shared-memory gather that bounds
 *       the kernel. Smaller matrices and ranks keep FP32 operands.
 *   SGL_PRECISION_FP32: F and v are gathered in FP32 (the round-1 kernel) whatever the matrix.
 *   SGL_PRECISION_MIXED16_ALWAYS: 16-bit staging for every matrix at padded ranks >= 32 (tests).
 * The environment variable SGL_PRECISION=fp32 | always | mixed16 sets the mode of handles created afterwards. */
#define SGL_PRECISION_MIXED16 0
#define SGL_PRECISION_FP32 1
#define SGL_PRECISION_MIXED16_ALWAYS 2
int sgl_set_precision(sgl_handle* h, int mode);
int sgl_get_precision(sgl_handle* h);
int sgl_synchronize(sgl_handle* h);
/* the cudaStream_t all work of this handle is enqueued on */
void* sgl_stream(sgl_handle* h);
/* kernel launches issued through this handle so far (bench.py's gpu_launches) */
int64_t sgl_launch_count(sgl_handle* h);

/* Per-kernel-kind device timing with CUDA events on the handle's stream (kind 0 = SpMM, 1 = NNLS,
 * 2 = Gram, 3 = other). sgl_profile_read synchronises, returns summed milliseconds, launch counts and
 * algorithmic bytes per kind since the last read, and resets them. */
int sgl_profile(sgl_handle* h, int enable);
int sgl_profile_read(sgl_handle* h, double* ms4, int64_t* counts4, int64_t* bytes4);

/* ---- host-facing entry points (one per reference routine) -------------------------------- */

/* c_nmf / c_nmf_sparse_list: src/singlet.cpp:638-672, 715-743 (RcppExports.cpp:97-116, 138-155).
 * w: k x m in (w_init) / out; d: k out; h: k x n out. iters_out / tol_out may be NULL.
 * At may be NULL (nAt = 0): the transpose is then built on the device (sgl_matrix_transpose). */
int sgl_nmf(sgl_handle* h, const sgl_csc* A, int nA, const sgl_csc* At, int nAt, double tol, uint16_t maxit,
            double L1_w, double L1_h, double L2_w, double L2_h, int k, double* w, double* d, double* h_out,
            int32_t* iters_out, double* tol_out, const sgl_callbacks* cb);

/* c_linked_nmf: src/singlet.cpp:1059-1086 with predict_link :416-433 (RcppExports.R:74-76). link_h is
 * lh_rows x lh_cols, link_w is lw_rows x lw_cols (column-major doubles); like the reference a side is linked only
 * when its matrix has one column per cell (link_h) / per gene (link_w), otherwise that half-iteration is plain. */
int sgl_linked_nmf(sgl_handle* h, const sgl_csc* A, const sgl_csc* At, double tol, uint16_t maxit, double L1, double L2,
                   int k, double* w, double* d, double* h_out, const double* link_h, int lh_rows, int64_t lh_cols,
                   const double* link_w, int lw_rows, int64_t lw_cols, int32_t* iters_out, double* tol_out,
                   const sgl_callbacks* cb);

/* c_ard_nmf / c_ard_nmf_sparse_list: src/singlet.cpp:1090-1234 (RcppExports.cpp:283-327). */
int sgl_ard_nmf(sgl_handle* h, const sgl_csc* A, int nA, const sgl_csc* At, int nAt, double tol, uint16_t maxit,
                double L1, double L2, int k, double* w, double* d, double* h_out, uint64_t seed,
                uint64_t inv_density, double overfit_threshold, uint16_t trace_test_mse, sgl_trace* trace,
                const sgl_callbacks* cb);

/* Rank-search batching: the fits of a cross-validation sweep -- the loop over
 * (rank, replicate) of R/cross_validate_nmf.R:69-97 and R/ard_nmf.R:95-159, one c_ard_nmf call each -- are
 * independent given A, so the library runs `concurrency` of them at a time on private streams (each of
 * these small fits is latency-bound and leaves most of the chip idle). A and At are uploaded once and
 * shared; every worker keeps its own mask and factor buffers. Results are bit-identical to n_jobs
 * sequential sgl_ard_nmf calls. concurrency <= 0 lets the library choose (bounded by free device
 * memory). cb->poll_interrupt is polled on the calling thread while the workers run; on_iter is not used. */
typedef struct sgl_fit_job {
    int32_t k;        /* rank of this fit */
    int32_t status;   /* out: SGL_OK or this fit's error code */
    uint64_t seed;    /* mask seed (rng state) of this fit */
    double* w;        /* k x m in (w_init) / out */
    double* d;        /* k out */
    double* h;        /* k x n out */
    sgl_trace* trace; /* out, capacity set by the caller */
} sgl_fit_job;
int sgl_ard_nmf_batch(sgl_handle* h, const sgl_csc* A, int nA, const sgl_csc* At, int nAt, double tol, uint16_t maxit,
                      double L1, double L2, uint64_t inv_density, double overfit_threshold, uint16_t trace_test_mse,
                      sgl_fit_job* jobs, int32_t n_jobs, int32_t concurrency, const sgl_callbacks* cb);

/* Dense-input variants c_nmf_dense / c_ard_nmf_dense: src/singlet.cpp:1051-1054, 1357-1361 (RcppExports.cpp:241-260,
 * 329-350) with predict / predict_mask / mse_test on Eigen::MatrixXd (:370-381, 506-531, 610-634). A is m x n, At is
 * n x m, column-major doubles. Every entry (zeros included) takes part, no column is skipped -- like the reference. */
int sgl_nmf_dense(sgl_handle* h, const double* A, const double* At, int64_t m, int64_t n, double tol, uint16_t maxit,
                  double L1_w, double L1_h, double L2_w, double L2_h, int k, double* w, double* d, double* h_out,
                  int32_t* iters_out, double* tol_out, const sgl_callbacks* cb);
int sgl_ard_nmf_dense(sgl_handle* h, const double* A, const double* At, int64_t m, int64_t n, double tol, uint16_t maxit,
                      double L1, double L2, int k, double* w, double* d, double* h_out, uint64_t seed, uint64_t inv_density,
                      double overfit_threshold, uint16_t trace_test_mse, sgl_trace* trace, const sgl_callbacks* cb);

/* c_project_model: src/singlet.cpp:405-413 (RcppExports.cpp:82-95). w is w_rows x w_cols column-major;
 * it is transposed when w_rows == nrow(A) exactly as the reference does. h: k x n out, d: k out. */
int sgl_project_model(sgl_handle* h, const sgl_csc* A, int nA, const double* w, int64_t w_rows, int64_t w_cols,
                      double L1, double L2, double* h_out, double* d_out);

/* Rcpp_predict: src/singlet.cpp:350-367 (RcppExports.cpp:67-80): one H update from h = 0, no scaling. */
int sgl_predict(sgl_handle* h, const sgl_csc* A, int nA, const double* w, int64_t w_rows, int64_t w_cols, double L1,
                double L2, double* h_out);

/* ---- test hooks for the bit-exact parts ----------------------------------------------------
 * rng::rand(i,j) / rng::draw (src/singlet.cpp:47-64, 91-95) evaluated ON THE DEVICE for n pairs. */
int sgl_mask_rand(sgl_handle* h, uint64_t seed, const uint64_t* i, const uint64_t* j, int64_t n, uint64_t* out);
int sgl_mask_draw(sgl_handle* h, uint64_t seed, uint64_t inv_density, const uint64_t* i, const uint64_t* j,
                  int64_t n, uint8_t* out);

/* ---- device-level building blocks ---------------------------------------------------------
 * Factor buffers are float [cols][KP] (KP = sgl_padded_rank(k): k rounded up to 4,8,16,32,64,128;
 * padding lanes hold 0). Scalars that are reduced across GPUs are double. All work is enqueued on
 * the handle's stream; nothing synchronises unless stated. */
int sgl_padded_rank(int k);

/* Upload a chunk list (concatenated by columns) and build the gather-tile index. */
int sgl_matrix_upload(sgl_handle* h, const sgl_csc* chunks, int n_chunks, sgl_matrix** out);
/* Device-side transpose (replaces `Matrix::t(A)` of R/run_nmf.R:40, R/cross_validate_nmf.R:58,
 * R/ard_nmf.R:81): a new device matrix holding m^T with sorted row indices, bit-identical to uploading the host
 * transpose. One pass over the records per 57,856 rows of m (a single pass for A, whose rows are genes). The host-facing
 * entry points (sgl_nmf, sgl_linked_nmf, sgl_ard_nmf, sgl_ard_nmf_batch) use it when At is NULL / nAt is 0. */
int sgl_matrix_transpose(sgl_handle* h, const sgl_matrix* m, sgl_matrix** out);
/* Deterministic synthetic sparse counts (exact rules in singlet_b200/synth.py),
 * generated on the device. Orientation 0: columns = cells [col0, col0+ncol) of the m x n matrix;
 * orientation 1: columns = genes [col0, col0+ncol) of its transpose. values_table: 8 floats. */
int sgl_matrix_synth(sgl_handle* h, int64_t m_genes, int64_t n_cells, double density, uint64_t data_seed,
                     int orientation, int64_t col0, int64_t ncol, const float* values_table, sgl_matrix** out);
/* Same, restricted to the rows [row0, row0 + nrows) of that orientation and re-based to 0 (a rank's block of the
 * transpose over its own cells only). */
int sgl_matrix_synth_block(sgl_handle* h, int64_t m_genes, int64_t n_cells, double density, uint64_t data_seed,
                           int orientation, int64_t col0, int64_t ncol, int64_t row0, int64_t nrows,
                           const float* values_table, sgl_matrix** out);
/* Copy the column pointers (int64[ncol + 1]) into a caller-owned DEVICE buffer (stream-ordered). */
int sgl_matrix_colptr(sgl_handle* h, const sgl_matrix* m, int64_t* dst_device);
int sgl_matrix_free(sgl_handle* h, sgl_matrix* m);
int sgl_matrix_info(const sgl_matrix* m, int64_t* nrow, int64_t* ncol, int64_t* nnz);
/* Copy back to host in dgCMatrix form (p may be int32 only if nnz < 2^31). Synchronises. */
int sgl_matrix_download(sgl_handle* h, const sgl_matrix* m, int32_t* p, int32_t* i, double* x);

/* double k x cols (column-major, host) <-> float [cols][KP] (device) */
int sgl_factor_upload(sgl_handle* h, const double* host_kxc, int k, int64_t cols, float* dev);
int sgl_factor_download(sgl_handle* h, const float* dev, int k, int64_t cols, double* host_kxc);

/* AAt (src/singlet.cpp:200-206) partial: gram[KP*KP] (double, device) = sum over the given columns
 * of f f^T. add_jitter != 0 adds the 1e-15 diagonal (do it once, after any all-reduce). */
int sgl_dev_gram(sgl_handle* h, const float* F, int k, int64_t cols, double* gram, int add_jitter);
int sgl_dev_gram_jitter(sgl_handle* h, int k, double* gram);

/* predict (src/singlet.cpp:333-347) for the columns of X: b = F_in . X[:, c] then coordinate-descent
 * NNLS against `gram` (double [KP*KP], jitter included) with warm start / result in F_out; empty
 * columns are skipped. rowsum[KP] (double, device) receives the row sums of the new F_out
 * (the local part of `scale`'s d, without the 1e-15). */
int sgl_dev_update(sgl_handle* h, const sgl_matrix* X, const float* F_in, float* F_out, int k, const double* gram,
                   double L1, double L2, double* rowsum);

/* The two halves of sgl_dev_update, for layouts where the right-hand sides are summed across GPUs before the
 * solve: sgl_dev_rhs writes B[ncol][KP] = F_in . X (float, device); sgl_dev_solve runs the NNLS on B for `ncol`
 * columns. colptr_like is any int64[ncol + 1] device array with colptr_like[c] == colptr_like[c+1] exactly for the
 * columns that are empty in the GLOBAL matrix (they are skipped like src/singlet.cpp:340). */
int sgl_dev_rhs(sgl_handle* h, const sgl_matrix* X, const float* F_in, int k, float* B_out);
int sgl_dev_solve(sgl_handle* h, const float* B, const int64_t* colptr_like, int64_t ncol, float* F_out, int k,
                  const double* gram, double L1, double L2, double* rowsum);

/* sgl_dev_update in two calls on the handle's own right-hand-side scratch, so that a driver can enqueue the product
 * (which only reads F_in) ahead of time: sgl_dev_update_rhs leaves b = F_in . X in the handle and returns a ticket;
 * sgl_dev_update_solve runs the NNLS of sgl_dev_update on it and fails with SGL_EINVAL when another product on this
 * handle has overwritten the scratch since (compare the ticket with sgl_dev_rhs_epoch first and redo the product). */
int sgl_dev_update_rhs(sgl_handle* h, const sgl_matrix* X, const float* F_in, int k, uint64_t* ticket_out);
uint64_t sgl_dev_rhs_epoch(const sgl_handle* h);
int sgl_dev_update_solve(sgl_handle* h, const sgl_matrix* X, uint64_t ticket, float* F_out, int k, const double* gram,
                         double L1, double L2, double* rowsum);

/* scale (src/singlet.cpp:219-225): F[c][f] /= d[f]; d is double[KP] on the device (already
 * all-reduced and with the 1e-15 added -- see sgl_dev_finish_d). */
int sgl_dev_finish_d(sgl_handle* h, int k, double* d_inout);
int sgl_dev_scale(sgl_handle* h, float* F, int k, int64_t cols, const double* d);
/* sgl_dev_finish_d, and in the same launch the Gram of the factor BEFORE scaling (sgl_dev_gram without jitter, summed
 * over the ranks) becomes the Gram of the scaled factor: gram[i][j] / (d[i] d[j]), 1e-15 added to the diagonal
 * (src/singlet.cpp:206). With it the row sums of `scale` and the partial Grams travel in ONE all-reduce. */
int sgl_dev_finish_d_rescale_gram(sgl_handle* h, int k, double* d_inout, double* gram_inout);

/* cor (src/singlet.cpp:184-197): the five running sums over the given columns -> sums[5] (double,
 * device): sum x, sum y, sum xy, sum x^2, sum y^2. sgl_cor_from_sums finishes on the host. */
int sgl_dev_cor_sums(sgl_handle* h, const float* X, const float* Y, int k, int64_t cols, double* sums);
double sgl_cor_from_sums(const double* sums5, double n_elems);

/* Speckled mask (src/singlet.cpp:436-466, 536-568), materialised once per (seed, inv_density):
 * a training copy of X with held-out non-zeros zeroed plus the per-column held-out index lists.
 * mask_t = 0: columns are cells (hash(col+col_offset, row+row_offset)); 1: columns are genes. */
typedef struct sgl_mask sgl_mask;
int sgl_mask_build(sgl_handle* h, const sgl_matrix* X, uint64_t seed, uint64_t inv_density, int mask_t,
                   int64_t col_offset, int64_t row_offset, sgl_mask** out);
int sgl_mask_free(sgl_handle* h, sgl_mask* m);
int sgl_mask_info(const sgl_mask* m, int64_t* n_masked, int64_t* n_masked_nonzero);
/* held-out row indices of one column (host copy; for bit-exactness tests). Returns count. */
int64_t sgl_mask_column(sgl_handle* h, const sgl_mask* m, int64_t col, int32_t* rows_out, int64_t capacity);

/* predict_mask (src/singlet.cpp:436-466) */
int sgl_dev_update_masked(sgl_handle* h, const sgl_matrix* X, const sgl_mask* mask, const float* F_in, float* F_out,
                          int k, const double* gram, double L1, double L2, double* rowsum);
/* mse_test (src/singlet.cpp:536-568) over the cell columns of `mask` (mask_t = 0): writes the SUM of
 * per-column losses to loss_sum[0] (double, device); divide by the global n afterwards.
 * which = 0: held-out entries (test); 1: the entries that are not held out (train, harness-defined); 2: both in ONE pass
 * over the held-out lists and the non-zeros (the fused train/test loss kernel): loss_sum[0] = test, loss_sum[1] = train. */
int sgl_dev_mse(sgl_handle* h, const sgl_matrix* A, const sgl_mask* mask, const float* W, const double* d,
                const float* H, int k, int which, double* loss_sum);

/* ---- IVSparse wire formats (csrc/ivsparse.cpp) -------------------------------
 * Host-side codec for the file images of the reference's vendored IVSparse library: IVCSC (compression level 3,
 * inst/include/src/IVCSC/IVCSC_Methods.hpp:72-95, IVCSC_Private_Methods.hpp:128-299) and VCSC (level 2,
 * inst/include/src/VCSC/VCSC_Methods.hpp:77-109), as written by write_IVCSC / save_IVSparse / build_IVCSC2 and read by
 * read_IVSparse / run_nmf_on_sparsematrix_list (src/singlet.cpp:783-995). `image` is the whole file in memory. */
int sgl_ivsparse_info(const void* image, uint64_t bytes, int32_t* level, int64_t* nrow, int64_t* ncol, int64_t* nnz,
                      int32_t* value_bytes);
/* Columns [col0, col0 + ncol) as dgCMatrix slots (p has ncol + 1 entries, rows ascending within a column). Returns the
 * non-zeros of the range; with i = x = NULL only p is filled (sizing call). A range must hold < 2^31 non-zeros: decode an
 * atlas-scale file as a list of column chunks and hand that list to sgl_nmf / sgl_multi_nmf. */
int64_t sgl_ivsparse_decode(const void* image, uint64_t bytes, int64_t col0, int64_t ncol, int32_t* p, int32_t* i, double* x,
                            int64_t capacity);
/* The image the reference's IVCSC (level 3) / VCSC (level 2) type writes for this chunk list (float values, 8-byte index
 * type in the metadata). Returns its size; writes it when out != NULL. */
int64_t sgl_ivsparse_encode(const sgl_csc* chunks, int n_chunks, int level, void* out, uint64_t capacity);

/* ---- multi-GPU (csrc/multi.cu) ---------------------------------------------------
 * The reference's chunk-list entry points (c_nmf_sparse_list src/singlet.cpp:715-743, c_ard_nmf_sparse_list :1162-1234,
 * chunks and "distributed transpose" blocks built at R/cross_validate_nmf.R:37-50) map 1:1 onto GPUs: cells are sharded
 * for the H update, genes for the W update; NCCL (bound at run time with dlopen) carries the k x m / k x n factor
 * all-gathers, the reduce-scatter of the W-update right-hand sides and the all-reduces of Gram / row sums / loss.
 * Shards are always the contiguous ranges of sgl_shard_bounds. */
void sgl_shard_bounds(int64_t total, int world, int rank, int64_t* lo, int64_t* hi, int64_t* per);

/* One rank of a GPU group: an sgl_handle plus an NCCL communicator on the handle's stream. */
typedef struct sgl_comm sgl_comm;
/* one-process-per-GPU jobs: rank 0 makes the 128-byte id, every rank receives it (any host transport) and joins */
int sgl_comm_unique_id(void* id128);
int sgl_comm_init_rank(sgl_handle* h, int device, int world, int rank, const void* id128, sgl_comm** out);
int sgl_comm_destroy(sgl_comm* c);
int sgl_comm_rank(const sgl_comm* c);
int sgl_comm_world(const sgl_comm* c);
sgl_handle* sgl_comm_handle(const sgl_comm* c);
int64_t sgl_comm_collectives(const sgl_comm* c); /* NCCL calls issued so far */

/* Device state of one sharded fit on one rank. masked == 0 (c_nmf): A_loc = this rank's cells (m x n_loc), At_loc = the
 * transpose of that block (n_loc x m) or NULL to build it on the device. masked != 0 (c_ard_nmf): At_loc = this rank's
 * genes over ALL cells (n_total x g_loc), required. w_init: k x m, identical on every rank. */
typedef struct sgl_fit sgl_fit;
int sgl_fit_create(sgl_comm* c, const sgl_matrix* A_loc, const sgl_matrix* At_loc, int64_t n_total, int k, const double* w_init,
                   int masked, uint64_t seed, uint64_t inv_density, sgl_fit** out);
/* one trip of src/singlet.cpp:648-659 (:1108-1114 when masked), collectives included; tol_out = 1 - cor(w, w_prev).
 * stop_flag (optional, in/out): set to 1 on any rank to make EVERY rank leave with 1 (agreed through the all-reduce). */
int sgl_fit_iterate(sgl_fit* f, double L1_w, double L1_h, double L2_w, double L2_h, double* tol_out, int* stop_flag);
/* Plain fits enqueue the part of the NEXT iteration that only reads W (its Gram, the H-update product) before the host waits
 * for tol, so the device does not idle across the host round trip; results are unchanged. on = 0 before the last iteration
 * of a fit avoids one wasted product (sgl_nmf_rank does). Default: on. */
int sgl_fit_set_lookahead(sgl_fit* f, int on);
int sgl_fit_test_mse(sgl_fit* f, double* out);                            /* mse_test, all-reduced */
int sgl_fit_download(sgl_fit* f, double* w, double* d, double* h_local);  /* k x m, k, k x n_loc; any may be NULL */
int sgl_fit_shard(const sgl_fit* f, int64_t* c0, int64_t* c1, int64_t* g0, int64_t* g1);
int sgl_fit_destroy(sgl_fit* f);

/* Whole fits on one rank of a one-process-per-GPU job: every rank makes the same call with its shards; w (in/out), d, the
 * iteration count and the trace are identical on all ranks, h_local is this rank's k x n_loc block. */
int sgl_nmf_rank(sgl_comm* c, const sgl_matrix* A_loc, const sgl_matrix* At_loc, int64_t n_total, double tol, uint16_t maxit,
                 double L1_w, double L1_h, double L2_w, double L2_h, int k, double* w, double* d, double* h_local,
                 int32_t* iters_out, double* tol_out, const sgl_callbacks* cb);
int sgl_ard_nmf_rank(sgl_comm* c, const sgl_matrix* A_loc, const sgl_matrix* At_loc, int64_t n_total, double tol, uint16_t maxit,
                     double L1, double L2, int k, double* w, double* d, double* h_local, uint64_t seed, uint64_t inv_density,
                     double overfit_threshold, uint16_t trace_test_mse, sgl_trace* trace, const sgl_callbacks* cb);

/* All devices of ONE process (what an R session is): ncclCommInitAll, one host thread per device, callbacks on the
 * calling thread only. devices == NULL: 0 .. n_devices - 1. */
typedef struct sgl_multi sgl_multi;
int sgl_multi_create(int n_devices, const int* devices, sgl_multi** out);
int sgl_multi_destroy(sgl_multi* mg);
int sgl_multi_size(const sgl_multi* mg);
sgl_comm* sgl_multi_rank(const sgl_multi* mg, int rank);
int sgl_multi_set_precision(sgl_multi* mg, int mode);
/* c_nmf_sparse_list / c_ard_nmf_sparse_list over the devices: the chunk lists are re-cut into one shard per device without
 * copying (column-range views). Same arguments and results as sgl_nmf / sgl_ard_nmf; sgl_multi_nmf ignores At (every
 * device transposes its own cell block), sgl_multi_ard_nmf needs the gene-block list. */
int sgl_multi_nmf(sgl_multi* mg, const sgl_csc* A, int nA, const sgl_csc* At, int nAt, double tol, uint16_t maxit, double L1_w,
                  double L1_h, double L2_w, double L2_h, int k, double* w, double* d, double* h_out, int32_t* iters_out,
                  double* tol_out, const sgl_callbacks* cb);
int sgl_multi_ard_nmf(sgl_multi* mg, const sgl_csc* A, int nA, const sgl_csc* At, int nAt, double tol, uint16_t maxit, double L1,
                      double L2, int k, double* w, double* d, double* h_out, uint64_t seed, uint64_t inv_density,
                      double overfit_threshold, uint16_t trace_test_mse, sgl_trace* trace, const sgl_callbacks* cb);

#ifdef __cplusplus
}
#endif
#endif /* SINGLET_CUDA_H */
