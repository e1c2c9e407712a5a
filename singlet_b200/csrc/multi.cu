// multi.cu -- the multi-GPU drivers of libsinglet_cuda.so, behind the C ABI (SURVEY.md 8e and the "Threading" row of 8b).
//
// The reference has one parallel loop over columns (src/singlet.cpp:336-346) and a serial loop over column chunks
// (:384-402, :469-503) whose "distributed transpose" gene blocks are built in R (R/cross_validate_nmf.R:37-50). Those
// chunk lists map 1:1 onto GPUs: cells (columns of A) are sharded for the H update, genes for the W update.
//
//   sgl_comm   one rank of a group of GPUs: an sgl_handle + an NCCL communicator on the handle's stream. Created
//              either by every process of a one-process-per-GPU job (sgl_comm_init_rank, unique id from rank 0) or for
//              all devices of ONE process at once (sgl_multi_create -> ncclCommInitAll), which is what an R session needs.
//   sgl_fit    the device state of one sharded fit on one rank; sgl_fit_iterate is one trip of src/singlet.cpp:648-659
//              (or :1108-1114 with the speckled mask), collectives included. Layouts:
//                plain   "B": only the cells are sharded. The rank holds its cell block A_loc (m x n_loc) and the transpose
//                        of that SAME block (n_loc x m, built on the device when not given); the W-update right-hand sides
//                        of all m genes are formed from the local cells, reduce-scattered (k x m floats), every rank solves
//                        one gene shard, W is all-gathered. H never leaves its rank.
//                masked  "A": the rank holds its cell block and its gene block over ALL cells (exactly the reference's chunk
//                        list and distributed-transpose block); the per-column Gram corrections stay local; H and W are
//                        all-gathered after each half-iteration.
//              Per half-iteration the only other exchanges are all-reduces of k (+ k^2) doubles: the row sums of `scale`
//              (:220) and the partial Gram (:200-206); `cor` runs redundantly on the replicated W so every rank takes the
//              same stopping decision; the CV loss adds one all-reduced double.
//   sgl_multi  single-process front end: sgl_multi_nmf / sgl_multi_ard_nmf take the reference's chunk lists (A_ = column
//              chunks, At_ = gene-block transposes), cut them into one shard per device WITHOUT copying (column-range views),
//              and drive every device from its own host thread. Callbacks only ever run on the calling thread.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy already loaded by the host application -- e.g. torch's --
// or the system one), so the library has no link-time dependency on it and single-GPU users never load it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

using namespace sgl;

// ---------------------------------------------------------------------------------------------
// NCCL, bound at run time
// ---------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};
NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // Order: an explicit path (SGL_NCCL_LIB; the Python binding points it at the NCCL bundled with torch), a libnccl the
        // process has already loaded, the system library. The loader keys libraries by soname, so whichever libnccl.so.2 comes
        // first is the one a later `import torch` binds to as well: loading the system 2.27 before torch (built against its
        // bundled 2.28) made torch fail with "undefined symbol: ncclDevCommCreate". RTLD_LOCAL: nothing is exported.
        const char* env = getenv("SGL_NCCL_LIB");
        if (env && env[0]) api.lib = dlopen(env, RTLD_NOW | RTLD_LOCAL);
        if (!api.lib) api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (api.lib) break;
            api.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        }
        if (!api.lib) {
            api.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found");
            return;
        }
        bool ok = true;
        auto bind = [&](const char* sym) {
            void* p = dlsym(api.lib, sym);
            if (!p) { ok = false; api.error = std::string("libnccl lacks ") + sym; }
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))bind("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))bind("ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll))bind("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))bind("ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce))bind("ncclAllReduce");
        api.ReduceScatter = (decltype(api.ReduceScatter))bind("ncclReduceScatter");
        api.AllGather = (decltype(api.AllGather))bind("ncclAllGather");
        api.GroupStart = (decltype(api.GroupStart))bind("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))bind("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))bind("ncclGetErrorString");
        if (!ok) { dlclose(api.lib); api.lib = nullptr; }
    });
    return &api;
}
int need_nccl(NcclApi** out) {
    NcclApi* a = nccl_api();
    if (!a->lib) return fail(SGL_ENODEVICE, "multi-GPU entry point needs NCCL: %s", a->error.c_str());
    *out = a;
    return SGL_OK;
}
#define SGL_NCCL(api, expr)                                                                                         \
    do {                                                                                                            \
        ncclResult_t _r = (expr);                                                                                   \
        if (_r != ncclSuccess) return fail(SGL_ECUDA, "%s failed: %s (%s:%d)", #expr, (api)->GetErrorString(_r), __FILE__, __LINE__); \
    } while (0)

__global__ void column_counts_kernel(const int64_t* __restrict__ colptr, int64_t ncol, int64_t* __restrict__ counts) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncol) counts[c] = colptr[c + 1] - colptr[c];
}
}  // namespace

extern "C" void sgl_shard_bounds(int64_t total, int world, int rank, int64_t* lo, int64_t* hi, int64_t* per) {
    const int64_t p = (total + world - 1) / world;
    const int64_t l = rank * p < total ? rank * p : total;
    const int64_t u = l + p < total ? l + p : total;
    if (lo) *lo = l;
    if (hi) *hi = u;
    if (per) *per = p;
}

// ---------------------------------------------------------------------------------------------
// sgl_comm
// ---------------------------------------------------------------------------------------------
struct sgl_comm {
    sgl_handle* h = nullptr;
    bool own_handle = false;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    cudaStream_t stream = nullptr;
    double* pinned = nullptr;  // 256 doubles (d of up to SGL_MAX_RANK factors is read back through it)
    int64_t collectives = 0;
};

static int comm_finish(sgl_comm* c) {
    c->stream = (cudaStream_t)sgl_stream(c->h);
    if (cudaMallocHost(&c->pinned, sizeof(double) * 256) != cudaSuccess) return fail(SGL_ENOMEM, "cudaMallocHost failed");
    return SGL_OK;
}

extern "C" int sgl_comm_unique_id(void* id128) {
    if (!id128) return fail(SGL_EINVAL, "NULL id buffer");
    NcclApi* api = nullptr;
    SGL_TRY(need_nccl(&api));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    SGL_NCCL(api, api->GetUniqueId(&id));
    std::memcpy(id128, &id, 128);
    return SGL_OK;
}

extern "C" int sgl_comm_init_rank(sgl_handle* h, int device, int world, int rank, const void* id128, sgl_comm** out) {
    if (!h || !out || world < 1 || rank < 0 || rank >= world) return fail(SGL_EINVAL, "sgl_comm_init_rank: bad argument");
    sgl_comm* c = new sgl_comm();
    c->h = h;
    c->rank = rank;
    c->world = world;
    c->device = device;
    SGL_CUDA(cudaSetDevice(device));
    if (world > 1) {
        if (!id128) { delete c; return fail(SGL_EINVAL, "sgl_comm_init_rank: NULL unique id"); }
        NcclApi* api = nullptr;
        int rc = need_nccl(&api);
        if (rc != SGL_OK) { delete c; return rc; }
        ncclUniqueId id;
        std::memcpy(&id, id128, 128);
        ncclResult_t r = api->CommInitRank(&c->comm, world, id, rank);
        if (r != ncclSuccess) { delete c; return fail(SGL_ECUDA, "ncclCommInitRank failed: %s", api->GetErrorString(r)); }
    }
    int rc = comm_finish(c);
    if (rc != SGL_OK) { delete c; return rc; }
    *out = c;
    return SGL_OK;
}

extern "C" int sgl_comm_destroy(sgl_comm* c) {
    if (!c) return SGL_OK;
    cudaSetDevice(c->device);
    if (c->comm) {
        NcclApi* api = nccl_api();
        if (api->lib) api->CommDestroy(c->comm);
    }
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->own_handle) sgl_destroy(c->h);
    delete c;
    return SGL_OK;
}
extern "C" int sgl_comm_rank(const sgl_comm* c) { return c ? c->rank : -1; }
extern "C" int sgl_comm_world(const sgl_comm* c) { return c ? c->world : -1; }
extern "C" sgl_handle* sgl_comm_handle(const sgl_comm* c) { return c ? c->h : nullptr; }
extern "C" int64_t sgl_comm_collectives(const sgl_comm* c) { return c ? c->collectives : 0; }

static int all_reduce_f64(sgl_comm* c, double* buf, size_t n) {
    if (c->world == 1) return SGL_OK;
    NcclApi* api = nccl_api();
    SGL_NCCL(api, api->AllReduce(buf, buf, n, ncclDouble, ncclSum, c->comm, c->stream));
    ++c->collectives;
    return SGL_OK;
}
static int all_reduce_i64(sgl_comm* c, int64_t* buf, size_t n) {
    if (c->world == 1) return SGL_OK;
    NcclApi* api = nccl_api();
    SGL_NCCL(api, api->AllReduce(buf, buf, n, ncclInt64, ncclSum, c->comm, c->stream));
    ++c->collectives;
    return SGL_OK;
}
// full: [per * world][KP] floats; every rank has filled its rows [rank * per, (rank + 1) * per): in-place all-gather
static int all_gather_rows(sgl_comm* c, float* full, int64_t per, int KP) {
    if (c->world == 1) return SGL_OK;
    NcclApi* api = nccl_api();
    const size_t count = (size_t)per * KP;
    SGL_NCCL(api, api->AllGather(full + (size_t)c->rank * count, full, count, ncclFloat, c->comm, c->stream));
    ++c->collectives;
    return SGL_OK;
}
// sum `full` over ranks; rank r ends up with rows [r * per, (r + 1) * per) summed (in place)
static int reduce_scatter_rows(sgl_comm* c, float* full, int64_t per, int KP) {
    if (c->world == 1) return SGL_OK;
    NcclApi* api = nccl_api();
    const size_t count = (size_t)per * KP;
    SGL_NCCL(api, api->ReduceScatter(full, full + (size_t)c->rank * count, count, ncclFloat, ncclSum, c->comm, c->stream));
    ++c->collectives;
    return SGL_OK;
}

// ---------------------------------------------------------------------------------------------
// sgl_fit: one sharded fit on one rank
// ---------------------------------------------------------------------------------------------
struct sgl_fit {
    sgl_comm* c = nullptr;
    const sgl_matrix* A = nullptr;   // m x n_loc (this rank's cells)
    const sgl_matrix* At = nullptr;  // plain: n_loc x m; masked: n x g_loc (this rank's genes over all cells)
    sgl_matrix* At_own = nullptr;    // transpose built here (plain fits without an At)
    sgl_mask *mA = nullptr, *mAt = nullptr;
    bool masked = false;
    int k = 0, KP = 0;
    int64_t m = 0, n = 0;            // global shape
    int64_t c0 = 0, c1 = 0, c_per = 0, g0 = 0, g1 = 0, g_per = 0;
    float *W = nullptr, *Wprev = nullptr, *H = nullptr, *Bw = nullptr;  // W: [g_per * world][KP]; H: plain [n_loc][KP], masked [c_per * world][KP]
    double *gram = nullptr, *dvec = nullptr, *sums = nullptr;
    int64_t* gene_ptr = nullptr;     // plain: int64[m + 1], equal consecutive entries where a gene is empty in the GLOBAL matrix
    int64_t iterations = 0;
    // plain fits: the part of the next iteration that only READS W (Wprev <- W, Gram of W, right-hand sides of the H update) is
    // enqueued before the host waits for this iteration's tol, so the device never idles across the host round trip
    bool lookahead = true, head_ready = false, flag_slot_dirty = false;
    uint64_t head_ticket = 0;
    cudaEvent_t ev_done = nullptr;
};

static void fit_release(sgl_fit* f) {
    if (!f) return;
    cudaSetDevice(f->c->device);
    if (f->mA) sgl_mask_free(f->c->h, f->mA);
    if (f->mAt) sgl_mask_free(f->c->h, f->mAt);
    if (f->At_own) sgl_matrix_free(f->c->h, f->At_own);
    cudaFree(f->W); cudaFree(f->Wprev); cudaFree(f->H); cudaFree(f->Bw);
    cudaFree(f->gram); cudaFree(f->gene_ptr);  // dvec and sums live behind gram
    if (f->ev_done) cudaEventDestroy(f->ev_done);
    delete f;
}

extern "C" int sgl_fit_destroy(sgl_fit* f) {
    fit_release(f);
    return SGL_OK;
}

// masked == 0: A_loc = this rank's cells [c0, c1) of shard_bounds(n_total), At_loc = transpose of that block or NULL.
// masked != 0: At_loc = this rank's genes [g0, g1) of shard_bounds(m) over ALL n_total cells (required).
extern "C" int sgl_fit_create(sgl_comm* c, const sgl_matrix* A_loc, const sgl_matrix* At_loc, int64_t n_total, int k, const double* w_init,
                              int masked, uint64_t seed, uint64_t inv_density, sgl_fit** out) {
    if (!c || !A_loc || !w_init || !out) return fail(SGL_EINVAL, "sgl_fit_create: NULL argument");
    const int KP = sgl_padded_rank(k);
    if (KP < 0) return fail(SGL_EINVAL, "rank k=%d outside [1, %d]", k, SGL_MAX_RANK);
    SGL_CUDA(cudaSetDevice(c->device));
    sgl_fit* f = new sgl_fit();
    f->c = c;
    f->k = k;
    f->KP = KP;
    f->masked = masked != 0;
    int64_t m = 0, n_loc = 0, nnz = 0;
    sgl_matrix_info(A_loc, &m, &n_loc, &nnz);
    f->m = m;
    f->n = n_total;
    sgl_shard_bounds(n_total, c->world, c->rank, &f->c0, &f->c1, &f->c_per);
    sgl_shard_bounds(m, c->world, c->rank, &f->g0, &f->g1, &f->g_per);
    int rc = SGL_OK;
    do {
        if (n_loc != f->c1 - f->c0) {
            rc = fail(SGL_EINVAL, "rank %d holds %lld cells but its shard of %lld cells over %d ranks is [%lld, %lld)", c->rank, (long long)n_loc,
                      (long long)n_total, c->world, (long long)f->c0, (long long)f->c1);
            break;
        }
        f->A = A_loc;
        if (f->masked) {
            if (!At_loc) { rc = fail(SGL_EINVAL, "a masked sharded fit needs this rank's gene block of At over all cells"); break; }
            int64_t tr = 0, tc = 0;
            sgl_matrix_info(At_loc, &tr, &tc, nullptr);
            if (tr != n_total || tc != f->g1 - f->g0) {
                rc = fail(SGL_EINVAL, "rank %d: At block is %lld x %lld, expected %lld x %lld", c->rank, (long long)tr, (long long)tc, (long long)n_total,
                          (long long)(f->g1 - f->g0));
                break;
            }
            f->At = At_loc;
            if ((rc = sgl_mask_build(c->h, f->A, seed, inv_density, 0, f->c0, 0, &f->mA)) != SGL_OK) break;
            if ((rc = sgl_mask_build(c->h, f->At, seed, inv_density, 1, f->g0, 0, &f->mAt)) != SGL_OK) break;
        } else if (At_loc) {
            int64_t tr = 0, tc = 0;
            sgl_matrix_info(At_loc, &tr, &tc, nullptr);
            if (tr != n_loc || tc != m) { rc = fail(SGL_EINVAL, "rank %d: the transpose of the local block must be %lld x %lld", c->rank, (long long)n_loc, (long long)m); break; }
            f->At = At_loc;
        } else {
            if ((rc = sgl_matrix_transpose(c->h, f->A, &f->At_own)) != SGL_OK) break;
            f->At = f->At_own;
        }
        const size_t w_rows = (size_t)f->g_per * c->world, h_rows = f->masked ? (size_t)f->c_per * c->world : (size_t)(n_loc > 0 ? n_loc : 1);
        cudaError_t e = cudaSuccess;
        auto alloc = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes > 0 ? bytes : 16); if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, bytes > 0 ? bytes : 16, c->stream); };
        alloc((void**)&f->W, sizeof(float) * w_rows * KP);
        alloc((void**)&f->Wprev, sizeof(float) * w_rows * KP);
        alloc((void**)&f->H, sizeof(float) * h_rows * KP);
        if (!f->masked) alloc((void**)&f->Bw, sizeof(float) * w_rows * KP);
        // [Gram KP x KP | d KP | stop flag | pad]: one buffer, so that the partial Gram, the row sums and the flag of the
        // H update travel in ONE all-reduce
        alloc((void**)&f->gram, sizeof(double) * (KP * KP + KP + 8));
        if (e == cudaSuccess) f->dvec = f->gram + (size_t)KP * KP;
        if (e == cudaSuccess) f->sums = f->dvec + KP + 1;  // [flag | 5 sums of cor | loss]: one read-back per iteration
        if (!f->masked) alloc((void**)&f->gene_ptr, sizeof(int64_t) * (w_rows + 2));
        if (e != cudaSuccess) { rc = fail(SGL_ENOMEM, "sgl_fit_create: cudaMalloc failed: %s", cudaGetErrorString(e)); break; }
        if ((rc = sgl_factor_upload(c->h, w_init, k, m, f->W)) != SGL_OK) break;  // every rank holds the full w_init
        std::vector<double> ones((size_t)KP, 1.0);
        cudaMemcpyAsync(f->dvec, ones.data(), sizeof(double) * KP, cudaMemcpyHostToDevice, c->stream);
        if (!f->masked) {
            // which genes are empty in the GLOBAL matrix (src/singlet.cpp:340 skips them): column counts of the local
            // transposes, summed over the ranks; gene_ptr = running count of the non-empty genes
            int64_t* cnt = f->gene_ptr + 1;
            if ((rc = sgl_matrix_colptr(c->h, f->At, f->gene_ptr)) != SGL_OK) break;
            std::vector<int64_t> cp((size_t)m + 1);
            cudaMemcpyAsync(cp.data(), f->gene_ptr, sizeof(int64_t) * (m + 1), cudaMemcpyDeviceToHost, c->stream);
            if (cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = fail(SGL_ECUDA, "sgl_fit_create: copy failed"); break; }
            std::vector<int64_t> counts((size_t)m + 1, 0);
            for (int64_t g = 0; g < m; ++g) counts[(size_t)g] = cp[(size_t)g + 1] - cp[(size_t)g];
            cudaMemcpyAsync(cnt, counts.data(), sizeof(int64_t) * m, cudaMemcpyHostToDevice, c->stream);
            if ((rc = all_reduce_i64(c, cnt, (size_t)m)) != SGL_OK) break;
            cudaMemcpyAsync(counts.data(), cnt, sizeof(int64_t) * m, cudaMemcpyDeviceToHost, c->stream);
            if (cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = fail(SGL_ECUDA, "sgl_fit_create: all-reduce of the gene counts failed"); break; }
            std::vector<int64_t> gp(w_rows + 2, 0);
            for (size_t g = 0; g < w_rows + 1; ++g) gp[g + 1] = gp[g] + ((g < (size_t)m && counts[g] > 0) ? 1 : 0);
            cudaMemcpyAsync(f->gene_ptr, gp.data(), sizeof(int64_t) * (w_rows + 2), cudaMemcpyHostToDevice, c->stream);
        }
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = fail(SGL_ECUDA, "sgl_fit_create: %s", cudaGetErrorString(cudaGetLastError())); break; }
    } while (0);
    if (rc != SGL_OK) {
        fit_release(f);
        return rc;
    }
    *out = f;
    return SGL_OK;
}

// The head of a plain iteration: everything that only READS the replicated W -- Wprev <- W, the Gram of W (no collective: W is
// replicated) and the right-hand sides of the H update, left in the handle's scratch under a ticket.
static int plain_head(sgl_fit* f) {
    sgl_comm* c = f->c;
    sgl_handle* h = c->h;
    const size_t w_rows = (size_t)f->g_per * c->world;
    SGL_CUDA(cudaMemcpyAsync(f->Wprev, f->W, sizeof(float) * w_rows * f->KP, cudaMemcpyDeviceToDevice, c->stream));
    SGL_TRY(sgl_dev_gram(h, f->W, f->k, f->m, f->gram, 1));
    SGL_TRY(sgl_dev_update_rhs(h, f->A, f->W, f->k, &f->head_ticket));
    f->head_ready = true;
    return SGL_OK;
}

extern "C" int sgl_fit_set_lookahead(sgl_fit* f, int on) {
    if (!f) return fail(SGL_EINVAL, "NULL fit");
    f->lookahead = on != 0;
    return SGL_OK;
}

// One ALS iteration. stop_flag (optional, in/out): every rank passes 0 or 1; the values are summed in an all-reduce the
// iteration makes anyway, so all ranks leave with the same non-zero value when ANY rank asked to stop (interrupts in a
// one-process job).
// Collectives of a plain iteration on more than one rank: ONE all-reduce of [partial Gram of the unscaled H | row sums of H |
// stop flag] (the Gram is rescaled by 1 / (d_i d_j) afterwards, sgl_dev_finish_d_rescale_gram), the reduce-scatter of the
// W-update right-hand sides, and ONE grouped launch of {all-reduce of the W row sums, all-gather of the unscaled W shard};
// every rank then scales the whole W by the same d. SGL_FIT_COLLECTIVES=split restores the five separate collectives of
// the first version (A/B runs).
extern "C" int sgl_fit_iterate(sgl_fit* f, double L1_w, double L1_h, double L2_w, double L2_h, double* tol_out, int* stop_flag) {
    if (!f) return fail(SGL_EINVAL, "NULL fit");
    sgl_comm* c = f->c;
    sgl_handle* h = c->h;
    const int k = f->k, KP = f->KP;
    const int64_t m = f->m, n_loc = f->c1 - f->c0, g_loc = f->g1 - f->g0;
    SGL_CUDA(cudaSetDevice(c->device));
    const size_t w_rows = (size_t)f->g_per * c->world;
    float* H_loc = f->masked ? f->H + (size_t)f->c0 * KP : f->H;
    float* W_loc = f->W + (size_t)f->g0 * KP;
    static const bool split_env = [] { const char* e = getenv("SGL_FIT_COLLECTIVES"); return e && strcmp(e, "split") == 0; }();
    const bool merged = !f->masked && c->world > 1 && !split_env;
    const double flag = (stop_flag && *stop_flag) ? 1.0 : 0.0;
    c->pinned[40] = flag;
    if (!f->masked) {
        if (!(f->head_ready && f->head_ticket == sgl_dev_rhs_epoch(h))) SGL_TRY(plain_head(f));
        f->head_ready = false;
        // ---- H update over the local cells against the replicated W ----
        SGL_TRY(sgl_dev_update_solve(h, f->A, f->head_ticket, H_loc, k, f->gram, L1_h, L2_h, f->dvec));
        if (merged) {
            SGL_TRY(sgl_dev_gram(h, H_loc, k, n_loc, f->gram, 0));  // of the UNSCALED H
            if (flag != 0.0 || f->flag_slot_dirty)                  // the slot holds 0 otherwise
                SGL_CUDA(cudaMemcpyAsync(f->dvec + KP, c->pinned + 40, sizeof(double), cudaMemcpyHostToDevice, c->stream));
            SGL_TRY(all_reduce_f64(c, f->gram, (size_t)KP * KP + KP + 1));
            SGL_TRY(sgl_dev_finish_d_rescale_gram(h, k, f->dvec, f->gram));
            SGL_TRY(sgl_dev_scale(h, H_loc, k, n_loc, f->dvec));
        } else {
            SGL_TRY(all_reduce_f64(c, f->dvec, (size_t)KP));
            SGL_TRY(sgl_dev_finish_d(h, k, f->dvec));
            SGL_TRY(sgl_dev_scale(h, H_loc, k, n_loc, f->dvec));
            SGL_TRY(sgl_dev_gram(h, H_loc, k, n_loc, f->gram, 0));
            SGL_TRY(all_reduce_f64(c, f->gram, (size_t)KP * KP));
            SGL_TRY(sgl_dev_gram_jitter(h, k, f->gram));
        }
        // ---- W update: partial right-hand sides of ALL genes from the local cells, summed onto the gene shards ----
        SGL_TRY(sgl_dev_rhs(h, f->At, H_loc, k, f->Bw));
        SGL_TRY(reduce_scatter_rows(c, f->Bw, f->g_per, KP));
        SGL_TRY(sgl_dev_solve(h, f->Bw + (size_t)f->g0 * KP, f->gene_ptr + f->g0, g_loc, W_loc, k, f->gram, L1_w, L2_w, f->dvec));
    } else {
        SGL_CUDA(cudaMemcpyAsync(f->Wprev, f->W, sizeof(float) * w_rows * KP, cudaMemcpyDeviceToDevice, c->stream));
        // ---- H update: Gram of W summed over the ranks' gene shards; masked solve of the local cells; H all-gathered ----
        SGL_TRY(sgl_dev_gram(h, W_loc, k, g_loc, f->gram, 0));
        SGL_TRY(all_reduce_f64(c, f->gram, (size_t)KP * KP));
        SGL_TRY(sgl_dev_gram_jitter(h, k, f->gram));
        SGL_TRY(sgl_dev_update_masked(h, f->A, f->mA, f->W, H_loc, k, f->gram, L1_h, L2_h, f->dvec));
        SGL_TRY(all_reduce_f64(c, f->dvec, (size_t)KP));
        SGL_TRY(sgl_dev_finish_d(h, k, f->dvec));
        SGL_TRY(sgl_dev_scale(h, H_loc, k, n_loc, f->dvec));
        SGL_TRY(all_gather_rows(c, f->H, f->c_per, KP));
        // ---- W update over the local genes against the replicated H ----
        SGL_TRY(sgl_dev_gram(h, H_loc, k, n_loc, f->gram, 0));
        SGL_TRY(all_reduce_f64(c, f->gram, (size_t)KP * KP));
        SGL_TRY(sgl_dev_gram_jitter(h, k, f->gram));
        SGL_TRY(sgl_dev_update_masked(h, f->At, f->mAt, f->H, W_loc, k, f->gram, L1_w, L2_w, f->dvec));
    }
    if (merged) {
        // row sums of the W update and the unscaled shards in one NCCL launch; every rank scales the whole W by the same d
        NcclApi* api = nccl_api();
        SGL_NCCL(api, api->GroupStart());
        int rc_a = all_reduce_f64(c, f->dvec, (size_t)KP);
        int rc_b = rc_a == SGL_OK ? all_gather_rows(c, f->W, f->g_per, KP) : rc_a;
        SGL_NCCL(api, api->GroupEnd());
        if (rc_b != SGL_OK) return rc_b;
        SGL_TRY(sgl_dev_finish_d(h, k, f->dvec));
        SGL_TRY(sgl_dev_scale(h, f->W, k, m, f->dvec));
    } else {
        // row sums of the W update (+ the stop flag of every rank in the spare slot behind them)
        SGL_CUDA(cudaMemcpyAsync(f->dvec + KP, c->pinned + 40, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        SGL_TRY(all_reduce_f64(c, f->dvec, (size_t)KP + 1));
        SGL_TRY(sgl_dev_finish_d(h, k, f->dvec));
        SGL_TRY(sgl_dev_scale(h, W_loc, k, g_loc, f->dvec));
        SGL_TRY(all_gather_rows(c, f->W, f->g_per, KP));
    }
    SGL_TRY(sgl_dev_cor_sums(h, f->W, f->Wprev, k, m, f->sums));  // replicated W: the same value on every rank
    SGL_CUDA(cudaMemcpyAsync(c->pinned + 41, f->dvec + KP, sizeof(double) * 6, cudaMemcpyDeviceToHost, c->stream));  // flag, 5 sums
    if (!f->masked && f->lookahead) {
        // the host waits for this iteration's numbers only; the head of the next one is already queued behind them
        if (!f->ev_done) SGL_CUDA(cudaEventCreateWithFlags(&f->ev_done, cudaEventDisableTiming));
        SGL_CUDA(cudaEventRecord(f->ev_done, c->stream));
        SGL_TRY(plain_head(f));
        SGL_CUDA(cudaEventSynchronize(f->ev_done));
    } else {
        SGL_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (tol_out) *tol_out = sgl_cor_from_sums(c->pinned + 42, (double)k * (double)m);
    const bool stopped = c->pinned[41] > 0.5;
    f->flag_slot_dirty = stopped;
    if (stop_flag) *stop_flag = stopped ? 1 : 0;
    ++f->iterations;
    return SGL_OK;
}

// mse_test (src/singlet.cpp:536-568) of a masked fit: mean over ALL cells of the per-cell held-out loss
extern "C" int sgl_fit_test_mse(sgl_fit* f, double* out) {
    if (!f || !out) return fail(SGL_EINVAL, "NULL argument");
    if (!f->masked) return fail(SGL_EINVAL, "test MSE needs a masked fit");
    sgl_comm* c = f->c;
    SGL_CUDA(cudaSetDevice(c->device));
    SGL_TRY(sgl_dev_mse(c->h, f->A, f->mA, f->W, f->dvec, f->H + (size_t)f->c0 * f->KP, f->k, 0, f->sums + 5));
    SGL_TRY(all_reduce_f64(c, f->sums + 5, 1));
    SGL_CUDA(cudaMemcpyAsync(c->pinned + 8, f->sums + 5, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SGL_CUDA(cudaStreamSynchronize(c->stream));
    *out = c->pinned[8] / (double)f->n;
    return SGL_OK;
}

// w: k x m (every rank gets the same), d: k, h_local: k x (this rank's cells). Any of them may be NULL.
extern "C" int sgl_fit_download(sgl_fit* f, double* w, double* d, double* h_local) {
    if (!f) return fail(SGL_EINVAL, "NULL fit");
    sgl_comm* c = f->c;
    SGL_CUDA(cudaSetDevice(c->device));
    if (w) SGL_TRY(sgl_factor_download(c->h, f->W, f->k, f->m, w));
    if (h_local && f->c1 > f->c0) SGL_TRY(sgl_factor_download(c->h, f->masked ? f->H + (size_t)f->c0 * f->KP : f->H, f->k, f->c1 - f->c0, h_local));
    if (d) {
        SGL_CUDA(cudaMemcpyAsync(c->pinned, f->dvec, sizeof(double) * f->k, cudaMemcpyDeviceToHost, c->stream));
        SGL_CUDA(cudaStreamSynchronize(c->stream));
        for (int q = 0; q < f->k; ++q) d[q] = c->pinned[q];
    }
    return SGL_OK;
}
extern "C" int sgl_fit_shard(const sgl_fit* f, int64_t* c0, int64_t* c1, int64_t* g0, int64_t* g1) {
    if (!f) return fail(SGL_EINVAL, "NULL fit");
    if (c0) *c0 = f->c0;
    if (c1) *c1 = f->c1;
    if (g0) *g0 = f->g0;
    if (g1) *g1 = f->g1;
    return SGL_OK;
}

// ---- whole fits on one rank: c_nmf_base / c_ard_nmf_base (src/singlet.cpp:638-666, 1091-1152) over shards ----
struct IterEvent { int iter; double tol, overfit; };
struct RankShared {  // one-process jobs: rank 0 reports its iterations to the calling thread, which owns the callbacks
    std::mutex mu;
    std::vector<IterEvent> events;
    std::atomic<int> stop{0};
};

static int nmf_rank(sgl_comm* c, const sgl_matrix* A_loc, const sgl_matrix* At_loc, int64_t n_total, double tol, uint16_t maxit, double L1_w,
                    double L1_h, double L2_w, double L2_h, int k, double* w, double* d, double* h_local, int32_t* iters_out, double* tol_out,
                    const sgl_callbacks* cb, RankShared* shared) {
    sgl_fit* f = nullptr;
    // SGL_TIMING=1: wall-clock phases of this rank on stderr (developer aid, like the single-GPU entry points)
    const bool timing = getenv("SGL_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!timing) return;
        cudaStreamSynchronize(c->stream);
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[sgl timing] rank %d %-34s %9.3f ms\n", c->rank, what, std::chrono::duration<double, std::milli>(t1 - t_last).count());
        t_last = t1;
    };
    SGL_TRY(sgl_fit_create(c, A_loc, At_loc, n_total, k, w, 0, 0, 0, &f));
    mark("fit_create (transpose, gene counts)");
    double tol_ = 1;
    uint16_t iter_ = 0;
    int rc = SGL_OK;
    for (; iter_ < maxit && tol_ > tol; ++iter_) {  // src/singlet.cpp:647
        if (iter_ == 1) mark("first iteration (+ tiles)");
        int stop = 0;
        if (cb && cb->poll_interrupt && cb->poll_interrupt(cb->user)) stop = 1;
        if (shared && shared->stop.load()) stop = 1;
        f->lookahead = (int)iter_ + 1 < (int)maxit;  // no product for an iteration that will not run
        if ((rc = sgl_fit_iterate(f, L1_w, L1_h, L2_w, L2_h, &tol_, &stop)) != SGL_OK) break;
        if (cb && cb->on_iter) cb->on_iter(cb->user, iter_ + 1, tol_, NAN);
        if (shared && c->rank == 0) {
            std::lock_guard<std::mutex> lock(shared->mu);
            shared->events.push_back({iter_ + 1, tol_, NAN});
        }
        if (stop) { rc = fail(SGL_EINTERRUPT, "interrupted"); break; }
    }
    mark("other iterations");
    if (rc == SGL_OK) {
        if (iters_out) *iters_out = iter_;
        if (tol_out) *tol_out = tol_;
        rc = sgl_fit_download(f, w, d, h_local);
    }
    mark("download");
    fit_release(f);
    return rc;
}

static int ard_rank(sgl_comm* c, const sgl_matrix* A_loc, const sgl_matrix* At_loc, int64_t n_total, double tol, uint16_t maxit, double L1, double L2,
                    int k, double* w, double* d, double* h_local, uint64_t seed, uint64_t inv_density, double overfit_threshold,
                    uint16_t trace_test_mse, sgl_trace* tr, const sgl_callbacks* cb, RankShared* shared) {
    if (trace_test_mse == 0) return fail(SGL_EINVAL, "trace_test_mse must be >= 1 (the reference divides by it)");
    if (inv_density == 0) return fail(SGL_EINVAL, "inv_density must be >= 1");
    if (!tr || tr->capacity < (int)maxit / (int)trace_test_mse + 2) return fail(SGL_EINVAL, "trace capacity too small");
    sgl_fit* f = nullptr;
    SGL_TRY(sgl_fit_create(c, A_loc, At_loc, n_total, k, w, 1, seed, inv_density, &f));
    tr->length = 0;
    auto push = [&](double mse, int it, double ft) {
        const int q = tr->length;
        tr->test_mse[q] = mse;
        tr->iter[q] = it;
        tr->tol[q] = ft;
        double mn = tr->test_mse[0];
        for (int t = 1; t <= q; ++t) mn = tr->test_mse[t] < mn ? tr->test_mse[t] : mn;
        tr->score_overfit[q] = (mse - mn) / (mse + mn);
        tr->length = q + 1;
    };
    double tol_ = 1;
    uint16_t iter_ = 0;
    int rc = SGL_OK;
    for (; iter_ < maxit && tol_ > tol; ++iter_) {  // src/singlet.cpp:1107
        int stop = 0;
        if (cb && cb->poll_interrupt && cb->poll_interrupt(cb->user)) stop = 1;
        if (shared && shared->stop.load()) stop = 1;
        if ((rc = sgl_fit_iterate(f, L1, L1, L2, L2, &tol_, &stop)) != SGL_OK) break;
        double overfit = NAN;
        bool done = false;
        if (iter_ % trace_test_mse == 0) {
            double mse = 0;
            if ((rc = sgl_fit_test_mse(f, &mse)) != SGL_OK) break;
            push(mse, iter_, tol_);
            overfit = tr->score_overfit[tr->length - 1];
            done = overfit > overfit_threshold;  // the same value on every rank (all-reduced loss)
        }
        if (cb && cb->on_iter) cb->on_iter(cb->user, iter_ + 1, tol_, overfit);
        if (shared && c->rank == 0) {
            std::lock_guard<std::mutex> lock(shared->mu);
            shared->events.push_back({iter_ + 1, tol_, overfit});
        }
        if (stop) { rc = fail(SGL_EINTERRUPT, "interrupted"); break; }
        if (done) break;  // iter_ is not incremented on break
    }
    if (rc == SGL_OK && iter_ % trace_test_mse != 0) {
        double mse = 0;
        rc = sgl_fit_test_mse(f, &mse);
        if (rc == SGL_OK) push(mse, iter_, tol_);
    }
    if (rc == SGL_OK) rc = sgl_fit_download(f, w, d, h_local);
    fit_release(f);
    return rc;
}

extern "C" int sgl_nmf_rank(sgl_comm* c, const sgl_matrix* A_loc, const sgl_matrix* At_loc, int64_t n_total, double tol, uint16_t maxit,
                            double L1_w, double L1_h, double L2_w, double L2_h, int k, double* w, double* d, double* h_local,
                            int32_t* iters_out, double* tol_out, const sgl_callbacks* cb) {
    if (!c || !A_loc || !w) return fail(SGL_EINVAL, "sgl_nmf_rank: NULL argument");
    return nmf_rank(c, A_loc, At_loc, n_total, tol, maxit, L1_w, L1_h, L2_w, L2_h, k, w, d, h_local, iters_out, tol_out, cb, nullptr);
}
extern "C" int sgl_ard_nmf_rank(sgl_comm* c, const sgl_matrix* A_loc, const sgl_matrix* At_loc, int64_t n_total, double tol, uint16_t maxit, double L1,
                                double L2, int k, double* w, double* d, double* h_local, uint64_t seed, uint64_t inv_density,
                                double overfit_threshold, uint16_t trace_test_mse, sgl_trace* trace, const sgl_callbacks* cb) {
    if (!c || !A_loc || !At_loc || !w || !trace) return fail(SGL_EINVAL, "sgl_ard_nmf_rank: NULL argument");
    return ard_rank(c, A_loc, At_loc, n_total, tol, maxit, L1, L2, k, w, d, h_local, seed, inv_density, overfit_threshold, trace_test_mse, trace,
                    cb, nullptr);
}

// ---------------------------------------------------------------------------------------------
// sgl_multi: all devices of one process
// ---------------------------------------------------------------------------------------------
struct sgl_multi {
    std::vector<sgl_comm*> ranks;
};

extern "C" int sgl_multi_create(int n_devices, const int* devices, sgl_multi** out) {
    if (!out || n_devices < 1) return fail(SGL_EINVAL, "sgl_multi_create: bad argument");
    const int have = sgl_device_count();
    if (have <= 0) return fail(SGL_ENODEVICE, "no CUDA device available (this library has no CPU fallback)");
    if (n_devices > have) return fail(SGL_EINVAL, "%d devices requested, %d present", n_devices, have);
    std::vector<int> devs((size_t)n_devices);
    for (int q = 0; q < n_devices; ++q) devs[(size_t)q] = devices ? devices[q] : q;
    sgl_multi* mg = new sgl_multi();
    std::vector<ncclComm_t> comms((size_t)n_devices, nullptr);
    if (n_devices > 1) {
        NcclApi* api = nullptr;
        int rc = need_nccl(&api);
        if (rc != SGL_OK) { delete mg; return rc; }
        ncclResult_t r = api->CommInitAll(comms.data(), n_devices, devs.data());
        if (r != ncclSuccess) { delete mg; return fail(SGL_ECUDA, "ncclCommInitAll failed: %s", api->GetErrorString(r)); }
    }
    for (int q = 0; q < n_devices; ++q) {
        sgl_comm* c = new sgl_comm();
        c->rank = q;
        c->world = n_devices;
        c->device = devs[(size_t)q];
        c->comm = comms[(size_t)q];
        c->own_handle = true;
        int rc = sgl_create(c->device, nullptr, &c->h);
        if (rc == SGL_OK) rc = comm_finish(c);
        mg->ranks.push_back(c);
        if (rc != SGL_OK) {
            for (sgl_comm* r : mg->ranks) sgl_comm_destroy(r);
            delete mg;
            return rc;
        }
    }
    *out = mg;
    return SGL_OK;
}
extern "C" int sgl_multi_destroy(sgl_multi* mg) {
    if (!mg) return SGL_OK;
    for (sgl_comm* r : mg->ranks) sgl_comm_destroy(r);
    delete mg;
    return SGL_OK;
}
extern "C" int sgl_multi_size(const sgl_multi* mg) { return mg ? (int)mg->ranks.size() : 0; }
extern "C" sgl_comm* sgl_multi_rank(const sgl_multi* mg, int rank) {
    return (mg && rank >= 0 && rank < (int)mg->ranks.size()) ? mg->ranks[(size_t)rank] : nullptr;
}
extern "C" int sgl_multi_set_precision(sgl_multi* mg, int mode) {
    if (!mg) return fail(SGL_EINVAL, "NULL argument");
    for (sgl_comm* r : mg->ranks) SGL_TRY(sgl_set_precision(r->h, mode));
    return SGL_OK;
}

// the columns [lo, hi) of a chunk list as a list of zero-copy column-range views
static std::vector<sgl_csc> slice_columns(const sgl_csc* chunks, int n_chunks, int64_t lo, int64_t hi) {
    std::vector<sgl_csc> out;
    int64_t off = 0;
    for (int q = 0; q < n_chunks; ++q) {
        const int64_t b = off, e = off + chunks[q].ncol;
        off = e;
        const int64_t s = lo > b ? lo : b, t = hi < e ? hi : e;
        if (s >= t) continue;
        sgl_csc v = chunks[q];
        v.p = chunks[q].p + (s - b);
        v.ncol = t - s;
        out.push_back(v);
    }
    if (out.empty() && n_chunks > 0) {  // an empty shard: zero columns of the right height
        sgl_csc v = chunks[0];
        v.ncol = 0;
        out.push_back(v);
    }
    return out;
}

struct MultiJob {
    int rc = SGL_OK;
    std::string err;
};

template <typename RankFn>
static int run_ranks(sgl_multi* mg, const sgl_callbacks* cb, RankShared& shared, RankFn fn) {
    const int G = (int)mg->ranks.size();
    std::vector<MultiJob> jobs((size_t)G);
    std::atomic<int> running(G);
    std::vector<std::thread> pool;
    for (int r = 0; r < G; ++r)
        pool.emplace_back([&, r] {
            jobs[(size_t)r].rc = fn(r);
            if (jobs[(size_t)r].rc != SGL_OK) {
                jobs[(size_t)r].err = last_error();  // thread-local: hand it to the calling thread
                shared.stop.store(1);                  // let the other ranks leave their loops at the next iteration
            }
            running.fetch_sub(1);
        });
    size_t seen = 0;
    auto drain = [&] {  // callbacks belong to the calling thread
        std::vector<IterEvent> ev;
        {
            std::lock_guard<std::mutex> lock(shared.mu);
            ev.assign(shared.events.begin() + (long)seen, shared.events.end());
            seen = shared.events.size();
        }
        if (cb && cb->on_iter)
            for (const IterEvent& e : ev) cb->on_iter(cb->user, e.iter, e.tol, e.overfit);
    };
    while (running.load() > 0) {
        if (cb && cb->poll_interrupt && !shared.stop.load() && cb->poll_interrupt(cb->user)) shared.stop.store(1);
        drain();
        std::this_thread::sleep_for(std::chrono::microseconds(500));
    }
    for (auto& th : pool) th.join();
    drain();
    for (int r = 0; r < G; ++r)
        if (jobs[(size_t)r].rc != SGL_OK && jobs[(size_t)r].rc != SGL_EINTERRUPT) return fail(jobs[(size_t)r].rc, "rank %d: %s", r, jobs[(size_t)r].err.c_str());
    for (int r = 0; r < G; ++r)
        if (jobs[(size_t)r].rc == SGL_EINTERRUPT) return fail(SGL_EINTERRUPT, "interrupted");
    return SGL_OK;
}

// c_nmf_sparse_list (src/singlet.cpp:715-743) over the devices of this process. A_ = column chunks of A (any number, any
// sizes; they are re-cut into one contiguous cell shard per device without copying). At_ is not needed (every device
// transposes its own cell block); it is accepted for signature parity and ignored. w: k x m in/out, h_out: k x n.
extern "C" int sgl_multi_nmf(sgl_multi* mg, const sgl_csc* A_, int nA, const sgl_csc* At_, int nAt, double tol, uint16_t maxit, double L1_w,
                             double L1_h, double L2_w, double L2_h, int k, double* w, double* d, double* h_out, int32_t* iters_out,
                             double* tol_out, const sgl_callbacks* cb) {
    (void)At_;
    (void)nAt;
    if (!mg || !A_ || nA < 1 || !w || !d || !h_out) return fail(SGL_EINVAL, "sgl_multi_nmf: NULL argument");
    if (k < 1 || k > SGL_MAX_RANK) return fail(SGL_EINVAL, "rank k=%d outside [1, %d]", k, SGL_MAX_RANK);
    const int G = (int)mg->ranks.size();
    int64_t n = 0;
    const int64_t m = A_[0].nrow;
    for (int q = 0; q < nA; ++q) n += A_[q].ncol;
    RankShared shared;
    std::vector<std::vector<double>> w_rank((size_t)G, std::vector<double>(w, w + (size_t)k * (size_t)m));
    std::vector<int32_t> iters((size_t)G, 0);
    std::vector<double> tols((size_t)G, 1.0), d_rank((size_t)G * (size_t)k, 0.0);
    const int rc = run_ranks(mg, cb, shared, [&](int r) -> int {
        sgl_comm* c = mg->ranks[(size_t)r];
        if (cudaSetDevice(c->device) != cudaSuccess) return fail(SGL_ECUDA, "cudaSetDevice failed");
        int64_t c0, c1;
        sgl_shard_bounds(n, G, r, &c0, &c1, nullptr);
        std::vector<sgl_csc> view = slice_columns(A_, nA, c0, c1);
        sgl_matrix* A_loc = nullptr;
        int rr = sgl_matrix_upload(c->h, view.data(), (int)view.size(), &A_loc);
        if (rr == SGL_OK)
            rr = nmf_rank(c, A_loc, nullptr, n, tol, maxit, L1_w, L1_h, L2_w, L2_h, k, w_rank[(size_t)r].data(), d_rank.data() + (size_t)r * k,
                          h_out + (size_t)c0 * k, &iters[(size_t)r], &tols[(size_t)r], nullptr, &shared);
        else
            shared.stop.store(1);
        if (A_loc) sgl_matrix_free(c->h, A_loc);
        return rr;
    });
    if (rc != SGL_OK) return rc;
    std::memcpy(w, w_rank[0].data(), sizeof(double) * (size_t)k * (size_t)m);
    std::memcpy(d, d_rank.data(), sizeof(double) * (size_t)k);
    if (iters_out) *iters_out = iters[0];
    if (tol_out) *tol_out = tols[0];
    return SGL_OK;
}

// c_ard_nmf_sparse_list (src/singlet.cpp:1162-1234) over the devices of this process: A_ = column chunks of A, At_ = the
// reference's "distributed transpose" (R/cross_validate_nmf.R:37-50): column chunks of t(A), i.e. gene blocks over all cells.
// Both lists are re-cut into one contiguous shard per device without copying.
extern "C" int sgl_multi_ard_nmf(sgl_multi* mg, const sgl_csc* A_, int nA, const sgl_csc* At_, int nAt, double tol, uint16_t maxit, double L1,
                                 double L2, int k, double* w, double* d, double* h_out, uint64_t seed, uint64_t inv_density,
                                 double overfit_threshold, uint16_t trace_test_mse, sgl_trace* trace, const sgl_callbacks* cb) {
    if (!mg || !A_ || nA < 1 || !At_ || nAt < 1 || !w || !d || !h_out || !trace)
        return fail(SGL_EINVAL, "sgl_multi_ard_nmf: NULL argument (the masked multi-GPU fit needs the gene-block list At_)");
    if (k < 1 || k > SGL_MAX_RANK) return fail(SGL_EINVAL, "rank k=%d outside [1, %d]", k, SGL_MAX_RANK);
    const int G = (int)mg->ranks.size();
    int64_t n = 0, mt = 0;
    const int64_t m = A_[0].nrow;
    for (int q = 0; q < nA; ++q) n += A_[q].ncol;
    for (int q = 0; q < nAt; ++q) mt += At_[q].ncol;
    if (mt != m || At_[0].nrow != n) return fail(SGL_EINVAL, "At_ (%lld x %lld) is not the transpose shape of A_ (%lld x %lld)", (long long)At_[0].nrow, (long long)mt, (long long)m, (long long)n);
    RankShared shared;
    std::vector<std::vector<double>> w_rank((size_t)G, std::vector<double>(w, w + (size_t)k * (size_t)m));
    std::vector<double> d_rank((size_t)G * (size_t)k, 0.0);
    // every rank fills an identical trace (all-reduced losses); rank 0 writes the caller's, the others private copies
    const int cap = trace->capacity;
    std::vector<std::vector<double>> t_mse((size_t)G, std::vector<double>((size_t)cap)), t_tol = t_mse, t_so = t_mse;
    std::vector<std::vector<int32_t>> t_it((size_t)G, std::vector<int32_t>((size_t)cap));
    std::vector<sgl_trace> traces((size_t)G);
    for (int r = 0; r < G; ++r) {
        traces[(size_t)r] = sgl_trace{t_mse[(size_t)r].data(), t_it[(size_t)r].data(), t_tol[(size_t)r].data(), t_so[(size_t)r].data(), cap, 0};
        if (r == 0) traces[0] = *trace;
    }
    const int rc = run_ranks(mg, cb, shared, [&](int r) -> int {
        sgl_comm* c = mg->ranks[(size_t)r];
        if (cudaSetDevice(c->device) != cudaSuccess) return fail(SGL_ECUDA, "cudaSetDevice failed");
        int64_t c0, c1, g0, g1;
        sgl_shard_bounds(n, G, r, &c0, &c1, nullptr);
        sgl_shard_bounds(m, G, r, &g0, &g1, nullptr);
        std::vector<sgl_csc> va = slice_columns(A_, nA, c0, c1), vt = slice_columns(At_, nAt, g0, g1);
        sgl_matrix *A_loc = nullptr, *At_loc = nullptr;
        int rr = sgl_matrix_upload(c->h, va.data(), (int)va.size(), &A_loc);
        if (rr == SGL_OK) rr = sgl_matrix_upload(c->h, vt.data(), (int)vt.size(), &At_loc);
        if (rr == SGL_OK)
            rr = ard_rank(c, A_loc, At_loc, n, tol, maxit, L1, L2, k, w_rank[(size_t)r].data(), d_rank.data() + (size_t)r * k, h_out + (size_t)c0 * k,
                          seed, inv_density, overfit_threshold, trace_test_mse, &traces[(size_t)r], nullptr, &shared);
        else
            shared.stop.store(1);
        if (A_loc) sgl_matrix_free(c->h, A_loc);
        if (At_loc) sgl_matrix_free(c->h, At_loc);
        return rr;
    });
    if (rc != SGL_OK) return rc;
    trace->length = traces[0].length;
    std::memcpy(w, w_rank[0].data(), sizeof(double) * (size_t)k * (size_t)m);
    std::memcpy(d, d_rank.data(), sizeof(double) * (size_t)k);
    return SGL_OK;
}
