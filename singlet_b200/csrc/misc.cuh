// misc.cuh -- the small dense kernels of the ALS loop (Gram, scale, cor), format conversion, the
// speckled-mask materialisation, the test/train MSE and the synthetic-data generator.
#pragma once
#include "common.cuh"

namespace sgl {

// ----------------------------------------------------------------------------------------------
// Gram  a = X X^T  (AAt, reference src/singlet.cpp:200-206) for float X [cols][KP], accumulated in
// FP64 across tiles (FP32 within a tile of <= 128 columns). Each CTA reduces a strided set of tiles staged in shared memory; every thread
// owns a TM x TM micro-tile of the KP x KP output. Per-CTA partials are reduced in fixed order by
// reduce_partials_kernel (deterministic), which is also where multi-GPU partials would be added.
// ----------------------------------------------------------------------------------------------
template <int KP>
struct GramCfg {
    static constexpr int TM = (KP >= 32) ? KP / 16 : 1;  // micro-tile edge: 32->2, 64->4, 128->8
    static constexpr int TPB = (KP / TM) * (KP / TM);    // threads with work (<= 256)
    static constexpr int THREADS = 256;
    static constexpr int TILE_COLS = (KP <= 32) ? 128 : ((KP == 64) ? 64 : 32);
};

template <int KP>
__global__ void __launch_bounds__(256)
gram_partial_kernel(const float* __restrict__ X, int64_t cols, double* __restrict__ part /*[grid][KP*KP]*/) {
    using C = GramCfg<KP>;
    constexpr int TM = C::TM, TC = C::TILE_COLS, LD = KP + 4;  // padded row: conflict-light float reads
    __shared__ __align__(16) float tile[TC * LD];
    const int tid = threadIdx.x;
    const int ti = (tid / (KP / TM)) * TM, tj = (tid % (KP / TM)) * TM;
    const bool worker = tid < C::TPB;
    double acc[TM][TM];
#pragma unroll
    for (int a = 0; a < TM; ++a)
#pragma unroll
        for (int b = 0; b < TM; ++b) acc[a][b] = 0.0;

    for (int64_t c0 = (int64_t)blockIdx.x * TC; c0 < cols; c0 += (int64_t)gridDim.x * TC) {
        const int nc = (int)min((int64_t)TC, cols - c0);
        for (int e = tid; e < TC * (KP / 4); e += C::THREADS) {
            const int c = e / (KP / 4), f4 = e % (KP / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < nc) v = reinterpret_cast<const float4*>(X + (c0 + c) * KP)[f4];
            *reinterpret_cast<float4*>(&tile[c * LD + 4 * f4]) = v;
        }
        __syncthreads();
        if (worker) {
            // FP32 products summed over the <= TILE_COLS columns of this tile, then folded into the FP64
            // accumulators: the long sum (up to 10^6 columns) is carried in double, the inner loop stays
            // on the FP32 pipe (the FP64 pipe made this kernel 6x slower)
            float part32[TM][TM];
#pragma unroll
            for (int a = 0; a < TM; ++a)
#pragma unroll
                for (int b = 0; b < TM; ++b) part32[a][b] = 0.f;
            for (int c = 0; c < nc; ++c) {
                float xi[TM], xj[TM];
#pragma unroll
                for (int a = 0; a < TM; ++a) { xi[a] = tile[c * LD + ti + a]; xj[a] = tile[c * LD + tj + a]; }
#pragma unroll
                for (int a = 0; a < TM; ++a)
#pragma unroll
                    for (int b = 0; b < TM; ++b) part32[a][b] = fmaf(xi[a], xj[b], part32[a][b]);
            }
#pragma unroll
            for (int a = 0; a < TM; ++a)
#pragma unroll
                for (int b = 0; b < TM; ++b) acc[a][b] += (double)part32[a][b];
        }
        __syncthreads();
    }
    if (worker) {
#pragma unroll
        for (int a = 0; a < TM; ++a)
#pragma unroll
            for (int b = 0; b < TM; ++b) part[(int64_t)blockIdx.x * KP * KP + (ti + a) * KP + (tj + b)] = acc[a][b];
    }
}

// gram (double, with jitter) -> float copy without the jitter on padding, reciprocal diagonal
__global__ void gram_finish_kernel(const double* __restrict__ gram, int k, int KP, float* __restrict__ gram_f,
                                   float* __restrict__ gram_f_nojit, float* __restrict__ inv_diag,
                                   unsigned long long* __restrict__ zero4 /* work counters of the solver that follows, or NULL */) {
    if (zero4 && threadIdx.x < 4) zero4[threadIdx.x] = 0ull;
    for (int t = threadIdx.x; t < KP * KP; t += blockDim.x) {
        const int r = t / KP, c = t % KP;
        const double g = gram[t];
        gram_f[t] = (float)g;
        gram_f_nojit[t] = (float)((r == c && r < k) ? g - 1e-15 : g);
        if (r == c) inv_diag[r] = (r < k) ? (float)(1.0 / g) : 0.f;
    }
}
__global__ void add_jitter_kernel(double* gram, int k, int KP) {
    const int r = threadIdx.x;
    if (r < k) gram[r * KP + r] += 1e-15;
}
__global__ void finish_d_kernel(double* d, int k, int KP) {
    const int r = threadIdx.x;
    if (r < KP) d[r] = (r < k) ? d[r] + 1e-15 : 1.0;
}
// finish_d, then the Gram of the UNSCALED factor becomes the Gram of the scaled one, G[i][j] / (d[i] d[j]), plus the
// jitter (src/singlet.cpp:206): lets the sharded driver all-reduce the row sums and the partial Grams in one collective
__global__ void finish_d_rescale_gram_kernel(double* d, double* gram, int k, int KP) {
    __shared__ double sd[128];
    for (int r = threadIdx.x; r < KP; r += blockDim.x) {
        const double v = (r < k) ? d[r] + 1e-15 : 1.0;
        d[r] = v;
        sd[r] = v;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < KP * KP; t += blockDim.x) {
        const int r = t / KP, c = t % KP;
        if (r < k && c < k) {
            const double g = gram[t] / (sd[r] * sd[c]);
            gram[t] = (r == c) ? g + 1e-15 : g;
        }
    }
}

// scale (src/singlet.cpp:222-224): X[c][f] /= d[f]
__global__ void scale_kernel(float* __restrict__ X, int64_t n_elems, int KP, const double* __restrict__ d) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_elems) {
        const int f = (int)(e & (KP - 1));
        X[e] = (float)((double)X[e] / d[f]);
    }
}

// out[e] = sum over splits of parts[s][e] (fixed order)
__global__ void sum_splits_kernel(const float* __restrict__ parts, int splits, int64_t n, float* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) {
        float s = 0.f;
        for (int q = 0; q < splits; ++q) s += parts[(int64_t)q * n + e];
        out[e] = s;
    }
}

// predict_link (src/singlet.cpp:429-430): out[c][f] = (sum over splits of parts[s][c][f]) * link[c][f]
// (link is float [cols][KP] with 1.0 in the rows beyond link_rows)
__global__ void link_rhs_kernel(const float* __restrict__ parts, int splits, int64_t n, const float* __restrict__ link,
                                float* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) {
        float s = 0.f;
        for (int q = 0; q < splits; ++q) s += parts[(int64_t)q * n + e];
        out[e] = s * link[e];
    }
}
// link matrix (double, link_rows x cols column-major) -> float [cols][KP], rows >= link_rows = 1
__global__ void link_to_dev_kernel(const double* __restrict__ src, int link_rows, int KP, int64_t cols,
                                   float* __restrict__ dst) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < cols * KP) {
        const int64_t c = e / KP;
        const int f = (int)(e % KP);
        dst[e] = (f < link_rows) ? (float)src[c * link_rows + f] : 1.f;
    }
}

// cor (src/singlet.cpp:184-197): five running sums in FP64 -> per-CTA partials [grid][5]
__global__ void __launch_bounds__(256)
cor_partial_kernel(const float* __restrict__ X, const float* __restrict__ Y, int64_t n_elems,
                   double* __restrict__ part) {
    double s[5] = {0, 0, 0, 0, 0};
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_elems; e += (int64_t)gridDim.x * blockDim.x) {
        const double x = X[e], y = Y[e];
        s[0] += x; s[1] += y; s[2] += x * y; s[3] += x * x; s[4] += y * y;
    }
    __shared__ double sm[8][5];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const double v = warp_sum(s[q]);
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double v = 0;
        for (int w = 0; w < 8; ++w) v += sm[w][threadIdx.x];
        part[(int64_t)blockIdx.x * 5 + threadIdx.x] = v;
    }
}

// ----------------------------------------------------------------------------------------------
// format conversion
// ----------------------------------------------------------------------------------------------
__global__ void pack_records_kernel(const int32_t* __restrict__ idx, const double* __restrict__ val, int64_t n,
                                    uint2* __restrict__ rec) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) rec[e] = make_uint2((uint32_t)idx[e], __float_as_uint((float)val[e]));
}
__global__ void unpack_records_kernel(const uint2* __restrict__ rec, int64_t n, int32_t* __restrict__ idx,
                                      double* __restrict__ val) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) {
        idx[e] = (int32_t)rec[e].x;
        val[e] = (double)__uint_as_float(rec[e].y);
    }
}
__global__ void colptr_from_p32_kernel(const int32_t* __restrict__ p, int64_t ncol_chunk, int64_t nnz_offset,
                                       int64_t* __restrict__ colptr /*points at chunk's first column*/, int last) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncol_chunk + (last ? 1 : 0)) colptr[c] = nnz_offset + (int64_t)p[c];
}
// input validation after upload: rows in range and strictly ascending within every column (the tile index and the
// shared-memory gather rely on it; a dgCMatrix guarantees it). flag[0] |= 1 (range) | 2 (order).
__global__ void __launch_bounds__(256)
validate_records_kernel(const uint2* __restrict__ rec, const int64_t* __restrict__ colptr, int64_t ncol, int64_t nrow,
                        int* __restrict__ flag) {
    const int lane = threadIdx.x & 31;
    const int64_t col = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (col >= ncol) return;
    int bad = 0;
    const int64_t b = colptr[col], e = colptr[col + 1];
    for (int64_t t = b + lane; t < e; t += 32) {
        const uint32_t r = rec[t].x;
        if ((int64_t)r >= nrow) bad |= 1;
        if (t + 1 < e && rec[t + 1].x <= r) bad |= 2;
    }
    if (bad) atomicOr(flag, bad);
}

// dense-input variants (src/singlet.cpp:370-381, 506-531, 610-634): the dense loops visit EVERY row of a column,
// zeros included, and never skip a column; storing every entry as a record reproduces exactly that.
__global__ void dense_to_records_kernel(const double* __restrict__ D, int64_t nrow, int64_t ncol, uint2* __restrict__ rec,
                                        int64_t* __restrict__ colptr) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < nrow * ncol) rec[e] = make_uint2((uint32_t)(e % nrow), __float_as_uint((float)D[e]));
    if (e <= ncol) colptr[e] = e * nrow;
}

// double k x cols column-major (element (f,c) at c*k+f) -> float [cols][KP], zero padded
__global__ void factor_to_dev_kernel(const double* __restrict__ src, int k, int KP, int64_t cols,
                                     float* __restrict__ dst) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < cols * KP) {
        const int64_t c = e / KP;
        const int f = (int)(e % KP);
        dst[e] = (f < k) ? (float)src[c * k + f] : 0.f;
    }
}
__global__ void factor_to_host_kernel(const float* __restrict__ src, int k, int KP, int64_t cols,
                                      double* __restrict__ dst) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < cols * k) {
        const int64_t c = e / k;
        const int f = (int)(e % k);
        dst[e] = (double)src[c * KP + f];
    }
}

// ----------------------------------------------------------------------------------------------
// Device-side transpose (`Matrix::t(A)`, R/run_nmf.R:40): column-compressed records of X
// (nrow x ncol) -> column-compressed records of X^T with the rows of every column in ascending order, deterministic,
// no sort. The columns of X are cut into blocks; a CTA owns one block and keeps one 32-bit counter per ROW of X in
// shared memory (up to ~57k rows -- the gene dimension -- per pass).
//   1. count:   blockcnt[b][r] = non-zeros of row r inside column block b (shared-memory histogram)
//   2. offsets: per row an exclusive scan over the blocks (in place) + the row totals -> column pointers of X^T
//   3. scatter: the CTA walks its columns IN ORDER; the non-zeros of one column hit distinct rows, so they are placed
//               in parallel with plain read-modify-writes of the per-row cursors, and a barrier separates columns:
//               entries of a row arrive in ascending column order.
// ----------------------------------------------------------------------------------------------
// All three kernels work on the row range [row0, row0 + nr) of X (nr counters fit in shared memory); matrices with more
// rows are transposed in several passes over the records.
__global__ void __launch_bounds__(1024)
transpose_count_kernel(const uint2* __restrict__ rec, const int64_t* __restrict__ colptr, int64_t ncol, int64_t row0, int64_t nr,
                       int64_t cols_per_block, int32_t* __restrict__ blockcnt) {
    extern __shared__ int32_t s_cnt[];
    for (int64_t r = threadIdx.x; r < nr; r += blockDim.x) s_cnt[r] = 0;
    __syncthreads();
    const int64_t c0 = (int64_t)blockIdx.x * cols_per_block;
    const int64_t c1 = c0 + cols_per_block < ncol ? c0 + cols_per_block : ncol;
    if (c0 < c1) {
        const int64_t b = colptr[c0], e = colptr[c1];
        for (int64_t p = b + threadIdx.x; p < e; p += blockDim.x) {
            const int64_t r = (int64_t)rec[p].x - row0;
            if (r >= 0 && r < nr) atomicAdd(&s_cnt[r], 1);
        }
    }
    __syncthreads();
    for (int64_t r = threadIdx.x; r < nr; r += blockDim.x) blockcnt[(int64_t)blockIdx.x * nr + r] = s_cnt[r];
}
// per row: exclusive scan over the blocks (in place) and the row total
__global__ void transpose_offsets_kernel(int32_t* __restrict__ blockcnt, int64_t nr, int n_blocks, int64_t* __restrict__ rowtot) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nr) return;
    int64_t run = 0;
    for (int b = 0; b < n_blocks; ++b) {
        const int32_t c = blockcnt[(int64_t)b * nr + r];
        blockcnt[(int64_t)b * nr + r] = (int32_t)run;  // a row holds < 2^31 non-zeros (it has at most ncol <= 2^31 - 1 entries)
        run += c;
    }
    if (rowtot) rowtot[r] = run;
}
__global__ void __launch_bounds__(1024)
transpose_scatter_kernel(const uint2* __restrict__ rec, const int64_t* __restrict__ colptr, int64_t ncol, int64_t row0, int64_t nr,
                         int64_t cols_per_block, const int32_t* __restrict__ blockoff, const int64_t* __restrict__ tptr,
                         uint2* __restrict__ trec) {
    extern __shared__ int32_t s_cur[];
    for (int64_t r = threadIdx.x; r < nr; r += blockDim.x) s_cur[r] = blockoff[(int64_t)blockIdx.x * nr + r];
    __syncthreads();
    const int64_t c0 = (int64_t)blockIdx.x * cols_per_block;
    const int64_t c1 = c0 + cols_per_block < ncol ? c0 + cols_per_block : ncol;
    for (int64_t c = c0; c < c1; ++c) {
        const int64_t b = colptr[c], e = colptr[c + 1];
        for (int64_t p = b + threadIdx.x; p < e; p += blockDim.x) {
            const uint2 v = rec[p];
            const int64_t r = (int64_t)v.x - row0;
            if (r >= 0 && r < nr) {
                const int32_t slot = s_cur[r];  // rows of one column are distinct: no two threads touch the same cursor
                s_cur[r] = slot + 1;
                trec[tptr[v.x] + slot] = make_uint2((uint32_t)c, v.y);
            }
        }
        __syncthreads();
    }
}

// ----------------------------------------------------------------------------------------------
// single-CTA exclusive scan of int64 counts (column counts -> pointers). n <= a few million.
// out has n + 1 entries.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
exclusive_scan_kernel(const int64_t* __restrict__ counts, int64_t n, int64_t* __restrict__ out) {
    __shared__ int64_t warp_tot[32], warp_excl[32];
    __shared__ int64_t block_total, carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t idx = base + threadIdx.x;
        const int64_t v = (idx < n) ? counts[idx] : 0;
        int64_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int64_t w = warp_tot[lane];
            int64_t wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_excl[lane] = wi - w;
            if (lane == 31) block_total = wi;
        }
        __syncthreads();
        const int64_t carry = carry_s;
        if (idx < n) out[idx] = carry + warp_excl[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + block_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry_s;
}

// ----------------------------------------------------------------------------------------------
// ordered compaction, warp per column: candidates 0..n_cand-1 are tested 32 at a time; the survivors
// are written in candidate order (so row indices come out ascending). Used for the mask lists and
// for the synthetic generator. MODE 0 = count only, 1 = fill.
// ----------------------------------------------------------------------------------------------
struct MaskGen {  // held-out entries of one column
    uint64_t seed;
    ModP mod;
    int mask_t;  // 0: column = cell, candidate = gene; 1: column = gene, candidate = cell
    int64_t col_offset, row_offset;
    __device__ __forceinline__ bool test(int64_t col, int64_t cand) const {
        const uint64_t cell = (uint64_t)(mask_t ? cand + row_offset : col + col_offset);
        const uint64_t gene = (uint64_t)(mask_t ? col + col_offset : cand + row_offset);
        return is_multiple(hash_pair(seed, cell, gene), mod);
    }
};

template <int MODE>
__global__ void __launch_bounds__(256)
mask_lists_kernel(MaskGen gen, int64_t ncol, int64_t n_cand, const int64_t* __restrict__ colptr,
                  const uint2* __restrict__ rec, int64_t* __restrict__ counts, const int64_t* __restrict__ mptr,
                  uint2* __restrict__ mrec) {
    const int lane = threadIdx.x & 31;
    const int64_t col = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (col >= ncol) return;
    int64_t n_out = 0;
    const int64_t out0 = (MODE == 1) ? mptr[col] : 0;
    const int64_t rb = (MODE == 1) ? colptr[col] : 0, re = (MODE == 1) ? colptr[col + 1] : 0;
    for (int64_t c0 = 0; c0 < n_cand; c0 += 32) {
        const int64_t cand = c0 + lane;
        const bool hit = (cand < n_cand) && gen.test(col, cand);
        const uint32_t bal = __ballot_sync(0xffffffffu, hit);
        if (MODE == 1 && hit) {
            // value of X at (cand, col): binary search among the column's ascending rows
            int64_t lo = rb, hi = re;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if ((int64_t)rec[mid].x < cand) lo = mid + 1; else hi = mid;
            }
            const uint32_t vbits = (lo < re && (int64_t)rec[lo].x == cand) ? rec[lo].y : 0u;
            mrec[out0 + n_out + __popc(bal & ((1u << lane) - 1u))] = make_uint2((uint32_t)cand, vbits);
        }
        n_out += __popc(bal);
    }
    if (MODE == 0 && lane == 0) counts[col] = n_out;
}

// training copy of the record stream: held-out non-zeros get value 0 (they then add nothing to b,
// which is what skipping them does in src/singlet.cpp:452-457)
__global__ void __launch_bounds__(256)
mask_records_kernel(MaskGen gen, int64_t ncol, const int64_t* __restrict__ colptr, const uint2* __restrict__ rec,
                    uint2* __restrict__ rec_train, unsigned long long* __restrict__ n_held_nz) {
    const int lane = threadIdx.x & 31;
    const int64_t col = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (col >= ncol) return;
    unsigned long long held = 0;
    for (int64_t p = colptr[col] + lane; p < colptr[col + 1]; p += 32) {
        uint2 r = rec[p];
        if (gen.test(col, (int64_t)r.x)) { r.y = 0u; ++held; }
        rec_train[p] = r;
    }
    held = (unsigned long long)warp_sum((double)held);
    if (lane == 0 && held) atomicAdd(n_held_nz, held);
}

// ----------------------------------------------------------------------------------------------
// mse_test (src/singlet.cpp:536-568) / harness train MSE, warp per cell column -- the fused loss kernel.
//   test : mean over held-out genes g of (sum_f W[g][f] d[f] H[c][f] - A[g][c])^2   (0 if none)
//   train: the same over the genes that are NOT held out (all m genes when mask == NULL), via
//          S_all - S_heldout with S_all = hd^T (W^T W) hd - 2 sum_nz a*pred + sum_nz a^2.
// which = 0: test, 1: train, 2: both from one pass over the held-out list and the non-zeros.
// LPE = KP / 4 lanes cooperate on one entry: each reads 16 bytes of the gene's W row, so a group reads one contiguous
// 4 * KP-byte piece (a whole 128-byte line at KP = 32) and the warp covers 32 / LPE entries per step; the partial dot
// products are folded with log2(LPE) shuffles. (One entry per lane -- every lane walking a whole row of its own -- kept the
// L1 at 98 % of its wavefront rate: 2.17 ms for 7.5e7 held-out entries at k = 32, profiles/r2_summary.md section 4.)
// Writes per-column losses; the caller reduces them in fixed order.
// ----------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(128)
mse_kernel(const int64_t* __restrict__ colptr, const uint2* __restrict__ rec, const int64_t* __restrict__ mptr,
           const uint2* __restrict__ mrec, const float* __restrict__ W, const double* __restrict__ d,
           const float* __restrict__ H, const double* __restrict__ gram_w /*[KP*KP] W^T W, jitter irrelevant*/,
           int64_t m_genes, int64_t ncol, int k, int which, double* __restrict__ losses,
           double* __restrict__ losses_train /* which == 2: test loss -> losses, train loss -> losses_train, one pass */) {
    constexpr int LPE = KP / 4, EPS = 32 / LPE;  // lanes per entry, entries per warp step
    const int lane = threadIdx.x & 31;
    const int lig = lane % LPE, grp = lane / LPE;
    const int64_t col = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (col >= ncol) return;  // warp-uniform
    // my four entries of hd[f] = d[f] * h[c][f]
    float hd[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int f = 4 * lig + q;
        hd[q] = (f < k) ? (float)(d[f] * (double)H[col * KP + f]) : 0.f;
    }
    auto predict = [&](int64_t gene) {  // all lanes of the warp call it together; the result is on every lane of the group
        const float4 w4 = *reinterpret_cast<const float4*>(W + gene * KP + 4 * lig);
        float p = w4.x * hd[0];
        p = fmaf(w4.y, hd[1], p); p = fmaf(w4.z, hd[2], p); p = fmaf(w4.w, hd[3], p);
#pragma unroll
        for (int o = 1; o < LPE; o <<= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
        return p;
    };

    double s_held = 0.0;
    int64_t n_held = 0;
    if (mptr != nullptr) {
        const int64_t mb = mptr[col], me = mptr[col + 1];
        n_held = me - mb;
        for (int64_t base = mb; base < me; base += EPS) {  // warp-uniform trip count
            const int64_t p = base + grp;
            const bool ok = p < me;
            const uint2 r = ok ? mrec[p] : make_uint2(0u, 0u);
            const double res = (double)predict((int64_t)r.x) - (double)__uint_as_float(r.y);
            if (ok && lig == 0) s_held += res * res;
        }
        s_held = warp_sum(s_held);
    }
    if (which != 1 && lane == 0) losses[col] = n_held > 0 ? s_held / (double)n_held : 0.0;
    if (which == 0) return;
    // train: S_all
    double quad = 0.0;
    for (int i = lane; i < k; i += 32) {
        double row = 0.0;
        for (int j = 0; j < k; ++j) row += gram_w[i * KP + j] * (double)(float)(d[j] * (double)H[col * KP + j]);
        quad += row * (double)(float)(d[i] * (double)H[col * KP + i]);
    }
    double lin = 0.0;
    const int64_t cb = colptr[col], ce = colptr[col + 1];
    for (int64_t base = cb; base < ce; base += EPS) {
        const int64_t p = base + grp;
        const bool ok = p < ce;
        const uint2 r = ok ? rec[p] : make_uint2(0u, 0u);
        const double a = (double)__uint_as_float(r.y);
        const double pr = (double)predict((int64_t)r.x);
        if (ok && lig == 0) lin += a * a - 2.0 * a * pr;
    }
    const double s_all = warp_sum(quad + lin);
    const int64_t n_train = m_genes - n_held;
    if (lane == 0) (which == 2 ? losses_train : losses)[col] = n_train > 0 ? (s_all - s_held) / (double)n_train : 0.0;
}

// ----------------------------------------------------------------------------------------------
// synthetic generator, warp per column (see SynthSpec in common.cuh). orientation 0: column = cell,
// candidates = strata; orientation 1: column = gene, candidates = cells.
// ----------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256)
synth_kernel(SynthSpec sp, int orientation, int64_t col0, int64_t ncol, int64_t row0, int64_t nrows,
             int64_t* __restrict__ counts, const int64_t* __restrict__ colptr, uint2* __restrict__ rec) {
    const int lane = threadIdx.x & 31;
    const int64_t lc = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (lc >= ncol) return;
    const int64_t col = col0 + lc;
    // candidates: strata (orientation 0) or cells (orientation 1) that can hit a row of [row0, row0 + nrows)
    const int64_t c_begin = orientation == 0 ? row0 / sp.S : row0;
    const int64_t c_end = orientation == 0 ? (row0 + nrows + sp.S - 1) / sp.S : row0 + nrows;
    const int64_t out0 = (MODE == 1) ? colptr[lc] : 0;
    int64_t n_out = 0;
    for (int64_t c0 = c_begin; c0 < c_end; c0 += 32) {
        const int64_t cand = c0 + lane;
        bool hit = false;
        int64_t gene = 0;
        float value = 0.f;
        if (cand < c_end) {
            if (orientation == 0) {
                hit = synth_entry(sp, (uint64_t)col, (uint32_t)cand, gene, value) && gene >= row0 && gene < row0 + nrows;
            } else {
                hit = synth_entry(sp, (uint64_t)cand, (uint32_t)(col / sp.S), gene, value) && gene == col;
            }
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, hit);
        if (MODE == 1 && hit) {
            const uint32_t row = (uint32_t)((orientation == 0 ? gene : cand) - row0);
            rec[out0 + n_out + __popc(bal & ((1u << lane) - 1u))] = make_uint2(row, __float_as_uint(value));
        }
        n_out += __popc(bal);
    }
    if (MODE == 0 && lane == 0) counts[lc] = n_out;
}

// rng test hooks (sgl_mask_rand / sgl_mask_draw)
__global__ void hash_pairs_kernel(uint64_t seed, const uint64_t* __restrict__ i, const uint64_t* __restrict__ j,
                                  int64_t n, uint64_t* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) out[e] = hash_pair(seed, i[e], j[e]);
}
__global__ void draw_pairs_kernel(uint64_t seed, ModP mod, const uint64_t* __restrict__ i,
                                  const uint64_t* __restrict__ j, int64_t n, uint8_t* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) out[e] = is_multiple(hash_pair(seed, i[e], j[e]), mod) ? 1 : 0;
}

}  // namespace sgl
