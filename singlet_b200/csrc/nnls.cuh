// nnls.cuh -- sequential coordinate-descent NNLS with L1/L2 (reference src/singlet.cpp:229-250;
// exact rules in SURVEY.md App. A-5), FP32 state.
//
// Two kernels:
//  * nnls_cols_kernel<KP>: plain predict(). The Gram matrix is shared by every column, so ONE THREAD
//    solves ONE COLUMN: b[KP] lives in registers (the coordinate loop is fully unrolled so all
//    indices are static), x in a thread-private shared-memory column, and the Gram column a[:, i]
//    is broadcast to the warp from shared memory (LDS.128, one wavefront per 4 values). This
//    issues ~KP+20 instructions per coordinate step per 32 columns instead of ~10 per step per
//    single column for a warp-per-column layout (an ~6x saving at KP = 32).
//  * nnls_masked_kernel<KP>: predict_mask(). Every column has its own Gram a_i = a - W_M W_M^T
//    (reference src/singlet.cpp:460-462), so ONE WARP solves ONE COLUMN: lane j holds rows
//    j, j+32 of a_i in registers, builds the correction from the held-out rows, then runs the
//    coordinate loop with shuffles.
// Both write the row sums of the new solution (the local part of `scale`'s d, src/singlet.cpp:220)
// as per-CTA double partials, reduced in a fixed order afterwards (deterministic).
#pragma once
#include "common.cuh"
#include <type_traits>

namespace sgl {

constexpr int NNLS_MAX_SWEEPS = 100;  // uint8_t it < 100 (src/singlet.cpp:231)

// one coordinate step on scalars; returns delta to apply as b -= a[:, i] * delta.
// Mirrors App. A-5: diff = b_i / a_ii - L1 + L2 * x_i; clamp at zero; tol bookkeeping.
__device__ __forceinline__ float cd_step(float bi, float inv_aii, float& xi, float L1, float L2, float& tol) {
    float diff = bi * inv_aii;
    diff -= L1;                 // L1 == 0 -> exact no-op
    diff = fmaf(L2, xi, diff);  // L2 == 0 -> exact no-op
    float delta;
    if (-diff > xi) {
        delta = -xi;  // x_i == 0 -> delta = -0: b unchanged, tol unchanged (matches `if (x != 0)`)
        if (xi != 0.f) tol = 1.f;
        xi = 0.f;
    } else {
        delta = diff;
        if (diff != 0.f) {
            xi += diff;
            tol += fabsf(__fdividef(diff, xi + 1e-15f));
        }
    }
    return delta;
}

// ----------------------------------------------------------------------------------------------
// plain: thread per column
// ----------------------------------------------------------------------------------------------
template <int KP>
struct NnlsCfg {
    static constexpr int THREADS = 128;
    static constexpr int NCL = (KP <= 32) ? 2 : 1;  // columns per lane: 64 registers of right-hand sides either way
    static constexpr int MIN_CTAS = 3;              // register cap 170
};

// branch-free coordinate step (same arithmetic as cd_step): returns MINUS the delta, i.e. the
// multiplier m of  b += a[:, i] * m.
__device__ __forceinline__ float cd_step_nb(float bi, float inv_aii, float& xi, float L1, float L2, float& tol) {
    const float diff = fmaf(L2, xi, fmaf(bi, inv_aii, -L1));
    const bool clamp = (-diff > xi);
    const float xsum = xi + diff;
    float r;  // 1 / (x_new + 1e-15): single MUFU.RCP, evaluated on both paths so that no branch is needed
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(xsum + 1e-15f));
    const float tol_clamp = (xi != 0.f) ? 1.f : tol;  // `tol = 1` only when x_i actually changes
    const float tol_step = tol + fabsf(diff * r);     // diff == 0 adds exactly 0
    tol = clamp ? tol_clamp : tol_step;
    const float m = clamp ? xi : -diff;
    xi = clamp ? 0.f : xsum;
    return m;
}

__device__ __forceinline__ void ffma2_bcast(unsigned long long& acc, unsigned long long w, float v) {
    unsigned long long vv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(vv) : "f"(v));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(w), "l"(vv));
}
__device__ __forceinline__ float lo32(unsigned long long v) { return __uint_as_float((uint32_t)(v & 0xffffffffull)); }
__device__ __forceinline__ float hi32(unsigned long long v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}

// Gram matrix and reciprocal diagonal of the current half-iteration, in constant memory: every thread
// of the plain solver reads the same a[i][j] at the same time, so the loads go through the uniform
// datapath (LDCU -> uniform registers feeding FFMA2 directly) and never touch the shared-memory
// crossbar or the vector register file. Refreshed by cudaMemcpyToSymbolAsync before each launch.
// One symbol, one copy per solve: the Gram in the first 64 x 64 floats, the reciprocal diagonal behind it.
__constant__ __align__(16) float c_gram[64 * 64 + 64];
#define c_inv_diag (c_gram + 64 * 64)

// Persistent kernel with lane refill. Every lane solves NCL = 2 columns at a time (two independent
// dependency chains per thread, and every Gram row fetched once feeds both). Columns converge after
// very different numbers of sweeps (from ~30 up to the 100-sweep cap), so a slot whose column is done
// writes it back and claims the next unsolved column from a global counter instead of idling until the
// slowest lane of its warp finishes. Columns are independent, so the result does not depend on which
// lane solves which column. Row sums are taken afterwards by rowsum_partial_kernel.
// NT threads per CTA, NCL columns per lane: <THREADS, 2> for large column counts; <32, 1> when there are
// too few columns to fill the chip otherwise (e.g. a gene shard of the W update on 8 GPUs).
// MINB: CTAs per SM the register allocation must allow (default: 3 CTAs of 128 threads -- <= 170 registers, the fastest
// code when the columns make many rounds of the persistent grid -- or 8 of 32 threads).
template <int KP, int NT, int NCL, int MINB = ((NT == 32) ? 8 : NnlsCfg<KP>::MIN_CTAS)>
__global__ void __launch_bounds__(NT, MINB)
nnls_cols_kernel(const float* __restrict__ Bparts,  // [splits][ncol][KP]
                 int splits, float* __restrict__ X,  // [ncol][KP] warm start in / solution out
                 const int64_t* __restrict__ colptr, int64_t ncol, int k, float L1, float L2,
                 unsigned long long* __restrict__ next_col,  // global work counter (zeroed by the host)
                 unsigned long long* __restrict__ stats,     // optional: [0] += sweeps, [1] += columns solved
                 int64_t col_end = -1)                       // solve columns [0, col_end) only (-1: all ncol; see the tail split)
{
    __shared__ float sx[NCL * KP * NT];  // sx[(s * KP + i) * NT + tid]
    if (col_end < 0) col_end = ncol;

    const int lane = threadIdx.x & 31;
    float2 b[NCL][KP / 2];  // right-hand sides as FP32 pairs (the rank-1 update runs on FFMA2)
    int64_t col[NCL];
    float tol[NCL];
    int sweeps[NCL];
#pragma unroll
    for (int s = 0; s < NCL; ++s) {
        col[s] = -1;
        tol[s] = 1.f;
        sweeps[s] = 0;
#pragma unroll
        for (int j2 = 0; j2 < KP / 2; ++j2) b[s][j2] = make_float2(0.f, 0.f);
    }
    bool exhausted = false;  // the counter ran past ncol
    const float kf = (float)k;
#pragma unroll
    for (int j = 0; j < NCL * KP; ++j) sx[j * NT + threadIdx.x] = 0.f;

    while (true) {
        // ---- retire finished columns and claim new ones (per slot; the warp takes this path together) ----
#pragma unroll
        for (int s = 0; s < NCL; ++s) {
            const bool finished = (col[s] >= 0) && (sweeps[s] >= NNLS_MAX_SWEEPS || !(tol[s] / kf > 1e-8f));
            bool need = (col[s] < 0 || finished) && !exhausted;
            if (finished) {
                if (stats) { atomicAdd(&stats[0], (unsigned long long)sweeps[s]); atomicAdd(&stats[1], 1ull); }
                float4* xd = reinterpret_cast<float4*>(X + col[s] * KP);
#pragma unroll
                for (int j4 = 0; j4 < KP / 4; ++j4)
                    xd[j4] = make_float4(sx[(s * KP + 4 * j4 + 0) * NT + threadIdx.x], sx[(s * KP + 4 * j4 + 1) * NT + threadIdx.x],
                                         sx[(s * KP + 4 * j4 + 2) * NT + threadIdx.x], sx[(s * KP + 4 * j4 + 3) * NT + threadIdx.x]);
                col[s] = -1;
                // an empty slot is made inert instead of being predicated off in the sweep: with b = 0 and
                // x = 0 every coordinate step clamps at zero (diff = -L1 <= 0) and multiplies the update by 0
#pragma unroll
                for (int j2 = 0; j2 < KP / 2; ++j2) b[s][j2] = make_float2(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < KP; ++j) sx[(s * KP + j) * NT + threadIdx.x] = 0.f;
            }
            while (__any_sync(0xffffffffu, need)) {
                const uint32_t mask = __ballot_sync(0xffffffffu, need);
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(next_col, (unsigned long long)__popc(mask));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (need) {
                    const int64_t c = (int64_t)base + __popc(mask & ((1u << lane) - 1u));
                    if (c >= col_end) {
                        exhausted = true;
                        need = false;
                    } else if (colptr[c] != colptr[c + 1]) {  // empty columns are skipped (:340)
#pragma unroll
                        for (int j2 = 0; j2 < KP / 2; ++j2) b[s][j2] = make_float2(0.f, 0.f);
                        for (int sp = 0; sp < splits; ++sp) {
                            const float4* src = reinterpret_cast<const float4*>(Bparts + ((int64_t)sp * ncol + c) * KP);
#pragma unroll
                            for (int j4 = 0; j4 < KP / 4; ++j4) {
                                const float4 v = src[j4];
                                b[s][2 * j4].x += v.x; b[s][2 * j4].y += v.y; b[s][2 * j4 + 1].x += v.z; b[s][2 * j4 + 1].y += v.w;
                            }
                        }
                        const float4* xs = reinterpret_cast<const float4*>(X + c * KP);
#pragma unroll
                        for (int j4 = 0; j4 < KP / 4; ++j4) {
                            const float4 v = xs[j4];
                            sx[(s * KP + 4 * j4 + 0) * NT + threadIdx.x] = v.x; sx[(s * KP + 4 * j4 + 1) * NT + threadIdx.x] = v.y;
                            sx[(s * KP + 4 * j4 + 2) * NT + threadIdx.x] = v.z; sx[(s * KP + 4 * j4 + 3) * NT + threadIdx.x] = v.w;
                        }
                        col[s] = c;
                        tol[s] = 1.f;
                        sweeps[s] = 0;
                        need = false;
                    }  // else: empty column, claim another one
                }
            }
        }
        bool any = false;
#pragma unroll
        for (int s = 0; s < NCL; ++s) {
            any = any || (col[s] >= 0);
            tol[s] = 0.f;
            ++sweeps[s];
        }
        if (!__any_sync(0xffffffffu, any)) break;

        // ---- one sweep (src/singlet.cpp:231-248) for every slot that holds a column ----
        if constexpr (NCL == 1 && KP >= 64) {
            // One column per lane leaves ONE dependency chain per thread (b_i -> diff -> m -> b) and the in-order warp waits on
            // it at every coordinate (FP32 pipe 52 % busy, the FFMA / FSEL at the head of the chain hold 29 % of the stall
            // samples -- profiles/r2_summary.md; reading half of each Gram row from shared memory instead of constant memory
            // was tried first: -1.6 % with 8 pairs, +7.5 % with 16). Two coordinates per block: m_i, then b_{i+1} updated by a single scalar FMA,
            // then m_{i+1}; only then are both rank-1 updates applied to all of b, in coordinate order (bit-identical to
            // step-by-step), so that the next block's chain has 2 x KP/2 independent FFMA2 to hide behind.
            // x of the next block is read one block ahead, the two reciprocals of the tol bookkeeping are issued before the
            // rank-1 updates and consumed after them (the in-order warp otherwise waits on LDS / MUFU results -- short
            // scoreboard was the top stall -- in front of 64 independent FFMA2)
            float xp0 = sx[threadIdx.x], xp1 = sx[NT + threadIdx.x];
#pragma unroll
            for (int i = 0; i < KP; i += 2) {
                if (i + 1 < k) {  // uniform
                    const float x0 = xp0, x1 = xp1;
                    if (i + 2 < KP) { xp0 = sx[(i + 2) * NT + threadIdx.x]; xp1 = sx[((i + 3 < KP) ? i + 3 : i + 2) * NT + threadIdx.x]; }
                    // the arithmetic of cd_step_nb, twice, with the tol part deferred
                    const float d0 = fmaf(L2, x0, fmaf(b[0][i >> 1].x, c_inv_diag[i], -L1));
                    const bool cl0 = (-d0 > x0);
                    const float xs0 = x0 + d0;
                    const float m0 = cl0 ? x0 : -d0;
                    const float b1 = fmaf(c_gram[i * KP + i + 1], m0, b[0][i >> 1].y);
                    const float d1 = fmaf(L2, x1, fmaf(b1, c_inv_diag[i + 1], -L1));
                    const bool cl1 = (-d1 > x1);
                    const float xs1 = x1 + d1;
                    const float m1 = cl1 ? x1 : -d1;
                    float r0, r1;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(xs0 + 1e-15f));
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(xs1 + 1e-15f));
                    const float2 mm0 = make_float2(m0, m0), mm1 = make_float2(m1, m1);
                    const float2* a0 = reinterpret_cast<const float2*>(c_gram + i * KP);
                    const float2* a1 = reinterpret_cast<const float2*>(c_gram + (i + 1) * KP);
#pragma unroll
                    for (int j2 = 0; j2 < KP / 2; ++j2) {
                        b[0][j2] = __ffma2_rn(a0[j2], mm0, b[0][j2]);
                        b[0][j2] = __ffma2_rn(a1[j2], mm1, b[0][j2]);
                    }
                    float tl = tol[0];
                    tl = cl0 ? ((x0 != 0.f) ? 1.f : tl) : tl + fabsf(d0 * r0);
                    tl = cl1 ? ((x1 != 0.f) ? 1.f : tl) : tl + fabsf(d1 * r1);
                    tol[0] = tl;
                    sx[(i)*NT + threadIdx.x] = cl0 ? 0.f : xs0;
                    sx[(i + 1) * NT + threadIdx.x] = cl1 ? 0.f : xs1;
                } else if (i < k) {  // odd rank: the last coordinate alone
                    float xi = sx[(i)*NT + threadIdx.x], tl = tol[0];
                    const float mult = cd_step_nb(b[0][i >> 1].x, c_inv_diag[i], xi, L1, L2, tl);
                    tol[0] = tl;
                    sx[(i)*NT + threadIdx.x] = xi;
                    const float2 mm = make_float2(mult, mult);
                    const float2* ai = reinterpret_cast<const float2*>(c_gram + i * KP);
#pragma unroll
                    for (int j2 = 0; j2 < KP / 2; ++j2) b[0][j2] = __ffma2_rn(ai[j2], mm, b[0][j2]);
                }
            }
        } else {
#pragma unroll
        for (int i = 0; i < KP; ++i) {
            if (i < k) {  // uniform
                float2 mm[NCL];
#pragma unroll
                for (int s = 0; s < NCL; ++s) {
                    float xi = sx[(s * KP + i) * NT + threadIdx.x], tl = tol[s];
                    const float bi = (i & 1) ? b[s][i >> 1].y : b[s][i >> 1].x;
                    const float mult = cd_step_nb(bi, c_inv_diag[i], xi, L1, L2, tl);
                    tol[s] = tl;
                    sx[(s * KP + i) * NT + threadIdx.x] = xi;
                    mm[s] = make_float2(mult, mult);
                }
                const float2* ai = reinterpret_cast<const float2*>(c_gram + i * KP);
#pragma unroll
                for (int j2 = 0; j2 < KP / 2; ++j2) {
                    const float2 a2 = ai[j2];
#pragma unroll
                    for (int s = 0; s < NCL; ++s) b[s][j2] = __ffma2_rn(a2, mm[s], b[s][j2]);
                }
            }
        }
        }
    }
}

// row sums of X [cols][KP] in FP64: per-CTA partials [grid][KP], reduced in fixed order afterwards
template <int KP>
__global__ void __launch_bounds__(256)
rowsum_partial_kernel(const float* __restrict__ X, int64_t cols, double* __restrict__ part) {
    constexpr int CPI = 256 / KP;  // columns per pass (KP <= 128 -> >= 2)
    __shared__ double sm[256];
    const int f = threadIdx.x % KP, sub = threadIdx.x / KP;
    double s = 0.0;
    for (int64_t c = (int64_t)blockIdx.x * CPI + sub; c < cols; c += (int64_t)gridDim.x * CPI) s += (double)X[c * KP + f];
    sm[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < KP) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < CPI; ++q) t += sm[q * KP + threadIdx.x];
        part[(int64_t)blockIdx.x * KP + threadIdx.x] = t;
    }
}

// ----------------------------------------------------------------------------------------------
// generic fallback for KP = 128 (b and x both in thread-private shared memory; slow path for the
// rare ranks above 64)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
nnls_cols_big_kernel(const float* __restrict__ Bparts, int splits, float* __restrict__ X,
                     const float* __restrict__ gram_f, const float* __restrict__ inv_diag,
                     const int64_t* __restrict__ colptr, int64_t ncol, int k, int KP, float L1, float L2,
                     double* __restrict__ rowsum_part) {
    extern __shared__ float sm[];
    constexpr int NT = 32;
    float* sb = sm;                  // [KP][NT]
    float* sx = sm + (size_t)KP * NT;  // [KP][NT]
    const int64_t col = (int64_t)blockIdx.x * NT + threadIdx.x;
    const bool in_range = col < ncol;
    const bool solve = in_range && (colptr[col] != colptr[col + 1]);
    for (int j = 0; j < KP; ++j) {
        float bj = 0.f, xj = 0.f;
        if (in_range) {
            for (int s = 0; s < splits; ++s) bj += Bparts[((int64_t)s * ncol + col) * KP + j];
            xj = X[col * KP + j];
        }
        sb[j * NT + threadIdx.x] = bj;
        sx[j * NT + threadIdx.x] = xj;
    }
    float tol = 1.f;
    const float kf = (float)k;
    if (solve) {
        for (int sweep = 0; sweep < NNLS_MAX_SWEEPS && (tol / kf > 1e-8f); ++sweep) {
            tol = 0.f;
            for (int i = 0; i < k; ++i) {
                float xi = sx[i * NT + threadIdx.x];
                const float delta = cd_step(sb[i * NT + threadIdx.x], inv_diag[i], xi, L1, L2, tol);
                sx[i * NT + threadIdx.x] = xi;
                if (delta != 0.f)
                    for (int j = 0; j < k; ++j) sb[j * NT + threadIdx.x] = fmaf(-gram_f[i * KP + j], delta, sb[j * NT + threadIdx.x]);
            }
        }
    }
    for (int j = 0; j < KP; ++j) {
        const float xj = sx[j * NT + threadIdx.x];
        if (solve) X[col * KP + j] = xj;
        const double s = warp_sum((double)xj);
        if (threadIdx.x == 0) rowsum_part[(int64_t)blockIdx.x * KP + j] = s;
    }
}

// ----------------------------------------------------------------------------------------------
// masked: warp per column, per-column Gram correction
// ----------------------------------------------------------------------------------------------
template <int KP>
struct MaskedCfg {
    static constexpr int RPL = (KP + 31) / 32;  // Gram rows per lane
    static constexpr int WARPS = 4;
};

template <int KP>
__global__ void __launch_bounds__(MaskedCfg<KP>::WARPS * 32)
nnls_masked_kernel(const float* __restrict__ Bparts, int splits, float* __restrict__ X,
                   const float* __restrict__ gram_f,   // [KP][KP] float, jitter-free part is fine (see below)
                   const float* __restrict__ F,        // gather factor [rows][KP]
                   const int64_t* __restrict__ colptr, // X's column pointers (empty-column skip)
                   const int64_t* __restrict__ mptr,   // [ncol + 1] held-out list pointers
                   const uint2* __restrict__ mrec,     // held-out records {row, value bits}
                   int64_t ncol, int k, float L1, float L2, double* __restrict__ rowsum_part,
                   // corrections computed beforehand on the tensor cores (gramcorr.cuh), see nnls_masked_sub_kernel
                   const float* __restrict__ gm = nullptr, int64_t blk0 = 0)
{
    constexpr int RPL = MaskedCfg<KP>::RPL;
    constexpr int WARPS = MaskedCfg<KP>::WARPS;
    __shared__ double sred[WARPS][KP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t blk = (int64_t)blockIdx.x + blk0;
    const int64_t col = blk * WARPS + warp;
    const bool in_range = col < ncol;
    const bool solve = in_range && (colptr[col] != colptr[col + 1]);

    // lane holds rows r = lane + 32*c of a_i: a[c][i] = a_i[r][i]
    float a[RPL][KP];
    float b[RPL], x[RPL];
#pragma unroll
    for (int c = 0; c < RPL; ++c) {
        const int r = lane + 32 * c;
        b[c] = 0.f;
        x[c] = 0.f;
#pragma unroll
        for (int i = 0; i < KP; ++i) a[c][i] = 0.f;
        if (in_range && r < KP) {
            for (int s = 0; s < splits; ++s) b[c] += Bparts[((int64_t)s * ncol + col) * KP + r];
            x[c] = X[col * KP + r];
        }
    }

    if (solve) {
        // correction G_M = sum over held-out rows of f f^T  (reference: AAt(submat(w, idx)), :460-461)
        if (gm != nullptr) {  // precomputed: rows lane and lane + 32 of this column's KP x KP block
            const float* src = gm + (col - blk0 * WARPS) * (int64_t)(KP * KP);
#pragma unroll
            for (int c = 0; c < RPL; ++c) {
                const int r = lane + 32 * c;
                if (r < KP) {
#pragma unroll
                    for (int i4 = 0; i4 < KP / 4; ++i4) {
                        const float4 v = *reinterpret_cast<const float4*>(src + r * KP + 4 * i4);
                        a[c][4 * i4 + 0] = v.x; a[c][4 * i4 + 1] = v.y; a[c][4 * i4 + 2] = v.z; a[c][4 * i4 + 3] = v.w;
                    }
                }
            }
        }
        const int64_t mb = gm ? 0 : mptr[col], me = gm ? 0 : mptr[col + 1];
        for (int64_t p = mb; p < me; ++p) {
            const int64_t row = (int64_t)mrec[p].x;
            const float* fr = F + row * KP;
            float mine[RPL];
#pragma unroll
            for (int c = 0; c < RPL; ++c) mine[c] = (lane + 32 * c < KP) ? fr[lane + 32 * c] : 0.f;
            const float4* fr4 = reinterpret_cast<const float4*>(fr);
#pragma unroll
            for (int i4 = 0; i4 < KP / 4; ++i4) {
                const float4 f4 = fr4[i4];  // uniform address: broadcast
#pragma unroll
                for (int c = 0; c < RPL; ++c) {
                    a[c][4 * i4 + 0] = fmaf(mine[c], f4.x, a[c][4 * i4 + 0]);
                    a[c][4 * i4 + 1] = fmaf(mine[c], f4.y, a[c][4 * i4 + 1]);
                    a[c][4 * i4 + 2] = fmaf(mine[c], f4.z, a[c][4 * i4 + 2]);
                    a[c][4 * i4 + 3] = fmaf(mine[c], f4.w, a[c][4 * i4 + 3]);
                }
            }
        }
        // a_i = (G + eps I) - (G_M + eps I): the jitters cancel (App. A-11), so subtract from the
        // jitter-free FP32 copy of G.
        float inv[RPL];
#pragma unroll
        for (int c = 0; c < RPL; ++c) {
            const int r = lane + 32 * c;
            inv[c] = 0.f;
            if (r < KP) {
#pragma unroll
                for (int i = 0; i < KP; ++i) a[c][i] = gram_f[r * KP + i] - a[c][i];
            }
        }
        // diagonal reciprocals: a_i[r][r] is element a[c][r] of the lane owning row r
#pragma unroll
        for (int i = 0; i < KP; ++i) {
            const float dia = a[i / 32][i];  // valid on lane i % 32
            if ((i & 31) == lane) inv[i / 32] = 1.0f / dia;
        }

        // Coordinate loop. Every lane runs the (branch-free) scalar step on ITS OWN coordinate of slot i / 32 --
        // only the owner lane's result is meaningful -- and one shuffle broadcasts the owner's multiplier, so a
        // step costs one SHFL instead of three. The reference's running `tol` (reset to 1 by a clamp event, then
        // accumulating the later terms, src/singlet.cpp:240-246) is rebuilt at the end of the sweep from the
        // per-lane records: tol = [any event] + sum of the terms of the coordinates after the last event.
        float tol = 1.f;
        const float kf = (float)k;
        for (int sweep = 0; sweep < NNLS_MAX_SWEEPS && (tol / kf > 1e-8f); ++sweep) {
            float term[RPL];
            bool ev[RPL];
#pragma unroll
            for (int c = 0; c < RPL; ++c) { term[c] = 0.f; ev[c] = false; }
#pragma unroll
            for (int i = 0; i < KP; ++i) {
                if (i < k) {  // uniform
                    const int owner = i & 31, c_own = i >> 5;
                    float xi = x[c_own];
                    float t0 = 0.f;  // starts from 0: a clamp event leaves 1, a regular step leaves its term
                    const float xold = xi;
                    const float m_own = cd_step_nb(b[c_own], inv[c_own], xi, L1, L2, t0);
                    const float mult = __shfl_sync(0xffffffffu, m_own, owner);
                    if (lane == owner) {
                        const bool clamp_ev = (xi == 0.f) && (xold != 0.f) && (t0 == 1.f);
                        x[c_own] = xi;
                        ev[c_own] = clamp_ev;
                        term[c_own] = clamp_ev ? 0.f : t0;
                    }
#pragma unroll
                    for (int c = 0; c < RPL; ++c) b[c] = fmaf(a[c][i], mult, b[c]);
                }
            }
            // rebuild tol: position of the last clamp event in coordinate order i = 32 * c + lane
            int last = -1;
#pragma unroll
            for (int c = 0; c < RPL; ++c) {
                const uint32_t mk = __ballot_sync(0xffffffffu, ev[c]);
                if (mk) last = 32 * c + (31 - __clz(mk));
            }
            float part = 0.f;
#pragma unroll
            for (int c = 0; c < RPL; ++c) part += (32 * c + lane > last) ? term[c] : 0.f;
            tol = warp_sum(part) + (last >= 0 ? 1.f : 0.f);
        }
    }

#pragma unroll
    for (int c = 0; c < RPL; ++c) {
        const int r = lane + 32 * c;
        if (r < KP) {
            if (solve) X[col * KP + r] = x[c];
            sred[warp][r] = in_range ? (double)x[c] : 0.0;
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < KP; t += WARPS * 32) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) s += sred[w][t];
        rowsum_part[blk * KP + t] = s;
    }
}

// Sub-warp variant for KP <= 32: L lanes cooperate on one column, each holding RPT = KP / L consecutive
// rows of a_i in registers, so one warp solves G = 32 / L columns at once (L = 1 for KP <= 8: a whole
// column per thread). A coordinate step then costs the same ~20 instructions per WARP as in the
// warp-per-column kernel above but advances G columns, and every lane does useful FMAs in the
// Gram-correction phase. When there are too few columns to fill the chip (the H update of a small
// matrix: few cells, long held-out lists) the WS = 4 or 2 warps of a CTA share one column group: each
// accumulates the correction over its share of the held-out rows and warp 0 folds the partial sums
// through shared memory before solving. The host picks the largest WS whose grid is still resident
// in one wave.
// The same kernel with mptr == nullptr (no held-out lists, a_i = the Gram) is the PLAIN solver for small column
// counts (a gene shard of the W update on 8 GPUs, pbmc3k): there the thread-per-column kernel has one warp per SM and
// is bound by the latency of its per-coordinate chain (~150 cycles x 3200 steps), while four (eight) lanes per column with
// the blocked sweep need ~100 cycles per block of four coordinates.
#ifndef SGL_MASKED32_MINB
#define SGL_MASKED32_MINB 3
#endif
template <int KP>
struct MaskedSubCfg {
    static constexpr int L = (KP <= 8) ? 1 : KP / 4;  // lanes per column
    static constexpr int RPT = KP / L;                // rows of a_i per lane (4 or 8)
    static constexpr int G = 32 / L;                  // columns per warp
    static constexpr int WARPS = 4;
};

template <int KP, int WS>
__global__ void __launch_bounds__((WS == 2 ? 2 : MaskedSubCfg<KP>::WARPS) * 32, (WS == 2 ? 2 : 1) * SGL_MASKED32_MINB)
nnls_masked_sub_kernel(const float* __restrict__ Bparts, int splits, float* __restrict__ X,
                       const float* __restrict__ gram_f,   // [KP][KP] jitter-free FP32 Gram
                       const float* __restrict__ F,        // gather factor [rows][KP]
                       const int64_t* __restrict__ colptr, const int64_t* __restrict__ mptr,
                       const uint2* __restrict__ mrec, int64_t ncol, int k, float L1, float L2,
                       double* __restrict__ rowsum_part,
                       // corrections computed beforehand on the tensor cores (gramcorr.cuh): gm[(col - first column of
                       // this launch) * KP * KP + i * KP + j], with mptr == nullptr; blk0 = first CTA index of this launch
                       // (a launch covers a column chunk: blk0 * columns-per-CTA is its first column)
                       const float* __restrict__ gm = nullptr, int64_t blk0 = 0) {
    using C = MaskedSubCfg<KP>;
    constexpr int L = C::L, RPT = C::RPT, G = C::G;
    constexpr int WARPS = (WS == 2) ? 2 : C::WARPS;  // WS = 1: four column groups per CTA; WS = 2 / 4: the CTA shares one
    constexpr int CG = WARPS / WS;                   // column groups per CTA
    static_assert(WS == 1 || WS == WARPS, "a column group is owned by one warp or by the whole CTA");
    constexpr int D = (L == 1) ? 4 : 8;        // held-out rows in flight per column
    constexpr int CPL = (L > 1) ? 1 : KP / 4;  // 16-byte chunks of a factor row copied by one lane
    __shared__ double sred[WARPS][KP];
    __shared__ float sacc[(WS > 1) ? RPT * KP * 32 : 1];
    __shared__ __align__(16) float sring[WARPS][D][32 * CPL * 4];  // [warp][stage][column slot][KP]
    __shared__ __align__(8) uint2 sidx[WARPS][2][G][D];            // held-out records of two blocks of D entries
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lig = lane % L;  // lane inside its column's group: owns rows lig*RPT .. lig*RPT+RPT-1
    const int gi = lane / L;   // column slot inside the warp
    const int cgi = warp / WS, wsub = warp % WS;
    const int64_t blk = (int64_t)blockIdx.x + blk0;
    const int64_t col = (blk * CG + cgi) * G + gi;
    const bool in_range = col < ncol;
    const bool solve = in_range && (colptr[col] != colptr[col + 1]);

    // ---- correction G_M = sum over held-out rows of f f^T (reference :460-461), my RPT rows of it ----
    float a[RPT][KP];
#pragma unroll
    for (int c = 0; c < RPT; ++c)
#pragma unroll
        for (int i = 0; i < KP; ++i) a[c][i] = 0.f;
    int64_t mb = 0;
    int my_n = 0;  // my share: held-out entries mb + wsub + WS * t, t < my_n
    if (solve && mptr) {  // mptr == nullptr: no mask at all (plain solve of a few columns, see below)
        mb = mptr[col];
        const int64_t len = mptr[col + 1] - mb;
        my_n = (int)((len - wsub + WS - 1) / WS);
        if (my_n < 0) my_n = 0;
    }
    const int max_n = __reduce_max_sync(0xffffffffu, my_n);
    // The loads of a held-out entry form a dependent index -> factor-row chain with L2 latency on both links, and
    // loads into registers cannot be kept many entries deep (a warp has six scoreboards: waiting for an old load
    // waits for the young ones too). So both links go through shared memory with cp.async (LDGSTS): the factor
    // rows of D entries per column are in flight in a ring of D stages, and the records of block b + 2 (D entries)
    // are copied while block b is consumed. One commit group per entry; out-of-range entries are zero-filled
    // (they add nothing), so the consumer needs no predicates.
    const uint32_t ring = smem_u32(&sring[warp][0][0]);
    const uint32_t iring = smem_u32(&sidx[warp][0][0][0]);
    constexpr uint32_t STAGE_BYTES = 32 * CPL * 16;
    constexpr int IPL = (D + L - 1) / L;  // records copied per lane per block
    auto copy_idx = [&](int blk) {  // records of entries blk*D .. blk*D+D-1 of my column -> sidx[blk & 1][gi][*]
#pragma unroll
        for (int q = 0; q < IPL; ++q) {
            const int e = lig + q * L;
            if (e < D) {
                const int t = blk * D + e;
                const bool ok = t < my_n;
                cp_async8(iring + (uint32_t)((((blk & 1) * G + gi) * D + e) * 8), mrec + (ok ? mb + wsub + (int64_t)WS * t : 0), ok ? 8u : 0u);
            }
        }
    };
    auto issue = [&](int t, int stage) {  // my 16-byte chunk(s) of the factor row of my column's entry t
        const uint32_t row = sidx[warp][(t / D) & 1][gi][t % D].x;
        const bool ok = t < my_n;
        const float* src = F + (int64_t)row * KP + (L > 1 ? lig * 4 : 0);
#pragma unroll
        for (int q = 0; q < CPL; ++q)
            cp_async16(ring + (uint32_t)stage * STAGE_BYTES + (uint32_t)(lane * CPL + q) * 16u, src + 4 * q, ok ? 16u : 0u);
    };
    if (max_n > 0) {  // warp-uniform
        copy_idx(0);
        copy_idx(1);
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
#pragma unroll
        for (int d = 0; d < D; ++d) {
            issue(d, d);
            cp_async_commit();
        }
    }
    for (int tb = 0; tb < max_n; tb += D) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const int t = tb + d;
            if (t < max_n) {  // warp-uniform
                cp_async_wait<D - 1>();  // entry t (and, at d == 0, the records of the next block) has landed
                __syncwarp();
                const float4* st4 = reinterpret_cast<const float4*>(&sring[warp][d][0]);
                float f[KP], mine[RPT];
#pragma unroll
                for (int i4 = 0; i4 < KP / 4; ++i4) {
                    const float4 v = st4[gi * (KP / 4) + i4];  // same address on the L lanes of a group: broadcast
                    f[4 * i4 + 0] = v.x; f[4 * i4 + 1] = v.y; f[4 * i4 + 2] = v.z; f[4 * i4 + 3] = v.w;
                }
                if constexpr (L == 1) {
#pragma unroll
                    for (int c = 0; c < RPT; ++c) mine[c] = f[c];
                } else {
                    const float4 v = st4[lane];  // my RPT == 4 rows of f
                    mine[0] = v.x; mine[1] = v.y; mine[2] = v.z; mine[3] = v.w;
                }
#pragma unroll
                for (int c = 0; c < RPT; ++c)
#pragma unroll
                    for (int i = 0; i < KP; ++i) a[c][i] = fmaf(mine[c], f[i], a[c][i]);
                __syncwarp();  // the stage (and, at d == 0, the record slot of block tb / D) is free again
                issue(t + D, d);
                if (d == 0) copy_idx(tb / D + 2);
                cp_async_commit();
            }
        }
    }
    cp_async_wait<0>();
    if constexpr (WS > 1) {  // fold the WS partial corrections into warp 0, one warp at a time
        for (int w = 1; w < WS; ++w) {
            if (wsub == w) {
#pragma unroll
                for (int c = 0; c < RPT; ++c)
#pragma unroll
                    for (int i = 0; i < KP; ++i) sacc[(c * KP + i) * 32 + lane] = a[c][i];
            }
            __syncthreads();
            if (wsub == 0) {
#pragma unroll
                for (int c = 0; c < RPT; ++c)
#pragma unroll
                    for (int i = 0; i < KP; ++i) a[c][i] += sacc[(c * KP + i) * 32 + lane];
            }
            __syncthreads();
        }
    }

    if (gm != nullptr && solve) {  // G_M from gram_corr_mma_kernel: my RPT rows of it
        const float* src = gm + ((col - blk0 * (CG * G)) * KP + lig * RPT) * (int64_t)KP;
#pragma unroll
        for (int c = 0; c < RPT; ++c)
#pragma unroll
            for (int i4 = 0; i4 < KP / 4; ++i4) {
                const float4 v = *reinterpret_cast<const float4*>(src + c * KP + 4 * i4);
                a[c][4 * i4 + 0] = v.x; a[c][4 * i4 + 1] = v.y; a[c][4 * i4 + 2] = v.z; a[c][4 * i4 + 3] = v.w;
            }
    }

    // ---- right-hand side, warm start, a_i = G - G_M (the 1e-15 jitters cancel, App. A-11) ----
    float b[RPT], x[RPT], inv[RPT];
#pragma unroll
    for (int c = 0; c < RPT; ++c) { b[c] = 0.f; x[c] = 0.f; inv[c] = 0.f; }
    const bool owner_warp = (wsub == 0);
    if (in_range && owner_warp) {
#pragma unroll
        for (int c4 = 0; c4 < RPT / 4; ++c4) {
            for (int s = 0; s < splits; ++s) {
                const float4 v = *reinterpret_cast<const float4*>(Bparts + ((int64_t)s * ncol + col) * KP + lig * RPT + 4 * c4);
                b[4 * c4 + 0] += v.x; b[4 * c4 + 1] += v.y; b[4 * c4 + 2] += v.z; b[4 * c4 + 3] += v.w;
            }
            const float4 v = *reinterpret_cast<const float4*>(X + col * KP + lig * RPT + 4 * c4);
            x[4 * c4 + 0] = v.x; x[4 * c4 + 1] = v.y; x[4 * c4 + 2] = v.z; x[4 * c4 + 3] = v.w;
        }
    }
    bool active = solve && owner_warp;
    if (__any_sync(0xffffffffu, active)) {
#pragma unroll
        for (int c = 0; c < RPT; ++c) {
            const float4* g4 = reinterpret_cast<const float4*>(gram_f + (lig * RPT + c) * KP);
#pragma unroll
            for (int i4 = 0; i4 < KP / 4; ++i4) {
                const float4 v = g4[i4];
                a[c][4 * i4 + 0] = v.x - a[c][4 * i4 + 0];
                a[c][4 * i4 + 1] = v.y - a[c][4 * i4 + 1];
                a[c][4 * i4 + 2] = v.z - a[c][4 * i4 + 2];
                a[c][4 * i4 + 3] = v.w - a[c][4 * i4 + 3];
            }
        }
#pragma unroll
        for (int c = 0; c < RPT; ++c) {  // the diagonal element of my row lig*RPT + c is a[c][lig*RPT + c]: pick it with selects
            float dia = 1.f;
#pragma unroll
            for (int o = 0; o < L; ++o) dia = (lig == o) ? a[c][o * RPT + c] : dia;
            inv[c] = (lig * RPT + c < k) ? 1.0f / dia : 0.f;  // padding coordinates are inert
        }

        // Coordinate loop: at step i every lane runs the branch-free scalar step on its slot i % RPT, the
        // owner lane (i / RPT) of each group broadcasts its multiplier with one width-L shuffle, and every
        // lane applies it to its RPT entries of b. A lane's coordinates are consecutive, so it keeps the
        // reference's running tol (reset to 1 by a clamp event, src/singlet.cpp:240-246) locally and the
        // group's value is rebuilt after the sweep: the lanes from the last one with an event onwards add up.
        auto sweeps = [&](auto has_l2) {
        constexpr bool HAS_L2 = decltype(has_l2)::value;  // L2 == 0 (the default): c0 is the constant -L1
        float tol_g = 1.f;
        const float kf = (float)k;
        for (int sweep = 0; sweep < NNLS_MAX_SWEEPS; ++sweep) {
            active = active && (tol_g / kf > 1e-8f);
            if (!__any_sync(0xffffffffu, active)) break;
            // Critical path of a step: b -> diff -> m -> shuffle -> b, nothing else is computed inside the sweep:
            // m = min(-diff, x) is `clamp ? x : -diff` and x_new = max(x + diff, 0) is `clamp ? 0 : x + diff` in one
            // FMNMX each; c0 = L2 * x - L1 does not depend on b (exact no-ops when L1 / L2 are 0, App. A-5).
            // The owner lane keeps diff and the old x of its coordinates; the tol bookkeeping is replayed from
            // them after the sweep (each coordinate is visited exactly once per sweep).
            float xo[RPT], dv[RPT];
#pragma unroll
            for (int c = 0; c < RPT; ++c) { xo[c] = x[c]; dv[c] = 0.f; }
#pragma unroll
            for (int o = 0; o < L; ++o) {
                const bool mine = active && (lig == o);  // a finished column is frozen: x stays, its b is dead
                if constexpr (L == 1) {
#pragma unroll
                    for (int c = 0; c < RPT; ++c) {
                        if (c < k) {  // uniform
                            const float xi = x[c];
                            const float c0 = HAS_L2 ? fmaf(L2, xi, -L1) : -L1;
                            const float diff = fmaf(b[c], inv[c], c0);
                            const float mult = fminf(-diff, xi);
                            const float xnew = fmaxf(xi + diff, 0.f);
                            x[c] = mine ? xnew : xi;
                            dv[c] = mine ? diff : dv[c];
#pragma unroll
                            for (int cc = 0; cc < RPT; ++cc) b[cc] = fmaf(a[cc][c], mult, b[cc]);
                        }
                    }
                } else {
                    // Phase A: lane o runs its RPT consecutive coordinates on a private copy of its b entries, applying
                    // its own multipliers to the later entries directly, so no shuffle sits between two of its steps.
                    // Phase B: the RPT multipliers are broadcast (the shuffles overlap) and every lane, lane o included,
                    // applies them to b in coordinate order -- the values are bit-identical to the step-by-step order.
                    // A padding coordinate (i >= k, only in the last block) is inert: its b, x, inv and Gram entries are
                    // zero, so diff = -L1, m = 0 and nothing changes; only whole blocks are skipped, so that no branch
                    // separates the shuffles from each other.
                    if (o * RPT < k) {  // uniform
                        float bl[RPT], mo[RPT];
#pragma unroll
                        for (int c = 0; c < RPT; ++c) bl[c] = b[c];
#pragma unroll
                        for (int c = 0; c < RPT; ++c) {
                            const int i = o * RPT + c;
                            const float xi = x[c];
                            const float c0 = HAS_L2 ? fmaf(L2, xi, -L1) : -L1;
                            const float diff = fmaf(bl[c], inv[c], c0);
                            mo[c] = fminf(-diff, xi);
                            const float xnew = fmaxf(xi + diff, 0.f);
                            x[c] = mine ? xnew : xi;
                            dv[c] = mine ? diff : dv[c];
#pragma unroll
                            for (int cc = c + 1; cc < RPT; ++cc) bl[cc] = fmaf(a[cc][i], mo[c], bl[cc]);
                        }
                        float mult[RPT];
#pragma unroll
                        for (int c = 0; c < RPT; ++c) mult[c] = __shfl_sync(0xffffffffu, mo[c], o, L);
#pragma unroll
                        for (int c = 0; c < RPT; ++c)
#pragma unroll
                            for (int cc = 0; cc < RPT; ++cc) b[cc] = fmaf(a[cc][o * RPT + c], mult[c], b[cc]);
                    }
                }
            }
            // replay of the reference's running tol over my (consecutive) coordinates: a clamp event that changes
            // x resets it to 1, every other step adds |diff / (x_new + 1e-15)| (src/singlet.cpp:240-246)
            float ltol = 0.f;
            bool lev = false;
#pragma unroll
            for (int c = 0; c < RPT; ++c) {
                const bool clamp = (-dv[c] > xo[c]);
                const bool clamp_ev = clamp && (xo[c] != 0.f);
                float r;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x[c] + 1e-15f));
                const float term = clamp ? 0.f : fabsf(dv[c] * r);
                lev = lev || clamp_ev;
                ltol = clamp_ev ? 1.f : ltol + term;
            }
            float part = ltol;
            if constexpr (L > 1) {
                const uint32_t bal = __ballot_sync(0xffffffffu, lev);
                const uint32_t gm = (bal >> (gi * L)) & ((1u << L) - 1u);
                const int hl = gm ? (31 - __clz(gm)) : 0;  // last lane of my group with an event
                part = (lig >= hl) ? ltol : 0.f;
#pragma unroll
                for (int o = 1; o < L; o <<= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            }
            if (active) tol_g = part;
        }
        };
        if (L2 == 0.f) sweeps(std::false_type{}); else sweeps(std::true_type{});
    }

    // ---- write back, per-CTA row sums (the local part of scale's d) ----
    if (solve && owner_warp) {
#pragma unroll
        for (int c4 = 0; c4 < RPT / 4; ++c4)
            *reinterpret_cast<float4*>(X + col * KP + lig * RPT + 4 * c4) = make_float4(x[4 * c4], x[4 * c4 + 1], x[4 * c4 + 2], x[4 * c4 + 3]);
    }
#pragma unroll
    for (int c = 0; c < RPT; ++c) {
        double v = (in_range && owner_warp) ? (double)x[c] : 0.0;
#pragma unroll
        for (int o = L; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (gi == 0) sred[warp][lig * RPT + c] = v;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < KP; t += WARPS * 32) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) s += sred[w][t];
        rowsum_part[blk * KP + t] = s;
    }
}

// Padded rank 64, blocked sweep. Same solve as nnls_masked_kernel<64> (warp per column, the lane holds two rows of a_i) with the
// rows of a lane CONSECUTIVE (2 lane, 2 lane + 1) so that, as in the sub-warp kernel, the owner lane o runs the two coordinates
// 2o, 2o + 1 on a private copy of its b entries (phase A), the two multipliers are broadcast by two overlapping shuffles and every
// lane applies them in coordinate order (phase B) -- bit-identical to the step-by-step order, half the shuffle-bound steps.
// Padding coordinates (i >= k) are inert: b = x = inv = 0 gives a zero multiplier.
__global__ void __launch_bounds__(MaskedCfg<64>::WARPS * 32)
nnls_masked64_blocked_kernel(const float* __restrict__ Bparts, int splits, float* __restrict__ X, const float* __restrict__ gram_f,
                             const float* __restrict__ F, const int64_t* __restrict__ colptr, const int64_t* __restrict__ mptr,
                             const uint2* __restrict__ mrec, int64_t ncol, int k, float L1, float L2,
                             double* __restrict__ rowsum_part, const float* __restrict__ gm, int64_t blk0) {
    constexpr int KP = 64, WARPS = MaskedCfg<64>::WARPS;
    __shared__ double sred[WARPS][KP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t blk = (int64_t)blockIdx.x + blk0;
    const int64_t col = blk * WARPS + warp;
    const bool in_range = col < ncol;
    const bool solve = in_range && (colptr[col] != colptr[col + 1]);
    const int r0 = 2 * lane;  // my rows: r0, r0 + 1
    float a[2][KP];
    float b[2] = {0.f, 0.f}, x[2] = {0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < KP; ++i) a[c][i] = 0.f;
    if (in_range) {
        for (int s = 0; s < splits; ++s) {
            const float2 v = *reinterpret_cast<const float2*>(Bparts + ((int64_t)s * ncol + col) * KP + r0);
            b[0] += v.x;
            b[1] += v.y;
        }
        const float2 v = *reinterpret_cast<const float2*>(X + col * KP + r0);
        x[0] = v.x;
        x[1] = v.y;
    }
    if (solve) {
        if (gm != nullptr) {
            const float* src = gm + (col - blk0 * WARPS) * (int64_t)(KP * KP) + (int64_t)r0 * KP;
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int i4 = 0; i4 < KP / 4; ++i4) {
                    const float4 v = *reinterpret_cast<const float4*>(src + c * KP + 4 * i4);
                    a[c][4 * i4 + 0] = v.x; a[c][4 * i4 + 1] = v.y; a[c][4 * i4 + 2] = v.z; a[c][4 * i4 + 3] = v.w;
                }
        } else if (mptr != nullptr) {  // FP32 correction inside the solver (SGL_GRAMCORR=ffma)
            const int64_t mb = mptr[col], me = mptr[col + 1];
            for (int64_t p = mb; p < me; ++p) {
                const float* fr = F + (int64_t)mrec[p].x * KP;
                const float2 mine = *reinterpret_cast<const float2*>(fr + r0);
                const float4* fr4 = reinterpret_cast<const float4*>(fr);
#pragma unroll
                for (int i4 = 0; i4 < KP / 4; ++i4) {
                    const float4 f4 = fr4[i4];  // uniform address: broadcast
                    a[0][4 * i4 + 0] = fmaf(mine.x, f4.x, a[0][4 * i4 + 0]); a[0][4 * i4 + 1] = fmaf(mine.x, f4.y, a[0][4 * i4 + 1]);
                    a[0][4 * i4 + 2] = fmaf(mine.x, f4.z, a[0][4 * i4 + 2]); a[0][4 * i4 + 3] = fmaf(mine.x, f4.w, a[0][4 * i4 + 3]);
                    a[1][4 * i4 + 0] = fmaf(mine.y, f4.x, a[1][4 * i4 + 0]); a[1][4 * i4 + 1] = fmaf(mine.y, f4.y, a[1][4 * i4 + 1]);
                    a[1][4 * i4 + 2] = fmaf(mine.y, f4.z, a[1][4 * i4 + 2]); a[1][4 * i4 + 3] = fmaf(mine.y, f4.w, a[1][4 * i4 + 3]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int i4 = 0; i4 < KP / 4; ++i4) {
                const float4 v = *reinterpret_cast<const float4*>(gram_f + (r0 + c) * KP + 4 * i4);
                a[c][4 * i4 + 0] = v.x - a[c][4 * i4 + 0]; a[c][4 * i4 + 1] = v.y - a[c][4 * i4 + 1];
                a[c][4 * i4 + 2] = v.z - a[c][4 * i4 + 2]; a[c][4 * i4 + 3] = v.w - a[c][4 * i4 + 3];
            }
        // my diagonal elements a[c][r0 + c]: picked with selects (the index depends on the lane)
        float inv[2] = {0.f, 0.f};
#pragma unroll
        for (int o = 0; o < 32; ++o) {
            if (lane == o) {
                inv[0] = (2 * o < k) ? 1.0f / a[0][2 * o] : 0.f;
                inv[1] = (2 * o + 1 < k) ? 1.0f / a[1][2 * o + 1] : 0.f;
            }
        }
        float tol = 1.f;
        const float kf = (float)k;
        for (int sweep = 0; sweep < NNLS_MAX_SWEEPS && (tol / kf > 1e-8f); ++sweep) {
            float term[2] = {0.f, 0.f};
            bool ev[2] = {false, false};
#pragma unroll
            for (int o = 0; o < 32; ++o) {
                if (2 * o < k) {  // uniform: a block with at least one real coordinate
                    // phase A (every lane on its own data; only lane o's results are used)
                    float x0 = x[0], x1 = x[1], t0 = 0.f, t1 = 0.f;
                    const float m0 = cd_step_nb(b[0], inv[0], x0, L1, L2, t0);
                    const float bl1 = fmaf(a[1][2 * o], m0, b[1]);
                    const float m1 = cd_step_nb(bl1, inv[1], x1, L1, L2, t1);
                    const float mult0 = __shfl_sync(0xffffffffu, m0, o);
                    const float mult1 = __shfl_sync(0xffffffffu, m1, o);
                    if (lane == o) {
                        const bool e0 = (x0 == 0.f) && (x[0] != 0.f) && (t0 == 1.f);
                        const bool e1 = (x1 == 0.f) && (x[1] != 0.f) && (t1 == 1.f);
                        x[0] = x0; x[1] = x1;
                        ev[0] = e0; ev[1] = e1;
                        term[0] = e0 ? 0.f : t0;
                        term[1] = e1 ? 0.f : t1;
                    }
                    // phase B, in coordinate order
                    b[0] = fmaf(a[0][2 * o], mult0, b[0]);
                    b[1] = fmaf(a[1][2 * o], mult0, b[1]);
                    b[0] = fmaf(a[0][2 * o + 1], mult1, b[0]);
                    b[1] = fmaf(a[1][2 * o + 1], mult1, b[1]);
                }
            }
            // rebuild tol: the last clamp event in coordinate order (coordinate of slot c on this lane: 2 lane + c)
            const uint32_t bal0 = __ballot_sync(0xffffffffu, ev[0]), bal1 = __ballot_sync(0xffffffffu, ev[1]);
            const uint32_t any = bal0 | bal1;
            int last = -1;
            if (any) {
                const int hl = 31 - __clz(any);
                last = 2 * hl + (((bal1 >> hl) & 1u) ? 1 : 0);
            }
            float part = ((r0 > last) ? term[0] : 0.f) + ((r0 + 1 > last) ? term[1] : 0.f);
            tol = warp_sum(part) + (last >= 0 ? 1.f : 0.f);
        }
        *reinterpret_cast<float2*>(X + col * KP + r0) = make_float2(x[0], x[1]);
    }
    sred[warp][r0] = in_range ? (double)x[0] : 0.0;
    sred[warp][r0 + 1] = in_range ? (double)x[1] : 0.0;
    __syncthreads();
    for (int t = threadIdx.x; t < KP; t += WARPS * 32) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) s += sred[w][t];
        rowsum_part[blk * KP + t] = s;
    }
}

// generic masked fallback for KP = 128: a_i, b, x in shared memory (one warp per CTA, padded rows)
__global__ void __launch_bounds__(32)
nnls_masked_big_kernel(const float* __restrict__ Bparts, int splits, float* __restrict__ X,
                       const float* __restrict__ gram_f, const float* __restrict__ F,
                       const int64_t* __restrict__ colptr, const int64_t* __restrict__ mptr,
                       const uint2* __restrict__ mrec, int64_t ncol, int k, int KP, float L1, float L2,
                       double* __restrict__ rowsum_part) {
    extern __shared__ float sm[];
    const int LD = KP + 1;
    float* sa = sm;                      // [KP][LD]
    float* sb = sm + (size_t)KP * LD;    // [KP]
    float* sx = sb + KP;                 // [KP]
    float* sinv = sx + KP;               // [KP]
    const int lane = threadIdx.x;
    const int64_t col = blockIdx.x;
    const bool solve = colptr[col] != colptr[col + 1];
    for (int r = lane; r < KP; r += 32) {
        float bj = 0.f;
        for (int s = 0; s < splits; ++s) bj += Bparts[((int64_t)s * ncol + col) * KP + r];
        sb[r] = bj;
        sx[r] = X[col * KP + r];
        for (int i = 0; i < KP; ++i) sa[r * LD + i] = 0.f;
    }
    __syncwarp();
    if (solve) {
        for (int64_t p = mptr[col]; p < mptr[col + 1]; ++p) {
            const float* fr = F + (int64_t)mrec[p].x * KP;
            for (int r = lane; r < k; r += 32) {
                const float mine = fr[r];
                for (int i = 0; i < k; ++i) sa[r * LD + i] = fmaf(mine, fr[i], sa[r * LD + i]);
            }
        }
        for (int r = lane; r < KP; r += 32) {
            for (int i = 0; i < KP; ++i) sa[r * LD + i] = gram_f[r * KP + i] - sa[r * LD + i];
            sinv[r] = 1.0f / sa[r * LD + r];
        }
        __syncwarp();
        float tol = 1.f;
        const float kf = (float)k;
        for (int sweep = 0; sweep < NNLS_MAX_SWEEPS && (tol / kf > 1e-8f); ++sweep) {
            tol = 0.f;
            for (int i = 0; i < k; ++i) {
                float xi = sx[i];
                const float delta = cd_step(sb[i], sinv[i], xi, L1, L2, tol);
                __syncwarp();
                if (lane == 0) sx[i] = xi;
                for (int r = lane; r < k; r += 32) sb[r] = fmaf(-sa[r * LD + i], delta, sb[r]);
                __syncwarp();
            }
        }
    }
    for (int r = lane; r < KP; r += 32) {
        if (solve) X[col * KP + r] = sx[r];
        rowsum_part[(int64_t)blockIdx.x * KP + r] = (double)sx[r];
    }
}

// deterministic fixed-order reduction of per-CTA double partials: out[j] = sum_p part[p][j]
__global__ void reduce_partials_kernel(const double* __restrict__ part, int64_t n_parts, int width,
                                       double* __restrict__ out) {
    __shared__ double sm[256];
    const int j = blockIdx.x;  // one CTA per output element
    double s = 0.0;
    for (int64_t p = threadIdx.x; p < n_parts; p += blockDim.x) s += part[p * width + j];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[j] = sm[0];
}

}  // namespace sgl
