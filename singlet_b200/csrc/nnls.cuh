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
__constant__ __align__(16) float c_gram[64 * 64];
__constant__ float c_inv_diag[64];

// Persistent kernel with lane refill. Every lane solves NCL = 2 columns at a time (two independent
// dependency chains per thread, and every Gram row fetched once feeds both). Columns converge after
// very different numbers of sweeps (from ~30 up to the 100-sweep cap), so a slot whose column is done
// writes it back and claims the next unsolved column from a global counter instead of idling until the
// slowest lane of its warp finishes. Columns are independent, so the result does not depend on which
// lane solves which column. Row sums are taken afterwards by rowsum_partial_kernel.
// NT threads per CTA, NCL columns per lane: <THREADS, 2> for large column counts; <32, 1> when there are
// too few columns to fill the chip otherwise (e.g. a gene shard of the W update on 8 GPUs).
template <int KP, int NT, int NCL>
__global__ void __launch_bounds__(NT, (NT == 32) ? 8 : NnlsCfg<KP>::MIN_CTAS)
nnls_cols_kernel(const float* __restrict__ Bparts,  // [splits][ncol][KP]
                 int splits, float* __restrict__ X,  // [ncol][KP] warm start in / solution out
                 const int64_t* __restrict__ colptr, int64_t ncol, int k, float L1, float L2,
                 unsigned long long* __restrict__ next_col,  // global work counter (zeroed by the host)
                 unsigned long long* __restrict__ stats)     // optional: [0] += sweeps, [1] += columns solved
{
    __shared__ float sx[NCL * KP * NT];  // sx[(s * KP + i) * NT + tid]

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
                    if (c >= ncol) {
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
                   int64_t ncol, int k, float L1, float L2, double* __restrict__ rowsum_part)
{
    constexpr int RPL = MaskedCfg<KP>::RPL;
    constexpr int WARPS = MaskedCfg<KP>::WARPS;
    __shared__ double sred[WARPS][KP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t col = (int64_t)blockIdx.x * WARPS + warp;
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
        const int64_t mb = mptr[col], me = mptr[col + 1];
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
        rowsum_part[(int64_t)blockIdx.x * KP + t] = s;
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
