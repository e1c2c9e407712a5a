// spmm.cuh -- sparse right-hand-side product  B[:, c] = sum_{nz (r, v) in X[:, c]} v * F[r, :]
// (the `b += it.value() * w.col(it.row())` loop of predict, reference src/singlet.cpp:341-343).
//
// Layout. X is column-compressed with 8-byte records {int32 row, float value}; the gather operand F
// is float [rows][KP]. F does not fit in shared memory (k x m = 3.84 MB, k x n = 128 MB at the
// headline config), so rows are cut into tiles of `rb_rows` rows; a precomputed table
// tileptr[t][col] gives, for every column, where tile t starts inside the column's record range
// (rows are ascending within a column, so every (column, tile) sub-range is contiguous).
//
// One CTA owns WARPS*NC columns and walks a range of row tiles. Per tile the F tile
// (rb_rows x KP floats, contiguous in memory) is staged into shared memory by ONE bulk async copy
// (TMA, cp.async.bulk -> SASS UBLKCP) into a 2-deep ring guarded by mbarriers, so the load of tile
// t+1 overlaps the FMAs of tile t. Each warp keeps the accumulators of its NC columns in registers
// for the whole walk: LPN = min(8, KP/4) lanes cooperate on one non-zero (each lane owns FPL = KP/LPN
// factors as float4s), so one LDS.128 per lane fetches 32/LPN gathered rows with every 8-lane
// phase reading one contiguous 128-byte row (bank-conflict free), and the record {row, value} is
// fetched once per lane group straight from the HBM stream (read-once, L1 no-allocate).
//
// Roofline note: the HBM stream is 8 B per non-zero, the shared-memory gather is
// KP*4 B per non-zero; at KP = 32 the 128 B/clk/SM crossbar caps the kernel at ~1 non-zero/clk/SM.
#pragma once
#include "common.cuh"

namespace sgl {

template <int KP>
struct SpmmCfg {
    static constexpr int LPN = (KP / 4 < 8) ? (KP / 4) : 8;  // lanes per non-zero
    static constexpr int FPL = KP / LPN;                     // factors per lane (4, 8 or 16)
    static constexpr int NV = FPL / 4;                       // 16-byte loads per lane per non-zero
    static constexpr int SLOTS = 32 / LPN;                   // non-zeros per warp step
    static constexpr int NC = (64 / FPL);                    // columns per warp (64 accumulator regs)
    static constexpr int WARPS = 16;
    static constexpr int COLS_PER_CTA = WARPS * NC;
    // record loads issued back to back before the first use: one group covers ~48 non-zeros, i.e. a
    // typical (column, tile) sub-range, so each sub-range exposes one memory latency
    static constexpr int UNR = (48 / SLOTS) < 2 ? 2 : (48 / SLOTS);
};

// rows per staged tile: two stages must fit in 227 KB of shared memory
static inline int spmm_tile_rows(int kp) {
    const int budget = (227 * 1024 - 1024) / 2;  // bytes per stage
    int rows = budget / (kp * 4);
    rows &= ~7;
    // 800 rows at KP = 32: at 5 % density a (column, tile) sub-range then averages 40 records, so
    // one group of UNR * SLOTS = 48 record loads covers ~90 % of the sub-ranges in a single pass
    const int cap = 800 * 32 / kp;
    return rows < cap ? rows : cap;
}

// packed FP32 pair FMA (sm_100 FFMA2; SASS takes the scalar operand as a broadcast .F32):
//   acc.{lo,hi} += w.{lo,hi} * v
__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long w, float v) {
    unsigned long long vv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(vv) : "f"(v));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(w), "l"(vv));
}
__device__ __forceinline__ void lds_2x64(uint32_t addr, unsigned long long& a, unsigned long long& b) {
    asm("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}

template <int KP>
__global__ void __launch_bounds__(SpmmCfg<KP>::WARPS * 32, 1)
spmm_tiles_kernel(const uint2* __restrict__ rec, const int64_t* __restrict__ colptr,
                  const int32_t* __restrict__ tileptr,  // [n_tiles + 1][ncol_pad]
                  int64_t ncol, int64_t ncol_pad, int64_t nrow, int rb_rows, int n_tiles, int tiles_per_split,
                  const float* __restrict__ F,  // [nrow][KP]
                  float* __restrict__ Bout,     // [splits][ncol][KP]
                  int dbg)                      // debug knobs (SGL_SPMM_DEBUG): 1 = no L2 prefetch, 2 = fake records
{
    using C = SpmmCfg<KP>;
    constexpr int UNR = C::UNR;
    constexpr int GROUP = UNR * C::SLOTS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage0 = reinterpret_cast<float*>(smem_raw);
    const size_t stage_floats = (size_t)rb_rows * KP;
    float* stage1 = stage0 + stage_floats;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage1 + stage_floats);  // 2 "full" barriers

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = lane % C::LPN;  // factor chunk owned by this lane
    const int g = lane / C::LPN;  // non-zero slot inside a warp step
    const int64_t col0 = (int64_t)blockIdx.x * C::COLS_PER_CTA + (int64_t)warp * C::NC;
    const int t_begin = blockIdx.y * tiles_per_split;
    const int t_end = min(n_tiles, t_begin + tiles_per_split);

    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
    }
    __syncthreads();

    auto issue = [&](int t) {  // thread 0 only
        const int s = (t - t_begin) & 1;
        const int64_t r0 = (int64_t)t * rb_rows;
        const int64_t rows = min((int64_t)rb_rows, nrow - r0);
        const uint32_t bytes = (uint32_t)(rows * KP * 4);
        mbar_arrive_expect_tx(&bars[s], bytes);
        tma_load_1d(s ? stage1 : stage0, F + r0 * KP, bytes, &bars[s]);
    };
    if (threadIdx.x == 0 && t_begin < t_end) {
        issue(t_begin);
        if (t_begin + 1 < t_end) issue(t_begin + 1);
    }

    // accumulators: NC columns x FPL factors, as packed FP32 pairs
    unsigned long long acc[C::NC][C::FPL / 2];
#pragma unroll
    for (int j = 0; j < C::NC; ++j)
#pragma unroll
        for (int f = 0; f < C::FPL / 2; ++f) acc[j][f] = 0ull;

    // this lane's column (for the coalesced tile-pointer loads) and its record base
    const int64_t my_col = col0 + lane;
    const bool my_col_ok = (lane < C::NC) && (my_col < ncol);
    const int64_t my_base = my_col_ok ? colptr[my_col] : 0;

    // tileptr[t][col] is both the end of tile t-1 and the start of tile t. Each lane (< NC) keeps the
    // boundaries of its column three tiles deep so that, at the top of tile t, it can (a) fetch the
    // boundary needed two tiles later and (b) ask L2 to prefetch the column's records of tile t+1
    // with one bulk prefetch (the record loads of the next tile then hit L2 instead of HBM).
    int32_t p0 = 0, p1 = 0, p2 = 0;  // tileptr[t], tileptr[t+1], tileptr[t+2]
    if (my_col_ok && t_begin < t_end) {
        p0 = tileptr[(int64_t)t_begin * ncol_pad + my_col];
        p1 = tileptr[(int64_t)(t_begin + 1) * ncol_pad + my_col];
        p2 = (t_begin + 1 < t_end) ? tileptr[(int64_t)(t_begin + 2) * ncol_pad + my_col] : p1;
        l2_prefetch_records(rec + my_base + p0, p1 - p0);
    }

    uint32_t phase0 = 0, phase1 = 0;
    for (int t = t_begin; t < t_end; ++t) {
        const int s = (t - t_begin) & 1;
        // per-lane (lanes < NC): first record of this column in tile t, and how many
        const int64_t start_l = my_base + p0;
        const int32_t cnt_l = p1 - p0;
        int32_t p3 = p2;
        if (my_col_ok) {
            if (t + 2 < t_end) p3 = tileptr[(int64_t)(t + 3) * ncol_pad + my_col];
            if (t + 1 < t_end && !(dbg & 1)) l2_prefetch_records(rec + my_base + p1, p2 - p1);
        }
        mbar_wait(&bars[s], s ? phase1 : phase0);
        if (s) phase1 ^= 1u; else phase0 ^= 1u;
        // shared address of (row 0 of the matrix, chunk q) as seen through this tile: adding
        // row * KP * 4 for any row of the tile lands inside the stage (32-bit wrap-around is fine)
        const uint32_t tile_q = smem_u32(s ? stage1 : stage0) + (uint32_t)q * 16u - (uint32_t)(t * rb_rows) * (uint32_t)(KP * 4);

        const uint32_t pad_row = (uint32_t)(t * rb_rows);  // a valid row of this tile
#pragma unroll
        for (int j = 0; j < C::NC; ++j) {
            const int64_t start = __shfl_sync(0xffffffffu, start_l, j);
            const int32_t n = __shfl_sync(0xffffffffu, cnt_l, j);
            const uint2* lp = rec + start + g;  // this lane's first record
            int32_t rem = n - g;                // records left for this lane's slot (may be <= 0)
            for (int32_t done = 0; done < n; done += GROUP) {  // warp-uniform
                // slots past the end of the sub-range gather a valid row with value 0 (adds +0)
                uint2 r[UNR];
#pragma unroll
                for (int u = 0; u < UNR; ++u)
                    r[u] = (rem > u * C::SLOTS) ? ((dbg & 2) ? make_uint2(pad_row + (uint32_t)((rem * 37 + u * 11) & 511), 0x3f800000u) : ldg_stream_u2(lp + u * C::SLOTS)) : make_uint2(pad_row, 0u);
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    if (done + u * C::SLOTS >= n) break;  // warp-uniform: no step past the sub-range
                    const float v = __uint_as_float(r[u].y);
                    const uint32_t a = tile_q + r[u].x * (uint32_t)(KP * 4);
#pragma unroll
                    for (int c4 = 0; c4 < C::NV; ++c4) {
                        unsigned long long w01, w23;
                        lds_2x64(a + (uint32_t)(c4 * C::LPN * 16), w01, w23);
                        ffma2(acc[j][2 * c4 + 0], w01, v);
                        ffma2(acc[j][2 * c4 + 1], w23, v);
                    }
                }
                lp += GROUP;
                rem -= GROUP;
            }
        }
        p0 = p1; p1 = p2; p2 = p3;
        __syncthreads();  // every warp is done with stage s
        if (threadIdx.x == 0 && t + 2 < t_end) issue(t + 2);
    }

    // fold the SLOTS partial sums (lanes with equal q) and store
#pragma unroll
    for (int j = 0; j < C::NC; ++j) {
        float out[C::FPL];
#pragma unroll
        for (int f = 0; f < C::FPL / 2; ++f) {
            float lo = __uint_as_float((uint32_t)(acc[j][f] & 0xffffffffull));
            float hi = __uint_as_float((uint32_t)(acc[j][f] >> 32));
#pragma unroll
            for (int o = C::LPN; o < 32; o <<= 1) {
                lo += __shfl_xor_sync(0xffffffffu, lo, o);
                hi += __shfl_xor_sync(0xffffffffu, hi, o);
            }
            out[2 * f] = lo;
            out[2 * f + 1] = hi;
        }
        const int64_t col = col0 + j;
        if (g == 0 && col < ncol) {
            float4* dst = reinterpret_cast<float4*>(Bout + ((int64_t)blockIdx.y * ncol + col) * KP) + q;
#pragma unroll
            for (int u = 0; u < C::NV; ++u)
                dst[u * C::LPN] = make_float4(out[4 * u + 0], out[4 * u + 1], out[4 * u + 2], out[4 * u + 3]);
        }
    }
}

// tile index: tileptr[t][col] = first record of column `col` (relative to colptr[col]) whose row is
// >= t * rb_rows. One thread per (col, t).
__global__ void build_tileptr_kernel(const uint2* __restrict__ rec, const int64_t* __restrict__ colptr, int64_t ncol,
                                     int64_t ncol_pad, int rb_rows, int n_tiles, int32_t* __restrict__ tileptr) {
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (col >= ncol_pad) return;
    int32_t out = 0;
    if (col < ncol) {
        const int64_t b = colptr[col], e = colptr[col + 1];
        if (t >= n_tiles) {
            out = (int32_t)(e - b);
        } else {
            const int32_t target = t * rb_rows;
            int64_t lo = b, hi = e;  // first index with row >= target
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if ((int32_t)rec[mid].x < target) lo = mid + 1; else hi = mid;
            }
            out = (int32_t)(lo - b);
        }
    }
    tileptr[(int64_t)t * ncol_pad + col] = out;
}

}  // namespace sgl
