// spmm.cuh -- sparse right-hand-side product  B[:, c] = sum_{nz (r, v) in X[:, c]} v * F[r, :]
// (the `b += it.value() * w.col(it.row())` loop of predict, reference src/singlet.cpp:341-343).
//
// HBM layout ("warp streams"). The gather operand F (float [rows][KP]) does not fit in shared memory
// (k x m = 3.84 MB, k x n = 128 MB at the headline config), so rows are cut into tiles of rb_rows
// rows that are staged through shared memory, and the non-zeros are stored IN THE ORDER THE KERNEL
// CONSUMES THEM: columns are grouped NC at a time (one group = one warp); the stream of a group is
//     for tile t: for column j of the group: the records {int32 row, float value} of X[:, col] whose
//     row falls in tile t, padded with {first row of tile t, 0.0f} to a multiple of PAD records.
// goff[group][t] is the stream position of (group, t). A warp therefore reads ONE contiguous array
// front to back: fully coalesced, no per-record predicates, and trivially prefetchable. Which columns
// form a group is given by perm[group * NC + j] (-1 = no column): the identity for evenly filled
// matrices, a snake deal of the columns sorted by their non-zero count when the counts are skewed
// (genes of real data), so that the warps of a CTA -- and the CTAs -- finish together.
//
// Kernel. One CTA = 16 warps = 16 column groups; it walks a range of row tiles. Per tile the F tile
// (rb_rows x KP floats, contiguous) is staged by ONE bulk async copy (TMA: cp.async.bulk -> UBLKCP)
// into a 2-deep ring guarded by mbarriers, so the load of tile t+1 overlaps the FMAs of tile t. Each
// warp keeps a private ring of RC chunks (64 records each) in shared memory that it fills with
// cp.async (LDGSTS) RC-1 chunks ahead of its read position -- the HBM latency of the record stream
// never reaches the math -- and keeps the accumulators of its NC columns in registers for the whole
// walk. LPN = min(8, KP/4) lanes cooperate on one non-zero (each lane owns FPL = KP/LPN factors as
// 16-byte chunks), so one LDS.128 fetches 32/LPN gathered rows with every 8-lane phase reading one
// contiguous 128-byte row (bank-conflict free); products run on packed FFMA2.
//
// Roofline note: HBM sees 8 B per non-zero (+ padding), the shared-memory crossbar sees
// KP*4 B of gather per non-zero; at KP = 32 the 128 B/clk/SM crossbar caps the kernel at about one
// non-zero per clock per SM.
#pragma once
#include "common.cuh"

namespace sgl {

template <int KP>
struct SpmmCfg {
    static constexpr int LPN = (KP / 4 < 8) ? (KP / 4) : 8;  // lanes per non-zero
    static constexpr int FPL = KP / LPN;                     // factors per lane (4, 8 or 16)
    static constexpr int NV = FPL / 4;                       // 16-byte gathers per lane per non-zero
    static constexpr int SLOTS = 32 / LPN;                   // non-zeros per warp step
    static constexpr int NC = (64 / FPL);                    // columns per warp (64 accumulator regs)
    static constexpr int WARPS = 16;
    static constexpr int COLS_PER_CTA = WARPS * NC;
    static constexpr int DS = (KP <= 32) ? 2 : 1;            // warp steps per record fetch (LDS.128 / LDS.64)
    static constexpr int PAD = DS * SLOTS;                   // sub-ranges are padded to this many records
    static constexpr int CHUNK = 64;                         // records per cp.async chunk (512 B)
    static constexpr int RC = 4;                             // ring chunks per warp (2 KB)
    static constexpr int RING_BYTES = WARPS * RC * CHUNK * 8;
};
constexpr int SPMM_RING_BYTES = 16 * 4 * 64 * 8;

// rows per staged F tile: two stages + the record rings must fit in 227 KB of shared memory
static inline int spmm_tile_rows(int kp) {
    const int budget = (227 * 1024 - 1024 - SPMM_RING_BYTES) / 2;  // bytes per stage
    int rows = budget / (kp * 4);
    rows &= ~7;
    return rows;
}

__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long w, float v) {
    unsigned long long vv;  // SASS: FFMA2 takes the scalar as a broadcast .F32 operand, no move needed
    asm("mov.b64 %0, {%1, %1};" : "=l"(vv) : "f"(v));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(w), "l"(vv));
}
__device__ __forceinline__ void lds_2x64(uint32_t addr, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ uint2 lds_u2(uint32_t addr) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr));
    return r;
}
template <int KP>
__global__ void __launch_bounds__(SpmmCfg<KP>::WARPS * 32, 1)
spmm_stream_kernel(const uint2* __restrict__ stream,     // warp streams (see header)
                   const int64_t* __restrict__ goff,     // [n_groups][n_tiles + 1]
                   const int32_t* __restrict__ tileptr,  // [n_tiles + 1][ncol_pad]
                   const int32_t* __restrict__ perm,     // [n_groups * NC] column of every group slot, -1 = none
                   int64_t ncol, int64_t ncol_pad, int64_t nrow, int rb_rows, int n_tiles, int tiles_per_split,
                   const float* __restrict__ F,  // [nrow][KP]
                   float* __restrict__ Bout)     // [splits][ncol][KP]
{
    using C = SpmmCfg<KP>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage0 = reinterpret_cast<float*>(smem_raw);
    const size_t stage_floats = (size_t)rb_rows * KP;
    float* stage1 = stage0 + stage_floats;
    unsigned char* ring_all = reinterpret_cast<unsigned char*>(stage1 + stage_floats);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring_all + C::RING_BYTES);  // 2 "full" barriers

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = lane % C::LPN;  // factor chunk owned by this lane
    const int g = lane / C::LPN;  // non-zero slot inside a warp step
    const int64_t group = (int64_t)blockIdx.x * C::WARPS + warp;
    const int64_t col0 = group * C::NC;
    const int t_begin = blockIdx.y * tiles_per_split;
    const int t_end = min(n_tiles, t_begin + tiles_per_split);
    const bool group_ok = col0 < ncol;

    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
    }
    __syncthreads();

    auto issue = [&](int t) {  // thread 0 only: stage F tile t
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

    // ---- this warp's record stream and ring ----
    int64_t s_begin = 0, s_total = 0;
    if (group_ok && t_begin < t_end) {
        s_begin = goff[group * (n_tiles + 1) + t_begin];
        s_total = goff[group * (n_tiles + 1) + t_end] - s_begin;  // records, multiple of PAD
    }
    const uint2* sp = stream + s_begin;
    const int64_t n_chunks = (s_total + C::CHUNK - 1) / C::CHUNK;
    const uint32_t ring = smem_u32(ring_all) + (uint32_t)warp * (C::RC * C::CHUNK * 8);
    auto fill = [&](int64_t c) {  // whole warp: chunk c -> ring slot c % RC (2 records per lane)
        if (c < n_chunks) {
            const int64_t r = c * C::CHUNK + 2 * lane;
            const bool in = r < s_total;  // s_total and r are even: a lane's pair is in or out as a whole
            cp_async16(ring + (uint32_t)((c % C::RC) * C::CHUNK + 2 * lane) * 8u, sp + (in ? r : 0), in ? 16u : 0u);
        }
        cp_async_commit();  // always commit so that the group count stays uniform
    };
#pragma unroll
    for (int c = 0; c < C::RC; ++c) fill(c);
    cp_async_wait<C::RC - 1>();  // chunk 0 has landed
    __syncwarp();

    // accumulators: NC columns x FPL factors, as packed FP32 pairs
    unsigned long long acc[C::NC][C::FPL / 2];
#pragma unroll
    for (int j = 0; j < C::NC; ++j)
#pragma unroll
        for (int f = 0; f < C::FPL / 2; ++f) acc[j][f] = 0ull;

    // per-column record counts of a tile come from the tile index (lanes < NC hold one column each)
    const int64_t my_col = (lane < C::NC && group_ok) ? (int64_t)perm[col0 + lane] : -1;
    const bool my_col_ok = my_col >= 0;
    int32_t p0 = 0, p1 = 0;  // tileptr[t], tileptr[t+1]
    if (my_col_ok && t_begin < t_end) {
        p0 = tileptr[(int64_t)t_begin * ncol_pad + my_col];
        p1 = tileptr[(int64_t)(t_begin + 1) * ncol_pad + my_col];
    }

    uint32_t pos = 0;  // records consumed so far (multiple of PAD); ring offset = pos % (RC*CHUNK)
    auto ring_read = [&](uint32_t at) {  // this lane group's record(s) of the fetch step at `at`
        const uint32_t roff = ring + ((at & (C::RC * C::CHUNK - 1)) + (uint32_t)(C::DS * g)) * 8u;
        if constexpr (C::DS == 2) {
            return lds_u4(roff);  // {row0, val0, row1, val1}
        } else {
            const uint2 r = lds_u2(roff);
            return make_uint4(r.x, r.y, 0u, 0u);
        }
    };
    uint4 rr = ring_read(0);
    uint32_t phase0 = 0, phase1 = 0;
    for (int t = t_begin; t < t_end; ++t) {
        const int s = (t - t_begin) & 1;
        const int32_t steps_l = (p1 - p0 + C::PAD - 1) / C::PAD;  // fetch steps of my column in this tile
        int32_t p2 = p1;
        if (my_col_ok && t + 1 < t_end) p2 = tileptr[(int64_t)(t + 2) * ncol_pad + my_col];
        mbar_wait(&bars[s], s ? phase1 : phase0);
        if (s) phase1 ^= 1u; else phase0 ^= 1u;
        // shared address of (row 0 of the matrix, chunk q) as seen through this tile: adding
        // row * KP * 4 for any row of the tile lands inside the stage (32-bit wrap-around is fine)
        const uint32_t tile_q = smem_u32(s ? stage1 : stage0) + (uint32_t)q * 16u - (uint32_t)(t * rb_rows) * (uint32_t)(KP * 4);

#pragma unroll
        for (int j = 0; j < C::NC; ++j) {
            const int32_t steps = __shfl_sync(0xffffffffu, steps_l, j);
            for (int32_t st = 0; st < steps; ++st) {  // warp-uniform
                const uint4 cur = rr;
                // advance, make the chunk under the new position readable, and fetch its records
                // BEFORE the math of the current ones (software pipelining of the ring read)
                pos += C::PAD;
                if ((pos & (C::CHUNK - 1)) == 0) {  // entering the next chunk: refill the slot just freed
                    __syncwarp();
                    fill((int64_t)(pos / C::CHUNK) + C::RC - 1);
                    cp_async_wait<C::RC - 1>();
                    __syncwarp();
                }
                rr = ring_read(pos);
                const uint32_t a0 = tile_q + cur.x * (uint32_t)(KP * 4);
                if constexpr (C::DS == 2) {
                    const uint32_t a1 = tile_q + cur.z * (uint32_t)(KP * 4);
#pragma unroll
                    for (int c4 = 0; c4 < C::NV; ++c4) {
                        unsigned long long w01, w23, x01, x23;
                        lds_2x64(a0 + (uint32_t)(c4 * C::LPN * 16), w01, w23);
                        lds_2x64(a1 + (uint32_t)(c4 * C::LPN * 16), x01, x23);
                        ffma2(acc[j][2 * c4 + 0], w01, __uint_as_float(cur.y));
                        ffma2(acc[j][2 * c4 + 1], w23, __uint_as_float(cur.y));
                        ffma2(acc[j][2 * c4 + 0], x01, __uint_as_float(cur.w));
                        ffma2(acc[j][2 * c4 + 1], x23, __uint_as_float(cur.w));
                    }
                } else {
#pragma unroll
                    for (int c4 = 0; c4 < C::NV; ++c4) {
                        unsigned long long w01, w23;
                        lds_2x64(a0 + (uint32_t)(c4 * C::LPN * 16), w01, w23);
                        ffma2(acc[j][2 * c4 + 0], w01, __uint_as_float(cur.y));
                        ffma2(acc[j][2 * c4 + 1], w23, __uint_as_float(cur.y));
                    }
                }
            }
        }
        p0 = p1; p1 = p2;
        __syncthreads();  // every warp is done with stage s
        if (threadIdx.x == 0 && t + 2 < t_end) issue(t + 2);
    }
    cp_async_wait<0>();

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
        const int64_t col = __shfl_sync(0xffffffffu, my_col, j);
        if (g == 0 && col >= 0) {
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

// padded record count of every (group, tile), flattened as [n_groups][n_tiles + 1] with a zero in the
// last slot of each row: the exclusive scan of this array is goff.
__global__ void stream_counts_kernel(const int32_t* __restrict__ tileptr, const int32_t* __restrict__ perm, int64_t ncol,
                                     int64_t ncol_pad, int n_tiles, int nc, int pad, int64_t n_groups, int64_t* __restrict__ counts) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_groups * (n_tiles + 1)) return;
    const int64_t grp = e / (n_tiles + 1);
    const int t = (int)(e % (n_tiles + 1));
    int64_t tot = 0;
    if (t < n_tiles) {
        for (int j = 0; j < nc; ++j) {
            const int64_t col = perm[grp * nc + j];
            if (col >= 0) {
                const int32_t n = tileptr[(int64_t)(t + 1) * ncol_pad + col] - tileptr[(int64_t)t * ncol_pad + col];
                tot += (n + pad - 1) / pad * pad;
            }
        }
    }
    counts[e] = tot;
}

// fill the warp streams from column-compressed records: one warp per (group, tile)
__global__ void __launch_bounds__(256)
stream_fill_kernel(const uint2* __restrict__ rec, const int64_t* __restrict__ colptr, const int32_t* __restrict__ tileptr,
                   const int32_t* __restrict__ perm, const int64_t* __restrict__ goff, int64_t ncol, int64_t ncol_pad, int n_tiles, int rb_rows, int nc, int pad,
                   int64_t n_groups, uint2* __restrict__ stream) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= n_groups * n_tiles) return;
    const int64_t grp = w / n_tiles;
    const int t = (int)(w % n_tiles);
    int64_t dst = goff[grp * (n_tiles + 1) + t];
    const uint2 padrec = make_uint2((uint32_t)(t * rb_rows), 0u);
    for (int j = 0; j < nc; ++j) {
        const int64_t col = perm[grp * nc + j];
        if (col < 0) continue;
        const int32_t b = tileptr[(int64_t)t * ncol_pad + col], e = tileptr[(int64_t)(t + 1) * ncol_pad + col];
        const int32_t n = e - b, np = (n + pad - 1) / pad * pad;
        const uint2* src = rec + colptr[col] + b;
        for (int32_t r = lane; r < np; r += 32) stream[dst + r] = (r < n) ? src[r] : padrec;
        dst += np;
    }
}

}  // namespace sgl
