// spmm_h16.cuh -- the sparse right-hand-side product with 16-bit staged operands and FP32 accumulation:
//     B[:, c] = sum_{nz (r, v) in X[:, c]} v * F[r, :]        (reference src/singlet.cpp:341-343)
// F is staged as an FP16 shadow scaled by a power of two (F16 = half(F * 2^se)), the non-zero values as FP16 scaled
// by a power of two taken from max |v|; every product half x half is exact in FP32 and is added to an FP32
// accumulator by ONE instruction (PTX fma.rn.f32.f16 -> SASS FHFMA, full rate on sm_100a); the result is multiplied
// by the two inverse scales.
//
// Why: the FP32-operand kernel (spmm.cuh) is bound by the shared-memory crossbar -- every non-zero gathers its
// KP-float row of F (128 B at KP = 32) against 8 B of HBM stream, and the crossbar moves 128 B/clk/SM. A 16-bit
// operand halves the gather and the TMA tile fill, and the tile holds twice the rows, so the padding of the
// per-(column, tile) sub-ranges halves too. Keeping v in FP32 needs an HADD2.F32 conversion per gathered half; that
// instruction only issues on the fmaheavy pipe at half rate and became the bound (measured: profiles/r2_summary.md section 2),
// hence the FP16 value and the mixed-precision FMA.
//
// Layout differences from spmm.cuh (same "warp stream" idea, one contiguous record array per column group):
//   * record = 4 bytes: {uint16 (row inside its tile) * 32, half value}: HBM sees 4 B per non-zero;
//   * NC = 8 columns per warp (8 x 8 factors = 64 accumulator registers), 16 warps = 128 columns per CTA;
//     LPN = KP / 8 lanes cooperate on one non-zero (16 bytes = 8 halves each), SLOTS = 32 / LPN non-zeros per warp
//     step; the WORK of a (column, tile) sub-range is padded to whole steps (SLOTS records);
//   * its STORAGE is padded to blocks of four steps, slot-major: the four records a lane needs for four consecutive
//     steps are 16 contiguous bytes, fetched from the ring by ONE LDS.128 (one shared-memory wavefront per 4 * SLOTS
//     non-zeros instead of four, no per-step record addressing), and every sub-range starts on a block boundary so
//     that the kernel carries no position state from one column to the next;
//   * KP = 32: a row is 64 B, so an 8-lane LDS.128 phase covers TWO rows; they are conflict-free only when they sit
//     in opposite halves of the 128-byte bank line (row parity). The records of a sub-range are therefore stored
//     with even and odd rows alternating (the surplus parity at the end), which makes all phases but the surplus
//     conflict-free (microbench: scripts/microbench, profiles/r2_microbench.txt).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace sgl {

template <int KP>
struct H16Cfg {
    static_assert(KP >= 32, "the 16-bit operand kernel is built for padded ranks >= 32");
    static constexpr int LPN = KP / 8;          // lanes per non-zero
    static constexpr int SLOTS = 32 / LPN;      // non-zeros per warp step
    static constexpr int NC = 8;                // columns per warp
    static constexpr int WARPS = 16;
    static constexpr int COLS_PER_CTA = WARPS * NC;
    static constexpr int PAD = 4 * SLOTS;       // records per storage padding unit = one block
    static constexpr int BLOCK = 4 * SLOTS;     // records per stream block: four steps, slot-major
    static constexpr int BLOCK_BYTES = BLOCK * 4;
    static constexpr int CHUNK = 128;           // records per cp.async chunk (512 B, a whole number of blocks)
    static constexpr int RC = 4;                // ring chunks per warp
    static constexpr int RING_BYTES = WARPS * RC * CHUNK * 4;
    static constexpr int WARP_RING_BYTES = RC * CHUNK * 4;  // 2 KB, and every warp's ring is 2 KB aligned
    static constexpr int ROW_BYTES = KP * 2;
    static constexpr int ROW_SHIFT = (ROW_BYTES == 64) ? 1 : ((ROW_BYTES == 128) ? 2 : 3);  // ROW_BYTES = 32 << ROW_SHIFT
    static constexpr bool INTERLEAVE = (ROW_BYTES == 64);  // two rows per bank line: alternate row parity
};
constexpr int H16_RING_BYTES = 16 * 4 * 128 * 4;
constexpr int H16_SMEM_BYTES = 227 * 1024;

// shared memory: [<= 2 KB alignment slack][rings 32 KB][stage 0][stage 1][2 mbarriers]
static inline int h16_tile_rows(int kp) {
    const int budget = (H16_SMEM_BYTES - 2048 - H16_RING_BYTES - 64) / 2;  // bytes per stage
    int rows = budget / (kp * 2);
    rows &= ~7;
    if (rows > 2040) rows = 2040;  // the record holds row * 32 in 16 bits
    return rows;
}

// ---- FP16 shadow of a factor ------------------------------------------------------------------
// maxbits[0] = bit pattern of max |F| (non-negative floats order like unsigned integers)
__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ F, int64_t n, uint32_t* __restrict__ maxbits) {
    uint32_t mx = 0u;
    for (int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; e < n; e += (int64_t)gridDim.x * blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4*>(F + e);  // n is a multiple of KP >= 4
        mx = max(mx, __float_as_uint(v.x) & 0x7fffffffu);
        mx = max(mx, __float_as_uint(v.y) & 0x7fffffffu);
        mx = max(mx, __float_as_uint(v.z) & 0x7fffffffu);
        mx = max(mx, __float_as_uint(v.w) & 0x7fffffffu);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(maxbits, mx);
}
// F16 = half(F * 2^se) with se chosen so that max |F| * 2^se lies in [2^14, 2^15); inv_scale[0] = 2^-se
__global__ void __launch_bounds__(256)
shadow_kernel(const float* __restrict__ F, int64_t n, const uint32_t* __restrict__ maxbits, __half* __restrict__ F16,
              float* __restrict__ inv_scale) {
    const uint32_t mb = maxbits[0];
    int E = (int)(mb >> 23) - 127;  // max = 1.m * 2^E (subnormal / zero: E = -127)
    if (mb == 0u) E = 14;
    int se = 14 - E;
    se = se > 126 ? 126 : (se < -126 ? -126 : se);
    const float scale = __uint_as_float((uint32_t)(127 + se) << 23);
    if (blockIdx.x == 0 && threadIdx.x == 0) inv_scale[0] = __uint_as_float((uint32_t)(127 - se) << 23);
    for (int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; e < n; e += (int64_t)gridDim.x * blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4*>(F + e);
        const __half2 a = __floats2half2_rn(v.x * scale, v.y * scale);
        const __half2 b = __floats2half2_rn(v.z * scale, v.w * scale);
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&a);
        o.y = *reinterpret_cast<const uint32_t*>(&b);
        *reinterpret_cast<uint2*>(F16 + e) = o;
    }
}

// ---- stream construction ------------------------------------------------------------------------
// max |value| of the records (bit pattern, see absmax_kernel)
__global__ void __launch_bounds__(256)
rec_absmax_kernel(const uint2* __restrict__ rec, int64_t n, uint32_t* __restrict__ maxbits) {
    uint32_t mx = 0u;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        mx = max(mx, rec[e].y & 0x7fffffffu);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(maxbits, mx);
}
// power-of-two scale that maps max |x| (bit pattern) into [2^14, 2^15) -- the same rule on host and device
__host__ __device__ __forceinline__ int h16_scale_exp(uint32_t maxbits) {
    int E = (int)((maxbits & 0x7fffffffu) >> 23) - 127;
    if ((maxbits & 0x7fffffffu) == 0u) E = 14;
    int se = 14 - E;
    return se > 126 ? 126 : (se < -126 ? -126 : se);
}

// record = {low 16 bits: row_in_tile * 32 (the gather address is base + (that << ROW_SHIFT): one mask + one LEA),
//           high 16 bits: half(value * vscale)}
// The value is rounded to FP16 STOCHASTICALLY with a hash of the record's position in the matrix as the random
// number: up with probability equal to the discarded fraction (add 13 random bits below the FP16 mantissa, then
// truncate). Round-to-nearest would give every occurrence of a value the SAME error -- count data has few distinct
// values (the synthetic configs have eight) -- and the error of b = sum v * w would not average out over the
// non-zeros of a column (measured 1.6e-4 relative); with the dither the errors are independent with zero mean
// (1e-5). Values that are exact in FP16 (integer counts up to 2048) stay exact. Deterministic: same matrix, same stream.
__device__ __forceinline__ uint32_t h16_pack(uint32_t row_in_tile, uint32_t vbits, float vscale, uint64_t position) {
    uint32_t hsh = (uint32_t)position ^ (uint32_t)(position >> 32) * 0x9E3779B9u;
    hsh ^= hsh >> 16; hsh *= 0x85EBCA6Bu; hsh ^= hsh >> 13; hsh *= 0xC2B2AE35u; hsh ^= hsh >> 16;  // murmur3 finaliser
    const float x = __uint_as_float(vbits) * vscale;
    const uint32_t xb = __float_as_uint(x);
    const bool finite = (xb & 0x7f800000u) != 0x7f800000u;
    const float xd = finite ? __uint_as_float(xb + (hsh & 0x1fffu)) : x;
    const __half hv = __float2half_rz(xd);
    return ((row_in_tile << 5) & 0xffffu) | ((uint32_t)__half_as_ushort(hv) << 16);
}
// stream position of record `i` of a group's stream: blocks of four steps, slot-major (see the header)
__device__ __forceinline__ int64_t h16_stream_pos(int64_t i, int slots) {
    const int64_t step = i / slots;
    const int slot = (int)(i % slots);
    return (step >> 2) * (4 * slots) + slot * 4 + (step & 3);
}
// padded record count of every (group, tile), flattened as [n_groups][n_tiles + 1] with a zero in the last slot of
// each row (its exclusive scan is goff); every sub-range takes whole blocks
__global__ void stream_counts_h16_kernel(const int32_t* __restrict__ tileptr, const int32_t* __restrict__ perm, int64_t ncol_pad,
                                         int n_tiles, int nc, int block, int64_t n_groups, int64_t* __restrict__ counts) {
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
                tot += (n + block - 1) / block * block;
            }
        }
    }
    counts[e] = tot;
}

// one warp per (group, tile): packs the sub-ranges of the group's columns into 4-byte records {row in tile * 32, half
// value * vscale}, alternating row parity when `interleave`, padded to whole blocks with 0, each sub-range in block
// layout (h16_stream_pos relative to its own start)
__global__ void __launch_bounds__(256)
stream_fill_h16_kernel(const uint2* __restrict__ rec, const int64_t* __restrict__ colptr, const int32_t* __restrict__ tileptr,
                       const int32_t* __restrict__ perm, const int64_t* __restrict__ goff, int64_t ncol_pad, int n_tiles,
                       int rb_rows, int nc, int slots, int interleave, float vscale, int64_t n_groups,
                       uint32_t* __restrict__ stream) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= n_groups * n_tiles) return;
    const int64_t grp = w / n_tiles;
    const int t = (int)(w % n_tiles);
    const int block = 4 * slots;
    uint32_t* out = stream + goff[grp * (n_tiles + 1) + t];
    const uint32_t row0 = (uint32_t)t * (uint32_t)rb_rows;
    for (int j = 0; j < nc; ++j) {
        const int64_t col = perm[grp * nc + j];
        if (col < 0) continue;
        const int32_t b = tileptr[(int64_t)t * ncol_pad + col], e = tileptr[(int64_t)(t + 1) * ncol_pad + col];
        const int32_t n = e - b, np = (n + block - 1) / block * block;
        const uint2* src = rec + colptr[col] + b;
        if (!interleave) {
            for (int32_t r = lane; r < np; r += 32) {
                uint32_t o = 0u;
                if (r < n) {
                    const uint2 v = src[r];
                    o = h16_pack(v.x - row0, v.y, vscale, (uint64_t)(colptr[col] + b + r));
                }
                out[h16_stream_pos(r, slots)] = o;
            }
        } else {
            int32_t n_even = 0;
            for (int32_t r0 = 0; r0 < n; r0 += 32) {
                const int32_t r = r0 + lane;
                const bool ev = (r < n) && (((src[r].x - row0) & 1u) == 0u);
                n_even += __popc(__ballot_sync(0xffffffffu, ev));
            }
            const int32_t n_odd = n - n_even, mn = n_even < n_odd ? n_even : n_odd;
            int32_t ce = 0, co = 0;  // even / odd rows placed so far
            for (int32_t r0 = 0; r0 < n; r0 += 32) {
                const int32_t r = r0 + lane;
                uint2 v = make_uint2(0u, 0u);
                bool ev = false, od = false;
                if (r < n) {
                    v = src[r];
                    v.x -= row0;
                    ev = (v.x & 1u) == 0u;
                    od = !ev;
                }
                const uint32_t be = __ballot_sync(0xffffffffu, ev), bo = __ballot_sync(0xffffffffu, od);
                const uint32_t below = (1u << lane) - 1u;
                if (r < n) {
                    const int32_t ip = ev ? ce + __popc(be & below) : co + __popc(bo & below);  // index among its parity
                    const int32_t pos = ip < mn ? 2 * ip + (ev ? 0 : 1) : 2 * mn + (ip - mn);
                    out[h16_stream_pos(pos, slots)] = h16_pack(v.x, v.y, vscale, (uint64_t)(colptr[col] + b + r));
                }
                ce += __popc(be);
                co += __popc(bo);
            }
            for (int32_t r = n + lane; r < np; r += 32) out[h16_stream_pos(r, slots)] = 0u;
        }
        out += np;
    }
}

// ---- the kernel -----------------------------------------------------------------------------------
// acc[f] += half(w, f) * half(rec >> 16): 8 mixed-precision FMAs (FHFMA), one per gathered half
__device__ __forceinline__ void h16_fma8(float (&acc)[8], const uint4& w, uint32_t rec) {
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
    unsigned short vlo, vh;
    asm("mov.b32 {%0, %1}, %2;" : "=h"(vlo), "=h"(vh) : "r"(rec));
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        unsigned short lo, hi;
        asm("mov.b32 {%0, %1}, %2;" : "=h"(lo), "=h"(hi) : "r"(ww[p]));
        asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc[2 * p]) : "h"(lo), "h"(vh));
        asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc[2 * p + 1]) : "h"(hi), "h"(vh));
    }
}

template <int KP>
__global__ void __launch_bounds__(H16Cfg<KP>::WARPS * 32, 1)
spmm_h16_kernel(const uint32_t* __restrict__ stream,  // warp streams of 4-byte records
                const int64_t* __restrict__ goff,     // [n_groups][n_tiles + 1]
                const int32_t* __restrict__ tileptr,  // [n_tiles + 1][ncol_pad]
                const int32_t* __restrict__ perm,     // [n_groups * NC]
                int64_t ncol, int64_t ncol_pad, int64_t nrow, int rb_rows, int n_tiles, int tiles_per_split,
                const __half* __restrict__ F16,       // [nrow][KP] scaled shadow
                const float* __restrict__ inv_scale,  // 2^-se of the shadow (device)
                float inv_vscale,                     // 2^-sv of the record values
                float* __restrict__ Bout)             // [splits][ncol][KP]
{
    using C = H16Cfg<KP>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t stage_bytes = (uint32_t)rb_rows * C::ROW_BYTES;
    // rings first, on a 2 KB boundary of the shared window: a ring address is then (byte position & 2047) | base
    const uint32_t smem0 = smem_u32(smem_raw);
    const uint32_t ring_all = (smem0 + 2047u) & ~2047u;
    const uint32_t stage0 = ring_all + C::RING_BYTES, stage1 = stage0 + stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (stage1 + stage_bytes - smem0));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = lane % C::LPN;  // 16-byte chunk of the row owned by this lane
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

    auto issue = [&](int t) {  // thread 0 only: stage the F16 tile t
        const int s = (t - t_begin) & 1;
        const int64_t r0 = (int64_t)t * rb_rows;
        const int64_t rows = min((int64_t)rb_rows, nrow - r0);
        const uint32_t bytes = (uint32_t)(rows * C::ROW_BYTES);
        mbar_arrive_expect_tx(&bars[s], bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s ? stage1 : stage0),
                     "l"(reinterpret_cast<const unsigned char*>(F16) + r0 * C::ROW_BYTES), "r"(bytes), "r"(smem_u32(&bars[s]))
                     : "memory");
    };
    if (threadIdx.x == 0 && t_begin < t_end) {
        issue(t_begin);
        if (t_begin + 1 < t_end) issue(t_begin + 1);
    }

    // ---- this warp's record stream and ring ----
    int64_t s_begin = 0, s_total = 0;
    if (group_ok && t_begin < t_end) {
        s_begin = goff[group * (n_tiles + 1) + t_begin];
        s_total = goff[group * (n_tiles + 1) + t_end] - s_begin;  // records, a whole number of blocks
    }
    const uint32_t* sp = stream + s_begin;  // whole blocks: 16-byte aligned
    const uint32_t ring = ring_all + (uint32_t)warp * C::WARP_RING_BYTES;
    auto fill = [&](int64_t c) {  // whole warp: chunk c -> ring slot c % RC (4 records per lane); past the end: zeros
        const int64_t r = c * C::CHUNK + 4 * lane;
        const bool in = r < s_total;  // s_total and r are multiples of 4: a lane's quad is in or out as a whole
        cp_async16(ring + (uint32_t)((c % C::RC) * C::CHUNK + 4 * lane) * 4u, sp + (in ? r : 0), in ? 16u : 0u);
        cp_async_commit();
    };
#pragma unroll
    for (int c = 0; c < C::RC; ++c) fill(c);
    cp_async_wait<C::RC - 1>();  // chunk 0 has landed
    __syncwarp();

    float acc[C::NC][8];
#pragma unroll
    for (int j = 0; j < C::NC; ++j)
#pragma unroll
        for (int f = 0; f < 8; ++f) acc[j][f] = 0.f;

    // per-column record counts of a tile come from the tile index (lanes < NC hold one column each)
    const int64_t my_col = (lane < C::NC && group_ok) ? (int64_t)perm[col0 + lane] : -1;
    const bool my_col_ok = my_col >= 0;
    int32_t p0 = 0, p1 = 0;
    if (my_col_ok && t_begin < t_end) {
        p0 = tileptr[(int64_t)t_begin * ncol_pad + my_col];
        p1 = tileptr[(int64_t)(t_begin + 1) * ncol_pad + my_col];
    }

    // records: one LDS.128 fetches this lane's slot of the four steps of a block
    const uint32_t ring_g = ring + (uint32_t)g * 16u;
    uint32_t bpos = 0;  // ring byte position of the next block
    auto next_block = [&]() {  // the next block of the stream (waiting for / refilling ring chunks as they turn over)
        uint4 r;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                     : "r"((bpos & (uint32_t)(C::WARP_RING_BYTES - 1)) | ring_g));  // bpos never touches the bits of g * 16
        bpos += C::BLOCK_BYTES;
        if ((bpos & (uint32_t)(C::CHUNK * 4 - 1)) == 0) {  // the chunk just finished is free; the next one must have landed
            __syncwarp();
            fill((int64_t)(bpos / (C::CHUNK * 4)) + C::RC - 1);
            cp_async_wait<C::RC - 1>();
            __syncwarp();
        }
        return r;
    };
    auto gather = [&](uint32_t base, uint32_t rec) {  // the 16 bytes of the record's row owned by this lane
        uint4 w;
        const uint32_t addr = base + ((rec & 0xffffu) << C::ROW_SHIFT);
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "r"(addr));
        return w;
    };
    // (reading the records one block ahead and handing them over in registers was measured SLOWER: the four extra moves
    //  per block cost more issue slots than the hidden LDS latency gives back -- profiles/r2_summary.md section 2)
    uint32_t phase0 = 0, phase1 = 0;
    for (int t = t_begin; t < t_end; ++t) {
        const int s = (t - t_begin) & 1;
        const int32_t steps_l = (p1 - p0 + C::SLOTS - 1) / C::SLOTS;  // steps of my column in this tile
        int32_t p2 = p1;
        if (my_col_ok && t + 1 < t_end) p2 = tileptr[(int64_t)(t + 2) * ncol_pad + my_col];
        mbar_wait(&bars[s], s ? phase1 : phase0);
        if (s) phase1 ^= 1u; else phase0 ^= 1u;
        const uint32_t base = (s ? stage1 : stage0) + (uint32_t)q * 16u;

#pragma unroll
        for (int j = 0; j < C::NC; ++j) {
            int32_t left = __shfl_sync(0xffffffffu, steps_l, j);  // warp-uniform
#pragma unroll 1
            for (; left >= 4; left -= 4) {  // whole blocks: four gathers in flight, then 32 FMAs
                const uint4 rq = next_block();
                const uint4 wa = gather(base, rq.x);
                const uint4 wb = gather(base, rq.y);
                const uint4 wc = gather(base, rq.z);
                const uint4 wd = gather(base, rq.w);
                h16_fma8(acc[j], wa, rq.x);
                h16_fma8(acc[j], wb, rq.y);
                h16_fma8(acc[j], wc, rq.z);
                h16_fma8(acc[j], wd, rq.w);
            }
            if (left > 0) {  // the last, partly filled block (its unused steps hold zero records that are not executed)
                const uint4 rq = next_block();
                const uint4 wa = gather(base, rq.x);
                const uint4 wb = gather(base, rq.y);  // harmless when left == 1: a zero record gathers row 0
                h16_fma8(acc[j], wa, rq.x);
                if (left > 1) h16_fma8(acc[j], wb, rq.y);
                if (left > 2) {
                    const uint4 wc = gather(base, rq.z);
                    h16_fma8(acc[j], wc, rq.z);
                }
            }
        }
        p0 = p1; p1 = p2;
        __syncthreads();  // every warp is done with stage s
        if (threadIdx.x == 0 && t + 2 < t_end) issue(t + 2);
    }
    cp_async_wait<0>();

    // fold the SLOTS partial sums (lanes with equal q), undo the operand scales and store
    const float inv = inv_scale[0] * inv_vscale;
#pragma unroll
    for (int j = 0; j < C::NC; ++j) {
        float out[8];
#pragma unroll
        for (int f = 0; f < 8; ++f) {
            float a = acc[j][f];
#pragma unroll
            for (int o = C::LPN; o < 32; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            out[f] = a * inv;
        }
        const int64_t col = __shfl_sync(0xffffffffu, my_col, j);
        if (g == 0 && col >= 0) {
            float4* dst = reinterpret_cast<float4*>(Bout + ((int64_t)blockIdx.y * ncol + col) * KP + q * 8);
            dst[0] = make_float4(out[0], out[1], out[2], out[3]);
            dst[1] = make_float4(out[4], out[5], out[6], out[7]);
        }
    }
}

}  // namespace sgl
