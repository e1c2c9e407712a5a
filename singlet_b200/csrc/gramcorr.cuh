// gramcorr.cuh -- the masked solver's per-column Gram correction on the tensor cores:
//     G_M(c) = sum over the held-out rows r of column c of f_r f_r^T        (reference src/singlet.cpp:460-461)
// so that the solver can use a_c = G - G_M(c) (:462). k^2 multiply-adds per held-out entry: at cross-validation sizes this
// is 40 % of the masked solver's instructions, on a large matrix (30k x 100k, k = 32: 3e8 held-out entries per iteration)
// it is 94 % of the whole ALS iteration when done with FP32 FMAs (profiles/r2_summary.md section 4).
//
// One warp = one column. The operand is the OTHER factor F (k x rows, FP32). It is split once per half-iteration into two
// BF16 planes, hi = bf16(F) and mid = bf16(F - hi): 16 mantissa bits together. The held-out rows are gathered 16 at a time
// into a shared-memory ring with cp.async (64 + 64 contiguous bytes per row at KP = 32, the same bytes as the FP32 row), read as
// mma fragments with ldmatrix.trans (the A operand hi^T and the B operand hi / mid are the SAME registers: A[i][e] =
// F[e][i] = B[e][i]) and accumulated in FP32 by mma.sync.m16n8k16 (SASS HMMA.16816.F32.BF16):
//     D1 += hi^T hi,   D2 += hi^T mid,    G_M = D1 + D2 + D2^T            (mid^T mid ~ 2^-18 |G_M| is dropped)
// i.e. two tensor passes instead of three because the two cross terms are transposes of each other; the transpose is done
// once per column through the (then idle) ring. Error of G_M: ~2^-17 relative per product, averaging down over the list
// -- the same order as the FP32 accumulation error of the FFMA path (tests/test_gpu_parity.py compares both paths).
// Measured (scripts/microbench, profiles/r2_microbench.txt): 1.14 clk per (column, held-out row) per SM and pass against
// ~8 clk at the FP32 FMA peak. tcgen05.mma was considered and not used: its smallest tile is M = 64 x N = 8 per CTA with the
// operands described as whole shared-memory tiles; here every column has its OWN 32 x 32 output and its own gathered row list,
// so four columns would have to share one 128 x 128 accumulator of which only the diagonal blocks are wanted (75 % wasted
// MMA work), while the gather (128 B per entry from L2) already bounds the kernel near 2 clk per entry.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace sgl {

// F [rows][KP] FP32 -> BF16 pairs [rows][2][KP]: the hi plane of a row followed by its mid plane, so that the two 2*KP-byte
// halves a held-out entry gathers are ONE contiguous, aligned 4*KP-byte piece of L2 (a whole 128-byte line at KP = 32)
template <int KP>
__global__ void __launch_bounds__(256)
bf16_split_kernel(const float* __restrict__ F, int64_t n, uint16_t* __restrict__ pairs) {
    for (int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; e < n; e += (int64_t)gridDim.x * blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4*>(F + e);
        const float f[4] = {v.x, v.y, v.z, v.w};
        uint16_t h[4], m[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const __nv_bfloat16 bh = __float2bfloat16_rn(f[q]);
            const __nv_bfloat16 bm = __float2bfloat16_rn(f[q] - __bfloat162float(bh));
            h[q] = __bfloat16_as_ushort(bh);
            m[q] = __bfloat16_as_ushort(bm);
        }
        const int64_t row = e / KP;
        const int c = (int)(e % KP);
        uint16_t* dst = pairs + row * (2 * KP) + c;
        *reinterpret_cast<uint2*>(dst) = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
        *reinterpret_cast<uint2*>(dst + KP) = make_uint2((uint32_t)m[0] | ((uint32_t)m[1] << 16), (uint32_t)m[2] | ((uint32_t)m[3] << 16));
    }
}

// PLANES = 2: the BF16 hi | mid pairs above (FP32-equivalent corrections). PLANES = 1: ONE pass over the FP16 shadow of the
// factor that the 16-bit-operand SpMM of the same half-iteration has just built (spmm_h16.cuh: half(F * 2^se), [rows][KP]) --
// half the gathered bytes and half the MMAs; used where the precision policy already stages the sparse product's operands in
// 16 bits (large matrices, sgl_set_precision): an 11-bit operand perturbs G_M -- 5 % of a -- by ~1e-5, far below the 1.6e-4 the
// 16-bit product leaves in the right-hand sides.
template <int KP, int PLANES = 2>
struct GramCorrCfg {
    static_assert(KP == 16 || KP == 32 || KP == 64, "tensor-core Gram correction: padded ranks 16, 32 and 64");
    static_assert(PLANES == 2 || (PLANES == 1 && KP >= 32), "the FP16 shadow exists for padded ranks >= 32");
    static constexpr int MT = KP / 16, NT = KP / 8;     // 16-row and 8-column tiles of the KP x KP output
    // KP = 64: the 64 x 64 output does not fit one warp's registers; WPC = 2 warps share a column, each owning MTW = 2 of
    // the four 16-row tiles (both gather the whole list: the B operand needs every factor). With the output split, D2^T
    // would live in the other warp, so the split path adds the third pass (mid^T hi) into the SAME accumulator instead of
    // transposing: G_M = hi^T hi + hi^T mid + mid^T hi.
    static constexpr int WPC = (KP + 31) / 32 > 1 ? KP / 32 : 1;
    static constexpr int MTW = MT / WPC;
    static constexpr bool THREE_PASS = (WPC > 1) && (PLANES == 2);
    static constexpr int EB = KP * 2 * PLANES;          // bytes of one gathered entry as it lies in global memory (hi row | mid row)
    static constexpr int CPE = EB / 16;                 // 16-byte chunks per entry (8 / 4)
    static constexpr int EPL = (EB >= 128) ? 1 : 128 / EB;  // entries per 128-byte line of the ring
    static constexpr int LPE = (EB >= 128) ? EB / 128 : 1;  // lines per entry (2 at KP = 64 with two planes)
    static constexpr int SW_MASK = (CPE < 8 ? CPE : 8) - 1; // chunk c of entry e sits at chunk ((e % EPL) * CPE + c) ^ ((e / EPL) & SW_MASK) of
                                                        // line e / EPL: unpadded lines, so the 8 lanes that copy one line with cp.async write
                                                        // one conflict-free wavefront, AND the 8 row addresses of an ldmatrix tile (8
                                                        // consecutive entries, same logical chunk) fall into 8 distinct 16-byte bank groups
    static constexpr int BLK = 16;                      // held-out rows per block = K of one mma
    static constexpr int STAGE_BYTES = BLK * EB;
    static constexpr int STAGES = 4;
    static constexpr int WARPS = 4;
    static constexpr int WARP_BYTES = STAGES * STAGE_BYTES;
    __host__ __device__ static constexpr uint32_t offset(int e, int c) {  // byte offset of chunk c of entry e inside a stage
        return (uint32_t)(((e / EPL) * LPE + c / 8) * 128 + (((((e % EPL) * CPE + c) % 8) ^ ((e / EPL) & SW_MASK)) * 16));
    }
    static_assert(PLANES == 1 || THREE_PASS || WARP_BYTES >= KP * (KP + 1) * 4, "the transpose scratch reuses the ring");
};

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// gm[(col - col0) * KP * KP + i * KP + j] = G_M(col)[i][j] for col0 <= col < col0 + ncols (columns without non-zeros are
// skipped like in the solver: :444). pairs: the BF16 hi / mid planes of the gather factor, [rows][2][KP].
// PLANES = 1: `pairs` is the FP16 shadow [rows][KP] and inv_scale[0] = 2^-se its inverse scale (device): G_M = D1 * 2^-2se.
template <int KP, int PLANES>
__global__ void __launch_bounds__(GramCorrCfg<KP, PLANES>::WARPS * 32)
gram_corr_mma_kernel(const uint16_t* __restrict__ pairs, const int64_t* __restrict__ colptr,
                     const int64_t* __restrict__ mptr, const uint2* __restrict__ mrec, int64_t col0, int64_t ncols,
                     float* __restrict__ gm, const float* __restrict__ inv_scale) {
    using C = GramCorrCfg<KP, PLANES>;
    constexpr int MT = C::MTW, NT = C::NT;  // MT: the m-tiles of THIS warp
    extern __shared__ __align__(128) unsigned char ring_dyn[];  // [WARPS][WARP_BYTES] (64 KB at KP = 64 with two planes)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t cl = ((int64_t)blockIdx.x * C::WARPS + warp) / C::WPC;
    const int mt0 = ((int)(warp % C::WPC)) * MT;  // first m-tile of this warp (KP = 64: rows 0-31 or 32-63 of the output)
    if (cl >= ncols) return;  // warp-uniform; no CTA-wide barrier below
    const int64_t col = col0 + cl;
    if (colptr[col] == colptr[col + 1]) return;
    const int64_t mb = mptr[col];
    const int n = (int)(mptr[col + 1] - mb);
    const int nblk = (n + C::BLK - 1) / C::BLK;
    unsigned char* ring_mine = ring_dyn + (size_t)warp * C::WARP_BYTES;
    const uint32_t ring = smem_u32(ring_mine);

    constexpr bool TWO_ACC = (PLANES == 2) && !C::THREE_PASS;
    float d1[MT][NT][4], d2[TWO_ACC ? MT : 1][TWO_ACC ? NT : 1][4];
#pragma unroll
    for (int p = 0; p < MT; ++p)
#pragma unroll
        for (int q = 0; q < NT; ++q)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                d1[p][q][c] = 0.f;
                if constexpr (TWO_ACC) d2[p][q][c] = 0.f;
            }

    // A block is 16 entries of 4 * KP contiguous bytes (hi row | mid row). One cp.async instruction copies 32 / CPE whole
    // entries, CPE adjacent lanes covering the bytes of one entry in order, so that every request of the instruction is made
    // of whole 32-byte sectors and lands in shared memory as whole conflict-free lines. (One lane per 16 bytes with the two
    // planes 2 * KP bytes apart, as first written, fetched every sector twice: 264 B per entry over the L2 crossbar; a padded
    // 144-byte row split the shared-memory side of each request into ~10 wavefronts and still moved 200 B per entry --
    // profiles/r2_summary.md section 4.) The row index of entry e of the block is held by lane e.
    constexpr int CPE = C::CPE, EPI = 32 / CPE, NI = C::BLK / EPI;
    auto load_idx = [&](int blk) -> uint32_t {
        const int t = blk * C::BLK + (lane & 15);
        return (t < n) ? mrec[mb + t].x : 0u;
    };
    auto issue = [&](int blk, int stage, uint32_t idxreg) {
        const int chunk = lane % CPE;
#pragma unroll
        for (int q = 0; q < NI; ++q) {
            const int e = q * EPI + lane / CPE;
            const uint32_t row = __shfl_sync(0xffffffffu, idxreg, e);
            const bool ok = blk * C::BLK + e < n;  // entries past the end are zero-filled: they add nothing
            const uint16_t* src = pairs + (int64_t)row * (PLANES * KP) + chunk * 8;
            const uint32_t dst = ring + (uint32_t)(stage * C::STAGE_BYTES) + C::offset(e, chunk);
            cp_async16(dst, src, ok ? 16u : 0u);
        }
    };
#pragma unroll
    for (int s = 0; s < C::STAGES; ++s) {
        issue(s, s, load_idx(s));
        cp_async_commit();
    }
    // Row indices (streamed from HBM: 8 bytes per held-out entry, not L2-resident like the factor) are fetched two blocks
    // before the copy that needs them is issued. The loop is unrolled by U = 2 with one index register per residue: rotating
    // two registers made the compiler wait for the younger load at the rotation itself (55 % of the stall samples), one
    // block of distance left 17 % of the samples on the shuffle that consumes the index, two blocks 25 % of fewer samples
    // (1.15 -> 1.05 ms), and four blocks (U = 4) was slower again (profiles/r2_summary.md section 4).
    constexpr int U = 2;
    static_assert(C::STAGES % U == 0, "block b and block b + STAGES share an index register");
    uint32_t idx_r[U];
#pragma unroll
    for (int r = 0; r < U; ++r) idx_r[r] = load_idx(C::STAGES + r);
    // Two fragment orders of the same tiles: for the B operand an x4 load returns {P[2j], Q[2j], P[2j+1], Q[2j+1]} (tile ti:
    // entry half h = ti & 1, factor octet 2 j + (ti >> 1)) -- the register PAIRS mma wants; for the A operand {P[2j], P[2j+1],
    // Q[2j], Q[2j+1]} (h = ti >> 1, octet 2 j + (ti & 1)) -- the register QUAD of m-tile j. The hi plane is loaded in both
    // orders: two more LDSM per block instead of four register moves in front of every HMMA.
    const int lm_e = 8 * ((lane >> 3) & 1) + (lane & 7), lm_q = lane >> 4;
    const int la_e = 8 * (lane >> 4) + (lane & 7), la_q = (lane >> 3) & 1;
    auto block = [&](int b, uint32_t& idx_mine) {
        cp_async_wait<C::STAGES - 1>();
        __syncwarp();
        const uint32_t st = ring + (uint32_t)((b % C::STAGES) * C::STAGE_BYTES);
        // P[q] = {F[2t][8q+g], F[2t+1][8q+g]}, Q[q] = the same for entries 8 + 2t, 9 + 2t  (g = lane / 4, t = lane % 4)
        uint32_t Ah[MT][4], Am[C::THREE_PASS ? MT : 1][4], Bh[NT / 2][4], Bm[PLANES == 2 ? NT / 2 : 1][4];
#pragma unroll
        for (int pl = 0; pl < MT; ++pl) {
            ldmatrix_x4_trans(st + C::offset(la_e, 2 * (mt0 + pl) + la_q), Ah[pl][0], Ah[pl][1], Ah[pl][2], Ah[pl][3]);
            if constexpr (C::THREE_PASS)
                ldmatrix_x4_trans(st + C::offset(la_e, CPE / 2 + 2 * (mt0 + pl) + la_q), Am[pl][0], Am[pl][1], Am[pl][2], Am[pl][3]);
        }
#pragma unroll
        for (int j = 0; j < NT / 2; ++j) {
            ldmatrix_x4_trans(st + C::offset(lm_e, 2 * j + lm_q), Bh[j][0], Bh[j][1], Bh[j][2], Bh[j][3]);
            if constexpr (PLANES == 2) ldmatrix_x4_trans(st + C::offset(lm_e, CPE / 2 + 2 * j + lm_q), Bm[j][0], Bm[j][1], Bm[j][2], Bm[j][3]);
        }
#pragma unroll
        for (int p = 0; p < MT; ++p)
#pragma unroll
            for (int q = 0; q < NT; ++q) {
                if constexpr (C::THREE_PASS) {
                    mma_bf16_16816(d1[p][q], Ah[p][0], Ah[p][1], Ah[p][2], Ah[p][3], Bh[q / 2][2 * (q & 1)], Bh[q / 2][2 * (q & 1) + 1]);
                    mma_bf16_16816(d1[p][q], Ah[p][0], Ah[p][1], Ah[p][2], Ah[p][3], Bm[q / 2][2 * (q & 1)], Bm[q / 2][2 * (q & 1) + 1]);
                    mma_bf16_16816(d1[p][q], Am[p][0], Am[p][1], Am[p][2], Am[p][3], Bh[q / 2][2 * (q & 1)], Bh[q / 2][2 * (q & 1) + 1]);
                } else if constexpr (PLANES == 2) {
                    mma_bf16_16816(d1[p][q], Ah[p][0], Ah[p][1], Ah[p][2], Ah[p][3], Bh[q / 2][2 * (q & 1)], Bh[q / 2][2 * (q & 1) + 1]);
                    mma_bf16_16816(d2[p][q], Ah[p][0], Ah[p][1], Ah[p][2], Ah[p][3], Bm[q / 2][2 * (q & 1)], Bm[q / 2][2 * (q & 1) + 1]);
                } else {
                    mma_f16_16816(d1[p][q], Ah[p][0], Ah[p][1], Ah[p][2], Ah[p][3], Bh[q / 2][2 * (q & 1)], Bh[q / 2][2 * (q & 1) + 1]);
                }
            }
        __syncwarp();  // every lane has read the stage
        issue(b + C::STAGES, b % C::STAGES, idx_mine);
        cp_async_commit();
        idx_mine = load_idx(b + C::STAGES + U);
    };
    for (int b = 0; b < nblk; b += U) {
#pragma unroll
        for (int r = 0; r < U; ++r)
            if (b + r < nblk) block(b + r, idx_r[r]);  // warp-uniform
    }
    cp_async_wait<0>();
    __syncwarp();
    const int g = lane >> 2, t = lane & 3;
    float* out = gm + cl * (int64_t)(KP * KP);
    if constexpr (!TWO_ACC) {
        const float sc2 = (PLANES == 1) ? inv_scale[0] * inv_scale[0] : 1.f;
#pragma unroll
        for (int p = 0; p < MT; ++p)
#pragma unroll
            for (int q = 0; q < NT; ++q) {
                const int i = 16 * (mt0 + p) + g, j = 8 * q + 2 * t;
                *reinterpret_cast<float2*>(out + i * KP + j) = make_float2(d1[p][q][0] * sc2, d1[p][q][1] * sc2);
                *reinterpret_cast<float2*>(out + (i + 8) * KP + j) = make_float2(d1[p][q][2] * sc2, d1[p][q][3] * sc2);
            }
    } else {
    // G_M = D1 + D2 + D2^T: D2 goes through the ring transposed (row stride KP + 1: the four t-lanes of a row group hit
    // different banks)
    float* sc = reinterpret_cast<float*>(ring_mine);
#pragma unroll
    for (int p = 0; p < MT; ++p)
#pragma unroll
        for (int q = 0; q < NT; ++q) {
            const int i = 16 * p + g, j = 8 * q + 2 * t;
            sc[j * (KP + 1) + i] = d2[p][q][0];
            sc[(j + 1) * (KP + 1) + i] = d2[p][q][1];
            sc[j * (KP + 1) + i + 8] = d2[p][q][2];
            sc[(j + 1) * (KP + 1) + i + 8] = d2[p][q][3];
        }
    __syncwarp();
#pragma unroll
    for (int p = 0; p < MT; ++p)
#pragma unroll
        for (int q = 0; q < NT; ++q) {
            const int i = 16 * p + g, j = 8 * q + 2 * t;
            const float v0 = d1[p][q][0] + d2[p][q][0] + sc[i * (KP + 1) + j];
            const float v1 = d1[p][q][1] + d2[p][q][1] + sc[i * (KP + 1) + j + 1];
            const float v2 = d1[p][q][2] + d2[p][q][2] + sc[(i + 8) * (KP + 1) + j];
            const float v3 = d1[p][q][3] + d2[p][q][3] + sc[(i + 8) * (KP + 1) + j + 1];
            *reinterpret_cast<float2*>(out + i * KP + j) = make_float2(v0, v1);
            *reinterpret_cast<float2*>(out + (i + 8) * KP + j) = make_float2(v2, v3);
        }
    }
}

}  // namespace sgl
