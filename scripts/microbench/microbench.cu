// microbench.cu -- B200 micro-benchmarks behind the SpMM design choices. Developer tool, not product
// code: build with `make -C scripts/microbench`, run on the GPU box, results summarised under profiles/.
//   1. shared-memory gather: LDS.128 of 64-byte rows (FP16 operand, 4 lanes per row) with the two rows of every
//      8-lane phase in the same / opposite halves of the 128-byte bank line, and 128-byte rows (FP32 operand)
//   2. FP32-accumulate math on a 16-bit operand: HADD2.F32 + FFMA2 against FHFMA (fma.rn.f32.f16) against FFMA2 alone
//   3. tensor memory as a gather table: tcgen05.ld.32x32b.x1 with a data-dependent column
//   4. the masked solver's Gram correction G_M = sum_r f_r f_r^T (k = 32) per column: the FP32 rank-1 update of
//      nnls_masked_sub_kernel (8 lanes per column, 128 FFMA per lane and held-out row) against warp-level tensor-core
//      SYRK blocks of 8 held-out rows (8 x mma.sync.m16n8k8 TF32, single pass and 3-pass split) and of 16 rows in BF16
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- 1. gather ------------------------------------------------------------------------------------------------
// mode 0: 64 B rows, slot pairs in opposite halves; 1: 64 B rows, random; 2: 64 B rows, same half;
// mode 3: 128 B rows, 8 lanes per row (the FP32 layout)
template <int MODE>
__global__ void __launch_bounds__(512, 1) gather_kernel(int iters, unsigned long long* out_clk, float* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = (float)i;
    __syncthreads();
    const uint32_t base = smem_u32(smem);
    uint32_t state = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 17u;
    float acc = 0.f;
    const int lpn = (MODE == 3) ? 8 : 4;
    const int slot = lane / lpn, q = lane % lpn;
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            // a pseudo-random row per slot, identical for the lanes of a slot
            state = state * 1664525u + 1013904223u;
            uint32_t r = __shfl_sync(0xffffffffu, state >> 8, slot * lpn) % 1500u;
            if (MODE == 0) r = (r & ~1u) | (uint32_t)(slot & 1);
            if (MODE == 2) r = (r & ~1u);
            const uint32_t row_bytes = (MODE == 3) ? 128u : 64u;
            if (MODE == 3) r %= 750u;
            uint4 v;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + r * row_bytes + q * 16u));
            acc += __uint_as_float(v.x) + __uint_as_float(v.w);
        }
    }
    unsigned long long t1 = clock64();
    if (threadIdx.x == 0) out_clk[blockIdx.x] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
}

// ---- 2. math --------------------------------------------------------------------------------------------------
// mode 0: FFMA2 only (FP32 operand); 1: 2x HADD2.F32 + FFMA2 per half2; 2: 2x FHFMA per half2
template <int MODE>
__global__ void __launch_bounds__(512, 1) math_kernel(int iters, unsigned long long* out_clk, float* sink, const uint32_t* src) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = src[(threadIdx.x * 4 + i) & 1023];
    float v = __uint_as_float(src[threadIdx.x & 1023] | 0x3f000000u);
    unsigned long long acc2[8];
    float acc[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc2[i] = 0ull;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    const unsigned short vh = __half_as_ushort(__float2half(v));
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {  // 8 "non-zeros" x 8 factors
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                w[p] += 0x00010001u;  // loop-carried: the conversions cannot be hoisted or shared
                const uint32_t x = w[p];
                if (MODE == 0) {
                    unsigned long long ww, vv;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(ww) : "r"(x), "r"(x ^ 0x00010000u));
                    asm("mov.b64 %0, {%1, %1};" : "=l"(vv) : "f"(v));
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[p + 4 * (u & 1)]) : "l"(ww), "l"(vv));
                } else if (MODE == 1) {
                    unsigned short lo, hi;
                    asm("mov.b32 {%0, %1}, %2;" : "=h"(lo), "=h"(hi) : "r"(x));
                    float flo, fhi;
                    asm volatile("cvt.f32.f16 %0, %1;" : "=f"(flo) : "h"(lo));
                    asm volatile("cvt.f32.f16 %0, %1;" : "=f"(fhi) : "h"(hi));
                    unsigned long long ww, vv;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(ww) : "f"(flo), "f"(fhi));
                    asm("mov.b64 %0, {%1, %1};" : "=l"(vv) : "f"(v));
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[p + 4 * (u & 1)]) : "l"(ww), "l"(vv));
                } else {
                    unsigned short lo, hi;
                    asm("mov.b32 {%0, %1}, %2;" : "=h"(lo), "=h"(hi) : "r"(x));
                    asm volatile("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc[2 * p + 8 * (u & 1)]) : "h"(lo), "h"(vh));
                    asm volatile("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc[2 * p + 1 + 8 * (u & 1)]) : "h"(hi), "h"(vh));
                }
            }
        }
    }
    unsigned long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += __uint_as_float((uint32_t)acc2[i]) + __uint_as_float((uint32_t)(acc2[i] >> 32));
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (threadIdx.x == 0) out_clk[blockIdx.x] = t1 - t0;
    if (s == 123.456f) sink[0] = s;
}

// ---- 3. TMEM gather -------------------------------------------------------------------------------------------
// every warp fills its 32-lane quarter (32 lanes x 512 columns) and then reads single columns at data-dependent
// addresses: lane = factor, column = gathered row. One tcgen05.ld.32x32b.x1 moves 128 bytes.
__global__ void __launch_bounds__(512, 1) tmem_kernel(int iters, int warps_active, unsigned long long* out_clk, float* sink) {
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    if (warp < 4) {  // fill: column c of lane l holds c + l / 64
        for (int c = 0; c < 512; ++c) {
            const uint32_t val = __float_as_uint((float)c + (float)(threadIdx.x & 31) / 64.f);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tb + c), "r"(val));
        }
        asm volatile("tcgen05.wait::st.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    float acc = 0.f;
    uint32_t state = warp * 2654435761u + blockIdx.x * 40503u + 17u;
    unsigned long long t0 = clock64();
    if (warp < warps_active) {
        for (int it = 0; it < iters; ++it) {
            uint32_t r[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                state = state * 1664525u + 1013904223u;
                const uint32_t col = (state >> 10) & 511u;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[u]) : "r"(tb + col));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += __uint_as_float(r[u]);
        }
    }
    unsigned long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) atomicMax(&out_clk[blockIdx.x], t1 - t0);
    if (acc == 123.456f) sink[0] = acc;
    // correctness probe: column 100 of every quarter
    uint32_t probe;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(probe) : "r"(tb + 100u));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    if (blockIdx.x == 0 && threadIdx.x == 37) sink[1] = __uint_as_float(probe);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}


// ---- 4. Gram correction: FFMA rank-1 updates vs mma.sync SYRK ------------------------------------------------------
// Work unit = one held-out row of one column (k^2 = 1024 multiply-adds). Rows are staged in shared memory the way the
// solver's cp.async ring leaves them ([entry][KP] floats, padded to 40 floats per entry for the fragment reads).
// MODE 0: FP32 FFMA, 8 lanes per column (4 columns per warp), lane owns 4 rows of the 32 x 32 matrix
// MODE 1: TF32 mma.sync m16n8k8, one column per warp, 8 entries per block, single pass (8 LDS + 8 CVT + 8 MMA)
// MODE 2: 3xTF32 (hi*hi + hi*lo + lo*hi): 24 MMA per block
// MODE 3: BF16 mma.sync m16n8k16, 16 entries per block, single pass
template <int MODE>
__global__ void __launch_bounds__(512, 1) gramcorr_kernel(int iters, int warps_active, unsigned long long* out_clk, float* sink) {
    __shared__ __align__(16) float stage[16][16 * 40];  // per warp: 16 entries x 40 floats
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = lane; i < 16 * 40; i += 32) stage[warp][i] = 1.0f + (float)((i * 7 + warp) % 13) * 0.0625f;
    __syncwarp();
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    float acc2[96];
    if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 96; ++i) acc2[i] = 0.f;
    }
    const int g = lane >> 2, t = lane & 3;
    unsigned long long t0 = clock64();
    if (warp < warps_active) {
        for (int it = 0; it < iters; ++it) {
            if (MODE == 0) {
                // 4 columns per warp, 8 lanes each: per held-out row 8 broadcast LDS.128 + 1 own LDS.128 + 128 FFMA
                const int gi = lane >> 3, lig = lane & 7;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float4* st4 = reinterpret_cast<const float4*>(&stage[warp][(e * 4 + gi) * 40]);
                    float f[32];
#pragma unroll
                    for (int i4 = 0; i4 < 8; ++i4) {
                        const float4 v = st4[i4];
                        f[4 * i4] = v.x; f[4 * i4 + 1] = v.y; f[4 * i4 + 2] = v.z; f[4 * i4 + 3] = v.w;
                    }
                    const float4 mine = st4[lig];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        acc[i] = fmaf(mine.x, f[i], acc[i]);
                        acc2[i] = fmaf(mine.y, f[i], acc2[i]);
                        acc2[32 + i] = fmaf(mine.z, f[i], acc2[32 + i]);
                        acc2[64 + i] = fmaf(mine.w, f[i], acc2[64 + i]);
                    }
                }
            } else if (MODE == 1 || MODE == 2) {
                float v[4], u[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) { v[q] = stage[warp][t * 40 + g + 8 * q]; u[q] = stage[warp][(t + 4) * 40 + g + 8 * q]; }
                uint32_t vh[4], uh[4], vl[4], ul[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(vh[q]) : "f"(v[q]));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(uh[q]) : "f"(u[q]));
                    if (MODE == 2) {
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(vl[q]) : "f"(v[q] - __uint_as_float(vh[q])));
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(ul[q]) : "f"(u[q] - __uint_as_float(uh[q])));
                    }
                }
#pragma unroll
                for (int p = 0; p < 2; ++p)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float* d = &acc[(p * 4 + q) * 4];
                        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                                     : "r"(vh[2 * p]), "r"(vh[2 * p + 1]), "r"(uh[2 * p]), "r"(uh[2 * p + 1]), "r"(vh[q]), "r"(uh[q]));
                        if (MODE == 2) {
                            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                         : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                                         : "r"(vh[2 * p]), "r"(vh[2 * p + 1]), "r"(uh[2 * p]), "r"(uh[2 * p + 1]), "r"(vl[q]), "r"(ul[q]));
                            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                         : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                                         : "r"(vl[2 * p]), "r"(vl[2 * p + 1]), "r"(ul[2 * p]), "r"(ul[2 * p + 1]), "r"(vh[q]), "r"(uh[q]));
                        }
                    }
            } else {
                // BF16 m16n8k16: A frag 4 regs (2 bf16 each), B frag 2 regs; 16 entries per block
                uint32_t a[2][4], b[4][2];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float x0 = stage[warp][(2 * t) * 40 + g + 8 * q], x1 = stage[warp][(2 * t + 1) * 40 + g + 8 * q];
                    const float y0 = stage[warp][(2 * t + 8) * 40 + g + 8 * q], y1 = stage[warp][(2 * t + 9) * 40 + g + 8 * q];
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(b[q][0]) : "f"(x1), "f"(x0));
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(b[q][1]) : "f"(y1), "f"(y0));
                }
#pragma unroll
                for (int p = 0; p < 2; ++p) { a[p][0] = b[2 * p][0]; a[p][1] = b[2 * p + 1][0]; a[p][2] = b[2 * p][1]; a[p][3] = b[2 * p + 1][1]; }
#pragma unroll
                for (int p = 0; p < 2; ++p)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float* d = &acc[(p * 4 + q) * 4];
                        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                                     : "r"(a[p][0]), "r"(a[p][1]), "r"(a[p][2]), "r"(a[p][3]), "r"(b[q][0]), "r"(b[q][1]));
                    }
            }
        }
    }
    unsigned long long t1 = clock64();
    if (lane == 0) atomicMax(&out_clk[blockIdx.x], t1 - t0);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i];
    if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 96; ++i) s += acc2[i];
    }
    if (s == 123.456f) sink[0] = s;
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    unsigned long long* d_clk;
    float* d_sink;
    uint32_t* d_src;
    CK(cudaMalloc(&d_clk, sizeof(unsigned long long) * sms));
    CK(cudaMalloc(&d_sink, 16));
    CK(cudaMalloc(&d_src, 4096));
    std::vector<uint32_t> src(1024);
    for (int i = 0; i < 1024; ++i) src[i] = 0x3c003c00u + (uint32_t)i * 0x00010001u;
    CK(cudaMemcpy(d_src, src.data(), 4096, cudaMemcpyHostToDevice));
    std::vector<unsigned long long> clk(sms);
    auto report = [&](const char* name, double per_cta_ops, double bytes_per_op) {
        cudaMemcpy(clk.data(), d_clk, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
        double mx = 0;
        for (int i = 0; i < sms; ++i) mx = clk[i] > mx ? (double)clk[i] : mx;
        printf("%-44s %10.0f clk  %7.3f clk/warp-op  %7.1f B/clk/SM\n", name, mx, mx / per_cta_ops, bytes_per_op * per_cta_ops / mx);
    };
    const int iters = 2000;
    const size_t smem = 96 * 1024;
    CK(cudaFuncSetAttribute(gather_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(gather_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(gather_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(gather_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int rep = 0; rep < 2; ++rep) {
        const double ops = 16.0 * iters * 8;  // warp-level LDS.128 per CTA
        gather_kernel<0><<<sms, 512, smem>>>(iters, d_clk, d_sink); CK(cudaDeviceSynchronize());
        if (rep) report("LDS.128 64B rows, opposite halves", ops, 512);
        gather_kernel<1><<<sms, 512, smem>>>(iters, d_clk, d_sink); CK(cudaDeviceSynchronize());
        if (rep) report("LDS.128 64B rows, random halves", ops, 512);
        gather_kernel<2><<<sms, 512, smem>>>(iters, d_clk, d_sink); CK(cudaDeviceSynchronize());
        if (rep) report("LDS.128 64B rows, same half", ops, 512);
        gather_kernel<3><<<sms, 512, smem>>>(iters, d_clk, d_sink); CK(cudaDeviceSynchronize());
        if (rep) report("LDS.128 128B rows (FP32 layout)", ops, 512);
    }
    for (int rep = 0; rep < 2; ++rep) {
        const double ops = 16.0 * iters * 8;  // warp-level "non-zero x 8 factors" units per CTA
        math_kernel<0><<<sms, 512>>>(iters, d_clk, d_sink, d_src); CK(cudaDeviceSynchronize());
        if (rep) report("8 factors: 4 FFMA2 (FP32 operand)", ops, 0);
        math_kernel<1><<<sms, 512>>>(iters, d_clk, d_sink, d_src); CK(cudaDeviceSynchronize());
        if (rep) report("8 factors: 8 HADD2.F32 + 4 FFMA2", ops, 0);
        math_kernel<2><<<sms, 512>>>(iters, d_clk, d_sink, d_src); CK(cudaDeviceSynchronize());
        if (rep) report("8 factors: 8 FHFMA (f16 value)", ops, 0);
    }
    for (int wa : {1, 4, 8, 16}) {
        CK(cudaMemset(d_clk, 0, sizeof(unsigned long long) * sms));
        tmem_kernel<<<sms, 512>>>(iters, wa, d_clk, d_sink); CK(cudaDeviceSynchronize());
        CK(cudaMemset(d_clk, 0, sizeof(unsigned long long) * sms));
        tmem_kernel<<<sms, 512>>>(iters, wa, d_clk, d_sink); CK(cudaDeviceSynchronize());
        char name[64];
        snprintf(name, sizeof(name), "tcgen05.ld.32x32b.x1 dynamic column, %2d warps", wa);
        report(name, (double)wa * iters * 8, 128);
    }
    // 4. Gram correction: clk per (column, held-out row) unit per SM
    for (int wa : {4, 8, 16}) {
        auto run = [&](auto kern, const char* nm, double units_per_iter) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaMemset(d_clk, 0, sizeof(unsigned long long) * sms);
                kern<<<sms, 512>>>(iters, wa, d_clk, d_sink);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(clk.data(), d_clk, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
            double mx = 0;
            for (int i = 0; i < sms; ++i) mx = clk[i] > mx ? (double)clk[i] : mx;
            printf("%-40s %2d warps %10.0f clk  %7.3f clk per column-row per SM  (%6.1f FMA/clk/SM)\n", nm, wa, mx,
                   mx / (units_per_iter * wa * iters), 1024.0 * units_per_iter * wa * iters / mx);
        };
        run(gramcorr_kernel<0>, "Gram corr: FFMA, 8 lanes/column", 8.0);
        run(gramcorr_kernel<1>, "Gram corr: mma.sync TF32 1-pass", 8.0);
        run(gramcorr_kernel<2>, "Gram corr: mma.sync 3xTF32", 8.0);
        run(gramcorr_kernel<3>, "Gram corr: mma.sync BF16 1-pass", 16.0);
    }
    float probe[4];
    CK(cudaMemcpy(probe, d_sink, 16, cudaMemcpyDeviceToHost));
    printf("tmem probe (expect 100 + 5/64 = 100.078): %.4f\n", probe[1]);
    return 0;
}
