// common.cuh -- shared device helpers for libsinglet_cuda (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/singlet_cuda.h"

namespace sgl {

// ----------------------------------------------------------------------------------------------
// error plumbing (no exceptions across the C ABI)
// ----------------------------------------------------------------------------------------------
std::string& last_error();
int fail(int code, const char* fmt, ...);

#define SGL_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return ::sgl::fail(SGL_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                               __LINE__);                                                     \
    } while (0)

#define SGL_TRY(expr)            \
    do {                         \
        int _r = (expr);         \
        if (_r != SGL_OK) return _r; \
    } while (0)

static inline int kp_of(int k) {
    int kp = 4;
    while (kp < k) kp <<= 1;
    return kp;
}

// ----------------------------------------------------------------------------------------------
// speckled-mask hash: rng::rand(i), rng::rand(i,j), rng::draw  (reference src/singlet.cpp:30-64,
// 91-95). uint64 wrap-around arithmetic, bit-exact by construction.
// The first argument is always the CELL index, the second the GENE index (App. A-9).
// ----------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t hash_cell(uint64_t state, uint64_t i) {
    i ^= i << 19;
    i ^= i >> 7;
    i ^= i << 36;
    uint64_t x = state + i;
    x ^= x << 38;
    x ^= x >> 13;
    x ^= x << 23;
    return x;
}
__host__ __device__ __forceinline__ uint64_t premix_gene(uint64_t j) {
    j ^= j >> 7;
    j ^= j << 23;
    j ^= j >> 8;
    return j;
}
__host__ __device__ __forceinline__ uint64_t hash_finish(uint64_t cell_hash, uint64_t gene_mix) {
    uint64_t x = cell_hash + gene_mix;
    x ^= x >> 7;
    x ^= x << 53;
    x ^= x >> 4;
    return x;
}
__host__ __device__ __forceinline__ uint64_t hash_pair(uint64_t state, uint64_t cell, uint64_t gene) {
    return hash_finish(hash_cell(state, cell), premix_gene(gene));
}

// x % p == 0 for a runtime p. For p < 2^16 (the usual case: p = round(1/test_density)) everything
// stays in 32-bit arithmetic: x = hi*2^32 + lo, 2^32 mod p = c32 precomputed on the host.
struct ModP {
    uint64_t p;
    uint32_t p32, c32;  // c32 = 2^32 mod p (valid when small)
    int small;
};
static inline ModP make_modp(uint64_t p) {
    ModP m;
    m.p = p;
    m.small = (p < 65536ull) ? 1 : 0;
    m.p32 = (uint32_t)p;
    m.c32 = m.small ? (uint32_t)((1ull << 32) % p) : 0u;
    return m;
}
__device__ __forceinline__ bool is_multiple(uint64_t x, const ModP& m) {
    if (m.small) {
        const uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x;
        // (hi mod p) * c32 < 2^32 because both factors are < 2^16
        const uint32_t r = ((hi % m.p32) * m.c32) % m.p32 + (lo % m.p32);
        return (r % m.p32) == 0u;
    }
    return (x % m.p) == 0ull;
}

// ----------------------------------------------------------------------------------------------
// synthetic counts generator (stratified so that it is O(nnz); the same rules are
// restated in numpy in singlet_b200/synth.py). Column c (cell), stratum s of S consecutive genes:
//   u = splitmix64(seed ^ (c << 32 | s)); present iff low32(u) < q32; gene = s*S + (bits 32..47 of u) % S;
//   value = table[min(7, ctz(bits 48..63 of u | 0x80))].
// ----------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
struct SynthSpec {
    int64_t m, n;
    uint64_t seed;
    uint32_t S, q32;
    float table[8];
};
// returns true when (cell, stratum) holds a non-zero; gene/value receive its row and value
__device__ __forceinline__ bool synth_entry(const SynthSpec& sp, uint64_t cell, uint32_t stratum, int64_t& gene,
                                            float& value) {
    const uint64_t u = splitmix64(sp.seed ^ ((cell << 32) | (uint64_t)stratum));
    if ((uint32_t)u >= sp.q32) return false;
    gene = (int64_t)stratum * sp.S + (int64_t)(((uint32_t)(u >> 32) & 0xFFFFu) % sp.S);
    if (gene >= sp.m) return false;
    const uint32_t hi = (uint32_t)(u >> 48) | 0x80u;
    value = sp.table[__ffs(hi) - 1];
    return true;
}

// ----------------------------------------------------------------------------------------------
// PTX: mbarrier + bulk async copy (TMA, global -> shared), used by the SpMM to stage factor tiles
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 16-byte async copy global -> shared (LDGSTS), bypassing L1; src_bytes = 0 zero-fills
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// 8-byte variant (through L1); src_bytes = 0 zero-fills
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}


__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared; bytes % 16 == 0, both addresses 16-byte aligned. SASS: UBLKCP.
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ uint2 ldg_stream_u2(const uint2* p) {
    uint2 r;
    asm("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

// ask L2 to fetch `count` 8-byte records starting at p (rounded out to 16-byte granules)
__device__ __forceinline__ void l2_prefetch_records(const uint2* p, int32_t count) {
    if (count <= 0) return;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uintptr_t a0 = a & ~(uintptr_t)15;
    const uint32_t bytes = (uint32_t)(((a + (uintptr_t)count * 8 + 15) & ~(uintptr_t)15) - a0);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"(bytes) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace sgl
