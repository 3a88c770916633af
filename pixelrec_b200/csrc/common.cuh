// pixelrec_b200 -- shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pixelrec_b200.h"

namespace pr {

// ---------------------------------------------------------------- error plumbing (api.cu)
void set_last_error(const char* fmt, ...);
int sm_count();
int tune();                      // PR_TUNE bit mask (api.cu)
// Device-resident dropout seed offset (pr_set_seed_device; CUDA-graph replay: the host seed is frozen into the graph, the
// offset is bumped by a captured kernel).  Only builds with -DPR_SEED_DEV read it; the default build's kernels are unchanged.
const unsigned long long* seed_device();
#ifdef PR_SEED_DEV
#define PR_SEED(a) ((a).seed_dev ? (a).seed + *(a).seed_dev : (a).seed)
#define PR_SET_SEED_DEV(a) (a).seed_dev = pr::seed_device()
#else
#define PR_SEED(a) ((a).seed)
#define PR_SET_SEED_DEV(a) (void)0
#endif
#define PR_TUNE_LN_BWD_PIPE 1
#define PR_TUNE_LN_L2_PREFETCH 2
#define PR_TUNE_LN_FWD_PIPE 4
#define PR_TUNE_ATTN_PAIR 8
#define PR_TUNE_SCORE_V2 16      /* score_topk: branch-free 8-warp epilogue (r02: 0.497 -> 0.208 ms, profiles/r02a_bench_score.json) */
#define PR_TUNE_ATTN_LONG_TC 64  /* long-sequence attention forward on mma.sync TF32 (r02: 3.49 -> 0.77 ms at ViT-B/16 shape) */
#define PR_TUNE_SCORE_ARES 128    /* fp16 scoring: seq_out tile resident in shared memory (r02: no gain, off) */
#define PR_TUNE_SCORE_MCAST 32   /* + table tile TMA-multicast across the m-tiles of a cluster (r02: slower than plain v2, off) */
#define PR_TUNE_SCATTER_RING 256  /* scatter_add_rows as a TMA-staged ring (rows_ring.cuh) instead of the LDG warp-per-run kernel */
#define PR_TUNE_LN_FWD_ROWS2 512   /* LayerNorm forward (register kernel) with two rows in flight per warp */
#ifndef PR_TUNE_DEFAULT
/* measured: profiles/r01h_rowkernels_ab.md, r01l_attention_pair.md, r02a_bench_score.json, r02a_bench_attn_long.json,
   r02j_scatter_ab.json, r02k_bench_n1_rows2.json */
#define PR_TUNE_DEFAULT (PR_TUNE_LN_BWD_PIPE | PR_TUNE_ATTN_PAIR | PR_TUNE_SCORE_V2 | PR_TUNE_ATTN_LONG_TC | PR_TUNE_SCATTER_RING | PR_TUNE_LN_FWD_ROWS2)
#endif

#define PR_CHECK_ARG(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            pr::set_last_error(__VA_ARGS__);    \
            return PR_ERR_INVALID_ARGUMENT;     \
        }                                       \
    } while (0)

#define PR_CUDA_LAUNCH_CHECK(name)                                                         \
    do {                                                                                   \
        cudaError_t e__ = cudaGetLastError();                                              \
        if (e__ != cudaSuccess) {                                                          \
            pr::set_last_error("%s: %s", name, cudaGetErrorString(e__));                   \
            return (int)e__;                                                               \
        }                                                                                  \
    } while (0)

#define PR_CUDA_CALL(expr)                                                                 \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            pr::set_last_error("%s: %s", #expr, cudaGetErrorString(e__));                  \
            return (int)e__;                                                               \
        }                                                                                  \
    } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------- vector ld/st
__device__ __forceinline__ float4 ldg_stream(const float4* p) {  // read-once data: skip L1
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------- mbarrier / bulk-copy (TMA engine, 1-D) PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 4-byte asynchronous copy global -> shared (LDGSTS): no register holds the value, so nothing can stall on it before it is read back
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// global -> shared bulk copy (UBLKCP); bytes % 16 == 0, both addresses 16-B aligned; completes on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
// L2 prefetch of a contiguous span (no smem, no completion tracking); bytes % 16 == 0, 16-B aligned
__device__ __forceinline__ void l2_prefetch_bulk(const void* gmem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {  // smem of all but the N most recent groups may be reused
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------- Philox4x32-10 (dropout masks; oracle/philox_np.py restates it)
struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
    __device__ __forceinline__ uint4 operator()(uint64_t ctr, uint32_t stream) const {
        uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = stream, c3 = 0;
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            const uint32_t n0 = hi1 ^ c1 ^ a, n2 = hi0 ^ c3 ^ b;
            c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};
// dropout: an element is kept iff its 16-bit random field >= thr, thr = round(p * 2^16) (p in [0,1)).
// One Philox call (128 bits) serves 8 elements.
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
    double t = (double)p * 65536.0 + 0.5;
    if (t < 0) t = 0;
    if (t > 65535.0) t = 65535.0;
    return (uint32_t)t;
}
// 8 keep bits from one Philox draw: bit e (e<4) <- low half of word e, bit 4+e <- high half of word e
__device__ __forceinline__ uint32_t keep_bits8(const uint4& r, uint32_t thr) {
    uint32_t m = 0;
    m |= ((r.x & 0xffffu) >= thr) ? 1u : 0u;
    m |= ((r.y & 0xffffu) >= thr) ? 2u : 0u;
    m |= ((r.z & 0xffffu) >= thr) ? 4u : 0u;
    m |= ((r.w & 0xffffu) >= thr) ? 8u : 0u;
    m |= ((r.x >> 16) >= thr) ? 16u : 0u;
    m |= ((r.y >> 16) >= thr) ? 32u : 0u;
    m |= ((r.z >> 16) >= thr) ? 64u : 0u;
    m |= ((r.w >> 16) >= thr) ? 128u : 0u;
    return m;
}

}  // namespace pr
