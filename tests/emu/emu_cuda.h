// Host emulation of the handful of CUDA constructs our SIMT kernels use, so that their indexing / reductions can be
// checked against the oracle WITHOUT a GPU (test infrastructure only -- never linked into the product library).
// One CTA at a time; every CUDA thread is a pthread; __syncthreads / warp collectives are pthread barriers, so a
// collective reached by only part of a warp deadlocks here just as it would be undefined on the device.
#pragma once
#include <pthread.h>

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };

namespace emu {
struct Cta {
    pthread_barrier_t block_bar;
    pthread_barrier_t warp_bar[32];
    uint32_t xchg[32][32];
    int nthreads;
};
inline Cta*& cta() { static Cta* c = nullptr; return c; }
inline emu_dim3& block_dim() { static emu_dim3 d; return d; }
inline emu_dim3& grid_dim() { static emu_dim3 d; return d; }
inline unsigned char*& dyn_smem() { static unsigned char* p = nullptr; return p; }
inline thread_local emu_dim3 t_threadIdx, t_blockIdx;

// runs kernel() for every thread of every block; blocks sequentially
inline void launch(int grid, int block, size_t smem_bytes, const std::function<void()>& kernel) {
    grid_dim().x = (unsigned)grid;
    block_dim().x = (unsigned)block;
    void* mem = nullptr;
    if (posix_memalign(&mem, 1024, smem_bytes + 1024)) abort();
    dyn_smem() = (unsigned char*)mem;
    Cta c;
    c.nthreads = block;
    cta() = &c;
    const int nwarps = (block + 31) / 32;
    for (int b = 0; b < grid; ++b) {
        pthread_barrier_init(&c.block_bar, nullptr, (unsigned)block);
        for (int w = 0; w < nwarps; ++w) pthread_barrier_init(&c.warp_bar[w], nullptr, (unsigned)std::min(32, block - 32 * w));
        std::vector<std::thread> ts;
        ts.reserve(block);
        for (int t = 0; t < block; ++t)
            ts.emplace_back([&, t, b]() {
                t_threadIdx.x = (unsigned)t;
                t_blockIdx.x = (unsigned)b;
                kernel();
            });
        for (auto& t : ts) t.join();
        pthread_barrier_destroy(&c.block_bar);
        for (int w = 0; w < nwarps; ++w) pthread_barrier_destroy(&c.warp_bar[w]);
    }
    free(mem);
    dyn_smem() = nullptr;
}
inline int warp_id() { return (int)(t_threadIdx.x >> 5); }
inline int lane_id() { return (int)(t_threadIdx.x & 31); }
inline void warp_sync() { pthread_barrier_wait(&cta()->warp_bar[warp_id()]); }
template <typename T>
inline T warp_read(T v, int src_lane) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    uint32_t bits;
    memcpy(&bits, &v, 4);
    cta()->xchg[warp_id()][lane_id()] = bits;
    warp_sync();
    const uint32_t r = cta()->xchg[warp_id()][src_lane & 31];
    warp_sync();
    T out;
    memcpy(&out, &r, 4);
    return out;
}
}  // namespace emu

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::block_dim())
#define gridDim (emu::grid_dim())
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

static inline void __syncthreads() { pthread_barrier_wait(&emu::cta()->block_bar); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_sync(); }
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int off) { return emu::warp_read(v, emu::lane_id() ^ off); }
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src) { return emu::warp_read(v, src); }
static inline unsigned __ballot_sync(unsigned, bool p) {
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= (emu::warp_read<uint32_t>(p ? 1u : 0u, l) & 1u) << l;
    return m;
}
static inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <typename T>
static inline T __ldg(const T* p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }

#define PR_LDG4(p) (*(p))
#define PR_LDG4_STREAM(p) (*(p))
#define PR_DYN_SMEM_F4(name) float4* name = reinterpret_cast<float4*>(emu::dyn_smem())
