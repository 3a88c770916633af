// Host emulation of the handful of CUDA constructs our SIMT kernels use, so that their indexing / reductions can be
// checked against the oracle WITHOUT a GPU (test infrastructure only -- never linked into the product library).
// The CTAs of one cluster run concurrently, clusters one after the other; every CUDA thread is a pthread; __syncthreads / warp collectives are pthread barriers, so a
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
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };

namespace emu {
struct Cluster;
struct Cta {
    pthread_barrier_t block_bar;
    pthread_barrier_t warp_bar[32];
    uint32_t xchg[32][32];
    int nthreads = 0;
    unsigned char* smem = nullptr;     // this CTA's dynamic shared memory (1024-byte aligned)
    size_t smem_bytes = 0;
    int rank = 0;                      // rank in the cluster
    unsigned block_index = 0;
    Cluster* cluster = nullptr;
    float* tmem = nullptr;             // [128 lanes][512 columns] (emu_tc.h)
};
struct Cluster {
    std::vector<Cta*> ctas;
    pthread_barrier_t bar;             // every thread of every CTA of the cluster
};
inline thread_local Cta* t_cta = nullptr;
inline Cta*& cta() { return t_cta; }
inline emu_dim3& block_dim() { static emu_dim3 d; return d; }
inline emu_dim3& grid_dim() { static emu_dim3 d; return d; }
inline unsigned char* dyn_smem() { return t_cta->smem; }
inline thread_local emu_dim3 t_threadIdx, t_blockIdx;
inline std::function<void()>& after_launch_hook() { static std::function<void()> f; return f; }   // joins async engines (emu_tc.h)

// runs kernel() for every thread of every block; the `cluster` CTAs of a cluster run concurrently, clusters sequentially
inline void launch_cluster(int grid, int cluster, int block, size_t smem_bytes, const std::function<void()>& kernel) {
    if (grid % cluster) abort();
    grid_dim().x = (unsigned)grid;
    block_dim().x = (unsigned)block;
    const int nwarps = (block + 31) / 32;
    for (int b0 = 0; b0 < grid; b0 += cluster) {
        Cluster cl;
        pthread_barrier_init(&cl.bar, nullptr, (unsigned)(block * cluster));
        std::vector<Cta> ctas(cluster);
        std::vector<std::vector<float>> tmem(cluster);
        for (int r = 0; r < cluster; ++r) {
            Cta& c = ctas[r];
            c.nthreads = block;
            c.rank = r;
            c.block_index = (unsigned)(b0 + r);
            c.cluster = &cl;
            void* mem = nullptr;
            if (posix_memalign(&mem, 1024, smem_bytes + 1024)) abort();
            memset(mem, 0xAB, smem_bytes + 1024);
            c.smem = (unsigned char*)mem;
            c.smem_bytes = smem_bytes;
            tmem[r].assign((size_t)128 * 512, std::nanf(""));
            c.tmem = tmem[r].data();
            pthread_barrier_init(&c.block_bar, nullptr, (unsigned)block);
            for (int w = 0; w < nwarps; ++w) pthread_barrier_init(&c.warp_bar[w], nullptr, (unsigned)std::min(32, block - 32 * w));
            cl.ctas.push_back(&c);
        }
        std::vector<std::thread> ts;
        ts.reserve((size_t)block * cluster);
        for (int r = 0; r < cluster; ++r)
            for (int t = 0; t < block; ++t)
                ts.emplace_back([&, t, r]() {
                    t_cta = &ctas[r];
                    t_threadIdx.x = (unsigned)t;
                    t_blockIdx.x = ctas[r].block_index;
                    kernel();
                });
        for (auto& t : ts) t.join();
        if (after_launch_hook()) after_launch_hook()();
        for (int r = 0; r < cluster; ++r) {
            pthread_barrier_destroy(&ctas[r].block_bar);
            for (int w = 0; w < nwarps; ++w) pthread_barrier_destroy(&ctas[r].warp_bar[w]);
            free(ctas[r].smem);
        }
        pthread_barrier_destroy(&cl.bar);
    }
}
inline void launch(int grid, int block, size_t smem_bytes, const std::function<void()>& kernel) {
    launch_cluster(grid, 1, block, smem_bytes, kernel);
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
#define __shared__ static   /* CTAs of a launch run one after the other here, so one static array per kernel is a CTA's shared memory */

static inline void __syncthreads() { pthread_barrier_wait(&emu::cta()->block_bar); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_sync(); }
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int off) { return emu::warp_read(v, emu::lane_id() ^ off); }
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src) { return emu::warp_read(v, src); }
template <typename T>
static inline T __shfl_up_sync(unsigned, T v, int delta) {
    const int src = emu::lane_id() - delta;
    const T r = emu::warp_read(v, src < 0 ? emu::lane_id() : src);
    return src < 0 ? v : r;
}
static inline unsigned __match_any_sync(unsigned, uint32_t v) {
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= (emu::warp_read<uint32_t>(v, l) == v ? 1u : 0u) << l;
    return m;
}
static inline unsigned __ballot_sync(unsigned, bool p) {
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= (emu::warp_read<uint32_t>(p ? 1u : 0u, l) & 1u) << l;
    return m;
}
static inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
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
#define PR_DYN_SMEM_BYTES(name) unsigned char* name = emu::dyn_smem()
