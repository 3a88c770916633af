// pixelrec_b200 -- peer-memory (NVLink / NVSwitch P2P) exchange of the row-sharded item table.
//   The reference replicates the table on every GPU and all-reduces a dense [N,D] gradient (DDP, REC/run.py:40 around
//   nn.Embedding, REC/model/IDNet/sasrec.py:31,68).  Here the table is row-sharded (owner(i) = i % G, local row i / G)
//   and the two exchange steps are ONE kernel each over peer-mapped memory -- no staging buffers, no all_to_all:
//     pr_gather_rows_peers_f32  lookup + exchange: every row is read straight out of its owner's shard over NVLink
//     pr_push_rows_peers_f32    gradient rows are written straight into the owner's receive region of this rank,
//                               together with their local row ids; the owner then runs the ordinary
//                               pr_scatter_plan + pr_scatter_add_rows_f32 over its receive buffer
//   pr_shared_{alloc,open,close,free}: cudaMalloc + CUDA IPC handles, so that one process per GPU can map its peers'
//   shards and receive buffers (cudaIpcOpenMemHandle with lazy peer access).
// Byte movers: 128-bit accesses, several independent loads in flight per lane (NVLink round trips are ~2-4 us).
#include <algorithm>
#include <string.h>

#include "common.cuh"

#define PR_LDG4_STREAM(p) pr::ldg_stream(p)
#include "peer.cuh"

using namespace pr;

extern "C" int pr_gather_rows_peers_f32(const float* const* shards, int G, int64_t N, int64_t D, const int64_t* idx, int64_t R,
                                        float* out, int32_t* status, pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PR_CHECK_ARG(G >= 1 && N > 0 && D > 0 && R >= 0, "pr_gather_rows_peers_f32: bad shape G=%d N=%lld D=%lld R=%lld", G,
                 (long long)N, (long long)D, (long long)R);
    PR_CHECK_ARG(D % 4 == 0, "pr_gather_rows_peers_f32: D=%lld must be a multiple of 4", (long long)D);
    if (R == 0) return PR_OK;
    PR_CHECK_ARG(shards && idx && out, "pr_gather_rows_peers_f32: null pointer");
    PR_CHECK_ARG(aligned16(out), "pr_gather_rows_peers_f32: out must be 16-byte aligned");
    const long long ngroups = (R + PG_ROWS - 1) / PG_ROWS;
    const int grid = (int)std::min<long long>((ngroups + 7) / 8, (long long)sm_count() * 8);
    gather_rows_peers_kernel<<<grid, 256, 0, stream>>>((const float4* const*)shards, G, N, (int)(D / 4), (const long long*)idx, R,
                                                       (float4*)out, status);
    PR_CUDA_LAUNCH_CHECK("gather_rows_peers_kernel");
    return PR_OK;
}

extern "C" int pr_push_rows_peers_f32(const float* rows, const int64_t* ids, int64_t U, int64_t D, int G, int rank, int64_t cap,
                                      int64_t skip_id, float* const* recv_rows, int64_t* const* recv_ids, int32_t* counters,
                                      int32_t* status, pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PR_CHECK_ARG(G >= 1 && rank >= 0 && rank < G && cap > 0 && U >= 0, "pr_push_rows_peers_f32: bad G=%d rank=%d cap=%lld U=%lld",
                 G, rank, (long long)cap, (long long)U);
    PR_CHECK_ARG(D > 0 && D % 4 == 0, "pr_push_rows_peers_f32: D=%lld must be a positive multiple of 4", (long long)D);
    if (U == 0) return PR_OK;
    PR_CHECK_ARG(rows && ids && recv_rows && recv_ids && counters, "pr_push_rows_peers_f32: null pointer");
    PR_CHECK_ARG(aligned16(rows), "pr_push_rows_peers_f32: rows must be 16-byte aligned");
    const int grid = (int)std::max<long long>(1, std::min<long long>((U + 7) / 8, (long long)sm_count() * 8));
    push_rows_peers_kernel<<<grid, 256, 0, stream>>>((const float4*)rows, (const long long*)ids, U, (int)(D / 4), G, rank, cap,
                                                     skip_id, (float4* const*)recv_rows, (long long* const*)recv_ids, counters,
                                                     status);
    PR_CUDA_LAUNCH_CHECK("push_rows_peers_kernel");
    return PR_OK;
}

// ---- device-side plan variants: everything the exchange needs is derived from ONE pr_scatter_plan of the step's ids, with the
// number of distinct ids left in device memory -- no host synchronisation anywhere, so the whole multi-GPU step can be
// captured into a CUDA graph (the torch.unique-based plan needs the count on the host).
__global__ void __launch_bounds__(256) plan_inverse_kernel(const int* __restrict__ perm, const int* __restrict__ seg_start,
                                                           const int* __restrict__ n_uniq, long long R, long long pad_slot,
                                                           long long* __restrict__ inverse) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= R) return;
    const int U = *n_uniq;
    const int n_valid = seg_start[U];
    long long slot = pad_slot;                       // positions the plan dropped (padding id): the extra row of the pull
    if (p < n_valid) {
        int lo = 0, hi = U - 1;                      // largest u with seg_start[u] <= p
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (seg_start[mid] <= p) lo = mid; else hi = mid - 1;
        }
        slot = lo;
    }
    inverse[perm[p]] = slot;
}

// out[u] = owner's row of uniq_ids[u] for u < *n_uniq; out[pad_slot] = owner's row of pad_id (pad_slot >= 0)
__global__ void __launch_bounds__(256) gather_rows_peers_plan_kernel(const float4* const* __restrict__ shards, int G, long long N,
                                                                     int D4, const int* __restrict__ uniq_ids,
                                                                     const int* __restrict__ n_uniq, long long pad_id,
                                                                     long long pad_slot, float4* __restrict__ out,
                                                                     int* __restrict__ status) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long U = *n_uniq;
    const long long total = U + (pad_slot >= 0 ? 1 : 0);
    const long long ngroups = (total + PG_ROWS - 1) / PG_ROWS;
    for (long long g = warp; g < ngroups; g += nwarps) {
        const long long e0 = g * PG_ROWS;
        const float4* src[PG_ROWS];
        long long dst[PG_ROWS];
#pragma unroll
        for (int j = 0; j < PG_ROWS; ++j) {
            const long long e = e0 + j;
            const bool live = e < total;
            const long long id = !live ? 0 : (e < U ? (long long)__ldg(uniq_ids + e) : pad_id);
            const bool ok = live && id >= 0 && id < N;
            if (live && !ok && status && lane == 0) atomicOr(status, 1);
            src[j] = ok ? (shards[(int)(id % G)] + (id / G) * (long long)D4) : nullptr;
            dst[j] = !live ? -1 : (e < U ? e : pad_slot);
        }
        for (int c = lane; c < D4; c += 32) {
            float4 v[PG_ROWS];
#pragma unroll
            for (int j = 0; j < PG_ROWS; ++j) v[j] = src[j] ? PR_LDG4_STREAM(src[j] + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < PG_ROWS; ++j)
                if (dst[j] >= 0) out[dst[j] * (long long)D4 + c] = v[j];
        }
    }
}

__global__ void __launch_bounds__(256) push_rows_peers_plan_kernel(const float4* __restrict__ rows, const int* __restrict__ ids,
                                                                   const int* __restrict__ n_dev, int D4, int G, int rank,
                                                                   long long cap, float4* const* __restrict__ recv_rows,
                                                                   long long* const* __restrict__ recv_ids,
                                                                   int* __restrict__ counters, int* __restrict__ status) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long U = *n_dev;
    for (long long u = warp; u < U; u += nwarps) {
        const long long id = __ldg(ids + u);
        const int owner = (int)(id % G);
        int pos = 0;
        if (lane == 0) pos = atomicAdd(counters + owner, 1);
        pos = __shfl_sync(0xffffffffu, pos, 0);
        if (pos >= cap) {
            if (status && lane == 0) atomicOr(status, 2);
            continue;
        }
        const long long slot = (long long)rank * cap + pos;
        float4* dst = recv_rows[owner] + slot * (long long)D4;
        const float4* src = rows + u * (long long)D4;
        for (int c = lane; c < D4; c += 128) {
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = (c + 32 * j < D4) ? PR_LDG4_STREAM(src + c + 32 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c + 32 * j < D4) dst[c + 32 * j] = v[j];
        }
        if (lane == 0) recv_ids[owner][slot] = id / G;
    }
}

extern "C" int pr_plan_inverse(const int32_t* perm, const int32_t* seg_start, const int32_t* n_uniq, int64_t R, int64_t pad_slot,
                               int64_t* inverse, pr_stream_t stream_) {
    PR_CHECK_ARG(R >= 0 && (R == 0 || (perm && seg_start && n_uniq && inverse)), "pr_plan_inverse: bad arguments");
    if (R == 0) return PR_OK;
    plan_inverse_kernel<<<(unsigned)((R + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(perm, seg_start, n_uniq, R, pad_slot,
                                                                                         (long long*)inverse);
    PR_CUDA_LAUNCH_CHECK("plan_inverse_kernel");
    return PR_OK;
}

extern "C" int pr_gather_rows_peers_plan_f32(const float* const* shards, int G, int64_t N, int64_t D, const int32_t* uniq_ids,
                                             const int32_t* n_uniq, int64_t max_uniq, int64_t pad_id, int64_t pad_slot, float* out,
                                             int32_t* status, pr_stream_t stream_) {
    PR_CHECK_ARG(G >= 1 && N > 0 && D > 0 && D % 4 == 0 && max_uniq >= 1, "pr_gather_rows_peers_plan_f32: bad shape");
    PR_CHECK_ARG(shards && uniq_ids && n_uniq && out && aligned16(out), "pr_gather_rows_peers_plan_f32: null/unaligned pointer");
    const long long ngroups = (max_uniq + 1 + PG_ROWS - 1) / PG_ROWS;
    const int grid = (int)std::min<long long>((ngroups + 7) / 8, (long long)sm_count() * 8);
    gather_rows_peers_plan_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>((const float4* const*)shards, G, N, (int)(D / 4), uniq_ids,
                                                                            n_uniq, pad_id, pad_slot, (float4*)out, status);
    PR_CUDA_LAUNCH_CHECK("gather_rows_peers_plan_kernel");
    return PR_OK;
}

extern "C" int pr_push_rows_peers_plan_f32(const float* rows, const int32_t* ids, const int32_t* n_dev, int64_t max_n, int64_t D,
                                           int G, int rank, int64_t cap, float* const* recv_rows, int64_t* const* recv_ids,
                                           int32_t* counters, int32_t* status, pr_stream_t stream_) {
    PR_CHECK_ARG(G >= 1 && rank >= 0 && rank < G && cap > 0 && max_n >= 1 && D > 0 && D % 4 == 0,
                 "pr_push_rows_peers_plan_f32: bad arguments");
    PR_CHECK_ARG(rows && ids && n_dev && recv_rows && recv_ids && counters && aligned16(rows), "pr_push_rows_peers_plan_f32: null pointer");
    const int grid = (int)std::max<long long>(1, std::min<long long>((max_n + 7) / 8, (long long)sm_count() * 8));
    push_rows_peers_plan_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>((const float4*)rows, ids, n_dev, (int)(D / 4), G, rank, cap,
                                                                          (float4* const*)recv_rows, (long long* const*)recv_ids,
                                                                          counters, status);
    PR_CUDA_LAUNCH_CHECK("push_rows_peers_plan_kernel");
    return PR_OK;
}

// ---- barrier over peer-mapped flags ----------------------------------------------------------------------------------------
// Orders the peer kernels of all ranks without a collective library call: thread r of one CTA publishes `epoch` in peer r's flag
// array (slot = this rank) with a system-scope release -- after a system fence, so the rows this GPU pushed in earlier kernels
// are visible first -- and spins with system-scope acquire loads on its own array until peer r's flag has reached `epoch`.
// Replaces the two 4-byte NCCL all_reduces per step (0.16 ms each at N = 8, most of it launch latency and rank skew).  A peer
// that never arrives would hang the GPU, so the spin gives up after ~2 s of SM clocks and raises status bit 4.
__global__ void __launch_bounds__(32) peer_barrier_kernel(unsigned long long* const* __restrict__ flags, int G, int rank,
                                                          unsigned long long epoch, unsigned long long* __restrict__ epoch_dev,
                                                          int* __restrict__ status) {
    const int r = threadIdx.x;
    if (epoch_dev) {                                 // CUDA-graph mode: the count lives in device memory, bumped per call
        unsigned long long e = 0;
        if (r == 0) e = atomicAdd(epoch_dev, 1ULL) + 1ULL;
        epoch = __shfl_sync(0xffffffffu, e, 0);
    }
    if (r >= G) return;
    __threadfence_system();
    unsigned long long* theirs = flags[r] + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
    const unsigned long long* mine = flags[rank] + r;
    const long long t0 = clock64();
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
        if (v >= epoch) break;
        if (clock64() - t0 > 4000000000LL) {
            if (status) atomicOr(status, 4);
            break;
        }
        __nanosleep(100);
    }
}

extern "C" int pr_peer_barrier(uint64_t* const* flag_tables, int G, int rank, uint64_t epoch, uint64_t* epoch_dev, int32_t* status,
                               pr_stream_t stream_) {
    PR_CHECK_ARG(flag_tables && G >= 1 && G <= 32 && rank >= 0 && rank < G, "pr_peer_barrier: bad G=%d rank=%d", G, rank);
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream_>>>((unsigned long long* const*)flag_tables, G, rank,
                                                             (unsigned long long)epoch, (unsigned long long*)epoch_dev, status);
    PR_CUDA_LAUNCH_CHECK("peer_barrier_kernel");
    return PR_OK;
}

// ---- shareable device memory (CUDA IPC) ---------------------------------------------------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "pr_shared_* handles are 64 bytes");

extern "C" int pr_shared_alloc(size_t bytes, void** dptr, unsigned char* handle64) {
    PR_CHECK_ARG(bytes > 0 && dptr && handle64, "pr_shared_alloc: bad arguments");
    void* p = nullptr;
    PR_CUDA_CALL(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_last_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
        return (int)e;
    }
    memcpy(handle64, &h, 64);
    *dptr = p;
    return PR_OK;
}

extern "C" int pr_shared_free(void* dptr) {
    if (!dptr) return PR_OK;
    PR_CUDA_CALL(cudaFree(dptr));
    return PR_OK;
}

extern "C" int pr_shared_open(const unsigned char* handle64, void** dptr) {
    PR_CHECK_ARG(handle64 && dptr, "pr_shared_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    PR_CUDA_CALL(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dptr = p;
    return PR_OK;
}

extern "C" int pr_shared_close(void* dptr) {
    if (!dptr) return PR_OK;
    PR_CUDA_CALL(cudaIpcCloseMemHandle(dptr));
    return PR_OK;
}
