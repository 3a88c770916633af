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

namespace pr {

constexpr int PG_ROWS = 4;   // rows in flight per warp iteration

__global__ void __launch_bounds__(256) gather_rows_peers_kernel(const float4* const* __restrict__ shards, int G, long long N,
                                                                int D4, const long long* __restrict__ idx, long long R,
                                                                float4* __restrict__ out, int* __restrict__ status) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long ngroups = (R + PG_ROWS - 1) / PG_ROWS;
    for (long long g = warp; g < ngroups; g += nwarps) {
        const long long r0 = g * PG_ROWS;
        const float4* src[PG_ROWS];
        bool live[PG_ROWS];
#pragma unroll
        for (int j = 0; j < PG_ROWS; ++j) {
            live[j] = (r0 + j) < R;
            const long long id = live[j] ? __ldg(idx + r0 + j) : 0;
            const bool ok = (id >= 0) && (id < N);
            if (live[j] && !ok && status && lane == 0) atomicOr(status, 1);
            // owner(i) = i % G holds row i at local row i / G
            src[j] = ok ? (shards[(int)(id % G)] + (id / G) * (long long)D4) : nullptr;
        }
        for (int c = lane; c < D4; c += 32) {
            float4 v[PG_ROWS];
#pragma unroll
            for (int j = 0; j < PG_ROWS; ++j) v[j] = src[j] ? ldg_stream(src[j] + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < PG_ROWS; ++j)
                if (live[j]) out[(r0 + j) * (long long)D4 + c] = v[j];
        }
    }
}

// one warp per source row: claim a slot of this rank's region on the owner, copy the row there, record the local row id
__global__ void __launch_bounds__(256) push_rows_peers_kernel(const float4* __restrict__ rows, const long long* __restrict__ ids,
                                                              long long U, int D4, int G, int rank, long long cap,
                                                              long long skip_id, float4* const* __restrict__ recv_rows,
                                                              long long* const* __restrict__ recv_ids,
                                                              int* __restrict__ counters, int* __restrict__ status) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long u = warp; u < U; u += nwarps) {
        const long long id = __ldg(ids + u);
        if (id == skip_id) continue;                       // padding id: its gradient is dropped (nn.Embedding padding_idx)
        if (id < 0) {
            if (status && lane == 0) atomicOr(status, 1);
            continue;
        }
        const int owner = (int)(id % G);
        int pos = 0;
        if (lane == 0) pos = atomicAdd(counters + owner, 1);
        pos = __shfl_sync(0xffffffffu, pos, 0);
        if (pos >= cap) {                                  // receive region full: flagged, never written out of bounds
            if (status && lane == 0) atomicOr(status, 2);
            continue;
        }
        const long long slot = (long long)rank * cap + pos;
        float4* dst = recv_rows[owner] + slot * (long long)D4;
        const float4* src = rows + u * (long long)D4;
        for (int c = lane; c < D4; c += 128) {
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = (c + 32 * j < D4) ? ldg_stream(src + c + 32 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c + 32 * j < D4) dst[c + 32 * j] = v[j];
        }
        if (lane == 0) recv_ids[owner][slot] = id / G;
    }
}

}  // namespace pr

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
