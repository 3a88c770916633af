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
