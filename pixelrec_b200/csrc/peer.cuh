// pixelrec_b200 -- peer-memory row kernels (device code of peer.cu).  Also compiled for the HOST by tests/emu (PR_EMU:
// threads + barriers stand in for a CTA), so keep it free of inline PTX.
#pragma once

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
            for (int j = 0; j < PG_ROWS; ++j) v[j] = src[j] ? PR_LDG4_STREAM(src[j] + c) : make_float4(0.f, 0.f, 0.f, 0.f);
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
            for (int j = 0; j < 4; ++j) v[j] = (c + 32 * j < D4) ? PR_LDG4_STREAM(src + c + 32 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c + 32 * j < D4) dst[c + 32 * j] = v[j];
        }
        if (lane == 0) recv_ids[owner][slot] = id / G;
    }
}

}  // namespace pr
