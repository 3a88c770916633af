// pixelrec_b200 -- K2 index plan (device code of pr_scatter_plan): stable LSD radix sort of (item id, position) with digits of up
// to 10 bits, run boundaries and compaction.  N = 97 001 ids take TWO passes (9-bit digits); the id conversion rides in the first
// histogram pass, the run flags in the scan's reduce pass and the emission in its apply pass: 3 * passes + 3 launches (9 at C2; the
// first version's 8-bit digits and separate convert / flag / emit kernels took 15).  Output = what a stable sort defines, so every
// version is interchangeable bit for bit.  Also compiled for the HOST by tests/emu (emu_rows.cpp).
#pragma once

namespace pr {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 2048 keys per CTA
constexpr int RS_MAX_BINS = 1024;
constexpr int SCAN_THREADS = 1024;

// key of a request: its id, or the sentinel N (sorts last, dropped) for padding / out-of-range ids
__device__ __forceinline__ uint32_t plan_key(long long id, long long N, long long pad, int* status) {
    const bool oor = (id < 0) || (id >= N);
    if (oor && status) atomicOr(status, 1);
    return (oor || id == pad) ? (uint32_t)N : (uint32_t)id;
}

// per-tile digit histogram -> tile_hist[digit][tile].  idx != null (first pass): keys are made from the ids and stored.
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const long long* __restrict__ idx, long long N, long long pad,
                                                             int* __restrict__ status, uint32_t* __restrict__ keys, int R,
                                                             int shift, int bins, uint32_t* __restrict__ tile_hist, int T) {
    __shared__ uint32_t hist[RS_MAX_BINS];
    for (int d = threadIdx.x; d < bins; d += RS_THREADS) hist[d] = 0;
    __syncthreads();
    const int base = blockIdx.x * RS_TILE;
    const uint32_t dmask = (uint32_t)bins - 1u;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const int i = base + j * RS_THREADS + threadIdx.x;
        if (i < R) {
            uint32_t k;
            if (idx) {
                k = plan_key(idx[i], N, pad, status);
                keys[i] = k;
            } else {
                k = keys[i];
            }
            atomicAdd(&hist[(k >> shift) & dmask], 1u);
        }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < bins; d += RS_THREADS) tile_hist[(size_t)d * T + blockIdx.x] = hist[d];  // digit-major
}

// single-CTA in-place exclusive scan of a[0..n); optionally writes the grand total.  Warp w owns the contiguous segment
// [w * per_warp, +per_warp) and walks it 32 elements at a time, so every load / store is one coalesced 128-byte line (the first
// version gave each THREAD a contiguous chunk: 32 scattered lines per instruction, 13 us for the 21 K counters of a radix pass).
__global__ void __launch_bounds__(SCAN_THREADS) scan_single_kernel(uint32_t* __restrict__ a, int n,
                                                                   uint32_t* __restrict__ total_out) {
    __shared__ uint32_t warp_tot[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int per_warp = ((n + 31) / 32 + 31) / 32 * 32;          // multiple of 32
    const int lo = min(n, wid * per_warp), hi = min(n, lo + per_warp);
    uint32_t s = 0;
#pragma unroll 4
    for (int i = lo + lane; i < hi; i += 32) s += a[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) warp_tot[wid] = s;
    __syncthreads();
    if (wid == 0) {
        const uint32_t w = warp_tot[lane];
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        warp_tot[lane] = winc - w;  // exclusive offsets of the warps
        if (lane == 31 && total_out) *total_out = winc;
    }
    __syncthreads();
    uint32_t carry = warp_tot[wid];
    for (int i0 = lo; i0 < hi; i0 += 32) {
        const int i = i0 + lane;
        const uint32_t v = (i < hi) ? a[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (i < hi) a[i] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
}

// stable scatter of one pass.  vals_in == null (first pass): the value of key i is its position i.
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const uint32_t* __restrict__ keys_in,
                                                                const int* __restrict__ vals_in,
                                                                uint32_t* __restrict__ keys_out,
                                                                int* __restrict__ vals_out, int R, int shift, int bins,
                                                                const uint32_t* __restrict__ tile_base, int T) {
    __shared__ uint32_t whist[RS_THREADS / 32][RS_MAX_BINS];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int ww = 0; ww < RS_THREADS / 32; ++ww)
        for (int d = tid; d < bins; d += RS_THREADS) whist[ww][d] = 0;
    __syncthreads();
    const uint32_t dmask = (uint32_t)bins - 1u;
    // warp w owns the contiguous sub-tile [base, base + 256): order inside = (round j, lane) -> stable
    const int base = blockIdx.x * RS_TILE + w * (32 * RS_ITEMS);
    uint32_t key[RS_ITEMS];
    uint32_t local[RS_ITEMS];
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const int i = base + j * 32 + lane;
        const bool valid = i < R;
        key[j] = valid ? keys_in[i] : 0xffffffffu;
        const uint32_t d = (key[j] >> shift) & dmask;
        const unsigned peers = __match_any_sync(0xffffffffu, valid ? d : 0xffffffffu);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        const int leader = __ffs(peers) - 1;
        uint32_t b = 0;
        if (valid) b = whist[w][d];
        __syncwarp();
        if (valid && lane == leader) whist[w][d] = b + (uint32_t)__popc(peers);
        __syncwarp();
        local[j] = b + (uint32_t)rank;
    }
    __syncthreads();
    for (int d = tid; d < bins; d += RS_THREADS) {   // exclusive prefix over the 8 warps, per digit
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < RS_THREADS / 32; ++ww) {
            const uint32_t t = whist[ww][d];
            whist[ww][d] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const int i = base + j * 32 + lane;
        if (i < R) {
            const uint32_t d = (key[j] >> shift) & dmask;
            const uint32_t pos = tile_base[(size_t)d * T + blockIdx.x] + whist[w][d] + local[j];
            keys_out[pos] = key[j];
            vals_out[pos] = vals_in ? vals_in[i] : i;
        }
    }
}

// 1 at the first position of every run of a real key (< N)
__device__ __forceinline__ uint32_t seg_flag(const uint32_t* __restrict__ skeys, int i, uint32_t N) {
    const uint32_t k = skeys[i];
    return (k < N && (i == 0 || skeys[i - 1] != k)) ? 1u : 0u;
}

// run starts per tile (phase 1 of the multi-CTA exclusive scan of the flags; the flags themselves are not stored)
__global__ void __launch_bounds__(RS_THREADS) seg_reduce_kernel(const uint32_t* __restrict__ skeys, int R, uint32_t N,
                                                                uint32_t* __restrict__ tile_sum) {
    __shared__ uint32_t ws[RS_THREADS / 32];
    const int base = blockIdx.x * RS_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const int i = base + j * RS_THREADS + threadIdx.x;
        if (i < R) s += seg_flag(skeys, i, N);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < RS_THREADS / 32; ++w) t += ws[w];
        tile_sum[blockIdx.x] = t;
    }
}

// phase 3 + emission: run index of every position (tile offset + prefix inside the tile; a thread owns 8 consecutive positions),
// then uniq_ids / seg_start / row2slot at run starts, the closing boundary and n_uniq
__global__ void __launch_bounds__(RS_THREADS) seg_emit_kernel(const uint32_t* __restrict__ skeys, int R, uint32_t N,
                                                              const uint32_t* __restrict__ tile_off, int* __restrict__ uniq_ids,
                                                              int* __restrict__ seg_start, int* __restrict__ n_uniq,
                                                              int* __restrict__ row2slot) {
    __shared__ uint32_t ws[RS_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int base = blockIdx.x * RS_TILE + tid * RS_ITEMS;
    uint32_t f[RS_ITEMS], key[RS_ITEMS];
    uint32_t prev = (base > 0 && base - 1 < R) ? skeys[base - 1] : 0xffffffffu;
    const uint32_t before = prev;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const int i = base + j;
        key[j] = (i < R) ? skeys[i] : 0xffffffffu;
        f[j] = (i < R && key[j] < N && (i == 0 || prev != key[j])) ? 1u : 0u;
        prev = key[j];
        s += f[j];
    }
    uint32_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) ws[wid] = inc;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < wid; ++w) woff += ws[w];
    uint32_t u = tile_off[blockIdx.x] + woff + (inc - s);     // runs that start before position `base`
    prev = before;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const int i = base + j;
        if (i < R) {
            const uint32_t k = key[j];
            if (f[j]) {
                uniq_ids[u] = (int)k;
                seg_start[u] = i;
                if (row2slot) row2slot[k] = (int)u;
            }
            if (k >= N && (i == 0 || prev < N)) {  // first dropped row closes the last real run
                seg_start[u] = i;
                *n_uniq = (int)u;
            }
            if (i == R - 1 && k < N) {
                seg_start[u + f[j]] = R;
                *n_uniq = (int)(u + f[j]);
            }
            prev = k;
            u += f[j];
        }
    }
}

__global__ void plan_empty_kernel(int* seg_start, int* n_uniq) {
    seg_start[0] = 0;
    *n_uniq = 0;
}

}  // namespace pr
