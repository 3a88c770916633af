// pixelrec_b200 -- fused (dropout) + residual + LayerNorm (+ dropout) and the feed-forward activation.
//   replaces REC/model/IDNet/sasrec.py:77-83 and REC/model/layers.py:613-615, 651-660, 667-671.
// HBM-bound: one warp owns one row, the row lives in registers (float4 per lane, coalesced 512-B
// warp transactions), statistics by warp shuffles, dropout masks regenerated from Philox in the
// backward instead of being stored.
#include <algorithm>
#include <math.h>

#include "common.cuh"
#include "act.cuh"

namespace pr {

struct LnArgs {
    const float* h; long long h_seq_stride; long long rows_per_seq;
    const float* res; long long res_period;
    const float* gamma; const float* beta; float eps;
    long long rows; int D4;
    float p_pre, p_post; unsigned long long seed; unsigned stream_pre, stream_post;
    int l2_prefetch;   // PR_TUNE_LN_L2_PREFETCH: pull the warp's next row into L2 while this one is processed
    int h_is_z;        // backward only: `h` already holds z = drop_pre(h) + res (written by pr_gemm_tf32_drop's epilogue), res is null;
                       // the pre-dropout mask is then applied to the outgoing gradient only
#ifdef PR_SEED_DEV
    const unsigned long long* seed_dev;   // device-side seed offset (pr_set_seed_device)
#endif
};

// keep bits of one row for this lane: chunk j (= float4 column lane + 32*j) uses bits [4*(j&1), +4) of
// mk[j>>1], drawn from Philox(counter = row*D4 + lane + 64*(j>>1), stream)   (oracle/philox_np.py restates it)
template <int VPL>
__device__ __forceinline__ void row_keep_bits(const Philox& ph, unsigned stream, unsigned thr, long long row, int D4,
                                              int lane, unsigned (&mk)[(VPL + 1) / 2]) {
#pragma unroll
    for (int q = 0; q < (VPL + 1) / 2; ++q)
        mk[q] = keep_bits8(ph((unsigned long long)row * D4 + lane + 64 * q, stream), thr);
}

__device__ __forceinline__ float4 apply_keep(float4 v, unsigned bits4, float inv_keep) {
    v.x = (bits4 & 1u) ? v.x * inv_keep : 0.f;
    v.y = (bits4 & 2u) ? v.y * inv_keep : 0.f;
    v.z = (bits4 & 4u) ? v.z * inv_keep : 0.f;
    v.w = (bits4 & 8u) ? v.w * inv_keep : 0.f;
    return v;
}

template <int VPL>
__device__ __forceinline__ void load_z(const LnArgs& a, long long row, int lane, const unsigned (&mk_pre)[(VPL + 1) / 2],
                                       float inv_keep_pre, float4 (&z)[VPL]) {
    const long long s = row / a.rows_per_seq, t = row - s * a.rows_per_seq;
    const float4* h4 = reinterpret_cast<const float4*>(a.h + s * a.h_seq_stride) + t * a.D4;
    const long long rrow = a.res_period > 0 ? (row % a.res_period) : row;
    const float4* r4 = a.res ? reinterpret_cast<const float4*>(a.res) + rrow * a.D4 : nullptr;
    float4 hv[VPL], rv[VPL];
#pragma unroll
    for (int j = 0; j < VPL; ++j) {     // issue every load of the row before touching the data
        const int c = lane + 32 * j;
        hv[j] = (c < a.D4) ? h4[c] : make_float4(0.f, 0.f, 0.f, 0.f);
        rv[j] = (r4 && c < a.D4) ? r4[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
        float4 v = hv[j];
        if (a.p_pre > 0.f && !a.h_is_z) v = apply_keep(v, mk_pre[j >> 1] >> (4 * (j & 1)), inv_keep_pre);
        v.x += rv[j].x; v.y += rv[j].y; v.z += rv[j].z; v.w += rv[j].w;
        z[j] = v;
    }
}

// lanes 0..2 each ask the copy engine to pull one array's next row into L2 (no smem, nothing to wait for)
__device__ __forceinline__ void prefetch_row_l2(const LnArgs& a, long long row, int lane, const float* dy) {
    const unsigned bytes = (unsigned)a.D4 * 16u;
    if (lane == 0) {
        const long long s = row / a.rows_per_seq, t = row - s * a.rows_per_seq;
        l2_prefetch_bulk(a.h + s * a.h_seq_stride + t * (long long)a.D4 * 4, bytes);
    } else if (lane == 1) {
        if (a.res && a.res_period <= 0) l2_prefetch_bulk(a.res + row * (long long)a.D4 * 4, bytes);
    } else if (lane == 2) {
        if (dy) l2_prefetch_bulk(dy + row * (long long)a.D4 * 4, bytes);
    }
}

template <int VPL>
__device__ __forceinline__ void row_stats(const float4 (&z)[VPL], int lane, int D4, float& mean, float& var) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) s += (z[j].x + z[j].y) + (z[j].z + z[j].w);  // out-of-range chunks hold zeros
    const float invD = 1.0f / (float)(D4 * 4);
    mean = warp_sum(s) * invD;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
        if (lane + 32 * j < D4) {
            const float a = z[j].x - mean, b = z[j].y - mean, c = z[j].z - mean, d = z[j].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    }
    var = warp_sum(q) * invD;
}

template <int VPL>
__global__ void __launch_bounds__(256, (VPL <= 4) ? 4 : 1) add_ln_fwd_kernel(LnArgs a, float* __restrict__ y, float* __restrict__ mean_out,
                                                         float* __restrict__ rstd_out) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const Philox ph(PR_SEED(a));
    const unsigned thr_pre = drop_threshold(a.p_pre), thr_post = drop_threshold(a.p_post);
    const float ik_pre = 1.0f / (1.0f - a.p_pre), ik_post = 1.0f / (1.0f - a.p_post);
    const float4* g4 = reinterpret_cast<const float4*>(a.gamma);
    const float4* b4 = reinterpret_cast<const float4*>(a.beta);
    for (long long row = warp; row < a.rows; row += nwarps) {
        float4 z[VPL];
        unsigned mk_pre[(VPL + 1) / 2], mk_post[(VPL + 1) / 2];
        if (a.l2_prefetch && row + nwarps < a.rows) prefetch_row_l2(a, row + nwarps, lane, nullptr);
        if (a.p_pre > 0.f) row_keep_bits<VPL>(ph, a.stream_pre, thr_pre, row, a.D4, lane, mk_pre);
        if (a.p_post > 0.f) row_keep_bits<VPL>(ph, a.stream_post, thr_post, row, a.D4, lane, mk_post);
        load_z<VPL>(a, row, lane, mk_pre, ik_pre, z);
        float mean, var;
        row_stats<VPL>(z, lane, a.D4, mean, var);
        const float rstd = 1.0f / sqrtf(var + a.eps);
        float4* y4 = reinterpret_cast<float4*>(y) + row * a.D4;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const int c = lane + 32 * j;
            if (c < a.D4) {
                const float4 g = __ldg(g4 + c), b = __ldg(b4 + c);
                float4 o;
                o.x = (z[j].x - mean) * rstd * g.x + b.x;
                o.y = (z[j].y - mean) * rstd * g.y + b.y;
                o.z = (z[j].z - mean) * rstd * g.z + b.z;
                o.w = (z[j].w - mean) * rstd * g.w + b.w;
                if (a.p_post > 0.f) o = apply_keep(o, mk_post[j >> 1] >> (4 * (j & 1)), ik_post);
                y4[c] = o;
            }
        }
        if (lane == 0) {
            mean_out[row] = mean;
            rstd_out[row] = rstd;
        }
    }
}

// Forward, two rows per warp iteration (PR_TUNE_LN_FWD_ROWS2): the loads of both rows are issued before either is reduced, which
// doubles the bytes a warp keeps in flight (the one-row kernel alternates load / shuffle-reduce / store and leaves the SM short of
// outstanding requests once the row is a single tensor, i.e. behind pr_gemm_tf32_drop).  Per-row arithmetic is add_ln_fwd_kernel's.
template <int VPL>
__global__ void __launch_bounds__(256, (VPL <= 4) ? 3 : 1) add_ln_fwd_rows2_kernel(LnArgs a, float* __restrict__ y,
                                                                                   float* __restrict__ mean_out,
                                                                                   float* __restrict__ rstd_out) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const Philox ph(PR_SEED(a));
    const unsigned thr_pre = drop_threshold(a.p_pre), thr_post = drop_threshold(a.p_post);
    const float ik_pre = 1.0f / (1.0f - a.p_pre), ik_post = 1.0f / (1.0f - a.p_post);
    const float4* g4 = reinterpret_cast<const float4*>(a.gamma);
    const float4* b4 = reinterpret_cast<const float4*>(a.beta);
    for (long long row0 = 2 * warp; row0 < a.rows; row0 += 2 * nwarps) {
        const bool two = row0 + 1 < a.rows;
        float4 z[2][VPL];
        unsigned mk_pre[2][(VPL + 1) / 2], mk_post[2][(VPL + 1) / 2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (r == 0 || two) {
                if (a.p_pre > 0.f) row_keep_bits<VPL>(ph, a.stream_pre, thr_pre, row0 + r, a.D4, lane, mk_pre[r]);
                if (a.p_post > 0.f) row_keep_bits<VPL>(ph, a.stream_post, thr_post, row0 + r, a.D4, lane, mk_post[r]);
            }
        }
        load_z<VPL>(a, row0, lane, mk_pre[0], ik_pre, z[0]);
        if (two) load_z<VPL>(a, row0 + 1, lane, mk_pre[1], ik_pre, z[1]);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (r == 1 && !two) break;
            const long long row = row0 + r;
            float mean, var;
            row_stats<VPL>(z[r], lane, a.D4, mean, var);
            const float rstd = 1.0f / sqrtf(var + a.eps);
            float4* y4 = reinterpret_cast<float4*>(y) + row * a.D4;
#pragma unroll
            for (int j = 0; j < VPL; ++j) {
                const int c = lane + 32 * j;
                if (c < a.D4) {
                    const float4 g = __ldg(g4 + c), b = __ldg(b4 + c);
                    float4 o;
                    o.x = (z[r][j].x - mean) * rstd * g.x + b.x;
                    o.y = (z[r][j].y - mean) * rstd * g.y + b.y;
                    o.z = (z[r][j].z - mean) * rstd * g.z + b.z;
                    o.w = (z[r][j].w - mean) * rstd * g.w + b.w;
                    if (a.p_post > 0.f) o = apply_keep(o, mk_post[r][j >> 1] >> (4 * (j & 1)), ik_post);
                    y4[c] = o;
                }
            }
            if (lane == 0) {
                mean_out[row] = mean;
                rstd_out[row] = rstd;
            }
        }
    }
}

// Forward as per-warp TMA row pipelines (PR_TUNE_LN_FWD_PIPE; D == 128*VPL): stage = [h | res] rows, same arithmetic
template <int VPL, int STAGES>
__global__ void __launch_bounds__(256, (VPL <= 4) ? 2 : 1) add_ln_fwd_pipe_kernel(LnArgs a, float* __restrict__ y,
                                                                                float* __restrict__ mean_out,
                                                                                float* __restrict__ rstd_out) {
    extern __shared__ __align__(128) float4 sm_dyn[];
    constexpr int RF4 = VPL * 32;
    constexpr unsigned ROW_BYTES = RF4 * 16u;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float4* ring = sm_dyn + (size_t)wid * STAGES * 2 * RF4;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm_dyn + (size_t)nw * STAGES * 2 * RF4) + wid * STAGES;
    const long long warp = (long long)blockIdx.x * nw + wid;
    const long long nwarps = (long long)gridDim.x * nw;
    const bool has_res = a.res != nullptr;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncwarp();
    auto issue = [&](int s, long long row) {            // lane 0 only
        const long long sq = row / a.rows_per_seq, tt = row - sq * a.rows_per_seq;
        float4* st = ring + (size_t)s * 2 * RF4;
        mbar_arrive_expect_tx(&bars[s], (has_res ? 2u : 1u) * ROW_BYTES);
        bulk_g2s(st, a.h + sq * a.h_seq_stride + tt * (long long)(RF4 * 4), ROW_BYTES, &bars[s]);
        if (has_res) {
            const long long rrow = a.res_period > 0 ? (row % a.res_period) : row;
            bulk_g2s(st + RF4, a.res + rrow * (long long)(RF4 * 4), ROW_BYTES, &bars[s]);
        }
    };
    const long long mine = warp < a.rows ? (a.rows - warp + nwarps - 1) / nwarps : 0;
    if (lane == 0)
        for (int s = 0; s < STAGES && s < mine; ++s) issue(s, warp + s * nwarps);
    const Philox ph(PR_SEED(a));
    const unsigned thr_pre = drop_threshold(a.p_pre), thr_post = drop_threshold(a.p_post);
    const float ik_pre = 1.0f / (1.0f - a.p_pre), ik_post = 1.0f / (1.0f - a.p_post);
    const float4* g4 = reinterpret_cast<const float4*>(a.gamma);
    const float4* b4 = reinterpret_cast<const float4*>(a.beta);
    int s = 0;
    unsigned parity = 0;
    for (long long k = 0; k < mine; ++k) {
        const long long row = warp + k * nwarps;
        unsigned mk_pre[(VPL + 1) / 2], mk_post[(VPL + 1) / 2];
        if (a.p_pre > 0.f) row_keep_bits<VPL>(ph, a.stream_pre, thr_pre, row, RF4, lane, mk_pre);
        if (a.p_post > 0.f) row_keep_bits<VPL>(ph, a.stream_post, thr_post, row, RF4, lane, mk_post);
        mbar_wait(&bars[s], parity);
        const float4* st = ring + (size_t)s * 2 * RF4;
        float4 z[VPL];
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const int c = lane + 32 * j;
            float4 v = st[c];
            if (a.p_pre > 0.f) v = apply_keep(v, mk_pre[j >> 1] >> (4 * (j & 1)), ik_pre);
            if (has_res) {
                const float4 r = st[RF4 + c];
                v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
            }
            z[j] = v;
        }
        __syncwarp();
        if (lane == 0 && k + STAGES < mine) {
            fence_proxy_async();
            issue(s, row + (long long)STAGES * nwarps);
        }
        if (++s == STAGES) { s = 0; parity ^= 1u; }
        float mean, var;
        row_stats<VPL>(z, lane, RF4, mean, var);
        const float rstd = 1.0f / sqrtf(var + a.eps);
        float4* y4 = reinterpret_cast<float4*>(y) + row * RF4;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const int c = lane + 32 * j;
            const float4 g = __ldg(g4 + c), b = __ldg(b4 + c);
            float4 o;
            o.x = (z[j].x - mean) * rstd * g.x + b.x;
            o.y = (z[j].y - mean) * rstd * g.y + b.y;
            o.z = (z[j].z - mean) * rstd * g.z + b.z;
            o.w = (z[j].w - mean) * rstd * g.w + b.w;
            if (a.p_post > 0.f) o = apply_keep(o, mk_post[j >> 1] >> (4 * (j & 1)), ik_post);
            y4[c] = o;
        }
        if (lane == 0) {
            mean_out[row] = mean;
            rstd_out[row] = rstd;
        }
    }
}

// fixed-order sum of the warps' register accumulators through shared memory -> partials[{0,1,2}][blockIdx.x][D]
template <int VPL, bool DBIAS>
__device__ __forceinline__ void cta_reduce_partials(const float4 (&accg)[VPL], const float4 (&accb)[VPL],
                                                    const float4 (&accd)[DBIAS ? VPL : 1], float4* sm_acc, int D4,
                                                    float* __restrict__ partials) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int w = 0; w < nw; ++w) {
        if (wid == w) {
#pragma unroll
            for (int j = 0; j < VPL; ++j) {
                const int c = lane + 32 * j;
                if (c < D4) {
                    if (w == 0) {
                        sm_acc[c] = accg[j];
                        sm_acc[D4 + c] = accb[j];
                        if (DBIAS) sm_acc[2 * D4 + c] = accd[DBIAS ? j : 0];
                    } else {
                        if (DBIAS) {
                            float4 u = sm_acc[2 * D4 + c];
                            u.x += accd[DBIAS ? j : 0].x; u.y += accd[DBIAS ? j : 0].y; u.z += accd[DBIAS ? j : 0].z; u.w += accd[DBIAS ? j : 0].w;
                            sm_acc[2 * D4 + c] = u;
                        }
                        float4 t = sm_acc[c];
                        t.x += accg[j].x; t.y += accg[j].y; t.z += accg[j].z; t.w += accg[j].w;
                        sm_acc[c] = t;
                        t = sm_acc[D4 + c];
                        t.x += accb[j].x; t.y += accb[j].y; t.z += accb[j].z; t.w += accb[j].w;
                        sm_acc[D4 + c] = t;
                    }
                }
            }
        }
        __syncthreads();
    }
    float4* pg = reinterpret_cast<float4*>(partials) + (long long)blockIdx.x * D4;
    float4* pb = reinterpret_cast<float4*>(partials) + ((long long)gridDim.x + blockIdx.x) * D4;
    float4* pd = reinterpret_cast<float4*>(partials) + (2 * (long long)gridDim.x + blockIdx.x) * D4;
    for (int c = threadIdx.x; c < D4; c += blockDim.x) {
        pg[c] = sm_acc[c];
        pb[c] = sm_acc[D4 + c];
        if (DBIAS) pd[c] = sm_acc[2 * D4 + c];
    }
}

// backward.  dgamma/dbeta: per-lane register accumulators over the rows of this warp, then a
// fixed-order sum over the CTA's warps in shared memory -> partials[{0,1}][blockIdx.x][D]
template <int VPL, bool DBIAS>
__global__ void __launch_bounds__(256, (VPL <= 4) ? 2 : 1) add_ln_bwd_kernel(LnArgs a, const float* __restrict__ dy,
                                                         const float* __restrict__ mean_in,
                                                         const float* __restrict__ rstd_in, float* __restrict__ dh,
                                                         long long dh_seq_stride, int dh_accumulate,
                                                         float* __restrict__ dres, float* __restrict__ partials) {
    extern __shared__ float4 sm_acc[];  // [2 or 3][D4]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long warp = (long long)blockIdx.x * nw + wid;
    const long long nwarps = (long long)gridDim.x * nw;
    const Philox ph(PR_SEED(a));
    const unsigned thr_pre = drop_threshold(a.p_pre), thr_post = drop_threshold(a.p_post);
    const float ik_pre = 1.0f / (1.0f - a.p_pre), ik_post = 1.0f / (1.0f - a.p_post);
    const float4* g4 = reinterpret_cast<const float4*>(a.gamma);
    const float invD = 1.0f / (float)(a.D4 * 4);
    float4 accg[VPL], accb[VPL], accd[DBIAS ? VPL : 1];      // accd: column sums of dh = bias grad of the producing Linear
#pragma unroll
    for (int j = 0; j < VPL; ++j) accg[j] = accb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < (DBIAS ? VPL : 1); ++j) accd[j] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (long long row = warp; row < a.rows; row += nwarps) {
        float4 z[VPL];
        unsigned mk_pre[(VPL + 1) / 2], mk_post[(VPL + 1) / 2];
        if (a.l2_prefetch && row + nwarps < a.rows) prefetch_row_l2(a, row + nwarps, lane, dy);
        if (a.p_pre > 0.f) row_keep_bits<VPL>(ph, a.stream_pre, thr_pre, row, a.D4, lane, mk_pre);
        if (a.p_post > 0.f) row_keep_bits<VPL>(ph, a.stream_post, thr_post, row, a.D4, lane, mk_post);
        load_z<VPL>(a, row, lane, mk_pre, ik_pre, z);
        const float mean = mean_in[row], rstd = rstd_in[row];
        const float4* dy4 = reinterpret_cast<const float4*>(dy) + row * a.D4;
        float4 dxh[VPL];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const int c = lane + 32 * j;
            dxh[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < a.D4) {
                float4 d = ldg_stream(dy4 + c);
                if (a.p_post > 0.f) d = apply_keep(d, mk_post[j >> 1] >> (4 * (j & 1)), ik_post);
                const float4 g = __ldg(g4 + c);
                float4 xh;  // z[j] becomes xhat
                xh.x = (z[j].x - mean) * rstd; xh.y = (z[j].y - mean) * rstd;
                xh.z = (z[j].z - mean) * rstd; xh.w = (z[j].w - mean) * rstd;
                z[j] = xh;
                accg[j].x += d.x * xh.x; accg[j].y += d.y * xh.y; accg[j].z += d.z * xh.z; accg[j].w += d.w * xh.w;
                accb[j].x += d.x; accb[j].y += d.y; accb[j].z += d.z; accb[j].w += d.w;
                d.x *= g.x; d.y *= g.y; d.z *= g.z; d.w *= g.w;
                dxh[j] = d;
                s1 += (d.x + d.y) + (d.z + d.w);
                s2 += (d.x * xh.x + d.y * xh.y) + (d.z * xh.z + d.w * xh.w);
            }
        }
        s1 = warp_sum(s1) * invD;
        s2 = warp_sum(s2) * invD;
        float4* dh4;
        if (dh_seq_stride > 0) {
            const long long sq = row / a.rows_per_seq, tt = row - sq * a.rows_per_seq;
            dh4 = reinterpret_cast<float4*>(dh + sq * dh_seq_stride) + tt * a.D4;
        } else {
            dh4 = reinterpret_cast<float4*>(dh) + row * a.D4;
        }
        float4* dr4 = dres ? reinterpret_cast<float4*>(dres) + row * a.D4 : nullptr;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const int c = lane + 32 * j;
            if (c < a.D4) {
                float4 dz;
                dz.x = rstd * (dxh[j].x - s1 - z[j].x * s2);
                dz.y = rstd * (dxh[j].y - s1 - z[j].y * s2);
                dz.z = rstd * (dxh[j].z - s1 - z[j].z * s2);
                dz.w = rstd * (dxh[j].w - s1 - z[j].w * s2);
                if (dr4) dr4[c] = dz;
                if (a.p_pre > 0.f) dz = apply_keep(dz, mk_pre[j >> 1] >> (4 * (j & 1)), ik_pre);
                if (DBIAS) { accd[DBIAS ? j : 0].x += dz.x; accd[DBIAS ? j : 0].y += dz.y; accd[DBIAS ? j : 0].z += dz.z; accd[DBIAS ? j : 0].w += dz.w; }
                if (dh_accumulate) {
                    const float4 o = dh4[c];
                    dz.x += o.x; dz.y += o.y; dz.z += o.z; dz.w += o.w;
                }
                dh4[c] = dz;
            }
        }
    }
    cta_reduce_partials<VPL, DBIAS>(accg, accb, accd, sm_acc, a.D4, partials);
}

// Same backward as per-warp TMA row pipelines (PR_TUNE_LN_BWD_PIPE; needs D == 128*VPL so every lane column is live):
// each warp owns STAGES shared-memory stages of [dy | h | res] rows, lane 0 keeps STAGES rows of bulk copies in
// flight (cp.async.bulk + mbarrier complete_tx) while the warp works on the oldest one, so the bytes in flight per SM
// no longer depend on how many rows fit in registers.  Row -> warp mapping, arithmetic and summation order are those
// of add_ln_bwd_kernel.
template <int VPL, bool DBIAS, int STAGES, int NSRC>   // NSRC rows per stage: [dy | h | res], or [dy | z] (no residual / z mode)
__global__ void __launch_bounds__(256, (VPL <= 4) ? 2 : 1) add_ln_bwd_pipe_kernel(
    LnArgs a, const float* __restrict__ dy, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
    float* __restrict__ dh, long long dh_seq_stride, int dh_accumulate, float* __restrict__ dres,
    float* __restrict__ partials) {
    extern __shared__ __align__(128) float4 sm_dyn[];
    constexpr int RF4 = VPL * 32;                       // float4 per row
    constexpr unsigned ROW_BYTES = RF4 * 16u;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float4* ring = sm_dyn + (size_t)wid * STAGES * NSRC * RF4;
    float4* sm_acc = sm_dyn + (size_t)nw * STAGES * NSRC * RF4;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm_acc + 3 * RF4) + wid * STAGES;
    const long long warp = (long long)blockIdx.x * nw + wid;
    const long long nwarps = (long long)gridDim.x * nw;
    const bool has_res = NSRC == 3 && a.res != nullptr;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncwarp();
    auto issue = [&](int s, long long row) {            // lane 0 only
        const long long sq = row / a.rows_per_seq, tt = row - sq * a.rows_per_seq;
        float4* st = ring + (size_t)s * NSRC * RF4;
        mbar_arrive_expect_tx(&bars[s], (has_res ? 3u : 2u) * ROW_BYTES);
        bulk_g2s(st, dy + row * (long long)(RF4 * 4), ROW_BYTES, &bars[s]);
        bulk_g2s(st + RF4, a.h + sq * a.h_seq_stride + tt * (long long)(RF4 * 4), ROW_BYTES, &bars[s]);
        if (has_res) {
            const long long rrow = a.res_period > 0 ? (row % a.res_period) : row;
            bulk_g2s(st + 2 * RF4, a.res + rrow * (long long)(RF4 * 4), ROW_BYTES, &bars[s]);
        }
    };
    const long long mine = warp < a.rows ? (a.rows - warp + nwarps - 1) / nwarps : 0;
    if (lane == 0)
        for (int s = 0; s < STAGES && s < mine; ++s) issue(s, warp + s * nwarps);

    const Philox ph(PR_SEED(a));
    const unsigned thr_pre = drop_threshold(a.p_pre), thr_post = drop_threshold(a.p_post);
    const float ik_pre = 1.0f / (1.0f - a.p_pre), ik_post = 1.0f / (1.0f - a.p_post);
    const float4* g4 = reinterpret_cast<const float4*>(a.gamma);
    const float invD = 1.0f / (float)(RF4 * 4);
    float4 accg[VPL], accb[VPL], accd[DBIAS ? VPL : 1];
#pragma unroll
    for (int j = 0; j < VPL; ++j) accg[j] = accb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < (DBIAS ? VPL : 1); ++j) accd[j] = make_float4(0.f, 0.f, 0.f, 0.f);

    int s = 0;
    unsigned parity = 0;
    for (long long k = 0; k < mine; ++k) {
        const long long row = warp + k * nwarps;
        unsigned mk_pre[(VPL + 1) / 2], mk_post[(VPL + 1) / 2];
        if (a.p_pre > 0.f) row_keep_bits<VPL>(ph, a.stream_pre, thr_pre, row, RF4, lane, mk_pre);
        if (a.p_post > 0.f) row_keep_bits<VPL>(ph, a.stream_post, thr_post, row, RF4, lane, mk_post);
        const float mean = mean_in[row], rstd = rstd_in[row];
        mbar_wait(&bars[s], parity);
        const float4* st = ring + (size_t)s * NSRC * RF4;
        float4 d[VPL], z[VPL];
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const int c = lane + 32 * j;
            d[j] = st[c];
            float4 v = st[RF4 + c];
            if (a.p_pre > 0.f && !a.h_is_z) v = apply_keep(v, mk_pre[j >> 1] >> (4 * (j & 1)), ik_pre);
            if (has_res) {
                const float4 r = st[2 * RF4 + c];
                v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
            }
            z[j] = v;
        }
        __syncwarp();                                   // every lane has its copy: the stage may be refilled
        if (lane == 0 && k + STAGES < mine) {
            fence_proxy_async();
            issue(s, row + (long long)STAGES * nwarps);
        }
        if (++s == STAGES) { s = 0; parity ^= 1u; }

        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const int c = lane + 32 * j;
            float4 dd = d[j];
            if (a.p_post > 0.f) dd = apply_keep(dd, mk_post[j >> 1] >> (4 * (j & 1)), ik_post);
            const float4 g = __ldg(g4 + c);
            float4 xh;
            xh.x = (z[j].x - mean) * rstd; xh.y = (z[j].y - mean) * rstd;
            xh.z = (z[j].z - mean) * rstd; xh.w = (z[j].w - mean) * rstd;
            z[j] = xh;
            accg[j].x += dd.x * xh.x; accg[j].y += dd.y * xh.y; accg[j].z += dd.z * xh.z; accg[j].w += dd.w * xh.w;
            accb[j].x += dd.x; accb[j].y += dd.y; accb[j].z += dd.z; accb[j].w += dd.w;
            dd.x *= g.x; dd.y *= g.y; dd.z *= g.z; dd.w *= g.w;
            d[j] = dd;
            s1 += (dd.x + dd.y) + (dd.z + dd.w);
            s2 += (dd.x * xh.x + dd.y * xh.y) + (dd.z * xh.z + dd.w * xh.w);
        }
        s1 = warp_sum(s1) * invD;
        s2 = warp_sum(s2) * invD;
        float4* dh4;
        if (dh_seq_stride > 0) {
            const long long sq = row / a.rows_per_seq, tt = row - sq * a.rows_per_seq;
            dh4 = reinterpret_cast<float4*>(dh + sq * dh_seq_stride) + tt * RF4;
        } else {
            dh4 = reinterpret_cast<float4*>(dh) + row * RF4;
        }
        float4* dr4 = dres ? reinterpret_cast<float4*>(dres) + row * RF4 : nullptr;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const int c = lane + 32 * j;
            float4 dz;
            dz.x = rstd * (d[j].x - s1 - z[j].x * s2);
            dz.y = rstd * (d[j].y - s1 - z[j].y * s2);
            dz.z = rstd * (d[j].z - s1 - z[j].z * s2);
            dz.w = rstd * (d[j].w - s1 - z[j].w * s2);
            if (dr4) dr4[c] = dz;
            if (a.p_pre > 0.f) dz = apply_keep(dz, mk_pre[j >> 1] >> (4 * (j & 1)), ik_pre);
            if (DBIAS) { accd[DBIAS ? j : 0].x += dz.x; accd[DBIAS ? j : 0].y += dz.y; accd[DBIAS ? j : 0].z += dz.z; accd[DBIAS ? j : 0].w += dz.w; }
            if (dh_accumulate) {
                const float4 o = dh4[c];
                dz.x += o.x; dz.y += o.y; dz.z += o.z; dz.w += o.w;
            }
            dh4[c] = dz;
        }
    }
    __syncthreads();
    cta_reduce_partials<VPL, DBIAS>(accg, accb, accd, sm_acc, RF4, partials);
}

// out[m][c] = sum_p partials[m][p][c].  32 columns x 32 row-groups per CTA; group y adds rows p = y, y+32, ...
// in order, then the 32 group sums are added in order -> deterministic.
__global__ void __launch_bounds__(1024) colsum_kernel(const float* __restrict__ partials, int n_partials, long long D,
                                                      float* __restrict__ out) {
    __shared__ float sm[32][33];
    const long long c = (long long)blockIdx.x * 32 + threadIdx.x;
    const float* P = partials + (long long)blockIdx.y * n_partials * D;
    float s = 0.f;
    if (c < D)
        for (int p = threadIdx.y; p < n_partials; p += 32) s += P[(long long)p * D + c];
    sm[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < D) {
        float t = 0.f;
#pragma unroll
        for (int y = 0; y < 32; ++y) t += sm[y][threadIdx.x];
        out[(long long)blockIdx.y * D + c] = t;
    }
}

// ACT is a template parameter so the switch folds away; each thread keeps UNR independent 16-B loads in flight
template <int ACT, int UNR>
__global__ void __launch_bounds__(256) act_fwd_kernel(const float* __restrict__ x, long long n, float* __restrict__ y) {
    const long long n4 = n >> 2, stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    long long i = i0;
    for (; i + (UNR - 1) * stride < n4; i += UNR * stride) {
        float4 v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) v[u] = ldg_stream(x4 + i + u * stride);
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            v[u].x = act_f(v[u].x, ACT); v[u].y = act_f(v[u].y, ACT); v[u].z = act_f(v[u].z, ACT); v[u].w = act_f(v[u].w, ACT);
            y4[i + u * stride] = v[u];
        }
    }
    for (; i < n4; i += stride) {
        float4 v = x4[i];
        v.x = act_f(v.x, ACT); v.y = act_f(v.y, ACT); v.z = act_f(v.z, ACT); v.w = act_f(v.w, ACT);
        y4[i] = v;
    }
    for (long long t = (n4 << 2) + i0; t < n; t += stride) y[t] = act_f(x[t], ACT);
}

template <int ACT, int UNR>
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                      long long n, float* __restrict__ dx) {
    const long long n4 = n >> 2, stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* d4 = reinterpret_cast<const float4*>(dy);
    float4* o4 = reinterpret_cast<float4*>(dx);
    long long i = i0;
    for (; i + (UNR - 1) * stride < n4; i += UNR * stride) {
        float4 v[UNR], d[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            v[u] = ldg_stream(x4 + i + u * stride);
            d[u] = ldg_stream(d4 + i + u * stride);
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            d[u].x *= act_df(v[u].x, ACT); d[u].y *= act_df(v[u].y, ACT); d[u].z *= act_df(v[u].z, ACT); d[u].w *= act_df(v[u].w, ACT);
            o4[i + u * stride] = d[u];
        }
    }
    for (; i < n4; i += stride) {
        const float4 v = x4[i];
        float4 d = d4[i];
        d.x *= act_df(v.x, ACT); d.y *= act_df(v.y, ACT); d.z *= act_df(v.z, ACT); d.w *= act_df(v.w, ACT);
        o4[i] = d;
    }
    for (long long t = (n4 << 2) + i0; t < n; t += stride) dx[t] = dy[t] * act_df(x[t], ACT);
}

// dx = act'(x) * dy and per-CTA partial column sums of dx (= bias grad of the Linear that produced x): warp per row,
// the row is walked in chunks of CH float4 per lane with every load of a chunk issued before its math
template <int ACT, int VPL>
__global__ void __launch_bounds__(256, (VPL <= 8) ? 2 : 1) act_bwd_bias_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           long long rows, int C4, float* __restrict__ dx,
                                                           float* __restrict__ partials) {
    extern __shared__ float4 sm_acc[];  // [C4]
    constexpr int CH = VPL < 8 ? VPL : 8;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long warp = (long long)blockIdx.x * nw + wid, nwarps = (long long)gridDim.x * nw;
    float4 acc[VPL];
#pragma unroll
    for (int j = 0; j < VPL; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long row = warp; row < rows; row += nwarps) {
        const float4* x4 = reinterpret_cast<const float4*>(x) + row * C4;
        const float4* d4 = reinterpret_cast<const float4*>(dy) + row * C4;
        float4* o4 = reinterpret_cast<float4*>(dx) + row * C4;
#pragma unroll
        for (int jb = 0; jb < VPL; jb += CH) {
            float4 v[CH], d[CH];
#pragma unroll
            for (int u = 0; u < CH; ++u) {
                const int c = lane + 32 * (jb + u);
                v[u] = d[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c < C4) {
                    v[u] = ldg_stream(x4 + c);
                    d[u] = ldg_stream(d4 + c);
                }
            }
#pragma unroll
            for (int u = 0; u < CH; ++u) {
                const int c = lane + 32 * (jb + u);
                if (c < C4) {
                    float4 t = d[u];
                    t.x *= act_df(v[u].x, ACT); t.y *= act_df(v[u].y, ACT); t.z *= act_df(v[u].z, ACT); t.w *= act_df(v[u].w, ACT);
                    o4[c] = t;
                    acc[jb + u].x += t.x; acc[jb + u].y += t.y; acc[jb + u].z += t.z; acc[jb + u].w += t.w;
                }
            }
        }
    }
    for (int w = 0; w < nw; ++w) {
        if (wid == w) {
#pragma unroll
            for (int j = 0; j < VPL; ++j) {
                const int c = lane + 32 * j;
                if (c < C4) {
                    if (w == 0) sm_acc[c] = acc[j];
                    else {
                        float4 t = sm_acc[c];
                        t.x += acc[j].x; t.y += acc[j].y; t.z += acc[j].z; t.w += acc[j].w;
                        sm_acc[c] = t;
                    }
                }
            }
        }
        __syncthreads();
    }
    float4* p = reinterpret_cast<float4*>(partials) + (long long)blockIdx.x * C4;
    for (int c = threadIdx.x; c < C4; c += blockDim.x) p[c] = sm_acc[c];
}

static int ln_grid(long long rows) {
    const long long by_rows = (rows + 7) / 8;
    return (int)std::max<long long>(1, std::min<long long>(by_rows, (long long)sm_count() * 8));
}
// the pipelined backward needs every lane column live (D = 128 * VPL) and 2 stages of 3 rows per warp in smem
static bool ln_bwd_pipe_ok(long long D) {
    return (tune() & PR_TUNE_LN_BWD_PIPE) && (D == 128 || D == 256 || D == 512 || D == 1024);
}
static int ln_bwd_grid(long long rows, long long D) {
    const long long by_rows = (rows + 7) / 8;
    const long long per_sm = ln_bwd_pipe_ok(D) ? (D <= 512 ? 2 : 1) : 4;   // pipelined: persistent, all CTAs resident
    return (int)std::max<long long>(1, std::min<long long>(by_rows, (long long)sm_count() * per_sm));
}

static int check_ln_common(const char* who, const void* h, const void* gamma, long long rows, long long D,
                           long long rows_per_seq, long long h_seq_stride, float p_pre, float p_post) {
    PR_CHECK_ARG(rows >= 0 && D > 0 && D % 4 == 0, "%s: bad shape rows=%lld D=%lld (D %% 4 must be 0)", who, rows, D);
    PR_CHECK_ARG(D <= 4096, "%s: D=%lld > 4096 unsupported", who, D);
    PR_CHECK_ARG(rows_per_seq > 0 && h_seq_stride % 4 == 0, "%s: rows_per_seq must be > 0 and h_seq_stride %% 4 == 0", who);
    PR_CHECK_ARG(p_pre >= 0.f && p_pre < 1.f && p_post >= 0.f && p_post < 1.f, "%s: dropout p outside [0,1)", who);
    PR_CHECK_ARG(rows == 0 || (h && gamma), "%s: null pointer", who);
    return PR_OK;
}

}  // namespace pr

using namespace pr;

#define PR_DISPATCH_ACT(act, CALL)                              \
    do {                                                        \
        switch (act) {                                          \
            case PR_ACT_GELU: CALL(PR_ACT_GELU); break;         \
            case PR_ACT_RELU: CALL(PR_ACT_RELU); break;         \
            case PR_ACT_SWISH: CALL(PR_ACT_SWISH); break;       \
            case PR_ACT_TANH: CALL(PR_ACT_TANH); break;         \
            case PR_ACT_QUICK_GELU: CALL(PR_ACT_QUICK_GELU); break; \
            default: CALL(PR_ACT_SIGMOID); break;               \
        }                                                       \
    } while (0)

#define PR_DISPATCH_VPL(D4, CALL)         \
    do {                                  \
        if ((D4) <= 32) { CALL(1); }      \
        else if ((D4) <= 64) { CALL(2); } \
        else if ((D4) <= 128) { CALL(4); }\
        else if ((D4) <= 256) { CALL(8); }\
        else if ((D4) <= 512) { CALL(16); }\
        else { CALL(32); }                \
    } while (0)

extern "C" int pr_add_ln_fwd_f32(const float* h, int64_t h_seq_stride, int64_t rows_per_seq, const float* res,
                                 int64_t res_period, const float* gamma, const float* beta, float eps, int64_t rows,
                                 int64_t D, float p_pre, float p_post, uint64_t seed, uint32_t stream_pre,
                                 uint32_t stream_post, float* y, float* mean, float* rstd, pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_ln_common("pr_add_ln_fwd_f32", h, gamma, rows, D, rows_per_seq, h_seq_stride, p_pre, p_post);
    if (rc) return rc;
    if (rows == 0) return PR_OK;
    PR_CHECK_ARG(beta && y && mean && rstd, "pr_add_ln_fwd_f32: null pointer");
    PR_CHECK_ARG(aligned16(h) && aligned16(res) && aligned16(gamma) && aligned16(beta) && aligned16(y),
                 "pr_add_ln_fwd_f32: pointers must be 16-byte aligned");
    LnArgs a{h, h_seq_stride, rows_per_seq, res, res_period, gamma, beta, eps, rows, (int)(D / 4),
             p_pre, p_post, seed, stream_pre, stream_post, (tune() & PR_TUNE_LN_L2_PREFETCH) ? 1 : 0};
    PR_SET_SEED_DEV(a);
    if ((tune() & PR_TUNE_LN_FWD_PIPE) && (D == 128 || D == 256 || D == 512 || D == 1024)) {
        const int pgrid = (int)std::max<long long>(1, std::min<long long>((rows + 7) / 8, (long long)sm_count() * (D <= 512 ? 2 : 1)));
#define FPIPE(V, S)                                                                                            \
    do {                                                                                                       \
        const size_t sm = (size_t)8 * S * 2 * D * 4 + 8 * S * sizeof(uint64_t);                                 \
        PR_CUDA_CALL(cudaFuncSetAttribute(add_ln_fwd_pipe_kernel<V, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
        add_ln_fwd_pipe_kernel<V, S><<<pgrid, 256, sm, stream>>>(a, y, mean, rstd);                            \
    } while (0)
        if (D == 128) FPIPE(1, 12);
        else if (D == 256) FPIPE(2, 6);
        else if (D == 512) FPIPE(4, 3);
        else FPIPE(8, 3);
#undef FPIPE
        PR_CUDA_LAUNCH_CHECK("add_ln_fwd_pipe_kernel");
        return PR_OK;
    }
    if ((tune() & PR_TUNE_LN_FWD_ROWS2) && a.D4 <= 128) {
        const int grid2 = (int)std::max<long long>(1, std::min<long long>((rows + 15) / 16, (long long)sm_count() * 3));   // persistent: 3 CTAs per SM resident
#define CALL2(V) add_ln_fwd_rows2_kernel<V><<<grid2, 256, 0, stream>>>(a, y, mean, rstd)
        if (a.D4 <= 32) { CALL2(1); } else if (a.D4 <= 64) { CALL2(2); } else { CALL2(4); }
#undef CALL2
        PR_CUDA_LAUNCH_CHECK("add_ln_fwd_rows2_kernel");
        return PR_OK;
    }
    const int grid = ln_grid(rows);
#define CALL(V) add_ln_fwd_kernel<V><<<grid, 256, 0, stream>>>(a, y, mean, rstd)
    PR_DISPATCH_VPL(a.D4, CALL);
#undef CALL
    PR_CUDA_LAUNCH_CHECK("add_ln_fwd_kernel");
    return PR_OK;
}

extern "C" int pr_add_ln_bwd_partials(int64_t rows, int64_t D) {
    if (rows <= 0) return 1;
    return ln_bwd_grid(rows, D);
}

static int add_ln_bwd_impl(bool want_dbias, bool h_is_z, const float* dy, const float* h, int64_t h_seq_stride, int64_t rows_per_seq,
                                 const float* res, int64_t res_period, const float* gamma, const float* mean,
                                 const float* rstd, int64_t rows, int64_t D, float p_pre, float p_post, uint64_t seed,
                                 uint32_t stream_pre, uint32_t stream_post, float* dh, int64_t dh_seq_stride,
                                 int dh_accumulate, float* dres, float* partials, int n_partials,
                                 pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_ln_common("pr_add_ln_bwd_f32", h, gamma, rows, D, rows_per_seq, h_seq_stride, p_pre, p_post);
    if (rc) return rc;
    PR_CHECK_ARG(partials, "pr_add_ln_bwd_f32: partials is null");
    const int grid = ln_bwd_grid(rows, D);
    PR_CHECK_ARG(n_partials == (rows > 0 ? grid : 1), "pr_add_ln_bwd_f32: n_partials=%d, expected %d (pr_add_ln_bwd_partials)",
                 n_partials, rows > 0 ? grid : 1);
    if (rows == 0) {
        PR_CUDA_CALL(cudaMemsetAsync(partials, 0, sizeof(float) * (want_dbias ? 3 : 2) * (size_t)D, stream));
        return PR_OK;
    }
    PR_CHECK_ARG(dy && mean && rstd && dh, "pr_add_ln_bwd_f32: null pointer");
    PR_CHECK_ARG(dh_seq_stride >= 0 && dh_seq_stride % 4 == 0, "pr_add_ln_bwd_f32: dh_seq_stride must be >= 0 and %% 4 == 0");
    PR_CHECK_ARG(aligned16(dy) && aligned16(h) && aligned16(res) && aligned16(gamma) && aligned16(dh) && aligned16(dres) &&
                     aligned16(partials),
                 "pr_add_ln_bwd_f32: pointers must be 16-byte aligned");
    LnArgs a{h, h_seq_stride, rows_per_seq, res, res_period, gamma, nullptr, 0.f, rows, (int)(D / 4),
             p_pre, p_post, seed, stream_pre, stream_post, 0};
    PR_SET_SEED_DEV(a);
    a.l2_prefetch = (tune() & PR_TUNE_LN_L2_PREFETCH) ? 1 : 0;
    a.h_is_z = h_is_z ? 1 : 0;
    PR_CHECK_ARG(!h_is_z || !res, "pr_add_ln_bwd_bias_z_f32: z already contains the residual");
    if (ln_bwd_pipe_ok(D)) {
#define PIPE(V, S, NS)                                                                                               \
    do {                                                                                                             \
        const size_t sm = (size_t)8 * S * NS * D * 4 + (size_t)3 * D * 4 + 8 * S * sizeof(uint64_t);                  \
        auto kt = add_ln_bwd_pipe_kernel<V, true, S, NS>;                                                            \
        auto kf = add_ln_bwd_pipe_kernel<V, false, S, NS>;                                                           \
        PR_CUDA_CALL(cudaFuncSetAttribute(want_dbias ? kt : kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
        (want_dbias ? kt : kf)<<<grid, 256, sm, stream>>>(a, dy, mean, rstd, dh, dh_seq_stride, dh_accumulate, dres, partials); \
    } while (0)
        if (res) {                      // three source rows per stage
            if (D == 128) PIPE(1, 8, 3);
            else if (D == 256) PIPE(2, 4, 3);
            else if (D == 512) PIPE(4, 2, 3);
            else PIPE(8, 2, 3);
        } else {                        // two source rows per stage: the same shared memory holds 1.5x the stages
            if (D == 128) PIPE(1, 12, 2);
            else if (D == 256) PIPE(2, 6, 2);
            else if (D == 512) PIPE(4, 3, 2);
            else PIPE(8, 3, 2);
        }
#undef PIPE
        PR_CUDA_LAUNCH_CHECK("add_ln_bwd_pipe_kernel");
        return PR_OK;
    }
    const size_t smem = (size_t)(want_dbias ? 3 : 2) * D * sizeof(float);
#define CALL(V)                                                                                                     \
    do {                                                                                                            \
        if (want_dbias) add_ln_bwd_kernel<V, true><<<grid, 256, smem, stream>>>(a, dy, mean, rstd, dh, dh_seq_stride, dh_accumulate, dres, partials); \
        else add_ln_bwd_kernel<V, false><<<grid, 256, smem, stream>>>(a, dy, mean, rstd, dh, dh_seq_stride, dh_accumulate, dres, partials); \
    } while (0)
    PR_DISPATCH_VPL(a.D4, CALL);
#undef CALL
    PR_CUDA_LAUNCH_CHECK("add_ln_bwd_kernel");
    return PR_OK;
}

extern "C" int pr_add_ln_bwd_f32(const float* dy, const float* h, int64_t h_seq_stride, int64_t rows_per_seq,
                                 const float* res, int64_t res_period, const float* gamma, const float* mean,
                                 const float* rstd, int64_t rows, int64_t D, float p_pre, float p_post, uint64_t seed,
                                 uint32_t stream_pre, uint32_t stream_post, float* dh, int64_t dh_seq_stride,
                                 int dh_accumulate, float* dres, float* partials, int n_partials, pr_stream_t stream_) {
    return add_ln_bwd_impl(false, false, dy, h, h_seq_stride, rows_per_seq, res, res_period, gamma, mean, rstd, rows, D, p_pre, p_post,
                           seed, stream_pre, stream_post, dh, dh_seq_stride, dh_accumulate, dres, partials, n_partials, stream_);
}

extern "C" int pr_add_ln_bwd_bias_f32(const float* dy, const float* h, int64_t h_seq_stride, int64_t rows_per_seq,
                                      const float* res, int64_t res_period, const float* gamma, const float* mean,
                                      const float* rstd, int64_t rows, int64_t D, float p_pre, float p_post, uint64_t seed,
                                      uint32_t stream_pre, uint32_t stream_post, float* dh, int64_t dh_seq_stride,
                                      int dh_accumulate, float* dres, float* partials, int n_partials, pr_stream_t stream_) {
    return add_ln_bwd_impl(true, false, dy, h, h_seq_stride, rows_per_seq, res, res_period, gamma, mean, rstd, rows, D, p_pre, p_post,
                           seed, stream_pre, stream_post, dh, dh_seq_stride, dh_accumulate, dres, partials, n_partials, stream_);
}

extern "C" int pr_add_ln_bwd_bias_z_f32(const float* dy, const float* z, const float* gamma, const float* mean, const float* rstd,
                                        int64_t rows, int64_t D, float p_pre, uint64_t seed, uint32_t stream_pre, float* dh,
                                        float* dz, float* partials, int n_partials, pr_stream_t stream_) {
    return add_ln_bwd_impl(true, true, dy, z, 0, rows > 0 ? rows : 1, nullptr, 0, gamma, mean, rstd, rows, D, p_pre, 0.f, seed,
                           stream_pre, 0, dh, 0, 0, dz, partials, n_partials, stream_);
}

extern "C" int pr_colsum_f32(const float* partials, int n_mats, int n_partials, int64_t D, float* out,
                             pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PR_CHECK_ARG(n_partials >= 1 && D > 0 && n_mats >= 1 && n_mats <= 65535, "pr_colsum_f32: bad shape");
    PR_CHECK_ARG(partials && out, "pr_colsum_f32: null pointer");
    colsum_kernel<<<dim3((unsigned)((D + 31) / 32), (unsigned)n_mats), dim3(32, 32), 0, stream>>>(partials, n_partials, D, out);
    PR_CUDA_LAUNCH_CHECK("colsum_kernel");
    return PR_OK;
}

// ---- column sums of a row-major matrix (bias gradient of a Linear whose output gradient no kernel of ours produced with
// partials attached: the fused q|k|v projection, REC/model/layers.py:586-588).  Stage 1: CTA (row block, 128-column block), lane =
// float4 column group, 8 warps stride the rows with 4 loads in flight, fixed-order reduction over the warps -> partials
// [row_blocks, C]; stage 2: colsum_kernel.  Deterministic; reads the matrix once at HBM speed.
__global__ void __launch_bounds__(256) colsum_rows_kernel(const float* __restrict__ x, long long M, long long C, int rows_per_cta,
                                                          float* __restrict__ partials) {
    __shared__ float4 sm[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long c0 = ((long long)blockIdx.x * 32 + lane) * 4;
    const long long r0 = (long long)blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c0 < C) {
        long long r = r0 + warp;
        for (; r + 24 < r1; r += 32) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = ldg_stream(reinterpret_cast<const float4*>(x + (r + 8 * u) * C + c0));
#pragma unroll
            for (int u = 0; u < 4; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
        for (; r < r1; r += 8) {
            const float4 v = ldg_stream(reinterpret_cast<const float4*>(x + r * C + c0));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    sm[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && c0 < C) {
        float4 t = sm[0][lane];
#pragma unroll
        for (int w = 1; w < 8; ++w) { t.x += sm[w][lane].x; t.y += sm[w][lane].y; t.z += sm[w][lane].z; t.w += sm[w][lane].w; }
        *reinterpret_cast<float4*>(partials + (long long)blockIdx.y * C + c0) = t;
    }
}

extern "C" int pr_colsum_rows_partials(int64_t M, int64_t C) {
    if (M <= 0 || C <= 0) return 0;
    const long long col_blocks = (C + 127) / 128;
    long long row_blocks = std::max<long long>(1, (long long)sm_count() * 4 / col_blocks);
    row_blocks = std::min<long long>(row_blocks, (M + 63) / 64);
    return (int)row_blocks;
}

extern "C" int pr_colsum_rows_f32(const float* x, int64_t M, int64_t C, float* partials, int n_partials, float* out,
                                  pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PR_CHECK_ARG(M > 0 && C > 0 && C % 4 == 0, "pr_colsum_rows_f32: bad shape M=%lld C=%lld (C %% 4)", (long long)M, (long long)C);
    PR_CHECK_ARG(x && partials && out && aligned16(x) && aligned16(partials), "pr_colsum_rows_f32: null/unaligned pointer");
    PR_CHECK_ARG(n_partials == pr_colsum_rows_partials(M, C), "pr_colsum_rows_f32: n_partials=%d, expected %d", n_partials,
                 pr_colsum_rows_partials(M, C));
    const int rows_per_cta = (int)((M + n_partials - 1) / n_partials);
    colsum_rows_kernel<<<dim3((unsigned)((C + 127) / 128), (unsigned)n_partials), 256, 0, stream>>>(x, M, C, rows_per_cta, partials);
    PR_CUDA_LAUNCH_CHECK("colsum_rows_kernel");
    return pr_colsum_f32(partials, 1, n_partials, C, out, stream_);
}

extern "C" int pr_act_fwd_f32(const float* x, int64_t n, int act, float* y, pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PR_CHECK_ARG(n >= 0 && act >= 0 && act <= PR_ACT_QUICK_GELU, "pr_act_fwd_f32: bad n/act");
    if (n == 0) return PR_OK;
    PR_CHECK_ARG(x && y && aligned16(x) && aligned16(y), "pr_act_fwd_f32: null/unaligned pointer");
    const int grid = (int)std::max<long long>(1, std::min<long long>((n / 4 + 255) / 256, (long long)sm_count() * 16));
#define CALLF(A) act_fwd_kernel<A, 4><<<grid, 256, 0, stream>>>(x, n, y)
    PR_DISPATCH_ACT(act, CALLF);
#undef CALLF
    PR_CUDA_LAUNCH_CHECK("act_fwd_kernel");
    return PR_OK;
}

extern "C" int pr_act_bwd_f32(const float* x, const float* dy, int64_t n, int act, float* dx, pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PR_CHECK_ARG(n >= 0 && act >= 0 && act <= PR_ACT_QUICK_GELU, "pr_act_bwd_f32: bad n/act");
    if (n == 0) return PR_OK;
    PR_CHECK_ARG(x && dy && dx && aligned16(x) && aligned16(dy) && aligned16(dx), "pr_act_bwd_f32: null/unaligned pointer");
    const int grid = (int)std::max<long long>(1, std::min<long long>((n / 4 + 255) / 256, (long long)sm_count() * 16));
#define CALLB(A) act_bwd_kernel<A, 4><<<grid, 256, 0, stream>>>(x, dy, n, dx)
    PR_DISPATCH_ACT(act, CALLB);
#undef CALLB
    PR_CUDA_LAUNCH_CHECK("act_bwd_kernel");
    return PR_OK;
}

extern "C" int pr_act_bwd_bias_partials(int64_t rows, int64_t cols) {
    (void)cols;
    if (rows <= 0) return 1;
    return (int)std::max<long long>(1, std::min<long long>((rows + 7) / 8, (long long)sm_count() * 8));
}

extern "C" int pr_act_bwd_bias_f32(const float* x, const float* dy, int64_t rows, int64_t cols, int act, float* dx,
                                   float* partials, int n_partials, pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PR_CHECK_ARG(rows > 0 && cols > 0 && cols % 4 == 0 && cols <= 8192, "pr_act_bwd_bias_f32: bad shape rows=%lld cols=%lld", (long long)rows,
                 (long long)cols);
    PR_CHECK_ARG(act >= 0 && act <= PR_ACT_QUICK_GELU, "pr_act_bwd_bias_f32: bad act");
    PR_CHECK_ARG(x && dy && dx && partials && aligned16(x) && aligned16(dy) && aligned16(dx) && aligned16(partials),
                 "pr_act_bwd_bias_f32: null/unaligned pointer");
    const int grid = pr_act_bwd_bias_partials(rows, cols);
    PR_CHECK_ARG(n_partials == grid, "pr_act_bwd_bias_f32: n_partials=%d, expected %d", n_partials, grid);
    const int C4 = (int)(cols / 4);
    const size_t smem = (size_t)cols * sizeof(float);
#define CALLV(A, V) act_bwd_bias_kernel<A, V><<<grid, 256, smem, stream>>>(x, dy, rows, C4, dx, partials)
#define CALLA(A)                         \
    do {                                 \
        if (C4 <= 32) CALLV(A, 1);       \
        else if (C4 <= 64) CALLV(A, 2);  \
        else if (C4 <= 128) CALLV(A, 4); \
        else if (C4 <= 256) CALLV(A, 8); \
        else if (C4 <= 512) CALLV(A, 16);\
        else if (C4 <= 1024) CALLV(A, 32);\
        else CALLV(A, 64);               \
    } while (0)
    PR_DISPATCH_ACT(act, CALLA);
#undef CALLV
#undef CALLA
    PR_CUDA_LAUNCH_CHECK("act_bwd_bias_kernel");
    return PR_OK;
}
