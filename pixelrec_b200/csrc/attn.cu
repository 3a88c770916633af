// pixelrec_b200 -- K4+K6: SASRec causal self-attention core for tiny L (<= 64) and large batch.
//   replaces get_attention_mask (REC/model/IDNet/sasrec.py:119-126) and the attention core of
//   MultiHeadAttention.forward (REC/model/layers.py:590-612): ~8 launches per layer in the reference
//   (2 batched GEMMs, div, add-mask, softmax, dropout, permute/contiguous) and a materialised [B,1,L,L] mask.
//
// Shape regime: L in {10,20,50}, dh in {16..512}, B*h in the tens of thousands -> the op is HBM-bound
// (16*B*L*D bytes forward), the L x L score matrix never leaves the SM.
//
// Structure (warp-specialised, TMA-fed):
//   * warp NCW (producer): streams [L x CW] fp32 tiles (CW = min(dh,128) floats of one head) into a ring of
//     shared-memory stages with cp.async.bulk (one bulk copy per row, issued by one lane each), completion
//     counted by the stage's `full` mbarrier; waits on `empty` mbarriers before reusing a stage.
//   * warps 0..NCW-1 (consumers): each owns whole (batch, head) items; S = Q K^T accumulates over the
//     dh/CW column chunks in registers (4x8 lane grid, staggered float4 reads -> conflict-free LDS.128),
//     masking + softmax by 8-lane shuffles, P kept in a warp-private smem tile, then O = P V per chunk.
//   * mask is computed in registers from key_ids (never materialised); dropout masks come from Philox.
#include <algorithm>
#include <math.h>

#include "common.cuh"

namespace pr {

constexpr int ATT_MAX_STAGES = 32;

struct AttnArgs {
    const float* q; const float* k; const float* v; long long ld;
    const long long* key_ids;
    int B, L, h, dh, nc, causal;
    float p_drop; unsigned long long seed; unsigned rng_stream;
    float inv_sqrt;
    float* ctx; float* probs;                     // forward outputs
    const float* dctx; float* dq; float* dk; float* dv; long long ld_grad;   // backward
    int stages;
#ifdef PR_SEED_DEV
    const unsigned long long* seed_dev;   // device-side seed offset (pr_set_seed_device)
#endif
};

template <int LMAX>
struct AttnCfg {
    static constexpr int RI = (LMAX + 3) / 4;           // score rows per lane   (lane grid 4 x 8)
    static constexpr int CJ = (LMAX + 7) / 8;           // score cols per lane
    static constexpr int LP = ((LMAX + 7) / 8) * 8;     // padded row length of the warp-private L x L tiles
    static constexpr int NCW = (LMAX > 32) ? 4 : 8;     // consumer warps (forward)
    static constexpr int NCW_BWD = (LMAX > 32) ? 2 : 8; // consumer warps (backward; 3 private tiles each)
};

// acc[ii][jj] += sum_d X[i][d] * Y[j][d]  over one [L x CW] tile pair; i = a + 4*ii, j = b + 8*jj
template <int LMAX, int CW4>
__device__ __forceinline__ void tile_dot_accum(const float4* __restrict__ Xs, const float4* __restrict__ Ys, int a,
                                               int b, float (&acc)[AttnCfg<LMAX>::RI][AttnCfg<LMAX>::CJ]) {
    constexpr int RI = AttnCfg<LMAX>::RI, CJ = AttnCfg<LMAX>::CJ;
    int xrow[RI], yrow[CJ];
#pragma unroll
    for (int ii = 0; ii < RI; ++ii) xrow[ii] = min(a + 4 * ii, LMAX - 1) * CW4;
#pragma unroll
    for (int jj = 0; jj < CJ; ++jj) yrow[jj] = min(b + 8 * jj, LMAX - 1) * CW4;
#pragma unroll 2
    for (int cc = 0; cc < CW4; ++cc) {
        const int ch = (cc + b) % CW4;  // stagger: the 8 lanes of a quarter-warp hit 8 different 16-B bank groups
        float4 xv[RI], yv[CJ];
#pragma unroll
        for (int ii = 0; ii < RI; ++ii) xv[ii] = Xs[xrow[ii] + ch];
#pragma unroll
        for (int jj = 0; jj < CJ; ++jj) yv[jj] = Ys[yrow[jj] + ch];
#pragma unroll
        for (int ii = 0; ii < RI; ++ii)
#pragma unroll
            for (int jj = 0; jj < CJ; ++jj) {
                acc[ii][jj] = fmaf(xv[ii].x, yv[jj].x, acc[ii][jj]);
                acc[ii][jj] = fmaf(xv[ii].y, yv[jj].y, acc[ii][jj]);
                acc[ii][jj] = fmaf(xv[ii].z, yv[jj].z, acc[ii][jj]);
                acc[ii][jj] = fmaf(xv[ii].w, yv[jj].w, acc[ii][jj]);
            }
    }
}

// out[r][d] = sum_{s < L} M[s][r] * X[s][d]  for one [L x CW] tile X; M is a warp-private smem matrix
// [LMAX][LP] (contiguous in r).  Lane owns float4 column d4 and rows r = rg + RG*ii.  Rows r < L are
// stored to out + r*out_ld (+ d4).
template <int LMAX, int CW4>
__device__ __forceinline__ void tile_outer_store(const float* __restrict__ M, const float4* __restrict__ Xs, int L,
                                                 int lane, float* __restrict__ out, long long out_ld) {
    constexpr int LP = AttnCfg<LMAX>::LP;
    constexpr int RG = 32 / CW4;
    constexpr int NR = (LMAX + RG - 1) / RG;
    const int d4 = lane % CW4, rg = lane / CW4;
    float4 o[NR];
#pragma unroll
    for (int ii = 0; ii < NR; ++ii) o[ii] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < L; ++s) {
        const float4 x = Xs[s * CW4 + d4];
        const float* Mr = M + s * LP + rg;
#pragma unroll
        for (int ii = 0; ii < NR; ++ii) {
            const float m = Mr[RG * ii];
            o[ii].x = fmaf(m, x.x, o[ii].x);
            o[ii].y = fmaf(m, x.y, o[ii].y);
            o[ii].z = fmaf(m, x.z, o[ii].z);
            o[ii].w = fmaf(m, x.w, o[ii].w);
        }
    }
#pragma unroll
    for (int ii = 0; ii < NR; ++ii) {
        const int r = rg + RG * ii;
        if (r < L) *reinterpret_cast<float4*>(out + (long long)r * out_ld + d4 * 4) = o[ii];
    }
}

// NOTE on the `produced` counter: an mbarrier parity wait can only tell the current phase from the previous
// one.  Several consumer warps wait on the SAME stage barrier for different phases (tile t and tile t+S belong
// to different warps), so a warp may start waiting for tile t+S before tile t has even landed -- its parity
// wait would then succeed spuriously.  Consumers therefore first spin on `produced` (tiles armed so far,
// bumped in order by the producer AFTER it saw tile t-S released), which guarantees the barrier is already in
// tile t's phase when the parity wait starts.
struct TileRing {
    unsigned char* tiles; uint64_t* full; uint64_t* empty; volatile unsigned long long* produced; int S;
    uint32_t tile_bytes;
    __device__ __forceinline__ const float4* wait_full(long long t) const {
        const int s = (int)(t % S);
        while (*produced <= (unsigned long long)t) {
        }
        mbar_wait(&full[s], (uint32_t)((t / S) & 1));
        return reinterpret_cast<const float4*>(tiles + (size_t)s * tile_bytes);
    }
    __device__ __forceinline__ void release(long long t, int lane) const {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[(int)(t % S)]);
    }
};

// producer: copy rows [0,L) x [col0, col0+CW) of a [*, ld] fp32 matrix into stage t
__device__ __forceinline__ void produce_tile(const TileRing& ring, long long t, const float* src, long long ld, int L,
                                             int CW, int lane) {
    const int s = (int)(t % ring.S);
    mbar_wait(&ring.empty[s], (uint32_t)(((t / ring.S) & 1) ^ 1));
    const uint32_t row_bytes = (uint32_t)CW * 4u;
    if (lane == 0) mbar_arrive_expect_tx(&ring.full[s], row_bytes * (uint32_t)L);
    __syncwarp();
    unsigned char* dst = ring.tiles + (size_t)s * ring.tile_bytes;
    for (int r = lane; r < L; r += 32) bulk_g2s(dst + (size_t)r * row_bytes, src + (long long)r * ld, row_bytes, &ring.full[s]);
    if (lane == 0) *ring.produced = (unsigned long long)t + 1ull;
}

__device__ __forceinline__ TileRing ring_setup(unsigned char* smem, int S, uint32_t tile_bytes, size_t private_bytes,
                                               unsigned char** private_base) {
    // layout: [tiles S*tile_bytes][private][full 32][empty 32][produced]
    TileRing r;
    r.tiles = smem;
    r.S = S;
    r.tile_bytes = tile_bytes;
    *private_base = smem + (size_t)S * tile_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(*private_base + private_bytes);
    r.full = bars;
    r.empty = bars + ATT_MAX_STAGES;
    r.produced = reinterpret_cast<volatile unsigned long long*>(bars + 2 * ATT_MAX_STAGES);
    if (threadIdx.x == 0) {
        *r.produced = 0ull;
        for (int s = 0; s < S; ++s) {
            mbar_init(&r.full[s], 1);
            mbar_init(&r.empty[s], 1);
        }
        fence_mbar_init();
    }
    __syncthreads();
    return r;
}

// dropout keep-mask for score entry (row i, col j = b + 8*jj) of `item`: bit jj of keep_bits8(Philox(counter
// (item*L + i)*8 + b, stream))   (oracle/philox_np.py attn_keep_scale restates it; CJ <= 8)
template <int CJ>
__device__ __forceinline__ void attn_keep(const Philox& ph, unsigned stream, long long item, int L, int i, int b,
                                          unsigned thr, bool (&keep)[CJ]) {
    static_assert(CJ <= 8, "one Philox draw covers 8 score columns per lane");
    const unsigned m = keep_bits8(ph((unsigned long long)((item * L + i) * 8 + b), stream), thr);
#pragma unroll
    for (int jj = 0; jj < CJ; ++jj) keep[jj] = (m >> jj) & 1u;
}

// =============================================================================================== forward
template <int LMAX, int CW4>
__global__ void __launch_bounds__((AttnCfg<LMAX>::NCW + 1) * 32, 1) attn_fwd_kernel(AttnArgs A) {
    using C = AttnCfg<LMAX>;
    constexpr int RI = C::RI, CJ = C::CJ, LP = C::LP, NCW = C::NCW, CW = CW4 * 4;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* priv;
    const TileRing ring = ring_setup(smem, A.stages, (uint32_t)(LMAX * CW * 4), (size_t)NCW * LMAX * LP * 4, &priv);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = A.L, nc = A.nc, T = 3 * nc;
    const long long n_items = (long long)A.B * A.h;
    // contiguous block of items per CTA: the h heads of a sequence (adjacent 512-byte pieces of the same rows) and
    // consecutive sequences are read by the same SM back to back -> DRAM pages / L2 lines are reused while hot
    // (a round-robin assignment scattered every row over 12 SMs and ran at 0.28 of the HBM roofline)
    const long long items_per_cta = (n_items + gridDim.x - 1) / gridDim.x;
    const long long item_lo = (long long)blockIdx.x * items_per_cta;
    const long long item_hi = min(n_items, item_lo + items_per_cta);

    if (warp == NCW) {  // ------------------------------------------------ producer warp
        long long t = 0;
        for (long long item = item_lo; item < item_hi; ++item) {
            const long long b = item / A.h;
            const int hd = (int)(item - b * A.h);
            const long long off = b * L * A.ld + (long long)hd * A.dh;
            for (int c = 0; c < nc; ++c) {
                produce_tile(ring, t++, A.q + off + c * CW, A.ld, L, CW, lane);
                produce_tile(ring, t++, A.k + off + c * CW, A.ld, L, CW, lane);
            }
            for (int c = 0; c < nc; ++c) produce_tile(ring, t++, A.v + off + c * CW, A.ld, L, CW, lane);
        }
        return;
    }
    // ---------------------------------------------------------------------- consumer warps
    float* Pt = reinterpret_cast<float*>(priv) + (size_t)warp * LMAX * LP;  // Pt[j][i] = dropped P[i][j]
    const int a = lane >> 3, b8 = lane & 7;
    const Philox ph(PR_SEED(A));
    const unsigned thr = drop_threshold(A.p_drop);
    const float inv_keep = 1.0f / (1.0f - A.p_drop);
    long long n = 0;
    for (long long item = item_lo; item < item_hi; ++item, ++n) {
        if ((int)(n % NCW) != warp) continue;
        const long long bb = item / A.h;
        const int hd = (int)(item - bb * A.h);
        const long long t0 = n * T;
        float acc[RI][CJ];
#pragma unroll
        for (int ii = 0; ii < RI; ++ii)
#pragma unroll
            for (int jj = 0; jj < CJ; ++jj) acc[ii][jj] = 0.f;
        for (int c = 0; c < nc; ++c) {
            const float4* Qs = ring.wait_full(t0 + 2 * c);
            const float4* Ks = ring.wait_full(t0 + 2 * c + 1);
            tile_dot_accum<LMAX, CW4>(Qs, Ks, a, b8, acc);
            ring.release(t0 + 2 * c, lane);
            ring.release(t0 + 2 * c + 1, lane);
        }
        // ---- mask (sasrec.py:119-126) + softmax (layers.py:595-604) in registers
        bool kvalid[CJ];
#pragma unroll
        for (int jj = 0; jj < CJ; ++jj) {
            const int j = b8 + 8 * jj;
            kvalid[jj] = (j < L) && (A.key_ids == nullptr || A.key_ids[bb * L + j] != 0);
        }
#pragma unroll
        for (int ii = 0; ii < RI; ++ii) {
            const int i = a + 4 * ii;
            float mx = -INFINITY;
#pragma unroll
            for (int jj = 0; jj < CJ; ++jj) {
                const int j = b8 + 8 * jj;
                const bool ok = kvalid[jj] && (!A.causal || j <= i);
                float s = acc[ii][jj] * A.inv_sqrt + (ok ? 0.0f : -1e9f);
                if (j >= L) s = -INFINITY;
                acc[ii][jj] = s;
                mx = fmaxf(mx, s);
            }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
            float sum = 0.f;
#pragma unroll
            for (int jj = 0; jj < CJ; ++jj) {
                const float e = (b8 + 8 * jj < L) ? expf(acc[ii][jj] - mx) : 0.f;
                acc[ii][jj] = e;
                sum += e;
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            sum += __shfl_xor_sync(0xffffffffu, sum, 4);
            const float inv = 1.0f / sum;
            bool keep[CJ];
            if (A.p_drop > 0.f) attn_keep<CJ>(ph, A.rng_stream, item, L, i, b8, thr, keep);
#pragma unroll
            for (int jj = 0; jj < CJ; ++jj) {
                const int j = b8 + 8 * jj;
                const float p = acc[ii][jj] * inv;
                if (i < L && j < L) {
                    A.probs[(item * L + i) * L + j] = p;
                    float pd = p;
                    if (A.p_drop > 0.f) pd = keep[jj] ? p * inv_keep : 0.f;
                    Pt[j * LP + i] = pd;
                }
            }
        }
        __syncwarp();
        // ---- O = drop(P) V, one [L x CW] chunk of the head at a time (layers.py:609-612)
        for (int c = 0; c < nc; ++c) {
            const long long tv = t0 + 2 * nc + c;
            const float4* Vs = ring.wait_full(tv);
            float* out = A.ctx + bb * L * (long long)(A.h * A.dh) + (long long)hd * A.dh + c * CW;
            tile_outer_store<LMAX, CW4>(Pt, Vs, L, lane, out, (long long)A.h * A.dh);
            ring.release(tv, lane);
        }
        __syncwarp();
    }
}

// =============================================================================================== backward
//   dPd = dO V^T ; dP = drop'(dPd) ; dS = P o (dP - rowsum(dP o P)) / sqrt(dh)
//   dV = Pd^T dO ; dQ = dS K ; dK = dS^T Q
template <int LMAX, int CW4>
__global__ void __launch_bounds__((AttnCfg<LMAX>::NCW_BWD + 1) * 32, 1) attn_bwd_kernel(AttnArgs A) {
    using C = AttnCfg<LMAX>;
    constexpr int RI = C::RI, CJ = C::CJ, LP = C::LP, NCW = C::NCW_BWD, CW = CW4 * 4;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* priv;
    const TileRing ring = ring_setup(smem, A.stages, (uint32_t)(LMAX * CW * 4), (size_t)NCW * 3 * LMAX * LP * 4, &priv);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = A.L, nc = A.nc, T = 5 * nc;
    const long long n_items = (long long)A.B * A.h;
    // contiguous block of items per CTA: the h heads of a sequence (adjacent 512-byte pieces of the same rows) and
    // consecutive sequences are read by the same SM back to back -> DRAM pages / L2 lines are reused while hot
    // (a round-robin assignment scattered every row over 12 SMs and ran at 0.28 of the HBM roofline)
    const long long items_per_cta = (n_items + gridDim.x - 1) / gridDim.x;
    const long long item_lo = (long long)blockIdx.x * items_per_cta;
    const long long item_hi = min(n_items, item_lo + items_per_cta);
    const long long Dm = (long long)A.h * A.dh;

    if (warp == NCW) {  // producer: (dO_c, V_c) x nc, then (dO_c, K_c, Q_c) x nc
        long long t = 0;
        for (long long item = item_lo; item < item_hi; ++item) {
            const long long b = item / A.h;
            const int hd = (int)(item - b * A.h);
            const long long off = b * L * A.ld + (long long)hd * A.dh;
            const long long offo = b * L * Dm + (long long)hd * A.dh;
            for (int c = 0; c < nc; ++c) {
                produce_tile(ring, t++, A.dctx + offo + c * CW, Dm, L, CW, lane);
                produce_tile(ring, t++, A.v + off + c * CW, A.ld, L, CW, lane);
            }
            for (int c = 0; c < nc; ++c) {
                produce_tile(ring, t++, A.dctx + offo + c * CW, Dm, L, CW, lane);
                produce_tile(ring, t++, A.k + off + c * CW, A.ld, L, CW, lane);
                produce_tile(ring, t++, A.q + off + c * CW, A.ld, L, CW, lane);
            }
        }
        return;
    }
    float* Pd_s = reinterpret_cast<float*>(priv) + (size_t)warp * 3 * LMAX * LP;  // Pd_s[i][j]
    float* dS_s = Pd_s + LMAX * LP;                                                 // dS_s[i][j]
    float* dS_t = dS_s + LMAX * LP;                                                 // dS_t[j][i]
    const int a = lane >> 3, b8 = lane & 7;
    const Philox ph(PR_SEED(A));
    const unsigned thr = drop_threshold(A.p_drop);
    const float inv_keep = 1.0f / (1.0f - A.p_drop);
    long long n = 0;
    for (long long item = item_lo; item < item_hi; ++item, ++n) {
        if ((int)(n % NCW) != warp) continue;
        const long long bb = item / A.h;
        const int hd = (int)(item - bb * A.h);
        const long long t0 = n * T;
        float acc[RI][CJ];
#pragma unroll
        for (int ii = 0; ii < RI; ++ii)
#pragma unroll
            for (int jj = 0; jj < CJ; ++jj) acc[ii][jj] = 0.f;
        for (int c = 0; c < nc; ++c) {
            const float4* dOs = ring.wait_full(t0 + 2 * c);
            const float4* Vs = ring.wait_full(t0 + 2 * c + 1);
            tile_dot_accum<LMAX, CW4>(dOs, Vs, a, b8, acc);
            ring.release(t0 + 2 * c, lane);
            ring.release(t0 + 2 * c + 1, lane);
        }
#pragma unroll
        for (int ii = 0; ii < RI; ++ii) {
            const int i = a + 4 * ii;
            bool keep[CJ];
            if (A.p_drop > 0.f) attn_keep<CJ>(ph, A.rng_stream, item, L, i, b8, thr, keep);
            float p[CJ], dp[CJ];
            float rd = 0.f;
#pragma unroll
            for (int jj = 0; jj < CJ; ++jj) {
                const int j = b8 + 8 * jj;
                const bool in = (i < L) && (j < L);
                p[jj] = in ? A.probs[(item * L + i) * L + j] : 0.f;
                float d = in ? acc[ii][jj] : 0.f;
                if (A.p_drop > 0.f) d = keep[jj] ? d * inv_keep : 0.f;
                dp[jj] = d;
                rd = fmaf(d, p[jj], rd);
            }
            rd += __shfl_xor_sync(0xffffffffu, rd, 1);
            rd += __shfl_xor_sync(0xffffffffu, rd, 2);
            rd += __shfl_xor_sync(0xffffffffu, rd, 4);
#pragma unroll
            for (int jj = 0; jj < CJ; ++jj) {
                const int j = b8 + 8 * jj;
                if (i < L && j < L) {
                    const float ds = p[jj] * (dp[jj] - rd) * A.inv_sqrt;
                    float pd = p[jj];
                    if (A.p_drop > 0.f) pd = keep[jj] ? pd * inv_keep : 0.f;
                    Pd_s[i * LP + j] = pd;
                    dS_s[i * LP + j] = ds;
                    dS_t[j * LP + i] = ds;
                }
            }
        }
        __syncwarp();
        for (int c = 0; c < nc; ++c) {
            const long long tb = t0 + 2 * nc + 3 * c;
            const long long go = bb * L * A.ld_grad + (long long)hd * A.dh + c * CW;
            const float4* dOs = ring.wait_full(tb);
            tile_outer_store<LMAX, CW4>(Pd_s, dOs, L, lane, A.dv + go, A.ld_grad);   // dV[j] = sum_i Pd[i][j] dO[i]
            ring.release(tb, lane);
            const float4* Ks = ring.wait_full(tb + 1);
            tile_outer_store<LMAX, CW4>(dS_t, Ks, L, lane, A.dq + go, A.ld_grad);    // dQ[i] = sum_j dS[i][j] K[j]
            ring.release(tb + 1, lane);
            const float4* Qs = ring.wait_full(tb + 2);
            tile_outer_store<LMAX, CW4>(dS_s, Qs, L, lane, A.dk + go, A.ld_grad);    // dK[j] = sum_i dS[i][j] Q[i]
            ring.release(tb + 2, lane);
        }
        __syncwarp();
    }
}

// ----------------------------------------------------------------------------------------------- host
template <int LMAX, int CW4, bool BWD>
static int launch_attn(AttnArgs& A, cudaStream_t stream) {
    using C = AttnCfg<LMAX>;
    constexpr int NCW = BWD ? C::NCW_BWD : C::NCW;
    const size_t tile = (size_t)LMAX * CW4 * 16;
    const size_t priv = (size_t)NCW * (BWD ? 3 : 1) * LMAX * C::LP * 4;
    const size_t bars = (size_t)2 * ATT_MAX_STAGES * 8 + 16;
    const size_t budget = 220 * 1024;
    PR_CHECK_ARG(priv + bars + 4 * tile <= budget, "attention: L=%d dh=%d does not fit shared memory", A.L, A.dh);
    int S = (int)((budget - priv - bars) / tile);
    S = std::min(S, ATT_MAX_STAGES);
    const int tiles_per_item = (BWD ? 5 : 3) * A.nc;
    S = std::min(S, std::max(4, tiles_per_item * NCW * 2));  // never more than two items per consumer in flight
    A.stages = S;
    const size_t smem = (size_t)S * tile + priv + bars;
    auto kern = BWD ? attn_bwd_kernel<LMAX, CW4> : attn_fwd_kernel<LMAX, CW4>;
    PR_CUDA_CALL(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long n_items = (long long)A.B * A.h;
    const int grid = (int)std::max<long long>(1, std::min<long long>((n_items + NCW - 1) / NCW, sm_count()));
    kern<<<grid, (NCW + 1) * 32, smem, stream>>>(A);
    PR_CUDA_LAUNCH_CHECK(BWD ? "attn_bwd_kernel" : "attn_fwd_kernel");
    return PR_OK;
}

template <bool BWD>
static int dispatch_attn(AttnArgs& A, cudaStream_t stream) {
    const int cw4 = std::min(A.dh, 128) / 4;
    const int L = A.L;
#define PR_ATT_L(CW4)                                                         \
    do {                                                                      \
        if (L <= 12) return launch_attn<12, CW4, BWD>(A, stream);             \
        if (L <= 20) return launch_attn<20, CW4, BWD>(A, stream);             \
        if (L <= 32) return launch_attn<32, CW4, BWD>(A, stream);             \
        return launch_attn<64, CW4, BWD>(A, stream);                          \
    } while (0)
    switch (cw4) {
        case 4: PR_ATT_L(4);
        case 8: PR_ATT_L(8);
        case 16: PR_ATT_L(16);
        case 32: PR_ATT_L(32);
        default: break;
    }
#undef PR_ATT_L
    set_last_error("attention: head dim %d unsupported (need 16, 32, 64, 128 or a multiple of 128)", A.dh);
    return PR_ERR_UNSUPPORTED;
}

static int check_attn(const char* who, const float* q, const float* k, const float* v, long long ld, int B, int L, int h,
                      int dh, float p) {
    PR_CHECK_ARG(B > 0 && L > 0 && h > 0 && dh > 0, "%s: bad shape B=%d L=%d h=%d dh=%d", who, B, L, h, dh);
    PR_CHECK_ARG(L <= 64, "%s: L=%d > 64 unsupported", who, L);
    PR_CHECK_ARG(dh % 4 == 0 && (dh <= 128 || dh % 128 == 0), "%s: dh=%d must be <=128 (multiple of 4) or a multiple of 128", who, dh);
    PR_CHECK_ARG(ld % 4 == 0 && ld >= (long long)h * dh, "%s: ld=%lld must be a multiple of 4 and >= h*dh", who, ld);
    PR_CHECK_ARG(q && k && v && aligned16(q) && aligned16(k) && aligned16(v), "%s: q/k/v null or not 16-byte aligned", who);
    PR_CHECK_ARG(p >= 0.f && p < 1.f, "%s: dropout p outside [0,1)", who);
    return PR_OK;
}

}  // namespace pr

using namespace pr;

extern "C" int pr_sasrec_attn_fwd_f32(const float* q, const float* k, const float* v, int64_t ld,
                                      const int64_t* key_ids, int B, int L, int h, int dh, int causal, float p_drop,
                                      uint64_t seed, uint32_t rng_stream, float* ctx, float* probs,
                                      pr_stream_t stream_) {
    int rc = check_attn("pr_sasrec_attn_fwd_f32", q, k, v, ld, B, L, h, dh, p_drop);
    if (rc) return rc;
    PR_CHECK_ARG(ctx && probs && aligned16(ctx), "pr_sasrec_attn_fwd_f32: ctx/probs null or unaligned");
    AttnArgs A{};
    A.q = q; A.k = k; A.v = v; A.ld = ld; A.key_ids = (const long long*)key_ids;
    A.B = B; A.L = L; A.h = h; A.dh = dh; A.nc = (dh + 127) / 128; A.causal = causal;
    A.p_drop = p_drop; A.seed = seed; A.rng_stream = rng_stream;
    PR_SET_SEED_DEV(A);
    A.inv_sqrt = (float)(1.0 / sqrt((double)dh));
    A.ctx = ctx; A.probs = probs;
    return dispatch_attn<false>(A, (cudaStream_t)stream_);
}

extern "C" int pr_sasrec_attn_bwd_f32(const float* q, const float* k, const float* v, int64_t ld, const float* probs,
                                      const float* dctx, int B, int L, int h, int dh, int causal, float p_drop,
                                      uint64_t seed, uint32_t rng_stream, float* dq, float* dk, float* dv,
                                      int64_t ld_grad, pr_stream_t stream_) {
    int rc = check_attn("pr_sasrec_attn_bwd_f32", q, k, v, ld, B, L, h, dh, p_drop);
    if (rc) return rc;
    PR_CHECK_ARG(probs && dctx && dq && dk && dv, "pr_sasrec_attn_bwd_f32: null pointer");
    PR_CHECK_ARG(aligned16(dctx) && aligned16(dq) && aligned16(dk) && aligned16(dv), "pr_sasrec_attn_bwd_f32: unaligned pointer");
    PR_CHECK_ARG(ld_grad % 4 == 0 && ld_grad >= (int64_t)h * dh, "pr_sasrec_attn_bwd_f32: bad ld_grad");
    AttnArgs A{};
    A.q = q; A.k = k; A.v = v; A.ld = ld;
    A.B = B; A.L = L; A.h = h; A.dh = dh; A.nc = (dh + 127) / 128; A.causal = causal;
    A.p_drop = p_drop; A.seed = seed; A.rng_stream = rng_stream;
    PR_SET_SEED_DEV(A);
    A.inv_sqrt = (float)(1.0 / sqrt((double)dh));
    A.probs = const_cast<float*>(probs); A.dctx = dctx; A.dq = dq; A.dk = dk; A.dv = dv; A.ld_grad = ld_grad;
    return dispatch_attn<true>(A, (cudaStream_t)stream_);
}
