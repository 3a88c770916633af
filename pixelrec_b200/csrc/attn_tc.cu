// pixelrec_b200 -- K4+K6, tensor-core variant: the same warp-specialised, TMA-fed attention pipeline as attn.cu
// (REC/model/IDNet/sasrec.py:119-126 + REC/model/layers.py:590-612), but the two tiny per-head matrix products
// run on the tensor cores (mma.sync m16n8k8 TF32, fp32 accumulate) instead of FFMA.  attn.cu is issue-bound on
// FFMA (ncu: 34 % issue active, 9 warps/SM, 0.28 of the HBM roofline); here a (batch, head) item costs ~200
// tensor instructions instead of ~3400 FFMA, which moves the kernel onto the HBM roofline it belongs to.
//
// Used when TF32 is allowed for matrix products (the linear layers of the model already run TF32 -- what
// torch 1.10 did on Ampere); attn.cu remains the strict-fp32 path.  Dropout masks, masking semantics, probs
// layout and the saved tensors are identical to attn.cu, so forward/backward of the two paths interoperate.
//
// Fragment trick: the contraction index of a product is arbitrary as long as A and B agree.  For O = P V (and
// dQ = dS K) the key index is permuted (mma k = t <-> key 2t, k = t+4 <-> key 2t+1) so that the accumulator
// fragments of S = Q K^T are bit-for-bit the A fragments of the next product: P never leaves registers.
//
// Tiles are fed by 2-D TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B): one TMA operation moves an [L x 32-float] box
// (the first version issued one 512-byte cp.async.bulk per row and was bound by the TMA engine's per-operation
// rate: 1.8 TB/s at dh = 128 no matter how cheap the math was).  A [L x CW] tile is CW/32 such boxes; inside a box
// element (r, c) sits at r*128 B + (((c/4) ^ (r%8)) * 16 B) + (c%4)*4 B, which makes both fragment access patterns
// (row = g, col = k0 + t  and  row = 2t, col = n0 + g) bank-conflict-free without padding.
#include <cuda.h>

#include <algorithm>
#include <math.h>

#include "common.cuh"

namespace pr {

struct TcArgs {
    const float* q; const float* k; const float* v; long long ld;
    const long long* key_ids;
    int B, L, h, dh, nc, causal;
    float p_drop; unsigned long long seed; unsigned rng_stream;
    float inv_sqrt;
    float* ctx; float* probs;
    const float* dctx; float* dq; float* dk; float* dv; long long ld_grad;
#ifdef PR_SEED_DEV
    const unsigned long long* seed_dev;   // device-side seed offset (pr_set_seed_device)
#endif
};

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// word offset of element (r, c) inside a swizzled tile made of [ROWS x 32-float] boxes
__device__ __forceinline__ int sw_off(int r, int c, int box_words) {
    return (c >> 5) * box_words + r * 32 + (((((c >> 2) & 7) ^ (r & 7))) << 2) + (c & 3);
}
__device__ __forceinline__ void tma_box_2d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}

template <int MT, int NT>
struct TcCfg {
    static constexpr int ROWS = 8 * NT;                                  // smem rows of a box (keys; zero beyond L)
    static constexpr int LS = 16 * MT + 4;                               // row stride of the private L x L tiles (== 4 mod 16)
    static constexpr int PR = 8 * NT;                                     // rows of the private tiles
};

// One warp = one self-contained pipeline: it issues the 2-D TMA loads of its own (batch, head) items into private
// buffers, waits on its own mbarriers and computes.  (The first design fed 8 consumer warps from one producer warp
// through a shared in-order ring: ncu showed the consumers spinning on the ring ~60 % of the time and the producer
// blocked on full stages -- 0.28 of the HBM roofline no matter how cheap the math was.)
struct WarpBuf {
    float* buf[3];            // three [ROWS x CW] swizzled tiles
    uint64_t* bar[3];         // one mbarrier per tile
    unsigned phase[3];        // parity to wait for next
    // `leader`: the one thread of the pipeline that talks to the TMA engine (lane 0 of the warp, or of the pair's first warp)
    __device__ __forceinline__ void issue(int i, const CUtensorMap* tm, int col0, int row0, int L, int CW, int box_bytes,
                                          bool leader) {
        if (leader) {
            mbar_arrive_expect_tx(bar[i], (uint32_t)(CW * 4) * (uint32_t)L);
            for (int bx = 0; bx < CW / 32; ++bx)
                tma_box_2d(reinterpret_cast<unsigned char*>(buf[i]) + (size_t)bx * box_bytes, tm, col0 + 32 * bx, row0, bar[i]);
        }
    }
    __device__ __forceinline__ void wait(int i) {
        mbar_wait(bar[i], phase[i] & 1u);
        phase[i] ^= 1u;
    }
};

// all threads of one pipeline: a warp, or (PAIR) the two warps that share the pipeline's buffers -- named barrier 1 + pipe
template <bool PAIR>
__device__ __forceinline__ void pipe_sync(int pipe) {
    if (PAIR) asm volatile("bar.sync %0, 64;" ::"r"(pipe + 1) : "memory");
    else __syncwarp();
}

// carve per-warp buffers out of dynamic smem, zero them (rows >= L must read as 0), init the barriers
// (NW = pipelines per CTA, `warp` = index of this thread's pipeline: a warp, or a pair of warps sharing the buffers)
template <int NW>
__device__ __forceinline__ WarpBuf warp_setup(unsigned char* smem, int tile_bytes, size_t private_bytes_per_warp,
                                              unsigned char** priv, int warp) {
    const size_t per_warp = (size_t)3 * tile_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NW * per_warp + NW * private_bytes_per_warp);
    float4* z = reinterpret_cast<float4*>(smem);
    const size_t n16 = (NW * per_warp + NW * private_bytes_per_warp) / 16;
    for (size_t i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 3 * NW; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    fence_proxy_async();
    __syncthreads();
    WarpBuf w;
    for (int i = 0; i < 3; ++i) {
        w.buf[i] = reinterpret_cast<float*>(smem + warp * per_warp + (size_t)i * tile_bytes);
        w.bar[i] = &bars[3 * warp + i];
        w.phase[i] = 0;
    }
    *priv = smem + NW * per_warp + (size_t)warp * private_bytes_per_warp;
    return w;
}

// acc[mi][nj] += X[16mi + {g, g+8}][d] * Y[8nj + g][d]  over one CW-wide tile pair  ("row x row" product)
template <int MT, int NT>
__device__ __forceinline__ void tc_rowrow(const float* __restrict__ Xs, const float* __restrict__ Ys, int box_words, int CW,
                                          int g, int t, float (&acc)[MT][NT][4], int m0 = 0) {   // X rows start at m-tile m0
    // every row used here has (row & 7) == g, so the swizzle term is shared by all of them
#pragma unroll 4
    for (int ks = 0; ks < CW; ks += 8) {
        const int base = (ks >> 5) * box_words + t;
        const int ch0 = ((((ks >> 2) & 7) ^ g) << 2), ch1 = (((((ks >> 2) + 1) & 7) ^ g) << 2);
        uint32_t a[MT][4], b[NT][2];
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) {
            const float* r0 = Xs + base + (16 * (mi + m0) + g) * 32;
            // query rows >= 8*NT do not exist in the tile (and are >= L): alias them, their results are discarded
            const float* r1 = (16 * (mi + m0) + 8 < 8 * NT) ? r0 + 8 * 32 : r0;
            a[mi][0] = to_tf32(r0[ch0]);
            a[mi][1] = to_tf32(r1[ch0]);
            a[mi][2] = to_tf32(r0[ch1]);
            a[mi][3] = to_tf32(r1[ch1]);
        }
#pragma unroll
        for (int nj = 0; nj < NT; ++nj) {
            const float* r0 = Ys + base + (8 * nj + g) * 32;
            b[nj][0] = to_tf32(r0[ch0]);
            b[nj][1] = to_tf32(r0[ch1]);
        }
#pragma unroll
        for (int mi = 0; mi < MT; ++mi)
#pragma unroll
            for (int nj = 0; nj < NT; ++nj) mma_tf32(acc[mi][nj], a[mi][0], a[mi][1], a[mi][2], a[mi][3], b[nj][0], b[nj][1]);
    }
}

// out[16mi + {g,g+8}][d] = sum_j Afrag[mi][ks] (x) X[8ks + {2t, 2t+1}][d]   for every 8-wide column tile of X; rows < L
// are stored.  Afrag holds A fragments with the permuted contraction index (k = t <-> row 2t, k = t+4 <-> row 2t+1).
template <int MT, int KT>
__device__ __forceinline__ void tc_frag_times_tile(const uint32_t (&af)[MT][KT][4], const float* __restrict__ Xs, int box_words,
                                                   int CW, int L, int g, int t, float* __restrict__ out, long long out_ld,
                                                   int m0 = 0) {
    for (int nd = 0; nd < CW; nd += 8) {
        // column nd + g: 16-byte chunk ((nd>>2) + (g>>2)) & 7, word g & 3; rows 2t / 2t+1 (+8ks): (row & 7) = 2t / 2t+1
        const int cbase = (nd >> 5) * box_words + (g & 3);
        const int chunk = ((nd >> 2) + (g >> 2)) & 7;
        const int o0 = cbase + (2 * t) * 32 + ((chunk ^ (2 * t)) << 2);
        const int o1 = cbase + (2 * t + 1) * 32 + ((chunk ^ (2 * t + 1)) << 2);
        float acc[MT][4];
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) acc[mi][0] = acc[mi][1] = acc[mi][2] = acc[mi][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KT; ++ks) {
            const uint32_t b0 = to_tf32(Xs[o0 + ks * 256]);       // 8 rows further: same (row & 7), +8*32 words
            const uint32_t b1 = to_tf32(Xs[o1 + ks * 256]);
#pragma unroll
            for (int mi = 0; mi < MT; ++mi) mma_tf32(acc[mi], af[mi][ks][0], af[mi][ks][1], af[mi][ks][2], af[mi][ks][3], b0, b1);
        }
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) {
            const int r0 = 16 * (mi + m0) + g, r1 = r0 + 8;
            if (r0 < L) *reinterpret_cast<float2*>(out + (long long)r0 * out_ld + nd + 2 * t) = make_float2(acc[mi][0], acc[mi][1]);
            if (r1 < L) *reinterpret_cast<float2*>(out + (long long)r1 * out_ld + nd + 2 * t) = make_float2(acc[mi][2], acc[mi][3]);
        }
    }
}

// A fragments of M^T from a private smem matrix M[r][c] (row stride LS): out rows = c, contraction over r (permuted)
template <int MT, int KT>
__device__ __forceinline__ void tc_load_transposed(const float* __restrict__ M, int LS, int g, int t, uint32_t (&af)[MT][KT][4],
                                                   int m0 = 0) {
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
        for (int ks = 0; ks < KT; ++ks) {
            const float* r0 = M + (8 * ks + 2 * t) * LS + 16 * (mi + m0) + g;
            const float* r1 = r0 + LS;
            af[mi][ks][0] = to_tf32(r0[0]);      // (row c = 16mi+g,   k = t   <-> r = 8ks+2t)
            af[mi][ks][1] = to_tf32(r0[8]);      // (row c = 16mi+g+8, k = t)
            af[mi][ks][2] = to_tf32(r1[0]);      // (row c = 16mi+g,   k = t+4 <-> r = 8ks+2t+1)
            af[mi][ks][3] = to_tf32(r1[8]);
        }
}

// keep bits for score row i, columns j = 8*nj + 2t + e (e = 0,1): same Philox layout as attn.cu attn_keep
template <int NT>
__device__ __forceinline__ void tc_keep(const Philox& ph, unsigned stream, long long item, int L, int i, int t, unsigned thr,
                                        unsigned (&bits)[2]) {
    static_assert(NT <= 8, "one Philox draw covers 8 column tiles");
    bits[0] = keep_bits8(ph((unsigned long long)((item * L + i) * 8 + 2 * t), stream), thr);
    bits[1] = keep_bits8(ph((unsigned long long)((item * L + i) * 8 + 2 * t + 1), stream), thr);
}

// =============================================================================================== forward
// PAIR: two warps share one pipeline (one set of tiles, one set of mbarriers) and each owns one 16-row query tile of the
// item -- every product of the forward is row-local, so the pair only meets at the buffer hand-overs (pipe_sync).  Twice
// the warps per SM at the same shared-memory footprint; the kernel is latency-bound at 6 warps/SM (profiles/r01h_attention_ncu.md).
template <int MT, int NT, int NW, bool PAIR>
__global__ void __launch_bounds__(NW * (PAIR ? 64 : 32), 1) attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                 const __grid_constant__ CUtensorMap tmK,
                                                                 const __grid_constant__ CUtensorMap tmV, const TcArgs A) {
    using C = TcCfg<MT, NT>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzled boxes need 1 KiB alignment
    const int CW = min(A.dh, 128);
    constexpr int box_words = C::ROWS * 32, box_bytes = box_words * 4;
    static_assert(!PAIR || MT == 2, "a pair splits exactly two query tiles");
    constexpr int MW = PAIR ? MT / 2 : MT;                    // query tiles per warp
    const int lane = threadIdx.x & 31;
    const int warp = PAIR ? (threadIdx.x >> 6) : (threadIdx.x >> 5);          // pipeline index inside the CTA
    const int m0 = PAIR ? ((threadIdx.x >> 5) & 1) : 0;                       // first query tile of this warp
    const bool leader = (lane == 0) && (m0 == 0);
    unsigned char* priv;
    WarpBuf wb = warp_setup<NW>(smem, (CW / 32) * box_bytes, 0, &priv, warp);
    const int L = A.L, nc = A.nc;
    const long long n_items = (long long)A.B * A.h;
    // contiguous items per warp: the heads of a sequence and consecutive sequences stream through the same SM
    const long long n_warps = (long long)gridDim.x * NW;
    const long long per_warp = (n_items + n_warps - 1) / n_warps;
    const long long item_lo = ((long long)blockIdx.x * NW + warp) * per_warp;
    const long long item_hi = min(n_items, item_lo + per_warp);
    if (item_lo >= item_hi) return;
    const int g = lane >> 2, t = lane & 3;
    const Philox ph(PR_SEED(A));
    const unsigned thr = drop_threshold(A.p_drop);
    const float inv_keep = 1.0f / (1.0f - A.p_drop);
    auto coords = [&](long long item, int& row0, int& col0) {
        const long long b = item / A.h;
        row0 = (int)(b * L);
        col0 = (int)(item - b * A.h) * A.dh;
    };
    {   // prologue: first item's Q, K (chunk 0) and V (chunk 0)
        int r0, c0;
        coords(item_lo, r0, c0);
        wb.issue(0, &tmQ, c0, r0, L, CW, box_bytes, leader);
        wb.issue(1, &tmK, c0, r0, L, CW, box_bytes, leader);
        wb.issue(2, &tmV, c0, r0, L, CW, box_bytes, leader);
    }
    for (long long item = item_lo; item < item_hi; ++item) {
        const long long bb = item / A.h;
        const int hd = (int)(item - bb * A.h);
        int row0, col0, nrow0 = 0, ncol0 = 0;
        coords(item, row0, col0);
        const bool has_next = item + 1 < item_hi;
        if (has_next) coords(item + 1, nrow0, ncol0);
        float acc[MW][NT][4];
#pragma unroll
        for (int mi = 0; mi < MW; ++mi)
#pragma unroll
            for (int nj = 0; nj < NT; ++nj) acc[mi][nj][0] = acc[mi][nj][1] = acc[mi][nj][2] = acc[mi][nj][3] = 0.f;
        for (int c = 0; c < nc; ++c) {
            wb.wait(0);
            wb.wait(1);
            tc_rowrow<MW, NT>(wb.buf[0], wb.buf[1], box_words, CW, g, t, acc, m0);
            pipe_sync<PAIR>(warp);                           // everyone is done reading before the buffers are refilled
            if (c + 1 < nc) {
                wb.issue(0, &tmQ, col0 + (c + 1) * CW, row0, L, CW, box_bytes, leader);
                wb.issue(1, &tmK, col0 + (c + 1) * CW, row0, L, CW, box_bytes, leader);
            } else if (has_next) {                           // next item's Q, K load while this item does softmax + P V
                wb.issue(0, &tmQ, ncol0, nrow0, L, CW, box_bytes, leader);
                wb.issue(1, &tmK, ncol0, nrow0, L, CW, box_bytes, leader);
            }
        }
        // ---- mask + softmax on the accumulator fragments: thread holds rows {16mi+g, +8}, cols 8nj + 2t + {0,1}
        bool kvalid[NT][2];
#pragma unroll
        for (int nj = 0; nj < NT; ++nj)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = 8 * nj + 2 * t + e;
                kvalid[nj][e] = (j < L) && (A.key_ids == nullptr || A.key_ids[bb * L + j] != 0);
            }
        uint32_t pf[MW][NT][4];          // dropped P as A fragments of the next product
#pragma unroll
        for (int mi = 0; mi < MW; ++mi)
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                const int i = 16 * (mi + m0) + g + 8 * hrow;
                float sc[NT][2];
                float mx = -INFINITY;
#pragma unroll
                for (int nj = 0; nj < NT; ++nj)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = 8 * nj + 2 * t + e;
                        const bool ok = kvalid[nj][e] && (!A.causal || j <= i);
                        float x = acc[mi][nj][2 * hrow + e] * A.inv_sqrt + (ok ? 0.0f : -1e9f);
                        if (j >= L) x = -INFINITY;
                        sc[nj][e] = x;
                        mx = fmaxf(mx, x);
                    }
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                float sum = 0.f;
#pragma unroll
                for (int nj = 0; nj < NT; ++nj)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float ex = (8 * nj + 2 * t + e < L) ? expf(sc[nj][e] - mx) : 0.f;
                        sc[nj][e] = ex;
                        sum += ex;
                    }
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                const float inv = 1.0f / sum;
                unsigned kb[2] = {0xffu, 0xffu};
                if (A.p_drop > 0.f) tc_keep<NT>(ph, A.rng_stream, item, L, i, t, thr, kb);
#pragma unroll
                for (int nj = 0; nj < NT; ++nj) {
                    const int j = 8 * nj + 2 * t;
                    const float p0 = sc[nj][0] * inv, p1 = sc[nj][1] * inv;
                    if (i < L) {
                        float* pr = A.probs + (item * L + i) * L + j;
                        if (j < L) pr[0] = p0;
                        if (j + 1 < L) pr[1] = p1;
                    }
                    float d0 = p0, d1 = p1;
                    if (A.p_drop > 0.f) {
                        d0 = ((kb[0] >> nj) & 1u) ? p0 * inv_keep : 0.f;
                        d1 = ((kb[1] >> nj) & 1u) ? p1 * inv_keep : 0.f;
                    }
                    // accumulator (c0,c1 | c2,c3) -> A fragment (a0,a2 | a1,a3) of the key-permuted product
                    pf[mi][nj][hrow] = to_tf32(d0);
                    pf[mi][nj][2 + hrow] = to_tf32(d1);
                }
            }
        // ---- O = drop(P) V, chunk by chunk
        for (int c = 0; c < nc; ++c) {
            wb.wait(2);
            float* out = A.ctx + bb * L * (long long)(A.h * A.dh) + (long long)hd * A.dh + c * CW;
            tc_frag_times_tile<MW, NT>(pf, wb.buf[2], box_words, CW, L, g, t, out, (long long)A.h * A.dh, m0);
            pipe_sync<PAIR>(warp);
            if (c + 1 < nc) wb.issue(2, &tmV, col0 + (c + 1) * CW, row0, L, CW, box_bytes, leader);
            else if (has_next) wb.issue(2, &tmV, ncol0, nrow0, L, CW, box_bytes, leader);
        }
    }
}

// =============================================================================================== backward
//   buffers: 0 = dO (kept for both phases when nc == 1), 1 = V then K, 2 = Q
//   PAIR: as in the forward; dP/dS/dQ are split by query tile, dV/dK by key tile (they contract over ALL queries, which
//   the pair exchanges through the pipeline's private Pd / dS tiles -- one extra pipe_sync after they are written)
template <int MT, int NT, int NW, bool PAIR>
__global__ void __launch_bounds__(NW * (PAIR ? 64 : 32), 1) attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                 const __grid_constant__ CUtensorMap tmK,
                                                                 const __grid_constant__ CUtensorMap tmV,
                                                                 const __grid_constant__ CUtensorMap tmDO, const TcArgs A) {
    using C = TcCfg<MT, NT>;
    constexpr int LS = C::LS, PR = C::PR;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int CW = min(A.dh, 128);
    constexpr int box_words = C::ROWS * 32, box_bytes = box_words * 4;
    static_assert(!PAIR || MT == 2, "a pair splits exactly two query / key tiles");
    constexpr int MW = PAIR ? MT / 2 : MT;
    const int lane = threadIdx.x & 31;
    const int warp = PAIR ? (threadIdx.x >> 6) : (threadIdx.x >> 5);
    const int m0 = PAIR ? ((threadIdx.x >> 5) & 1) : 0;
    const bool leader = (lane == 0) && (m0 == 0);
    unsigned char* priv;
    WarpBuf wb = warp_setup<NW>(smem, (CW / 32) * box_bytes, (size_t)2 * PR * LS * 4, &priv, warp);
    const int L = A.L, nc = A.nc;
    const long long n_items = (long long)A.B * A.h;
    const long long n_warps = (long long)gridDim.x * NW;
    const long long per_warp = (n_items + n_warps - 1) / n_warps;
    const long long item_lo = ((long long)blockIdx.x * NW + warp) * per_warp;
    const long long item_hi = min(n_items, item_lo + per_warp);
    if (item_lo >= item_hi) return;
    float* Pd_s = reinterpret_cast<float*>(priv);             // Pd_s[i][j]   (i < 8NT rows, j < 16MT cols)
    float* dS_s = Pd_s + PR * LS;
    const int g = lane >> 2, t = lane & 3;
    const Philox ph(PR_SEED(A));
    const unsigned thr = drop_threshold(A.p_drop);
    const float inv_keep = 1.0f / (1.0f - A.p_drop);
    auto coords = [&](long long item, int& row0, int& col0) {
        const long long b = item / A.h;
        row0 = (int)(b * L);
        col0 = (int)(item - b * A.h) * A.dh;
    };
    {
        int r0, c0;
        coords(item_lo, r0, c0);
        wb.issue(0, &tmDO, c0, r0, L, CW, box_bytes, leader);
        wb.issue(1, &tmV, c0, r0, L, CW, box_bytes, leader);
        wb.issue(2, &tmQ, c0, r0, L, CW, box_bytes, leader);
    }
    for (long long item = item_lo; item < item_hi; ++item) {
        const long long bb = item / A.h;
        const int hd = (int)(item - bb * A.h);
        int row0, col0, nrow0 = 0, ncol0 = 0;
        coords(item, row0, col0);
        const bool has_next = item + 1 < item_hi;
        if (has_next) coords(item + 1, nrow0, ncol0);
        float acc[MW][NT][4];
#pragma unroll
        for (int mi = 0; mi < MW; ++mi)
#pragma unroll
            for (int nj = 0; nj < NT; ++nj) acc[mi][nj][0] = acc[mi][nj][1] = acc[mi][nj][2] = acc[mi][nj][3] = 0.f;
        for (int c = 0; c < nc; ++c) {              // dPd = dO V^T
            wb.wait(0);
            wb.wait(1);
            tc_rowrow<MW, NT>(wb.buf[0], wb.buf[1], box_words, CW, g, t, acc, m0);
            pipe_sync<PAIR>(warp);
            if (c + 1 < nc) {
                wb.issue(0, &tmDO, col0 + (c + 1) * CW, row0, L, CW, box_bytes, leader);
                wb.issue(1, &tmV, col0 + (c + 1) * CW, row0, L, CW, box_bytes, leader);
            } else {
                if (nc > 1) wb.issue(0, &tmDO, col0, row0, L, CW, box_bytes, leader);   // chunk 0 of dO again for phase 2
                wb.issue(1, &tmK, col0, row0, L, CW, box_bytes, leader);                 // K chunk 0 replaces V
            }
        }
        uint32_t dsf[MW][NT][4];                    // dS as A fragments (key-permuted) for dQ = dS K
#pragma unroll
        for (int mi = 0; mi < MW; ++mi)
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                const int i = 16 * (mi + m0) + g + 8 * hrow;
                unsigned kb[2] = {0xffu, 0xffu};
                if (A.p_drop > 0.f) tc_keep<NT>(ph, A.rng_stream, item, L, i, t, thr, kb);
                float p[NT][2], dp[NT][2];
                float rd = 0.f;
#pragma unroll
                for (int nj = 0; nj < NT; ++nj)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = 8 * nj + 2 * t + e;
                        const bool in = (i < L) && (j < L);
                        p[nj][e] = in ? A.probs[(item * L + i) * L + j] : 0.f;
                        float d = in ? acc[mi][nj][2 * hrow + e] : 0.f;
                        if (A.p_drop > 0.f) d = ((kb[e] >> nj) & 1u) ? d * inv_keep : 0.f;
                        dp[nj][e] = d;
                        rd = fmaf(d, p[nj][e], rd);
                    }
                rd += __shfl_xor_sync(0xffffffffu, rd, 1);
                rd += __shfl_xor_sync(0xffffffffu, rd, 2);
#pragma unroll
                for (int nj = 0; nj < NT; ++nj) {
                    float ds[2], pd[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        ds[e] = p[nj][e] * (dp[nj][e] - rd) * A.inv_sqrt;
                        pd[e] = p[nj][e];
                        if (A.p_drop > 0.f) pd[e] = ((kb[e] >> nj) & 1u) ? pd[e] * inv_keep : 0.f;
                    }
                    dsf[mi][nj][hrow] = to_tf32(ds[0]);
                    dsf[mi][nj][2 + hrow] = to_tf32(ds[1]);
                    if (i < PR) {                   // rows beyond 8NT are never read back (they are >= L anyway)
                        const int j = 8 * nj + 2 * t;
                        *reinterpret_cast<float2*>(Pd_s + i * LS + j) = make_float2(pd[0], pd[1]);
                        *reinterpret_cast<float2*>(dS_s + i * LS + j) = make_float2(ds[0], ds[1]);
                    }
                }
            }
        pipe_sync<PAIR>(warp);                                // all query rows of Pd / dS are in the private tiles
        uint32_t pT[MW][NT][4], dsT[MW][NT][4];
        tc_load_transposed<MW, NT>(Pd_s, LS, g, t, pT, m0);   // A = Pd^T  (rows = keys j, contraction over queries i)
        tc_load_transposed<MW, NT>(dS_s, LS, g, t, dsT, m0);  // A = dS^T
        pipe_sync<PAIR>(warp);                                // private tiles may be overwritten by the next item
        for (int c = 0; c < nc; ++c) {
            const long long go = bb * L * A.ld_grad + (long long)hd * A.dh + c * CW;
            const bool last = (c + 1 == nc);
            if (nc > 1 || c > 0) wb.wait(0);                   // (nc == 1: dO of phase 1 is still resident)
            tc_frag_times_tile<MW, NT>(pT, wb.buf[0], box_words, CW, L, g, t, A.dv + go, A.ld_grad, m0);     // dV = Pd^T dO
            pipe_sync<PAIR>(warp);
            if (!last) wb.issue(0, &tmDO, col0 + (c + 1) * CW, row0, L, CW, box_bytes, leader);
            else if (has_next) wb.issue(0, &tmDO, ncol0, nrow0, L, CW, box_bytes, leader);
            wb.wait(1);
            tc_frag_times_tile<MW, NT>(dsf, wb.buf[1], box_words, CW, L, g, t, A.dq + go, A.ld_grad, m0);    // dQ = dS K
            pipe_sync<PAIR>(warp);
            if (!last) wb.issue(1, &tmK, col0 + (c + 1) * CW, row0, L, CW, box_bytes, leader);
            else if (has_next) wb.issue(1, &tmV, ncol0, nrow0, L, CW, box_bytes, leader);
            wb.wait(2);
            tc_frag_times_tile<MW, NT>(dsT, wb.buf[2], box_words, CW, L, g, t, A.dk + go, A.ld_grad, m0);    // dK = dS^T Q
            pipe_sync<PAIR>(warp);
            if (!last) wb.issue(2, &tmQ, col0 + (c + 1) * CW, row0, L, CW, box_bytes, leader);
            else if (has_next) wb.issue(2, &tmQ, ncol0, nrow0, L, CW, box_bytes, leader);
        }
    }
}

// ----------------------------------------------------------------------------------------------- host
typedef CUresult (*TcEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TcEncodeFn tc_encode_fn() {
    static TcEncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (TcEncodeFn)p;
    }
    return fn;
}
// [rows, cols] fp32 view with row pitch ld floats -> boxes of [L rows x 32 floats], 128-byte swizzle
static int tc_make_map(CUtensorMap* tm, const float* base, long long rows, long long cols, long long ld, int L) {
    TcEncodeFn fn = tc_encode_fn();
    if (!fn) {
        set_last_error("cuTensorMapEncodeTiled entry point not available");
        return PR_ERR_UNSUPPORTED;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32u, (cuuint32_t)L};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return PR_ERR_INVALID_ARGUMENT;
    }
    return PR_OK;
}

template <int MT, int NT, int NW, bool BWD, bool PAIR>
static int launch_tc_nw(TcArgs& A, cudaStream_t stream) {
    using C = TcCfg<MT, NT>;
    const int CW = std::min(A.dh, 128);
    const size_t tile = (size_t)(CW / 32) * C::ROWS * 128;
    const size_t priv = BWD ? (size_t)2 * C::PR * C::LS * 4 : 0;
    const size_t smem = (size_t)NW * (3 * tile + priv) + (size_t)3 * NW * 8 + 1024;
    const long long rows = (long long)A.B * A.L, cols = (long long)A.h * A.dh;
    CUtensorMap tmQ, tmK, tmV, tmDO;
    int rc;
    if ((rc = tc_make_map(&tmQ, A.q, rows, cols, A.ld, A.L))) return rc;
    if ((rc = tc_make_map(&tmK, A.k, rows, cols, A.ld, A.L))) return rc;
    if ((rc = tc_make_map(&tmV, A.v, rows, cols, A.ld, A.L))) return rc;
    const long long n_items = (long long)A.B * A.h;
    const int grid = (int)std::max<long long>(1, std::min<long long>((n_items + NW - 1) / NW, sm_count()));
    if (BWD) {
        if ((rc = tc_make_map(&tmDO, A.dctx, rows, cols, cols, A.L))) return rc;
        PR_CUDA_CALL(cudaFuncSetAttribute(attn_tc_bwd_kernel<MT, NT, NW, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attn_tc_bwd_kernel<MT, NT, NW, PAIR><<<grid, NW * (PAIR ? 64 : 32), smem, stream>>>(tmQ, tmK, tmV, tmDO, A);
    } else {
        PR_CUDA_CALL(cudaFuncSetAttribute(attn_tc_fwd_kernel<MT, NT, NW, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attn_tc_fwd_kernel<MT, NT, NW, PAIR><<<grid, NW * (PAIR ? 64 : 32), smem, stream>>>(tmQ, tmK, tmV, A);
    }
    PR_CUDA_LAUNCH_CHECK(BWD ? "attn_tc_bwd_kernel" : "attn_tc_fwd_kernel");
    return PR_OK;
}

// warps per CTA: as many private pipelines as 220 KiB of shared memory hold (3 tiles [+ 2 private L x L tiles] per warp)
template <int MT, int NT, bool BWD>
static int launch_tc(TcArgs& A, cudaStream_t stream) {
    using C = TcCfg<MT, NT>;
    const int CW = std::min(A.dh, 128);
    const size_t per_warp = (size_t)3 * (CW / 32) * C::ROWS * 128 + (BWD ? (size_t)2 * C::PR * C::LS * 4 : 0) + 24;
    const int fit = (int)((220 * 1024 - 1024) / per_warp);
    PR_CHECK_ARG(fit >= 2, "attention(tf32): L=%d dh=%d does not fit shared memory", A.L, A.dh);
    constexpr bool CAN_PAIR = (MT == 2);
    if (CAN_PAIR && (tune() & PR_TUNE_ATTN_PAIR)) {          // two warps per pipeline (named barriers 1..NW: NW <= 8)
        if (fit >= 8) return launch_tc_nw<MT, NT, 8, BWD, CAN_PAIR>(A, stream);
        if (fit >= 6) return launch_tc_nw<MT, NT, 6, BWD, CAN_PAIR>(A, stream);
        if (fit >= 5) return launch_tc_nw<MT, NT, 5, BWD, CAN_PAIR>(A, stream);
        if (fit >= 4) return launch_tc_nw<MT, NT, 4, BWD, CAN_PAIR>(A, stream);
        return launch_tc_nw<MT, NT, 2, BWD, CAN_PAIR>(A, stream);
    }
    if (fit >= 8) return launch_tc_nw<MT, NT, 8, BWD, false>(A, stream);
    if (fit >= 6) return launch_tc_nw<MT, NT, 6, BWD, false>(A, stream);
    if (fit >= 5) return launch_tc_nw<MT, NT, 5, BWD, false>(A, stream);
    if (fit >= 4) return launch_tc_nw<MT, NT, 4, BWD, false>(A, stream);
    return launch_tc_nw<MT, NT, 2, BWD, false>(A, stream);
}

template <bool BWD>
static int dispatch_tc(TcArgs& A, cudaStream_t stream) {
    if (A.L <= 16) return launch_tc<1, 2, BWD>(A, stream);
    if (A.L <= 24) return launch_tc<2, 3, BWD>(A, stream);
    if (A.L <= 32) return launch_tc<2, 4, BWD>(A, stream);
    set_last_error("attention(tf32): L=%d > 32 unsupported (use the fp32 kernel)", A.L);
    return PR_ERR_UNSUPPORTED;
}

static int check_tc(const char* who, const float* q, const float* k, const float* v, long long ld, int B, int L, int h, int dh,
                    float p) {
    PR_CHECK_ARG(B > 0 && L > 0 && h > 0 && dh > 0, "%s: bad shape B=%d L=%d h=%d dh=%d", who, B, L, h, dh);
    PR_CHECK_ARG(L <= 32, "%s: L=%d > 32 unsupported by the tensor-core path", who, L);
    PR_CHECK_ARG(dh % 32 == 0 && (dh <= 128 || dh % 128 == 0), "%s: dh=%d must be 32, 64, 96, 128 or a multiple of 128", who, dh);
    PR_CHECK_ARG(ld % 4 == 0 && ld >= (long long)h * dh, "%s: ld=%lld must be a multiple of 4 and >= h*dh", who, ld);
    PR_CHECK_ARG(q && k && v && aligned16(q) && aligned16(k) && aligned16(v), "%s: q/k/v null or not 16-byte aligned", who);
    PR_CHECK_ARG(p >= 0.f && p < 1.f, "%s: dropout p outside [0,1)", who);
    return PR_OK;
}

}  // namespace pr

using namespace pr;

extern "C" int pr_sasrec_attn_fwd_tf32(const float* q, const float* k, const float* v, int64_t ld, const int64_t* key_ids,
                                       int B, int L, int h, int dh, int causal, float p_drop, uint64_t seed,
                                       uint32_t rng_stream, float* ctx, float* probs, pr_stream_t stream_) {
    int rc = check_tc("pr_sasrec_attn_fwd_tf32", q, k, v, ld, B, L, h, dh, p_drop);
    if (rc) return rc;
    PR_CHECK_ARG(ctx && probs && aligned16(ctx), "pr_sasrec_attn_fwd_tf32: ctx/probs null or unaligned");
    TcArgs A{};
    A.q = q; A.k = k; A.v = v; A.ld = ld; A.key_ids = (const long long*)key_ids;
    A.B = B; A.L = L; A.h = h; A.dh = dh; A.nc = (dh + 127) / 128; A.causal = causal;
    A.p_drop = p_drop; A.seed = seed; A.rng_stream = rng_stream;
    PR_SET_SEED_DEV(A);
    A.inv_sqrt = (float)(1.0 / sqrt((double)dh));
    A.ctx = ctx; A.probs = probs;
    return dispatch_tc<false>(A, (cudaStream_t)stream_);
}

extern "C" int pr_sasrec_attn_bwd_tf32(const float* q, const float* k, const float* v, int64_t ld, const float* probs,
                                       const float* dctx, int B, int L, int h, int dh, int causal, float p_drop,
                                       uint64_t seed, uint32_t rng_stream, float* dq, float* dk, float* dv, int64_t ld_grad,
                                       pr_stream_t stream_) {
    int rc = check_tc("pr_sasrec_attn_bwd_tf32", q, k, v, ld, B, L, h, dh, p_drop);
    if (rc) return rc;
    PR_CHECK_ARG(probs && dctx && dq && dk && dv, "pr_sasrec_attn_bwd_tf32: null pointer");
    PR_CHECK_ARG(aligned16(dctx) && aligned16(dq) && aligned16(dk) && aligned16(dv), "pr_sasrec_attn_bwd_tf32: unaligned pointer");
    PR_CHECK_ARG(ld_grad % 4 == 0 && ld_grad >= (int64_t)h * dh, "pr_sasrec_attn_bwd_tf32: bad ld_grad");
    TcArgs A{};
    A.q = q; A.k = k; A.v = v; A.ld = ld;
    A.B = B; A.L = L; A.h = h; A.dh = dh; A.nc = (dh + 127) / 128; A.causal = causal;
    A.p_drop = p_drop; A.seed = seed; A.rng_stream = rng_stream;
    PR_SET_SEED_DEV(A);
    A.inv_sqrt = (float)(1.0 / sqrt((double)dh));
    A.probs = const_cast<float*>(probs); A.dctx = dctx; A.dq = dq; A.dk = dk; A.dv = dv; A.ld_grad = ld_grad;
    return dispatch_tc<true>(A, (cudaStream_t)stream_);
}
