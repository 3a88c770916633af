// pixelrec_b200 -- K5: the encoder's nn.Linear layers and their backward on tcgen05 tensor cores.
//   forward      y  = x W^T + b            REC/model/layers.py:586-588 (query/key/value), :613 (dense), :666, :669 (feed-forward)
//   input grad   dx = dy W                 (autograd of the same lines)
//   weight grad  dW = dy^T x               (autograd of the same lines; contraction over the B*L rows, split-K)
// One persistent, warp-specialised kernel computes   out[M,N] = epilogue( A[M,K] . B[N,K]^T )   for all three:
//   * operands stay fp32 in HBM (the tensor core reads them as TF32); each may be K-major ([rows, K] row-major: activations in
//     the forward, nn.Linear weights) or MN-major ([K, rows] row-major: W in the input-gradient GEMM, dy and x in the
//     weight-gradient GEMM) -- no transposed copy is ever made.  K-major tiles use the 128-byte swizzle, MN-major TF32 tiles
//     the 32-byte-atom 128-byte swizzle (the only layout tcgen05 accepts for them), both written by TMA.
//   * CTA PAIRS (cta_group::2): two CTAs of a cluster hold the two 128-row halves of a 256 x 256 tile.  Each loads its half of
//     A and HALF of B per k-block (32 KiB instead of 48 KiB for a lone CTA): per-SM operand ingest is what bounded the
//     single-CTA pipeline (ncu, profiles/r02a_score_v2_ncu.md: tensor pipe 55-62 % active with L2 at 54 %).  The leader CTA's
//     single MMA thread issues tcgen05.mma.cta_group::2 (UMMA 256x256x8); tcgen05.commit multicasts "stage free" and
//     "accumulator ready" to both CTAs; both epilogues arrive on the leader's "TMEM buffer drained" barrier.
//   * persistent: grid = one CTA (pair) per SM (pair); work items (m-tile, n-tile, k-split) are walked with a static stride;
//     the smem ring and the two TMEM accumulator buffers run across item boundaries, so the epilogue of item i overlaps the
//     MMAs of item i+1.
//   * epilogue (8 warps; thread == output row, 32 columns at a time): tcgen05.ld -> bias / residual add / activation /
//     activation-backward -> swizzled shared-memory staging -> TMA store (coalesced 128-byte rows; out-of-range rows and
//     columns clipped by the tensor map).  Same-shaped epilogue inputs (residual, pre-activation) arrive by TMA too.
//     Optional per-warp column sums of the output (bias gradient) go to a partials matrix for pr_colsum_f32.
#include <cuda.h>

#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "act.cuh"
#include "common.cuh"
#include "tc_ptx.cuh"

namespace pr {

constexpr int GM_BM = 128;                 // rows per CTA (UMMA M = 128 * CG)
constexpr int GM_BN = 256;                 // columns per tile (UMMA N)
constexpr int GM_BK = 32;                  // floats per k-block: one 128-byte swizzle row
constexpr int GM_A_BYTES = GM_BM * GM_BK * 4;                  // 16 KiB
constexpr int GM_EPI_WARPS = 8;
constexpr int GM_THREADS = 64 + 32 * GM_EPI_WARPS;
constexpr int GM_CHUNK_BYTES = 32 * 32 * 4;                    // one epilogue chunk: 32 rows x 32 columns
constexpr int GM_TMEM_COLS = 512;

constexpr int GM_MAX_STAGES = 8;
constexpr int GM_BAR_BYTES = 512;
constexpr int GM_SMEM_LIMIT = 227 * 1024;
template <int CG> struct GemmCfg {
    static constexpr int B_ROWS = GM_BN / CG;                  // rows of B this CTA loads
    static constexpr int B_BYTES = B_ROWS * GM_BK * 4;
    static constexpr int STAGE_BYTES = GM_A_BYTES + B_BYTES;   // 48 KiB (CG 1) / 32 KiB (CG 2)
};
// shared memory: [ring: stages x (A | B)] [out staging: obuf x 8 warps x 4 KiB] [aux staging: 8 x 4 KiB, if used] [barriers]
static inline size_t gm_smem_bytes(int stage_bytes, int stages, int obuf, int aux) {
    return (size_t)stages * stage_bytes + (size_t)(obuf + (aux ? 1 : 0)) * GM_EPI_WARPS * GM_CHUNK_BYTES + GM_BAR_BYTES + 1024;
}

enum { GM_EPI_STORE = 0, GM_EPI_ADD = 1, GM_EPI_ACT = 2, GM_EPI_ACT_BWD = 3 };

struct GemmArgs {
    int m_tiles, n_tiles, splits;          // work items = m_tiles * n_tiles * splits (m-tile = 128*CG rows, n-tile = 256 columns)
    int kb_total, kb_per_split;            // k-blocks of 32
    int a_mn, b_mn;                        // operand majors: 0 = K-major, 1 = MN-major
    int epi, act;
    long long M, N;
    const float* bias;                     // [N] or null
    float* colsum;                         // [n_part_rows, N] or null: per-warp column sums of the stored output
    int has_out2;
    int stages, obuf;                      // ring depth; out-staging buffers per epilogue warp (1 or 2)
    int debias;                            // scale the accumulator by 1.00071 (TF32 truncation shrink): see pr_gemm_tf32 in the header
    float p_drop;                          // PR_GEMM_ADD only: out = drop(acc + bias) + aux with ln.cu's keep-bit layout (pr_gemm_tf32_drop)
    unsigned long long seed;
    unsigned rng_stream;
#ifdef PR_SEED_DEV
    const unsigned long long* seed_dev;
#endif
    int debug;                             // PR_GEMM_DEBUG (timing experiments only, results are garbage): 1 = no output stores,
                                           // 2 = no operand loads (MMAs on whatever the ring holds)
};

// ---- PTX forms of the CTA-pair pipeline (cute/arch/copy_sm100_tma.hpp, mma_sm100_umma.hpp, cutlass/arch/barrier.h)
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
    return r;
}
// tile -> own shared memory; the bytes are credited to an mbarrier given by its shared::cluster address (the pair leader's)
__device__ __forceinline__ void tma_load_2d_cg2(void* smem_dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar_cluster_addr) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(smem_u32(smem_src)),
                 "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, const void* smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tm),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
        : "memory");
}
// one arrive on the same-offset mbarrier of both CTAs of the pair once this thread's earlier MMAs have completed
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

// shared-memory matrix descriptors (cute::UMMA::SmemDescriptor)
//   K-major, 128-byte swizzle: rows 128 B apart, 8-row groups 1024 B apart (SBO); LBO unused
//   MN-major TF32, 128-byte swizzle with 32-byte atoms (layout type 1): a [32 k-rows x 128 B] block per 32 MN elements;
//   LBO = distance between those blocks (4096 B), SBO = distance between groups of 4 k-rows (512 B)
__device__ __forceinline__ uint64_t gm_desc(uint32_t smem_addr, int mn_major) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 46;
    if (mn_major) {
        d |= (uint64_t)(4096 >> 4) << 16;
        d |= (uint64_t)(512 >> 4) << 32;
        d |= (uint64_t)1 << 61;
    } else {
        d |= (uint64_t)1 << 16;
        d |= (uint64_t)(1024 >> 4) << 32;
        d |= (uint64_t)2 << 61;
    }
    return d;
}
__host__ __device__ constexpr uint32_t gm_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a_mn & 1) << 15) | ((uint32_t)(b_mn & 1) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct GmItem { int m_tile, n_tile, split, kb0, nkb; };
__device__ __forceinline__ GmItem gm_item(const GemmArgs& a, long long w) {
    GmItem it;
    it.split = (int)(w % a.splits);
    const long long t = w / a.splits;
    it.n_tile = (int)(t % a.n_tiles);
    it.m_tile = (int)(t / a.n_tiles);
    it.kb0 = it.split * a.kb_per_split;
    it.nkb = min(a.kb_per_split, a.kb_total - it.kb0);
    return it;
}

template <int CG>
__global__ void __launch_bounds__(GM_THREADS, 1) gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmB,
                                                                  const __grid_constant__ CUtensorMap tmOut,
                                                                  const __grid_constant__ CUtensorMap tmOut2,
                                                                  const __grid_constant__ CUtensorMap tmAux, const GemmArgs a) {
    using C = GemmCfg<CG>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int NST = a.stages;
    const bool use_aux_stage = (a.epi == GM_EPI_ADD || a.epi == GM_EPI_ACT_BWD || a.has_out2);
    unsigned char* ring = smem;
    unsigned char* out_stage = smem + (size_t)NST * C::STAGE_BYTES;               // [obuf][8 warps][4 KiB], 1 KiB aligned
    unsigned char* aux_stage = out_stage + (size_t)a.obuf * GM_EPI_WARPS * GM_CHUNK_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux_stage + (use_aux_stage ? GM_EPI_WARPS * GM_CHUNK_BYTES : 0));
    uint64_t* empty_bar = full_bar + GM_MAX_STAGES;
    uint64_t* tfull_bar = empty_bar + GM_MAX_STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* aux_bar = tempty_bar + 2;                                             // [8]
    uint32_t* tmem_slot_p = reinterpret_cast<uint32_t*>(aux_bar + GM_EPI_WARPS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (CG == 2) ? (int)cluster_ctarank() : 0;
    const long long group = blockIdx.x / CG, n_groups = gridDim.x / CG;
    const long long n_items = (long long)a.m_tiles * a.n_tiles * a.splits;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], CG * GM_EPI_WARPS); }
        for (int w = 0; w < GM_EPI_WARPS; ++w) mbar_init(&aux_bar[w], 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        if (CG == 2) tmem_alloc_cg2(tmem_slot_p, GM_TMEM_COLS);
        else tmem_alloc(tmem_slot_p, GM_TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();            // the peer's barriers are initialised before any remote arrive / TMA credit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_p;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (one thread in every CTA)
        if (lane == 0) {
            // ring position as (stage, phase) counters: no division on the per-k-block path of any role
            int s = 0;
            uint32_t ph = 0;
            for (long long w = group; w < n_items; w += n_groups) {
                const GmItem wi = gm_item(a, w);
                const int m0 = (a.debug & 8) ? rank * GM_BM : wi.m_tile * (GM_BM * CG) + rank * GM_BM;   // this CTA's rows of A
                const int n0 = wi.n_tile * GM_BN + rank * C::B_ROWS;                // this CTA's rows of B
                for (int kb = 0; kb < wi.nkb; ++kb, s = (s + 1 == NST) ? 0 : s + 1, ph ^= (s == 0) ? 1u : 0u) {
                    const int k0 = (a.debug & 8) ? 0 : (wi.kb0 + kb) * GM_BK;      // 8 = always the same (L2-hot) k-block
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    unsigned char* st = ring + (size_t)s * C::STAGE_BYTES;
                    if (a.debug & 2) {
                        if (rank == 0) mbar_arrive(&full_bar[s]);
                        continue;
                    }
                    if (CG == 2) {
                        // the pair leader's barrier collects the bytes of both CTAs
                        const uint32_t bar = mapa_u32(smem_u32(&full_bar[s]), 0);
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * C::STAGE_BYTES);
                        if (a.a_mn) {
                            for (int q = 0; q < GM_BM / 32; ++q)
                                tma_load_2d_cg2(st + q * 4096, &tmA, m0 + q * 32, k0, bar);
                        } else {
                            tma_load_2d_cg2(st, &tmA, k0, m0, bar);
                        }
                        if (a.b_mn) {
                            for (int q = 0; q < C::B_ROWS / 32; ++q)
                                tma_load_2d_cg2(st + GM_A_BYTES + q * 4096, &tmB, n0 + q * 32, k0, bar);
                        } else {
                            tma_load_2d_cg2(st + GM_A_BYTES, &tmB, k0, n0, bar);
                        }
                    } else {
                        mbar_arrive_expect_tx(&full_bar[s], C::STAGE_BYTES);
                        if (a.a_mn) {
                            for (int q = 0; q < GM_BM / 32; ++q) tma_load_2d(st + q * 4096, &tmA, m0 + q * 32, k0, &full_bar[s]);
                        } else {
                            tma_load_2d(st, &tmA, k0, m0, &full_bar[s]);
                        }
                        if (a.b_mn) {
                            for (int q = 0; q < C::B_ROWS / 32; ++q)
                                tma_load_2d(st + GM_A_BYTES + q * 4096, &tmB, n0 + q * 32, k0, &full_bar[s]);
                        } else {
                            tma_load_2d(st + GM_A_BYTES, &tmB, k0, n0, &full_bar[s]);
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (one thread of the pair leader)
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = gm_idesc(GM_BM * CG, GM_BN, a.a_mn, a.b_mn);
            const uint32_t a_step = a.a_mn ? (1024u >> 4) : (32u >> 4);            // 8 k-rows: two 512-B groups / 32 bytes along the row
            const uint32_t b_step = a.b_mn ? (1024u >> 4) : (32u >> 4);
            // descriptors differ between stages only in their 14-bit address field: add the stage offset to stage 0's
            const uint32_t ring0 = smem_u32(ring);
            const uint64_t adesc0 = gm_desc(ring0, a.a_mn), bdesc0 = gm_desc(ring0 + GM_A_BYTES, a.b_mn);
            constexpr uint32_t stage_step = (uint32_t)C::STAGE_BYTES >> 4;
            int s = 0;
            uint32_t ph = 0, tc = 0;
            for (long long w = group; w < n_items; w += n_groups, ++tc) {
                const GmItem wi = gm_item(a, w);
                const int buf = (int)(tc & 1u);
                mbar_wait(&tempty_bar[buf], ((tc >> 1) & 1u) ^ 1u);                 // both epilogues drained this TMEM buffer
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)buf * GM_BN;
                for (int kb = 0; kb < wi.nkb; ++kb, s = (s + 1 == NST) ? 0 : s + 1, ph ^= (s == 0) ? 1u : 0u) {
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint64_t adesc = adesc0 + (uint64_t)(stage_step * (uint32_t)s);
                    const uint64_t bdesc = bdesc0 + (uint64_t)(stage_step * (uint32_t)s);
#pragma unroll
                    for (int k4 = 0; k4 < GM_BK / 8; ++k4) {
                        if (CG == 2) umma_tf32_cg2(d_tmem, adesc + a_step * k4, bdesc + b_step * k4, idesc, (kb | k4) ? 1u : 0u);
                        else umma_tf32(d_tmem, adesc + a_step * k4, bdesc + b_step * k4, idesc, (kb | k4) ? 1u : 0u);
                    }
                    if (CG == 2) umma_commit_cg2(&empty_bar[s]);                    // stage s free in both CTAs
                    else umma_commit(&empty_bar[s]);
                }
                if (CG == 2) umma_commit_cg2(&tfull_bar[buf]);                      // accumulator complete, both CTAs
                else umma_commit(&tfull_bar[buf]);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue warps 2..9: thread == output row
        const int ew = warp - 2;
        const int q = warp & 3;                                                     // TMEM lane quadrant of this warp
        // chunk c of this warp = columns [64c + 32*half, +32) of the tile's 256: float4 columns f and f + 32 of a 512-byte
        // group then belong to the SAME thread (chunks c and c + 2), which is what lets one Philox draw serve both (dropout)
        const int half = ew >> 2;
        float* ost0 = reinterpret_cast<float*>(out_stage + ew * GM_CHUNK_BYTES);
        float* ast = reinterpret_cast<float*>(aux_stage + ew * GM_CHUNK_BYTES);
        const int obuf_stride = (a.obuf > 1) ? GM_EPI_WARPS * GM_CHUNK_BYTES / 4 : 0;   // floats between this warp's two buffers
        int ob = 0;
        uint64_t* abar = &aux_bar[ew];
        const bool has_aux = (a.epi == GM_EPI_ADD || a.epi == GM_EPI_ACT_BWD);
        const uint32_t tempty_leader = (CG == 2) ? mapa_u32(smem_u32(&tempty_bar[0]), 0) : 0u;
        uint32_t aux_par = 0;
        const int sw = lane & 7;
        long long tc = 0;
        if (has_aux && lane == 0 && group < n_items) {                              // aux chunk of the first item
            const GmItem wi = gm_item(a, group);
            mbar_arrive_expect_tx(abar, GM_CHUNK_BYTES);
            tma_load_2d(ast, &tmAux, wi.n_tile * GM_BN + half * 32, wi.m_tile * (GM_BM * CG) + rank * GM_BM + q * 32, abar);
        }
        const bool drop = a.p_drop > 0.f;
        const Philox ph(drop ? PR_SEED(a) : 0ull);
        const unsigned drop_thr = drop_threshold(a.p_drop);
        const float drop_ik = 1.0f / (1.0f - a.p_drop);
        const unsigned long long D4 = (unsigned long long)(a.N >> 2);
        for (long long w = group; w < n_items; w += n_groups, ++tc) {
            const GmItem wi = gm_item(a, w);
            const int buf = (int)(tc & 1);
            const int row0 = wi.m_tile * (GM_BM * CG) + rank * GM_BM + q * 32;
            const int colb = wi.n_tile * GM_BN + half * 32;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * GM_BN + half * 32);
            // this warp's 128 bias values: lane l holds the 4 columns at (chunk l / 8, float4 l % 8); broadcast by shuffles below
            // (one global load per lane and tile, issued before the wait for the accumulator)
            float4 bl = make_float4(0.f, 0.f, 0.f, 0.f);
            {
                const int bc = colb + (lane >> 3) * 64 + (lane & 7) * 4;
                if (a.bias && bc < a.N) bl = __ldg(reinterpret_cast<const float4*>(a.bias + bc));
            }
            unsigned keep_hi0 = 0, keep_hi1 = 0;                                   // high nibbles of chunks 0 / 1, used by chunks 2 / 3
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int col0 = colb + c * 64;
                float x[32];
                if (has_aux) {
                    mbar_wait(abar, aux_par);
                    aux_par ^= 1u;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 t = *reinterpret_cast<const float4*>(ast + lane * 32 + ((j ^ sw) << 2));
                        x[4 * j] = t.x; x[4 * j + 1] = t.y; x[4 * j + 2] = t.z; x[4 * j + 3] = t.w;
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {                                               // next chunk's aux (this item or the next one)
                        int nr = row0, nc = col0 + 64;
                        bool more = true;
                        if (c == 3) {
                            const long long wn = w + n_groups;
                            more = wn < n_items;
                            if (more) {
                                const GmItem nx = gm_item(a, wn);
                                nr = nx.m_tile * (GM_BM * CG) + rank * GM_BM + q * 32;
                                nc = nx.n_tile * GM_BN + half * 32;
                            }
                        }
                        if (more) {
                            mbar_arrive_expect_tx(abar, GM_CHUNK_BYTES);
                            tma_load_2d(ast, &tmAux, nc, nr, abar);
                        }
                    }
                }
                if (c == 0) {
                    mbar_wait(&tfull_bar[buf], (uint32_t)((tc >> 1) & 1));
                    tc_fence_after();
                }
                float v[32];
                __syncwarp();                                                      // tcgen05.ld is warp-collective
                tmem_ld32(taddr + c * 64, v);
                if (c == 3) {                                                      // last TMEM read of this item: hand the buffer back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (CG == 2) mbar_arrive_cluster(tempty_leader + (uint32_t)buf * 8u);
                        else mbar_arrive(&tempty_bar[buf]);
                    }
                }
                if (a.debias) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= 1.00071f;
                }
                if (a.bias) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int src = c * 8 + j;
                        v[4 * j] += __shfl_sync(0xffffffffu, bl.x, src);
                        v[4 * j + 1] += __shfl_sync(0xffffffffu, bl.y, src);
                        v[4 * j + 2] += __shfl_sync(0xffffffffu, bl.z, src);
                        v[4 * j + 3] += __shfl_sync(0xffffffffu, bl.w, src);
                    }
                }
                if (a.epi == GM_EPI_ADD) {
                    if (drop) {
                        // ln.cu's layout: float4 column f of row r takes nibble (f / 32) & 1 of keep_bits8(Philox(r * D4 + f % 32 + 64 * (f / 64)));
                        // here f = 64 * n_tile + 16 * c + 8 * half + j, so chunks c and c + 2 share their eight draws
                        unsigned kb = (c & 1) ? keep_hi1 : keep_hi0;
                        if (c < 2) {
                            const unsigned long long ctr = (unsigned long long)(row0 + lane) * D4 + (unsigned)(64 * wi.n_tile + 16 * c + 8 * half);
                            unsigned lo = 0, hi = 0;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const unsigned m = keep_bits8(ph(ctr + j, a.rng_stream), drop_thr);
                                lo |= (m & 15u) << (4 * j);
                                hi |= (m >> 4) << (4 * j);
                            }
                            kb = lo;
                            if (c == 0) keep_hi0 = hi; else keep_hi1 = hi;
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = ((kb >> j) & 1u) ? v[j] * drop_ik : 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += x[j];
                } else if (a.epi == GM_EPI_ACT_BWD) {
                    if (a.act == PR_ACT_GELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= dgelu_fast(x[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= act_df(x[j], a.act);
                    }
                }
                if (a.epi == GM_EPI_ACT && a.has_out2) {                           // pre-activation copy (the backward needs it)
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(ast + lane * 32 + ((j ^ sw) << 2)) =
                            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmOut2, ast, col0, row0);
                        bulk_commit();
                    }
                }
                if (a.epi == GM_EPI_ACT) {
                    if (a.act == PR_ACT_GELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
                    } else if (a.act == PR_ACT_RELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                    } else if (a.act == PR_ACT_QUICK_GELU) {                       // x * sigmoid(1.702 x), CLIP ViT (fc1)
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __fdividef(v[j], 1.0f + __expf(-1.702f * v[j]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = act_f(v[j], a.act);
                    }
                }
                // staging buffer free again once the TMA store that last read it has done so (two buffers: the store before last)
                float* ost = ost0 + ob * obuf_stride;
                if (a.debug & 4) continue;                                          // 4 = no staging at all (TMEM drain only)
                if (lane == 0) {
                    if (a.obuf > 1 && !a.has_out2) bulk_wait_read<1>();
                    else bulk_wait_read<0>();
                }
                ob ^= 1;
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(ost + lane * 32 + ((j ^ sw) << 2)) =
                        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0 && !(a.debug & 1)) {
                    if (a.splits > 1) tma_store_3d(&tmOut, ost, col0, row0, wi.split);
                    else tma_store_2d(&tmOut, ost, col0, row0);
                    bulk_commit();
                }
                if (a.colsum) {                                                    // column `lane` of the staged chunk, rows in order
                    float s = 0.f;
#pragma unroll
                    for (int r = 0; r < 32; ++r) s += ost[r * 32 + ((((lane >> 2) ^ (r & 7)) << 2) | (lane & 3))];
                    const long long col = (long long)col0 + lane;
                    if (col < a.N) a.colsum[(long long)(row0 / 32) * a.N + col] = s;
                }
            }
        }
        if (lane == 0) bulk_wait_all<0>();                                          // stores complete before the CTA exits
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();            // no CTA exits (or frees TMEM) while its peer can still reach it
    if (warp == 1) {
        if (CG == 2) tmem_dealloc_cg2(tmem_base, GM_TMEM_COLS);
        else tmem_dealloc(tmem_base, GM_TMEM_COLS);
    }
}

// out[i] = sum_s partials[s][i], s ascending (deterministic); n % 4 == 0
__global__ void __launch_bounds__(256) gemm_splitk_reduce_kernel(const float4* __restrict__ part, int splits, long long n4,
                                                                 float4* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 s = part[i];
        for (int p = 1; p < splits; ++p) {
            const float4 t = part[(long long)p * n4 + i];
            s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        out[i] = s;
    }
}

// ---- host --------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn gm_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
// fp32 tensor [d2][d1][d0] (d0 contiguous; strides in elements), box [b2][b1][b0]
static int gm_map(CUtensorMap* tm, const float* base, int rank, const long long* dims, const long long* strides,
                  const int* box, CUtensorMapSwizzle swz, const char* what) {
    EncodeTiledFn fn = gm_encode_fn();
    if (!fn) {
        set_last_error("cuTensorMapEncodeTiled entry point not available");
        return PR_ERR_UNSUPPORTED;
    }
    cuuint64_t gdim[3];
    cuuint64_t gstride[2];
    cuuint32_t bx[3], estr[3] = {1, 1, 1};
    for (int i = 0; i < rank; ++i) { gdim[i] = (cuuint64_t)dims[i]; bx[i] = (cuuint32_t)box[i]; }
    for (int i = 1; i < rank; ++i) gstride[i - 1] = (cuuint64_t)strides[i] * 4;
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void*)base, gdim, gstride, bx, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
        return PR_ERR_INVALID_ARGUMENT;
    }
    return PR_OK;
}
// operand with `rows` rows (M or N) and contraction length K.  K-major: memory [rows][K] (ld elements per row);
// MN-major: memory [K][rows] (ld elements per k-row).
static int gm_operand_map(CUtensorMap* tm, const float* p, long long rows, long long K, long long ld, int mn_major,
                          int box_rows, const char* what) {
    if (mn_major) {
        const long long dims[2] = {rows, K}, strides[2] = {1, ld};
        const int box[2] = {32, GM_BK};
        return gm_map(tm, p, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, what);
    }
    const long long dims[2] = {K, rows}, strides[2] = {1, ld};
    const int box[2] = {GM_BK, box_rows};
    return gm_map(tm, p, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, what);
}

static int gm_cg() {            // CTA-pair pipeline unless PR_GEMM_CG=1 (A/B against the single-CTA form)
    static int cg = 0;
    if (!cg) {
        const char* e = getenv("PR_GEMM_CG");
        cg = (e && atoi(e) == 1) ? 1 : 2;
    }
    return cg;
}

template <int CG>
static int gm_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const CUtensorMap& tmOut2,
                     const CUtensorMap& tmAux, const GemmArgs& a, cudaStream_t stream) {
    auto kern = gemm_tf32_kernel<CG>;
    const size_t smem = gm_smem_bytes(GemmCfg<CG>::STAGE_BYTES, a.stages, a.obuf,
                                      a.epi == GM_EPI_ADD || a.epi == GM_EPI_ACT_BWD || a.has_out2);
    static bool attr_set = false;
    if (!attr_set) {
        PR_CUDA_CALL(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM_LIMIT));
        attr_set = true;
    }
    const long long n_items = (long long)a.m_tiles * a.n_tiles * a.splits;
    long long groups = std::min<long long>(n_items, sm_count() / CG);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3((unsigned)(groups * CG));
    cfg.blockDim = dim3(GM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    PR_CUDA_CALL(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmOut, tmOut2, tmAux, a));
    PR_CUDA_LAUNCH_CHECK("gemm_tf32_kernel");
    return PR_OK;
}

}  // namespace pr

using namespace pr;

extern "C" int pr_gemm_colsum_rows(int64_t M) {
    const int cg = gm_cg();
    const long long mt = (M + GM_BM * cg - 1) / (GM_BM * cg);
    return (int)(mt * (GM_BM * cg) / 32);
}

static int gemm_impl(const float* A, int a_mn, int64_t lda, const float* B, int b_mn, int64_t ldb, int64_t M, int64_t N,
                     int64_t K, const float* bias, const float* aux, int epi, int act, float* out, float* out2,
                     int splits, float* colsum_partials, int flags, float p_drop, uint64_t seed, uint32_t rng_stream,
                     pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PR_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "pr_gemm_tf32_drop: p_drop=%f outside [0, 1)", (double)p_drop);
    PR_CHECK_ARG(p_drop == 0.f || epi == GM_EPI_ADD, "pr_gemm_tf32_drop: dropout belongs to the PR_GEMM_ADD epilogue");
    PR_CHECK_ARG(M > 0 && N > 0 && K > 0, "pr_gemm_tf32: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
    // a K tail needs no code: the tensor maps carry the true K and TMA zero-fills what lies beyond it
    PR_CHECK_ARG((a_mn || K % 4 == 0) && (b_mn || K % 4 == 0), "pr_gemm_tf32: a K-major operand needs K %% 4 == 0 (K=%lld)", (long long)K);
    PR_CHECK_ARG(N % 4 == 0, "pr_gemm_tf32: N=%lld must be a multiple of 4", (long long)N);
    PR_CHECK_ARG(!a_mn || M % 4 == 0, "pr_gemm_tf32: an MN-major A needs M %% 4 == 0 (M=%lld)", (long long)M);
    PR_CHECK_ARG(A && B && out, "pr_gemm_tf32: null pointer");
    PR_CHECK_ARG(lda % 4 == 0 && ldb % 4 == 0 && lda >= (a_mn ? M : K) && ldb >= (b_mn ? N : K),
                 "pr_gemm_tf32: bad leading dimensions lda=%lld ldb=%lld", (long long)lda, (long long)ldb);
    PR_CHECK_ARG(aligned16(A) && aligned16(B) && aligned16(out) && aligned16(out2) && aligned16(bias) && aligned16(aux),
                 "pr_gemm_tf32: pointers must be 16-byte aligned");
    PR_CHECK_ARG(epi >= GM_EPI_STORE && epi <= GM_EPI_ACT_BWD, "pr_gemm_tf32: epi=%d", epi);
    PR_CHECK_ARG((epi != GM_EPI_ADD && epi != GM_EPI_ACT_BWD) || aux, "pr_gemm_tf32: this epilogue needs aux");
    PR_CHECK_ARG((epi != GM_EPI_ACT && epi != GM_EPI_ACT_BWD) || (act >= 0 && act <= 5), "pr_gemm_tf32: act=%d", act);
    PR_CHECK_ARG(splits >= 1 && (splits == 1 || (epi == GM_EPI_STORE && !bias && !colsum_partials)),
                 "pr_gemm_tf32: split-K stores raw partial sums (no bias / epilogue / column sums)");
    PR_CHECK_ARG(!colsum_partials || !bias, "pr_gemm_tf32: column sums are taken of the stored output; combine with bias is unsupported");
    PR_CHECK_ARG(M < (1LL << 31) - 512 && N < (1LL << 31) - 512 && K < (1LL << 31) - 512, "pr_gemm_tf32: shape too large");
    const int cg = gm_cg();
    GemmArgs a;
    a.m_tiles = (int)((M + GM_BM * cg - 1) / (GM_BM * cg));
    a.n_tiles = (int)((N + GM_BN - 1) / GM_BN);
    a.kb_total = (int)((K + GM_BK - 1) / GM_BK);
    splits = std::min(splits, a.kb_total);
    a.kb_per_split = (a.kb_total + splits - 1) / splits;
    a.splits = (a.kb_total + a.kb_per_split - 1) / a.kb_per_split;
    PR_CHECK_ARG(a.splits == splits, "pr_gemm_tf32: splits=%d does not divide %d k-blocks into non-empty parts", splits, a.kb_total);
    a.a_mn = a_mn ? 1 : 0; a.b_mn = b_mn ? 1 : 0;
    a.epi = epi; a.act = act; a.M = M; a.N = N; a.bias = bias; a.colsum = colsum_partials; a.has_out2 = out2 ? 1 : 0;
    a.debias = (flags & 1) ? 1 : 0;
    a.p_drop = p_drop; a.seed = seed; a.rng_stream = rng_stream;
    PR_SET_SEED_DEV(a);
    {
        // deepest ring that fits beside the staging buffers (PR_GEMM_STAGES / PR_GEMM_OBUF override, for A/B runs)
        const int stage_bytes = (cg == 2) ? GemmCfg<2>::STAGE_BYTES : GemmCfg<1>::STAGE_BYTES;
        const int aux_st = (epi == GM_EPI_ADD || epi == GM_EPI_ACT_BWD || out2) ? 1 : 0;
        static int env_obuf = -1, env_stages = -1;
        if (env_obuf < 0) {
            const char* e = getenv("PR_GEMM_OBUF");
            env_obuf = e ? std::max(1, std::min(2, atoi(e))) : 0;
            const char* f = getenv("PR_GEMM_STAGES");
            env_stages = f ? std::max(2, std::min(GM_MAX_STAGES, atoi(f))) : 0;
        }
        static int env_debug = -1;
        if (env_debug < 0) {
            const char* d = getenv("PR_GEMM_DEBUG");
            env_debug = d ? atoi(d) : 0;
        }
        a.debug = env_debug;
        a.obuf = env_obuf ? env_obuf : 1;
        int st = GM_MAX_STAGES;
        while (st > 2 && gm_smem_bytes(stage_bytes, st, a.obuf, aux_st) > (size_t)GM_SMEM_LIMIT) --st;
        if (env_stages) st = std::min(st, env_stages);
        a.stages = st;
        PR_CHECK_ARG(gm_smem_bytes(stage_bytes, st, a.obuf, aux_st) <= (size_t)GM_SMEM_LIMIT, "pr_gemm_tf32: shared memory budget");
    }
    CUtensorMap tmA, tmB, tmOut, tmOut2, tmAux;
    int rc = gm_operand_map(&tmA, A, M, K, lda, a.a_mn, GM_BM, "A");
    if (rc) return rc;
    rc = gm_operand_map(&tmB, B, N, K, ldb, a.b_mn, GM_BN / cg, "B");
    if (rc) return rc;
    {
        const int box[3] = {32, 32, 1};
        if (a.splits > 1) {
            const long long dims[3] = {N, M, a.splits}, strides[3] = {1, N, M * N};
            rc = gm_map(&tmOut, out, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, "out (split-K partials)");
        } else {
            const long long dims[2] = {N, M}, strides[2] = {1, N};
            rc = gm_map(&tmOut, out, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, "out");
        }
        if (rc) return rc;
        const long long dims[2] = {N, M}, strides[2] = {1, N};
        tmOut2 = tmOut;                                    // unused maps are never dereferenced: skip their (host-side) encoding
        tmAux = tmOut;
        if (out2) {
            rc = gm_map(&tmOut2, out2, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, "out2");
            if (rc) return rc;
        }
        if (aux) {
            rc = gm_map(&tmAux, aux, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, "aux");
            if (rc) return rc;
        }
    }
    return (cg == 2) ? gm_launch<2>(tmA, tmB, tmOut, tmOut2, tmAux, a, stream) : gm_launch<1>(tmA, tmB, tmOut, tmOut2, tmAux, a, stream);
}

extern "C" int pr_gemm_tf32(const float* A, int a_mn, int64_t lda, const float* B, int b_mn, int64_t ldb, int64_t M, int64_t N,
                            int64_t K, const float* bias, const float* aux, int epi, int act, float* out, float* out2,
                            int splits, float* colsum_partials, int flags, pr_stream_t stream) {
    return gemm_impl(A, a_mn, lda, B, b_mn, ldb, M, N, K, bias, aux, epi, act, out, out2, splits, colsum_partials, flags, 0.f, 0, 0,
                     stream);
}

extern "C" int pr_gemm_tf32_drop(const float* A, int a_mn, int64_t lda, const float* B, int b_mn, int64_t ldb, int64_t M, int64_t N,
                                 int64_t K, const float* bias, const float* aux, float* out, int flags, float p_drop,
                                 uint64_t seed, uint32_t rng_stream, pr_stream_t stream) {
    return gemm_impl(A, a_mn, lda, B, b_mn, ldb, M, N, K, bias, aux, PR_GEMM_ADD, -1, out, nullptr, 1, nullptr, flags, p_drop, seed,
                     rng_stream, stream);
}

extern "C" int pr_gemm_splitk_reduce_f32(const float* partials, int splits, int64_t n, float* out, pr_stream_t stream_) {
    PR_CHECK_ARG(splits >= 1 && n >= 0 && n % 4 == 0, "pr_gemm_splitk_reduce_f32: splits=%d n=%lld", splits, (long long)n);
    if (n == 0) return PR_OK;
    PR_CHECK_ARG(partials && out && aligned16(partials) && aligned16(out), "pr_gemm_splitk_reduce_f32: null or unaligned pointer");
    const long long n4 = n / 4;
    const int grid = (int)std::max<long long>(1, std::min<long long>((n4 + 255) / 256, (long long)sm_count() * 8));
    gemm_splitk_reduce_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>((const float4*)partials, splits, n4, (float4*)out);
    PR_CUDA_LAUNCH_CHECK("gemm_splitk_reduce_kernel");
    return PR_OK;
}
