// pixelrec_b200 -- K9: full-catalog scoring GEMM fused with masking and per-row top-k, on tcgen05 tensor cores.
//   replaces  scores = seq_output @ item_feature.T            REC/model/IDNet/sasrec.py:112
//             scores[:,0] = -inf ; scores[history] = -inf     REC/trainer/trainer.py:334-336
//             torch.topk(scores, max(topk))                   REC/evaluator/collector.py:133
//   (reference: cuBLAS SGEMM -> 397 MB [1024, 97K] logits written, re-read by index_put and by topk).
//
// This is the one genuine dense contraction of the path: 2*B_e*D*N FLOP (101.7 GFLOP at C2) -> tensor cores.
//   * operands stay fp32 in HBM/L2; TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) stages [128 x 32] / [256 x 32]
//     K-major tiles into a 4-stage shared-memory ring; tcgen05.mma.kind::tf32 (UMMA 128x256x8, single issuing
//     thread) accumulates a [128 users x 256 items] fp32 tile in TMEM; two TMEM buffers (2 x 256 columns) let the
//     epilogue of tile t overlap the MMAs of tile t+1.
//   * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue (one TMEM lane
//     quadrant each; thread == user row).  Epilogue: tcgen05.ld 32 columns at a time, apply the mask bitmap word
//     (pad column 0, history, columns >= N), keep a sorted top-K list per thread in registers.  Logits are never
//     written to memory.
//   * grid = m_tiles x n_splits (<= #SMs); every CTA walks its share of the N tiles; a small second kernel merges
//     the n_splits candidate lists per row (ties -> lower item id, like a stable sort).
//
// v2 (score_topk2_kernel, PR_TUNE_SCORE_V2; + PR_TUNE_SCORE_MCAST) -- same MMA pipeline, different epilogue and operand feed:
//   * v1's epilogue is what bounds it (ncu: tensor pipe 16-20 % active, L2 15 %): every candidate insertion is a divergent
//     branch that serialises the warp (~3.5 K insert events per warp and CTA), and one epilogue warp per scheduler cannot
//     hide the latency of the shared-memory candidate walk.  v2 builds a 32-bit candidate word per 32-column chunk with
//     compares only, then runs a WARP-UNIFORM loop (trip count = max candidates of any lane, ~1-2 per chunk) whose body
//     is a predicated register select tree + predicated sorted insert: no branches, no shared memory.  8 epilogue warps
//     (two per TMEM lane quadrant, 128 columns each) double the issue slots; every thread keeps its own list.
//   * MCAST: the m-tiles of an eval batch (8 at B_e=1024) all need the same [256 x 32] table tile.  They form a thread
//     block cluster; CTA r loads rows [r*256/CL, (r+1)*256/CL) of the tile and TMA-multicasts them into every CTA of
//     the cluster (empty barriers collect one tcgen05.commit.multicast arrive per CTA before a stage is overwritten).
//     L2->SM operand traffic per k-block drops from 48 KiB to 16 + 32/CL KiB (CL=8: 20 KiB).
#include <cuda.h>

#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

#include "tc_ptx.cuh"

#define PR_DYN_SMEM_BYTES(name) extern __shared__ __align__(1024) unsigned char name[]
#define PR_LDG4(p) __ldg(p)
#include "score_kernels.cuh"

namespace pr {

template <int K>
__global__ void __launch_bounds__(SC_THREADS, 1) score_topk_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmB,
                                                                   const ScoreArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // SWIZZLE_128B tiles (TMA destination == UMMA operand) must sit on 1024-byte boundaries
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t full_bar[SC_STAGES], empty_bar[SC_STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_slot;
    float* cand_smem = reinterpret_cast<float*>(smem + (size_t)SC_STAGES * SC_STAGE_BYTES);   // [128 epilogue threads][33]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x % a.m_tiles, split = blockIdx.x / a.m_tiles;
    const int t_begin = split * a.tiles_per_split;
    const int t_end = min(a.n_tiles, t_begin + a.tiles_per_split);
    const int n_my = t_end - t_begin;

    if (threadIdx.x == 0) {
        for (int s = 0; s < SC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 4); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(&tmem_slot, SC_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            long long it = 0;
            for (int t = 0; t < n_my; ++t) {
                const int n0 = (t_begin + t) * SC_BN;
                for (int kb = 0; kb < a.kblocks; ++kb, ++it) {
                    const int s = (int)(it % SC_STAGES);
                    mbar_wait(&empty_bar[s], (uint32_t)(((it / SC_STAGES) & 1) ^ 1));
                    mbar_arrive_expect_tx(&full_bar[s], SC_STAGE_BYTES);
                    unsigned char* st = smem + (size_t)s * SC_STAGE_BYTES;
                    tma_load_2d(st, &tmA, kb * SC_BK, m_tile * SC_BM, &full_bar[s]);
                    tma_load_2d(st + SC_A_BYTES, &tmB, kb * SC_BK, n0, &full_bar[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc = tf32_idesc(SC_BM, SC_BN);
            long long it = 0;
            for (int t = 0; t < n_my; ++t) {
                const int buf = t & 1;
                mbar_wait(&tempty_bar[buf], (uint32_t)(((t >> 1) & 1) ^ 1));   // epilogue drained this TMEM buffer
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)buf * SC_BN;
                for (int kb = 0; kb < a.kblocks; ++kb, ++it) {
                    const int s = (int)(it % SC_STAGES);
                    mbar_wait(&full_bar[s], (uint32_t)((it / SC_STAGES) & 1));
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)s * SC_STAGE_BYTES);
                    const uint64_t adesc = sw128_kmajor_desc(sa), bdesc = sw128_kmajor_desc(sa + SC_A_BYTES);
#pragma unroll
                    for (int k4 = 0; k4 < SC_BK / 8; ++k4)   // UMMA K = 8 tf32 = 32 bytes -> +2 in the 16-byte address field
                        umma_tf32(d_tmem, adesc + 2 * k4, bdesc + 2 * k4, idesc, (kb | k4) ? 1u : 0u);
                    umma_commit(&empty_bar[s]);                // stage reusable once these MMAs have read it
                }
                umma_commit(&tfull_bar[buf]);                   // accumulator tile complete
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue warps 2..5 (thread == user row)
        const int q = warp & 3;                                 // TMEM lane quadrant this warp may access
        const int row = m_tile * SC_BM + q * 32 + lane;
        float val[K];
        int idx[K];
#pragma unroll
        for (int i = 0; i < K; ++i) { val[i] = -INFINITY; idx[i] = -1; }
        const uint32_t* mrow = a.mask + (size_t)row * a.n_words;
        for (int t = 0; t < n_my; ++t) {
            const int buf = t & 1;
            const int tile = t_begin + t;
            const uint4 mw0 = *reinterpret_cast<const uint4*>(mrow + tile * 8);
            const uint4 mw1 = *reinterpret_cast<const uint4*>(mrow + tile * 8 + 4);
            const uint32_t mw[8] = {mw0.x, mw0.y, mw0.z, mw0.w, mw1.x, mw1.y, mw1.z, mw1.w};
            mbar_wait(&tfull_bar[buf], (uint32_t)((t >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * SC_BN;
#pragma unroll 1
            for (int c = 0; c < SC_BN / 32; ++c) {
                float v[32];
                __syncwarp();                                   // tcgen05.ld is warp-collective (.sync.aligned)
                tmem_ld32(taddr + c * 32, v);
                const uint32_t w = mw[c];
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    v[j] = ((w >> j) & 1u) ? -INFINITY : v[j];
                    mx = fmaxf(mx, v[j]);
                }
                if (mx > val[K - 1]) {
                    // rare path (after the first tiles only a few chunks per row hold a candidate): park the chunk in this
                    // thread's shared-memory row and walk it with ONE copy of the insertion code.  (Inlining 32 unrolled
                    // insertions made the epilogue ~40 KB of SASS: ncu showed 32 % no_instruction stalls -- I-cache misses --
                    // and the tensor pipe only 20 % active because the TMEM buffers were drained too slowly.)
                    const int c0 = tile * SC_BN + c * 32;
                    float* myrow = cand_smem + (threadIdx.x - 64) * 33;
#pragma unroll
                    for (int j = 0; j < 32; ++j) myrow[j] = v[j];
#pragma unroll 1
                    for (int j = 0; j < 32; ++j) {
                        const float x = myrow[j];
                        if (x > val[K - 1]) topk_insert<K>(val, idx, x, c0 + j);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);
        }
        float* cv = a.cand_val + ((size_t)row * a.n_splits + split) * K;
        int* ci = a.cand_idx + ((size_t)row * a.n_splits + split) * K;
#pragma unroll
        for (int i = 0; i < K; ++i) { cv[i] = val[i]; ci[i] = idx[i]; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, SC_TMEM_COLS);
}

// ---- host --------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
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

// [rows, D] fp32 row-major -> boxes of [box_rows x 32 floats], 128-byte swizzle, zero fill out of bounds
static int make_map(CUtensorMap* tm, const float* base, long long rows, long long D, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_last_error("cuTensorMapEncodeTiled entry point not available");
        return PR_ERR_UNSUPPORTED;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)D * 4};
    cuuint32_t box[2] = {(cuuint32_t)SC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld D=%lld)", (int)r, rows, D);
        return PR_ERR_INVALID_ARGUMENT;
    }
    return PR_OK;
}

// [rows, D] fp16 row-major -> boxes of [box_rows x 64 halves] (the same 128-byte swizzled rows as the fp32 map)
static int make_map_f16(CUtensorMap* tm, const void* base, long long rows, long long D, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_last_error("cuTensorMapEncodeTiled entry point not available");
        return PR_ERR_UNSUPPORTED;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)D * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled(fp16) failed with CUresult %d (rows=%lld D=%lld)", (int)r, rows, D);
        return PR_ERR_INVALID_ARGUMENT;
    }
    return PR_OK;
}

struct ScorePlan {
    int m_tiles, n_tiles, n_splits, tiles_per_split, n_words, K;
    size_t mask_bytes, cand_bytes, total;
};
static ScorePlan score_plan(long long B_e, long long N, int k) {
    ScorePlan p;
    p.K = (k <= 16) ? 16 : 32;
    p.m_tiles = (int)((B_e + SC_BM - 1) / SC_BM);
    p.n_tiles = (int)((N + SC_BN - 1) / SC_BN);
    int sms = sm_count();
    int splits = std::max(1, sms / std::max(1, p.m_tiles));
    splits = std::min(splits, p.n_tiles);
    splits = std::min(splits, 1024 / p.K);               // merge kernel handles <= 1024 candidates per row
    p.tiles_per_split = (p.n_tiles + splits - 1) / splits;
    p.n_splits = (p.n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
    p.n_words = p.n_tiles * 8;
    const size_t rows = (size_t)p.m_tiles * SC_BM;
    p.mask_bytes = (rows * p.n_words * 4 + 255) / 256 * 256;
    // sized for either kernel variant (the tuning mask may change between the workspace query and the call):
    // v1 keeps n_splits lists per row, v2 two lists (column halves) for each of <= n_splits splits
    p.cand_bytes = (rows * p.n_splits * 2 * p.K * 4 + 255) / 256 * 256;
    p.total = p.mask_bytes + 2 * p.cand_bytes;
    return p;
}

// v2 launch: 8 epilogue warps, optional cluster multicast of the table tile across the m-tiles
template <int K, int MODE, bool F16 = false, bool ARES = false>
static int launch_v2(const CUtensorMap& tmA, const void* W, long long N, long long D, ScoreArgs a, const ScorePlan& p,
                     bool mcast, cudaStream_t stream, int* n_lists) {
    auto kern = score_topk2_kernel<K, MODE, F16, ARES>;
    // ring (+ resident seq_out tile) + barriers + 1 KiB alignment slack
    const size_t smem = (ARES ? (size_t)a.kblocks * SC_A_BYTES + (size_t)3 * SC_B_BYTES : (size_t)SC_STAGES * SC_STAGE_BYTES) +
                        SC2_BAR_BYTES + 1024;
    PR_CUDA_CALL(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int splits = std::min(p.n_splits, 1024 / (2 * K));                      // two lists per split and row
    int CL = 1;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(SC2_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (mcast) {
        // largest power of two <= 8 (or PR_SCORE_CLUSTER, for A/B: clusters of 8 leave SMs idle when a GPC's SM count is not a
        // multiple of 8) that divides the number of m-tiles
        int cl_max = 8;
        if (const char* e = getenv("PR_SCORE_CLUSTER")) cl_max = std::max(1, std::min(8, atoi(e)));
        for (int c = 8; c > 1; c >>= 1)
            if (c <= cl_max && p.m_tiles % c == 0) { CL = c; break; }
        if (CL > 1) {
            // all clusters co-resident in one wave: the multicast couples the CTAs of a cluster, a second wave would idle SMs
            at[0].val.clusterDim.x = CL;
            cfg.gridDim = dim3(p.m_tiles * splits);
            int max_clusters = 0;
            const bool ok = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) == cudaSuccess;
            if (!ok || max_clusters < 1 || max_clusters < p.m_tiles / CL) {
                (void)cudaGetLastError();
                CL = 1;                                                     // cluster shape not schedulable here: plain v2
            } else {
                splits = std::max(1, std::min(splits, max_clusters / (p.m_tiles / CL)));
            }
        }
    }
    a.cluster = CL;
    a.tiles_per_split = (p.n_tiles + splits - 1) / splits;
    a.n_splits = (p.n_tiles + a.tiles_per_split - 1) / a.tiles_per_split;
    *n_lists = a.n_splits * 2;
    CUtensorMap tmB;
    int rc = F16 ? make_map_f16(&tmB, W, N, D, SC_BN / CL)                  // box = this CTA's slice of the table tile
                 : make_map(&tmB, (const float*)W, N, D, SC_BN / CL);
    if (rc) return rc;
    at[0].val.clusterDim.x = CL;
    cfg.gridDim = dim3(p.m_tiles * a.n_splits);
    PR_CUDA_CALL(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, a));
    PR_CUDA_LAUNCH_CHECK("score_topk2_kernel");
    return PR_OK;
}

}  // namespace pr

using namespace pr;

extern "C" size_t pr_score_topk_workspace_bytes(int64_t B_e, int64_t N, int k) {
    if (B_e <= 0 || N <= 0 || k <= 0 || k > 32) return 0;
    return score_plan(B_e, N, k).total;
}

// k_lists: depth of the per-thread lists (16 or 32; >= k unless bound_out certifies what the lists cannot hold)
static int score_topk_impl(const float* seq_out, int64_t B_e, const float* W, int64_t N, int64_t D, const int64_t* hist_u,
                           const int64_t* hist_i, int64_t n_hist, int mask_col0, int k, int k_lists, float* bound_out,
                           float* topk_val, int64_t* topk_idx, void* workspace, size_t workspace_bytes, pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PR_CHECK_ARG(B_e > 0 && N > 0 && D > 0, "pr_score_topk_f32: bad shape B_e=%lld N=%lld D=%lld", (long long)B_e,
                 (long long)N, (long long)D);
    PR_CHECK_ARG(D % SC_BK == 0, "pr_score_topk_f32: D=%lld must be a multiple of %d", (long long)D, SC_BK);
    PR_CHECK_ARG(k >= 1 && k <= 32 && k <= N, "pr_score_topk_f32: k=%d outside [1, min(32, N)]", k);
    PR_CHECK_ARG(N < (1LL << 31) - 512, "pr_score_topk_f32: N too large");
    PR_CHECK_ARG(seq_out && W && topk_val && topk_idx && workspace, "pr_score_topk_f32: null pointer");
    PR_CHECK_ARG(aligned16(seq_out) && aligned16(W), "pr_score_topk_f32: seq_out / W must be 16-byte aligned");
    PR_CHECK_ARG(n_hist >= 0 && (n_hist == 0 || (hist_u && hist_i)), "pr_score_topk_f32: bad history arguments");
    const ScorePlan p = score_plan(B_e, N, k_lists);
    PR_CHECK_ARG(workspace_bytes >= p.total, "pr_score_topk_f32: workspace %zu < required %zu", workspace_bytes, p.total);
    PR_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "pr_score_topk_f32: workspace must be 256-byte aligned");

    char* ws = (char*)workspace;
    uint32_t* mask = (uint32_t*)ws;
    float* cand_val = (float*)(ws + p.mask_bytes);
    int* cand_idx = (int*)(ws + p.mask_bytes + p.cand_bytes);
    const long long rows = (long long)p.m_tiles * SC_BM;
    const long long nmask = rows * p.n_words;
    score_mask_base_kernel<<<(int)((nmask + 255) / 256), 256, 0, stream>>>(mask, rows, p.n_words, N, mask_col0);
    if (n_hist > 0)
        score_mask_hist_kernel<<<(int)((n_hist + 255) / 256), 256, 0, stream>>>(mask, p.n_words, B_e, N, (const long long*)hist_u,
                                                                                (const long long*)hist_i, n_hist);
    PR_CUDA_LAUNCH_CHECK("score_mask kernels");

    CUtensorMap tmA, tmB;
    int rc = make_map(&tmA, seq_out, B_e, D, SC_BM);
    if (rc) return rc;

    ScoreArgs a;
    a.kblocks = (int)(D / SC_BK);
    a.m_tiles = p.m_tiles; a.n_tiles = p.n_tiles; a.tiles_per_split = p.tiles_per_split; a.n_splits = p.n_splits;
    a.n_words = p.n_words; a.mask = mask; a.cand_val = cand_val; a.cand_idx = cand_idx; a.cluster = 1;
    a.target = nullptr; a.n_rows = B_e; a.ce_part = nullptr;
    int n_lists = p.n_splits;
    if (tune() & PR_TUNE_SCORE_V2) {
        const bool mcast = (tune() & PR_TUNE_SCORE_MCAST) != 0;
        rc = (p.K == 16) ? launch_v2<16, 0>(tmA, W, N, D, a, p, mcast, stream, &n_lists)
                         : launch_v2<32, 0>(tmA, W, N, D, a, p, mcast, stream, &n_lists);
        if (rc) return rc;
    } else {
        rc = make_map(&tmB, W, N, D, SC_BN);
        if (rc) return rc;
        const size_t smem = (size_t)SC_STAGES * SC_STAGE_BYTES + 128 * 33 * 4 + 1024;     // ring + candidate rows + 1 KiB alignment slack
        const int grid = p.m_tiles * p.n_splits;
        if (p.K == 16) {
            PR_CUDA_CALL(cudaFuncSetAttribute(score_topk_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            score_topk_kernel<16><<<grid, SC_THREADS, smem, stream>>>(tmA, tmB, a);
        } else {
            PR_CUDA_CALL(cudaFuncSetAttribute(score_topk_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            score_topk_kernel<32><<<grid, SC_THREADS, smem, stream>>>(tmA, tmB, a);
        }
        PR_CUDA_LAUNCH_CHECK("score_topk_kernel");
    }
    score_merge_kernel<<<(int)((B_e + 3) / 4), 128, 0, stream>>>(cand_val, cand_idx, n_lists * p.K, B_e, k, topk_val,
                                                                 (long long*)topk_idx, p.K, bound_out);
    PR_CUDA_LAUNCH_CHECK("score_merge_kernel");
    return PR_OK;
}

extern "C" int pr_score_topk_f32(const float* seq_out, int64_t B_e, const float* W, int64_t N, int64_t D,
                                 const int64_t* hist_u, const int64_t* hist_i, int64_t n_hist, int mask_col0, int k,
                                 float* topk_val, int64_t* topk_idx, void* workspace, size_t workspace_bytes,
                                 pr_stream_t stream_) {
    return score_topk_impl(seq_out, B_e, W, N, D, hist_u, hist_i, n_hist, mask_col0, k, k, nullptr, topk_val, topk_idx, workspace,
                           workspace_bytes, stream_);
}

// ---- id-exact ranking (collector.py:133 ranks fp32 scores): TF32 candidates re-scored in fp32 ---------------------------------
// The tensor core reads TF32 (operands truncated to 10 mantissa bits), so two items whose fp32 scores differ by less than the
// TF32 error can come out in the wrong order (0.7 % of the top-10 ids at C2, profiles/r01f).  Exact mode keeps that pipeline as a
// candidate generator: (1) the usual pass with 16-deep per-thread lists, whose merge returns the best 32 candidates of the row
// AND v_bound = the largest "last kept value" of any full list: an item that is in no list has TF32 score <= v_bound; (2) one
// warp per row recomputes the 32 candidates' scores in fp32 and ranks them; (3) proof of completeness: everything outside the
// 32 has TF32 score <= max(v_bound, 32nd candidate value) =: v_out and hence fp32 score <= v_out + eps,
// eps = 1.25 * 2^-9 * |seq_row| * max_j |W_j| (truncation of both operands, Cauchy-Schwarz) -- if the k-th fp32 score is above
// that, the fp32 top-k is certain; (4) rows that fail the test (near-ties, rare) are flagged and re-ranked over the whole
// catalog in fp32 by score_exact_rows_kernel.
namespace pr {

constexpr int SCX_C = 32;                 // candidates per row

// out_max (zero-initialised) <- max_j |W_j|   (non-negative floats order like their bit patterns)
__global__ void __launch_bounds__(256) rownorm_max_kernel(const float* __restrict__ W, long long N, int D4, float* __restrict__ out_max) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    float best = 0.f;
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < N; r += warps) {
        const float4* p = reinterpret_cast<const float4*>(W) + r * D4;
        float s = 0.f;
        for (int c = lane; c < D4; c += 32) {
            const float4 v = __ldg(p + c);
            s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        best = fmaxf(best, warp_sum(s));
    }
    if (lane == 0) atomicMax(reinterpret_cast<int*>(out_max), __float_as_int(sqrtf(best)));
}

// fp32 dot of two rows, the summation order exact mode defines: lane-strided float4 partial sums, xor-shuffle reduction
__device__ __forceinline__ float warp_dot(const float4* __restrict__ a, const float4* __restrict__ b, int D4, int lane) {
    float s = 0.f;
    for (int c = lane; c < D4; c += 32) {
        const float4 x = a[c], y = __ldg(b + c);
        s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
    }
    return warp_sum(s);
}

__global__ void __launch_bounds__(128) score_rescore_kernel(const float* __restrict__ seq, const float* __restrict__ W, int D4,
                                                            long long B_e, int k, const float* __restrict__ cand_val,
                                                            const long long* __restrict__ cand_idx,
                                                            const float* __restrict__ list_bound,
                                                            const float* __restrict__ w_norm_max, float* __restrict__ out_val,
                                                            long long* __restrict__ out_idx, int* __restrict__ flagged) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= B_e) return;
    const float4* a = reinterpret_cast<const float4*>(seq) + row * D4;
    const long long my_id = cand_idx[row * SCX_C + lane];
    // best TF32 score anything OUTSIDE the 32 candidates can have: the 32nd candidate (if the merged list is full) or an item
    // that fell off one of the per-thread lists
    const bool list_full = cand_idx[row * SCX_C + SCX_C - 1] >= 0;
    const float v_last = fmaxf(list_full ? cand_val[row * SCX_C + SCX_C - 1] : -INFINITY, list_bound[row]);
    float my = -INFINITY;
    for (int c = 0; c < SCX_C; ++c) {
        const long long id = __shfl_sync(0xffffffffu, my_id, c);
        if (id < 0) break;                                            // the list is sorted: the rest is empty too
        const float d = warp_dot(a, reinterpret_cast<const float4*>(W) + id * D4, D4, lane);
        if (lane == c) my = d;
    }
    const float an = sqrtf(warp_dot(a, a, D4, lane));
    int rank = 0;
#pragma unroll
    for (int j = 0; j < SCX_C; ++j) {
        const float sj = __shfl_sync(0xffffffffu, my, j);
        const long long ij = __shfl_sync(0xffffffffu, my_id, j);
        if (ij >= 0 && my_id >= 0 && (sj > my || (sj == my && ij < my_id))) ++rank;
    }
    if (my_id < 0) rank = SCX_C;                                      // empty slots sort last
    if (rank < k) {
        out_val[row * k + rank] = my;
        out_idx[row * k + rank] = my_id;
    }
    const unsigned valid = __ballot_sync(0xffffffffu, my_id >= 0);
    const int n_valid = __popc(valid);
    if (lane >= n_valid && lane < k) {                                // fewer unmasked items than k
        out_val[row * k + lane] = -INFINITY;
        out_idx[row * k + lane] = -1;
    }
    // completeness test against everything that is NOT in the list
    const unsigned kth_lane = __ballot_sync(0xffffffffu, rank == k - 1);
    const float s_k = kth_lane ? __shfl_sync(0xffffffffu, my, __ffs((int)kth_lane) - 1) : -INFINITY;
    const float eps = 1.25f * 0.001953125f * an * w_norm_max[0];
    if (lane == 0 && v_last > -INFINITY && !(s_k > v_last + eps)) {
        const int pos = atomicAdd(flagged, 1);
        flagged[1 + pos] = (int)row;
    }
}

// exact fp32 ranking of one flagged row over the whole catalog (one CTA per row; rare path)
template <int K>
__global__ void __launch_bounds__(256) score_exact_rows_kernel(const float* __restrict__ seq, const float* __restrict__ W, int D4,
                                                               long long N, int k, const uint32_t* __restrict__ mask, int n_words,
                                                               const int* __restrict__ flagged, float* __restrict__ out_val,
                                                               long long* __restrict__ out_idx) {
    extern __shared__ __align__(16) unsigned char sx_smem[];
    float4* a_s = reinterpret_cast<float4*>(sx_smem);                                  // [D4]
    float* lv = reinterpret_cast<float*>(a_s + D4);                                     // [8 warps][K]
    int* li = reinterpret_cast<int*>(lv + 8 * K);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_flag = flagged[0];
    for (int f = blockIdx.x; f < n_flag; f += gridDim.x) {
        const long long row = flagged[1 + f];
        __syncthreads();
        for (int c = threadIdx.x; c < D4; c += blockDim.x) a_s[c] = reinterpret_cast<const float4*>(seq)[row * D4 + c];
        __syncthreads();
        float val[K];
        int idx[K];
#pragma unroll
        for (int i = 0; i < K; ++i) { val[i] = -INFINITY; idx[i] = -1; }
        const uint32_t* mrow = mask + (size_t)row * n_words;
        for (long long i = warp; i < N; i += 8) {                                       // ascending ids: ties keep the lower id
            if ((mrow[i >> 5] >> (i & 31)) & 1u) continue;
            const float d = warp_dot(a_s, reinterpret_cast<const float4*>(W) + i * D4, D4, lane);
            if (d > val[K - 1]) topk_insert<K>(val, idx, d, (int)i);                   // every lane keeps the same list
        }
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < K; ++i) { lv[warp * K + i] = val[i]; li[warp * K + i] = idx[i]; }
        }
        __syncthreads();
        if (warp == 0) {                                                               // merge 8 x K candidates: k rounds of arg-max
            for (int r = 0; r < k; ++r) {
                float bv = -INFINITY;
                int bi = 0x7fffffff, bs = -1;
                for (int c = lane; c < 8 * K; c += 32) {
                    const float v = lv[c];
                    const int id = li[c];
                    if (id >= 0 && (v > bv || (v == bv && id < bi))) { bv = v; bi = id; bs = c; }
                }
                float wv = bv;
                int wi = bi;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
                    if (ov > wv || (ov == wv && oi < wi)) { wv = ov; wi = oi; }
                }
                if (bs >= 0 && wi == bi) li[bs] = -1;                                  // consumed (ids are unique)
                __syncwarp();
                if (lane == 0) {
                    const bool none = (wi == 0x7fffffff);
                    out_val[row * k + r] = none ? -INFINITY : wv;
                    out_idx[row * k + r] = none ? -1 : (long long)wi;
                }
            }
        }
    }
}

}  // namespace pr

static size_t scx_extra_bytes(int64_t B_e) {      // top-32 (val, idx) per row + list bound + [count | flagged rows] + max norm
    return ((size_t)B_e * SCX_C * 12 + (size_t)B_e * 4 + ((size_t)B_e + 2) * 4 + 16 + 255) / 256 * 256;
}

extern "C" size_t pr_score_topk_exact_workspace_bytes(int64_t B_e, int64_t N, int k) {
    if (B_e <= 0 || N <= 0 || k <= 0 || k > 16) return 0;
    return score_plan(B_e, N, 16).total + scx_extra_bytes(B_e);
}

extern "C" int pr_table_norm_max_f32(const float* W, int64_t N, int64_t D, float* out_max, pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PR_CHECK_ARG(W && out_max && N > 0 && D > 0 && D % 4 == 0 && aligned16(W), "pr_table_norm_max_f32: bad arguments");
    PR_CUDA_CALL(cudaMemsetAsync(out_max, 0, 4, stream));
    const int grid = (int)std::min<long long>((N + 7) / 8, (long long)sm_count() * 8);
    rownorm_max_kernel<<<grid, 256, 0, stream>>>(W, N, (int)(D / 4), out_max);
    PR_CUDA_LAUNCH_CHECK("rownorm_max_kernel");
    return PR_OK;
}

extern "C" int pr_score_topk_exact_f32(const float* seq_out, int64_t B_e, const float* W, int64_t N, int64_t D,
                                       const int64_t* hist_u, const int64_t* hist_i, int64_t n_hist, int mask_col0, int k,
                                       const float* w_norm_max, float* topk_val, int64_t* topk_idx, int32_t* n_fallback_rows,
                                       void* workspace, size_t workspace_bytes, pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PR_CHECK_ARG(k >= 1 && k <= 16 && k <= N, "pr_score_topk_exact_f32: k=%d outside [1, min(16, N)]", k);
    PR_CHECK_ARG(B_e > 0 && N > 0 && D > 0 && D % SC_BK == 0, "pr_score_topk_exact_f32: bad shape B_e=%lld N=%lld D=%lld",
                 (long long)B_e, (long long)N, (long long)D);
    PR_CHECK_ARG(seq_out && W && topk_val && topk_idx && workspace, "pr_score_topk_exact_f32: null pointer");
    const size_t need = pr_score_topk_exact_workspace_bytes(B_e, N, k);
    PR_CHECK_ARG(workspace_bytes >= need, "pr_score_topk_exact_f32: workspace %zu < required %zu", workspace_bytes, need);
    const ScorePlan p = score_plan(B_e, N, 16);
    char* extra = (char*)workspace + p.total;
    float* c_val = (float*)extra;
    long long* c_idx = (long long*)(extra + (size_t)B_e * SCX_C * 4);
    float* bound = (float*)(extra + (size_t)B_e * SCX_C * 12);
    int* flagged = (int*)(bound + B_e);
    float* own_max = (float*)(flagged + B_e + 2);
    // (1) TF32 candidates: the ordinary fused pass with 32-deep lists
    PR_CHECK_ARG(N >= SCX_C, "pr_score_topk_exact_f32: N=%lld < %d candidates (use pr_score_topk_f32)", (long long)N, SCX_C);
    int rc = score_topk_impl(seq_out, B_e, W, N, D, hist_u, hist_i, n_hist, mask_col0, SCX_C, 16, bound, c_val, (int64_t*)c_idx,
                             workspace, p.total, stream_);
    if (rc) return rc;
    if (!w_norm_max) {
        rc = pr_table_norm_max_f32(W, N, D, own_max, stream_);
        if (rc) return rc;
        w_norm_max = own_max;
    }
    PR_CUDA_CALL(cudaMemsetAsync(flagged, 0, 4, stream));
    score_rescore_kernel<<<(int)((B_e + 3) / 4), 128, 0, stream>>>(seq_out, W, (int)(D / 4), B_e, k, c_val, c_idx, bound, w_norm_max,
                                                                   topk_val, (long long*)topk_idx, flagged);
    PR_CUDA_LAUNCH_CHECK("score_rescore_kernel");
    // (4) flagged rows: whole-catalog fp32 ranking (grid sized for the worst case; CTAs without a row exit at once)
    const size_t smem = (size_t)D * 4 + 8 * 16 * 8;
    const int grid = (int)std::min<long long>(B_e, (long long)sm_count() * 4);
    PR_CUDA_CALL(cudaFuncSetAttribute(score_exact_rows_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    score_exact_rows_kernel<16><<<grid, 256, smem, stream>>>(seq_out, W, (int)(D / 4), N, k, (const uint32_t*)workspace, p.n_words,
                                                             flagged, topk_val, (long long*)topk_idx);
    PR_CUDA_LAUNCH_CHECK("score_exact_rows_kernel");
    if (n_fallback_rows) PR_CUDA_CALL(cudaMemcpyAsync(n_fallback_rows, flagged, 4, cudaMemcpyDeviceToDevice, stream));
    return PR_OK;
}

// ---- full-catalog softmax cross-entropy (extension: the reference trains with sampled negatives, sasrec.py:88-92) ---------
extern "C" size_t pr_score_ce_workspace_bytes(int64_t B_e, int64_t N) {
    if (B_e <= 0 || N <= 0) return 0;
    const ScorePlan p = score_plan(B_e, N, 1);
    const size_t rows = (size_t)p.m_tiles * SC_BM;
    return p.mask_bytes + (rows * p.n_splits * 2 * 4 * 4 + 255) / 256 * 256;
}

extern "C" int pr_score_ce_f32(const float* seq_out, int64_t B_e, const float* W, int64_t N, int64_t D, const int64_t* target,
                               int mask_col0, float* lse, float* tgt_logit, float* nll, void* workspace, size_t workspace_bytes,
                               pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PR_CHECK_ARG(B_e > 0 && N > 0 && D > 0, "pr_score_ce_f32: bad shape B_e=%lld N=%lld D=%lld", (long long)B_e, (long long)N,
                 (long long)D);
    PR_CHECK_ARG(D % SC_BK == 0, "pr_score_ce_f32: D=%lld must be a multiple of %d", (long long)D, SC_BK);
    PR_CHECK_ARG(N < (1LL << 31) - 512, "pr_score_ce_f32: N too large");
    PR_CHECK_ARG(seq_out && W && workspace && (lse || nll), "pr_score_ce_f32: null pointer");
    PR_CHECK_ARG(target || !(tgt_logit || nll), "pr_score_ce_f32: tgt_logit / nll need target ids");
    PR_CHECK_ARG(aligned16(seq_out) && aligned16(W), "pr_score_ce_f32: seq_out / W must be 16-byte aligned");
    PR_CHECK_ARG(workspace_bytes >= pr_score_ce_workspace_bytes(B_e, N), "pr_score_ce_f32: workspace %zu < required %zu",
                 workspace_bytes, pr_score_ce_workspace_bytes(B_e, N));
    PR_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "pr_score_ce_f32: workspace must be 256-byte aligned");
    const ScorePlan p = score_plan(B_e, N, 1);
    char* ws = (char*)workspace;
    uint32_t* mask = (uint32_t*)ws;
    float* part = (float*)(ws + p.mask_bytes);
    const long long rows = (long long)p.m_tiles * SC_BM;
    const long long nmask = rows * p.n_words;
    score_mask_base_kernel<<<(int)((nmask + 255) / 256), 256, 0, stream>>>(mask, rows, p.n_words, N, mask_col0);
    PR_CUDA_LAUNCH_CHECK("score_mask_base_kernel");
    CUtensorMap tmA;
    int rc = make_map(&tmA, seq_out, B_e, D, SC_BM);
    if (rc) return rc;
    ScoreArgs a;
    a.kblocks = (int)(D / SC_BK);
    a.m_tiles = p.m_tiles; a.n_tiles = p.n_tiles; a.tiles_per_split = p.tiles_per_split; a.n_splits = p.n_splits;
    a.n_words = p.n_words; a.mask = mask; a.cand_val = nullptr; a.cand_idx = nullptr; a.cluster = 1;
    a.target = (const long long*)target; a.n_rows = B_e; a.ce_part = part;
    int n_lists = 0;
    rc = launch_v2<16, 1>(tmA, W, N, D, a, p, (tune() & PR_TUNE_SCORE_MCAST) != 0, stream, &n_lists);
    if (rc) return rc;
    score_ce_merge_kernel<<<(int)((B_e + 3) / 4), 128, 0, stream>>>(part, n_lists, B_e, lse, tgt_logit, nll);
    PR_CUDA_LAUNCH_CHECK("score_ce_merge_kernel");
    return PR_OK;
}

// ---- fp16-operand scoring (same mantissa width as TF32, twice the MMA rate, half the operand bytes) -----------------------
extern "C" int pr_score_prepare_f16(const float* src, int64_t n, void* dst_f16, int32_t* status, pr_stream_t stream_) {
    PR_CHECK_ARG(n >= 0 && n % 4 == 0, "pr_score_prepare_f16: n=%lld must be a non-negative multiple of 4", (long long)n);
    if (n == 0) return PR_OK;
    PR_CHECK_ARG(src && dst_f16 && aligned16(src) && (reinterpret_cast<uintptr_t>(dst_f16) & 7u) == 0,
                 "pr_score_prepare_f16: null or unaligned pointer");
    const long long n4 = n / 4;
    const int grid = (int)std::max<long long>(1, std::min<long long>((n4 + 255) / 256, (long long)sm_count() * 8));
    score_to_f16_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>((const float4*)src, n4, (uint2*)dst_f16, status);
    PR_CUDA_LAUNCH_CHECK("score_to_f16_kernel");
    return PR_OK;
}

static size_t f16_seq_bytes(int64_t B_e, int64_t D) { return ((size_t)B_e * D * 2 + 255) / 256 * 256; }

extern "C" size_t pr_score_topk_f16_workspace_bytes(int64_t B_e, int64_t N, int64_t D, int k) {
    if (B_e <= 0 || N <= 0 || D <= 0 || k <= 0 || k > 32) return 0;
    return score_plan(B_e, N, k).total + f16_seq_bytes(B_e, D);
}

extern "C" int pr_score_topk_f16(const float* seq_out, int64_t B_e, const void* W16, int64_t N, int64_t D, const int64_t* hist_u,
                                 const int64_t* hist_i, int64_t n_hist, int mask_col0, int k, float* topk_val, int64_t* topk_idx,
                                 void* workspace, size_t workspace_bytes, int32_t* status, pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PR_CHECK_ARG(B_e > 0 && N > 0 && D > 0, "pr_score_topk_f16: bad shape B_e=%lld N=%lld D=%lld", (long long)B_e, (long long)N,
                 (long long)D);
    PR_CHECK_ARG(D % 64 == 0, "pr_score_topk_f16: D=%lld must be a multiple of 64", (long long)D);
    PR_CHECK_ARG(k >= 1 && k <= 32 && k <= N, "pr_score_topk_f16: k=%d outside [1, min(32, N)]", k);
    PR_CHECK_ARG(N < (1LL << 31) - 512, "pr_score_topk_f16: N too large");
    PR_CHECK_ARG(seq_out && W16 && topk_val && topk_idx && workspace, "pr_score_topk_f16: null pointer");
    PR_CHECK_ARG(aligned16(seq_out) && aligned16(W16), "pr_score_topk_f16: seq_out / W16 must be 16-byte aligned");
    PR_CHECK_ARG(n_hist >= 0 && (n_hist == 0 || (hist_u && hist_i)), "pr_score_topk_f16: bad history arguments");
    const ScorePlan p = score_plan(B_e, N, k);
    PR_CHECK_ARG(workspace_bytes >= p.total + f16_seq_bytes(B_e, D), "pr_score_topk_f16: workspace %zu < required %zu",
                 workspace_bytes, p.total + f16_seq_bytes(B_e, D));
    PR_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "pr_score_topk_f16: workspace must be 256-byte aligned");
    char* ws = (char*)workspace;
    uint32_t* mask = (uint32_t*)ws;
    float* cand_val = (float*)(ws + p.mask_bytes);
    int* cand_idx = (int*)(ws + p.mask_bytes + p.cand_bytes);
    void* seq16 = ws + p.total;
    int rc = pr_score_prepare_f16(seq_out, B_e * D, seq16, status, stream_);
    if (rc) return rc;
    const long long rows = (long long)p.m_tiles * SC_BM;
    const long long nmask = rows * p.n_words;
    score_mask_base_kernel<<<(int)((nmask + 255) / 256), 256, 0, stream>>>(mask, rows, p.n_words, N, mask_col0);
    if (n_hist > 0)
        score_mask_hist_kernel<<<(int)((n_hist + 255) / 256), 256, 0, stream>>>(mask, p.n_words, B_e, N, (const long long*)hist_u,
                                                                                (const long long*)hist_i, n_hist);
    PR_CUDA_LAUNCH_CHECK("score_mask kernels");
    CUtensorMap tmA;
    rc = make_map_f16(&tmA, seq16, B_e, D, SC_BM);
    if (rc) return rc;
    ScoreArgs a;
    a.kblocks = (int)(D / 64);
    a.m_tiles = p.m_tiles; a.n_tiles = p.n_tiles; a.tiles_per_split = p.tiles_per_split; a.n_splits = p.n_splits;
    a.n_words = p.n_words; a.mask = mask; a.cand_val = cand_val; a.cand_idx = cand_idx; a.cluster = 1;
    a.target = nullptr; a.n_rows = B_e; a.ce_part = nullptr;
    int n_lists = 0;
    const bool mcast = (tune() & PR_TUNE_SCORE_MCAST) != 0;
    if ((tune() & PR_TUNE_SCORE_ARES) && D <= 512)       // seq_out tile resident in shared memory (<= 128 KiB)
        rc = (p.K == 16) ? launch_v2<16, 0, true, true>(tmA, W16, N, D, a, p, mcast, stream, &n_lists)
                         : launch_v2<32, 0, true, true>(tmA, W16, N, D, a, p, mcast, stream, &n_lists);
    else
        rc = (p.K == 16) ? launch_v2<16, 0, true>(tmA, W16, N, D, a, p, mcast, stream, &n_lists)
                         : launch_v2<32, 0, true>(tmA, W16, N, D, a, p, mcast, stream, &n_lists);
    if (rc) return rc;
    score_merge_kernel<<<(int)((B_e + 3) / 4), 128, 0, stream>>>(cand_val, cand_idx, n_lists * p.K, B_e, k, topk_val,
                                                                 (long long*)topk_idx);
    PR_CUDA_LAUNCH_CHECK("score_merge_kernel");
    return PR_OK;
}
