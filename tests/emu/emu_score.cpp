// Host build of csrc/score_kernels.cuh (v2 scoring kernel, candidate merge, mask kernels) on the emulated Blackwell pipeline
// (emu_tc.h).  emu_score_topk_v2 mirrors the launch sequence of pr_score_topk_f32 with an explicit split count and cluster
// size.  Loaded by tests/test_emu_kernels.py through ctypes.  Test infrastructure only.
#include "emu_tc.h"

#include "../../pixelrec_b200/csrc/score_kernels.cuh"

using namespace pr;

template <int K>
static void run_v2(const CUtensorMap& tmA, const CUtensorMap& tmB, const ScoreArgs& a, int grid, int cluster) {
    const size_t smem = (size_t)SC_STAGES * SC_STAGE_BYTES + SC2_BAR_BYTES + 1024;
    emu::after_launch_hook() = emu::join_async;
    emu::launch_cluster(grid, cluster, SC2_THREADS, smem, [&]() { score_topk2_kernel<K>(tmA, tmB, a); });
    emu::after_launch_hook() = nullptr;
}

extern "C" int emu_score_topk_v2(const float* seq, long long B_e, const float* W, long long N, long long D, const long long* hist_u,
                                 const long long* hist_i, long long n_hist, int mask_col0, int k, int splits_req, int cluster,
                                 float* out_val, long long* out_idx) {
    const int K = (k <= 16) ? 16 : 32;
    ScoreArgs a{};
    a.kblocks = (int)(D / SC_BK);
    a.m_tiles = (int)((B_e + SC_BM - 1) / SC_BM);
    a.n_tiles = (int)((N + SC_BN - 1) / SC_BN);
    if (a.m_tiles % cluster) return -1;
    a.tiles_per_split = (a.n_tiles + splits_req - 1) / splits_req;
    a.n_splits = (a.n_tiles + a.tiles_per_split - 1) / a.tiles_per_split;
    a.n_words = a.n_tiles * 8;
    a.cluster = cluster;
    const long long rows = (long long)a.m_tiles * SC_BM;
    std::vector<uint32_t> mask((size_t)rows * a.n_words);
    const int n_lists = a.n_splits * 2;
    std::vector<float> cand_val((size_t)rows * n_lists * K, -1234.5f);
    std::vector<int> cand_idx((size_t)rows * n_lists * K, -77);
    a.mask = mask.data();
    a.cand_val = cand_val.data();
    a.cand_idx = cand_idx.data();
    const long long nmask = rows * a.n_words;
    emu::launch((int)((nmask + 255) / 256), 256, 0, [&]() { score_mask_base_kernel(mask.data(), rows, a.n_words, N, mask_col0); });
    if (n_hist > 0)
        emu::launch((int)((n_hist + 255) / 256), 256, 0,
                    [&]() { score_mask_hist_kernel(mask.data(), a.n_words, B_e, N, hist_u, hist_i, n_hist); });
    const CUtensorMap tmA{seq, B_e, D, SC_BM, 4}, tmB{W, N, D, SC_BN / cluster, 4};
    if (K == 16) run_v2<16>(tmA, tmB, a, a.m_tiles * a.n_splits, cluster);
    else run_v2<32>(tmA, tmB, a, a.m_tiles * a.n_splits, cluster);
    emu::launch((int)((B_e + 3) / 4), 128, 0, [&]() {
        score_merge_kernel(cand_val.data(), cand_idx.data(), n_lists * K, B_e, k, out_val, out_idx);
    });
    return a.n_splits;
}

extern "C" int emu_score_ce_v2(const float* seq, long long B_e, const float* W, long long N, long long D, const long long* target,
                               int mask_col0, int splits_req, int cluster, float* lse, float* tgt, float* nll) {
    ScoreArgs a{};
    a.kblocks = (int)(D / SC_BK);
    a.m_tiles = (int)((B_e + SC_BM - 1) / SC_BM);
    a.n_tiles = (int)((N + SC_BN - 1) / SC_BN);
    if (a.m_tiles % cluster) return -1;
    a.tiles_per_split = (a.n_tiles + splits_req - 1) / splits_req;
    a.n_splits = (a.n_tiles + a.tiles_per_split - 1) / a.tiles_per_split;
    a.n_words = a.n_tiles * 8;
    a.cluster = cluster;
    const long long rows = (long long)a.m_tiles * SC_BM;
    std::vector<uint32_t> mask((size_t)rows * a.n_words);
    const int n_lists = a.n_splits * 2;
    std::vector<float> part((size_t)rows * n_lists * 4, -99.f);
    a.mask = mask.data();
    a.target = target;
    a.n_rows = B_e;
    a.ce_part = part.data();
    const long long nmask = rows * a.n_words;
    emu::launch((int)((nmask + 255) / 256), 256, 0, [&]() { score_mask_base_kernel(mask.data(), rows, a.n_words, N, mask_col0); });
    const CUtensorMap tmA{seq, B_e, D, SC_BM, 4}, tmB{W, N, D, SC_BN / cluster, 4};
    const size_t smem = (size_t)SC_STAGES * SC_STAGE_BYTES + SC2_BAR_BYTES + 1024;
    emu::after_launch_hook() = emu::join_async;
    emu::launch_cluster(a.m_tiles * a.n_splits, cluster, SC2_THREADS, smem, [&]() { score_topk2_kernel<16, 1>(tmA, tmB, a); });
    emu::after_launch_hook() = nullptr;
    emu::launch((int)((B_e + 3) / 4), 128, 0, [&]() { score_ce_merge_kernel(part.data(), n_lists, B_e, lse, tgt, nll); });
    return a.n_splits;
}

// fp16-operand variant: mirrors pr_score_prepare_f16 + pr_score_topk_f16
extern "C" int emu_score_topk_f16(const float* seq, long long B_e, const float* W, long long N, long long D, const long long* hist_u,
                                  const long long* hist_i, long long n_hist, int mask_col0, int k, int splits_req, int cluster,
                                  float* out_val, long long* out_idx, int* status, int ares) {
    const int K = (k <= 16) ? 16 : 32;
    std::vector<uint16_t> seq16((size_t)B_e * D), W16((size_t)N * D);
    emu::launch(4, 256, 0, [&]() { score_to_f16_kernel((const float4*)seq, B_e * D / 4, (uint2*)seq16.data(), status); });
    emu::launch(4, 256, 0, [&]() { score_to_f16_kernel((const float4*)W, N * D / 4, (uint2*)W16.data(), status); });
    ScoreArgs a{};
    a.kblocks = (int)(D / 64);
    a.m_tiles = (int)((B_e + SC_BM - 1) / SC_BM);
    a.n_tiles = (int)((N + SC_BN - 1) / SC_BN);
    if (a.m_tiles % cluster) return -1;
    a.tiles_per_split = (a.n_tiles + splits_req - 1) / splits_req;
    a.n_splits = (a.n_tiles + a.tiles_per_split - 1) / a.tiles_per_split;
    a.n_words = a.n_tiles * 8;
    a.cluster = cluster;
    a.n_rows = B_e;
    const long long rows = (long long)a.m_tiles * SC_BM;
    std::vector<uint32_t> mask((size_t)rows * a.n_words);
    const int n_lists = a.n_splits * 2;
    std::vector<float> cand_val((size_t)rows * n_lists * K, -1234.5f);
    std::vector<int> cand_idx((size_t)rows * n_lists * K, -77);
    a.mask = mask.data();
    a.cand_val = cand_val.data();
    a.cand_idx = cand_idx.data();
    const long long nmask = rows * a.n_words;
    emu::launch((int)((nmask + 255) / 256), 256, 0, [&]() { score_mask_base_kernel(mask.data(), rows, a.n_words, N, mask_col0); });
    if (n_hist > 0)
        emu::launch((int)((n_hist + 255) / 256), 256, 0,
                    [&]() { score_mask_hist_kernel(mask.data(), a.n_words, B_e, N, hist_u, hist_i, n_hist); });
    const CUtensorMap tmA{seq16.data(), B_e, D, SC_BM, 2}, tmB{W16.data(), N, D, SC_BN / cluster, 2};
    const size_t smem = (ares ? (size_t)a.kblocks * SC_A_BYTES + (size_t)3 * SC_B_BYTES : (size_t)SC_STAGES * SC_STAGE_BYTES) +
                        SC2_BAR_BYTES + 1024;
    if (ares && D > 512) return -2;
    emu::after_launch_hook() = emu::join_async;
    const int grid = a.m_tiles * a.n_splits;
    if (ares) {
        if (K == 16) emu::launch_cluster(grid, cluster, SC2_THREADS, smem, [&]() { score_topk2_kernel<16, 0, true, true>(tmA, tmB, a); });
        else emu::launch_cluster(grid, cluster, SC2_THREADS, smem, [&]() { score_topk2_kernel<32, 0, true, true>(tmA, tmB, a); });
    } else {
        if (K == 16) emu::launch_cluster(grid, cluster, SC2_THREADS, smem, [&]() { score_topk2_kernel<16, 0, true>(tmA, tmB, a); });
        else emu::launch_cluster(grid, cluster, SC2_THREADS, smem, [&]() { score_topk2_kernel<32, 0, true>(tmA, tmB, a); });
    }
    emu::after_launch_hook() = nullptr;
    emu::launch((int)((B_e + 3) / 4), 128, 0, [&]() {
        score_merge_kernel(cand_val.data(), cand_idx.data(), n_lists * K, B_e, k, out_val, out_idx);
    });
    return a.n_splits;
}

extern "C" int emu_f32_to_f16(const float* src, long long n, uint16_t* dst) {
    bool sat = false;
    for (long long i = 0; i < n; ++i) dst[i] = f32_to_f16_rn(src[i], &sat);
    return sat ? 1 : 0;
}
