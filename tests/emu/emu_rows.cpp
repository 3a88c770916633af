// Host build of csrc/rows_ring.cuh (TMA-staged segment reduce of the table gradient) on the emulated mbarrier / bulk-copy
// layer (emu_tc.h).  Loaded by tests/test_emu_kernels.py through ctypes.  Test infrastructure only.
#include "emu_tc.h"

#include "../../pixelrec_b200/csrc/rows_ring.cuh"
#include "../../pixelrec_b200/csrc/rows_plan.cuh"

using namespace pr;

template <int VPL, int RPS>
static void run_ring2(const float* dOut, int D, int gr, const int* perm, const int* uniq_ids, const int* seg_start, const int* n_uniq,
                      long long max_uniq, float scale, float* out_rows, float* dense_G, int grid, int nst) {
    const size_t smem = (size_t)nst * RPS * D * 4 + SR_BAR_BYTES + SR_Q_BYTES;
    emu::after_launch_hook() = emu::join_async;
    emu::launch(grid, 32, smem, [&]() {
        scatter_add_rows_ring2_kernel<VPL, RPS>(dOut, D, gr, perm, uniq_ids, seg_start, n_uniq, max_uniq, scale, out_rows, dense_G, nst);
    });
    emu::after_launch_hook() = nullptr;
}

template <int VPL, int RPS>
static void run_ring(const float* dOut, int D, int gr, const int* perm, const int* uniq_ids, const int* seg_start, const int* n_uniq,
                     long long max_uniq, float scale, float* out_rows, float* dense_G, int grid, int nst) {
    const size_t smem = (size_t)nst * RPS * D * 4 + SR_BAR_BYTES;
    emu::after_launch_hook() = emu::join_async;
    emu::launch(grid, 32, smem, [&]() {
        scatter_add_rows_ring_kernel<VPL, RPS>(dOut, D, gr, perm, uniq_ids, seg_start, n_uniq, max_uniq, scale, out_rows, dense_G, nst);
    });
    emu::after_launch_hook() = nullptr;
}

// same VPL / RPS choice as pr_scatter_add_rows_f32 (rows.cu); gr, grid and the ring depth are explicit so that tests can force
// group switches and both ring depths; big != 0: the 4-rows-per-stage A/B geometry at VPL = 4
extern "C" int emu_scatter_add_rows_ring(const float* dOut, int D, int gr, const int* perm, const int* uniq_ids, const int* seg_start,
                                         const int* n_uniq, long long max_uniq, float scale, float* out_rows, float* dense_G,
                                         int grid, int nst, int big) {
    int vpl = 1;
    while (32 * vpl < D / 4) vpl *= 2;
#define RUN(V, R) run_ring<V, R>(dOut, D, gr, perm, uniq_ids, seg_start, n_uniq, max_uniq, scale, out_rows, dense_G, grid, nst)
#define RUN2(V, R) run_ring2<V, R>(dOut, D, gr, perm, uniq_ids, seg_start, n_uniq, max_uniq, scale, out_rows, dense_G, grid, nst)
    const bool v2 = big == 2;       // big: 0 = ring v1, 1 = ring v1 with 4 rows per stage, 2 = ring v2 (look-ahead queue in shared memory)
    switch (vpl) {
        case 1: if (v2) RUN2(1, 8); else RUN(1, 8); break;
        case 2: if (v2) RUN2(2, 4); else RUN(2, 4); break;
        case 4: if (v2) RUN2(4, 2); else if (big) RUN(4, 4); else RUN(4, 2); break;
        case 8: if (v2) RUN2(8, 1); else RUN(8, 1); break;
        default: return -1;
    }
#undef RUN
#undef RUN2
    return vpl;
}

// the launch sequence of pr_scatter_plan (rows.cu) on the emulated kernels; `bits_per_digit` = 0 picks the production digits
extern "C" int emu_scatter_plan(const long long* idx, int R, long long N, long long pad, int* perm, int* uniq_ids, int* seg_start,
                                int* n_uniq, int* row2slot, int* status, int bits_per_digit) {
    if (R == 0) {
        emu::launch(1, 1, 0, [&]() { plan_empty_kernel(seg_start, n_uniq); });
        return 0;
    }
    const int T = (R + RS_TILE - 1) / RS_TILE;
    std::vector<uint32_t> k0(R), k1(R), tile_hist((size_t)RS_MAX_BINS * T), tile_sum(T + 1);
    std::vector<int> tmpv(R);
    uint32_t* kbuf[2] = {k0.data(), k1.data()};
    int bits = 1;
    while ((1LL << bits) <= N) ++bits;
    int passes = (bits + 9) / 10;
    int rb = (bits + passes - 1) / passes;
    if (bits_per_digit) { rb = bits_per_digit; passes = (bits + rb - 1) / rb; }
    const int bins = 1 << rb;
    int* vbuf[2];
    if (passes % 2 == 0) { vbuf[0] = perm; vbuf[1] = tmpv.data(); } else { vbuf[0] = tmpv.data(); vbuf[1] = perm; }
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        const int shift = rb * p;
        emu::launch(T, RS_THREADS, 0, [&]() {
            rs_hist_kernel(p == 0 ? idx : nullptr, N, pad, status, kbuf[cur], R, shift, bins, tile_hist.data(), T);
        });
        emu::launch(1, SCAN_THREADS, 0, [&]() { scan_single_kernel(tile_hist.data(), bins * T, nullptr); });
        emu::launch(T, RS_THREADS, 0, [&]() {
            rs_scatter_kernel(kbuf[cur], p == 0 ? nullptr : vbuf[cur], kbuf[cur ^ 1], vbuf[cur ^ 1], R, shift, bins, tile_hist.data(), T);
        });
        cur ^= 1;
    }
    const uint32_t* skeys = kbuf[cur];
    emu::launch(T, RS_THREADS, 0, [&]() { seg_reduce_kernel(skeys, R, (uint32_t)N, tile_sum.data()); });
    emu::launch(1, SCAN_THREADS, 0, [&]() { scan_single_kernel(tile_sum.data(), T, nullptr); });
    emu::launch(T, RS_THREADS, 0, [&]() { seg_emit_kernel(skeys, R, (uint32_t)N, tile_sum.data(), uniq_ids, seg_start, n_uniq, row2slot); });
    return passes;
}
