// Host build of csrc/rows_ring.cuh (TMA-staged segment reduce of the table gradient) on the emulated mbarrier / bulk-copy
// layer (emu_tc.h).  Loaded by tests/test_emu_kernels.py through ctypes.  Test infrastructure only.
#include "emu_tc.h"

#include "../../pixelrec_b200/csrc/rows_ring.cuh"

using namespace pr;

template <int VPL, int RPS>
static void run_ring(const float* dOut, int D, int gr, const int* perm, const int* uniq_ids, const int* seg_start, const int* n_uniq,
                     long long max_uniq, float scale, float* out_rows, float* dense_G, int grid) {
    const size_t smem = (size_t)SR_STAGES * RPS * D * 4 + SR_BAR_BYTES;
    emu::after_launch_hook() = emu::join_async;
    emu::launch(grid, 32, smem, [&]() {
        scatter_add_rows_ring_kernel<VPL, RPS>(dOut, D, gr, perm, uniq_ids, seg_start, n_uniq, max_uniq, scale, out_rows, dense_G, 1);
    });
    emu::after_launch_hook() = nullptr;
}

// same VPL / RPS choice as pr_scatter_add_rows_f32 (rows.cu); gr and grid are explicit so that tests can force group switches
extern "C" int emu_scatter_add_rows_ring(const float* dOut, int D, int gr, const int* perm, const int* uniq_ids, const int* seg_start,
                                         const int* n_uniq, long long max_uniq, float scale, float* out_rows, float* dense_G,
                                         int grid) {
    int vpl = 1;
    while (32 * vpl < D / 4) vpl *= 2;
    switch (vpl) {
        case 1: run_ring<1, 16>(dOut, D, gr, perm, uniq_ids, seg_start, n_uniq, max_uniq, scale, out_rows, dense_G, grid); break;
        case 2: run_ring<2, 8>(dOut, D, gr, perm, uniq_ids, seg_start, n_uniq, max_uniq, scale, out_rows, dense_G, grid); break;
        case 4: run_ring<4, 4>(dOut, D, gr, perm, uniq_ids, seg_start, n_uniq, max_uniq, scale, out_rows, dense_G, grid); break;
        case 8: run_ring<8, 2>(dOut, D, gr, perm, uniq_ids, seg_start, n_uniq, max_uniq, scale, out_rows, dense_G, grid); break;
        case 16: run_ring<16, 1>(dOut, D, gr, perm, uniq_ids, seg_start, n_uniq, max_uniq, scale, out_rows, dense_G, grid); break;
        default: return -1;
    }
    return vpl;
}
