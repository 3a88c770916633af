// Host build of the SIMT kernels in pixelrec_b200/csrc/{attn_long,peer}.cuh on the emulation layer (emu_cuda.h).
// Loaded by tests/test_emu_kernels.py through ctypes.  Test infrastructure only.
#include "emu_cuda.h"
#include "emu_mma.h"

#include "../../pixelrec_b200/csrc/attn_long.cuh"
#include "../../pixelrec_b200/csrc/attn_long_tc.cuh"
#include "../../pixelrec_b200/csrc/peer.cuh"

using namespace pr;

static LongAttnArgs make_args(const float* q, const float* k, const float* v, long long ld, const long long* key_ids, int B,
                              int L, int h, int dh, int causal) {
    LongAttnArgs A{};
    A.q = q; A.k = k; A.v = v; A.ld = ld; A.key_ids = key_ids;
    A.B = B; A.L = L; A.h = h; A.dh = dh; A.causal = causal;
    A.scale = (float)(1.0 / std::sqrt((double)dh));
    return A;
}

extern "C" void emu_attn_long_fwd(const float* q, const float* k, const float* v, long long ld, const long long* key_ids, int B,
                                  int L, int h, int dh, int causal, float* ctx, float* lse, int grid) {
    LongAttnArgs A = make_args(q, k, v, ld, key_ids, B, L, h, dh, causal);
    A.ctx = ctx; A.lse = lse;
    emu::launch(grid, AL_THREADS, long_smem_float4(L, dh) * 16, [&]() { attn_long_fwd_kernel(A); });
}

extern "C" int emu_attn_long_tc_fwd(const float* q, const float* k, const float* v, long long ld, const long long* key_ids, int B,
                                   int L, int h, int dh, int causal, float* ctx, float* lse, int grid) {
    LongAttnArgs A = make_args(q, k, v, ld, key_ids, B, L, h, dh, causal);
    A.ctx = ctx; A.lse = lse;
    if (dh == 32) emu::launch(grid, ALT_THREADS, long_tc_smem_floats<32>(L) * 4, [&]() { attn_long_tc_fwd_kernel<32>(A); });
    else if (dh == 64) emu::launch(grid, ALT_THREADS, long_tc_smem_floats<64>(L) * 4, [&]() { attn_long_tc_fwd_kernel<64>(A); });
    else if (dh == 128) emu::launch(grid, ALT_THREADS, long_tc_smem_floats<128>(L) * 4, [&]() { attn_long_tc_fwd_kernel<128>(A); });
    else return -1;
    return 0;
}

template <int DH>
static void run_tc_bwd(const LongAttnArgs& A, int grid) {
    emu::launch(grid, ALT_THREADS, long_tc_bwd_smem_floats<DH>(A.L) * 4, [&]() { attn_long_tc_bwd_dq_kernel<DH>(A); });
    emu::launch(grid, ALT_THREADS, long_tc_bwd_smem_floats<DH>(A.L) * 4, [&]() { attn_long_tc_bwd_dkv_kernel<DH>(A); });
}

extern "C" int emu_attn_long_tc_bwd(const float* q, const float* k, const float* v, long long ld, const long long* key_ids,
                                    const float* ctx, const float* lse, const float* dctx, int B, int L, int h, int dh, int causal,
                                    float* dq, float* dk, float* dv, long long ld_grad, float* delta, int grid) {
    LongAttnArgs A = make_args(q, k, v, ld, key_ids, B, L, h, dh, causal);
    A.lse = const_cast<float*>(lse); A.ctx_in = ctx; A.dctx = dctx; A.delta = delta;
    A.dq = dq; A.dk = dk; A.dv = dv; A.ld_grad = ld_grad;
    if (dh == 32) run_tc_bwd<32>(A, grid);
    else if (dh == 64) run_tc_bwd<64>(A, grid);
    else return -1;
    return 0;
}

extern "C" void emu_attn_long_bwd(const float* q, const float* k, const float* v, long long ld, const long long* key_ids,
                                  const float* ctx, const float* lse, const float* dctx, int B, int L, int h, int dh, int causal,
                                  float* dq, float* dk, float* dv, long long ld_grad, float* delta, int grid) {
    LongAttnArgs A = make_args(q, k, v, ld, key_ids, B, L, h, dh, causal);
    A.lse = const_cast<float*>(lse); A.ctx_in = ctx; A.dctx = dctx; A.delta = delta;
    A.dq = dq; A.dk = dk; A.dv = dv; A.ld_grad = ld_grad;
    emu::launch(grid, AL_THREADS, long_smem_float4(L, dh) * 16, [&]() { attn_long_bwd_dq_kernel(A); });
    emu::launch(grid, AL_THREADS, long_smem_float4(L, dh) * 16, [&]() { attn_long_bwd_dkv_kernel(A); });
}

extern "C" void emu_gather_rows_peers(const float* const* shards, int G, long long N, int D, const long long* idx, long long R,
                                      float* out, int* status, int grid) {
    emu::launch(grid, 256, 0, [&]() {
        gather_rows_peers_kernel((const float4* const*)shards, G, N, D / 4, idx, R, (float4*)out, status);
    });
}

extern "C" void emu_push_rows_peers(const float* rows, const long long* ids, long long U, int D, int G, int rank, long long cap,
                                    long long skip_id, float* const* recv_rows, long long* const* recv_ids, int* counters,
                                    int* status, int grid) {
    emu::launch(grid, 256, 0, [&]() {
        push_rows_peers_kernel((const float4*)rows, ids, U, D / 4, G, rank, cap, skip_id, (float4* const*)recv_rows, recv_ids,
                               counters, status);
    });
}
