// pixelrec_b200 -- K6b: attention core for long sequences (64 < L <= 256), host side of attn_long.cuh.
//   ViT-B/16 item encoder (197 tokens, 12 heads, dh 64): REC/model/load.py:90-99 -> HF CLIPVisionModel encoder layers.
#include <algorithm>
#include <math.h>

#include "common.cuh"

#define PR_LDG4(p) __ldg(p)
#define PR_DYN_SMEM_F4(name) extern __shared__ __align__(16) float4 name[]
#include "attn_long.cuh"

namespace pr {
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// D = A(16x8, row) * B(8x8, col) + D, TF32 operands, fp32 accumulate
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
}  // namespace pr
#include "attn_long_tc.cuh"

using namespace pr;

namespace {

int check_long(const char* who, const float* q, const float* k, const float* v, long long ld, int B, int L, int h, int dh) {
    PR_CHECK_ARG(B > 0 && L > 0 && h > 0 && dh > 0, "%s: bad shape B=%d L=%d h=%d dh=%d", who, B, L, h, dh);
    PR_CHECK_ARG(L <= 32 * AL_JJ, "%s: L=%d > %d unsupported", who, L, 32 * AL_JJ);
    PR_CHECK_ARG(dh == 4 || dh == 8 || dh == 16 || dh == 32 || dh == 64 || dh == 128, "%s: dh=%d must be a power of two in [4,128]",
                 who, dh);
    PR_CHECK_ARG(ld % 4 == 0 && ld >= (long long)h * dh, "%s: ld=%lld must be a multiple of 4 and >= h*dh", who, ld);
    PR_CHECK_ARG(q && k && v && aligned16(q) && aligned16(k) && aligned16(v), "%s: q/k/v null or not 16-byte aligned", who);
    PR_CHECK_ARG(long_smem_float4(L, dh) * 16 <= 220 * 1024, "%s: L=%d dh=%d does not fit shared memory", who, L, dh);
    return PR_OK;
}

template <int DH>
int launch_long_tc(const LongAttnArgs& A, cudaStream_t stream) {
    const size_t smem = long_tc_smem_floats<DH>(A.L) * 4;
    if (smem > 220 * 1024) return PR_ERR_UNSUPPORTED;      // caller falls back to the fp32 kernel
    PR_CUDA_CALL(cudaFuncSetAttribute(attn_long_tc_fwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long n_items = (long long)A.B * A.h;
    const int grid = (int)std::max<long long>(1, std::min<long long>(n_items, (long long)sm_count()));
    attn_long_tc_fwd_kernel<DH><<<grid, ALT_THREADS, smem, stream>>>(A);
    PR_CUDA_LAUNCH_CHECK("attn_long_tc_fwd_kernel");
    return PR_OK;
}

template <int DH>
int launch_long_tc_bwd(const LongAttnArgs& A, cudaStream_t stream) {
    const size_t smem = long_tc_bwd_smem_floats<DH>(A.L) * 4;
    if (smem > 220 * 1024) return PR_ERR_UNSUPPORTED;
    PR_CUDA_CALL(cudaFuncSetAttribute(attn_long_tc_bwd_dq_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PR_CUDA_CALL(cudaFuncSetAttribute(attn_long_tc_bwd_dkv_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long n_items = (long long)A.B * A.h;
    const int grid = (int)std::max<long long>(1, std::min<long long>(n_items, (long long)sm_count()));
    attn_long_tc_bwd_dq_kernel<DH><<<grid, ALT_THREADS, smem, stream>>>(A);
    PR_CUDA_LAUNCH_CHECK("attn_long_tc_bwd_dq_kernel");
    attn_long_tc_bwd_dkv_kernel<DH><<<grid, ALT_THREADS, smem, stream>>>(A);
    PR_CUDA_LAUNCH_CHECK("attn_long_tc_bwd_dkv_kernel");
    return PR_OK;
}

template <typename Kern>
int launch_long(Kern kern, const char* name, const LongAttnArgs& A, cudaStream_t stream) {
    const size_t smem = long_smem_float4(A.L, A.dh) * 16;
    PR_CUDA_CALL(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long n_items = (long long)A.B * A.h;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(2, (220 * 1024) / (smem + 1024)));
    const int grid = (int)std::max<long long>(1, std::min<long long>(n_items, (long long)sm_count() * per_sm));
    kern<<<grid, AL_THREADS, smem, stream>>>(A);
    PR_CUDA_LAUNCH_CHECK(name);
    return PR_OK;
}

}  // namespace

extern "C" int pr_attn_long_fwd_f32(const float* q, const float* k, const float* v, int64_t ld, const int64_t* key_ids, int B,
                                    int L, int h, int dh, int causal, float* ctx, float* lse, pr_stream_t stream_) {
    int rc = check_long("pr_attn_long_fwd_f32", q, k, v, ld, B, L, h, dh);
    if (rc) return rc;
    PR_CHECK_ARG(ctx && lse && aligned16(ctx), "pr_attn_long_fwd_f32: ctx/lse null or unaligned");
    LongAttnArgs A{};
    A.q = q; A.k = k; A.v = v; A.ld = ld; A.key_ids = (const long long*)key_ids;
    A.B = B; A.L = L; A.h = h; A.dh = dh; A.causal = causal;
    A.scale = (float)(1.0 / sqrt((double)dh));
    A.ctx = ctx; A.lse = lse;
    if (tune() & PR_TUNE_ATTN_LONG_TC) {                   // tensor-core forward (TF32); same outputs, the backward is unchanged
        rc = PR_ERR_UNSUPPORTED;
        if (dh == 32) rc = launch_long_tc<32>(A, (cudaStream_t)stream_);
        else if (dh == 64) rc = launch_long_tc<64>(A, (cudaStream_t)stream_);
        else if (dh == 128) rc = launch_long_tc<128>(A, (cudaStream_t)stream_);
        if (rc != PR_ERR_UNSUPPORTED) return rc;
    }
    return launch_long(attn_long_fwd_kernel, "attn_long_fwd_kernel", A, (cudaStream_t)stream_);
}

extern "C" int pr_attn_long_bwd_f32(const float* q, const float* k, const float* v, int64_t ld, const int64_t* key_ids,
                                    const float* ctx, const float* lse, const float* dctx, int B, int L, int h, int dh,
                                    int causal, float* dq, float* dk, float* dv, int64_t ld_grad, float* delta_ws,
                                    pr_stream_t stream_) {
    int rc = check_long("pr_attn_long_bwd_f32", q, k, v, ld, B, L, h, dh);
    if (rc) return rc;
    PR_CHECK_ARG(ctx && lse && dctx && dq && dk && dv && delta_ws, "pr_attn_long_bwd_f32: null pointer");
    PR_CHECK_ARG(aligned16(ctx) && aligned16(dctx) && aligned16(dq) && aligned16(dk) && aligned16(dv),
                 "pr_attn_long_bwd_f32: unaligned pointer");
    PR_CHECK_ARG(ld_grad % 4 == 0 && ld_grad >= (int64_t)h * dh, "pr_attn_long_bwd_f32: bad ld_grad");
    LongAttnArgs A{};
    A.q = q; A.k = k; A.v = v; A.ld = ld; A.key_ids = (const long long*)key_ids;
    A.B = B; A.L = L; A.h = h; A.dh = dh; A.causal = causal;
    A.scale = (float)(1.0 / sqrt((double)dh));
    A.lse = const_cast<float*>(lse); A.ctx_in = ctx; A.dctx = dctx; A.delta = delta_ws;
    A.dq = dq; A.dk = dk; A.dv = dv; A.ld_grad = ld_grad;
    if (tune() & PR_TUNE_ATTN_LONG_TC) {
        rc = PR_ERR_UNSUPPORTED;
        if (dh == 32) rc = launch_long_tc_bwd<32>(A, (cudaStream_t)stream_);
        else if (dh == 64) rc = launch_long_tc_bwd<64>(A, (cudaStream_t)stream_);      // dh = 128 would spill: fp32 kernels
        if (rc != PR_ERR_UNSUPPORTED) return rc;
    }
    rc = launch_long(attn_long_bwd_dq_kernel, "attn_long_bwd_dq_kernel", A, (cudaStream_t)stream_);
    if (rc) return rc;
    return launch_long(attn_long_bwd_dkv_kernel, "attn_long_bwd_dkv_kernel", A, (cudaStream_t)stream_);
}
