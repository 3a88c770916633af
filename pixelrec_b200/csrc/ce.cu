// pixelrec_b200 -- backward of the full-catalog softmax cross-entropy (extension; forward = pr_score_ce_f32 on the tcgen05
// scoring pipeline, csrc/score.cu).  The backward never holds the [B_e, N] probabilities either: the host loop walks the
// catalog in column chunks, recomputes a chunk of logits with pr_gemm_tf32, turns it IN PLACE into dS with the kernel below
// and feeds it to the two gradient GEMMs (dX += dS W_c, dW_c = dS^T X) on the same tensor-core kernel.
//   dS[r, j] = dnll[r] * (exp(S[r, j] - lse[r]) - [c0 + j == target[r]]),  0 for the masked padding column
#include <algorithm>

#include "common.cuh"

namespace pr {

__global__ void __launch_bounds__(256) ce_grad_chunk_kernel(float* __restrict__ S, long long ld, long long rows, int C, long long c0,
                                                            const float* __restrict__ lse, const long long* __restrict__ target,
                                                            const float* __restrict__ dnll, int mask_col0) {
    const int C4 = C >> 2;
    const long long n4 = rows * C4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const long long r = i / C4;
        const int j = (int)(i - r * C4) << 2;
        float4* p = reinterpret_cast<float4*>(S + r * ld + j);
        float4 s = *p;
        const float l = lse[r], g = dnll[r];
        const long long tj = target[r] - c0 - j;       // position of the target inside this float4, if 0..3
        float4 o;
        o.x = g * (__expf(s.x - l) - (tj == 0 ? 1.f : 0.f));
        o.y = g * (__expf(s.y - l) - (tj == 1 ? 1.f : 0.f));
        o.z = g * (__expf(s.z - l) - (tj == 2 ? 1.f : 0.f));
        o.w = g * (__expf(s.w - l) - (tj == 3 ? 1.f : 0.f));
        if (mask_col0 && c0 == 0 && j == 0) o.x = 0.f;
        *p = o;
    }
}

}  // namespace pr

using namespace pr;

extern "C" int pr_ce_grad_chunk_f32(float* S, int64_t ld, int64_t rows, int64_t C, int64_t c0, const float* lse,
                                    const int64_t* target, const float* dnll, int mask_col0, pr_stream_t stream_) {
    PR_CHECK_ARG(rows >= 0 && C > 0 && C % 4 == 0 && ld >= C && ld % 4 == 0 && c0 >= 0,
                 "pr_ce_grad_chunk_f32: bad shape rows=%lld C=%lld ld=%lld c0=%lld", (long long)rows, (long long)C, (long long)ld,
                 (long long)c0);
    if (rows == 0) return PR_OK;
    PR_CHECK_ARG(S && lse && target && dnll && aligned16(S), "pr_ce_grad_chunk_f32: null or unaligned pointer");
    PR_CHECK_ARG(C < (1LL << 31), "pr_ce_grad_chunk_f32: chunk too wide");
    const long long n4 = rows * (C / 4);
    const int grid = (int)std::max<long long>(1, std::min<long long>((n4 + 255) / 256, (long long)sm_count() * 8));
    ce_grad_chunk_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(S, ld, rows, (int)C, c0, lse, (const long long*)target, dnll,
                                                                   mask_col0 ? 1 : 0);
    PR_CUDA_LAUNCH_CHECK("ce_grad_chunk_kernel");
    return PR_OK;
}
