// pixelrec_b200 -- K8: sampled-negative pairwise (BPR-style) loss, forward and backward, one launch each.
//   replaces REC/model/IDNet/sasrec.py:88-92 (== gru4rec.py:63-67, PixelNet/mosasrec.py:89-93), which is
//   ~10 elementwise/reduction launches in the reference.  HBM-bound: one warp per (b,t) position reads the
//   three D-float rows once (128-bit loads), positions with mask == 0 are never read.
#include <algorithm>
#include <math.h>

#include "common.cuh"

namespace pr {

// VPL > 0: the three rows of a position live in registers (VPL float4 per lane each, D4 <= 32*VPL) and all of their loads are
// issued before the first multiply -- with the runtime-bounded column loop (VPL == 0, any D) a warp had three 512-byte requests in
// flight at a time and half the warps (masked positions) none: 0.52 / 0.61 of the HBM peak.  Same summation order either way.
template <int VPL>
__global__ void __launch_bounds__(256) bpr_fwd_kernel(const float* __restrict__ out, const float* __restrict__ tp,
                                                      const float* __restrict__ tn, long long t_seq_stride,
                                                      const long long* __restrict__ mask, long long B, int L, int D4,
                                                      float* __restrict__ pos_score, float* __restrict__ neg_score,
                                                      float* __restrict__ coef, float* __restrict__ terms) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long npos = B * L;
    const float invB = 1.0f / (float)B;
    for (long long p = warp; p < npos; p += nwarps) {
        const long long b = p / L;
        const int t = (int)(p - b * L);
        const float m = (float)mask[p];
        float ps = 0.f, ns = 0.f, cf = 0.f, term = 0.f;
        if (m != 0.f) {
            const float4* o4 = reinterpret_cast<const float4*>(out) + p * D4;
            const float4* p4 = reinterpret_cast<const float4*>(tp + b * t_seq_stride) + (long long)t * D4;
            const float4* n4 = reinterpret_cast<const float4*>(tn + b * t_seq_stride) + (long long)t * D4;
            if constexpr (VPL > 0) {
                float4 o[VPL > 0 ? VPL : 1], a[VPL > 0 ? VPL : 1], n[VPL > 0 ? VPL : 1];
#pragma unroll
                for (int j = 0; j < VPL; ++j) {
                    const int c = lane + 32 * j;
                    const bool in = c < D4;
                    o[j] = in ? ldg_stream(o4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    a[j] = in ? __ldg(p4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    n[j] = in ? __ldg(n4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int j = 0; j < VPL; ++j) {
                    if (lane + 32 * j < D4) {
                        ps += (o[j].x * a[j].x + o[j].y * a[j].y) + (o[j].z * a[j].z + o[j].w * a[j].w);
                        ns += (o[j].x * n[j].x + o[j].y * n[j].y) + (o[j].z * n[j].z + o[j].w * n[j].w);
                    }
                }
            } else {
                for (int c = lane; c < D4; c += 32) {
                    const float4 o = ldg_stream(o4 + c), a = __ldg(p4 + c), n = __ldg(n4 + c);
                    ps += (o.x * a.x + o.y * a.y) + (o.z * a.z + o.w * a.w);
                    ns += (o.x * n.x + o.y * n.y) + (o.z * n.z + o.w * n.w);
                }
            }
            ps = warp_sum(ps);
            ns = warp_sum(ns);
            const float s = ps - ns;
            const float sig = 1.0f / (1.0f + expf(-s));
            term = -logf(sig + 1e-8f) * m;
            cf = -(m * invB) * sig * (1.0f - sig) / (sig + 1e-8f);
        }
        if (lane == 0) {
            if (pos_score) pos_score[p] = ps;
            if (neg_score) neg_score[p] = ns;
            coef[p] = cf;
            terms[p] = term;
        }
    }
}

// loss = (1/B) * sum_p terms[p], fixed summation order (thread-strided partials, then a shared-memory tree)
__global__ void __launch_bounds__(1024) bpr_reduce_kernel(const float* __restrict__ terms, long long n, float invB,
                                                          float* __restrict__ loss) {
    __shared__ float sm[1024];
    float s = 0.f;
    for (long long i = threadIdx.x; i < n; i += 1024) s += terms[i];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = sm[0] * invB;
}

template <int VPL>
__global__ void __launch_bounds__(256) bpr_bwd_kernel(const float* __restrict__ out, const float* __restrict__ tp,
                                                      const float* __restrict__ tn, long long t_seq_stride,
                                                      const float* __restrict__ coef, const float* __restrict__ dloss,
                                                      long long B, int L, int D4, float* __restrict__ d_out,
                                                      float* __restrict__ d_tp, float* __restrict__ d_tn,
                                                      long long d_seq_stride) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long npos = B * L;
    const float g = dloss ? *dloss : 1.0f;
    for (long long p = warp; p < npos; p += nwarps) {
        const long long b = p / L;
        const int t = (int)(p - b * L);
        const float c = g * coef[p];
        float4* do4 = reinterpret_cast<float4*>(d_out) + p * D4;
        float4* dp4 = reinterpret_cast<float4*>(d_tp + b * d_seq_stride) + (long long)t * D4;
        float4* dn4 = reinterpret_cast<float4*>(d_tn + b * d_seq_stride) + (long long)t * D4;
        if (c == 0.f) {  // masked position: gradients are exactly zero, nothing to read
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int cc = lane; cc < D4; cc += 32) {
                do4[cc] = z; dp4[cc] = z; dn4[cc] = z;
            }
            continue;
        }
        const float4* o4 = reinterpret_cast<const float4*>(out) + p * D4;
        const float4* p4 = reinterpret_cast<const float4*>(tp + b * t_seq_stride) + (long long)t * D4;
        const float4* n4 = reinterpret_cast<const float4*>(tn + b * t_seq_stride) + (long long)t * D4;
        if constexpr (VPL > 0) {
            float4 o[VPL > 0 ? VPL : 1], a[VPL > 0 ? VPL : 1], n[VPL > 0 ? VPL : 1];
#pragma unroll
            for (int j = 0; j < VPL; ++j) {
                const int cc = lane + 32 * j;
                const bool in = cc < D4;
                o[j] = in ? ldg_stream(o4 + cc) : make_float4(0.f, 0.f, 0.f, 0.f);
                a[j] = in ? __ldg(p4 + cc) : make_float4(0.f, 0.f, 0.f, 0.f);
                n[j] = in ? __ldg(n4 + cc) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < VPL; ++j) {
                const int cc = lane + 32 * j;
                if (cc < D4) {
                    float4 r;
                    r.x = c * (a[j].x - n[j].x); r.y = c * (a[j].y - n[j].y); r.z = c * (a[j].z - n[j].z); r.w = c * (a[j].w - n[j].w);
                    do4[cc] = r;
                    r.x = c * o[j].x; r.y = c * o[j].y; r.z = c * o[j].z; r.w = c * o[j].w;
                    dp4[cc] = r;
                    r.x = -r.x; r.y = -r.y; r.z = -r.z; r.w = -r.w;
                    dn4[cc] = r;
                }
            }
        } else {
        for (int cc = lane; cc < D4; cc += 32) {
            const float4 o = ldg_stream(o4 + cc), a = __ldg(p4 + cc), n = __ldg(n4 + cc);
            float4 r;
            r.x = c * (a.x - n.x); r.y = c * (a.y - n.y); r.z = c * (a.z - n.z); r.w = c * (a.w - n.w);
            do4[cc] = r;
            r.x = c * o.x; r.y = c * o.y; r.z = c * o.z; r.w = c * o.w;
            dp4[cc] = r;
            r.x = -r.x; r.y = -r.y; r.z = -r.z; r.w = -r.w;
            dn4[cc] = r;
        }
        }
    }
}

}  // namespace pr

using namespace pr;

extern "C" int pr_bpr_loss_fwd_f32(const float* out, const float* tp, const float* tn, int64_t t_seq_stride,
                                   const int64_t* mask, int64_t B, int64_t L, int64_t D, float* pos_score,
                                   float* neg_score, float* coef, float* loss_terms, float* loss,
                                   pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PR_CHECK_ARG(B > 0 && L > 0 && D > 0 && D % 4 == 0, "pr_bpr_loss_fwd_f32: bad shape B=%lld L=%lld D=%lld",
                 (long long)B, (long long)L, (long long)D);
    PR_CHECK_ARG(L < (1 << 30) && t_seq_stride % 4 == 0, "pr_bpr_loss_fwd_f32: bad L / stride");
    PR_CHECK_ARG(out && tp && tn && mask && coef && loss_terms && loss, "pr_bpr_loss_fwd_f32: null pointer");
    PR_CHECK_ARG(aligned16(out) && aligned16(tp) && aligned16(tn), "pr_bpr_loss_fwd_f32: pointers must be 16-byte aligned");
    const long long npos = B * L;
    const int grid = (int)std::max<long long>(1, std::min<long long>((npos + 7) / 8, (long long)sm_count() * 8));
    const int D4 = (int)(D / 4);
#define PR_BPR_FWD(V) bpr_fwd_kernel<V><<<grid, 256, 0, stream>>>(out, tp, tn, t_seq_stride, (const long long*)mask, B, (int)L, D4, \
                                                                  pos_score, neg_score, coef, loss_terms)
    if (D4 <= 32) PR_BPR_FWD(1);
    else if (D4 <= 64) PR_BPR_FWD(2);
    else if (D4 <= 128) PR_BPR_FWD(4);
    else if (D4 <= 256) PR_BPR_FWD(8);
    else PR_BPR_FWD(0);
#undef PR_BPR_FWD
    bpr_reduce_kernel<<<1, 1024, 0, stream>>>(loss_terms, npos, 1.0f / (float)B, loss);
    PR_CUDA_LAUNCH_CHECK("bpr_fwd_kernel");
    return PR_OK;
}

extern "C" int pr_bpr_loss_bwd_f32(const float* out, const float* tp, const float* tn, int64_t t_seq_stride,
                                   const float* coef, const float* dloss, int64_t B, int64_t L, int64_t D,
                                   float* d_out, float* d_tp, float* d_tn, int64_t d_seq_stride,
                                   pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PR_CHECK_ARG(B > 0 && L > 0 && D > 0 && D % 4 == 0, "pr_bpr_loss_bwd_f32: bad shape");
    PR_CHECK_ARG(t_seq_stride % 4 == 0 && d_seq_stride % 4 == 0, "pr_bpr_loss_bwd_f32: strides must be multiples of 4");
    PR_CHECK_ARG(out && tp && tn && coef && d_out && d_tp && d_tn, "pr_bpr_loss_bwd_f32: null pointer");
    PR_CHECK_ARG(aligned16(out) && aligned16(tp) && aligned16(tn) && aligned16(d_out) && aligned16(d_tp) && aligned16(d_tn),
                 "pr_bpr_loss_bwd_f32: pointers must be 16-byte aligned");
    const long long npos = B * L;
    const int grid = (int)std::max<long long>(1, std::min<long long>((npos + 7) / 8, (long long)sm_count() * 8));
    const int D4 = (int)(D / 4);
#define PR_BPR_BWD(V) bpr_bwd_kernel<V><<<grid, 256, 0, stream>>>(out, tp, tn, t_seq_stride, coef, dloss, B, (int)L, D4, d_out, d_tp, \
                                                                  d_tn, d_seq_stride)
    if (D4 <= 32) PR_BPR_BWD(1);
    else if (D4 <= 64) PR_BPR_BWD(2);
    else if (D4 <= 128) PR_BPR_BWD(4);
    else if (D4 <= 256) PR_BPR_BWD(8);
    else PR_BPR_BWD(0);
#undef PR_BPR_BWD
    PR_CUDA_LAUNCH_CHECK("bpr_bwd_kernel");
    return PR_OK;
}
