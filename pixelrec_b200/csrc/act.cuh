// pixelrec_b200 -- activation functions of the feed-forward block (REC/model/layers.py:640-660) and their derivatives,
// shared by the elementwise kernels (csrc/ln.cu) and the GEMM epilogues (csrc/gemm.cu).
#pragma once
#include "common.cuh"

namespace pr {

// ------------------------------------------------------------------ activations (layers.py:640-660)
__device__ __forceinline__ float act_f(float x, int act) {
    switch (act) {
        case PR_ACT_GELU: return x * 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
        case PR_ACT_RELU: return fmaxf(x, 0.f);
        case PR_ACT_SWISH: return x / (1.0f + expf(-x));
        case PR_ACT_TANH: return tanhf(x);
        case PR_ACT_QUICK_GELU: return x / (1.0f + expf(-1.702f * x));
        default: return 1.0f / (1.0f + expf(-x));
    }
}
__device__ __forceinline__ float act_df(float x, int act) {
    switch (act) {
        case PR_ACT_GELU: {
            const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
            const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
            return cdf + x * pdf;
        }
        case PR_ACT_RELU: return x > 0.f ? 1.f : 0.f;
        case PR_ACT_SWISH: {
            const float s = 1.0f / (1.0f + expf(-x));
            return s + x * s * (1.0f - s);
        }
        case PR_ACT_TANH: {
            const float t = tanhf(x);
            return 1.0f - t * t;
        }
        case PR_ACT_QUICK_GELU: {
            const float s = 1.0f / (1.0f + expf(-1.702f * x));
            return s + 1.702f * x * s * (1.0f - s);
        }
        default: {
            const float s = 1.0f / (1.0f + expf(-x));
            return s * (1.0f - s);
        }
    }
}

// erf-GELU and its derivative for the GEMM epilogues (csrc/gemm.cu), where 128 activations per thread and tile have to fit
// under the tile's MMA time: erff + expf (~65 instructions per element) made those epilogues the bottleneck.
// Phi(x) = 0.5 (1 + erf(x / sqrt 2)) through Abramowitz & Stegun 7.1.26 (|error of erf| <= 1.5e-7) sharing ONE exponential
// u = exp(-x^2 / 2) between Phi and the density:  1 - erf(z) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-z^2),
// t = 1 / (1 + p z), z = |x| / sqrt 2.  The absolute error of Phi (<= 1e-7) is multiplied by x in gelu(x) = x Phi(x), i.e. it is
// a relative error of ~2e-7 for x > 0 and an absolute one below 1e-7 |x| for x < 0 -- inside fp32 rounding of the layer.
__device__ __forceinline__ void gelu_cdf_pdf_fast(float x, float& cdf, float& pdf) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
    const float u = __expf(-0.5f * x * x);
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float half_tail = 0.5f * p * t * u;                   // 0.5 (1 - erf(z)) = Phi(-|x|)
    cdf = (x >= 0.f) ? 1.0f - half_tail : half_tail;
    pdf = 0.39894228040143267794f * u;
}
__device__ __forceinline__ float gelu_fast(float x) {
    float c, d;
    gelu_cdf_pdf_fast(x, c, d);
    return x * c;
}
__device__ __forceinline__ float dgelu_fast(float x) {
    float c, d;
    gelu_cdf_pdf_fast(x, c, d);
    return fmaf(x, d, c);
}

}  // namespace pr
