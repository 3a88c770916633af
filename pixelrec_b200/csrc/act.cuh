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

}  // namespace pr
