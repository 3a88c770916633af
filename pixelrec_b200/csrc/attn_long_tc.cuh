// pixelrec_b200 -- long-sequence attention FORWARD on tensor cores (mma.sync m16n8k8 TF32, fp32 accumulate); device code.
//   Same contract as attn_long_fwd_kernel (attn_long.cuh): ctx [B, L, h*dh], lse [B*h, L]; 64 < L <= 256 (any L <= 256 works),
//   dh in {32, 64, 128}, optional causal / key-padding mask with the reference's additive -1e9, no dropout.
//   PR_TUNE_ATTN_LONG_TC selects it; the backward stays the fp32 two-pass kernels, which only need ctx and lse.
//
// One CTA per (sequence, head) at a time, K and V of the head resident in shared memory (TF32-rounded once, rows padded by one
// float4: the B-fragment reads of both products are bank-conflict free), 8 warps, each owning 16-row query tiles.  Per tile:
// Q (scaled by 1/sqrt(dh)) is staged through a warp-private buffer into A fragments that stay in registers; keys are walked
// in blocks of 64 with an online softmax (running row max / sum, FlashAttention-2 style); P never leaves registers: the
// accumulator fragment of S is reused as the A fragment of P.V through a permutation of the key index inside each 8-key step
// (logical k = t <-> key 2t, k = t+4 <-> key 2t+1), applied to the V rows the B fragments read.
//
// Written against to_tf32() / mma_tf32(), which attn_long.cu defines with PTX and tests/emu/emu_mma.h emulates lane-exactly
// (PTX ISA fragment layouts), so the fragment index arithmetic is checked against the oracle without a GPU.
#pragma once

namespace pr {

constexpr int ALT_WARPS = 8;
constexpr int ALT_THREADS = ALT_WARPS * 32;
constexpr int ALT_KB = 64;                  // keys per online-softmax block

template <int DH>
__host__ __device__ inline size_t long_tc_smem_floats(int L) {
    return (size_t)2 * L * (DH + 4) + (size_t)ALT_WARPS * 16 * (DH + 4);
}

template <int DH>
__global__ void __launch_bounds__(ALT_THREADS, 1) attn_long_tc_fwd_kernel(const LongAttnArgs A) {
    constexpr int RS = DH + 4;              // row stride in floats (K, V and the Q staging tiles)
    constexpr int KS = DH / 8;              // k-steps of Q.K^T  == n-tiles of the output
    PR_DYN_SMEM_F4(smem4);
    float* Ks = reinterpret_cast<float*>(smem4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int L = A.L;
    float* Vs = Ks + (size_t)L * RS;
    float* Qs = Vs + (size_t)L * RS + (size_t)warp * 16 * RS;
    const long long Dm = (long long)A.h * DH;
    const long long n_items = (long long)A.B * A.h;
    const int n_kb = (L + ALT_KB - 1) / ALT_KB;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const long long b = item / A.h;
        const int hd = (int)(item - b * A.h);
        const long long base = b * L * A.ld + (long long)hd * DH;
        __syncthreads();                                       // the previous item's K/V are no longer read
        for (int idx = threadIdx.x; idx < L * (DH / 4); idx += ALT_THREADS) {
            const int r = idx / (DH / 4), c = idx - r * (DH / 4);
            const float4 kv = PR_LDG4(reinterpret_cast<const float4*>(A.k + base + (long long)r * A.ld) + c);
            const float4 vv = PR_LDG4(reinterpret_cast<const float4*>(A.v + base + (long long)r * A.ld) + c);
            float* kd = Ks + r * RS + 4 * c;
            float* vd = Vs + r * RS + 4 * c;
            kd[0] = __uint_as_float(to_tf32(kv.x)); kd[1] = __uint_as_float(to_tf32(kv.y));
            kd[2] = __uint_as_float(to_tf32(kv.z)); kd[3] = __uint_as_float(to_tf32(kv.w));
            vd[0] = __uint_as_float(to_tf32(vv.x)); vd[1] = __uint_as_float(to_tf32(vv.y));
            vd[2] = __uint_as_float(to_tf32(vv.z)); vd[3] = __uint_as_float(to_tf32(vv.w));
        }
        __syncthreads();
        for (int i0 = warp * 16; i0 < L; i0 += ALT_WARPS * 16) {
            // ---- Q tile -> warp-private staging (scaled, TF32) -> A fragments in registers
            for (int idx = lane; idx < 16 * (DH / 4); idx += 32) {
                const int r = idx / (DH / 4), c = idx - r * (DH / 4);
                const int i = min(i0 + r, L - 1);
                const float4 qv = PR_LDG4(reinterpret_cast<const float4*>(A.q + base + (long long)i * A.ld) + c);
                float* qd = Qs + r * RS + 4 * c;
                qd[0] = __uint_as_float(to_tf32(qv.x * A.scale)); qd[1] = __uint_as_float(to_tf32(qv.y * A.scale));
                qd[2] = __uint_as_float(to_tf32(qv.z * A.scale)); qd[3] = __uint_as_float(to_tf32(qv.w * A.scale));
            }
            __syncwarp();
            uint32_t qf[KS][4];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                qf[ks][0] = __float_as_uint(Qs[g * RS + 8 * ks + t]);
                qf[ks][1] = __float_as_uint(Qs[(g + 8) * RS + 8 * ks + t]);
                qf[ks][2] = __float_as_uint(Qs[g * RS + 8 * ks + t + 4]);
                qf[ks][3] = __float_as_uint(Qs[(g + 8) * RS + 8 * ks + t + 4]);
            }
            __syncwarp();                                      // staging may be overwritten by this warp's next tile
            float o[KS][4];
#pragma unroll
            for (int dn = 0; dn < KS; ++dn) { o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f; }
            float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;       // rows i0+g and i0+g+8
            const int r0 = i0 + g, r1 = i0 + g + 8;
            for (int kb = 0; kb < n_kb; ++kb) {
                float s[8][4];
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) {
                        const int key = min(kb * ALT_KB + nt * 8 + g, L - 1);          // B[k = d][n = key] = K[key][d]
                        const uint32_t b0 = __float_as_uint(Ks[key * RS + 8 * ks + t]);
                        const uint32_t b1 = __float_as_uint(Ks[key * RS + 8 * ks + t + 4]);
                        mma_tf32(s[nt], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b0, b1);
                    }
                }
                // ---- mask + online softmax; this lane holds columns 2t, 2t+1 of every 8-key n-tile for rows r0, r1
                float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = kb * ALT_KB + nt * 8 + 2 * t + e;
                        const bool inb = j < L;
                        const bool kv_ok = inb && (A.key_ids == nullptr || A.key_ids[b * L + j] != 0);
                        const float bias0 = (kv_ok && (!A.causal || j <= r0)) ? 0.0f : -1e9f;
                        const float bias1 = (kv_ok && (!A.causal || j <= r1)) ? 0.0f : -1e9f;
                        s[nt][e] = inb ? s[nt][e] + bias0 : -INFINITY;
                        s[nt][2 + e] = inb ? s[nt][2 + e] + bias1 : -INFINITY;
                        bm0 = fmaxf(bm0, s[nt][e]);
                        bm1 = fmaxf(bm1, s[nt][2 + e]);
                    }
                }
                bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
                bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
                bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
                bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
                const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);     // finite: every block has a key < L
                const float c0 = expf(m0 - mn0), c1 = expf(m1 - mn1);       // exp(-inf) = 0 on the first block
                m0 = mn0; m1 = mn1;
                l0 *= c0; l1 *= c1;
#pragma unroll
                for (int dn = 0; dn < KS; ++dn) { o[dn][0] *= c0; o[dn][1] *= c0; o[dn][2] *= c1; o[dn][3] *= c1; }
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const float p00 = expf(s[nt][0] - mn0), p01 = expf(s[nt][1] - mn0);
                    const float p10 = expf(s[nt][2] - mn1), p11 = expf(s[nt][3] - mn1);
                    l0 += p00 + p01;
                    l1 += p10 + p11;
                    // accumulator fragment -> A fragment of P.V: logical k = t is key 2t, k = t+4 is key 2t+1 (of this n-tile)
                    const uint32_t a0 = to_tf32(p00), a1 = to_tf32(p10), a2 = to_tf32(p01), a3 = to_tf32(p11);
                    const int key0 = min(kb * ALT_KB + nt * 8 + 2 * t, L - 1), key1 = min(kb * ALT_KB + nt * 8 + 2 * t + 1, L - 1);
#pragma unroll
                    for (int dn = 0; dn < KS; ++dn) {                        // B[k][n = d] = V[key(k)][d]
                        const uint32_t b0 = __float_as_uint(Vs[key0 * RS + 8 * dn + g]);
                        const uint32_t b1 = __float_as_uint(Vs[key1 * RS + 8 * dn + g]);
                        mma_tf32(o[dn], a0, a1, a2, a3, b0, b1);
                    }
                }
            }
            l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
            l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
            l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
            l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
            const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
            float* out0 = A.ctx + (b * L + r0) * Dm + (long long)hd * DH;
            float* out1 = A.ctx + (b * L + r1) * Dm + (long long)hd * DH;
#pragma unroll
            for (int dn = 0; dn < KS; ++dn) {
                if (r0 < L) *reinterpret_cast<float2*>(out0 + 8 * dn + 2 * t) = make_float2(o[dn][0] * inv0, o[dn][1] * inv0);
                if (r1 < L) *reinterpret_cast<float2*>(out1 + 8 * dn + 2 * t) = make_float2(o[dn][2] * inv1, o[dn][3] * inv1);
            }
            if (t == 0) {
                if (r0 < L) A.lse[item * L + r0] = m0 + logf(l0);
                if (r1 < L) A.lse[item * L + r1] = m1 + logf(l1);
            }
        }
    }
}

// ============================================================================================ backward on tensor cores
// Recomputes S with exactly the forward's operands (Q scaled then TF32-rounded, K TF32-rounded, same k-step order), so
// P = exp(S - lse) is consistent with the forward's lse.  Two passes, like the fp32 kernels:
//   A: K,V resident; per 16-row query tile and 8-key n-tile: S, dP = dO V^T -> dS -> dQ += dS K       (+ delta_i = <dO_i, O_i>)
//   B: Q,dO resident; per 16-row KEY tile and 8-query n-tile: S^T, dP^T -> P^T, dS^T -> dV += P^T dO, dK += dS^T Q
// The S / dP fragments of one n-tile are turned into A fragments of the second product by the same key (query) permutation
// as in the forward, so nothing but the resident matrices and the staged tiles touches shared memory.
template <int DH>
__host__ __device__ inline size_t long_tc_bwd_smem_floats(int L) {
    return (size_t)2 * L * (DH + 4) + (size_t)ALT_WARPS * 2 * 16 * (DH + 4) + (size_t)2 * ((L + 3) / 4 * 4);
}

// stage rows [i0, i0+16) (clamped to L-1) of a [*, ld] matrix, times `mul`, TF32-rounded, into a warp-private tile
template <int DH>
__device__ __forceinline__ void alt_stage_tile(float* dst, const float* src, long long ld, int i0, int L, float mul, int lane) {
    constexpr int RS = DH + 4;
    for (int idx = lane; idx < 16 * (DH / 4); idx += 32) {
        const int r = idx / (DH / 4), c = idx - r * (DH / 4);
        const int i = min(i0 + r, L - 1);
        const float4 x = PR_LDG4(reinterpret_cast<const float4*>(src + (long long)i * ld) + c);
        float* d = dst + r * RS + 4 * c;
        d[0] = __uint_as_float(to_tf32(x.x * mul)); d[1] = __uint_as_float(to_tf32(x.y * mul));
        d[2] = __uint_as_float(to_tf32(x.z * mul)); d[3] = __uint_as_float(to_tf32(x.w * mul));
    }
}
template <int DH>
__device__ __forceinline__ void alt_load_afrag(const float* tile, int g, int t, uint32_t (&f)[DH / 8][4]) {
    constexpr int RS = DH + 4;
#pragma unroll
    for (int ks = 0; ks < DH / 8; ++ks) {
        f[ks][0] = __float_as_uint(tile[g * RS + 8 * ks + t]);
        f[ks][1] = __float_as_uint(tile[(g + 8) * RS + 8 * ks + t]);
        f[ks][2] = __float_as_uint(tile[g * RS + 8 * ks + t + 4]);
        f[ks][3] = __float_as_uint(tile[(g + 8) * RS + 8 * ks + t + 4]);
    }
}
// whole CTA: resident[r][:] = tf32(src[r][:] * mul)
template <int DH>
__device__ __forceinline__ void alt_load_resident(float* dst, const float* src, long long ld, int L, float mul) {
    constexpr int RS = DH + 4;
    for (int idx = threadIdx.x; idx < L * (DH / 4); idx += ALT_THREADS) {
        const int r = idx / (DH / 4), c = idx - r * (DH / 4);
        const float4 x = PR_LDG4(reinterpret_cast<const float4*>(src + (long long)r * ld) + c);
        float* d = dst + r * RS + 4 * c;
        d[0] = __uint_as_float(to_tf32(x.x * mul)); d[1] = __uint_as_float(to_tf32(x.y * mul));
        d[2] = __uint_as_float(to_tf32(x.z * mul)); d[3] = __uint_as_float(to_tf32(x.w * mul));
    }
}

template <int DH>
__global__ void __launch_bounds__(ALT_THREADS, 1) attn_long_tc_bwd_dq_kernel(const LongAttnArgs A) {
    constexpr int RS = DH + 4, KS = DH / 8;
    PR_DYN_SMEM_F4(smem4);
    float* Ks = reinterpret_cast<float*>(smem4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int L = A.L;
    float* Vs = Ks + (size_t)L * RS;
    float* Qt = Vs + (size_t)L * RS + (size_t)warp * 2 * 16 * RS;      // this warp's Q tile, then its dO tile
    float* Gt = Qt + 16 * RS;
    const long long Dm = (long long)A.h * DH;
    const long long n_items = (long long)A.B * A.h;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const long long b = item / A.h;
        const int hd = (int)(item - b * A.h);
        const long long base = b * L * A.ld + (long long)hd * DH;
        const long long obase = b * L * Dm + (long long)hd * DH;
        __syncthreads();
        alt_load_resident<DH>(Ks, A.k + base, A.ld, L, 1.0f);
        alt_load_resident<DH>(Vs, A.v + base, A.ld, L, 1.0f);
        __syncthreads();
        for (int i0 = warp * 16; i0 < L; i0 += ALT_WARPS * 16) {
            // delta_i = <dO_i, O_i> in fp32: two lanes per row, each half of the head dimension
            float dl;
            {
                const int r = lane >> 1, i = min(i0 + r, L - 1);
                const float4* go = reinterpret_cast<const float4*>(A.dctx + obase + (long long)i * Dm) + (lane & 1) * (DH / 8);
                const float4* oo = reinterpret_cast<const float4*>(A.ctx_in + obase + (long long)i * Dm) + (lane & 1) * (DH / 8);
                float part = 0.f;
                for (int c = 0; c < DH / 8; ++c) {
                    const float4 x = PR_LDG4(go + c), y = PR_LDG4(oo + c);
                    part = fmaf(x.x, y.x, part); part = fmaf(x.y, y.y, part); part = fmaf(x.z, y.z, part); part = fmaf(x.w, y.w, part);
                }
                dl = part + __shfl_xor_sync(0xffffffffu, part, 1);
                if ((lane & 1) == 0 && i0 + r < L) A.delta[item * L + i0 + r] = dl;
            }
            const float d0 = __shfl_sync(0xffffffffu, dl, 2 * g), d1 = __shfl_sync(0xffffffffu, dl, 2 * (g + 8));
            alt_stage_tile<DH>(Qt, A.q + base, A.ld, i0, L, A.scale, lane);
            alt_stage_tile<DH>(Gt, A.dctx + obase, Dm, i0, L, 1.0f, lane);
            __syncwarp();
            uint32_t qf[KS][4], gf[KS][4];
            alt_load_afrag<DH>(Qt, g, t, qf);
            alt_load_afrag<DH>(Gt, g, t, gf);
            __syncwarp();
            const int r0 = i0 + g, r1 = i0 + g + 8;
            const float lse0 = A.lse[item * L + min(r0, L - 1)], lse1 = A.lse[item * L + min(r1, L - 1)];
            float o[KS][4];
#pragma unroll
            for (int dn = 0; dn < KS; ++dn) { o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f; }
            for (int n0 = 0; n0 < L; n0 += 8) {                       // one 8-key n-tile at a time
                float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
                const int key = min(n0 + g, L - 1);
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    mma_tf32(s, qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], __float_as_uint(Ks[key * RS + 8 * ks + t]),
                             __float_as_uint(Ks[key * RS + 8 * ks + t + 4]));
                    mma_tf32(dp, gf[ks][0], gf[ks][1], gf[ks][2], gf[ks][3], __float_as_uint(Vs[key * RS + 8 * ks + t]),
                             __float_as_uint(Vs[key * RS + 8 * ks + t + 4]));
                }
                float ds[4];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = n0 + 2 * t + e;
                    const bool inb = j < L;
                    const bool kv_ok = inb && (A.key_ids == nullptr || A.key_ids[b * L + j] != 0);
                    const float b0 = (kv_ok && (!A.causal || j <= r0)) ? 0.0f : -1e9f;
                    const float b1 = (kv_ok && (!A.causal || j <= r1)) ? 0.0f : -1e9f;
                    const float p0 = inb ? expf(s[e] + b0 - lse0) : 0.f;
                    const float p1 = inb ? expf(s[2 + e] + b1 - lse1) : 0.f;
                    ds[e] = p0 * (dp[e] - d0) * A.scale;
                    ds[2 + e] = p1 * (dp[2 + e] - d1) * A.scale;
                }
                const uint32_t a0 = to_tf32(ds[0]), a1 = to_tf32(ds[2]), a2 = to_tf32(ds[1]), a3 = to_tf32(ds[3]);
                const int key0 = min(n0 + 2 * t, L - 1), key1 = min(n0 + 2 * t + 1, L - 1);
#pragma unroll
                for (int dn = 0; dn < KS; ++dn)                           // dQ += dS K
                    mma_tf32(o[dn], a0, a1, a2, a3, __float_as_uint(Ks[key0 * RS + 8 * dn + g]),
                             __float_as_uint(Ks[key1 * RS + 8 * dn + g]));
            }
            float* out0 = A.dq + (b * L + r0) * A.ld_grad + (long long)hd * DH;
            float* out1 = A.dq + (b * L + r1) * A.ld_grad + (long long)hd * DH;
#pragma unroll
            for (int dn = 0; dn < KS; ++dn) {
                if (r0 < L) *reinterpret_cast<float2*>(out0 + 8 * dn + 2 * t) = make_float2(o[dn][0], o[dn][1]);
                if (r1 < L) *reinterpret_cast<float2*>(out1 + 8 * dn + 2 * t) = make_float2(o[dn][2], o[dn][3]);
            }
        }
    }
}

template <int DH>
__global__ void __launch_bounds__(ALT_THREADS, 1) attn_long_tc_bwd_dkv_kernel(const LongAttnArgs A) {
    constexpr int RS = DH + 4, KS = DH / 8;
    PR_DYN_SMEM_F4(smem4);
    float* Qs = reinterpret_cast<float*>(smem4);                       // resident: scale * Q  and dO  (TF32)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int L = A.L, Lp = (L + 3) / 4 * 4;
    float* Gs = Qs + (size_t)L * RS;
    float* Kt = Gs + (size_t)L * RS + (size_t)warp * 2 * 16 * RS;      // this warp's K tile and V tile
    float* Vt = Kt + 16 * RS;
    float* lse_s = Gs + (size_t)L * RS + (size_t)ALT_WARPS * 2 * 16 * RS;
    float* del_s = lse_s + Lp;
    const long long Dm = (long long)A.h * DH;
    const long long n_items = (long long)A.B * A.h;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const long long b = item / A.h;
        const int hd = (int)(item - b * A.h);
        const long long base = b * L * A.ld + (long long)hd * DH;
        const long long obase = b * L * Dm + (long long)hd * DH;
        __syncthreads();
        alt_load_resident<DH>(Qs, A.q + base, A.ld, L, A.scale);
        alt_load_resident<DH>(Gs, A.dctx + obase, Dm, L, 1.0f);
        for (int i = threadIdx.x; i < L; i += ALT_THREADS) {
            lse_s[i] = A.lse[item * L + i];
            del_s[i] = A.delta[item * L + i];
        }
        __syncthreads();
        for (int j0 = warp * 16; j0 < L; j0 += ALT_WARPS * 16) {
            alt_stage_tile<DH>(Kt, A.k + base, A.ld, j0, L, 1.0f, lane);
            alt_stage_tile<DH>(Vt, A.v + base, A.ld, j0, L, 1.0f, lane);
            __syncwarp();
            uint32_t kf[KS][4], vf[KS][4];
            alt_load_afrag<DH>(Kt, g, t, kf);
            alt_load_afrag<DH>(Vt, g, t, vf);
            __syncwarp();
            const int ja = j0 + g, jb = j0 + g + 8;                        // this lane's two key rows
            const bool va = ja < L && (A.key_ids == nullptr || A.key_ids[b * L + ja] != 0);
            const bool vb = jb < L && (A.key_ids == nullptr || A.key_ids[b * L + jb] != 0);
            float dk[KS][4], dv[KS][4];
#pragma unroll
            for (int dn = 0; dn < KS; ++dn) {
                dk[dn][0] = dk[dn][1] = dk[dn][2] = dk[dn][3] = 0.f;
                dv[dn][0] = dv[dn][1] = dv[dn][2] = dv[dn][3] = 0.f;
            }
            for (int n0 = 0; n0 < L; n0 += 8) {                           // one 8-query n-tile at a time
                float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
                const int qi = min(n0 + g, L - 1);
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    mma_tf32(s, kf[ks][0], kf[ks][1], kf[ks][2], kf[ks][3], __float_as_uint(Qs[qi * RS + 8 * ks + t]),
                             __float_as_uint(Qs[qi * RS + 8 * ks + t + 4]));      // S^T[key][query] = k . (scale q)
                    mma_tf32(dp, vf[ks][0], vf[ks][1], vf[ks][2], vf[ks][3], __float_as_uint(Gs[qi * RS + 8 * ks + t]),
                             __float_as_uint(Gs[qi * RS + 8 * ks + t + 4]));      // dP^T[key][query] = v . dO
                }
                float p[4], dsn[4];                                            // dsn = dS / scale (the resident Q carries the scale)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int i = n0 + 2 * t + e;
                    const bool inb = i < L;
                    const float lse_i = lse_s[min(i, L - 1)], del_i = del_s[min(i, L - 1)];
                    const float ba = (va && (!A.causal || ja <= i)) ? 0.0f : -1e9f;
                    const float bb = (vb && (!A.causal || jb <= i)) ? 0.0f : -1e9f;
                    p[e] = (inb && ja < L) ? expf(s[e] + ba - lse_i) : 0.f;
                    p[2 + e] = (inb && jb < L) ? expf(s[2 + e] + bb - lse_i) : 0.f;
                    dsn[e] = p[e] * (dp[e] - del_i);
                    dsn[2 + e] = p[2 + e] * (dp[2 + e] - del_i);
                }
                const uint32_t pa0 = to_tf32(p[0]), pa1 = to_tf32(p[2]), pa2 = to_tf32(p[1]), pa3 = to_tf32(p[3]);
                const uint32_t sa0 = to_tf32(dsn[0]), sa1 = to_tf32(dsn[2]), sa2 = to_tf32(dsn[1]), sa3 = to_tf32(dsn[3]);
                const int q0 = min(n0 + 2 * t, L - 1), q1 = min(n0 + 2 * t + 1, L - 1);
#pragma unroll
                for (int dn = 0; dn < KS; ++dn) {
                    mma_tf32(dv[dn], pa0, pa1, pa2, pa3, __float_as_uint(Gs[q0 * RS + 8 * dn + g]),
                             __float_as_uint(Gs[q1 * RS + 8 * dn + g]));          // dV += P^T dO
                    mma_tf32(dk[dn], sa0, sa1, sa2, sa3, __float_as_uint(Qs[q0 * RS + 8 * dn + g]),
                             __float_as_uint(Qs[q1 * RS + 8 * dn + g]));          // dK += (dS^T / scale) (scale Q)
                }
            }
            float* dka = A.dk + (b * L + ja) * A.ld_grad + (long long)hd * DH;
            float* dkb = A.dk + (b * L + jb) * A.ld_grad + (long long)hd * DH;
            float* dva = A.dv + (b * L + ja) * A.ld_grad + (long long)hd * DH;
            float* dvb = A.dv + (b * L + jb) * A.ld_grad + (long long)hd * DH;
#pragma unroll
            for (int dn = 0; dn < KS; ++dn) {
                if (ja < L) {
                    *reinterpret_cast<float2*>(dka + 8 * dn + 2 * t) = make_float2(dk[dn][0], dk[dn][1]);
                    *reinterpret_cast<float2*>(dva + 8 * dn + 2 * t) = make_float2(dv[dn][0], dv[dn][1]);
                }
                if (jb < L) {
                    *reinterpret_cast<float2*>(dkb + 8 * dn + 2 * t) = make_float2(dk[dn][2], dk[dn][3]);
                    *reinterpret_cast<float2*>(dvb + 8 * dn + 2 * t) = make_float2(dv[dn][2], dv[dn][3]);
                }
            }
        }
    }
}

}  // namespace pr
