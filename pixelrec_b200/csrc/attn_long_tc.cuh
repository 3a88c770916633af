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

}  // namespace pr
