// pixelrec_b200 -- attention core for LONG sequences (64 < L <= 256): device code.
//   replaces the attention core of HF CLIPVisionModel's encoder layers when the item encoder is a ViT with more than
//   64 tokens (ViT-B/16: 197 tokens, dh = 64; REC/model/load.py:90-99, REC/model/PixelNet/mosasrec.py:69) and the same
//   core of MultiHeadAttention.forward (REC/model/layers.py:590-612) for MAX_ITEM_LIST_LENGTH > 64.
//
// Strict fp32 (FFMA), no dropout, optional causal / key-padding mask with the reference's additive -1e9 semantics.
// One CTA per (sequence, head) at a time; the two [L x dh] matrices a phase contracts against stay in shared memory
// (rows padded by one float4 -> conflict-free LDS.128 when lanes walk rows), every warp owns 4 rows of the other side:
//   forward      K,V resident; per 4 query rows: S = Q K^T (lane <-> keys), softmax, O = P V (lane <-> columns)
//   backward A   K,V resident; per 4 query rows: S, dP = dO V^T, dS -> dQ = dS K; also delta_i = <dO_i, O_i>
//   backward B   Q,dO resident; per 4 key rows:  S^T, dP^T, P^T, dS^T -> dV = P^T dO, dK = dS^T Q
// P is never written to HBM: the backward recomputes it from the saved log-sum-exp.
//
// This file is also compiled for the HOST by tests/emu (PR_EMU: threads + barriers stand in for a CTA) so that the
// indexing is checked against the oracle without a GPU; keep it free of inline PTX.
#pragma once

namespace pr {

constexpr int AL_WARPS = 8;      // warps per CTA
constexpr int AL_ROWS = 4;       // rows of the streamed side per warp pass
constexpr int AL_JJ = 8;         // resident rows per lane -> L <= 256
constexpr int AL_THREADS = AL_WARPS * 32;

struct LongAttnArgs {
    const float* q; const float* k; const float* v; long long ld;      // fused qkv rows: element (b, i, head*dh + d) at [(b*L+i)*ld + head*dh + d]
    const long long* key_ids;                                          // [B, L] or null: key j of sequence b is valid iff key_ids != 0
    int B, L, h, dh, causal;
    float scale;                                                       // 1 / sqrt(dh)
    float* ctx; float* lse;                                            // forward out: ctx [B, L, h*dh], lse [B*h, L]
    const float* ctx_in; const float* dctx;                            // backward in
    float* delta;                                                      // [B*h, L] scratch written by phase A, read by phase B
    float* dq; float* dk; float* dv; long long ld_grad;
};

__device__ __forceinline__ float al_dot4(const float4& a, const float4& b, float acc) {
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    return fmaf(a.w, b.w, acc);
}
__device__ __forceinline__ void al_axpy4(float4& o, float p, const float4& x) {
    o.x = fmaf(p, x.x, o.x);
    o.y = fmaf(p, x.y, o.y);
    o.z = fmaf(p, x.z, o.z);
    o.w = fmaf(p, x.w, o.w);
}

// shared-memory carve-up (float4 units): two resident matrices [L][dh4+1], per-warp row buffers, per-warp weight tiles
struct LongSmem {
    float4* M0; float4* M1;          // resident [L][RS]
    float4* rowA; float4* rowB;      // this warp's 4 streamed rows, two operands: [AL_ROWS][dh4]
    float* w0; float* w1;            // this warp's weights [Lp][AL_ROWS] (row r fastest): P / dS
    float* aux0; float* aux1;        // [Lp] per-CTA vectors (phase B: lse, delta of the resident side)
    int RS;
};
__host__ __device__ inline size_t long_smem_float4(int L, int dh) {
    const int dh4 = dh / 4, RS = dh4 + 1, Lp = (L + 3) / 4 * 4;
    return (size_t)2 * L * RS + (size_t)AL_WARPS * 2 * AL_ROWS * dh4 + (size_t)AL_WARPS * 2 * Lp + (size_t)2 * (Lp / 4);
}
__device__ __forceinline__ LongSmem long_smem_carve(float4* base, int L, int dh, int warp) {
    const int dh4 = dh / 4, Lp = (L + 3) / 4 * 4;
    LongSmem s;
    s.RS = dh4 + 1;
    s.M0 = base;
    s.M1 = s.M0 + (size_t)L * s.RS;
    float4* rows = s.M1 + (size_t)L * s.RS;
    s.rowA = rows + (size_t)warp * 2 * AL_ROWS * dh4;
    s.rowB = s.rowA + AL_ROWS * dh4;
    float4* wts = rows + (size_t)AL_WARPS * 2 * AL_ROWS * dh4;
    s.w0 = reinterpret_cast<float*>(wts + (size_t)warp * 2 * Lp);
    s.w1 = s.w0 + 4 * Lp;
    float4* aux = wts + (size_t)AL_WARPS * 2 * Lp;
    s.aux0 = reinterpret_cast<float*>(aux);
    s.aux1 = s.aux0 + Lp;
    return s;
}

// whole CTA: resident[r][c] = src[r*ld + 4c]  for r < L, c < dh4
__device__ __forceinline__ void al_load_resident(float4* dst, int RS, const float* src, long long ld, int L, int dh4) {
    for (int idx = threadIdx.x; idx < L * dh4; idx += AL_THREADS) {
        const int r = idx / dh4, c = idx - r * dh4;
        dst[r * RS + c] = PR_LDG4(reinterpret_cast<const float4*>(src + (long long)r * ld) + c);
    }
}
// one warp: rows[r][c] = src[min(i0+r, L-1)*ld + 4c]
__device__ __forceinline__ void al_load_rows(float4* dst, const float* src, long long ld, int i0, int L, int dh4, int lane) {
    for (int idx = lane; idx < AL_ROWS * dh4; idx += 32) {
        const int r = idx / dh4, c = idx - r * dh4;
        const int i = min(i0 + r, L - 1);
        dst[idx] = PR_LDG4(reinterpret_cast<const float4*>(src + (long long)i * ld) + c);
    }
}
// out[r] (float4 column c of this lane's group) = sum_j w[j][r] * M[j][c]; lanes are (group g, column c), the groups
// split the rows j and are summed with xor-shuffles.  Every lane returns the full sums.
__device__ __forceinline__ void al_weighted_sum(const float* w, const float4* M, int RS, int L, int dh4, int lane,
                                                float4 (&o)[AL_ROWS]) {
    const int g = lane / dh4, c = lane - g * dh4, ng = 32 / dh4;
#pragma unroll
    for (int r = 0; r < AL_ROWS; ++r) o[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = g; j < L; j += ng) {
        const float4 p = *reinterpret_cast<const float4*>(w + 4 * j);
        const float4 x = M[j * RS + c];
        al_axpy4(o[0], p.x, x);
        al_axpy4(o[1], p.y, x);
        al_axpy4(o[2], p.z, x);
        al_axpy4(o[3], p.w, x);
    }
    for (int off = dh4; off < 32; off <<= 1) {
#pragma unroll
        for (int r = 0; r < AL_ROWS; ++r) {
            o[r].x += __shfl_xor_sync(0xffffffffu, o[r].x, off);
            o[r].y += __shfl_xor_sync(0xffffffffu, o[r].y, off);
            o[r].z += __shfl_xor_sync(0xffffffffu, o[r].z, off);
            o[r].w += __shfl_xor_sync(0xffffffffu, o[r].w, off);
        }
    }
}
__device__ __forceinline__ float al_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float al_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ============================================================================================ forward
__global__ void __launch_bounds__(AL_THREADS, 1) attn_long_fwd_kernel(const LongAttnArgs A) {
    PR_DYN_SMEM_F4(smem4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = A.L, dh4 = A.dh / 4;
    const LongSmem S = long_smem_carve(smem4, L, A.dh, warp);
    const long long Dm = (long long)A.h * A.dh;
    const long long n_items = (long long)A.B * A.h;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const long long b = item / A.h;
        const int hd = (int)(item - b * A.h);
        const long long base = b * L * A.ld + (long long)hd * A.dh;
        __syncthreads();                                       // the previous item's K/V are no longer read
        al_load_resident(S.M0, S.RS, A.k + base, A.ld, L, dh4);
        al_load_resident(S.M1, S.RS, A.v + base, A.ld, L, dh4);
        __syncthreads();
        bool kvalid[AL_JJ];
#pragma unroll
        for (int jj = 0; jj < AL_JJ; ++jj) {
            const int j = lane + 32 * jj;
            kvalid[jj] = (j < L) && (A.key_ids == nullptr || A.key_ids[b * L + j] != 0);
        }
        for (int i0 = warp * AL_ROWS; i0 < L; i0 += AL_WARPS * AL_ROWS) {
            al_load_rows(S.rowA, A.q + base, A.ld, i0, L, dh4, lane);
            __syncwarp();
            float acc[AL_ROWS][AL_JJ];
#pragma unroll
            for (int r = 0; r < AL_ROWS; ++r)
#pragma unroll
                for (int jj = 0; jj < AL_JJ; ++jj) acc[r][jj] = 0.f;
            for (int c = 0; c < dh4; ++c) {
                float4 qv[AL_ROWS];
#pragma unroll
                for (int r = 0; r < AL_ROWS; ++r) qv[r] = S.rowA[r * dh4 + c];
#pragma unroll
                for (int jj = 0; jj < AL_JJ; ++jj) {
                    if (32 * jj < L) {                           // warp-uniform: skips the key blocks beyond L
                        const float4 kv = S.M0[min(lane + 32 * jj, L - 1) * S.RS + c];
#pragma unroll
                        for (int r = 0; r < AL_ROWS; ++r) acc[r][jj] = al_dot4(qv[r], kv, acc[r][jj]);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < AL_ROWS; ++r) {
                const int i = i0 + r;
                float mx = -INFINITY;
#pragma unroll
                for (int jj = 0; jj < AL_JJ; ++jj) {
                    const int j = lane + 32 * jj;
                    const bool ok = kvalid[jj] && (!A.causal || j <= i);
                    float s = acc[r][jj] * A.scale + (ok ? 0.0f : -1e9f);
                    if (j >= L) s = -INFINITY;
                    acc[r][jj] = s;
                    mx = fmaxf(mx, s);
                }
                mx = al_warp_max(mx);
                float sum = 0.f;
#pragma unroll
                for (int jj = 0; jj < AL_JJ; ++jj) {
                    const float e = (lane + 32 * jj < L) ? expf(acc[r][jj] - mx) : 0.f;
                    acc[r][jj] = e;
                    sum += e;
                }
                sum = al_warp_sum(sum);
                const float inv = 1.0f / sum;
#pragma unroll
                for (int jj = 0; jj < AL_JJ; ++jj) {
                    const int j = lane + 32 * jj;
                    if (j < L) S.w0[4 * j + r] = acc[r][jj] * inv;
                }
                if (lane == 0 && i < L) A.lse[item * L + i] = mx + logf(sum);
            }
            __syncwarp();
            float4 o[AL_ROWS];
            al_weighted_sum(S.w0, S.M1, S.RS, L, dh4, lane, o);           // O = P V
            if (lane < dh4) {
#pragma unroll
                for (int r = 0; r < AL_ROWS; ++r)
                    if (i0 + r < L)
                        *reinterpret_cast<float4*>(A.ctx + (b * L + i0 + r) * Dm + (long long)hd * A.dh + 4 * lane) = o[r];
            }
            __syncwarp();                                      // w0 / rowA are rewritten by the next pass
        }
    }
}

// ============================================================================================ backward, phase A: dQ (+ delta)
__global__ void __launch_bounds__(AL_THREADS, 1) attn_long_bwd_dq_kernel(const LongAttnArgs A) {
    PR_DYN_SMEM_F4(smem4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = A.L, dh4 = A.dh / 4;
    const LongSmem S = long_smem_carve(smem4, L, A.dh, warp);
    const long long Dm = (long long)A.h * A.dh;
    const long long n_items = (long long)A.B * A.h;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const long long b = item / A.h;
        const int hd = (int)(item - b * A.h);
        const long long base = b * L * A.ld + (long long)hd * A.dh;
        const long long obase = b * L * Dm + (long long)hd * A.dh;
        __syncthreads();
        al_load_resident(S.M0, S.RS, A.k + base, A.ld, L, dh4);
        al_load_resident(S.M1, S.RS, A.v + base, A.ld, L, dh4);
        __syncthreads();
        bool kvalid[AL_JJ];
#pragma unroll
        for (int jj = 0; jj < AL_JJ; ++jj) {
            const int j = lane + 32 * jj;
            kvalid[jj] = (j < L) && (A.key_ids == nullptr || A.key_ids[b * L + j] != 0);
        }
        for (int i0 = warp * AL_ROWS; i0 < L; i0 += AL_WARPS * AL_ROWS) {
            al_load_rows(S.rowA, A.q + base, A.ld, i0, L, dh4, lane);
            al_load_rows(S.rowB, A.dctx + obase, Dm, i0, L, dh4, lane);
            __syncwarp();
            // delta_i = <dO_i, O_i>
            float delta[AL_ROWS];
#pragma unroll
            for (int r = 0; r < AL_ROWS; ++r) {
                const int i = min(i0 + r, L - 1);
                float part = 0.f;
                for (int c = lane; c < dh4; c += 32) {
                    const float4 o4 = PR_LDG4(reinterpret_cast<const float4*>(A.ctx_in + obase + (long long)i * Dm) + c);
                    part = al_dot4(S.rowB[r * dh4 + c], o4, part);
                }
                delta[r] = al_warp_sum(part);
                if (lane == 0 && i0 + r < L) A.delta[item * L + i0 + r] = delta[r];
            }
            float acs[AL_ROWS][AL_JJ], acd[AL_ROWS][AL_JJ];
#pragma unroll
            for (int r = 0; r < AL_ROWS; ++r)
#pragma unroll
                for (int jj = 0; jj < AL_JJ; ++jj) { acs[r][jj] = 0.f; acd[r][jj] = 0.f; }
            for (int c = 0; c < dh4; ++c) {
                float4 qv[AL_ROWS], gv[AL_ROWS];
#pragma unroll
                for (int r = 0; r < AL_ROWS; ++r) { qv[r] = S.rowA[r * dh4 + c]; gv[r] = S.rowB[r * dh4 + c]; }
#pragma unroll
                for (int jj = 0; jj < AL_JJ; ++jj) {
                    if (32 * jj < L) {
                        const int jr = min(lane + 32 * jj, L - 1) * S.RS + c;
                        const float4 kv = S.M0[jr], vv = S.M1[jr];
#pragma unroll
                        for (int r = 0; r < AL_ROWS; ++r) {
                            acs[r][jj] = al_dot4(qv[r], kv, acs[r][jj]);
                            acd[r][jj] = al_dot4(gv[r], vv, acd[r][jj]);
                        }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < AL_ROWS; ++r) {
                const int i = i0 + r;
                const float lse_i = A.lse[item * L + min(i, L - 1)];
#pragma unroll
                for (int jj = 0; jj < AL_JJ; ++jj) {
                    const int j = lane + 32 * jj;
                    if (j < L) {
                        const bool ok = kvalid[jj] && (!A.causal || j <= i);
                        const float s = acs[r][jj] * A.scale + (ok ? 0.0f : -1e9f);
                        const float p = expf(s - lse_i);
                        S.w0[4 * j + r] = p * (acd[r][jj] - delta[r]) * A.scale;       // dS
                    }
                }
            }
            __syncwarp();
            float4 o[AL_ROWS];
            al_weighted_sum(S.w0, S.M0, S.RS, L, dh4, lane, o);           // dQ = dS K
            if (lane < dh4) {
#pragma unroll
                for (int r = 0; r < AL_ROWS; ++r)
                    if (i0 + r < L)
                        *reinterpret_cast<float4*>(A.dq + (b * L + i0 + r) * A.ld_grad + (long long)hd * A.dh + 4 * lane) = o[r];
            }
            __syncwarp();
        }
    }
}

// ============================================================================================ backward, phase B: dK, dV
__global__ void __launch_bounds__(AL_THREADS, 1) attn_long_bwd_dkv_kernel(const LongAttnArgs A) {
    PR_DYN_SMEM_F4(smem4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = A.L, dh4 = A.dh / 4;
    const LongSmem S = long_smem_carve(smem4, L, A.dh, warp);
    const long long Dm = (long long)A.h * A.dh;
    const long long n_items = (long long)A.B * A.h;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const long long b = item / A.h;
        const int hd = (int)(item - b * A.h);
        const long long base = b * L * A.ld + (long long)hd * A.dh;
        const long long obase = b * L * Dm + (long long)hd * A.dh;
        __syncthreads();
        al_load_resident(S.M0, S.RS, A.q + base, A.ld, L, dh4);           // resident side = queries
        al_load_resident(S.M1, S.RS, A.dctx + obase, Dm, L, dh4);
        for (int i = threadIdx.x; i < L; i += AL_THREADS) {
            S.aux0[i] = A.lse[item * L + i];
            S.aux1[i] = A.delta[item * L + i];
        }
        __syncthreads();
        for (int j0 = warp * AL_ROWS; j0 < L; j0 += AL_WARPS * AL_ROWS) {
            al_load_rows(S.rowA, A.k + base, A.ld, j0, L, dh4, lane);
            al_load_rows(S.rowB, A.v + base, A.ld, j0, L, dh4, lane);
            __syncwarp();
            bool kvalid[AL_ROWS];
#pragma unroll
            for (int r = 0; r < AL_ROWS; ++r) {
                const int j = j0 + r;
                kvalid[r] = (j < L) && (A.key_ids == nullptr || A.key_ids[b * L + j] != 0);
            }
            float acs[AL_ROWS][AL_JJ], acd[AL_ROWS][AL_JJ];
#pragma unroll
            for (int r = 0; r < AL_ROWS; ++r)
#pragma unroll
                for (int ii = 0; ii < AL_JJ; ++ii) { acs[r][ii] = 0.f; acd[r][ii] = 0.f; }
            for (int c = 0; c < dh4; ++c) {
                float4 kv[AL_ROWS], vv[AL_ROWS];
#pragma unroll
                for (int r = 0; r < AL_ROWS; ++r) { kv[r] = S.rowA[r * dh4 + c]; vv[r] = S.rowB[r * dh4 + c]; }
#pragma unroll
                for (int ii = 0; ii < AL_JJ; ++ii) {
                    if (32 * ii < L) {
                        const int ir = min(lane + 32 * ii, L - 1) * S.RS + c;
                        const float4 qv = S.M0[ir], gv = S.M1[ir];
#pragma unroll
                        for (int r = 0; r < AL_ROWS; ++r) {
                            acs[r][ii] = al_dot4(kv[r], qv, acs[r][ii]);
                            acd[r][ii] = al_dot4(vv[r], gv, acd[r][ii]);
                        }
                    }
                }
            }
#pragma unroll
            for (int ii = 0; ii < AL_JJ; ++ii) {
                const int i = lane + 32 * ii;
                if (i < L) {
                    const float lse_i = S.aux0[i], delta_i = S.aux1[i];
#pragma unroll
                    for (int r = 0; r < AL_ROWS; ++r) {
                        const int j = j0 + r;
                        const bool ok = kvalid[r] && (!A.causal || j <= i);
                        const float s = acs[r][ii] * A.scale + (ok ? 0.0f : -1e9f);
                        const float p = (j < L) ? expf(s - lse_i) : 0.f;
                        S.w0[4 * i + r] = p;                                            // P^T
                        S.w1[4 * i + r] = p * (acd[r][ii] - delta_i) * A.scale;         // dS^T
                    }
                }
            }
            __syncwarp();
            float4 o[AL_ROWS];
            al_weighted_sum(S.w0, S.M1, S.RS, L, dh4, lane, o);           // dV_j = sum_i P_ij dO_i
            if (lane < dh4) {
#pragma unroll
                for (int r = 0; r < AL_ROWS; ++r)
                    if (j0 + r < L)
                        *reinterpret_cast<float4*>(A.dv + (b * L + j0 + r) * A.ld_grad + (long long)hd * A.dh + 4 * lane) = o[r];
            }
            al_weighted_sum(S.w1, S.M0, S.RS, L, dh4, lane, o);           // dK_j = sum_i dS_ij Q_i
            if (lane < dh4) {
#pragma unroll
                for (int r = 0; r < AL_ROWS; ++r)
                    if (j0 + r < L)
                        *reinterpret_cast<float4*>(A.dk + (b * L + j0 + r) * A.ld_grad + (long long)hd * A.dh + 4 * lane) = o[r];
            }
            __syncwarp();
        }
    }
}

}  // namespace pr
