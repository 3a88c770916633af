// pixelrec_b200 -- K9 device code that does not depend on inline PTX: tile constants, kernel arguments, the per-thread sorted
// top-k list, the v2 scoring kernel (written against the primitives score.cu defines: mbar_*, tma_load_2d[_mcast],
// umma_tf32, umma_commit[_mcast], tmem_*, cluster_*), the candidate merge and the mask kernels.
// tests/emu compiles this file for the HOST with emulated primitives (tests/emu/emu_tc.h: mbarriers with phases and
// transaction counts, asynchronous TMA incl. multicast, deferred MMAs into an emulated TMEM, concurrent CTAs of a cluster),
// which checks the pipeline protocol and every index computation against the oracle without a GPU.
#pragma once

namespace pr {

constexpr int SC_BM = 128;       // users per tile (UMMA M)
constexpr int SC_BN = 256;       // items per tile (UMMA N)
constexpr int SC_BK = 32;        // floats per k-block = one 128-byte swizzle row
constexpr int SC_STAGES = 4;
constexpr int SC_A_BYTES = SC_BM * SC_BK * 4;            // 16 KiB
constexpr int SC_B_BYTES = SC_BN * SC_BK * 4;            // 32 KiB
constexpr int SC_STAGE_BYTES = SC_A_BYTES + SC_B_BYTES;  // 48 KiB
constexpr int SC_THREADS = 192;
constexpr int SC_TMEM_COLS = 512;

struct ScoreArgs {
    int kblocks;          // D / 32
    int m_tiles, n_tiles, tiles_per_split, n_splits;
    int n_words;          // mask words per row = n_tiles * 8
    const uint32_t* mask; // [m_tiles*128][n_words]
    float* cand_val;      // [m_tiles*128][n_splits][K]   (v2: [m_tiles*128][n_splits*2][K])
    int* cand_idx;
    int cluster;          // v2: CTAs per cluster sharing each table tile by TMA multicast (1 = none)
    const long long* target;   // CE mode: [B_e] target item per row
    long long n_rows;          // CE mode: B_e
    float* ce_part;            // CE mode: [m_tiles*128][n_splits*2][4] = (running max, sum of exp, target logit or -inf, unused)
};

constexpr int SC2_EPI_WARPS = 8;
constexpr int SC2_THREADS = 64 + 32 * SC2_EPI_WARPS;
constexpr int SC2_BAR_BYTES = 128;   // full[4] empty[4] tfull[2] tempty[2] mbarriers + the TMEM base slot, behind the ring

// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, rows 128 B apart, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout [61,64))
__device__ __forceinline__ uint64_t sw128_kmajor_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=tf32 [7,10)=2, B=tf32 [10,13)=2,
// A,B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t tf32_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::f16 with fp16 A and B: format fields [7,10) and [10,13) are 0
__host__ __device__ constexpr uint32_t f16_idesc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// fp32 -> fp16 bit pattern, round to nearest even, saturating to +-65504 (sets *sat); plain integer code so that the host
// emulation converts identically
__host__ __device__ inline uint16_t f32_to_f16_rn(float f, bool* sat) {
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) return (uint16_t)(sign | (ax > 0x7f800000u ? 0x7e00u : 0x7c00u));      // nan / inf
    if (ax >= 0x477ff000u) {                                  // rounds to >= 65520: saturate
        if (sat) *sat = true;
        return (uint16_t)(sign | 0x7bffu);
    }
    if (ax < 0x33000001u) return (uint16_t)sign;              // < 2^-25 (half of the smallest subnormal): zero
    if (ax < 0x38800000u) {                                   // subnormal half: value = m * 2^-24
        const int e = (int)(ax >> 23);                        // biased fp32 exponent, 102..112
        const uint32_t m = (ax & 0x7fffffu) | 0x800000u;      // 24-bit significand
        const int shift = 126 - e;                            // 14..24
        uint32_t h = m >> shift;
        const uint32_t rem = m & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (h & 1u))) ++h;
        return (uint16_t)(sign | h);
    }
    uint32_t h = ((ax >> 23) - 112u) << 10 | ((ax >> 13) & 0x3ffu);
    const uint32_t rem = ax & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;    // carry may bump the exponent: still correct
    return (uint16_t)(sign | h);
}

__global__ void __launch_bounds__(256) score_to_f16_kernel(const float4* __restrict__ src, long long n4, uint2* __restrict__ dst,
                                                           int* __restrict__ status) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    bool sat = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = src[i];
        uint2 o;
        o.x = (uint32_t)f32_to_f16_rn(v.x, &sat) | ((uint32_t)f32_to_f16_rn(v.y, &sat) << 16);
        o.y = (uint32_t)f32_to_f16_rn(v.z, &sat) | ((uint32_t)f32_to_f16_rn(v.w, &sat) << 16);
        dst[i] = o;
    }
    if (sat && status) atomicOr(status, 4);
}

// sorted (descending) insert; strict '>' keeps the earlier (lower) column on ties
template <int K>
__device__ __forceinline__ void topk_insert(float (&val)[K], int (&idx)[K], float v, int c) {
#pragma unroll
    for (int i = K - 1; i > 0; --i) {
        const bool shift = v > val[i - 1];
        const bool here = v > val[i];
        idx[i] = shift ? idx[i - 1] : (here ? c : idx[i]);
        val[i] = shift ? val[i - 1] : (here ? v : val[i]);
    }
    if (v > val[0]) { val[0] = v; idx[0] = c; }
}

// v[j] for a per-lane j without local memory: 31 selects
__device__ __forceinline__ float pick32(const float (&v)[32], int j) {
    float a[16], b[8], c[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (j & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
    const float d0 = (j & 8) ? c[1] : c[0], d1 = (j & 8) ? c[3] : c[2];
    return (j & 16) ? d1 : d0;
}

// (The linear layers of the encoder ran on this pipeline as a MODE 2 in round 1 -- 155-405 TFLOP/s, bound by its uncoalesced
// row-per-thread stores; they now have their own CTA-pair kernel with a TMA-store epilogue, csrc/gemm.cu.)
// MODE 0: masked per-row top-k (eval ranking).  MODE 1: online log-sum-exp over the unmasked catalog + the target item's
// logit = full-catalog softmax cross-entropy (K unused).
// F16: operands are fp16 copies of seq_out / the table (pr_score_prepare_f16): same 10 explicit mantissa bits as TF32 -- and
// rounded to nearest, where the TF32 datapath reads truncated fp32 words -- at twice the MMA rate and half the operand bytes.
// A 128-byte swizzle row then holds 64 elements and one MMA covers K = 16; every byte offset of the pipeline is unchanged.
// ARES (needs F16, D <= 512): the CTA's 128 x D seq_out tile (<= 128 KiB in fp16) is loaded ONCE and stays resident; the
// ring then carries only table tiles (3 stages of 32 KiB).  Without it the tile is re-streamed from L2 for every one of the
// CTA's ~20 table tiles, which at the fp16 MMA rate would exceed the chip's L2 throughput.
template <int K, int MODE = 0, bool F16 = false, bool ARES = false>
__global__ void __launch_bounds__(SC2_THREADS, 1) score_topk2_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                     const __grid_constant__ CUtensorMap tmB,
                                                                     const ScoreArgs a) {
    PR_DYN_SMEM_BYTES(smem_raw);
    // SWIZZLE_128B tiles (TMA destination == UMMA operand) must sit on 1024-byte boundaries
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    static_assert(!ARES || F16, "the resident seq_out tile only fits in fp16");
    constexpr int NST = ARES ? 3 : SC_STAGES;                // ring stages
    constexpr int STB = ARES ? SC_B_BYTES : SC_STAGE_BYTES;  // bytes per stage
    constexpr int B_OFF = ARES ? 0 : SC_A_BYTES;             // table tile inside a stage
    unsigned char* ring = smem + (ARES ? (size_t)a.kblocks * SC_A_BYTES : 0);      // resident A k-blocks come first
    // barriers live behind the ring, at the same offset in every CTA of a cluster (remote arrives address them by offset)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + (size_t)NST * STB);
    uint64_t* empty_bar = full_bar + SC_STAGES;
    uint64_t* tfull_bar = empty_bar + SC_STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* ares_bar = tempty_bar + 2;
    uint32_t* tmem_slot_p = reinterpret_cast<uint32_t*>(ares_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x % a.m_tiles, split = blockIdx.x / a.m_tiles;
    const int t_begin = split * a.tiles_per_split;
    const int t_end = min(a.n_tiles, t_begin + a.tiles_per_split);
    const int n_my = t_end - t_begin;                      // identical in every CTA of a cluster (same split)
    constexpr int BKE = F16 ? 64 : SC_BK;                  // operand elements per 128-byte k-block
    const int CL = a.cluster;
    const uint16_t cl_mask = (uint16_t)((1u << CL) - 1u);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], (uint32_t)CL); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], SC2_EPI_WARPS); }
        if (ARES) mbar_init(ares_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot_p, SC_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                        // peers' barriers are initialised before any remote arrive / multicast
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_p;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            const int rank = (CL > 1) ? (int)cluster_ctarank() : 0;
            const int slice_rows = SC_BN / CL;
            const int slice_bytes = SC_B_BYTES / CL;
            long long it = 0;
            if (ARES && n_my > 0) {                                    // the whole seq_out tile, once
                mbar_arrive_expect_tx(ares_bar, (uint32_t)a.kblocks * SC_A_BYTES);
                for (int kb = 0; kb < a.kblocks; ++kb)
                    tma_load_2d(smem + (size_t)kb * SC_A_BYTES, &tmA, kb * BKE, m_tile * SC_BM, ares_bar);
            }
            for (int t = 0; t < n_my; ++t) {
                const int n0 = (t_begin + t) * SC_BN;
                for (int kb = 0; kb < a.kblocks; ++kb, ++it) {
                    const int s = (int)(it % NST);
                    // CL arrivals: every CTA of the cluster has finished reading stage s of the previous round
                    mbar_wait(&empty_bar[s], (uint32_t)(((it / NST) & 1) ^ 1));
                    mbar_arrive_expect_tx(&full_bar[s], STB);          // (own A box +) CL slices of the table tile
                    unsigned char* st = ring + (size_t)s * STB;
                    if (!ARES) tma_load_2d(st, &tmA, kb * BKE, m_tile * SC_BM, &full_bar[s]);
                    if (CL > 1)
                        tma_load_2d_mcast(st + B_OFF + rank * slice_bytes, &tmB, kb * BKE, n0 + rank * slice_rows,
                                          &full_bar[s], cl_mask);
                    else
                        tma_load_2d(st + B_OFF, &tmB, kb * BKE, n0, &full_bar[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc = F16 ? f16_idesc(SC_BM, SC_BN) : tf32_idesc(SC_BM, SC_BN);
            long long it = 0;
            if (ARES && n_my > 0) {
                mbar_wait(ares_bar, 0);
                tc_fence_after();
            }
            for (int t = 0; t < n_my; ++t) {
                const int buf = t & 1;
                mbar_wait(&tempty_bar[buf], (uint32_t)(((t >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)buf * SC_BN;
                for (int kb = 0; kb < a.kblocks; ++kb, ++it) {
                    const int s = (int)(it % NST);
                    mbar_wait(&full_bar[s], (uint32_t)((it / NST) & 1));
                    tc_fence_after();
                    const uint32_t sb = smem_u32(ring + (size_t)s * STB + B_OFF);
                    const uint32_t sa = ARES ? smem_u32(smem + (size_t)kb * SC_A_BYTES) : smem_u32(ring + (size_t)s * STB);
                    const uint64_t adesc = sw128_kmajor_desc(sa), bdesc = sw128_kmajor_desc(sb);
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {                  // 4 MMAs of 32 operand bytes per 128-byte k-block
                        if constexpr (F16) umma_f16(d_tmem, adesc + 2 * k4, bdesc + 2 * k4, idesc, (kb | k4) ? 1u : 0u);
                        else umma_tf32(d_tmem, adesc + 2 * k4, bdesc + 2 * k4, idesc, (kb | k4) ? 1u : 0u);
                    }
                    if (CL > 1) umma_commit_mcast(&empty_bar[s], cl_mask);   // frees stage s in every CTA that writes into it
                    else umma_commit(&empty_bar[s]);
                }
                umma_commit(&tfull_bar[buf]);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue warps 2..9
        // thread == (user row, column half): warp%4 = TMEM lane quadrant, (warp-2)/4 = which 128 of the tile's 256 columns
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = m_tile * SC_BM + q * 32 + lane;
        {
        const uint32_t* mrow = a.mask + (size_t)row * a.n_words + half * 4;
        if constexpr (MODE == 0) {
        float val[K];
        int idx[K];
#pragma unroll
        for (int i = 0; i < K; ++i) { val[i] = -INFINITY; idx[i] = -1; }
        for (int t = 0; t < n_my; ++t) {
            const int buf = t & 1;
            const int tile = t_begin + t;
            const uint4 mw = *reinterpret_cast<const uint4*>(mrow + tile * 8);
            mbar_wait(&tfull_bar[buf], (uint32_t)((t >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * SC_BN + half * 128);
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                float v[32];
                __syncwarp();                                   // tcgen05.ld is warp-collective (.sync.aligned)
                tmem_ld32(taddr + c * 32, v);
                const uint32_t w = (c == 0) ? mw.x : (c == 1) ? mw.y : (c == 2) ? mw.z : mw.w;
                const float thr = val[K - 1];
                uint32_t bits = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j) bits |= (v[j] > thr) ? (1u << j) : 0u;
                bits &= ~w;                                     // pad column, history, columns >= N
                if (__any_sync(0xffffffffu, bits != 0u)) {
                    const int c0 = tile * SC_BN + half * 128 + c * 32;
                    do {                                        // warp-uniform trip count; lanes without a candidate insert -inf (no-op)
                        const bool has = bits != 0u;
                        const int j = has ? (__ffs((int)bits) - 1) : 0;
                        bits &= bits - 1u;
                        const float x = has ? pick32(v, j) : -INFINITY;
                        topk_insert<K>(val, idx, x, c0 + j);    // ascending j: on ties the lower column stays ahead
                    } while (__any_sync(0xffffffffu, bits != 0u));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);
        }
        const size_t list = (size_t)row * (a.n_splits * 2) + (size_t)(split * 2 + half);
        float* cv = a.cand_val + list * K;
        int* ci = a.cand_idx + list * K;
#pragma unroll
        for (int i = 0; i < K; ++i) { cv[i] = val[i]; ci[i] = idx[i]; }
        } else {
        // ---- CE mode: running (max, sum of exp) over this thread's columns, and the target's logit if it is among them
        float mx = -INFINITY, sum = 0.f, tlogit = -INFINITY;
        const long long tgt = (row < a.n_rows && a.target) ? a.target[row] : -1;      // rows >= B_e are tile padding
        for (int t = 0; t < n_my; ++t) {
            const int buf = t & 1;
            const int tile = t_begin + t;
            const uint4 mw = *reinterpret_cast<const uint4*>(mrow + tile * 8);
            mbar_wait(&tfull_bar[buf], (uint32_t)((t >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * SC_BN + half * 128);
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                float v[32];
                __syncwarp();
                tmem_ld32(taddr + c * 32, v);
                const uint32_t w = (c == 0) ? mw.x : (c == 1) ? mw.y : (c == 2) ? mw.z : mw.w;
                const long long c0 = (long long)tile * SC_BN + half * 128 + c * 32;
                const long long tcol = tgt - c0;
                if (tcol >= 0 && tcol < 32 && !((w >> (int)tcol) & 1u)) tlogit = pick32(v, (int)tcol);
                float cm = -INFINITY;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    v[j] = ((w >> j) & 1u) ? -INFINITY : v[j];
                    cm = fmaxf(cm, v[j]);
                }
                const float mnew = fmaxf(mx, cm);
                if (mnew > -INFINITY) {                         // (all-masked chunks before the first live one: nothing to add)
                    float part = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) part += expf(v[j] - mnew);      // exp(-inf) = 0 for masked columns
                    sum = sum * expf(mx - mnew) + part;
                    mx = mnew;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);
        }
        const size_t list = (size_t)row * (a.n_splits * 2) + (size_t)(split * 2 + half);
        *reinterpret_cast<float4*>(a.ce_part + list * 4) = make_float4(mx, sum, tlogit, 0.f);
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();       // no CTA may exit while a peer can still arrive on its barriers
    if (warp == 1) tmem_dealloc(tmem_base, SC_TMEM_COLS);
}

// merge the n_splits sorted candidate lists of a row: k rounds of warp arg-max (value desc, then item id asc)
// out_bound (optional): max over the row's lists of the LAST value a full list kept (list_len entries each) -- every item that is
// in no list scores at most that (the id-exact mode's completeness bound); -inf when no list is full.
__global__ void __launch_bounds__(128) score_merge_kernel(const float* __restrict__ cand_val, const int* __restrict__ cand_idx,
                                                          int n_cand, long long B_e, int k, float* __restrict__ out_val,
                                                          long long* __restrict__ out_idx, int list_len = 0,
                                                          float* __restrict__ out_bound = nullptr) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= B_e) return;
    constexpr int MAXC = 32;                      // up to 1024 candidates per row
    float v[MAXC];
    int id[MAXC];
    float bound = -INFINITY;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = lane + 32 * i;
        const bool ok = c < n_cand;
        v[i] = ok ? cand_val[row * n_cand + c] : -INFINITY;
        id[i] = ok ? cand_idx[row * n_cand + c] : -1;
        if (id[i] < 0) v[i] = -INFINITY;
        if (out_bound && ok && list_len > 0 && (c % list_len) == list_len - 1 && id[i] >= 0) bound = fmaxf(bound, v[i]);
    }
    if (out_bound) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bound = fmaxf(bound, __shfl_xor_sync(0xffffffffu, bound, o));
        if (lane == 0) out_bound[row] = bound;
    }
    for (int r = 0; r < k; ++r) {
        float bv = -INFINITY;
        int bi = 0x7fffffff, bslot = -1;
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const bool better = (id[i] >= 0) && (v[i] > bv || (v[i] == bv && id[i] < bi));
            if (better) { bv = v[i]; bi = id[i]; bslot = i; }
        }
        float wv = bv;
        int wi = bi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
            if (ov > wv || (ov == wv && oi < wi)) { wv = ov; wi = oi; }
        }
        if (bslot >= 0 && wi == bi && wv == bv) {
#pragma unroll
            for (int i = 0; i < MAXC; ++i)
                if (i == bslot) id[i] = -1;          // consumed (item ids are unique per row)
        }
        if (lane == 0) {
            const bool none = (wi == 0x7fffffff);
            out_val[row * k + r] = none ? -INFINITY : wv;
            out_idx[row * k + r] = none ? -1 : (long long)wi;
        }
    }
}

// CE merge: lse[row] = log sum_c exp(logit[row,c]) over the unmasked catalog from the n_lists partial (max, sum) pairs,
// tgt[row] = the target's logit (-inf if the target is masked or out of range), nll = lse - tgt.  One warp per row.
__global__ void __launch_bounds__(128) score_ce_merge_kernel(const float* __restrict__ part, int n_lists, long long B_e,
                                                             float* __restrict__ lse, float* __restrict__ tgt,
                                                             float* __restrict__ nll) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= B_e) return;
    float m = -INFINITY, tl = -INFINITY;
    for (int i = lane; i < n_lists; i += 32) {
        const float4 p = *reinterpret_cast<const float4*>(part + ((size_t)row * n_lists + i) * 4);
        m = fmaxf(m, p.x);
        tl = fmaxf(tl, p.z);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        tl = fmaxf(tl, __shfl_xor_sync(0xffffffffu, tl, o));
    }
    float s = 0.f;
    for (int i = lane; i < n_lists; i += 32) {
        const float4 p = *reinterpret_cast<const float4*>(part + ((size_t)row * n_lists + i) * 4);
        if (p.x > -INFINITY) s += p.y * expf(p.x - m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        const float l = m + logf(s);
        if (lse) lse[row] = l;
        if (tgt) tgt[row] = tl;
        if (nll) nll[row] = l - tl;
    }
}

// mask bitmap: bit (row, col) set -> score forced to -inf.  base: pad column 0 (optional) and columns >= N
__global__ void __launch_bounds__(256) score_mask_base_kernel(uint32_t* __restrict__ mask, long long rows, int n_words,
                                                              long long N, int mask_col0) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * n_words) return;
    const int w = (int)(i % n_words);
    const long long c0 = (long long)w * 32;
    uint32_t bits = 0;
    if (c0 + 32 > N) bits = (c0 >= N) ? 0xffffffffu : (0xffffffffu << (int)(N - c0));
    if (w == 0 && mask_col0) bits |= 1u;
    mask[i] = bits;
}
__global__ void __launch_bounds__(256) score_mask_hist_kernel(uint32_t* __restrict__ mask, int n_words, long long B_e,
                                                              long long N, const long long* __restrict__ hu,
                                                              const long long* __restrict__ hi, long long n_hist) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_hist) return;
    const long long u = hu[p], i = hi[p];
    if (u < 0 || u >= B_e || i < 0 || i >= N) return;
    atomicOr(mask + u * n_words + (i >> 5), 1u << (i & 31));
}

}  // namespace pr
