// pixelrec_b200 -- table row kernels for sm_100a:
//   K1  pr_gather_rows_f32       (REC/model/IDNet/sasrec.py:31,68  nn.Embedding forward)
//   K2  pr_scatter_plan / pr_scatter_add_rows_f32  (autograd embedding_dense_backward of sasrec.py:68)
//   K10 pr_adamw_rows_f32 / pr_adamw_dense_f32     (trainer/trainer.py:100-103,125 torch.optim.AdamW)
// All of this is HBM-bound byte/row movement: 128-bit accesses, rows staged through shared memory
// by the TMA engine (cp.async.bulk) for the gather, grids sized in multiples of the SM count.
#include <algorithm>

#include "common.cuh"

#define PR_DYN_SMEM_BYTES(name) extern __shared__ __align__(128) unsigned char name[]
#include "rows_ring.cuh"
#include "rows_plan.cuh"

namespace pr {

// =====================================================================================
// K1 gather, path 1: LDG.128 / STG.128, one warp per 4 rows, all loads issued before the stores
// =====================================================================================
constexpr int G_ROWS = 4;  // rows in flight per warp iteration

__global__ void __launch_bounds__(256) gather_rows_ldg_kernel(const float4* __restrict__ W, long long N, int D4,
                                                              const long long* __restrict__ idx, long long R,
                                                              float4* __restrict__ out, int* __restrict__ status) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long ngroups = (R + G_ROWS - 1) / G_ROWS;
    for (long long g = warp; g < ngroups; g += nwarps) {
        const long long r0 = g * G_ROWS;
        const float4* src[G_ROWS];
        bool live[G_ROWS];
#pragma unroll
        for (int j = 0; j < G_ROWS; ++j) {
            live[j] = (r0 + j) < R;
            long long id = live[j] ? __ldg(idx + r0 + j) : 0;
            const bool ok = (id >= 0) && (id < N);
            if (live[j] && !ok) {
                if (status && lane == 0) atomicOr(status, 1);
            }
            src[j] = ok ? (W + id * (long long)D4) : nullptr;
        }
        for (int c = lane; c < D4; c += 32) {
            float4 v[G_ROWS];
#pragma unroll
            for (int j = 0; j < G_ROWS; ++j) v[j] = src[j] ? __ldg(src[j] + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < G_ROWS; ++j)
                if (live[j]) stg_stream(out + (r0 + j) * (long long)D4 + c, v[j]);
        }
    }
}

// =====================================================================================
// K1 gather, path 2: TMA engine.  One warp per CTA drives a ring of shared-memory stages:
//   lane l issues cp.async.bulk global->shared for row l of the stage (2 KB at D=512), the
//   stage's mbarrier counts the bytes, then ONE cp.async.bulk shared->global writes the whole
//   stage (its rows are contiguous in `out`).  No data ever touches registers.
// =====================================================================================
constexpr int GB_STAGES = 6;
constexpr int GB_STAGE_BYTES = 16 * 1024;  // target bytes per stage (8 rows at D=512)

__global__ void __launch_bounds__(32) gather_rows_bulk_kernel(const float* __restrict__ W, long long N, int D,
                                                              const long long* __restrict__ idx, long long R,
                                                              float* __restrict__ out, int rows_per_stage,
                                                              int* __restrict__ status) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[GB_STAGES];
    const int lane = threadIdx.x;
    const uint32_t row_bytes = (uint32_t)D * 4u;
    const uint32_t stage_bytes = row_bytes * (uint32_t)rows_per_stage;
    if (lane == 0) {
        for (int s = 0; s < GB_STAGES; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    __syncwarp();
    const long long nchunks = (R + rows_per_stage - 1) / rows_per_stage;
    long long n_my = 0;
    if ((long long)blockIdx.x < nchunks) n_my = (nchunks - 1 - blockIdx.x) / gridDim.x + 1;
    // ids of the next chunk are prefetched one iteration ahead so the bulk issue never waits on them
    long long id_next = -1;
    if (n_my > 0) {
        const long long r = (long long)blockIdx.x * rows_per_stage + lane;
        if (lane < rows_per_stage && r < R) id_next = __ldg(idx + r);
    }

    for (long long it = 0; it < n_my + (GB_STAGES - 1); ++it) {
        // ---- drain: chunk jt landed -> one bulk store of the whole stage
        const long long jt = it - (GB_STAGES - 1);
        if (jt >= 0) {
            const int s = (int)(jt % GB_STAGES);
            const uint32_t parity = (uint32_t)((jt / GB_STAGES) & 1);
            mbar_wait(&full_bar[s], parity);
            const long long chunk = (long long)blockIdx.x + jt * gridDim.x;
            const long long r0 = chunk * rows_per_stage;
            const int nrows = (int)min((long long)rows_per_stage, R - r0);
            if (lane == 0) {
                bulk_s2g(out + r0 * (long long)D, smem_raw + (size_t)s * stage_bytes, (uint32_t)nrows * row_bytes);
                bulk_commit();
            }
        }
        // ---- fill: issue the row loads of chunk `it`
        if (it < n_my) {
            const int s = (int)(it % GB_STAGES);
            // the stage was last read by the store committed one loop iteration ago (2 groups back now)
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
            const long long chunk = (long long)blockIdx.x + it * gridDim.x;
            const long long r0 = chunk * rows_per_stage;
            const int nrows = (int)min((long long)rows_per_stage, R - r0);
            const long long id = id_next;
            {
                const long long rn = (chunk + gridDim.x) * rows_per_stage + lane;
                id_next = (it + 1 < n_my && lane < rows_per_stage && rn < R) ? __ldg(idx + rn) : -1;
            }
            const bool ok = (lane < nrows) && id >= 0 && id < N;
            const unsigned okmask = __ballot_sync(0xffffffffu, ok);
            unsigned char* slot = smem_raw + (size_t)s * stage_bytes + (size_t)lane * row_bytes;
            if (lane < nrows && !ok) {  // out-of-range id: zero row + flag (reference raises IndexError)
                if (status) atomicOr(status, 1);
                float4* z = reinterpret_cast<float4*>(slot);
                for (int c = 0; c < D / 4; ++c) z[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                fence_proxy_async();
            }
            if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], (uint32_t)__popc(okmask) * row_bytes);
            __syncwarp();
            if (ok) bulk_g2s(slot, W + id * (long long)D, row_bytes, &full_bar[s]);
        }
    }
    if (lane == 0) bulk_wait_all<0>();
}

// =====================================================================================
// K2 segment reduce, LDG form (rows outside the ring kernel's range, A/B under PR_TUNE without bit 256): one warp per
// (unique id, 32*VPL-float4 column block); rows of a run are added sequentially in ascending position -> deterministic,
// matches the oracle bit for bit.  The default path is scatter_add_rows_ring_kernel (rows_ring.cuh).
// =====================================================================================
template <int VPL>
__global__ void __launch_bounds__(256) scatter_add_rows_kernel(const float4* __restrict__ dOut, int D4, int ncb,
                                                               const int* __restrict__ perm,
                                                               const int* __restrict__ uniq_ids,
                                                               const int* __restrict__ seg_start,
                                                               const int* __restrict__ n_uniq, long long max_uniq,
                                                               float scale, float4* __restrict__ out_rows,
                                                               float4* __restrict__ dense_G) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    long long U = *n_uniq;
    if (U > max_uniq) U = max_uniq;
    const long long nwork = U * ncb;
    for (long long wk = warp; wk < nwork; wk += nwarps) {
        const int u = (int)(wk / ncb);
        const int c0 = (int)(wk % ncb) * (32 * VPL) + lane;
        const int s = seg_start[u], e = seg_start[u + 1];
        float4 acc[VPL];
        {
            const float4* row = dOut + (long long)perm[s] * D4;
#pragma unroll
            for (int j = 0; j < VPL; ++j) {
                const int c = c0 + 32 * j;
                acc[j] = (c < D4) ? ldg_stream(row + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        for (int k0 = s + 1; k0 < e; k0 += 32) {
            const int nk = min(32, e - k0);
            const int myp = (lane < nk) ? perm[k0 + lane] : 0;
#pragma unroll 2
            for (int kk = 0; kk < nk; ++kk) {
                const float4* row = dOut + (long long)__shfl_sync(0xffffffffu, myp, kk) * D4;
                float4 v[VPL];
#pragma unroll
                for (int j = 0; j < VPL; ++j) {
                    const int c = c0 + 32 * j;
                    v[j] = (c < D4) ? ldg_stream(row + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int j = 0; j < VPL; ++j) {
                    acc[j].x += v[j].x; acc[j].y += v[j].y; acc[j].z += v[j].z; acc[j].w += v[j].w;
                }
            }
        }
        const long long id = uniq_ids[u];
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const int c = c0 + 32 * j;
            if (c < D4) {
                float4 r = acc[j];
                if (scale != 1.0f) { r.x *= scale; r.y *= scale; r.z *= scale; r.w *= scale; }
                if (out_rows) out_rows[(long long)u * D4 + c] = r;
                if (dense_G) dense_G[id * D4 + c] = r;
            }
        }
    }
}

// =====================================================================================
// K10 AdamW (torch.optim.AdamW semantics), dense over every row, gradient looked up through row2slot
// =====================================================================================
struct AdamConsts {
    float decay, beta1, one_m_beta1, beta2, one_m_beta2, step_size, inv_sqrt_bc2, eps, gscale;
};

__device__ __forceinline__ AdamConsts adam_consts(float lr, float b1, float b2, float eps, float wd, float gscale,
                                                  long long step, const long long* step_dev) {
    if (step_dev) step = *step_dev;
    const double bc1 = 1.0 - pow((double)b1, (double)step);
    const double bc2 = 1.0 - pow((double)b2, (double)step);
    AdamConsts c;
    c.decay = 1.0f - lr * wd;
    c.beta1 = b1; c.one_m_beta1 = 1.0f - b1;
    c.beta2 = b2; c.one_m_beta2 = 1.0f - b2;
    c.step_size = (float)((double)lr / bc1);
    c.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    c.eps = eps;
    c.gscale = gscale;
    return c;
}

__device__ __forceinline__ void adam_elem(float& w, float& m, float& v, float g, const AdamConsts& c) {
    g *= c.gscale;
    w *= c.decay;
    m = m * c.beta1 + g * c.one_m_beta1;
    v = v * c.beta2 + (g * g) * c.one_m_beta2;
    const float denom = sqrtf(v) * c.inv_sqrt_bc2 + c.eps;
    w -= c.step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adamw_rows_kernel(float4* __restrict__ W, float4* __restrict__ M,
                                                         float4* __restrict__ V, long long N, int D4,
                                                         const float4* __restrict__ grad_rows,
                                                         int* __restrict__ row2slot, float lr, float b1, float b2,
                                                         float eps, float wd, float gscale, long long step,
                                                         const long long* __restrict__ step_dev) {
    const AdamConsts c = adam_consts(lr, b1, b2, eps, wd, gscale, step, step_dev);
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long row = warp; row < N; row += nwarps) {
        int slot = -1;
        if (row2slot) {
            if (lane == 0) {
                slot = row2slot[row];
                if (slot >= 0) row2slot[row] = -1;
            }
            slot = __shfl_sync(0xffffffffu, slot, 0);
        }
        const float4* g4 = (slot >= 0) ? grad_rows + (long long)slot * D4 : nullptr;
        float4* w4 = W + row * D4;
        float4* m4 = M + row * D4;
        float4* v4 = V + row * D4;
        for (int cidx = lane; cidx < D4; cidx += 32) {
            float4 w = w4[cidx], m = m4[cidx], v = v4[cidx];
            const float4 g = g4 ? ldg_stream(g4 + cidx) : make_float4(0.f, 0.f, 0.f, 0.f);
            adam_elem(w.x, m.x, v.x, g.x, c);
            adam_elem(w.y, m.y, v.y, g.y, c);
            adam_elem(w.z, m.z, v.z, g.z, c);
            adam_elem(w.w, m.w, v.w, g.w, c);
            w4[cidx] = w; m4[cidx] = m; v4[cidx] = v;
        }
    }
}

__global__ void __launch_bounds__(256) adamw_dense_kernel(float* __restrict__ w, const float* __restrict__ g,
                                                          float* __restrict__ m, float* __restrict__ v, long long n,
                                                          float lr, float b1, float b2, float eps, float wd,
                                                          float gscale, long long step,
                                                          const long long* __restrict__ step_dev) {
    const AdamConsts c = adam_consts(lr, b1, b2, eps, wd, gscale, step, step_dev);
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    float4* w4 = reinterpret_cast<float4*>(w);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 ww = w4[i], mm = m4[i], vv = v4[i];
        const float4 gg = g4[i];
        adam_elem(ww.x, mm.x, vv.x, gg.x, c);
        adam_elem(ww.y, mm.y, vv.y, gg.y, c);
        adam_elem(ww.z, mm.z, vv.z, gg.z, c);
        adam_elem(ww.w, mm.w, vv.w, gg.w, c);
        w4[i] = ww; m4[i] = mm; v4[i] = vv;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float ww = w[i], mm = m[i], vv = v[i];
        adam_elem(ww, mm, vv, g[i], c);
        w[i] = ww; m[i] = mm; v[i] = vv;
    }
}

}  // namespace pr

using namespace pr;

// ------------------------------------------------------------------------------------------ C ABI
extern "C" int pr_gather_rows_f32(const float* W, int64_t N, int64_t D, const int64_t* idx, int64_t R, float* out,
                                  int32_t* status, int impl, pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PR_CHECK_ARG(N > 0 && D > 0 && R >= 0, "pr_gather_rows_f32: bad shape N=%lld D=%lld R=%lld", (long long)N,
                 (long long)D, (long long)R);
    PR_CHECK_ARG(D % 4 == 0, "pr_gather_rows_f32: D=%lld must be a multiple of 4", (long long)D);
    PR_CHECK_ARG(impl >= 0 && impl <= 2, "pr_gather_rows_f32: impl must be 0,1,2");
    if (R == 0) return PR_OK;
    PR_CHECK_ARG(W && idx && out, "pr_gather_rows_f32: null pointer");
    PR_CHECK_ARG(aligned16(W) && aligned16(out), "pr_gather_rows_f32: W/out must be 16-byte aligned");
    const int sms = sm_count();
    const long long row_bytes = D * 4;
    if (impl == 0) impl = (row_bytes >= 512 && row_bytes <= 32 * 1024) ? 2 : 1;
    if (impl == 2) {
        PR_CHECK_ARG(row_bytes <= 32 * 1024, "pr_gather_rows_f32: bulk path needs rows <= 32 KiB");
        int rps = (int)(GB_STAGE_BYTES / row_bytes);
        if (rps < 1) rps = 1;
        if (rps > 32) rps = 32;
        const size_t smem = (size_t)GB_STAGES * rps * row_bytes;
        if (smem > 48 * 1024)
            PR_CUDA_CALL(cudaFuncSetAttribute(gather_rows_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem));
        const long long nchunks = (R + rps - 1) / rps;
        int ctas_per_sm = (int)((220 * 1024) / (smem + 1024));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        if (ctas_per_sm > 16) ctas_per_sm = 16;
        const int grid = (int)std::min<long long>(nchunks, (long long)sms * ctas_per_sm);
        gather_rows_bulk_kernel<<<grid, 32, smem, stream>>>(W, N, (int)D, (const long long*)idx, R, out, rps, status);
        PR_CUDA_LAUNCH_CHECK("gather_rows_bulk_kernel");
    } else {
        const long long ngroups = (R + G_ROWS - 1) / G_ROWS;
        const int grid = (int)std::min<long long>((ngroups + 7) / 8, (long long)sms * 8);
        gather_rows_ldg_kernel<<<grid, 256, 0, stream>>>((const float4*)W, N, (int)(D / 4), (const long long*)idx, R,
                                                         (float4*)out, status);
        PR_CUDA_LAUNCH_CHECK("gather_rows_ldg_kernel");
    }
    return PR_OK;
}

namespace {
struct PlanLayout {
    size_t keys0, keys1, tmpv, tile_hist, tile_sum, total;
    int T;
};
PlanLayout plan_layout(int64_t R) {
    PlanLayout L;
    const size_t r = (size_t)((R + 63) / 64 * 64);
    L.T = (int)((R + RS_TILE - 1) / RS_TILE);
    if (L.T < 1) L.T = 1;
    size_t o = 0;
    auto take = [&](size_t n) { size_t at = o; o += (n * 4 + 255) / 256 * 256; return at; };
    L.keys0 = take(r);
    L.keys1 = take(r);
    L.tmpv = take(r);
    L.tile_hist = take((size_t)RS_MAX_BINS * L.T);
    L.tile_sum = take((size_t)L.T + 1);
    L.total = o;
    return L;
}
}  // namespace

extern "C" size_t pr_scatter_plan_workspace_bytes(int64_t R, int64_t N) {
    (void)N;
    if (R < 0) return 0;
    return plan_layout(R).total;
}

extern "C" int pr_scatter_plan(const int64_t* idx, int64_t R, int64_t N, int64_t padding_idx, int32_t* perm,
                               int32_t* uniq_ids, int32_t* seg_start, int32_t* n_uniq, int32_t* row2slot,
                               void* workspace, size_t workspace_bytes, int32_t* status, pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PR_CHECK_ARG(R >= 0 && R <= (1 << 24), "pr_scatter_plan: R=%lld outside [0, 2^24]", (long long)R);
    PR_CHECK_ARG(N > 0 && N < 0x7fffffffLL, "pr_scatter_plan: N=%lld outside (0, 2^31)", (long long)N);
    PR_CHECK_ARG(seg_start && n_uniq, "pr_scatter_plan: null output");
    if (R == 0) {
        plan_empty_kernel<<<1, 1, 0, stream>>>(seg_start, n_uniq);
        PR_CUDA_LAUNCH_CHECK("plan_empty_kernel");
        return PR_OK;
    }
    PR_CHECK_ARG(idx && perm && uniq_ids && workspace, "pr_scatter_plan: null pointer");
    const PlanLayout L = plan_layout(R);
    PR_CHECK_ARG(workspace_bytes >= L.total, "pr_scatter_plan: workspace %zu < required %zu", workspace_bytes, L.total);
    char* ws = (char*)workspace;
    uint32_t* kbuf[2] = {(uint32_t*)(ws + L.keys0), (uint32_t*)(ws + L.keys1)};
    int* tmpv = (int*)(ws + L.tmpv);
    uint32_t* tile_hist = (uint32_t*)(ws + L.tile_hist);
    uint32_t* tile_sum = (uint32_t*)(ws + L.tile_sum);
    const int T = L.T;
    int bits = 1;
    while ((1LL << bits) <= N) ++bits;  // keys take values 0..N (N = sentinel)
    const int passes = (bits + 9) / 10;                 // digits of up to 10 bits: 2 passes up to N = 2^20 - 1
    const int rb = (bits + passes - 1) / passes;        // bits per digit, passes * rb >= bits
    const int bins = 1 << rb;
    // ping-pong so that the final pass lands the positions in `perm`
    int* vbuf[2];
    if (passes % 2 == 0) { vbuf[0] = perm; vbuf[1] = tmpv; } else { vbuf[0] = tmpv; vbuf[1] = perm; }
    const int iR = (int)R;
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        const int shift = rb * p;
        // pass 0 makes the keys from the ids on the fly and needs no value array (value of key i = i)
        rs_hist_kernel<<<T, RS_THREADS, 0, stream>>>(p == 0 ? (const long long*)idx : nullptr, N, padding_idx, status, kbuf[cur],
                                                     iR, shift, bins, tile_hist, T);
        scan_single_kernel<<<1, SCAN_THREADS, 0, stream>>>(tile_hist, bins * T, nullptr);
        rs_scatter_kernel<<<T, RS_THREADS, 0, stream>>>(kbuf[cur], p == 0 ? nullptr : vbuf[cur], kbuf[cur ^ 1], vbuf[cur ^ 1], iR,
                                                        shift, bins, tile_hist, T);
        cur ^= 1;
    }
    PR_CUDA_LAUNCH_CHECK("radix sort passes");
    const uint32_t* skeys = kbuf[cur];
    seg_reduce_kernel<<<T, RS_THREADS, 0, stream>>>(skeys, iR, (uint32_t)N, tile_sum);
    scan_single_kernel<<<1, SCAN_THREADS, 0, stream>>>(tile_sum, T, nullptr);
    seg_emit_kernel<<<T, RS_THREADS, 0, stream>>>(skeys, iR, (uint32_t)N, tile_sum, uniq_ids, seg_start, n_uniq, row2slot);
    PR_CUDA_LAUNCH_CHECK("segment kernels");
    return PR_OK;
}

extern "C" int pr_scatter_add_rows_f32(const float* dOut, int64_t R, int64_t D, const int32_t* perm,
                                       const int32_t* uniq_ids, const int32_t* seg_start, const int32_t* n_uniq,
                                       int64_t max_uniq, float scale, float* out_rows, float* dense_G,
                                       pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PR_CHECK_ARG(D > 0 && D % 4 == 0, "pr_scatter_add_rows_f32: D=%lld must be a positive multiple of 4", (long long)D);
    PR_CHECK_ARG(R >= 0 && max_uniq >= 0, "pr_scatter_add_rows_f32: negative size");
    if (R == 0 || max_uniq == 0) return PR_OK;
    PR_CHECK_ARG(dOut && perm && uniq_ids && seg_start && n_uniq, "pr_scatter_add_rows_f32: null pointer");
    PR_CHECK_ARG(out_rows || dense_G, "pr_scatter_add_rows_f32: no output given");
    PR_CHECK_ARG(aligned16(dOut) && aligned16(out_rows) && aligned16(dense_G),
                 "pr_scatter_add_rows_f32: pointers must be 16-byte aligned");
    const int D4 = (int)(D / 4);
    const int sms = sm_count();
    const long long work_hint = std::min<long long>(max_uniq, R);
    if ((tune() & PR_TUNE_SCATTER_RING) && D >= 64 && D <= 1024) {
        // TMA-staged ring (rows_ring.cuh): VPL float4 per lane cover a row, 4 KiB stages, 6 of them per one-warp CTA, ~9 CTAs per SM.
        // Measured (profiles/r02j_scatter_ab.json, r02k_scatter_variants.json): 1.4-2.5x the LDG kernel up to D = 1024; from 8 KiB
        // rows on the LDG kernel's several warps per row are as fast or faster.
        static int variant = -1;            // PR_SCATTER_VARIANT (A/B runs): 1 = 8 KiB stages (4 rows at D = 512), 4 = 4-stage ring,
        if (variant < 0) {                  // 8 = look-ahead queue in registers (ring v1) instead of shared memory (ring v2)
            const char* e = getenv("PR_SCATTER_VARIANT");
            variant = e ? atoi(e) : 0;
        }
        int vpl = 1;
        while (32 * vpl < D4) vpl *= 2;
        const bool big = (variant & 1) && vpl == 4;
        const bool v2 = !(variant & 8) && !big;
        const int nst = v2 ? ((variant & 4) ? 4 : 3) : ((variant & 4) ? 4 : SR_STAGES);
        const int rps = big ? 4 : std::max(1, 8 / vpl);
        const size_t smem = (size_t)nst * rps * D * 4 + SR_BAR_BYTES + (v2 ? SR_Q_BYTES : 0);
        int ctas_per_sm = (int)((220 * 1024) / (smem + 1024));
        ctas_per_sm = std::max(1, std::min(16, ctas_per_sm));
        const long long max_ctas = (long long)sms * ctas_per_sm;
        // runs per group: every CTA should get several groups, and the number of distinct ids is only known on the device --
        // assume a third of the rows (long-tail batches: 0.3-0.55), which errs towards more, smaller groups
        int gr = 32;
        while (gr > 2 && work_hint / (3 * gr) < 4 * max_ctas) gr >>= 1;
        const int grid = (int)std::max<long long>(1, std::min<long long>((work_hint + gr - 1) / gr, max_ctas));
#define PR_LAUNCH_RING(VPL, RPS)                                                                                    \
    do {                                                                                                            \
        if (smem > 48 * 1024)                                                                                       \
            PR_CUDA_CALL(cudaFuncSetAttribute(scatter_add_rows_ring_kernel<VPL, RPS>,                               \
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
        scatter_add_rows_ring_kernel<VPL, RPS><<<grid, 32, smem, stream>>>(dOut, (int)D, gr, perm, uniq_ids, seg_start, \
                                                                          n_uniq, max_uniq, scale, out_rows, dense_G, nst); \
    } while (0)
#define PR_LAUNCH_RING2(VPL, RPS)                                                                                   \
    scatter_add_rows_ring2_kernel<VPL, RPS><<<grid, 32, smem, stream>>>(dOut, (int)D, gr, perm, uniq_ids, seg_start, n_uniq,  \
                                                                       max_uniq, scale, out_rows, dense_G, nst)
        switch (vpl) {
            case 1: if (v2) PR_LAUNCH_RING2(1, 8); else PR_LAUNCH_RING(1, 8); break;
            case 2: if (v2) PR_LAUNCH_RING2(2, 4); else PR_LAUNCH_RING(2, 4); break;
            case 4: if (v2) PR_LAUNCH_RING2(4, 2); else if (big) PR_LAUNCH_RING(4, 4); else PR_LAUNCH_RING(4, 2); break;
            default: if (v2) PR_LAUNCH_RING2(8, 1); else PR_LAUNCH_RING(8, 1); break;
        }
#undef PR_LAUNCH_RING
#undef PR_LAUNCH_RING2
        PR_CUDA_LAUNCH_CHECK("scatter_add_rows_ring_kernel");
        return PR_OK;
    }
#define PR_LAUNCH_SCATTER(VPL)                                                                                      \
    do {                                                                                                            \
        const int ncb = (D4 + 32 * VPL - 1) / (32 * VPL);                                                           \
        const int grid = (int)std::max<long long>(1, std::min<long long>((work_hint * ncb + 7) / 8, (long long)sms * 8)); \
        scatter_add_rows_kernel<VPL><<<grid, 256, 0, stream>>>((const float4*)dOut, D4, ncb, perm, uniq_ids,        \
                                                               seg_start, n_uniq, max_uniq, scale, (float4*)out_rows, \
                                                               (float4*)dense_G);                                   \
    } while (0)
    if (D4 <= 32) PR_LAUNCH_SCATTER(1);
    else if (D4 <= 64) PR_LAUNCH_SCATTER(2);
    else if (D4 <= 128) PR_LAUNCH_SCATTER(4);
    else PR_LAUNCH_SCATTER(8);
#undef PR_LAUNCH_SCATTER
    PR_CUDA_LAUNCH_CHECK("scatter_add_rows_kernel");
    return PR_OK;
}

extern "C" int pr_adamw_rows_f32(float* W, float* M, float* V, int64_t N, int64_t D, const float* grad_rows,
                                 int32_t* row2slot, float lr, float beta1, float beta2, float eps, float weight_decay,
                                 float grad_scale, int64_t step, const int64_t* step_dev, pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PR_CHECK_ARG(N > 0 && D > 0 && D % 4 == 0, "pr_adamw_rows_f32: bad shape N=%lld D=%lld", (long long)N, (long long)D);
    PR_CHECK_ARG(W && M && V, "pr_adamw_rows_f32: null state pointer");
    PR_CHECK_ARG((row2slot == nullptr) || grad_rows, "pr_adamw_rows_f32: row2slot given without grad_rows");
    PR_CHECK_ARG(step >= 1 || step_dev, "pr_adamw_rows_f32: step must be >= 1");
    PR_CHECK_ARG(aligned16(W) && aligned16(M) && aligned16(V) && aligned16(grad_rows), "pr_adamw_rows_f32: alignment");
    const int grid = (int)std::min<long long>((N + 7) / 8, (long long)sm_count() * 8);
    adamw_rows_kernel<<<grid, 256, 0, stream>>>((float4*)W, (float4*)M, (float4*)V, N, (int)(D / 4),
                                                (const float4*)grad_rows, row2slot, lr, beta1, beta2, eps,
                                                weight_decay, grad_scale, step, (const long long*)step_dev);
    PR_CUDA_LAUNCH_CHECK("adamw_rows_kernel");
    return PR_OK;
}

extern "C" int pr_adamw_dense_f32(float* w, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                                  float beta2, float eps, float weight_decay, float grad_scale, int64_t step,
                                  const int64_t* step_dev, pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PR_CHECK_ARG(n >= 0, "pr_adamw_dense_f32: negative n");
    if (n == 0) return PR_OK;
    PR_CHECK_ARG(w && g && m && v, "pr_adamw_dense_f32: null pointer");
    PR_CHECK_ARG(step >= 1 || step_dev, "pr_adamw_dense_f32: step must be >= 1");
    PR_CHECK_ARG(aligned16(w) && aligned16(g) && aligned16(m) && aligned16(v), "pr_adamw_dense_f32: alignment");
    const int grid = (int)std::max<long long>(1, std::min<long long>((n / 4 + 255) / 256, (long long)sm_count() * 8));
    adamw_dense_kernel<<<grid, 256, 0, stream>>>(w, g, m, v, n, lr, beta1, beta2, eps, weight_decay, grad_scale, step,
                                                 (const long long*)step_dev);
    PR_CUDA_LAUNCH_CHECK("adamw_dense_kernel");
    return PR_OK;
}
