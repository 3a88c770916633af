// pixelrec_b200 -- K2 segment reduce, TMA-staged: device code of pr_scatter_add_rows_f32's default path
// (backward of REC/model/IDNet/sasrec.py:68: autograd's embedding_dense_backward without the dense [N,D] gradient).
//
// One warp per CTA streams the gradient rows of a strided set of run groups through a ring of shared-memory stages:
//   fill   lane l issues ONE cp.async.bulk global->shared for row perm[k + l] of the chunk (2 KB at D=512); the stage's
//          mbarrier counts the bytes; the perm entries of the next SR_AHEAD chunks and the boundaries of the next group
//          are already in registers, so the issue never waits on an index load;
//   drain  nst-1 chunks later the warp adds the staged rows, lane = 4-float column slice, IN ASCENDING SORTED
//          POSITION (the oracle's order: results are bit-identical to oracle.scatter_add_rows), and writes a reduced
//          row whenever a run ends.
// The drain side (one warp: shared-memory loads, adds, run bookkeeping, row stores) is what bounds a CTA, so rings are kept SMALL
// (24 KiB: 6 stages of 2 rows at D=512) and many CTAs share an SM: 16 KiB stages with 2 CTAs per SM measured 0.160 ms, 8 KiB
// stages with 4 CTAs 0.102 ms at C2 / B=4096 (profiles/r02k_scatter_variants.json).
// Rows in flight therefore do not depend on the run structure: a run of one row and a run of 500 duplicates of a hot
// item stream at the same rate (the LDG kernel it replaces walked a run with two loads in flight: 0.40 of HBM peak).
// Written against the primitives rows.cu defines (mbar_*, bulk_g2s, PR_DYN_SMEM_BYTES); tests/emu compiles this file
// for the HOST on emulated mbarriers / asynchronous bulk copies (tests/emu/emu_rows.cpp).
#pragma once

namespace pr {

constexpr int SR_STAGES = 6;
constexpr int SR_BAR_BYTES = 64;   // SR_STAGES mbarriers behind the ring
constexpr int SR_AHEAD = 4;        // chunks of index look-ahead on the fill side

// VPL = float4 per lane (row of up to 128*VPL floats), RPS = rows per stage, gr = runs per group (<= 32)
template <int VPL, int RPS>
__global__ void __launch_bounds__(32) scatter_add_rows_ring_kernel(const float* __restrict__ dOut, int D, int gr,
                                                                   const int* __restrict__ perm,
                                                                   const int* __restrict__ uniq_ids,
                                                                   const int* __restrict__ seg_start,
                                                                   const int* __restrict__ n_uniq, long long max_uniq,
                                                                   float scale, float* __restrict__ out_rows,
                                                                   float* __restrict__ dense_G, int nst) {
    PR_DYN_SMEM_BYTES(smem_raw);
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    const int D4 = D >> 2;
    const uint32_t row_bytes = (uint32_t)D * 4u;
    const uint32_t stage_bytes = row_bytes * (uint32_t)RPS;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)nst * stage_bytes);   // nst <= SR_STAGES ring stages
    if (lane == 0) {
        for (int s = 0; s < nst; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    __syncwarp();
    long long Ull = *n_uniq;
    if (Ull > max_uniq) Ull = max_uniq;
    const int U = (int)Ull;
    const int ngroups = (U + gr - 1) / gr;
    const int G = (int)gridDim.x;
    // sorted position where run i of group g starts (i == gr: where the group ends); runs beyond U clamp to the end
    auto bound = [&](int g, int i) { return seg_start[min(U, g * gr + min(i, gr))]; };

    // ---- fill side: a cursor (group pg, positions [pk, pe) left in it, next group's range) walks this CTA's chunks SR_AHEAD
    // chunks ahead of the bulk-copy issue and loads their perm entries, so the issue never waits on an index load
    // (with one chunk of look-ahead every iteration paid a full load latency: 3.2 TB/s, profiles/r02j_scatter_ab.json)
    int pg = (int)blockIdx.x;
    bool p_live = pg < ngroups;
    int pk = 0, pe = 0, nk = 0, ne = 0;
    if (p_live) {
        pk = bound(pg, 0);
        pe = bound(pg, gr);
        if (pg + G < ngroups) { nk = bound(pg + G, 0); ne = bound(pg + G, gr); }
    }
    int qn[SR_AHEAD], qperm[SR_AHEAD];   // rows of / this lane's row in the next SR_AHEAD chunks (qn == 0: no more chunks)
    auto next_chunk = [&](int& n_out, int& perm_out) {
        n_out = p_live ? min(RPS, pe - pk) : 0;
        perm_out = (lane < n_out) ? perm[pk + lane] : 0;
        if (p_live) {
            pk += n_out;
            if (pk >= pe) {
                pg += G;
                if (pg >= ngroups) {
                    p_live = false;
                } else {
                    pk = nk;
                    pe = ne;
                    if (pg + G < ngroups) { nk = bound(pg + G, 0); ne = bound(pg + G, gr); }
                }
            }
        }
    };
    const int ck0 = pk, ce0 = pe, ckn0 = nk;
    const bool live0 = p_live;
#pragma unroll
    for (int i = 0; i < SR_AHEAD; ++i) next_chunk(qn[i], qperm[i]);
    int ps = 0;
    // ---- drain side: lane l holds the end of run l of group cg; run ci is being accumulated
    bool c_live = live0;
    int cg = (int)blockIdx.x, ck = ck0, ce = ce0, ckn = ckn0, ci = 0;
    int cb1 = 0, cb1n = 0;
    if (c_live) {
        cb1 = bound(cg, lane + 1);
        if (cg + G < ngroups) cb1n = bound(cg + G, lane + 1);
    }
    int cur_end = __shfl_sync(FULL, cb1, 0);
    bool first = true;
    int cs = 0;
    uint32_t cphase = 0;
    float4 acc[VPL];
#pragma unroll
    for (int j = 0; j < VPL; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);

    // One iteration = drain the oldest staged chunk, then issue the chunk held in queue slot (qn_s, qperm_s) and refill that slot
    // with the chunk SR_AHEAD further on.  The slots are used round-robin by the unrolled loop below, so a perm entry is never
    // touched between its load and its use (a shifting register queue MOVed the pending loads every iteration, and every chunk
    // then waited a full memory latency: ncu, profiles/r02k_rowkernels_ncu.md).
    int it = 0;
    auto iteration = [&](int& qn_s, int& qperm_s) -> bool {
        if (it >= nst - 1) {
            if (!c_live) return true;
            const int n = min(RPS, ce - ck);
            mbar_wait(&full_bar[cs], cphase);
            const unsigned char* st = smem_raw + (size_t)cs * stage_bytes;
            float4 v[RPS][VPL];
#pragma unroll
            for (int r = 0; r < RPS; ++r) {
                if (r < n) {
                    const float4* row = reinterpret_cast<const float4*>(st + (size_t)r * row_bytes);
#pragma unroll
                    for (int j = 0; j < VPL; ++j) {
                        const int c = lane + 32 * j;
                        v[r][j] = (c < D4) ? row[c] : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < RPS; ++r) {
                if (r < n) {
                    if (first) {
#pragma unroll
                        for (int j = 0; j < VPL; ++j) acc[j] = v[r][j];
                        first = false;
                    } else {
#pragma unroll
                        for (int j = 0; j < VPL; ++j) {
                            acc[j].x += v[r][j].x; acc[j].y += v[r][j].y; acc[j].z += v[r][j].z; acc[j].w += v[r][j].w;
                        }
                    }
                    ++ck;
                    if (ck == cur_end) {   // run complete: one coalesced row store
                        const long long u = (long long)cg * gr + ci;
                        const long long id = dense_G ? (long long)uniq_ids[u] : 0;
#pragma unroll
                        for (int j = 0; j < VPL; ++j) {
                            const int c = lane + 32 * j;
                            if (c < D4) {
                                float4 o = acc[j];
                                if (scale != 1.0f) { o.x *= scale; o.y *= scale; o.z *= scale; o.w *= scale; }
                                if (out_rows) reinterpret_cast<float4*>(out_rows)[u * D4 + c] = o;
                                if (dense_G) reinterpret_cast<float4*>(dense_G)[id * D4 + c] = o;
                            }
                        }
                        first = true;
                        ++ci;
                        cur_end = __shfl_sync(FULL, cb1, ci & 31);
                    }
                }
            }
            __syncwarp();   // every lane is done reading the stage before it is refilled below
            if (++cs == nst) { cs = 0; cphase ^= 1u; }
            if (ck == ce) {   // group exhausted
                cg += G;
                if (cg >= ngroups) {
                    c_live = false;
                } else {
                    cb1 = cb1n;
                    ck = ckn;
                    ce = __shfl_sync(FULL, cb1, 31);
                    ci = 0;
                    cur_end = __shfl_sync(FULL, cb1, 0);
                    if (cg + G < ngroups) { cb1n = bound(cg + G, lane + 1); ckn = bound(cg + G, 0); }
                }
            }
        }
        if (qn_s > 0) {
            // stage ps held the chunk drained one iteration ago
            unsigned char* st = smem_raw + (size_t)ps * stage_bytes;
            if (lane == 0) mbar_arrive_expect_tx(&full_bar[ps], (uint32_t)qn_s * row_bytes);
            __syncwarp();
            if (lane < qn_s) bulk_g2s(st + (size_t)lane * row_bytes, dOut + (long long)qperm_s * D, row_bytes, &full_bar[ps]);
            if (++ps == nst) ps = 0;
            next_chunk(qn_s, qperm_s);
        }
        ++it;
        return false;
    };
    static_assert(SR_AHEAD == 4, "the loop below is unrolled over the queue slots");
    for (;;) {
        if (iteration(qn[0], qperm[0])) break;
        if (iteration(qn[1], qperm[1])) break;
        if (iteration(qn[2], qperm[2])) break;
        if (iteration(qn[3], qperm[3])) break;
    }
}

// Same pipeline with the index look-ahead held in SHARED memory: the perm entries of the next SR_AHEAD chunks are fetched by 4-byte
// cp.async copies (no register is the destination of a pending load, so neither a queue shift nor the compiler's scheduling can make
// the warp wait on one), the loop is not unrolled (a quarter of the code: the unrolled form stalled on instruction fetch, ncu
// "no_instruction" 1.19 per issue) and the register budget allows 16 one-warp CTAs per SM.
constexpr int SR_Q_BYTES = SR_AHEAD * 32 * 4;   // look-ahead queue behind the barriers: [SR_AHEAD][32] ints

template <int VPL, int RPS>
__global__ void __launch_bounds__(32, 16) scatter_add_rows_ring2_kernel(const float* __restrict__ dOut, int D, int gr,
                                                                        const int* __restrict__ perm,
                                                                        const int* __restrict__ uniq_ids,
                                                                        const int* __restrict__ seg_start,
                                                                        const int* __restrict__ n_uniq, long long max_uniq,
                                                                        float scale, float* __restrict__ out_rows,
                                                                        float* __restrict__ dense_G, int nst) {
    PR_DYN_SMEM_BYTES(smem_raw);
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    const int D4 = D >> 2;
    const uint32_t row_bytes = (uint32_t)D * 4u;
    const uint32_t stage_bytes = row_bytes * (uint32_t)RPS;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)nst * stage_bytes);
    int* qs = reinterpret_cast<int*>(smem_raw + (size_t)nst * stage_bytes + SR_BAR_BYTES);
    if (lane == 0) {
        for (int s = 0; s < nst; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    __syncwarp();
    long long Ull = *n_uniq;
    if (Ull > max_uniq) Ull = max_uniq;
    const int U = (int)Ull;
    const int ngroups = (U + gr - 1) / gr;
    const int G = (int)gridDim.x;
    auto bound = [&](int g, int i) { return seg_start[min(U, g * gr + min(i, gr))]; };

    // ---- fill side cursor
    int pg = (int)blockIdx.x;
    bool p_live = pg < ngroups;
    int pk = 0, pe = 0, nk = 0, ne = 0;
    if (p_live) {
        pk = bound(pg, 0);
        pe = bound(pg, gr);
        if (pg + G < ngroups) { nk = bound(pg + G, 0); ne = bound(pg + G, gr); }
    }
    const int ck0 = pk, ce0 = pe, ckn0 = nk;
    const bool live0 = p_live;
    // chunk the cursor points at: its rows' perm entries start their way into queue slot `slot`; returns its row count
    auto prefetch = [&](int slot) -> int {
        const int n_out = p_live ? min(RPS, pe - pk) : 0;
        if (lane < n_out) cp_async4(qs + slot * 32 + lane, perm + pk + lane);
        cp_async_commit();
        if (p_live) {
            pk += n_out;
            if (pk >= pe) {
                pg += G;
                if (pg >= ngroups) {
                    p_live = false;
                } else {
                    pk = nk;
                    pe = ne;
                    if (pg + G < ngroups) { nk = bound(pg + G, 0); ne = bound(pg + G, gr); }
                }
            }
        }
        return n_out;
    };
    int qn[SR_AHEAD];      // row counts of the next SR_AHEAD chunks (plain arithmetic: shifting them costs nothing)
#pragma unroll
    for (int i = 0; i < SR_AHEAD; ++i) qn[i] = prefetch(i);
    int ps = 0, qslot = 0;
    // ---- drain side
    bool c_live = live0;
    int cg = (int)blockIdx.x, ck = ck0, ce = ce0, ckn = ckn0, ci = 0;
    int cb1 = 0, cb1n = 0;
    if (c_live) {
        cb1 = bound(cg, lane + 1);
        if (cg + G < ngroups) cb1n = bound(cg + G, lane + 1);
    }
    int cur_end = __shfl_sync(FULL, cb1, 0);
    bool first = true;
    int cs = 0;
    uint32_t cphase = 0;
    float4 acc[VPL];
#pragma unroll
    for (int j = 0; j < VPL; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int it = 0;; ++it) {
        if (it >= nst - 1) {
            if (!c_live) break;
            const int n = min(RPS, ce - ck);
            mbar_wait(&full_bar[cs], cphase);
            const unsigned char* st = smem_raw + (size_t)cs * stage_bytes;
#pragma unroll
            for (int r = 0; r < RPS; ++r) {
                if (r < n) {
                    const float4* row = reinterpret_cast<const float4*>(st + (size_t)r * row_bytes);
                    if (first) {
#pragma unroll
                        for (int j = 0; j < VPL; ++j) {
                            const int c = lane + 32 * j;
                            acc[j] = (c < D4) ? row[c] : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        first = false;
                    } else {
#pragma unroll
                        for (int j = 0; j < VPL; ++j) {
                            const int c = lane + 32 * j;
                            if (c < D4) {
                                const float4 v = row[c];
                                acc[j].x += v.x; acc[j].y += v.y; acc[j].z += v.z; acc[j].w += v.w;
                            }
                        }
                    }
                    ++ck;
                    if (ck == cur_end) {   // run complete: one coalesced row store
                        const long long u = (long long)cg * gr + ci;
                        const long long id = dense_G ? (long long)uniq_ids[u] : 0;
#pragma unroll
                        for (int j = 0; j < VPL; ++j) {
                            const int c = lane + 32 * j;
                            if (c < D4) {
                                float4 o = acc[j];
                                if (scale != 1.0f) { o.x *= scale; o.y *= scale; o.z *= scale; o.w *= scale; }
                                if (out_rows) reinterpret_cast<float4*>(out_rows)[u * D4 + c] = o;
                                if (dense_G) reinterpret_cast<float4*>(dense_G)[id * D4 + c] = o;
                            }
                        }
                        first = true;
                        ++ci;
                        cur_end = __shfl_sync(FULL, cb1, ci & 31);
                    }
                }
            }
            __syncwarp();   // every lane is done reading the stage before it is refilled below
            if (++cs == nst) { cs = 0; cphase ^= 1u; }
            if (ck == ce) {   // group exhausted
                cg += G;
                if (cg >= ngroups) {
                    c_live = false;
                } else {
                    cb1 = cb1n;
                    ck = ckn;
                    ce = __shfl_sync(FULL, cb1, 31);
                    ci = 0;
                    cur_end = __shfl_sync(FULL, cb1, 0);
                    if (cg + G < ngroups) { cb1n = bound(cg + G, lane + 1); ckn = bound(cg + G, 0); }
                }
            }
        }
        if (qn[0] > 0) {
            cp_async_wait<SR_AHEAD - 1>();   // this chunk's entries have landed (each lane reads back the one it requested)
            const int myperm = (lane < qn[0]) ? qs[qslot * 32 + lane] : 0;
            unsigned char* st = smem_raw + (size_t)ps * stage_bytes;
            if (lane == 0) mbar_arrive_expect_tx(&full_bar[ps], (uint32_t)qn[0] * row_bytes);
            __syncwarp();
            if (lane < qn[0]) bulk_g2s(st + (size_t)lane * row_bytes, dOut + (long long)myperm * D, row_bytes, &full_bar[ps]);
            if (++ps == nst) ps = 0;
#pragma unroll
            for (int i = 0; i + 1 < SR_AHEAD; ++i) qn[i] = qn[i + 1];
            qn[SR_AHEAD - 1] = prefetch(qslot);          // the slot just read takes the chunk SR_AHEAD further on
            qslot = (qslot + 1) & (SR_AHEAD - 1);
        }
    }
    cp_async_wait<0>();
}

}  // namespace pr
