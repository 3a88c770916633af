// Host emulation of the Blackwell pipeline primitives csrc/score_kernels.cuh is written against (test infrastructure only):
//   * mbarriers with arrival counts, transaction counts and phase parity (a wait that can never be satisfied aborts after
//     a timeout instead of hanging the test run);
//   * TMA box loads (zero fill out of bounds) performed ASYNCHRONOUSLY by helper threads after a random delay, incl. the
//     multicast form that writes the same shared-memory offset of every CTA in the mask and signals each one's barrier;
//   * tcgen05.mma as a DEFERRED matrix product: operands are read from shared memory only when the issuing thread commits,
//     so a stage that is overwritten too early produces wrong numbers; accumulators live in an emulated TMEM
//     [128 lanes x 512 columns]; tcgen05.ld enforces the lane-quadrant rule (warp w may touch lanes 32*(w%4)..+31);
//   * commit (plain and cluster-multicast) = run the pending MMAs, then arrive.
// Operand tiles are kept UNSWIZZLED (row r of a box at +128*r bytes) on both the TMA and the MMA side: descriptor bit
// layouts and the 128-byte swizzle are hardware facts that only a GPU run can confirm (they are shared with the v1 kernel,
// which has run); what this layer checks is the protocol and every index computation around them.
#pragma once
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <mutex>
#include <random>
#include <unordered_map>

#include "emu_cuda.h"

struct CUtensorMap {            // stand-in: [rows, cols] row-major of 4-byte (fp32) or 2-byte (fp16) elements, boxes of [box_rows x 128 bytes]
    const void* base; long long rows, cols; int box_rows; int elem_bytes;
};
#define __grid_constant__

namespace emu {
struct Mbar { int expected = 0, pending = 0; long long tx = 0; int phase = 0; };
inline std::mutex& g_m() { static std::mutex m; return m; }
inline std::condition_variable& g_cv() { static std::condition_variable c; return c; }
inline std::unordered_map<const void*, Mbar>& bars() { static std::unordered_map<const void*, Mbar> m; return m; }
inline std::vector<std::thread>& async_ops() { static std::vector<std::thread> v; return v; }
inline std::mutex& async_m() { static std::mutex m; return m; }
inline void fail(const char* what) {
    fprintf(stderr, "emu_tc: %s\n", what);
    fflush(stderr);
    abort();
}
inline void finish_phase(Mbar& b) {
    if (b.pending == 0 && b.tx == 0) {
        b.phase ^= 1;
        b.pending = b.expected;
        g_cv().notify_all();
    }
}
inline void arrive_at(const void* bar, long long expect_tx = 0) {
    std::lock_guard<std::mutex> lk(g_m());
    auto it = bars().find(bar);
    if (it == bars().end()) fail("arrive on an uninitialised mbarrier");
    Mbar& b = it->second;
    if (b.pending <= 0) fail("more arrivals than the mbarrier expects in this phase");
    b.tx += expect_tx;
    b.pending--;
    finish_phase(b);
}
inline void complete_tx_at(const void* bar, long long bytes) {
    std::lock_guard<std::mutex> lk(g_m());
    auto it = bars().find(bar);
    if (it == bars().end()) fail("complete_tx on an uninitialised mbarrier");
    it->second.tx -= bytes;
    finish_phase(it->second);
}
inline void* peer_ptr(const void* p, int rank) {
    Cta* me = t_cta;
    const size_t off = (size_t)((const unsigned char*)p - me->smem);
    if (off >= me->smem_bytes + 1024) fail("remote access to an address outside dynamic shared memory");
    return me->cluster->ctas[rank]->smem + off;
}
inline void run_async(std::function<void()> fn) {
    static thread_local std::mt19937 rng(12345u + (unsigned)(size_t)t_cta);
    const int us = (int)(rng() % 300);
    std::lock_guard<std::mutex> lk(async_m());
    async_ops().emplace_back([fn, us]() {
        std::this_thread::sleep_for(std::chrono::microseconds(us));
        fn();
    });
}
inline void join_async() {
    std::vector<std::thread> ops;
    {
        std::lock_guard<std::mutex> lk(async_m());
        ops.swap(async_ops());
    }
    for (auto& t : ops) t.join();
    std::lock_guard<std::mutex> lk(g_m());
    bars().clear();
}
struct MmaOp { uint32_t d; uint64_t adesc, bdesc; uint32_t idesc, acc; bool f16; };
inline float half_bits_to_float(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    float f;
    if (e == 0) {
        f = std::ldexp((float)m, -24);
    } else if (e == 31) {
        f = m ? std::nanf("") : INFINITY;
    } else {
        f = std::ldexp((float)(m | 0x400u), (int)e - 25);
    }
    uint32_t bits;
    memcpy(&bits, &f, 4);
    bits |= sign;
    memcpy(&f, &bits, 4);
    return f;
}
inline thread_local std::vector<MmaOp> t_pending;
inline void copy_box(unsigned char* dst, CUtensorMap tm, int c0, int c1) {
    const int eb = tm.elem_bytes, inner = 128 / eb;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(tm.base);
    for (int r = 0; r < tm.box_rows; ++r)
        for (int c = 0; c < inner; ++c) {
            const long long row = (long long)c1 + r, col = (long long)c0 + c;
            unsigned char* o = dst + (size_t)r * 128 + (size_t)c * eb;
            if (row >= 0 && row < tm.rows && col >= 0 && col < tm.cols) memcpy(o, src + ((size_t)row * tm.cols + col) * eb, eb);
            else memset(o, 0, eb);
        }
}
inline void run_pending_mmas() {
    Cta* me = t_cta;
    for (const MmaOp& op : t_pending) {
        const int M = (int)((op.idesc >> 24) & 0x1f) << 4, N = (int)((op.idesc >> 17) & 0x3f) << 3;
        if ((op.d >> 16) != 0 || M != 128) fail("emulated UMMA supports M = 128 at TMEM lane 0 only");
        const uint32_t a_addr = (uint32_t)(op.adesc & 0x3FFF) << 4, b_addr = (uint32_t)(op.bdesc & 0x3FFF) << 4;
        const unsigned char* Ab = me->smem + (a_addr - 1024);
        const unsigned char* Bb = me->smem + (b_addr - 1024);
        const int col0 = (int)(op.d & 0xffff);
        if (col0 + N > 512) fail("UMMA accumulator beyond 512 TMEM columns");
        const bool fmt16 = ((op.idesc >> 7) & 7u) == 0 && ((op.idesc >> 10) & 7u) == 0;
        if (fmt16 != op.f16) fail("instruction descriptor operand format does not match the MMA kind");
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                float s = 0.f;
                if (op.f16) {                                                   // K = 16 halves = 32 bytes; rows are 128 bytes apart
                    const uint16_t* A = reinterpret_cast<const uint16_t*>(Ab + (size_t)m * 128);
                    const uint16_t* B = reinterpret_cast<const uint16_t*>(Bb + (size_t)n * 128);
                    for (int k = 0; k < 16; ++k) s += half_bits_to_float(A[k]) * half_bits_to_float(B[k]);
                } else {                                                        // K = 8 fp32 words
                    const float* A = reinterpret_cast<const float*>(Ab + (size_t)m * 128);
                    const float* B = reinterpret_cast<const float*>(Bb + (size_t)n * 128);
                    for (int k = 0; k < 8; ++k) s += A[k] * B[k];
                }
                float& d = me->tmem[(size_t)m * 512 + col0 + n];
                d = op.acc ? d + s : s;
            }
    }
    t_pending.clear();
}
}  // namespace emu

namespace pr {
// shared-window address: offset in this CTA's dynamic shared memory, biased so that alignment arithmetic is meaningful
inline uint32_t smem_u32(const void* p) { return (uint32_t)((const unsigned char*)p - emu::t_cta->smem) + 1024u; }
inline void mbar_init(uint64_t* bar, uint32_t count) {
    std::lock_guard<std::mutex> lk(emu::g_m());
    emu::Mbar b;
    b.expected = b.pending = (int)count;
    emu::bars()[bar] = b;
}
inline void fence_mbar_init() {}
inline void mbar_arrive(uint64_t* bar) { emu::arrive_at(bar); }
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) { emu::arrive_at(bar, bytes); }
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    std::lock_guard<std::mutex> lk(emu::g_m());
    auto it = emu::bars().find(bar);
    if (it == emu::bars().end()) emu::fail("wait on an uninitialised mbarrier");
    return it->second.phase != (int)parity;
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
    std::unique_lock<std::mutex> lk(emu::g_m());
    const bool ok = emu::g_cv().wait_for(lk, std::chrono::seconds(180), [&]() {
        auto it = emu::bars().find(bar);
        if (it == emu::bars().end()) emu::fail("wait on an uninitialised mbarrier");
        return it->second.phase != (int)parity;
    });
    if (!ok) emu::fail("DEADLOCK: an mbarrier wait was not satisfied within 180 s");
}
// 1-D bulk copy global -> shared (cp.async.bulk): asynchronous, completes `bytes` on the barrier
inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    if (bytes % 16 || ((uintptr_t)dst & 15) || ((uintptr_t)src & 15)) emu::fail("bulk copy needs 16-byte aligned addresses and size");
    emu::run_async([=]() {
        memcpy(dst, src, bytes);
        emu::complete_tx_at(bar, (long long)bytes);
    });
}
// cp.async (LDGSTS) 4-byte copies: performed when the issuing thread waits for the group (the latest moment the hardware allows),
// so code that reads the destination without waiting sees the 0xAB fill of fresh shared memory
struct PendingCp { void* dst; const void* src; };
inline thread_local std::vector<std::vector<PendingCp>> t_cp_groups{std::vector<PendingCp>{}};
inline void cp_async4(void* dst, const void* src) {
    if (((uintptr_t)dst & 3) || ((uintptr_t)src & 3)) emu::fail("cp.async 4-byte copy needs 4-byte aligned addresses");
    t_cp_groups.back().push_back(PendingCp{dst, src});
}
inline void cp_async_commit() { t_cp_groups.emplace_back(); }
template <int N>
inline void cp_async_wait() {       // all but the N most recently committed groups complete
    while ((int)t_cp_groups.size() - 1 > N) {
        for (const PendingCp& c : t_cp_groups.front()) memcpy(c.dst, c.src, 4);
        t_cp_groups.erase(t_cp_groups.begin());
    }
}
inline void l2_prefetch_bulk(const void* src, uint32_t bytes) {          // a hint: only its arguments can be wrong
    if (bytes % 16 || ((uintptr_t)src & 15)) emu::fail("bulk prefetch needs a 16-byte aligned address and size");
}
inline void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    const CUtensorMap m = *tm;
    unsigned char* d = (unsigned char*)dst;
    emu::run_async([=]() {
        emu::copy_box(d, m, c0, c1);
        emu::complete_tx_at(bar, (long long)m.box_rows * 128);
    });
}
inline void tma_load_2d_mcast(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar, uint16_t mask) {
    const CUtensorMap m = *tm;
    const int n = (int)emu::t_cta->cluster->ctas.size();
    for (int r = 0; r < 16; ++r) {
        if (!((mask >> r) & 1)) continue;
        if (r >= n) emu::fail("multicast mask names a CTA outside the cluster");
        unsigned char* d = (unsigned char*)emu::peer_ptr(dst, r);
        const void* b = emu::peer_ptr(bar, r);
        emu::run_async([=]() {
            emu::copy_box(d, m, c0, c1);
            emu::complete_tx_at(b, (long long)m.box_rows * 128);
        });
    }
}
inline void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    if (ncols != 512) emu::fail("emulated TMEM allocates all 512 columns");
    if (emu::lane_id() == 0) *slot = 0;     // warp-collective on the device
    emu::warp_sync();
}
inline void tmem_dealloc(uint32_t, uint32_t) { emu::warp_sync(); }
inline void tc_fence_before() {}
inline void tc_fence_after() {}
inline void umma_tf32(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    emu::t_pending.push_back(emu::MmaOp{d, adesc, bdesc, idesc, acc, false});
}
inline void umma_f16(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    emu::t_pending.push_back(emu::MmaOp{d, adesc, bdesc, idesc, acc, true});
}
inline void umma_commit(uint64_t* bar) {
    emu::run_pending_mmas();
    emu::arrive_at(bar);
}
inline void umma_commit_mcast(uint64_t* bar, uint16_t mask) {
    emu::run_pending_mmas();
    const int n = (int)emu::t_cta->cluster->ctas.size();
    for (int r = 0; r < 16; ++r)
        if ((mask >> r) & 1) {
            if (r >= n) emu::fail("multicast mask names a CTA outside the cluster");
            emu::arrive_at(emu::peer_ptr(bar, r));
        }
}
inline void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    const int lane0 = (int)(taddr >> 16), col = (int)(taddr & 0xffff);
    if (lane0 != 32 * (emu::warp_id() & 3)) emu::fail("tcgen05.ld outside the warp's TMEM lane quadrant");
    if (col + 32 > 512) emu::fail("tcgen05.ld beyond 512 TMEM columns");
    for (int i = 0; i < 32; ++i) v[i] = emu::t_cta->tmem[(size_t)(lane0 + emu::lane_id()) * 512 + col + i];
}
inline uint32_t cluster_ctarank() { return (uint32_t)emu::t_cta->rank; }
inline void cluster_sync_all() { pthread_barrier_wait(&emu::t_cta->cluster->bar); }
}  // namespace pr
