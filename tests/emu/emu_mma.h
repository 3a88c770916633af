// Host emulation of mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 and cvt.rna.tf32.f32 with the PTX ISA fragment
// layouts (lane = 4*g + t):  A: a0=(g, t) a1=(g+8, t) a2=(g, t+4) a3=(g+8, t+4);  B: b0=(k=t, n=g) b1=(k=t+4, n=g);
// C/D: c0=(g, 2t) c1=(g, 2t+1) c2=(g+8, 2t) c3=(g+8, 2t+1).  Warp-collective: every lane must call it (a partial warp
// deadlocks on the warp barrier).  Test infrastructure only.
#pragma once
#include "emu_cuda.h"

struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

namespace pr {
inline uint32_t to_tf32(float x) {                  // round to nearest, ties away from zero, 10-bit mantissa
    uint32_t u = __float_as_uint(x);
    if ((u & 0x7f800000u) == 0x7f800000u) return u; // inf / nan
    return (u + 0x1000u) & ~0x1fffu;
}
inline void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    static uint32_t frag[32][32][6];                // [warp][lane][a0..a3, b0, b1] of the one CTA that runs at a time (cluster size 1)
    const int w = emu::warp_id(), lane = emu::lane_id();
    uint32_t* mine = frag[w][lane];
    mine[0] = a0; mine[1] = a1; mine[2] = a2; mine[3] = a3; mine[4] = b0; mine[5] = b1;
    emu::warp_sync();
    const int g = lane >> 2, t = lane & 3;
    auto Aat = [&](int r, int k) { return __uint_as_float(frag[w][(r & 7) * 4 + (k & 3)][(r >> 3) + 2 * (k >> 2)]); };
    auto Bat = [&](int k, int n) { return __uint_as_float(frag[w][n * 4 + (k & 3)][4 + (k >> 2)]); };
    float d[4];
    const int rows[4] = {g, g, g + 8, g + 8}, cols[4] = {2 * t, 2 * t + 1, 2 * t, 2 * t + 1};
    for (int e = 0; e < 4; ++e) {
        float s = c[e];
        for (int k = 0; k < 8; ++k) s += Aat(rows[e], k) * Bat(k, cols[e]);
        d[e] = s;
    }
    emu::warp_sync();
    for (int e = 0; e < 4; ++e) c[e] = d[e];
}
}  // namespace pr
