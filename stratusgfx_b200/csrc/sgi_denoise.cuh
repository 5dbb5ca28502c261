// sgi_denoise.cuh — pieces of the à-trous filter shared by its two kernels (k_denoise.cu: clamped global loads; k_atrous_tma.cu:
// taps staged through shared memory by TMA): the weight, the tap accumulation and the final pass's blend epilogue.  One source for
// the arithmetic is what keeps the two kernels' outputs the same bytes.
#pragma once
#include "sgi_internal.h"
#include "sgi_math.cuh"

namespace sgi {

constexpr float kDenoisePreventDivByZero = 0.00001f;
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kLat = 4;                       // à-trous: outputs per thread, consecutive on the tap lattice of its column

SGI_DEV float lg2Fast(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
SGI_DEV float ex2Fast(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

__host__ __device__ inline bool slabPushes(const BandLinks& L, int lo, int hi, int yOffset) {      // band-local rows [lo, hi)
    const int a = lo + yOffset, b = hi + yOffset;
    return (pushTarget(L.up) && a < L.up.lastRow && b > L.up.firstRow) || (pushTarget(L.down) && a < L.down.lastRow && b > L.down.firstRow);
}
// CTA rows in launch order when the kernel signals its neighbours itself: the first, the last and the second-last row of the grid
// (the ones that hold the band's boundary rows) before the interior
__device__ __forceinline__ int boundaryFirst(int p, int n) {
    if (n <= 3) return p;
    return p == 0 ? 0 : (p == 1 ? n - 1 : (p == 2 ? n - 2 : p - 2));
}

// The tap spacing of the reference's passes (2, 4, 8: multiplier 1, 3, 7, BE:1920) is a template parameter; S = 0 reads it from the
// pass parameters (other iteration counts).  A thread whose (kLat+4) x 5 taps all lie inside the held rows and columns — every
// thread but those within two taps of an edge — addresses them as one row pointer per plane, advanced by a constant stride, plus
// IMMEDIATE column offsets; the others take the clamped path below (CLAMP_TO_EDGE).  Both paths read the same taps in the same
// order with the same arithmetic, so their outputs are the same bytes.  (Round 2 A/B on c4, profiles/r02a: 0.166 -> 0.148 ms per
// non-final pass against per-tap clamps and 64-bit index arithmetic on every thread.)

// one loaded tap (guide gq, reservoir rv) at tap line c (0 .. kLat+3) and column d (0 .. 4) into the kLat outputs it serves
SGI_DEV void atrousTap(const float4 (&cen)[kLat], float (&acc)[kLat][4], int c, int d, float4 gq, uint2 rv) {
    // the input of every pass >= 1 is the previous pass's vec4(shadowFactor, 1.0): alpha is 1 exactly
    const float rr = h2f((unsigned short)(rv.x & 0xffffu)), rg = h2f((unsigned short)(rv.x >> 16)), rb = h2f((unsigned short)(rv.y & 0xffffu));
    const float tq = gq.w * kLog2e;
    #pragma unroll
    for (int k = 0; k < kLat; ++k) {
        const int dy = c - 2 - k;
        if (dy < -2 || dy > 2) continue;
        float w = 1.0f;                              // centre tap (vpl_pbr_gi_denoise.fs:156)
        if (dy != 0 || d != 2) {
            // max(0, n.n')^128 * exp(-|z - z'|)   (sigmaN = 128, sigmaZ = 1, :93-98, :169-173)
            const float wn = fmaxf(__fmaf_rn(cen[k].z, gq.z, __fmaf_rn(cen[k].y, gq.y, __fmul_rn(cen[k].x, gq.x))), 0.0f);
            w = ex2Fast(__fmaf_rn(128.0f, lg2Fast(wn), -fabsf(cen[k].w - tq)));
        }
        acc[k][0] = __fmaf_rn(w, rr, acc[k][0]); acc[k][1] = __fmaf_rn(w, rg, acc[k][1]);
        acc[k][2] = __fmaf_rn(w, rb, acc[k][2]); acc[k][3] += w;
    }
}

// filtered output k of a thread: shadowFactor = sum(w * rgb) / max(1e-5, sum(w))   (:329), radiance tier
SGI_DEV f3 atrousResolve(const float (&a)[4]) {
    const float invDenom = __fdividef(1.0f, vmax(kDenoisePreventDivByZero, a[3]));
    return mk3(a[0] * invDenom, a[1] * invDenom, a[2] * invDenom);
}

// ---- final-pass epilogue: blend with the reprojected history and re-modulate (vpl_pbr_gi_denoise.fs:424-434) --------------
SGI_DEV void finalEpilogue(const BandDims& bd, const DenoiseIO& io, int x, int ly, f3 shadowFactor) {
    const int W = bd.W;
    const int y = ly + bd.yOffset;
    const size_t p = (size_t)ly * W + x;
    const float4 t = __ldg(io.temporal + p);                 // (prevGi.rgb, 1 / history) from temporalKernel
    const f4 screenColor = loadHalf4(io.screen, p);
    const f4 gi4 = loadHalf4(io.illum, p);
    const f3 gi = mk3(gi4.x, gi4.y, gi4.z);
    const f3 illumAvg = mix(mk3(t.x, t.y, t.z), shadowFactor, t.w);
    const f3 combined = mk3(screenColor.x, screenColor.y, screenColor.z) + gi * illumAvg;
    const uint2 cv = packHalf4(combined.x, combined.y, combined.z, 1.0f), gv = packHalf4(illumAvg.x, illumAvg.y, illumAvg.z, 1.0f);
    io.combinedOut[p] = cv;
    io.giOut[p] = gv;
    // row bands: this frame's GI boundary rows go to the neighbours' halos (reprojection may land there next frame; the history
    // plane is only ever read at the pixel itself); the owner band also stores its own lit rows into its frame image
    pushGi(io.links, y, x, W, gv);
    if (io.links.image != nullptr) io.links.image[(size_t)y * W + x] = cv;
}


}  // namespace sgi
