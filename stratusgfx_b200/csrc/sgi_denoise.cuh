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


// ---- passes >= 1: 5x5 à-trous (filterInput :142-190 inside the tap loop :309-326), clamped global-memory taps ------------------
constexpr int kDnX = 32, kDnY = 8;            // threads of a CTA: 32 columns x 8 lattice thread rows
SGI_DEV void prefetchL2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Rows a CTA row `by` of the à-trous grid can output (band-local, a superset for spacings > kDnY): thread rows kDnY*by .. +kDnY-1
// are lattice groups g = t / s, each kLat*s rows tall.  Host and device use the same formula to decide which CTAs push rows to
// a neighbour (the producers of the in-kernel signal).
__host__ __device__ inline void atrousSlab(int by, int s, int c0, int cEnd, int* lo, int* hi) {
    const int gLo = (by * kDnY) / s, gHi = (by * kDnY + kDnY - 1) / s;
    *lo = c0 + gLo * kLat * s;
    const int h = c0 + (gHi + 1) * kLat * s;
    *hi = h < cEnd ? h : cEnd;
}
template <int S>
SGI_DEV void atrousTapsInterior(const DenoiseIO& io, int W, int firstRow, int x, const float4 (&cen)[kLat], float (&acc)[kLat][4]) {
    const float4* __restrict__ rowG = io.guide + ((size_t)firstRow * W + x);
    const uint2* __restrict__ rowS = io.shadowsIn + ((size_t)firstRow * W + x);
    const size_t stride = (size_t)S * W;
    #pragma unroll
    for (int c = 0; c < kLat + 4; ++c) {
        #pragma unroll
        for (int d = 0; d < 5; ++d) atrousTap(cen, acc, c, d, __ldg(rowG + (d - 2) * S), __ldg(rowS + (d - 2) * S));
        rowG += stride; rowS += stride;
    }
}
// The border CTAs of a pass whose interior goes through the TMA tiles, enumerated without a table: the CTA rows above the covered
// block, the two flanks of the covered rows, the CTA rows below.  Host (count) and device (index -> CTA) use the same order.
__host__ __device__ inline int borderCtaCount(const DenoiseIO& io) {
    const int flank = io.borderBx0 + (io.borderGridX - io.borderBx1);
    return io.borderBy0 * io.borderGridX + (io.borderBy1 - io.borderBy0) * flank + (io.borderGridY - io.borderBy1) * io.borderGridX;
}
__host__ __device__ inline void borderCta(const DenoiseIO& io, int i, int* bx, int* by) {
    const int gx = io.borderGridX, top = io.borderBy0 * gx;
    if (i < top) { *by = i / gx; *bx = i - *by * gx; return; }
    i -= top;
    const int flank = io.borderBx0 + (gx - io.borderBx1), mid = (io.borderBy1 - io.borderBy0) * flank;
    if (i < mid) {
        const int r = i / flank, j = i - r * flank;
        *by = io.borderBy0 + r; *bx = j < io.borderBx0 ? j : io.borderBx1 + (j - io.borderBx0);
        return;
    }
    i -= mid;
    const int r = i / gx;
    *by = io.borderBy1 + r; *bx = i - r * gx;
}

// The CTAs of the clamped grid (gridX x gridY, spacing s, computed rows [c0, c0 + crows)) that lie entirely inside the pixel
// rectangle [x0, x1) x [y0, y1): x tiles [bx0, bx1) of the CTA rows [by0, by1) (contiguous: a slab's rows grow with by).
__host__ inline void borderGeometry(const BandDims& bd, int s, int gridX, int gridY, int x0, int x1, int y0, int y1, DenoiseIO* io) {
    io->skipX0 = x0; io->skipX1 = x1; io->skipY0 = y0; io->skipY1 = y1;
    io->borderGridX = gridX; io->borderGridY = gridY;
    io->borderBx0 = (x0 + kDnX - 1) / kDnX; io->borderBx1 = x1 / kDnX;
    if (io->borderBx1 < io->borderBx0) io->borderBx1 = io->borderBx0;
    io->borderBy0 = 0; io->borderBy1 = 0;
    bool open_ = false;
    for (int by = 0; by < gridY; ++by) {
        int lo, hi;
        atrousSlab(by, s, bd.c0, bd.c0 + bd.crows, &lo, &hi);
        const bool inside = lo < hi && lo >= y0 && hi <= y1;
        if (inside && !open_) { io->borderBy0 = by; io->borderBy1 = by + 1; open_ = true; }
        else if (inside && io->borderBy1 == by) io->borderBy1 = by + 1;
    }
}

// One CTA (bx, by) of the clamped à-trous grid, thread (tx, ty) of kDnX x kDnY.  It is the body of atrousKernel (k_denoise.cu: row
// bands that exchange halos, planes the TMA kernel does not take) and of the border CTAs inside atrousTmaKernel's launch
// (k_atrous_tma.cu: the part of a plane its tiles do not cover).
template <bool FINAL, int S>
SGI_DEV void atrousCta(const BandDims& bd, const DenoiseIO& io, int bx, int by, int tx, int ty) {
    const int W = bd.W, rows = bd.rows;
    const int s = S ? S : 1 + io.multiplier;                 // tap spacing: ivec2(dx,dy) + ivec2(dx,dy)*multiplier (:318)
    const int x = bx * kDnX + tx;
    const int t = by * kDnY + ty;                            // lattice column-of-rows id: residue r, group g
    const int g = t / s, r = t - g * s;
    const int m0 = g * kLat;                                 // lattice index of output 0: computed row c0 + (m0 + k)*s + r
    const int c0 = bd.c0, cEnd = bd.c0 + bd.crows;           // only the band's own rows are computed; taps may read its halo rows
    if (io.skipX1 > io.skipX0) {
        // whole-frame handles: the interior was written by the TMA-staged kernel (k_atrous_tma.cu); a CTA whose outputs all lie
        // inside that rectangle leaves at once (uniform), the others skip the covered outputs one by one
        int lo, hi;
        atrousSlab(by, s, bd.c0, bd.c0 + bd.crows, &lo, &hi);
        if (bx * kDnX >= io.skipX0 && bx * kDnX + kDnX <= io.skipX1 && lo >= io.skipY0 && hi <= io.skipY1) return;
    }
    bool live = x < W && c0 + m0 * s + r < cEnd;
    // (border launch) a thread whose four outputs all lie inside the TMA kernel's rectangle has nothing to do
    if (io.skipX1 > io.skipX0 && x >= io.skipX0 && x < io.skipX1 && c0 + m0 * s + r >= io.skipY0 && c0 + (m0 + kLat - 1) * s + r < io.skipY1) live = false;
    // row bands: taps of lattice rows m0-2 .. m0+kLat+1 reach into the halo only at the band's edges; those CTAs wait for the
    // neighbour's previous pass (whose kernel stored the halo rows), every other CTA starts at once
    bandWaitCta(io.links, live && m0 < 2, live && c0 + (m0 + kLat + 1) * s + r >= cEnd);
    if (live) {

    if (FINAL) {
        // the epilogue reads three more planes per output, first touched in this kernel: ask L2 for them now
        #pragma unroll
        for (int k = 0; k < kLat; ++k) {
            const size_t p = (size_t)min(c0 + (m0 + k) * s + r, cEnd - 1) * W + x;
            prefetchL2(io.temporal + p); prefetchL2(io.screen + p); prefetchL2(io.illum + p);
        }
    }
    float4 cen[kLat];                                        // centre guides (normal, depth*log2e)
    float acc[kLat][4];
    #pragma unroll
    for (int k = 0; k < kLat; ++k) {
        const int ly = min(c0 + (m0 + k) * s + r, rows - 1);
        cen[k] = __ldg(io.guide + (size_t)ly * W + x);
        cen[k].w *= kLog2e;
        acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.0f;
    }
    const int firstRow = c0 + (m0 - 2) * s + r;              // row of tap line c = 0; the last one is firstRow + (kLat + 3) * s
    const bool interior = S != 0 && x >= 2 * s && x + 2 * s < W && firstRow >= 0 && firstRow + (kLat + 3) * s < rows;
    if (interior) atrousTapsInterior<(S ? S : 1)>(io, W, firstRow, x, cen, acc);
    else {
        int txs[5];
        #pragma unroll
        for (int d = 0; d < 5; ++d) txs[d] = iclamp(x + (d - 2) * s, 0, W - 1);
        #pragma unroll
        for (int c = 0; c < kLat + 4; ++c) {
            // band-local clamp == CLAMP_TO_EDGE on a whole frame
            const size_t rowBase = (size_t)iclamp(c0 + (m0 + c - 2) * s + r, 0, rows - 1) * W;
            const float4* __restrict__ rowG = io.guide + rowBase;
            const uint2* __restrict__ rowS = io.shadowsIn + rowBase;
            #pragma unroll
            for (int d = 0; d < 5; ++d) atrousTap(cen, acc, c, d, __ldg(rowG + txs[d]), __ldg(rowS + txs[d]));
        }
    }
    #pragma unroll
    for (int k = 0; k < kLat; ++k) {
        const int ly = c0 + (m0 + k) * s + r;
        if (ly >= cEnd) break;
        if (x >= io.skipX0 && x < io.skipX1 && ly >= io.skipY0 && ly < io.skipY1) continue;
        const f3 shadowFactor = atrousResolve(acc[k]);
        const uint2 rv = packHalf4(shadowFactor.x, shadowFactor.y, shadowFactor.z, 1.0f);
        io.reservoirOut[(size_t)ly * W + x] = rv;
        pushRes(io.links, ly + bd.yOffset, x, W, rv);
        if (FINAL) finalEpilogue(bd, io, x, ly, shadowFactor);
    }
    }   // live
    if (io.links.doneCounter != nullptr) {                   // (uniform per CTA) a producer CTA counts itself once all its threads have stored
        int lo, hi;
        atrousSlab(by, s, c0, cEnd, &lo, &hi);
        if (lo < hi && slabPushes(io.links, lo, hi, bd.yOffset)) {
            __syncthreads();
            if (tx == 0 && ty == 0) producerDone(io.links);
        }
    }
}

}  // namespace sgi
