// k_denoise.cu — spatiotemporal à-trous denoiser (a6).
//
// Replaces vpl_pbr_gi_denoise.fs:275-435 (filterInput :142-190, computeMergedReservoir :192-273,
// sampleNormalWithOffset common.glsl:216-220) as driven by the host loop
// StratusRendererBackend.cpp:1879-1952:
//   pass 0            mergeReservoirs: centre texel only -> reservoir = (rgb / max(1e-5, a), 1)
//   passes 1..I-2     5x5 à-trous, tap spacing (1 + multiplier), weight = max(0,n.n')^128 * exp(-|z-z'|)
//   pass I-1 (final)  same filter, then temporal accumulation against the reprojected history and
//                     re-modulation with the albedo plane: combined = screen + albedoPlane * illum
// Texel addressing follows SURVEY.md Appendix A1: taps land on texel centres (exact fetch,
// CLAMP_TO_EDGE); only the reprojected previous-frame reads are bilinear (8-bit filter weights),
// prevIds is NEAREST.
//
// Non-final passes write only the reservoir attachment: the other three attachments of the
// reference's FBO are dead until the final pass (combined/gi are overwritten, history is passed
// through unchanged, and `indirectIllumination` is the unchanged albedo plane of the gather).
//
// Arithmetic tiers.  The reference evaluates the 25 tap weights through the texture units, a
// normalize, pow(x,128) and exp per tap; here the decoded normal and the depth of every pixel are
// stored once per frame in a float4 "guide" plane (bit-identical to normalize(rgb*2-1)) and a weight
// is  ex2(128*lg2(max(0,n.n')) - |z-z'|*log2(e))  with FMA and the MUFU approximations.  That is
// radiance arithmetic (tolerance 1e-3 relative / 60 dB PSNR); the one decision in this shader — the
// history reset `similarity < 0.99` — stays in exact IEEE arithmetic in glm order.
//
// Work mapping of the à-trous passes.  With tap spacing s, pixel (x,y) only ever reads rows
// y + dy*s.  A thread therefore owns kLat outputs on ONE lattice column: rows (g*kLat + k)*s + r,
// k = 0..kLat-1.  Those outputs share tap rows (output k, tap dy <-> lattice row k+dy), so the thread
// loads (kLat+4)*5 taps for kLat*25 weights: 2.5x fewer L1 requests and half->float conversions than
// one-pixel-per-thread, which is what bounds this stencil on a B200 (L1/tex + issue, not HBM).  Lanes
// of a warp are consecutive in x, so every request is fully coalesced.  Each output accumulates its
// taps in the same order (dy ascending, dx ascending) wherever it sits in a band, so a row band with an
// apron reproduces the whole-frame bytes.
#include "sgi_internal.h"
#include "sgi_math.cuh"
#include "sgi_merge.cuh"
#include "sgi_denoise.cuh"

namespace sgi {
namespace {

struct Bilin { int i0, i1, j0, j1; float a, b; };
SGI_DEV Bilin bilinSetup(float u, float v, int W, int H) {
    const float fu = fsub(fmul(u, (float)W), 0.5f), fv = fsub(fmul(v, (float)H), 0.5f);
    const float i0f = floorf(fu), j0f = floorf(fv);
    Bilin r;
    r.a = fmul(floorf(fadd(fmul(fsub(fu, i0f), 256.0f), 0.5f)), 1.0f / 256.0f);
    r.b = fmul(floorf(fadd(fmul(fsub(fv, j0f), 256.0f), 0.5f)), 1.0f / 256.0f);
    r.i0 = iclamp((int)i0f, 0, W - 1); r.i1 = iclamp((int)i0f + 1, 0, W - 1);
    r.j0 = iclamp((int)j0f, 0, H - 1); r.j1 = iclamp((int)j0f + 1, 0, H - 1);
    return r;
}
SGI_DEV float bilinMix(const Bilin& k, float t00, float t10, float t01, float t11) {
    const float na = fsub(1.0f, k.a), nb = fsub(1.0f, k.b);
    float r = fmul(fmul(na, nb), t00);
    r = fadd(r, fmul(fmul(k.a, nb), t10));
    r = fadd(r, fmul(fmul(na, k.b), t01));
    r = fadd(r, fmul(fmul(k.a, k.b), t11));
    return r;
}

__global__ void __launch_bounds__(256)
prepareGuideKernel(size_t n, const uint32_t* __restrict__ normal, const float* __restrict__ depth, float4* __restrict__ guide) {
    const size_t p = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (p >= n) return;
    const f3 nn = decodeNormal(__ldg(normal + p));
    guide[p] = make_float4(nn.x, nn.y, nn.z, __ldg(depth + p));
}

// ---- temporal pre-pass (vpl_pbr_gi_denoise.fs:337-425 without the terms that depend on this frame's filtered radiance) --------
// Everything of the final pass's temporal half that does NOT need the filter's result — the reprojected reads of the previous
// frame (depth, normal, GI: LINEAR; ids: NEAREST), the history-reset decision, the new history length — is a per-pixel function of
// planes that exist when the frame starts.  It runs as its own streaming kernel on the side stream, under the gather (which is
// bound by instruction issue, not by DRAM), and leaves (prevGi.rgb, 1 / history) in a float4 plane; the final pass then only
// blends and re-modulates (finalEpilogue below: three loads, two stores).  Inside the final à-trous pass the same code cost ten
// dependent loads per output at 35 % occupancy (80 registers) — half of that kernel's 0.31 ms.
// The arithmetic is unchanged: the decision `similarity < 0.99` (:410-418) and the stored history are exact tier.
__global__ void __launch_bounds__(256)
temporalKernel(const __grid_constant__ FrameConsts fc, BandDims bd, const __grid_constant__ DenoiseIO io) {
    const int W = bd.W, rows = bd.rows;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ly = bd.c0 + blockIdx.y * 8 + (threadIdx.x >> 5);        // computed rows only
    if (x >= W || ly >= bd.c0 + bd.crows) return;
    const int y = ly + bd.yOffset;
    const size_t p = (size_t)ly * W + x;
    // the reprojected taps sit at or next to this pixel: ask L2 for those lines while the velocity (which decides the exact
    // addresses) is still on its way, so the kernel pays one DRAM latency per thread instead of two
    prefetchL2(io.prevDepth + p); prefetchL2(io.prevNormal + p); prefetchL2(io.prevGi + p); prefetchL2(io.prevIds + p);
    const float u = fdiv(fadd((float)x, 0.5f), (float)W), v = fdiv(fadd((float)y, 0.5f), (float)bd.fullH);
    const uint32_t vel = __ldg(io.velocity + p);
    float historyAccum = h2f(__ldg(io.historyIn + p));
    const float centerDepth = __ldg(io.depth + p);
    const uint32_t normalBits = __ldg(io.normal + p);
    const float currId = __ldg(io.ids + p);
    const float pu = fsub(u, h2f((unsigned short)(vel & 0xffffu))), pv = fsub(v, h2f((unsigned short)(vel >> 16)));
    float accumMultiplier = 1.0f;

    // reprojected reads; rows are band-local (previous-frame planes hold the same rows)
    Bilin k = bilinSetup(pu, pv, W, bd.fullH);
    // a band whose halo is shorter than the frame's vertical motion reads a clamped row here: counted, so the caller can widen
    // the halo (the held rows are the caller's choice, include/stratusgi.h "Row bands")
    bool miss = k.j0 < bd.yOffset || k.j0 >= bd.yOffset + rows || (k.b != 0.0f && k.j1 >= bd.yOffset + rows);
    k.j0 = iclamp(k.j0 - bd.yOffset, 0, rows - 1); k.j1 = iclamp(k.j1 - bd.yOffset, 0, rows - 1);
    const size_t q00 = (size_t)k.j0 * W + k.i0, q10 = (size_t)k.j0 * W + k.i1, q01 = (size_t)k.j1 * W + k.i0, q11 = (size_t)k.j1 * W + k.i1;
    float prevCenterDepth; f3 pn, prevGi;
    uint32_t prevNormalBits = 0;
    const bool onTexel = k.a == 0.0f && k.b == 0.0f;
    if (onTexel) {
        // the reprojected tap sits on a texel centre (always, on a static camera): the bilinear blend
        // (1*1)*t00 + 0*t10 + 0*t01 + 0*t11 is t00 exactly, so three of the four fetches are skipped
        prevCenterDepth = __ldg(io.prevDepth + q00);
        const uint32_t n00 = __ldg(io.prevNormal + q00);
        prevNormalBits = n00;
        pn = mk3(unorm8(n00 & 0xffu), unorm8((n00 >> 8) & 0xffu), unorm8((n00 >> 16) & 0xffu));
        const f4 g00 = loadHalf4(io.prevGi, q00);
        prevGi = mk3(g00.x, g00.y, g00.z);
    } else {
        prevCenterDepth = bilinMix(k, __ldg(io.prevDepth + q00), __ldg(io.prevDepth + q10), __ldg(io.prevDepth + q01), __ldg(io.prevDepth + q11));
        const uint32_t n00 = __ldg(io.prevNormal + q00), n10 = __ldg(io.prevNormal + q10), n01 = __ldg(io.prevNormal + q01), n11 = __ldg(io.prevNormal + q11);
        pn.x = bilinMix(k, unorm8(n00 & 0xffu), unorm8(n10 & 0xffu), unorm8(n01 & 0xffu), unorm8(n11 & 0xffu));
        pn.y = bilinMix(k, unorm8((n00 >> 8) & 0xffu), unorm8((n10 >> 8) & 0xffu), unorm8((n01 >> 8) & 0xffu), unorm8((n11 >> 8) & 0xffu));
        pn.z = bilinMix(k, unorm8((n00 >> 16) & 0xffu), unorm8((n10 >> 16) & 0xffu), unorm8((n01 >> 16) & 0xffu), unorm8((n11 >> 16) & 0xffu));
        const f4 g00 = loadHalf4(io.prevGi, q00), g10 = loadHalf4(io.prevGi, q10), g01 = loadHalf4(io.prevGi, q01), g11 = loadHalf4(io.prevGi, q11);
        prevGi = mk3(bilinMix(k, g00.x, g10.x, g01.x, g11.x), bilinMix(k, g00.y, g10.y, g01.y, g11.y), bilinMix(k, g00.z, g10.z, g01.z, g11.z));
    }
    const int ni = iclamp((int)floorf(fmul(pu, (float)W)), 0, W - 1);
    const int njFrame = iclamp((int)floorf(fmul(pv, (float)bd.fullH)), 0, bd.fullH - 1);
    const int nj = iclamp(njFrame - bd.yOffset, 0, rows - 1);
    miss = miss || njFrame < bd.yOffset || njFrame >= bd.yOffset + rows;
    if (miss && io.links.reprojMisses != nullptr) atomicAdd(io.links.reprojMisses, 1);
    const float prevId = __ldg(io.prevIds + (size_t)nj * W + ni);

    // the history-reset decision `similarity < 0.99` (:410-418), exact.  Shortcut: when the reprojected tap is the same texel with
    // the same normal bits, depth and id as the current pixel (every pixel of a static camera), prevCenterNormal is bit-identical
    // to centerNormal, so wn = fl(n.n) = 1 +- 3e-7, wz = exp(-0) = 1, wid = 1 and the test cannot fire.
    const bool sameTexel = onTexel && prevNormalBits == normalBits && prevCenterDepth == centerDepth && currId == prevId;
    if (!sameTexel) {
        const f3 centerNormal = decodeNormal(normalBits);        // bit-identical to the guide plane's normal
        const f3 prevCenterNormal = normalize(pn * 2.0f - 1.0f);
        const float wn = vmax(0.0f, dot(centerNormal, prevCenterNormal));
        const float wz = expf(-fabsf(fsub(centerDepth, prevCenterDepth)));
        const float wid = currId != prevId ? 0.0f : 1.0f;
        const float similarity = fmul(fmul(wn, wz), wid);
        if (similarity < 0.99f) accumMultiplier = 0.0f;
    }
    historyAccum = vmin(fadd(1.0f, fmul(historyAccum, accumMultiplier)), fc.framesPerSecond);
    const float maxAccumulationFactor = __fdividef(1.0f, historyAccum);      // radiance tier (the stored history itself is exact)
    io.historyOut[p] = f2h(historyAccum);
    io.temporal[p] = make_float4(prevGi.x, prevGi.y, prevGi.z, maxAccumulationFactor);
}

// ---- pass 0: mergeReservoirs, the centre texel only (computeMergedReservoir with minmaxNearest = 0, :259-264) ----
template <bool FINAL>
__global__ void __launch_bounds__(kDnX * kDnY)
mergeKernel(const __grid_constant__ FrameConsts fc, BandDims bd, const __grid_constant__ DenoiseIO io) {
    const int x = blockIdx.x * kDnX + threadIdx.x;
    const int ly = bd.c0 + blockIdx.y * kDnY + threadIdx.y;          // computed rows only
    if (x >= bd.W || ly >= bd.c0 + bd.crows) return;
    const int y = ly + bd.yOffset;
    const size_t p = (size_t)ly * bd.W + x;
    const float4 gc = __ldg(io.guide + p);
    const f3 centerNormal = mk3(gc.x, gc.y, gc.z);
    const float centerDepth = gc.w;
    const f3 shadowFactor = mergeCentreTexel(centerNormal, centerDepth, loadHalf4(io.shadowsIn, p), fadd((float)x, 0.5f), fadd((float)y, 0.5f),
                                             fc.time, __ldg(io.numVisible));
    const uint2 rv = packHalf4(shadowFactor.x, shadowFactor.y, shadowFactor.z, 1.0f);
    io.reservoirOut[p] = rv;
    pushRes(io.links, y, x, bd.W, rv);
    if (FINAL) finalEpilogue(bd, io, x, ly, shadowFactor);
}

// ---- passes >= 1: 5x5 à-trous (filterInput :142-190 inside the tap loop :309-326) -------------------------------

template <bool FINAL, int S>
__global__ void __launch_bounds__(kDnX * kDnY, FINAL ? 3 : 4)
atrousKernel(const __grid_constant__ FrameConsts /*fc*/, BandDims bd, const __grid_constant__ DenoiseIO io) {
    int bx = (int)blockIdx.x;
    int by = io.links.doneCounter != nullptr ? boundaryFirst((int)blockIdx.y, (int)gridDim.y) : (int)blockIdx.y;
    atrousCta<FINAL, S>(bd, io, bx, by, (int)threadIdx.x, (int)threadIdx.y);
}

}  // namespace

// CUDA loads kernels lazily on first launch, and that load can stall behind a kernel that is spinning on a band flag
// (sgi_band.cu) — so every kernel of this file is loaded when the handle is created.
bool preloadDenoiseKernels() {
    cudaFuncAttributes a; bool ok = true;
    ok = ok && cudaFuncGetAttributes(&a, prepareGuideKernel) == cudaSuccess;
    ok = ok && cudaFuncGetAttributes(&a, temporalKernel) == cudaSuccess;
    ok = ok && cudaFuncGetAttributes(&a, mergeKernel<false>) == cudaSuccess;
    ok = ok && cudaFuncGetAttributes(&a, mergeKernel<true>) == cudaSuccess;
    ok = ok && cudaFuncGetAttributes(&a, atrousKernel<false, 0>) == cudaSuccess;
    ok = ok && cudaFuncGetAttributes(&a, atrousKernel<true, 0>) == cudaSuccess;
    ok = ok && cudaFuncGetAttributes(&a, atrousKernel<false, 2>) == cudaSuccess;
    ok = ok && cudaFuncGetAttributes(&a, atrousKernel<true, 2>) == cudaSuccess;
    ok = ok && cudaFuncGetAttributes(&a, atrousKernel<false, 4>) == cudaSuccess;
    ok = ok && cudaFuncGetAttributes(&a, atrousKernel<true, 4>) == cudaSuccess;
    ok = ok && cudaFuncGetAttributes(&a, atrousKernel<false, 8>) == cudaSuccess;
    ok = ok && cudaFuncGetAttributes(&a, atrousKernel<true, 8>) == cudaSuccess;
    return ok;
}


int launchPrepareGuide(cudaStream_t stream, const BandDims& bd, const uint32_t* normal, const float* depth, float4* guide) {
    const size_t n = (size_t)bd.W * bd.rows;
    prepareGuideKernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, normal, depth, guide);
    return 1;
}

int launchTemporal(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, const DenoiseIO& io) {
    dim3 grid((bd.W + 31) / 32, (bd.crows + 7) / 8);
    temporalKernel<<<grid, 256, 0, stream>>>(fc, bd, io);
    return 1;
}

int launchDenoisePass(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, const DenoiseIO& ioIn, bool merge, bool final_, bool* signalled) {
    dim3 block(kDnX, kDnY);
    if (signalled) *signalled = false;
    DenoiseIO io = ioIn;
    if (merge) {
        io.links.doneCounter = nullptr;          // the centre-only merge pass keeps the host-launched signal
        dim3 grid((bd.W + kDnX - 1) / kDnX, (bd.crows + kDnY - 1) / kDnY);
        if (final_) mergeKernel<true><<<grid, block, 0, stream>>>(fc, bd, io);
        else mergeKernel<false><<<grid, block, 0, stream>>>(fc, bd, io);
    } else {
        // thread rows enumerate (group, residue) pairs of the tap lattice: ceil(rows / (s*kLat)) groups of s residues
        const int s = 1 + io.multiplier;
        const int threadRows = ((bd.crows + s * kLat - 1) / (s * kLat)) * s;
        dim3 grid((bd.W + kDnX - 1) / kDnX, (threadRows + kDnY - 1) / kDnY);
#ifdef __CUDACC__                        // (the per-kernel host-emulation builds of tests/cuda_on_host run atrousKernel on everything; host_tu.py rule 7)
        if (io.links.up.res == nullptr && io.links.down.res == nullptr && io.links.waitUp == nullptr && io.links.waitDown == nullptr) {
            // No reservoir rows leave this pass and it waits for nobody: whole-frame handles and row bands that recompute their halo
            // (SGI_HALO_RECOMPUTE; the final pass's GI and image rows are stored by the epilogue both kernels share).  The lattice
            // tiles inside the computed rows go through the TMA-staged body (taps from shared memory: 0.150 -> 0.122 ms per pass on
            // c4, same bytes; profiles/r02i), the CTAs of this grid that reach outside them (CLAMP_TO_EDGE taps at the plane's edges,
            // leftover rows) run first IN THE SAME LAUNCH.  The neighbours' flags are then released by the signal kernel.
            int x0, x1, y0, y1;
            if (planAtrousTma(bd, s, &x0, &x1, &y0, &y1)) {
                io.links.doneCounter = nullptr;
                borderGeometry(bd, s, (int)grid.x, (int)grid.y, x0, x1, y0, y1, &io);
                if (launchAtrousTma(stream, bd, io, final_, borderCtaCount(io))) return 1;
                io = ioIn;      // (no tensor map: everything through atrousKernel)
            }
        }
#endif
        if (io.links.doneCounter != nullptr) {   // producers of the in-kernel signal: the CTA rows whose slab holds rows pushed to a neighbour
            int producerRows = 0;
            for (int by = 0; by < (int)grid.y; ++by) {
                int lo, hi;
                atrousSlab(by, s, bd.c0, bd.c0 + bd.crows, &lo, &hi);
                if (lo < hi && slabPushes(io.links, lo, hi, bd.yOffset)) ++producerRows;
            }
            io.links.producers = producerRows * (int)grid.x;
            if (io.links.producers == 0) io.links.doneCounter = nullptr;
            else if (signalled) *signalled = true;
        }
#define SGI_ATROUS_LAUNCH(S_)                                                                             \
        do {                                                                                              \
            void (*kf)(const FrameConsts, BandDims, const DenoiseIO) = final_ ? atrousKernel<true, S_> : atrousKernel<false, S_>; \
            kf<<<grid, block, 0, stream>>>(fc, bd, io);                                   \
        } while (0)
        if (s == 2) SGI_ATROUS_LAUNCH(2); else if (s == 4) SGI_ATROUS_LAUNCH(4); else if (s == 8) SGI_ATROUS_LAUNCH(8); else SGI_ATROUS_LAUNCH(0);
#undef SGI_ATROUS_LAUNCH
    }
    return 1;
}

}  // namespace sgi
