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
// stored once per frame in a float4 "guide" plane (bit-identical to normalize(rgb*2-1)), x^128 is
// seven squarings, exp is ex2.approx, and the weighted sum uses FMA.  That is radiance arithmetic
// (tolerance 1e-3 relative / 60 dB PSNR); the one decision in this shader — the history reset
// `similarity < 0.99` — stays in exact IEEE arithmetic in glm order.
#include "sgi_internal.h"
#include "sgi_math.cuh"

namespace sgi {
namespace {

constexpr float PREVENT_DIV_BY_ZERO = 0.00001f;
constexpr int kDnX = 32, kDnY = 8;

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

template <bool MERGE, bool FINAL>
__global__ void __launch_bounds__(kDnX * kDnY)
denoiseKernel(const __grid_constant__ FrameConsts fc, BandDims bd, const __grid_constant__ DenoiseIO io) {
    const int x = blockIdx.x * kDnX + threadIdx.x;
    const int ly = blockIdx.y * kDnY + threadIdx.y;
    if (x >= bd.W || ly >= bd.rows) return;
    const int W = bd.W, rows = bd.rows;
    const int y = ly + bd.yOffset;
    const size_t p = (size_t)ly * W + x;

    const float4 gc = __ldg(io.guide + p);
    const f3 centerNormal = mk3(gc.x, gc.y, gc.z);
    const float centerDepth = gc.w;

    float ax = 0.0f, ay = 0.0f, az = 0.0f, aw = 0.0f;
    if (MERGE) {
        // computeMergedReservoir with minmaxNearest = 0 (:259-264): the centre texel only
        const float depthCutoff = fmul(0.1f, centerDepth);
        float runningSum = 0.0f;
        float cw = 0.0f;
        const float probabilisticWeight = fdiv(1.0f, (float)__ldg(io.numVisible));
        const bool normalReject = dot(centerNormal, centerNormal) < 0.906f;
        const bool depthReject = fabsf(fsub(centerDepth, centerDepth)) > depthCutoff;
        if (!normalReject && !depthReject) {
            const f4 currReservoir = loadHalf4(io.shadowsIn, p);
            const float randUpdate = random3(fadd((float)x, 0.5f), fadd((float)y, 0.5f), fc.time);
            const float probability = fmul(currReservoir.w, probabilisticWeight);
            cw = fadd(cw, probability);
            if (randUpdate < fdiv(probability, cw)) {
                ax = currReservoir.x; ay = currReservoir.y; az = currReservoir.z;
                runningSum = fadd(runningSum, currReservoir.w);
            }
        }
        aw = runningSum;
    } else {
        const int step = 1 + io.multiplier;
        int tys[5];
        #pragma unroll
        for (int k = 0; k < 5; ++k) tys[k] = iclamp(ly + (k - 2) * step, 0, rows - 1) * W;   // band-local clamp == CLAMP_TO_EDGE on a whole frame
        #pragma unroll
        for (int dx = -2; dx <= 2; ++dx) {
            const int tx = iclamp(x + dx * step, 0, W - 1);
            #pragma unroll
            for (int dy = -2; dy <= 2; ++dy) {
                const size_t q = (size_t)(tys[dy + 2] + tx);
                const f4 r = loadHalf4(io.shadowsIn, q);
                float w = 1.0f;
                if (dx != 0 || dy != 0) {
                    const float4 g = __ldg(io.guide + q);
                    float wn = __fmaf_rn(gc.z, g.z, __fmaf_rn(gc.y, g.y, __fmul_rn(gc.x, g.x)));
                    wn = fmaxf(wn, 0.0f);
                    wn = wn * wn; wn = wn * wn; wn = wn * wn; wn = wn * wn; wn = wn * wn; wn = wn * wn; wn = wn * wn;   // ^128
                    w = wn * __expf(-fabsf(gc.w - g.w));
                }
                ax = __fmaf_rn(w, r.x, ax); ay = __fmaf_rn(w, r.y, ay); az = __fmaf_rn(w, r.z, az); aw = __fmaf_rn(w, r.w, aw);
            }
        }
    }
    const float denom = vmax(PREVENT_DIV_BY_ZERO, aw);
    const f3 shadowFactor = mk3(ax, ay, az) / denom;
    storeHalf4(io.reservoirOut, p, shadowFactor.x, shadowFactor.y, shadowFactor.z, 1.0f);

    if (FINAL) {
        const float u = fdiv(fadd((float)x, 0.5f), (float)W), v = fdiv(fadd((float)y, 0.5f), (float)bd.fullH);
        const f4 screenColor = loadHalf4(io.screen, p);
        const f4 gi4 = loadHalf4(io.illum, p);
        const f3 gi = mk3(gi4.x, gi4.y, gi4.z);
        const uint32_t vel = __ldg(io.velocity + p);
        const float pu = fsub(u, h2f((unsigned short)(vel & 0xffffu))), pv = fsub(v, h2f((unsigned short)(vel >> 16)));
        float historyAccum = h2f(__ldg(io.historyIn + p));
        float accumMultiplier = 1.0f;
        const f3 currGi = shadowFactor;

        // reprojected reads; rows are band-local (previous-frame planes hold the same rows)
        Bilin k = bilinSetup(pu, pv, W, bd.fullH);
        k.j0 = iclamp(k.j0 - bd.yOffset, 0, rows - 1); k.j1 = iclamp(k.j1 - bd.yOffset, 0, rows - 1);
        const size_t q00 = (size_t)k.j0 * W + k.i0, q10 = (size_t)k.j0 * W + k.i1, q01 = (size_t)k.j1 * W + k.i0, q11 = (size_t)k.j1 * W + k.i1;
        float prevCenterDepth; f3 pn, prevGi;
        if (k.a == 0.0f && k.b == 0.0f) {
            // the reprojected tap sits on a texel centre (always, on a static camera): the bilinear blend
            // (1*1)*t00 + 0*t10 + 0*t01 + 0*t11 is t00 exactly, so three of the four fetches are skipped
            prevCenterDepth = __ldg(io.prevDepth + q00);
            const uint32_t n00 = __ldg(io.prevNormal + q00);
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
        const f3 prevCenterNormal = normalize(pn * 2.0f - 1.0f);
        const float currId = __ldg(io.ids + p);
        const int ni = iclamp((int)floorf(fmul(pu, (float)W)), 0, W - 1);
        const int nj = iclamp(iclamp((int)floorf(fmul(pv, (float)bd.fullH)), 0, bd.fullH - 1) - bd.yOffset, 0, rows - 1);
        const float prevId = __ldg(io.prevIds + (size_t)nj * W + ni);

        // the history-reset decision: exact IEEE arithmetic (expf(-0) == 1 exactly on a static camera)
        const float wn = vmax(0.0f, dot(centerNormal, prevCenterNormal));
        const float wz = expf(-fabsf(fsub(centerDepth, prevCenterDepth)));
        const float wid = currId != prevId ? 0.0f : 1.0f;
        const float similarity = fmul(fmul(wn, wz), wid);
        if (similarity < 0.99f) accumMultiplier = 0.0f;
        historyAccum = vmin(fadd(1.0f, fmul(historyAccum, accumMultiplier)), fc.framesPerSecond);
        const float maxAccumulationFactor = fdiv(1.0f, historyAccum);
        const f3 illumAvg = mix(prevGi, currGi, maxAccumulationFactor);
        const f3 combined = mk3(screenColor.x, screenColor.y, screenColor.z) + gi * illumAvg;
        storeHalf4(io.combinedOut, p, combined.x, combined.y, combined.z, 1.0f);
        storeHalf4(io.giOut, p, illumAvg.x, illumAvg.y, illumAvg.z, 1.0f);
        io.historyOut[p] = f2h(historyAccum);
    }
}

}  // namespace

int launchPrepareGuide(cudaStream_t stream, const BandDims& bd, const uint32_t* normal, const float* depth, float4* guide) {
    const size_t n = (size_t)bd.W * bd.rows;
    prepareGuideKernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, normal, depth, guide);
    return 1;
}

int launchDenoisePass(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, const DenoiseIO& io, bool merge, bool final_) {
    dim3 block(kDnX, kDnY);
    dim3 grid((bd.W + kDnX - 1) / kDnX, (bd.rows + kDnY - 1) / kDnY);
    if (merge && final_) denoiseKernel<true, true><<<grid, block, 0, stream>>>(fc, bd, io);
    else if (merge) denoiseKernel<true, false><<<grid, block, 0, stream>>>(fc, bd, io);
    else if (final_) denoiseKernel<false, true><<<grid, block, 0, stream>>>(fc, bd, io);
    else denoiseKernel<false, false><<<grid, block, 0, stream>>>(fc, bd, io);
    return 1;
}

}  // namespace sgi
