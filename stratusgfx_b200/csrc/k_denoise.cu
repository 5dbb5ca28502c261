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

template <bool MERGE, bool FINAL>
__global__ void __launch_bounds__(kDnX * kDnY)
denoiseKernel(const __grid_constant__ FrameConsts fc, BandDims bd, const __grid_constant__ DenoiseIO io) {
    const int x = blockIdx.x * kDnX + threadIdx.x;
    const int ly = blockIdx.y * kDnY + threadIdx.y;
    if (x >= bd.W || ly >= bd.rows) return;
    const int W = bd.W, rows = bd.rows;
    const int y = ly + bd.yOffset;
    const size_t p = (size_t)ly * W + x;

    const f3 centerNormal = decodeNormal(__ldg(io.normal + p));
    const float centerDepth = __ldg(io.depth + p);

    f4 reservoirFiltered = mk4(0.0f, 0.0f, 0.0f, 0.0f);
    if (MERGE) {
        // computeMergedReservoir with minmaxNearest = 0 (:259-264): the centre texel only
        const float depthCutoff = fmul(0.1f, centerDepth);
        float runningSum = 0.0f;
        f4 centerReservoir = mk4(0.0f, 0.0f, 0.0f, 0.0f);
        const float probabilisticWeight = fdiv(1.0f, (float)__ldg(io.numVisible));
        const bool normalReject = dot(centerNormal, centerNormal) < 0.906f;
        const bool depthReject = fabsf(fsub(centerDepth, centerDepth)) > depthCutoff;
        if (!normalReject && !depthReject) {
            const f4 currReservoir = loadHalf4(io.shadowsIn, p);
            const float randUpdate = random3(fadd((float)x, 0.5f), fadd((float)y, 0.5f), fc.time);
            const float probability = fmul(currReservoir.w, probabilisticWeight);
            centerReservoir.w = fadd(centerReservoir.w, probability);
            if (randUpdate < fdiv(probability, centerReservoir.w)) {
                centerReservoir.x = fadd(centerReservoir.x, currReservoir.x);
                centerReservoir.y = fadd(centerReservoir.y, currReservoir.y);
                centerReservoir.z = fadd(centerReservoir.z, currReservoir.z);
                runningSum = fadd(runningSum, currReservoir.w);
            }
        }
        centerReservoir.w = runningSum;
        reservoirFiltered = centerReservoir;
    } else {
        const int step = 1 + io.multiplier;
        #pragma unroll
        for (int dx = -2; dx <= 2; ++dx) {
            const int tx = iclamp(x + dx * step, 0, W - 1);
            #pragma unroll
            for (int dy = -2; dy <= 2; ++dy) {
                const int ty = iclamp(ly + dy * step, 0, rows - 1);   // band-local clamp == CLAMP_TO_EDGE on a whole frame
                const size_t q = (size_t)ty * W + tx;
                const f4 reservoir = loadHalf4(io.shadowsIn, q);
                float filtered;
                if (dx == 0 && dy == 0) filtered = 1.0f;
                else {
                    const float currDepth = __ldg(io.depth + q);
                    const float wz = fdiv(expf(-fabsf(fsub(centerDepth, currDepth))), 1.0f);
                    const f3 currNormal = decodeNormal(__ldg(io.normal + q));
                    float wn = vmax(0.0f, dot(centerNormal, currNormal));
                    wn = powf(wn, 128.0f);
                    filtered = fmul(fmul(wn, wz), 1.0f);
                }
                reservoirFiltered = reservoirFiltered + filtered * reservoir;
            }
        }
    }
    const float denom = vmax(PREVENT_DIV_BY_ZERO, reservoirFiltered.w);
    const f3 shadowFactor = mk3(reservoirFiltered.x, reservoirFiltered.y, reservoirFiltered.z) / denom;
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
        const float prevCenterDepth = bilinMix(k, __ldg(io.prevDepth + q00), __ldg(io.prevDepth + q10), __ldg(io.prevDepth + q01), __ldg(io.prevDepth + q11));
        const uint32_t n00 = __ldg(io.prevNormal + q00), n10 = __ldg(io.prevNormal + q10), n01 = __ldg(io.prevNormal + q01), n11 = __ldg(io.prevNormal + q11);
        f3 pn;
        pn.x = bilinMix(k, unorm8(n00 & 0xffu), unorm8(n10 & 0xffu), unorm8(n01 & 0xffu), unorm8(n11 & 0xffu));
        pn.y = bilinMix(k, unorm8((n00 >> 8) & 0xffu), unorm8((n10 >> 8) & 0xffu), unorm8((n01 >> 8) & 0xffu), unorm8((n11 >> 8) & 0xffu));
        pn.z = bilinMix(k, unorm8((n00 >> 16) & 0xffu), unorm8((n10 >> 16) & 0xffu), unorm8((n01 >> 16) & 0xffu), unorm8((n11 >> 16) & 0xffu));
        const f3 prevCenterNormal = normalize(pn * 2.0f - 1.0f);
        const f4 g00 = loadHalf4(io.prevGi, q00), g10 = loadHalf4(io.prevGi, q10), g01 = loadHalf4(io.prevGi, q01), g11 = loadHalf4(io.prevGi, q11);
        const f3 prevGi = mk3(bilinMix(k, g00.x, g10.x, g01.x, g11.x), bilinMix(k, g00.y, g10.y, g01.y, g11.y), bilinMix(k, g00.z, g10.z, g01.z, g11.z));
        const float currId = __ldg(io.ids + p);
        const int ni = iclamp((int)floorf(fmul(pu, (float)W)), 0, W - 1);
        const int nj = iclamp(iclamp((int)floorf(fmul(pv, (float)bd.fullH)), 0, bd.fullH - 1) - bd.yOffset, 0, rows - 1);
        const float prevId = __ldg(io.prevIds + (size_t)nj * W + ni);

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
