// sgi_merge.cuh — denoise pass 0 (mergeReservoirs) for one pixel, shared by the stand-alone merge kernel
// (k_denoise.cu) and the gather kernel (k_gather.cu), which applies it to the reservoir it has just produced
// so that a whole frame needs no separate pass-0 launch.
//
// computeMergedReservoir with minmaxNearest = 0 (vpl_pbr_gi_denoise.fs:192-273, loop bounds :259-264): only the
// centre texel is visited, so the merged reservoir is that texel's (rgb, a) whenever the reservoir-sampling
// draw `randUpdate < probability / cw` succeeds — and with a single candidate probability / cw is 1.
// Exact IEEE arithmetic in glm order: the draw is a decision.
#pragma once
#include "sgi_math.cuh"

namespace sgi {

SGI_DEV f3 mergeCentreTexel(f3 centerNormal, float centerDepth, f4 currReservoir, float fragX, float fragY, float time, int numVisible) {
    const float PREVENT_DIV_BY_ZERO = 0.00001f;
    float ax = 0.0f, ay = 0.0f, az = 0.0f;
    const float depthCutoff = fmul(0.1f, centerDepth);
    float runningSum = 0.0f;
    float cw = 0.0f;
    const float probabilisticWeight = fdiv(1.0f, (float)numVisible);
    const bool normalReject = dot(centerNormal, centerNormal) < 0.906f;
    const bool depthReject = fabsf(fsub(centerDepth, centerDepth)) > depthCutoff;
    if (!normalReject && !depthReject) {
        const float randUpdate = random3(fragX, fragY, time);
        const float probability = fmul(currReservoir.w, probabilisticWeight);
        cw = fadd(cw, probability);
        if (randUpdate < fdiv(probability, cw)) {
            ax = currReservoir.x; ay = currReservoir.y; az = currReservoir.z;
            runningSum = fadd(runningSum, currReservoir.w);
        }
    }
    const float denom = vmax(PREVENT_DIV_BY_ZERO, runningSum);
    return mk3(ax, ay, az) / denom;     // vec4(shadowFactor, 1.0) is what the pass stores (:329, :433)
}

}  // namespace sgi
