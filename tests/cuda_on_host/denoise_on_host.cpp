// denoise_on_host.cpp — csrc/k_denoise.cu executed on the host (translated copy in _build/, see host_tu.py): the pass loop of
// sgi_denoise_pass (csrc/sgi_abi.cu) on a whole-frame handle — guide plane, pass 0 (centre merge), à-trous passes, final pass.
#include <vector>
#include "cuda_host_shim.h"
#include "k_denoise.cu"            // tests/cuda_on_host/_build/k_denoise.cu (-I _build comes first)

extern "C" int host_denoise(const SgiFrameParams* fp, int W, int H, int iters, const uint32_t* normal, const float* depth, const void* screenH4,
                            const uint32_t* velocityH2, const float* ids, const uint32_t* prevNormal, const float* prevIds, const float* prevDepth,
                            const void* illumH4, const void* reservoirH4, const void* prevGiH4, const unsigned short* historyIn,
                            const int32_t* visible, void* passReservoirsH4, void* combinedOutH4, void* giOutH4, unsigned short* historyOut) {
    sgi::FrameConsts fc;
    std::memset(&fc, 0, sizeof(fc));
    fc.time = fp->time;                                        // the fields of sgi::fillFrameConsts (csrc/k_vpl.cu) this file reads
    fc.framesPerSecond = (float)(1.0 / fp->deltaSeconds);
    sgi::BandDims bd;
    bd.W = W; bd.fullH = H; bd.yOffset = 0; bd.rows = H; bd.cubeRes = 0; bd.csmRes = 0; bd.c0 = 0; bd.crows = H;
    const size_t P = (size_t)W * H;
    std::vector<float4> guide(P);
    std::vector<uint2> res[2] = {std::vector<uint2>(P), std::vector<uint2>(P)};
    std::vector<float4> temporal(P);
    sgi::launchPrepareGuide(nullptr, bd, normal, depth, guide.data());
    for (int pass = 0; pass < iters; ++pass) {
        const bool merge = pass < 1, final_ = pass + 1 == iters;
        sgi::DenoiseIO io;
        std::memset(&io, 0, sizeof(io));
        io.screen = (const uint2*)screenH4; io.velocity = velocityH2; io.normal = normal; io.ids = ids; io.depth = depth;
        io.prevNormal = prevNormal; io.prevIds = prevIds; io.prevDepth = prevDepth;
        io.illum = (const uint2*)illumH4;
        io.shadowsIn = pass == 0 ? (const uint2*)reservoirH4 : res[(pass - 1) % 2].data();
        io.reservoirOut = res[pass % 2].data();
        io.prevGi = (const uint2*)prevGiH4; io.historyIn = historyIn; io.numVisible = visible;
        io.combinedOut = (uint2*)combinedOutH4; io.giOut = (uint2*)giOutH4; io.historyOut = historyOut;
        io.multiplier = (1 << pass) - 1;
        io.guide = guide.data();
        io.temporal = temporal.data();
        if (final_) sgi::launchTemporal(nullptr, fc, bd, io);      // the temporal pre-pass (sgi_frame runs it on the side stream at the start of the frame)
        sgi::launchDenoisePass(nullptr, fc, bd, io, merge, final_, nullptr);
        std::memcpy((uint2*)passReservoirsH4 + (size_t)pass * P, res[pass % 2].data(), P * sizeof(uint2));
    }
    return 0;
}

// The border CTAs of a pass whose interior goes through the TMA tiles (csrc/sgi_denoise.cuh: borderGeometry, borderCtaCount,
// borderCta): the enumeration the launch uses, and — per CTA of the clamped grid — whether atrousCta's own test sends it home.
extern "C" int host_border_enumeration(int W, int rows, int c0, int crows, int s, int x0, int x1, int y0, int y1, int* listBxBy, int cap,
                                       int* gridXY, unsigned char* insideFlags /* [gridY][gridX] */) {
    sgi::BandDims bd;
    bd.W = W; bd.fullH = rows; bd.yOffset = 0; bd.rows = rows; bd.cubeRes = 0; bd.csmRes = 0; bd.c0 = c0; bd.crows = crows;
    const int threadRows = ((crows + s * sgi::kLat - 1) / (s * sgi::kLat)) * s;
    const int gx = (W + sgi::kDnX - 1) / sgi::kDnX, gy = (threadRows + sgi::kDnY - 1) / sgi::kDnY;
    gridXY[0] = gx; gridXY[1] = gy;
    sgi::DenoiseIO io;
    std::memset(&io, 0, sizeof(io));
    sgi::borderGeometry(bd, s, gx, gy, x0, x1, y0, y1, &io);
    for (int by = 0; by < gy; ++by) {
        int lo, hi;
        sgi::atrousSlab(by, s, c0, c0 + crows, &lo, &hi);
        for (int bx = 0; bx < gx; ++bx)      // the early exit of atrousCta
            insideFlags[by * gx + bx] = (bx * sgi::kDnX >= io.skipX0 && bx * sgi::kDnX + sgi::kDnX <= io.skipX1 && lo >= io.skipY0 && hi <= io.skipY1) ? 1 : 0;
    }
    const int n = sgi::borderCtaCount(io);
    for (int i = 0; i < n && i < cap; ++i) sgi::borderCta(io, i, listBxBy + 2 * i, listBxBy + 2 * i + 1);
    return n;
}
