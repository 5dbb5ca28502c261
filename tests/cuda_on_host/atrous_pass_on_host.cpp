// atrous_pass_on_host.cpp — ONE à-trous pass, non-final or final with its blend epilogue, (csrc/k_denoise.cu launchDenoisePass) on caller-supplied guide and reservoir
// planes of a handle that holds `rows` rows and computes [c0, c0 + crows), in lockstep mode.  Built twice by
// tests/test_denoise_kernel_on_host.py: with -DSGI_HOST_TMA (next to atrous_tma_tu_on_host.cpp: the pass goes through the lattice
// tiles of csrc/k_atrous_tma.cu over the shim's tensor maps + the border CTAs) and without (the clamped kernel on everything).
// The two must write the same bytes to the same pixels and leave every other pixel alone.
#define SGI_HOST_LOCKSTEP
#include "cuda_host_shim.h"
#include "k_denoise.cu"            // tests/cuda_on_host/_build/k_denoise.cu (-I _build comes first)

extern "C" int host_atrous_pass(int W, int rows, int c0, int crows, int multiplier, const float4* guide, const uint2* shadowsIn, uint2* reservoirOut,
                                // the final pass's blend epilogue (all null: a non-final pass): (reprojected GI, 1 / history), screen, demodulation plane -> lit image, GI
                                const float4* temporal, const uint2* screen, const uint2* illum, uint2* combinedOut, uint2* giOut) {
    sgi::FrameConsts fc;
    std::memset(&fc, 0, sizeof(fc));
    sgi::BandDims bd;
    bd.W = W; bd.fullH = rows; bd.yOffset = 0; bd.rows = rows; bd.cubeRes = 0; bd.csmRes = 0; bd.c0 = c0; bd.crows = crows;
    sgi::DenoiseIO io;
    std::memset(&io, 0, sizeof(io));
    io.guide = guide; io.shadowsIn = shadowsIn; io.reservoirOut = reservoirOut; io.multiplier = multiplier;
    const bool final_ = temporal != nullptr;
    io.temporal = const_cast<float4*>(temporal); io.screen = screen; io.illum = illum; io.combinedOut = combinedOut; io.giOut = giOut;
    return sgi::launchDenoisePass(nullptr, fc, bd, io, false, final_, nullptr);
}

#ifdef SGI_HOST_TMA
extern "C" long long host_tma_boxes() { return sgihost::tmaLoads; }
#else
extern "C" long long host_tma_boxes() { return -1; }
#endif
