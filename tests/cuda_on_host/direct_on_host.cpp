// direct_on_host.cpp — csrc/k_direct.cu executed on the host: includes the translated copy host_tu.py wrote into _build/.
#include "cuda_host_shim.h"
#include "k_direct.cu"             // tests/cuda_on_host/_build/k_direct.cu (-I _build comes first)

extern "C" int host_direct_lighting(const SgiFrameParams* fp, int W, int H, int cubeRes, int csmRes, const float* depth, const uint32_t* normal,
                                    const uint32_t* albedo, const uint16_t* refl, const uint32_t* rma, const unsigned short* ssao, const float* csm,
                                    const unsigned short* const* shadowAtlases, int numAtlases, const SgiDirectParams* dp, void* outHalf4) {
    sgi::FrameConsts fc;
    std::memset(&fc, 0, sizeof(fc));
    // the part of sgi::fillFrameConsts (csrc/k_vpl.cu) this kernel reads
    for (int i = 0; i < 16; ++i) fc.invPV[i] = fp->invProjectionView[i];
    for (int i = 0; i < 3; ++i) { fc.viewPos[i] = fp->viewPosition[i]; fc.fogColor[i] = fp->fogColor[i]; fc.sunDir[i] = fp->infiniteLightDirection[i]; fc.sunColor[i] = fp->infiniteLightColor[i]; }
    fc.minRoughness = fp->minRoughness; fc.usePerceptualRoughness = fp->usePerceptualRoughness; fc.fogDensity = fp->fogDensity;
    for (int c = 0; c < 4; ++c) for (int i = 0; i < 16; ++i) fc.cascadeProjViews[c][i] = fp->cascadeProjViews[c][i];
    for (int c = 0; c < 3; ++c) for (int i = 0; i < 4; ++i) fc.cascadePlanes[c][i] = fp->cascadePlanes[c][i];
    sgi::BandDims bd;
    bd.W = W; bd.fullH = H; bd.yOffset = 0; bd.rows = H; bd.cubeRes = cubeRes; bd.csmRes = csmRes; bd.c0 = 0; bd.crows = H;
    sgi::AtlasTable atl;
    for (int i = 0; i < SGI_MAX_TOTAL_SHADOW_ATLASES; ++i) { atl.shadow[i] = i < numAtlases ? shadowAtlases[i] : nullptr; atl.diffuse[i] = nullptr; }
    return sgi::launchDirectLighting(nullptr, fc, bd, depth, normal, albedo, refl, rma, ssao, csm, atl, *dp, (uint2*)outHalf4);
}
