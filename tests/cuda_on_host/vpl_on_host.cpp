// vpl_on_host.cpp — csrc/k_vpl.cu (visibility cull with ordered compaction, VPL colouring, fillFrameConsts) and csrc/k_prepass.cu
// are separate translation units in csrc/ with their own anonymous namespaces; here k_vpl.cu only (see prepass_on_host.cpp).
// Lockstep mode: every CUDA thread of a CTA is an OS thread, barriers / shuffles / votes are real (cuda_host_shim.h).
#define SGI_HOST_LOCKSTEP
#include <vector>
#include "cuda_host_shim.h"
#include "k_vpl.cu"                // tests/cuda_on_host/_build/k_vpl.cu

static sgi::AtlasTable atlasTable(const void* const* shadow, const void* const* diffuse, int n) {
    sgi::AtlasTable t;
    for (int i = 0; i < SGI_MAX_TOTAL_SHADOW_ATLASES; ++i) {
        t.shadow[i] = i < n ? (const unsigned short*)shadow[i] : nullptr;
        t.diffuse[i] = i < n && diffuse ? (const uchar4*)diffuse[i] : nullptr;
    }
    return t;
}
static sgi::BandDims wholeFrame(int W, int H, int cubeRes, int csmRes) {
    sgi::BandDims bd;
    bd.W = W; bd.fullH = H; bd.yOffset = 0; bd.rows = H; bd.cubeRes = cubeRes; bd.csmRes = csmRes; bd.c0 = 0; bd.crows = H;
    return bd;
}

// sgi_cull (colorToo = 0) / the fused cull + colour of sgi_frame (colorToo = 1), csrc/sgi_abi.cu cullImpl
extern "C" int host_cull(const SgiFrameParams* fp, int W, int H, int cubeRes, int csmRes, const SgiVplData* vplData, const SgiAtlasEntry* atlasByVpl,
                         const float* csm, SgiVplData* updated, int32_t* visible, SgiAtlasEntry* entries, const void* const* shadow,
                         const void* const* diffuse, int numAtlases, int colorToo, const int32_t* countDevice) {
    sgi::FrameConsts fc; sgi::fillFrameConsts(fp, &fc);
    std::vector<uint32_t> ballots(512, 0u);
    unsigned int ticket = 0;
    const sgi::AtlasTable atl = atlasTable(shadow, diffuse, numAtlases);
    return sgi::launchCull(nullptr, fc, wholeFrame(W, H, cubeRes, csmRes), vplData, atlasByVpl, csm, updated, visible, entries, ballots.data(), &ticket,
                           colorToo ? &atl : nullptr, countDevice);
}
// sgi_color
extern "C" int host_color(const SgiFrameParams* fp, int W, int H, int cubeRes, int csmRes, SgiVplData* updated, const int32_t* visible,
                          const SgiAtlasEntry* entries, const void* const* shadow, const void* const* diffuse, int numAtlases) {
    sgi::FrameConsts fc; sgi::fillFrameConsts(fp, &fc);
    return sgi::launchColor(nullptr, fc, wholeFrame(W, H, cubeRes, csmRes), updated, visible, entries, atlasTable(shadow, diffuse, numAtlases));
}
