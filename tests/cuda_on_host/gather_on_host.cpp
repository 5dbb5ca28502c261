// gather_on_host.cpp — csrc/k_gather.cu in lockstep mode: the persistent stochastic gather (VPL table staged in dynamic shared
// memory, row segments fetched from a work counter) with the fused denoise pass 0.
#define SGI_HOST_LOCKSTEP
#include <vector>
#include "cuda_host_shim.h"
#include "k_gather.cu"             // tests/cuda_on_host/_build/k_gather.cu

namespace sgi { void fillFrameConsts(const SgiFrameParams* fp, FrameConsts* fc); }      // vpl_on_host.cpp (csrc/k_vpl.cu)

// sgi_gather (csrc/sgi_abi.cu) on a whole-frame handle, stochastic mode
extern "C" int host_gather(const SgiFrameParams* fp, int W, int H, int cubeRes, int csmRes, const float* depth, const uint32_t* normal,
                           const uint32_t* albedo, const uint16_t* refl, const uint32_t* rma, const unsigned short* ssao, const SgiVplData* updated,
                           const int32_t* visible, const SgiAtlasEntry* entries, const void* const* shadow, int numAtlases,
                           void* colorOutH4, void* reservoirOutH4, void* mergedOutH4) {
    sgi::FrameConsts fc; sgi::fillFrameConsts(fp, &fc);
    sgi::BandDims bd;
    bd.W = W; bd.fullH = H; bd.yOffset = 0; bd.rows = H; bd.cubeRes = cubeRes; bd.csmRes = csmRes; bd.c0 = 0; bd.crows = H;
    sgi::AtlasTable atl;
    for (int i = 0; i < SGI_MAX_TOTAL_SHADOW_ATLASES; ++i) { atl.shadow[i] = i < numAtlases ? (const unsigned short*)shadow[i] : nullptr; atl.diffuse[i] = nullptr; }
    std::vector<float4> guide((size_t)W * H);
    sgi::BandLinks links;
    std::memset(&links, 0, sizeof(links));
    int counters[2] = {0, 0};
    int64_t launches = 0;
    bool signalled = false;
    std::vector<float> coordU((size_t)W);
    std::vector<float2> coordRow((size_t)H);
    sgi::launchCoordTables(nullptr, bd, coordU.data(), coordRow.data());
    return sgi::launchGather(nullptr, fc, bd, depth, normal, albedo, refl, rma, ssao, updated, visible, entries, atl, (uint2*)colorOutH4,
                             (uint2*)reservoirOutH4, guide.data(), (uint2*)mergedOutH4, links, nullptr, nullptr, counters, &launches, &signalled,
                             coordU.data(), coordRow.data());
}
