// tiles_on_host.cpp — csrc/k_tiles.cu in lockstep mode: stage 1 (per-tile averages, patch size classes, the packed VPL table) and
// the persistent stage 2 (patch counter, bucketed candidates in shared memory, 4 lanes per tile with shuffled thresholds).
#define SGI_HOST_LOCKSTEP
#include <vector>
#include "cuda_host_shim.h"
#include "k_tiles.cu"              // tests/cuda_on_host/_build/k_tiles.cu

namespace sgi { void fillFrameConsts(const SgiFrameParams* fp, FrameConsts* fc); }      // vpl_on_host.cpp (csrc/k_vpl.cu)

// sgi_tile_lists (csrc/sgi_abi.cu tileListsOn)
extern "C" int host_tile_lists(const SgiFrameParams* fp, int W, int H, const float* depth, const uint32_t* normal, const SgiVplData* updated,
                               const int32_t* visible, SgiVplStage1PerTileOutputs* stage1, SgiVplStage2PerTileOutputs* stage2) {
    sgi::FrameConsts fc; sgi::fillFrameConsts(fp, &fc);
    sgi::BandDims bd;
    bd.W = W; bd.fullH = H; bd.yOffset = 0; bd.rows = H; bd.cubeRes = 0; bd.csmRes = 0; bd.c0 = 0; bd.crows = H;
    std::vector<int> patchScratch((size_t)sgi::patchScratchInts(bd), 0);
    std::vector<float4> posRadius(SGI_MAX_TOTAL_VPLS_PER_FRAME);
    const size_t tiles = (size_t)(W / SGI_TILE_DIVISOR) * (H / SGI_TILE_DIVISOR);
    std::vector<unsigned short> tilePos((tiles ? tiles : 1) * SGI_MAX_VPLS_PER_TILE);
    std::vector<int32_t> visibleCopy(SGI_MAX_TOTAL_VPLS_PER_FRAME + 1);
    int n = sgi::launchStage1(nullptr, fc, bd, depth, normal, stage1, patchScratch.data(), updated, visible, posRadius.data(), visibleCopy.data());
    n += sgi::launchStage2(nullptr, bd, stage1, posRadius.data(), visibleCopy.data(), stage2, tilePos.data(), patchScratch.data());
    return n;
}
