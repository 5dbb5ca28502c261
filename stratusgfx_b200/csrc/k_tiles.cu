// k_tiles.cu — per-tile VPL list construction (a3 + a4).
//
// a3 replaces vpl_tiled_deferred_culling_stage1.cs:50-90: mean world position and mean decoded
// normal of each 5x5-pixel tile, summed in gl_LocalInvocationIndex order (row-major in the tile).
// a4 replaces vpl_tiled_deferred_culling_stage2.cs:61-129: per tile, the 12 nearest VPLs by
// length(l-p)/radius that face the tile normal, in visible-list order with strict '<' insertion
// (ties keep the earlier light).  Both are dormant in the reference backend
// (StratusRendererBackend.cpp:1733-1783) but are what north_star calls "per-tile VPL lists".
//
// Stage 2 keeps the 12-entry sorted list in registers (fully unrolled carry insertion, no local
// memory), stages the visible VPLs' (position, radius, index) through shared memory in chunks,
// and rejects most candidates with one compare against the current 12th-best ratio before
// touching the list.
#include "sgi_internal.h"
#include "sgi_math.cuh"

namespace sgi {
namespace {

constexpr int kTileThreads = 128;
constexpr int kStage2Chunk = 1024;   // VPLs staged per shared-memory round: 1024 * 20 B = 20 KB

__global__ void __launch_bounds__(kTileThreads)
stage1Kernel(const __grid_constant__ FrameConsts fc, BandDims bd, const float* __restrict__ gDepth, const uint32_t* __restrict__ gNormal,
             SgiVplStage1PerTileOutputs* __restrict__ out) {
    const int tilesX = bd.W / SGI_TILE_DIVISOR, tilesYLocal = bd.rows / SGI_TILE_DIVISOR;
    const int tile = blockIdx.x * kTileThreads + threadIdx.x;
    if (tile >= tilesX * tilesYLocal) return;
    const int tx = tile % tilesX, ty = tile / tilesX;
    const m4 invPV = loadMat(fc.invPV);
    // viewportWidthHeight = gl_NumWorkGroups.xy * gl_WorkGroupSize.xy (stage1.cs:54)
    const float vw = (float)(tilesX * SGI_TILE_DIVISOR), vh = (float)((bd.fullH / SGI_TILE_DIVISOR) * SGI_TILE_DIVISOR);
    f3 sumP = mk3(0.0f), sumN = mk3(0.0f);
    #pragma unroll 5
    for (int i = 0; i < 25; ++i) {
        const int px = tx * 5 + (i % 5), ly = ty * 5 + (i / 5);
        const int py = ly + bd.yOffset;
        const float u = fdiv(fadd((float)px, 0.5f), vw), v = fdiv(fadd((float)py, 0.5f), vh);
        const size_t p = (size_t)ly * bd.W + px;
        const f3 fragPos = worldPositionFromDepth(u, v, __ldg(gDepth + p), invPV);
        const f3 normal = decodeNormal(__ldg(gNormal + p));
        sumP = sumP + fragPos;
        sumN = sumN + normal;
    }
    sumP = sumP / 25.0f;
    sumN = sumN / 25.0f;
    sumN = normalize(sumN);
    float4* o = reinterpret_cast<float4*>(out + tile);
    o[0] = make_float4(sumP.x, sumP.y, sumP.z, 1.0f);
    o[1] = make_float4(sumN.x, sumN.y, sumN.z, 1.0f);
}

struct StagedVpl { float x, y, z, r2m; };   // position and radius^2 * (1 + margin), the pre-test threshold at ratio 1

__global__ void __launch_bounds__(kTileThreads)
stage2Kernel(int numTiles, const SgiVplStage1PerTileOutputs* __restrict__ stage1, const SgiVplData* __restrict__ lightData,
             const int32_t* __restrict__ visible, SgiVplStage2PerTileOutputs* __restrict__ out) {
    __shared__ StagedVpl sVpl[kStage2Chunk];
    __shared__ float sRadius[kStage2Chunk];
    __shared__ int sIdx[kStage2Chunk];
    const int tile = blockIdx.x * kTileThreads + threadIdx.x;
    const bool active = tile < numTiles;
    const int numVisible = __ldg(visible);            // slot 0 = count, list starts at slot 1
    f3 fragPos = mk3(0.0f), normal = mk3(0.0f);
    if (active) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(stage1 + tile));
        const float4 b = __ldg(reinterpret_cast<const float4*>(stage1 + tile) + 1);
        fragPos = mk3(a.x, a.y, a.z); normal = mk3(b.x, b.y, b.z);
    }
    const float FLOAT_MAX_ = 3.402823466e+38f;
    float dist[SGI_MAX_VPLS_PER_TILE];
    int idx[SGI_MAX_VPLS_PER_TILE];
    #pragma unroll
    for (int i = 0; i < SGI_MAX_VPLS_PER_TILE; ++i) { dist[i] = FLOAT_MAX_; idx[i] = i; }
    int count = 0;
    float w2 = 1.0f;          // min(1, current 12th-best ratio)^2: only lights with ratio below it can change the list

    for (int base = 0; base < numVisible; base += kStage2Chunk) {
        const int n = min(kStage2Chunk, numVisible - base);
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += kTileThreads) {
            const int lightIndex = __ldg(visible + 1 + base + j);
            const float4 p = __ldg(reinterpret_cast<const float4*>(lightData + lightIndex));
            const float r = __ldg(&lightData[lightIndex].radius);
            sVpl[j].x = p.x; sVpl[j].y = p.y; sVpl[j].z = p.z; sVpl[j].r2m = r * r * (1.0f + 4e-5f);
            sRadius[j] = r;
            sIdx[j] = lightIndex;
        }
        __syncthreads();
        if (!active) continue;
        #pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const StagedVpl L = sVpl[j];
            const f3 lightMinusFrag = mk3(L.x, L.y, L.z) - fragPos;
            // Cheap conservative pre-test (FMA estimate, 4e-5 relative margin = >100x the rounding error of either
            // evaluation): a light that is surely farther than min(1, current 12th-best ratio) cannot change the list.
            const float d2e = __fmaf_rn(lightMinusFrag.z, lightMinusFrag.z, __fmaf_rn(lightMinusFrag.y, lightMinusFrag.y, lightMinusFrag.x * lightMinusFrag.x));
            if (d2e > L.r2m * w2) continue;
            // ... nor can one that is surely behind the tile
            const float se = __fmaf_rn(normal.z, lightMinusFrag.z, __fmaf_rn(normal.y, lightMinusFrag.y, normal.x * lightMinusFrag.x));
            if (se < 0.0f && se * se > 1e-10f * d2e) continue;
            // exact evaluation in the shader's operation order (stage2.cs:86-94)
            const float sideCheck = dot(normal, lightMinusFrag);
            if (sideCheck < 0.0f) continue;
            const float ratio = fdiv(length(lightMinusFrag), sRadius[j]);
            if (ratio > 1.0f) continue;
            if (!(ratio < dist[SGI_MAX_VPLS_PER_TILE - 1])) continue;   // sorted list: no slot can take it
            // insert before the first strictly larger entry, shifting the rest down (stage2.cs:100-122)
            float cd = ratio; int ci = sIdx[j]; bool inserted = false;
            #pragma unroll
            for (int s = 0; s < SGI_MAX_VPLS_PER_TILE; ++s) {
                const bool take = inserted || (cd < dist[s]);
                const float td = dist[s]; const int ti = idx[s];
                dist[s] = take ? cd : td; idx[s] = take ? ci : ti;
                cd = take ? td : cd; ci = take ? ti : ci;
                inserted = take;
            }
            if (count < SGI_MAX_VPLS_PER_TILE) ++count;
            const float w = fminf(dist[SGI_MAX_VPLS_PER_TILE - 1], 1.0f);
            w2 = w * w;
        }
    }
    if (!active) return;
    int32_t* o = reinterpret_cast<int32_t*>(out + tile);
    o[0] = count;
    #pragma unroll
    for (int i = 0; i < SGI_MAX_VPLS_PER_TILE; ++i) if (i < count) o[1 + i] = idx[i];
}

}  // namespace

int launchStage1(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, const float* depth, const uint32_t* normal,
                 SgiVplStage1PerTileOutputs* out) {
    const int tiles = (bd.W / SGI_TILE_DIVISOR) * (bd.rows / SGI_TILE_DIVISOR);
    if (tiles <= 0) return 0;
    stage1Kernel<<<(tiles + kTileThreads - 1) / kTileThreads, kTileThreads, 0, stream>>>(fc, bd, depth, normal, out);
    return 1;
}

int launchStage2(cudaStream_t stream, const BandDims& bd, const SgiVplStage1PerTileOutputs* stage1, const SgiVplData* lightData,
                 const int32_t* visible, SgiVplStage2PerTileOutputs* out) {
    const int tiles = (bd.W / SGI_TILE_DIVISOR) * (bd.rows / SGI_TILE_DIVISOR);
    if (tiles <= 0) return 0;
    stage2Kernel<<<(tiles + kTileThreads - 1) / kTileThreads, kTileThreads, 0, stream>>>(tiles, stage1, lightData, visible, out);
    return 1;
}

}  // namespace sgi
