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
// memory).  A CTA owns a compact 16x8 patch of tiles and examines the visible VPLs 128 at a time:
// a CTA-level bounding-ball test drops lights that cannot beat any tile's current 12th-best ratio,
// the survivors are compacted in list order into shared memory, and each thread then rejects most
// of those with one compare against its own 12th-best ratio before touching its list.
#include "sgi_internal.h"
#include "sgi_math.cuh"

namespace sgi {
namespace {

constexpr int kTileThreads = 128;

__global__ void __launch_bounds__(kTileThreads)
stage1Kernel(const __grid_constant__ FrameConsts fc, BandDims bd, const float* __restrict__ gDepth, const uint32_t* __restrict__ gNormal,
             SgiVplStage1PerTileOutputs* __restrict__ out) {
    // tiles of the band's computed rows only; outputs stay indexed by held tile row
    const int tilesX = bd.W / SGI_TILE_DIVISOR, tilesYCompute = bd.crows / SGI_TILE_DIVISOR;
    const int t = blockIdx.x * kTileThreads + threadIdx.x;
    if (t >= tilesX * tilesYCompute) return;
    const int tx = t % tilesX, ty = t / tilesX + bd.c0 / SGI_TILE_DIVISOR;
    const int tile = ty * tilesX + tx;
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

constexpr int kS2X = 16, kS2Y = 8;            // tiles per CTA: a compact 16x8 patch (80x40 pixels), one tile per thread
constexpr int kS2Threads = kS2X * kS2Y;
constexpr int kS2Chunk = kS2Threads;          // visible VPLs examined per round: one per thread in the CTA-level filter

SGI_DEV float warpMax(float v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
SGI_DEV float warpMin(float v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Stage 2, two levels of rejection, both conservative, so the lists are those of the plain loop:
//   CTA level   : the patch's tiles lie in a ball (c, R).  A light whose ratio lower bound (|l-c| - R) / radius exceeds
//                 the largest "12th-best ratio so far" of any tile in the patch cannot enter any of its lists.  Survivors
//                 are compacted (ballot + prefix, list order kept, so ties still go to the earlier light) into shared memory.
//   thread level: FMA estimate against the tile's own 12th-best ratio, then the exact chain in the shader's operation order.
// Margins (1e-4 relative at CTA level, 4e-5 at thread level) are >100x the rounding error of either evaluation.
__global__ void __launch_bounds__(kS2Threads)
stage2Kernel(int tilesX, int tileY0, int tilesY, const SgiVplStage1PerTileOutputs* __restrict__ stage1, const SgiVplData* __restrict__ lightData,
             const int32_t* __restrict__ visible, SgiVplStage2PerTileOutputs* __restrict__ out) {
    __shared__ StagedVpl sVpl[kS2Chunk];
    __shared__ float sRadius[kS2Chunk];
    __shared__ int sIdx[kS2Chunk];
    __shared__ int sCnt[kS2Threads / 32];
    __shared__ float sRed[kS2Threads / 32][6];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = blockIdx.x * kS2X + (tid % kS2X), ty = tileY0 + blockIdx.y * kS2Y + (tid / kS2X);   // computed tile rows [tileY0, tileY0 + tilesY)
    const bool active = tx < tilesX && ty < tileY0 + tilesY;
    const int tile = ty * tilesX + tx;
    const int numVisible = __ldg(visible);            // slot 0 = count, list starts at slot 1
    f3 fragPos = mk3(0.0f), normal = mk3(0.0f);
    if (active) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(stage1 + tile));
        const float4 b = __ldg(reinterpret_cast<const float4*>(stage1 + tile) + 1);
        fragPos = mk3(a.x, a.y, a.z); normal = mk3(b.x, b.y, b.z);
    }
    // ---- bounding ball of the patch's tile positions ----
    const float big = 3.0e38f;
    float lo[3] = {active ? fragPos.x : big, active ? fragPos.y : big, active ? fragPos.z : big};
    float hi[3] = {active ? fragPos.x : -big, active ? fragPos.y : -big, active ? fragPos.z : -big};
    #pragma unroll
    for (int k = 0; k < 3; ++k) { lo[k] = warpMin(lo[k]); hi[k] = warpMax(hi[k]); }
    if (lane == 0) { for (int k = 0; k < 3; ++k) { sRed[warp][k] = lo[k]; sRed[warp][3 + k] = hi[k]; } }
    __syncthreads();
    #pragma unroll
    for (int k = 0; k < 3; ++k) {
        lo[k] = fminf(fminf(sRed[0][k], sRed[1][k]), fminf(sRed[2][k], sRed[3][k]));
        hi[k] = fmaxf(fmaxf(sRed[0][3 + k], sRed[1][3 + k]), fmaxf(sRed[2][3 + k], sRed[3][3 + k]));
    }
    const float cx = 0.5f * (lo[0] + hi[0]), cy = 0.5f * (lo[1] + hi[1]), cz = 0.5f * (lo[2] + hi[2]);
    const float hx = hi[0] - cx, hy = hi[1] - cy, hz = hi[2] - cz;
    const float R = sqrtf(hx * hx + hy * hy + hz * hz) * 1.0001f;        // half diagonal of the box: every tile is inside
    __syncthreads();                                                      // sRed is reused below

    const float FLOAT_MAX_ = 3.402823466e+38f;
    float dist[SGI_MAX_VPLS_PER_TILE];
    int idx[SGI_MAX_VPLS_PER_TILE];
    #pragma unroll
    for (int i = 0; i < SGI_MAX_VPLS_PER_TILE; ++i) { dist[i] = FLOAT_MAX_; idx[i] = i; }
    int count = 0;
    float w = 1.0f, w2 = 1.0f;   // min(1, current 12th-best ratio) and its square: only lights with ratio below it can change the list
    float wPatch = 1.0f;         // max of w over the patch's tiles

    for (int base = 0; base < numVisible; base += kS2Chunk) {
        // ---- CTA-level filter: one candidate per thread, ordered compaction of the survivors ----
        const int j = base + tid;
        bool keep = false;
        float4 lp = make_float4(0.f, 0.f, 0.f, 0.f); float lr = 0.0f; int lightIndex = 0;
        if (j < numVisible) {
            lightIndex = __ldg(visible + 1 + j);
            lp = __ldg(reinterpret_cast<const float4*>(lightData + lightIndex));
            lr = __ldg(&lightData[lightIndex].radius);
            const float dx = lp.x - cx, dy = lp.y - cy, dz = lp.z - cz;
            const float d = sqrtf(dx * dx + dy * dy + dz * dz);
            keep = (d - R) <= wPatch * lr * 1.0001f + 1e-4f * (d + R);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) sCnt[warp] = __popc(bal);
        __syncthreads();
        int warpBase = 0, n = 0;
        #pragma unroll
        for (int k = 0; k < kS2Threads / 32; ++k) { const int c = sCnt[k]; if (k < warp) warpBase += c; n += c; }
        if (keep) {
            const int pos = warpBase + __popc(bal & ((1u << lane) - 1u));
            sVpl[pos].x = lp.x; sVpl[pos].y = lp.y; sVpl[pos].z = lp.z; sVpl[pos].r2m = lr * lr * (1.0f + 4e-5f);
            sRadius[pos] = lr;
            sIdx[pos] = lightIndex;
        }
        __syncthreads();
        // ---- thread level: this tile against the survivors, in list order ----
        if (active) {
            for (int q = 0; q < n; ++q) {
                const StagedVpl L = sVpl[q];
                const f3 lightMinusFrag = mk3(L.x, L.y, L.z) - fragPos;
                // cheap conservative pre-test: a light surely farther than min(1, current 12th-best ratio) cannot change the list
                const float d2e = __fmaf_rn(lightMinusFrag.z, lightMinusFrag.z, __fmaf_rn(lightMinusFrag.y, lightMinusFrag.y, lightMinusFrag.x * lightMinusFrag.x));
                if (d2e > L.r2m * w2) continue;
                // ... nor can one that is surely behind the tile
                const float se = __fmaf_rn(normal.z, lightMinusFrag.z, __fmaf_rn(normal.y, lightMinusFrag.y, normal.x * lightMinusFrag.x));
                if (se < 0.0f && se * se > 1e-10f * d2e) continue;
                // exact evaluation in the shader's operation order (stage2.cs:86-94)
                const float sideCheck = dot(normal, lightMinusFrag);
                if (sideCheck < 0.0f) continue;
                const float ratio = fdiv(length(lightMinusFrag), sRadius[q]);
                if (ratio > 1.0f) continue;
                if (!(ratio < dist[SGI_MAX_VPLS_PER_TILE - 1])) continue;   // sorted list: no slot can take it
                // insert before the first strictly larger entry, shifting the rest down (stage2.cs:100-122)
                float cd = ratio; int ci = sIdx[q]; bool inserted = false;
                #pragma unroll
                for (int sl = 0; sl < SGI_MAX_VPLS_PER_TILE; ++sl) {
                    const bool take = inserted || (cd < dist[sl]);
                    const float td = dist[sl]; const int ti = idx[sl];
                    dist[sl] = take ? cd : td; idx[sl] = take ? ci : ti;
                    cd = take ? td : cd; ci = take ? ti : ci;
                    inserted = take;
                }
                if (count < SGI_MAX_VPLS_PER_TILE) ++count;
                w = fminf(dist[SGI_MAX_VPLS_PER_TILE - 1], 1.0f);
                w2 = w * w;
            }
        }
        // ---- the patch's loosest threshold for the next round ----
        const float wm = warpMax(active ? w : 0.0f);
        if (lane == 0) sRed[warp][0] = wm;
        __syncthreads();
        wPatch = fmaxf(fmaxf(sRed[0][0], sRed[1][0]), fmaxf(sRed[2][0], sRed[3][0]));
    }
    if (!active) return;
    int32_t* o = reinterpret_cast<int32_t*>(out + tile);
    o[0] = count;
    #pragma unroll
    for (int i = 0; i < SGI_MAX_VPLS_PER_TILE; ++i) if (i < count) o[1 + i] = idx[i];
}

}  // namespace

// CUDA loads kernels lazily on first launch, and that load can stall behind a kernel that is spinning on a band flag
// (sgi_band.cu) — so every kernel of this file is loaded when the handle is created.
bool preloadTileKernels() {
    cudaFuncAttributes a; bool ok = true;
    ok = ok && cudaFuncGetAttributes(&a, stage1Kernel) == cudaSuccess;
    ok = ok && cudaFuncGetAttributes(&a, stage2Kernel) == cudaSuccess;
    return ok;
}

int launchStage1(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, const float* depth, const uint32_t* normal,
                 SgiVplStage1PerTileOutputs* out) {
    const int tiles = (bd.W / SGI_TILE_DIVISOR) * (bd.crows / SGI_TILE_DIVISOR);
    if (tiles <= 0) return 0;
    stage1Kernel<<<(tiles + kTileThreads - 1) / kTileThreads, kTileThreads, 0, stream>>>(fc, bd, depth, normal, out);
    return 1;
}

int launchStage2(cudaStream_t stream, const BandDims& bd, const SgiVplStage1PerTileOutputs* stage1, const SgiVplData* lightData,
                 const int32_t* visible, SgiVplStage2PerTileOutputs* out) {
    const int tilesX = bd.W / SGI_TILE_DIVISOR, tilesY = bd.crows / SGI_TILE_DIVISOR;
    if (tilesX <= 0 || tilesY <= 0) return 0;
    dim3 grid((tilesX + kS2X - 1) / kS2X, (tilesY + kS2Y - 1) / kS2Y);
    stage2Kernel<<<grid, kS2Threads, 0, stream>>>(tilesX, bd.c0 / SGI_TILE_DIVISOR, tilesY, stage1, lightData, visible, out);
    return 1;
}

}  // namespace sgi
