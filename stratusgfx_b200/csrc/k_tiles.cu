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
// memory).  A CTA owns a compact 16x8 patch of tiles, buckets the visible VPLs by a lower bound of
// their distance ratio to the patch and visits them nearest first, so the lists converge after a
// few dozen candidates and the rest of the VPLs is never looked at (see stage2Kernel).
#include "sgi_internal.h"
#include "sgi_math.cuh"

namespace sgi {
namespace {

constexpr int kS2X = 16, kS2Y = 8;            // tiles per CTA: a compact 16x8 patch (80x40 pixels), one tile per thread
constexpr int kS2Threads = kS2X * kS2Y;
constexpr int kS2Bins = 32;                   // candidate VPLs are bucketed by a lower bound of their ratio to the patch

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

// Stage 1, one thread per tile, CTAs shaped like stage 2's patches (16x8 tiles).  Besides the per-tile means it files its
// patch under one of kPatchClasses size classes (half diagonal of the patch's bounding box in world units, largest first):
// patches that straddle a depth discontinuity are the expensive ones in stage 2 (their ball is loose), and stage 2 launches
// them first so they do not form the tail of the grid.
constexpr int kPatchClasses = 8;

__global__ void __launch_bounds__(kS2Threads)
stage1Kernel(const __grid_constant__ FrameConsts fc, BandDims bd, const float* __restrict__ gDepth, const uint32_t* __restrict__ gNormal,
             SgiVplStage1PerTileOutputs* __restrict__ out, int* __restrict__ classCount, int* __restrict__ classList, int maxPatches) {
    __shared__ float sRed[kS2Threads / 32][6];
    // tiles of the band's computed rows only; outputs stay indexed by held tile row
    const int tilesX = bd.W / SGI_TILE_DIVISOR, tileY0 = bd.c0 / SGI_TILE_DIVISOR, tilesYCompute = bd.crows / SGI_TILE_DIVISOR;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = blockIdx.x * kS2X + (tid % kS2X), ty = tileY0 + blockIdx.y * kS2Y + (tid / kS2X);
    const bool active = tx < tilesX && ty < tileY0 + tilesYCompute;
    f3 sumP = mk3(0.0f), sumN = mk3(0.0f);
    if (active) {
        const int tile = ty * tilesX + tx;
        const m4 invPV = loadMat(fc.invPV);
        // viewportWidthHeight = gl_NumWorkGroups.xy * gl_WorkGroupSize.xy (stage1.cs:54)
        const float vw = (float)(tilesX * SGI_TILE_DIVISOR), vh = (float)((bd.fullH / SGI_TILE_DIVISOR) * SGI_TILE_DIVISOR);
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
    // ---- size class of the patch ----
    const float big = 3.0e38f;
    float lo[3] = {active ? sumP.x : big, active ? sumP.y : big, active ? sumP.z : big};
    float hi[3] = {active ? sumP.x : -big, active ? sumP.y : -big, active ? sumP.z : -big};
    #pragma unroll
    for (int k = 0; k < 3; ++k) { lo[k] = warpMin(lo[k]); hi[k] = warpMax(hi[k]); }
    if (lane == 0) { for (int k = 0; k < 3; ++k) { sRed[warp][k] = lo[k]; sRed[warp][3 + k] = hi[k]; } }
    __syncthreads();
    if (tid == 0) {
        float d2 = 0.0f;
        for (int k = 0; k < 3; ++k) {
            const float l = fminf(fminf(sRed[0][k], sRed[1][k]), fminf(sRed[2][k], sRed[3][k]));
            const float h = fmaxf(fmaxf(sRed[0][3 + k], sRed[1][3 + k]), fmaxf(sRed[2][3 + k], sRed[3][3 + k]));
            d2 += (h - l) * (h - l);
        }
        const float R = 0.5f * sqrtf(d2);
        int cls = kPatchClasses - 1;                                       // class 0: R >= 128, 1: >= 64, ... 7: < 2
        for (int k = 0; k < kPatchClasses - 1; ++k) if (R >= (float)(128 >> k)) { cls = k; break; }
        const int slot = atomicAdd(classCount + cls, 1);
        classList[cls * maxPatches + slot] = blockIdx.y * gridDim.x + blockIdx.x;
    }
}

// Stage 2.  The shader visits the visible VPLs in list order and keeps, per tile, the 12 smallest ratios with strict '<'
// insertion — i.e. the 12 smallest (ratio, list position) pairs in lexicographic order.  That set does not depend on the
// visiting order, so a CTA (a compact 16x8 patch of tiles inside a ball (c, R)) visits the VPLs NEAREST FIRST:
//   1. every visible VPL gets a lower bound of its ratio to any tile of the patch, lb = max(0, |l-c| - R) / radius, and is
//      bucketed by sqrt(lb) into 32 bins (shared-memory histogram + scatter of list positions); lb > 1 is dropped;
//   2. the bins are visited in order, each staged through shared memory; each thread tests the bin's VPLs against its tile (FMA pre-test against its own
//      12th-best ratio, then the exact chain in the shader's operation order, stage2.cs:86-94) and inserts with the
//      lexicographic comparison;
//   3. after a bin, if the largest 12th-best ratio of the patch is below the next bin's lower bound, nothing that
//      remains can enter any list and the CTA stops.
// Margins (1e-4 relative on the bounds, 4e-5 on the pre-test) are >100x the rounding error of either evaluation.
// `lightTable` is vplUpdatedData: entry k is a bit copy of lightData[visible[k+1]] (position and radius are not touched by
// the colouring stage), which saves the indirection through the index list.
__global__ void __launch_bounds__(kS2Threads)
stage2Kernel(int tilesX, int tileY0, int tilesY, int patchesX, const SgiVplStage1PerTileOutputs* __restrict__ stage1, const SgiVplData* __restrict__ lightTable,
             const int32_t* __restrict__ visible, SgiVplStage2PerTileOutputs* __restrict__ out, unsigned short* __restrict__ tilePos,
             const int* __restrict__ classCount, const int* __restrict__ classList, int maxPatches) {
    __shared__ unsigned short sOrder[SGI_MAX_TOTAL_VPLS_PER_FRAME];
    __shared__ int sBinCount[kS2Bins], sBinStart[kS2Bins + 1], sBinFill[kS2Bins];
    __shared__ float sRed[kS2Threads / 32][6];
    __shared__ float4 sCand[kS2Threads];
    __shared__ int sCandPos[kS2Threads];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // CTA b takes the b-th patch in size-class order (largest bounding boxes first, filed by stage 1)
    int patch = -1;
    {
        int b = blockIdx.x;
        #pragma unroll
        for (int k = 0; k < kPatchClasses; ++k) {
            const int c = __ldg(classCount + k);
            if (patch < 0 && b < c) patch = __ldg(classList + k * maxPatches + b);
            b -= c;
        }
    }
    if (patch < 0) return;                                                 // (uniform) cannot happen when stage 1 ran on the same dimensions
    const int pbx = patch % patchesX, pby = patch / patchesX;
    const int tx = pbx * kS2X + (tid % kS2X), ty = tileY0 + pby * kS2Y + (tid / kS2X);   // computed tile rows [tileY0, tileY0 + tilesY)
    const bool active = tx < tilesX && ty < tileY0 + tilesY;
    const int tile = ty * tilesX + tx;
    const int numVisible = min(__ldg(visible), SGI_MAX_TOTAL_VPLS_PER_FRAME);            // slot 0 = count, list starts at slot 1
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
    if (tid < kS2Bins) sBinCount[tid] = 0;
    __syncthreads();
    #pragma unroll
    for (int k = 0; k < 3; ++k) {
        lo[k] = fminf(fminf(sRed[0][k], sRed[1][k]), fminf(sRed[2][k], sRed[3][k]));
        hi[k] = fmaxf(fmaxf(sRed[0][3 + k], sRed[1][3 + k]), fmaxf(sRed[2][3 + k], sRed[3][3 + k]));
    }
    const float cx = 0.5f * (lo[0] + hi[0]), cy = 0.5f * (lo[1] + hi[1]), cz = 0.5f * (lo[2] + hi[2]);
    const float hx = hi[0] - cx, hy = hi[1] - cy, hz = hi[2] - cz;
    const float R = sqrtf(hx * hx + hy * hy + hz * hz) * 1.0001f;        // half diagonal of the box: every tile is inside

    // ---- 1. bucket the visible VPLs by the lower bound of their ratio (two sweeps: count, then scatter) ----
    auto binOf = [&](int j) -> int {
        const float4 lp = __ldg(reinterpret_cast<const float4*>(lightTable + j));
        const float lr = __ldg(&lightTable[j].radius);
        const float dx = lp.x - cx, dy = lp.y - cy, dz = lp.z - cz;
        const float d = sqrtf(dx * dx + dy * dy + dz * dz);
        const float gap = d - R - 1e-4f * (d + R);                        // conservative (smaller) distance to the ball
        if (gap > lr * 1.0001f) return -1;                                // ratio > 1 for every tile: the shader skips it
        const float lb = fmaxf(gap, 0.0f) / (lr * 1.0001f);               // <= the true ratio of every tile of the patch
        return min(kS2Bins - 1, (int)(sqrtf(lb) * (float)kS2Bins));
    };
    for (int j = tid; j < numVisible; j += kS2Threads) { const int b = binOf(j); if (b >= 0) atomicAdd(&sBinCount[b], 1); }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int b = 0; b < kS2Bins; ++b) { sBinStart[b] = acc; sBinFill[b] = acc; acc += sBinCount[b]; }
        sBinStart[kS2Bins] = acc;
    }
    __syncthreads();
    for (int j = tid; j < numVisible; j += kS2Threads) { const int b = binOf(j); if (b >= 0) sOrder[atomicAdd(&sBinFill[b], 1)] = (unsigned short)j; }
    __syncthreads();

    // ---- 2./3. nearest bins first ----
    const float FLOAT_MAX_ = 3.402823466e+38f;
    float dist[SGI_MAX_VPLS_PER_TILE];
    int pos[SGI_MAX_VPLS_PER_TILE];            // list positions; turned into VPL indices at the end
    #pragma unroll
    for (int i = 0; i < SGI_MAX_VPLS_PER_TILE; ++i) { dist[i] = FLOAT_MAX_; pos[i] = 0x7fffffff; }
    int count = 0;
    float w = 1.0f, w2 = 1.0f;   // min(1, current 12th-best ratio) and its square: only lights with ratio at or below it can change the list
    // a bin's candidates are staged through shared memory (one per thread), so the per-tile loop reads broadcast LDS.128
    // operands instead of chasing index -> table -> radius through global memory
    for (int b = 0; b < kS2Bins; ++b) {
      const int qEnd = sBinStart[b + 1];
      for (int q0 = sBinStart[b]; q0 < qEnd; q0 += kS2Threads) {          // (uniform) usually one round per non-empty bin
        const int n = min(kS2Threads, qEnd - q0);
        __syncthreads();                                                   // the previous round's readers are done
        if (tid < n) {
            const int j = sOrder[q0 + tid];
            const float4 lp = __ldg(reinterpret_cast<const float4*>(lightTable + j));
            sCand[tid] = make_float4(lp.x, lp.y, lp.z, __ldg(&lightTable[j].radius));
            sCandPos[tid] = j;
        }
        __syncthreads();
        if (active) {
            #pragma unroll 4
            for (int q = 0; q < n; ++q) {
                const float4 L = sCand[q];
                const float lr = L.w;
                const f3 lightMinusFrag = mk3(L.x, L.y, L.z) - fragPos;
                // cheap conservative pre-test: a light surely farther than min(1, current 12th-best ratio) cannot change the list
                const float d2e = __fmaf_rn(lightMinusFrag.z, lightMinusFrag.z, __fmaf_rn(lightMinusFrag.y, lightMinusFrag.y, lightMinusFrag.x * lightMinusFrag.x));
                if (d2e > lr * lr * (1.0f + 4e-5f) * w2) continue;
                // ... nor can one that is surely behind the tile
                const float se = __fmaf_rn(normal.z, lightMinusFrag.z, __fmaf_rn(normal.y, lightMinusFrag.y, normal.x * lightMinusFrag.x));
                if (se < 0.0f && se * se > 1e-10f * d2e) continue;
                // exact evaluation in the shader's operation order (stage2.cs:86-94)
                const float sideCheck = dot(normal, lightMinusFrag);
                if (sideCheck < 0.0f) continue;
                const float ratio = fdiv(length(lightMinusFrag), lr);
                if (ratio > 1.0f) continue;
                const int j = sCandPos[q];
                const float dl = dist[SGI_MAX_VPLS_PER_TILE - 1];
                if (!(ratio < dl || (ratio == dl && j < pos[SGI_MAX_VPLS_PER_TILE - 1]))) continue;   // worse than the 12th: no slot can take it
                // insert before the first entry that is larger in (ratio, list position) order, shifting the rest down:
                // the state the shader's in-order strict-'<' insertion (stage2.cs:100-122) reaches for the same set
                float cd = ratio; int cp = j; bool inserted = false;
                #pragma unroll
                for (int sl = 0; sl < SGI_MAX_VPLS_PER_TILE; ++sl) {
                    const bool take = inserted || cd < dist[sl] || (cd == dist[sl] && cp < pos[sl]);
                    const float td = dist[sl]; const int tp = pos[sl];
                    dist[sl] = take ? cd : td; pos[sl] = take ? cp : tp;
                    cd = take ? td : cd; cp = take ? tp : cp;
                    inserted = take;
                }
                if (count < SGI_MAX_VPLS_PER_TILE) ++count;
                w = fminf(dist[SGI_MAX_VPLS_PER_TILE - 1], 1.0f);
                w2 = w * w * (1.0f + 1e-6f);        // ties at the 12th ratio are decided by list position: keep them in the pre-test
            }
        }
      }
      if (sBinStart[b] == qEnd) continue;                                  // empty bin: nothing changed
      // the patch's loosest threshold: can anything in the remaining bins still matter?
      const float wm = warpMax(active ? w : 0.0f);
      __syncthreads();
      if (lane == 0) sRed[warp][0] = wm;
      __syncthreads();
      const float wPatch = fmaxf(fmaxf(sRed[0][0], sRed[1][0]), fmaxf(sRed[2][0], sRed[3][0]));
      const float nextLb = (float)(b + 1) / (float)kS2Bins;               // bins > b hold lights with sqrt(lb) >= (b+1)/32
      if (nextLb * nextLb > wPatch * 1.0001f) break;                       // uniform
    }
    if (!active) return;
    int32_t* o = reinterpret_cast<int32_t*>(out + tile);
    o[0] = count;
    #pragma unroll
    for (int i = 0; i < SGI_MAX_VPLS_PER_TILE; ++i) if (i < count) o[1 + i] = __ldg(visible + 1 + pos[i]);
    // the same list as positions in the compacted table, for the tile-list gather (library-owned side output)
    #pragma unroll
    for (int i = 0; i < SGI_MAX_VPLS_PER_TILE; ++i) if (i < count) tilePos[(size_t)tile * SGI_MAX_VPLS_PER_TILE + i] = (unsigned short)pos[i];
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

int patchScratchInts(const BandDims& bd) {        // ints of scratch the two launchers need: class counters + class lists
    const int tilesX = bd.W / SGI_TILE_DIVISOR, tilesY = bd.rows / SGI_TILE_DIVISOR;
    const int maxPatches = ((tilesX + kS2X - 1) / kS2X) * ((tilesY + kS2Y - 1) / kS2Y);
    return kPatchClasses * (1 + (maxPatches > 0 ? maxPatches : 1));
}

static int maxPatchesOf(const BandDims& bd) { return patchScratchInts(bd) / kPatchClasses - 1; }

int launchStage1(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, const float* depth, const uint32_t* normal,
                 SgiVplStage1PerTileOutputs* out, int* patchScratch) {
    const int tilesX = bd.W / SGI_TILE_DIVISOR, tilesY = bd.crows / SGI_TILE_DIVISOR;
    if (tilesX <= 0 || tilesY <= 0) return 0;
    cudaMemsetAsync(patchScratch, 0, kPatchClasses * sizeof(int), stream);
    dim3 grid((tilesX + kS2X - 1) / kS2X, (tilesY + kS2Y - 1) / kS2Y);
    stage1Kernel<<<grid, kS2Threads, 0, stream>>>(fc, bd, depth, normal, out, patchScratch, patchScratch + kPatchClasses, maxPatchesOf(bd));
    return 1;
}

int launchStage2(cudaStream_t stream, const BandDims& bd, const SgiVplStage1PerTileOutputs* stage1, const SgiVplData* lightTable,
                 const int32_t* visible, SgiVplStage2PerTileOutputs* out, unsigned short* tilePos, const int* patchScratch) {
    const int tilesX = bd.W / SGI_TILE_DIVISOR, tilesY = bd.crows / SGI_TILE_DIVISOR;
    if (tilesX <= 0 || tilesY <= 0) return 0;
    const int patchesX = (tilesX + kS2X - 1) / kS2X, patchesY = (tilesY + kS2Y - 1) / kS2Y;
    stage2Kernel<<<patchesX * patchesY, kS2Threads, 0, stream>>>(tilesX, bd.c0 / SGI_TILE_DIVISOR, tilesY, patchesX, stage1, lightTable, visible, out, tilePos,
                                                                 patchScratch, patchScratch + kPatchClasses, maxPatchesOf(bd));
    return 1;
}

}  // namespace sgi
