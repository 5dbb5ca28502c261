// k_tiles.cu — per-tile VPL list construction (a3 + a4).
//
// a3 replaces vpl_tiled_deferred_culling_stage1.cs:50-90: mean world position and mean decoded
// normal of each 5x5-pixel tile, summed in gl_LocalInvocationIndex order (row-major in the tile).
// a4 replaces vpl_tiled_deferred_culling_stage2.cs:61-129: per tile, the 12 nearest VPLs by
// length(l-p)/radius that face the tile normal, in visible-list order with strict '<' insertion
// (ties keep the earlier light).  Both are dormant in the reference backend
// (StratusRendererBackend.cpp:1733-1783) but are what north_star calls "per-tile VPL lists".
//
// Stage 2 keeps 12-entry sorted lists in registers (fully unrolled carry insertion, no local
// memory).  A CTA owns a compact 8x4 patch of tiles, buckets the visible VPLs by a lower bound of
// their distance ratio to the patch and visits them nearest first, four lanes per tile, so the lists
// converge after a few dozen candidates and the rest of the VPLs is never looked at (see stage2Kernel).
#include "sgi_internal.h"
#include "sgi_math.cuh"

namespace sgi {
namespace {

constexpr int kS2X = 8, kS2Y = 8;             // stage 2: tiles per CTA, a compact 8x8 patch (40x40 pixels) ...
constexpr int kS2Lanes = 4;                   // ... with 4 lanes per tile: lane g of a tile scans every 4th candidate VPL
constexpr int kS2Threads = kS2X * kS2Y * kS2Lanes;
constexpr int kS2Warps = kS2Threads / 32;
constexpr int kS2Bins = 32;                   // candidate VPLs are bucketed by a lower bound of their ratio to the patch
constexpr int kS2Round = 128;                 // candidates staged through shared memory per round
constexpr int kS2Exchange = 4;                // candidates a lane scans between two threshold exchanges with its tile's other lanes
constexpr int kS1X = 16, kS1Y = 8;            // stage 1: tiles per CTA (80x40 pixels), one tile per thread
constexpr int kS1Threads = kS1X * kS1Y;
constexpr int kSubX = kS1X / kS2X, kSubY = kS1Y / kS2Y;      // stage-2 patches per stage-1 CTA
static_assert(kS2X == 8 && kS2X * kS2Lanes == 32 && kS1X == 2 * kS2X && kS1Y % kS2Y == 0 && (kS1Threads / 32) % kSubY == 0 && kS2Round <= kS2Threads,
              "a stage-2 warp is one patch row; bit 3 of a stage-1 lane selects the patch column");

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
// reductions over the lanes of a warp that share bit 3 of the lane id (one half of each of the two tile rows a stage-1 warp holds)
SGI_DEV float halfRowMax(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16)); v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2)); v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return v;
}
SGI_DEV float halfRowMin(float v) {
    v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 16)); v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 4));
    v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 2)); v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return v;
}
SGI_DEV float sqrtApprox(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// Stage 1, one thread per tile, CTAs of 16x8 tiles = kSubX x kSubY of stage 2's patches.  Besides the per-tile means it files
// each of its patches under one of kPatchClasses size classes (half diagonal of the patch's bounding box in world units,
// largest first): patches that straddle a depth discontinuity are the expensive ones in stage 2 (their ball is loose), and
// stage 2 launches them first so they do not form the tail of the grid.  CTA (0,0) also packs the (position, radius) of the
// visible VPLs into a 16-byte-stride table for stage 2 (the 64-byte VplData records cost four times the sectors).
constexpr int kPatchClasses = 8;

__global__ void __launch_bounds__(kS1Threads)
stage1Kernel(const __grid_constant__ FrameConsts fc, BandDims bd, const float* __restrict__ gDepth, const uint32_t* __restrict__ gNormal,
             SgiVplStage1PerTileOutputs* __restrict__ out, int* __restrict__ classCount, int* __restrict__ classList, int maxPatches,
             const SgiVplData* __restrict__ lightTable, const int32_t* __restrict__ visible, float4* __restrict__ posRadius,
             int32_t* __restrict__ visibleCopy) {
    __shared__ float sRed[kS1Threads / 32][2][6];
    // tiles of the band's computed rows only; outputs stay indexed by held tile row
    const int tilesX = bd.W / SGI_TILE_DIVISOR, tileY0 = bd.c0 / SGI_TILE_DIVISOR, tilesYCompute = bd.crows / SGI_TILE_DIVISOR;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (blockIdx.x == 0 && blockIdx.y == 0) {
        const int V = min(__ldg(visible), SGI_MAX_TOTAL_VPLS_PER_FRAME);
        for (int j = tid; j < V; j += kS1Threads) {
            const float4 lp = __ldg(reinterpret_cast<const float4*>(lightTable + j));
            posRadius[j] = make_float4(lp.x, lp.y, lp.z, __ldg(&lightTable[j].radius));
        }
        // ... and copies the visible list: after this kernel the tile stages read library memory only, so the caller's tables
        // are free for the next frame's cull while stage 2 is still running (sgi_early_begin)
        for (int j = tid; j <= V; j += kS1Threads) visibleCopy[j] = j == 0 ? V : __ldg(visible + j);
    }
    const int tx = blockIdx.x * kS1X + (tid % kS1X), ty = tileY0 + blockIdx.y * kS1Y + (tid / kS1X);
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
    // ---- size class of each of the CTA's patches: sub-patch (sx, sy) = (bit 3 of the lane, tile row / kS2Y) ----
    const float big = 3.0e38f;
    float lo[3] = {active ? sumP.x : big, active ? sumP.y : big, active ? sumP.z : big};
    float hi[3] = {active ? sumP.x : -big, active ? sumP.y : -big, active ? sumP.z : -big};
    #pragma unroll
    for (int k = 0; k < 3; ++k) { lo[k] = halfRowMin(lo[k]); hi[k] = halfRowMax(hi[k]); }
    if ((lane & 23) == 0) { for (int k = 0; k < 3; ++k) { sRed[warp][lane >> 3][k] = lo[k]; sRed[warp][lane >> 3][3 + k] = hi[k]; } }
    __syncthreads();
    if (tid < kSubX * kSubY) {
        const int sx = tid % kSubX, sy = tid / kSubX;
        constexpr int warpsPerSub = (kS1Threads / 32) / kSubY;
        const int patchesX = (tilesX + kS2X - 1) / kS2X, patchesY = (tilesYCompute + kS2Y - 1) / kS2Y;
        const int px = blockIdx.x * kSubX + sx, py = blockIdx.y * kSubY + sy;
        if (px < patchesX && py < patchesY) {
            float d2 = 0.0f;
            for (int k = 0; k < 3; ++k) {
                float l = big, h = -big;
                for (int wv = 0; wv < warpsPerSub; ++wv) { l = fminf(l, sRed[sy * warpsPerSub + wv][sx][k]); h = fmaxf(h, sRed[sy * warpsPerSub + wv][sx][3 + k]); }
                d2 += (h - l) * (h - l);
            }
            const float R = 0.5f * sqrtf(d2);
            int cls = kPatchClasses - 1;                                       // class 0: R >= 64, 1: >= 32, ... 7: < 1
            for (int k = 0; k < kPatchClasses - 1; ++k) if (R >= (float)(64 >> k)) { cls = k; break; }
            const int slot = atomicAdd(classCount + cls, 1);
            classList[cls * maxPatches + slot] = py * patchesX + px;
        }
    }
}

// Stage 2.  The shader visits the visible VPLs in list order and keeps, per tile, the 12 smallest ratios with strict '<'
// insertion — i.e. the 12 smallest (ratio, list position) pairs in lexicographic order.  That set does not depend on the
// visiting order or on how the candidates are split, so:
//   1. a CTA owns a compact 8x8 patch of tiles inside a ball (c, R).  Every visible VPL gets a lower bound of its ratio to any
//      tile of the patch, lb = max(0, |l-c| - R) / radius, and is bucketed by sqrt(lb) into 32 bins (shared-memory histogram,
//      prefix sum, scatter of list positions); lb > 1 is dropped.  The result is the candidate list in NEAREST-FIRST order;
//   2. the list is walked in rounds of 128 candidates staged through shared memory.  A tile is served by 4 lanes; lane g tests
//      every 4th candidate of the round against the tile (FMA pre-test against the tile's current threshold, then the exact
//      chain in the shader's operation order, stage2.cs:86-94) and inserts into ITS OWN 12-entry list with the lexicographic
//      comparison.  Every 4 candidates the 4 lanes exchange a threshold that bounds the tile's final 12th-best ratio from
//      above: the smallest 12th entry of a lane, the largest 6th entry of a lane pair, the largest 3rd entry of the four
//      lanes — each proves that 12 candidates at or below it exist;
//   3. after a round, if the largest threshold of the patch is below the lower bound of the next unvisited candidate's bin,
//      nothing that remains can enter any list and the CTA is done with the patch (most patches need one round);
//   4. the four sorted lists of a tile are merged (12 rounds of a 4-lane lexicographic minimum), which is the shader's list.
// The grid is persistent: CTAs fetch patches from a global counter in size-class order (stage 1 files the loose patches
// first), so a small grid (a row band) has no second, almost empty wave and the few loose patches overlap with the many
// tight ones.  The serial chain of a tile is a quarter of a one-thread-per-tile scan.  The bounds of step 1 use approximate
// sqrt/rcp: they only have to be conservative, and their margins (1e-4 relative; 4e-5 on the pre-test) are >100x the error
// of either evaluation.  `posRadius` is stage 1's 16-byte packing of vplUpdatedData (entry k = (position, radius) of
// lightData[visible[k+1]]; neither is touched by the colouring stage).
__global__ void __launch_bounds__(kS2Threads)
stage2Kernel(int tilesX, int tileY0, int tilesY, int patchesX, int numPatches, const SgiVplStage1PerTileOutputs* __restrict__ stage1,
             const float4* __restrict__ posRadius, const int32_t* __restrict__ visible, SgiVplStage2PerTileOutputs* __restrict__ out,
             unsigned short* __restrict__ tilePos, const int* __restrict__ classCount, const int* __restrict__ classList, int maxPatches, int* __restrict__ nextPatch) {
    __shared__ unsigned short sOrder[SGI_MAX_TOTAL_VPLS_PER_FRAME];
    __shared__ unsigned char sBinOf[SGI_MAX_TOTAL_VPLS_PER_FRAME];
    __shared__ int sBinCount[kS2Bins], sBinFill[kS2Bins], sTotal, sNext;
    __shared__ float sRed[kS2Warps][6];
    __shared__ float4 sCand[kS2Round];
    __shared__ int sCandPos[kS2Round];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = tid & (kS2Lanes - 1), tl = tid / kS2Lanes;                 // lane of the tile, tile of the patch
    const int numVisible = min(__ldg(visible), SGI_MAX_TOTAL_VPLS_PER_FRAME);            // slot 0 = count, list starts at slot 1
    int classEnd[kPatchClasses];                                             // running totals of the size classes
    {
        int acc = 0;
        #pragma unroll
        for (int k = 0; k < kPatchClasses; ++k) { acc += __ldg(classCount + k); classEnd[k] = acc; }
    }
    const float FLOAT_MAX_ = 3.402823466e+38f;
    const float big = 3.0e38f;

    while (true) {
        // the next patch in size-class order (largest bounding boxes first, filed by stage 1)
        if (tid == 0) sNext = atomicAdd(nextPatch, 1);
        if (tid < kS2Bins) sBinCount[tid] = 0;
        __syncthreads();
        const int b = sNext;
        if (b >= numPatches || b >= classEnd[kPatchClasses - 1]) break;    // (uniform)
        int patch = 0;
        {
            int prev = 0; bool got = false;
            #pragma unroll
            for (int k = 0; k < kPatchClasses; ++k) {
                if (!got && b < classEnd[k]) { patch = __ldg(classList + k * maxPatches + (b - prev)); got = true; }
                prev = classEnd[k];
            }
        }
        const int pbx = patch % patchesX, pby = patch / patchesX;
        const int tx = pbx * kS2X + (tl % kS2X), ty = tileY0 + pby * kS2Y + (tl / kS2X);   // computed tile rows [tileY0, tileY0 + tilesY)
        const bool active = tx < tilesX && ty < tileY0 + tilesY;
        const int tile = ty * tilesX + tx;
        f3 fragPos = mk3(0.0f), normal = mk3(0.0f);
        if (active) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(stage1 + tile));
            const float4 c = __ldg(reinterpret_cast<const float4*>(stage1 + tile) + 1);
            fragPos = mk3(a.x, a.y, a.z); normal = mk3(c.x, c.y, c.z);
        }
        // ---- bounding ball of the patch's tile positions ----
        float lo[3] = {active ? fragPos.x : big, active ? fragPos.y : big, active ? fragPos.z : big};
        float hi[3] = {active ? fragPos.x : -big, active ? fragPos.y : -big, active ? fragPos.z : -big};
        #pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] = warpMin(lo[k]); hi[k] = warpMax(hi[k]); }
        if (lane == 0) { for (int k = 0; k < 3; ++k) { sRed[warp][k] = lo[k]; sRed[warp][3 + k] = hi[k]; } }
        __syncthreads();
        #pragma unroll
        for (int k = 0; k < 3; ++k) {
            float l = sRed[0][k], h = sRed[0][3 + k];
            #pragma unroll
            for (int wv = 1; wv < kS2Warps; ++wv) { l = fminf(l, sRed[wv][k]); h = fmaxf(h, sRed[wv][3 + k]); }
            lo[k] = l; hi[k] = h;
        }
        const float cx = 0.5f * (lo[0] + hi[0]), cy = 0.5f * (lo[1] + hi[1]), cz = 0.5f * (lo[2] + hi[2]);
        const float hx = hi[0] - cx, hy = hi[1] - cy, hz = hi[2] - cz;
        const float R = sqrtf(hx * hx + hy * hy + hz * hz) * 1.0001f;    // half diagonal of the box: every tile is inside

        // ---- 1. bucket the visible VPLs by the lower bound of their ratio (count, prefix, scatter) ----
        #pragma unroll 4
        for (int j = tid; j < numVisible; j += kS2Threads) {
            const float4 lp = __ldg(posRadius + j);
            const float dx = lp.x - cx, dy = lp.y - cy, dz = lp.z - cz;
            const float d = sqrtApprox(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx)));
            const float gap = d - R - 1e-4f * (d + R);                    // conservative (smaller) distance to the ball
            const float lr = lp.w * 1.0001f;
            int bin = 0xff;                                                // ratio > 1 for every tile: the shader skips it
            if (!(gap > lr)) {
                const float lb = __fdividef(fmaxf(gap, 0.0f), lr) * 0.9999f;   // <= the true ratio of every tile of the patch
                bin = min(kS2Bins - 1, (int)(sqrtApprox(lb) * (float)kS2Bins));
                atomicAdd(&sBinCount[bin], 1);
            }
            sBinOf[j] = (unsigned char)bin;
        }
        __syncthreads();
        if (tid < 32) {                                                    // exclusive prefix sum of the 32 bin counts
            const int c = sBinCount[tid];
            int inc = c;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (tid >= o) inc += t; }
            sBinFill[tid] = inc - c;
            if (tid == 31) sTotal = inc;
        }
        __syncthreads();
        for (int j = tid; j < numVisible; j += kS2Threads) { const int bin = sBinOf[j]; if (bin != 0xff) sOrder[atomicAdd(&sBinFill[bin], 1)] = (unsigned short)j; }
        const int M = sTotal;                                              // candidates, nearest bins first
        // (the first round's barrier orders the scatter before the reads of sOrder)

        // ---- 2./3. rounds of kS2Round candidates, nearest first ----
        float dist[SGI_MAX_VPLS_PER_TILE];
        int pos[SGI_MAX_VPLS_PER_TILE];        // list positions; turned into VPL indices at the end
        #pragma unroll
        for (int i = 0; i < SGI_MAX_VPLS_PER_TILE; ++i) { dist[i] = FLOAT_MAX_; pos[i] = 0x7fffffff; }
        int count = 0;
        float w = 1.0f, w2 = 1.0f;   // min(1, bound of the tile's 12th-best ratio) and its square: only lights with ratio at or below it can change the list
        for (int q0 = 0; q0 < M; q0 += kS2Round) {                        // (uniform) usually one round
            const int n = min(kS2Round, M - q0);
            __syncthreads();                                               // the previous round's readers are done
            if (tid < n) {
                const int j = sOrder[q0 + tid];
                sCand[tid] = __ldg(posRadius + j);
                sCandPos[tid] = j;
            }
            __syncthreads();
            for (int q1 = 0; q1 < n; q1 += kS2Lanes * kS2Exchange) {      // (uniform)
                #pragma unroll
                for (int e = 0; e < kS2Exchange; ++e) {
                    const int q = q1 + e * kS2Lanes + g;
                    if (q >= n || !active) continue;
                    const float4 L = sCand[q];
                    const float lr = L.w;
                    const f3 lightMinusFrag = mk3(L.x, L.y, L.z) - fragPos;
                    // cheap conservative pre-test: a light surely farther than the threshold cannot be in the tile's list
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
                    if (!(ratio < dl || (ratio == dl && j < pos[SGI_MAX_VPLS_PER_TILE - 1]))) continue;   // worse than this lane's 12th: no slot can take it
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
                }
                // threshold exchange between the 4 lanes of the tile (every lane of the warp takes part)
                const float a12 = dist[SGI_MAX_VPLS_PER_TILE - 1], a6 = dist[5], a3 = dist[2];
                float m12 = fminf(a12, __shfl_xor_sync(0xffffffffu, a12, 1)); m12 = fminf(m12, __shfl_xor_sync(0xffffffffu, m12, 2));
                float m6 = fmaxf(a6, __shfl_xor_sync(0xffffffffu, a6, 1)); m6 = fminf(m6, __shfl_xor_sync(0xffffffffu, m6, 2));
                float m3 = fmaxf(a3, __shfl_xor_sync(0xffffffffu, a3, 1)); m3 = fmaxf(m3, __shfl_xor_sync(0xffffffffu, m3, 2));
                w = fminf(fminf(m12, 1.0f), fminf(m6, m3));
                w2 = w * w * (1.0f + 1e-6f);    // ties at the 12th ratio are decided by list position: keep them in the pre-test
            }
            if (q0 + kS2Round >= M) break;                                 // (uniform) that was the last round
            // the patch's loosest threshold: can anything that remains still matter?
            const float wm = warpMax(active ? w : 0.0f);
            if (lane == 0) sRed[warp][0] = wm;
            __syncthreads();
            float wPatch = sRed[0][0];
            #pragma unroll
            for (int wv = 1; wv < kS2Warps; ++wv) wPatch = fmaxf(wPatch, sRed[wv][0]);
            const float nextLb = (float)sBinOf[sOrder[q0 + kS2Round]] / (float)kS2Bins;      // the rest has sqrt(lb) >= bin / 32
            if (nextLb * nextLb > wPatch * 1.0001f) break;                 // uniform
        }

        // ---- 4. merge the tile's four sorted lists: 12 rounds of the lexicographic minimum of the four heads ----
        int total = count + __shfl_xor_sync(0xffffffffu, count, 1);
        total += __shfl_xor_sync(0xffffffffu, total, 2);
        total = min(total, SGI_MAX_VPLS_PER_TILE);
        int mine[SGI_MAX_VPLS_PER_TILE / kS2Lanes];                        // lane g keeps (and stores) entries g, g+4, g+8 of the merged list
        #pragma unroll
        for (int r = 0; r < SGI_MAX_VPLS_PER_TILE; ++r) {
            float md = dist[0]; int mp = pos[0];
            #pragma unroll
            for (int o = 1; o < kS2Lanes; o <<= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, md, o); const int op = __shfl_xor_sync(0xffffffffu, mp, o);
                const bool take = od < md || (od == md && op < mp);
                md = take ? od : md; mp = take ? op : mp;
            }
            if (mp == pos[0] && md == dist[0]) {                           // this lane's head won (list positions are unique): pop it
                #pragma unroll
                for (int sl = 0; sl + 1 < SGI_MAX_VPLS_PER_TILE; ++sl) { dist[sl] = dist[sl + 1]; pos[sl] = pos[sl + 1]; }
                dist[SGI_MAX_VPLS_PER_TILE - 1] = FLOAT_MAX_; pos[SGI_MAX_VPLS_PER_TILE - 1] = 0x7fffffff;
            }
            if ((r & (kS2Lanes - 1)) == g) mine[r / kS2Lanes] = mp;
        }
        if (active) {
            int32_t* o = reinterpret_cast<int32_t*>(out + tile);
            if (g == 0) o[0] = total;
            #pragma unroll
            for (int i = 0; i < SGI_MAX_VPLS_PER_TILE / kS2Lanes; ++i) {
                const int r = i * kS2Lanes + g;
                if (r < total) {
                    o[1 + r] = __ldg(visible + 1 + mine[i]);
                    // the same list as positions in the compacted table, for the tile-list gather (library-owned side output)
                    tilePos[(size_t)tile * SGI_MAX_VPLS_PER_TILE + r] = (unsigned short)mine[i];
                }
            }
        }
        __syncthreads();                                                   // the shared arrays (and sNext) are reused by the next patch
    }
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

int patchScratchInts(const BandDims& bd) {        // ints of scratch the two launchers need: class counters + patch counter + class lists
    const int tilesX = bd.W / SGI_TILE_DIVISOR, tilesY = bd.rows / SGI_TILE_DIVISOR;
    const int maxPatches = ((tilesX + kS2X - 1) / kS2X) * ((tilesY + kS2Y - 1) / kS2Y);      // stage 2's patches
    return kPatchClasses * (2 + (maxPatches > 0 ? maxPatches : 1));
}

static int maxPatchesOf(const BandDims& bd) { return patchScratchInts(bd) / kPatchClasses - 2; }
// layout of the scratch: [0, 8) class counters, [8] stage 2's patch counter, [16, 16 + 8 * maxPatches) class lists

int launchStage1(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, const float* depth, const uint32_t* normal,
                 SgiVplStage1PerTileOutputs* out, int* patchScratch, const SgiVplData* lightTable, const int32_t* visible, float4* posRadius,
                 int32_t* visibleCopy) {
    const int tilesX = bd.W / SGI_TILE_DIVISOR, tilesY = bd.crows / SGI_TILE_DIVISOR;
    if (tilesX <= 0 || tilesY <= 0) return 0;
    cudaMemsetAsync(patchScratch, 0, 2 * kPatchClasses * sizeof(int), stream);
    dim3 grid((tilesX + kS1X - 1) / kS1X, (tilesY + kS1Y - 1) / kS1Y);
    stage1Kernel<<<grid, kS1Threads, 0, stream>>>(fc, bd, depth, normal, out, patchScratch, patchScratch + 2 * kPatchClasses, maxPatchesOf(bd),
                                                  lightTable, visible, posRadius, visibleCopy);
    return 1;
}

int launchStage2(cudaStream_t stream, const BandDims& bd, const SgiVplStage1PerTileOutputs* stage1, const float4* posRadius,
                 const int32_t* visible, SgiVplStage2PerTileOutputs* out, unsigned short* tilePos, int* patchScratch) {
    const int tilesX = bd.W / SGI_TILE_DIVISOR, tilesY = bd.crows / SGI_TILE_DIVISOR;
    if (tilesX <= 0 || tilesY <= 0) return 0;
    const int patchesX = (tilesX + kS2X - 1) / kS2X, patchesY = (tilesY + kS2Y - 1) / kS2Y;
    const int numPatches = patchesX * patchesY;
    static int slots = 0;                     // resident CTAs of the persistent grid (every device of a box is the same part)
    if (slots == 0) {
        int sms = 148, dev = 0, perSm = 1;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, stage2Kernel, kS2Threads, 0) != cudaSuccess || perSm < 1) perSm = 1;
        slots = sms * perSm;
    }
    const int grid = max(1, min(numPatches, slots));
    stage2Kernel<<<grid, kS2Threads, 0, stream>>>(tilesX, bd.c0 / SGI_TILE_DIVISOR, tilesY, patchesX, numPatches, stage1, posRadius, visible, out, tilePos,
                                                  patchScratch, patchScratch + 2 * kPatchClasses, maxPatchesOf(bd), patchScratch + kPatchClasses);
    return 1;
}

}  // namespace sgi
