// sgi_internal.h — handle layout and stage launchers shared by the .cu files of libstratusgi.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include "../../include/stratusgi.h"

namespace sgi {

// per-launch geometry of a row band (see include/stratusgi.h "Row bands")
struct BandDims {
    int W;            // frame width
    int fullH;        // frame height
    int yOffset;      // first frame row held by the planes
    int rows;         // rows held by the planes
    int cubeRes;
    int csmRes;
    int c0;           // first held row this handle computes (band-local index; 0 for a whole frame)
    int crows;        // rows computed (== rows unless the handle is a band with halo rows)
};

// where a band's kernels also store their boundary rows: the neighbour's halo rows (NVLink peer memory)
struct PushTarget {
    uint2* res;                // neighbour's reservoir plane written by this pass (half4), or nullptr
    uint2* gi;                 // neighbour's GI plane of this frame (half4; final pass)
    int yOffset;               // neighbour's first held frame row
    int firstRow, lastRow;     // frame rows [firstRow, lastRow) of mine that the neighbour holds as halo
};
struct BandLinks {
    PushTarget up, down;
    uint2* image;              // root's frame image (half4 [fullH][W]) or nullptr
    // in-kernel wait: CTAs that touch a halo (read it, or overwrite the neighbour's) spin until the neighbour's epoch
    // flag in THIS handle's slab reaches waitEpoch; all other CTAs of the launch start immediately
    const int32_t* waitUp;     // this handle's flags[0] (written by the band above) or nullptr
    const int32_t* waitDown;   // this handle's flags[1] (written by the band below) or nullptr
    int waitEpoch;
    int32_t* timeouts;
    // final pass of a band: reprojected previous-frame taps (vpl_pbr_gi_denoise.fs:280,352-360) that fall outside the rows this
    // handle holds are clamped to them and counted here (sgi_band_reprojection_misses); nullptr on whole-frame handles
    int32_t* reprojMisses;
    // in-kernel signal: the CTAs (gather: 32-pixel row segments) that store rows into a neighbour's halo are scheduled first and
    // counted; the last of them releases the neighbours' flags, so the neighbours see this pass's boundary rows a pass-time
    // before they need them and ranks no longer wait for each other stage by stage.  doneCounter == nullptr: the host
    // launches the one-thread signal kernel after the pass instead.
    int32_t* doneCounter;      // this handle's counter (device, zero between launches)
    int producers;             // how many producer CTAs / segments this launch has
    int32_t* sigUp;            // the upper neighbour's flag word for me (its flags[1]) or nullptr
    int32_t* sigDown;          // the lower neighbour's flag word for me (its flags[0]) or nullptr
    int sigEpoch;
};

constexpr int kGatherMaxDraws = 12;   // draws of one pixel's VPL sequence: <= 6 samples + 5 re-draws (vpl_pbr_gi.fs:138-171)

// uniforms in device-friendly form; passed by value as a __grid_constant__ kernel parameter
struct FrameConsts {
    float invPV[16];
    float viewPos[3];
    float time;
    float framesPerSecond;
    float minGiOcclusionFactor;
    float minRoughness;
    int   usePerceptualRoughness;
    float fogDensity;
    float fogColor[3];
    float sunDir[3];
    float sunColor[3];
    float cascadeProjViews[4][16];
    float cascadePlanes[3][4];
    int   totalNumLights;
    // hash(floatBitsToUint(time + 10000 k)) of the gather's k-th draw: the part of random(vec3(gl_FragCoord.xy, seed.z))
    // that is the same for every pixel (common.glsl:199-214, vpl_pbr_gi.fs:149-150)
    uint32_t hashZ[kGatherMaxDraws];
};

struct AtlasTable {
    const unsigned short* shadow[SGI_MAX_TOTAL_SHADOW_ATLASES];
    const uchar4* diffuse[SGI_MAX_TOTAL_SHADOW_ATLASES];
};

#ifdef __CUDACC__
__device__ __forceinline__ int ldAcquireSys(const int32_t* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stReleaseSys(int32_t* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
constexpr long long kSpinLimitCycles = 4000000000LL;      // ~2 s at 1.9 GHz: a dead neighbour must not hang the GPU
__device__ __forceinline__ bool spinUntil(const int32_t* flag, int epoch, long long t0) {
    while ((int)(ldAcquireSys(flag) - epoch) < 0) {           // wrap-safe "flag >= epoch"
        __nanosleep(100);
        if (clock64() - t0 > kSpinLimitCycles) return false;
    }
    return true;
}
// waits that found the flag behind are accounted in flags[5..7] of this handle (sgi_band_stats): how many, their summed and their
// longest duration in units of 1024 SM cycles — the evidence for "waiting for a neighbour" against "slower kernel"
__device__ __forceinline__ void bandSpinAccount(const BandLinks& L, long long t0) {
    const long long dt = clock64() - t0;
    if (dt > 2048) {
        int32_t* st = L.timeouts + 3;
        atomicAdd(st + 0, 1);
        atomicAdd(st + 1, (int)(dt >> 10));
        atomicMax(st + 2, (int)(dt >> 10));
    }
}
// Called by EVERY thread of a CTA.  If some thread needs a halo, thread 0 spins on the neighbour's flag and the CTA
// proceeds after a barrier (acquire + bar.sync orders the CTA's later loads/stores after the neighbour's release).
__device__ __forceinline__ void bandWaitCta(const BandLinks& L, bool needUp, bool needDown) {
    if (L.waitUp == nullptr && L.waitDown == nullptr) return;          // uniform across the launch
    const int up = __syncthreads_or(needUp && L.waitUp != nullptr);
    const int down = __syncthreads_or(needDown && L.waitDown != nullptr);
    if (up | down) {
        if (threadIdx.x == 0 && threadIdx.y == 0) {
            const long long t0 = clock64();
            bool ok = true;
            if (up) ok = spinUntil(L.waitUp, L.waitEpoch, t0) && ok;
            if (down) ok = spinUntil(L.waitDown, L.waitEpoch, t0) && ok;
            if (!ok) atomicAdd(L.timeouts, 1);
            bandSpinAccount(L, t0);
        }
        __syncthreads();
    }
}
// the last producer releases the neighbours' flags.  Called by ONE thread after a barrier that orders the CTA's (warp's) stores
// before it: fence.sys makes them visible system-wide before the count, and the thread that sees the full count fences again
// (acquire of the other producers' fenced stores) before the release stores.
__device__ __forceinline__ void producerDone(const BandLinks& L) {
    __threadfence_system();
    if (atomicAdd(L.doneCounter, 1) == L.producers - 1) {
        __threadfence_system();
        *L.doneCounter = 0;                                   // re-armed for the next launch on this stream
        if (L.sigUp != nullptr) stReleaseSys(L.sigUp, L.sigEpoch);
        if (L.sigDown != nullptr) stReleaseSys(L.sigDown, L.sigEpoch);
    }
}
__host__ __device__ inline bool pushTarget(const PushTarget& t) { return t.res != nullptr || t.gi != nullptr; }      // something of mine goes there
__device__ __forceinline__ bool pushesRow(const BandLinks& L, int y) {      // frame row y is stored into a neighbour's halo
    return (pushTarget(L.up) && y >= L.up.firstRow && y < L.up.lastRow) || (pushTarget(L.down) && y >= L.down.firstRow && y < L.down.lastRow);
}
// store a band's boundary pixel into the neighbours' halo rows (plain st.global on peer-mapped pointers: NVLink writes)
__device__ __forceinline__ void pushRes(const BandLinks& L, int y, int x, int W, uint2 v) {
    if (L.up.res != nullptr && y >= L.up.firstRow && y < L.up.lastRow) L.up.res[(size_t)(y - L.up.yOffset) * W + x] = v;
    if (L.down.res != nullptr && y >= L.down.firstRow && y < L.down.lastRow) L.down.res[(size_t)(y - L.down.yOffset) * W + x] = v;
}
__device__ __forceinline__ void pushGi(const BandLinks& L, int y, int x, int W, uint2 gi) {
    if (L.up.gi != nullptr && y >= L.up.firstRow && y < L.up.lastRow) L.up.gi[(size_t)(y - L.up.yOffset) * W + x] = gi;
    if (L.down.gi != nullptr && y >= L.down.firstRow && y < L.down.lastRow) L.down.gi[(size_t)(y - L.down.yOffset) * W + x] = gi;
}
#endif

struct DenoiseIO {
    const uint2* screen;          // half4
    const uint32_t* velocity;     // half2
    const uint32_t* normal;       // rgba8
    const float* ids;
    const float* depth;
    const uint32_t* prevNormal;
    const float* prevIds;
    const float* prevDepth;
    const float4* guide;          // library-owned: (decoded normal, depth) of every pixel of this frame
    float4* temporal;             // library-owned: (reprojected previous GI, 1 / history) of every pixel — temporalKernel writes, the final pass reads
    const uint2* illum;           // half4: the demodulation plane (vplGIFbo attachment 0)
    const uint2* shadowsIn;       // half4: reservoir input of this pass
    const uint2* prevGi;          // half4
    const unsigned short* historyIn;   // half
    const int32_t* numVisible;    // vplVisibleIndices (count in slot 0)
    BandLinks links;              // peer destinations (all null for a whole-frame handle)
    uint2* reservoirOut;          // half4
    uint2* combinedOut;           // half4 (final pass)
    uint2* giOut;                 // half4 (final pass)
    unsigned short* historyOut;   // half  (final pass)
    int multiplier;
    // outputs inside this pixel rectangle (band-local rows) are written by the TMA-staged tiles (k_atrous_tma.cu); the CTAs of the
    // clamped grid (gridX x gridY CTAs of kDnX x kDnY threads) that lie entirely inside it — x tiles [bx0, bx1) of CTA rows
    // [by0, by1) — do not run, the others (the "border CTAs") skip the covered outputs one by one
    int skipX0, skipX1, skipY0, skipY1;
    int borderGridX, borderGridY, borderBx0, borderBx1, borderBy0, borderBy1;
};

}  // namespace sgi

struct SgiHandle_ {
    SgiCreateParams cp;
    SgiBuffers buf;
    bool bound = false;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    int64_t launches = 0;
    std::string lastError;
    // sgi_frame runs the tile-list stages on this low-priority side stream next to the stochastic gather + denoiser (which do
    // not read the lists): forked after the cull, joined at the end of the frame
    cudaStream_t tileStream = nullptr;
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    // sgi_early_begin / sgi_cull_early: the cull (+ colouring) of the NEXT frame runs on a caller-chosen second stream under the
    // current frame's denoise passes.  evTablesFree: the current frame's last reader of the compacted tables and the atlases on
    // the main stream (the gather) is done; evStage1 covers the forked tile stages; evEarly: the early cull is complete.
    cudaEvent_t evTablesFree = nullptr, evEarly = nullptr;
    bool tablesFreeRecorded = false;
    bool earlyPending = false, earlyColored = false;
    SgiFrameParams earlyParams{};   // what the pending early cull ran with: the frame that consumes it must agree
    // row bands: a non-owner band sends its lit rows to the owner's frame image with a copy-engine transfer on this stream
    // after its final pass (the transfer overlaps the next frame), then raises its flag in the owner's slab
    cudaStream_t copyStream = nullptr;
    cudaStream_t auxStream = nullptr;     // the temporal pre-pass of sgi_frame (low priority, under the non-final à-trous passes)
    cudaEvent_t evFinal = nullptr, evCopyDone = nullptr;
    bool copyPending = false;
    int64_t imageWaitedFrame = 0;   // owner: last frame whose image the stream has waited for
    // library-owned planes
    void* giColor = nullptr;        // half4
    void* giReservoir = nullptr;    // half4
    void* resA = nullptr;           // half4 (vplGIDenoisedFbo1 attachment 2)
    void* resB = nullptr;           // half4 (vplGIDenoisedFbo2 attachment 2)
    void* combined = nullptr;       // half4 (used when lightingOut is not bound)
    void* giIllum[2] = {nullptr, nullptr};    // half4, [cur] / [cur^1] = previous frame
    void* history[2] = {nullptr, nullptr};    // half
    int cur = 0;
    int* patchScratch = nullptr;    // stage-1 -> stage-2 patch size classes (k_tiles.cu)
    float4* posRadius = nullptr;         // [SGI_MAX_TOTAL_VPLS_PER_FRAME] (position, radius) of the compacted VPLs, packed by stage 1 for stage 2
    int32_t* visibleCopy = nullptr;      // [SGI_MAX_TOTAL_VPLS_PER_FRAME + 1] stage 1's copy of the visible list for stage 2
    cudaEvent_t evStage1 = nullptr;      // stage 1 of the forked tile stages is done: the caller's tables are free (sgi_early_begin)
    bool stage1Outstanding = false;
    unsigned short* tilePos = nullptr;   // [tiles][12] per-tile lists as positions in the compacted VPL table (tile-list gather)
    bool tileListsValid = false;    // sgi_tile_lists ran since the last bind
    void* guide = nullptr;          // float4 (normalize(normal*2-1), depth), rebuilt once per frame
    void* temporal = nullptr;       // float4 (reprojected previous GI, 1 / history): the temporal pre-pass of this frame (k_denoise.cu)
    bool temporalDone = false;      // launched for the frame in flight (sgi_frame forks it; the final pass launches it if nobody did)
    bool temporalForked = false;    // ... on the side stream: the final pass waits for evTemporal
    bool gatherDoneThisFrame = false;   // (bands) frameNo already counts the frame in flight
    cudaEvent_t evTemporal = nullptr;
    bool guideValid = false;        // set by the gather (or the first denoise pass), cleared by bind and the final pass
    bool pass0Fused = false;        // the gather of this frame already wrote denoise pass 0 into resA, for `pass0Time`
    float pass0Time = 0.0f;
    // one cudaMalloc slab holds the planes other bands may write into (resA, resB, giIllum, history, flags, frame images)
    void* slab = nullptr;
    size_t slabBytes = 0;
    int64_t offRes[2] = {0, 0}, offGi[2] = {0, 0}, offHist[2] = {0, 0}, offFlags = 0, offImage[2] = {-1, -1};
    int32_t* flags = nullptr;       // [SGI_BAND_FLAGS] inside the slab
    int32_t* bandDone = nullptr;    // producer counter of the in-kernel signal (BandLinks::doneCounter)
    void* frameImage[2] = {nullptr, nullptr};
    bool banded = false;
    SgiBandParams band{};
    struct Peer { bool attached = false; bool ipc = false; char* base = nullptr; SgiPeerInfo info{}; } peer[3];
    int64_t frameNo = 0;            // frames started through sgi_gather since creation (band epochs)
    // scratch for the cull
    unsigned long long* prepassKeys = nullptr;   // [prepassCap] sort keys of the VPL pre-pass (k_prepass.cu), grown on demand
    int prepassCap = 0;
    int* prepassInside = nullptr;   // lights inside the frustum (before the 10000 cap)
    int* gatherCounters = nullptr;  // two work counters of the persistent gather, used alternately (k_gather.cu)
    float* taaTables = nullptr;     // [kTaaTableFloats] scratch of sgi_taa (k_taa.cu)
    float* coordU = nullptr;        // [width] (x + 0.5) / W, and [rows] ((y + 0.5) / fullH, hash(y + 0.5)): rebuilt with the planes
    float2* coordRow = nullptr;
    int64_t gatherLaunches = 0;
    uint32_t* cullBallots = nullptr;
    unsigned int* cullTicket = nullptr;
    int32_t* pinnedCount = nullptr;
};

namespace sgi {

void fillFrameConsts(const SgiFrameParams* fp, FrameConsts* fc);
bool preloadPrepassKernels(); int launchPrepass(cudaStream_t stream, const SgiPrepassParams& pp, unsigned long long* keys, int* inside);
bool preloadDirectKernels();
bool preloadCubemapKernels(); int launchCubemaps(cudaStream_t stream, const SgiCubeRenderParams& rp, const SgiVplData* lightData, const SgiAtlasEntry* atlasByVpl, int res, const AtlasTable& atlases);
constexpr int kTaaTableFloats = 512;     // tonemap / inverse-tonemap tables of the 256 texel values (k_taa.cu)
bool preloadTaaKernels(); int launchTaa(cudaStream_t stream, const BandDims& bd, const SgiTaaParams& tp, float* tables /* [kTaaTableFloats], device */);
bool preloadSsaoKernels(); int launchSsao(cudaStream_t stream, const BandDims& bd, const SgiSsaoParams& sp);
int launchDirectLighting(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, const float* depth, const uint32_t* normal, const uint32_t* albedo,
                         const uint16_t* refl, const uint32_t* rma, const unsigned short* ssao, const float* csm, const AtlasTable& atlases,
                         const SgiDirectParams& dp, uint2* out);
bool preloadAtrousTmaKernels(); bool preloadVplKernels(); bool preloadTileKernels(); bool preloadGatherKernels(); bool preloadDenoiseKernels(); bool preloadBandKernels();

// each launcher enqueues its kernels on `stream` and returns the number of launches made
int launchCull(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, const SgiVplData* lightData, const SgiAtlasEntry* atlasByVpl,
               const float* csm, SgiVplData* updated, int32_t* visible, SgiAtlasEntry* entries, uint32_t* ballots, unsigned int* ticket,
               const AtlasTable* colorAtlases /* non-null: the cull's last CTA also runs the colouring stage */,
               const int32_t* countDevice /* non-null: N = min(totalNumLights, *countDevice) */);
int launchColor(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, SgiVplData* updated, const int32_t* visible,
                const SgiAtlasEntry* entries, const AtlasTable& atlases);
// ---- row bands (sgi_band.cu) ----
BandLinks noLinks();
BandLinks bandLinks(SgiHandle_* h, int resIndex, bool final_, int waitEpoch);
void armInKernelSignal(SgiHandle_* h, BandLinks* L, int epoch);     // fills doneCounter / sigUp / sigDown / sigEpoch (producers is set by the launcher)
int signalOwnerRowsDone(SgiHandle_* h);                             // owner band: its own rows of the frame image are complete   // peer destinations of the pass writing reservoir resIndex; waitEpoch 0 = no wait
inline int bandEpoch(int64_t frameNo, int stage) { return (int)(uint32_t)(frameNo * 32 + stage); }   // stage k: output of pass k-1 pushed
int waitNeighbours(SgiHandle_* h, int epoch);
int signalNeighbours(SgiHandle_* h, int epoch, bool final_);
int waitAllBands(SgiHandle_* h, int64_t frameNo);
void detachPeers(SgiHandle_* h);
constexpr int kBandFlags = 64;       // int32 flags per handle: [0] up neighbour's epoch, [1] down neighbour's, [2] time-outs, [3] reprojection taps outside the held rows, [4] producer counter, [5..7] waits that spun: count, summed and longest duration (1024 cycles) of the in-kernel signal, [8+r] band r's image rows landed
int launchWaitFlags(cudaStream_t stream, const int32_t* f0, const int32_t* f1, int epoch, int32_t* timeouts);
int launchWaitAllBands(cudaStream_t stream, const int32_t* flags, int world, int epoch, int32_t* timeouts);
int launchSignal(cudaStream_t stream, int32_t* a, int32_t* b, int epoch, int32_t* c, int valueC);
int patchScratchInts(const BandDims& bd);
int launchStage1(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, const float* depth, const uint32_t* normal,
                 SgiVplStage1PerTileOutputs* out, int* patchScratch, const SgiVplData* lightTable, const int32_t* visible, float4* posRadius,
                 int32_t* visibleCopy /* [SGI_MAX_TOTAL_VPLS_PER_FRAME + 1]: stage 2 reads this copy, not the caller's list */);
int launchStage2(cudaStream_t stream, const BandDims& bd, const SgiVplStage1PerTileOutputs* stage1, const float4* posRadius,
                 const int32_t* visible, SgiVplStage2PerTileOutputs* out, unsigned short* tilePos, int* patchScratch);
int launchGather(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, const float* depth, const uint32_t* normal,
                 const uint32_t* albedo, const uint16_t* refl, const uint32_t* rma, const unsigned short* ssao,
                 const SgiVplData* updated, const int32_t* visible, const SgiAtlasEntry* entries, const AtlasTable& atlases,
                 uint2* colorOut, uint2* reservoirOut, float4* guideOut, uint2* mergedOut, const BandLinks& links,
                 const SgiVplStage2PerTileOutputs* tileLists /* non-null: tile-list mode */, const unsigned short* tilePos,
                 int* unitCounters /* two ints: work counters of the persistent gather */, int64_t* persistentLaunches /* selects the counter */,
                 bool* signalled /* out: the kernel releases the neighbours' flags itself */,
                 const float* coordU, const float2* coordRow /* launchCoordTables */);
int launchCoordTables(cudaStream_t stream, const BandDims& bd, float* u /* [W] */, float2* row /* [rows] */);   // k_gather.cu: per-column / per-row texture coordinates
int launchTemporal(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, const DenoiseIO& io);   // reads velocity / ids / normal / depth / prev* / prevGi / historyIn, writes historyOut + temporal
bool planAtrousTma(const BandDims& bd, int s, int* x0, int* x1, int* y0, int* y1);                              // k_atrous_tma.cu: the rectangle its tiles would cover
bool launchAtrousTma(cudaStream_t stream, const BandDims& bd, const DenoiseIO& io, bool final_, int nBorder);       // ... one launch: the border CTAs of io (borderCta), then the tiles
int launchPrepareGuide(cudaStream_t stream, const BandDims& bd, const uint32_t* normal, const float* depth, float4* guide);
int launchDenoisePass(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, const DenoiseIO& io, bool merge, bool final_,
                      bool* signalled /* out: the kernel releases the neighbours' flags itself */);

}  // namespace sgi
