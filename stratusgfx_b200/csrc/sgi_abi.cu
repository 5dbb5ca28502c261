// sgi_abi.cu — the C-ABI of include/stratusgi.h over the CUDA stages.
//
// Host-side sequencing mirrors the three reference methods
//   PerformVirtualPointLightCullingStage1_      (StratusRendererBackend.cpp:1578-1612)
//   PerformVirtualPointLightCullingStage2_      (:1666-1731, dormant tile stages :1733-1783)
//   ComputeVirtualPointLightGlobalIllumination_ (:1816-1952)
// with the fence + map round trip (:1609-1610, :1675-1701) removed: everything is enqueued on
// one stream and nothing blocks the host.  There is no CPU fallback anywhere in this library.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include "sgi_internal.h"

using namespace sgi;

#define SGI_CUDA(h, expr)                                                                         \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            char _b[512];                                                                         \
            snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            (h)->lastError = _b;                                                                  \
            return SGI_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

static thread_local std::string g_createError;

static int fail(SgiHandle h, int code, const char* msg) { h->lastError = msg; return code; }

static size_t pixels(const SgiHandle h) { return (size_t)h->cp.width * (size_t)h->cp.rows; }

static BandDims bandDims(const SgiHandle h);

static void freePlanes(SgiHandle h) {
    void** ps[] = {&h->giColor, &h->giReservoir, &h->combined, &h->guide, &h->temporal, &h->slab, (void**)&h->patchScratch, (void**)&h->tilePos,
                   (void**)&h->coordU, (void**)&h->coordRow};
    for (void** p : ps) { if (*p) cudaFree(*p); *p = nullptr; }
    h->resA = h->resB = nullptr;
    h->giIllum[0] = h->giIllum[1] = h->history[0] = h->history[1] = nullptr;
    h->frameImage[0] = h->frameImage[1] = nullptr;
    h->flags = nullptr; h->bandDone = nullptr;
}

// Planes that another band may store into (reservoir ping-pong, GI/history ping-pong, flags, frame images) live in ONE
// cudaMalloc slab, so a single CUDA IPC handle (or one peer-access base pointer) exposes them all (sgi_band.cu).
static int allocPlanes(SgiHandle h) {
    const size_t n = pixels(h);
    size_t off = 0;
    auto take = [&off](size_t bytes) { const size_t o = off; off = (off + bytes + 255) & ~size_t(255); return (int64_t)o; };
    h->offRes[0] = take(n * 8); h->offRes[1] = take(n * 8);
    h->offGi[0] = take(n * 8); h->offGi[1] = take(n * 8);
    h->offHist[0] = take(n * 2); h->offHist[1] = take(n * 2);
    h->offFlags = take(sgi::kBandFlags * sizeof(int32_t));
    h->offImage[0] = h->offImage[1] = -1;
    if (h->banded && h->band.owns_frame_image) {
        const size_t img = (size_t)h->cp.width * (size_t)h->cp.full_height * 8;
        h->offImage[0] = take(img); h->offImage[1] = take(img);
    }
    SGI_CUDA(h, cudaMalloc(&h->slab, off));
    h->slabBytes = off;
    char* base = (char*)h->slab;
    h->resA = base + h->offRes[0]; h->resB = base + h->offRes[1];
    for (int i = 0; i < 2; ++i) { h->giIllum[i] = base + h->offGi[i]; h->history[i] = base + h->offHist[i]; }
    h->flags = (int32_t*)(base + h->offFlags);
    h->bandDone = h->flags + 4;             // producer counter of the in-kernel signal (flags[4]; the slab's flags start at zero)
    for (int i = 0; i < 2; ++i) h->frameImage[i] = h->offImage[i] >= 0 ? base + h->offImage[i] : nullptr;
    void** half4s[] = {&h->giColor, &h->giReservoir, &h->combined};
    for (void** p : half4s) SGI_CUDA(h, cudaMalloc(p, n * 8));
    SGI_CUDA(h, cudaMalloc(&h->guide, n * 16));
    SGI_CUDA(h, cudaMalloc(&h->temporal, n * 16));
    SGI_CUDA(h, cudaMalloc((void**)&h->patchScratch, (size_t)patchScratchInts(bandDims(h)) * sizeof(int)));
    {
        const size_t tiles = (size_t)(h->cp.width / SGI_TILE_DIVISOR) * (size_t)(h->cp.rows / SGI_TILE_DIVISOR);
        SGI_CUDA(h, cudaMalloc((void**)&h->tilePos, (tiles > 0 ? tiles : 1) * SGI_MAX_VPLS_PER_TILE * sizeof(unsigned short)));
    }
    SGI_CUDA(h, cudaMalloc((void**)&h->coordU, (size_t)h->cp.width * sizeof(float)));
    SGI_CUDA(h, cudaMalloc((void**)&h->coordRow, (size_t)h->cp.rows * sizeof(float2)));
    sgi::launchCoordTables(h->stream, bandDims(h), h->coordU, h->coordRow);
    h->guideValid = false;
    h->tileListsValid = false;
    h->temporalDone = false;
    // per-frame clears to (0,0,0,1) only matter for planes that are read before written; the previous-frame FBO is
    // never cleared and starts at zero.  Reservoir halo rows start at zero too.
    SGI_CUDA(h, cudaMemsetAsync(h->slab, 0, off, h->stream));
    h->cur = 0;
    h->frameNo = 0;
    h->imageWaitedFrame = 0;
    h->copyPending = false;
    return SGI_OK;
}

int sgiAllocPlanesForBand(SgiHandle h) {     // sgi_band_setup (sgi_band.cu): the slab layout depends on the band parameters
    freePlanes(h);
    return allocPlanes(h);
}

static int checkDims(int width, int full_height, int y_offset, int rows) {
    if (width <= 0 || full_height <= 0 || rows <= 0 || y_offset < 0 || y_offset + rows > full_height) return SGI_ERR_INVALID_ARG;
    return SGI_OK;
}

static BandDims bandDims(const SgiHandle h) {
    BandDims bd;
    bd.W = h->cp.width; bd.fullH = h->cp.full_height; bd.yOffset = h->cp.y_offset; bd.rows = h->cp.rows;
    bd.cubeRes = h->cp.cube_res; bd.csmRes = h->cp.csm_res;
    bd.c0 = h->banded ? h->band.compute_y0 - h->cp.y_offset : 0;
    bd.crows = h->banded ? h->band.compute_rows : h->cp.rows;
    return bd;
}

// SGI_HALO_RECOMPUTE (include/stratusgi.h): pass `pass` of a band is computed on its own rows plus the rows the LATER passes of
// this band will read beyond them, sum_{j > pass} 2^(j+1), clipped to the held rows (a band at the frame's edge holds none there)
static bool recomputesHalo(const SgiHandle h) { return h->banded && h->band.halo_mode == SGI_HALO_RECOMPUTE; }
static BandDims bandDimsForPass(const SgiHandle h, int pass) {
    BandDims bd = bandDims(h);
    if (!recomputesHalo(h)) return bd;
    int ext = 0;
    for (int j = pass + 1; j < h->cp.denoise_iters; ++j) ext += 2 << j;
    const int up = ext < bd.c0 ? ext : bd.c0;
    const int below = bd.rows - (bd.c0 + bd.crows);
    const int down = ext < below ? ext : below;
    bd.c0 -= up; bd.crows += up + down;
    return bd;
}

static AtlasTable atlasTable(const SgiHandle h) {
    AtlasTable t;
    for (int i = 0; i < SGI_MAX_TOTAL_SHADOW_ATLASES; ++i) {
        t.shadow[i] = (const unsigned short*)h->buf.shadowCubeMaps[i];
        t.diffuse[i] = (const uchar4*)h->buf.diffuseCubeMaps[i];
    }
    return t;
}

extern "C" {

const char* sgi_version(void) { return "stratusgi-b200 0.1 (sm_100a)"; }

int sgi_create(const SgiCreateParams* params, SgiHandle* out) {
    if (!params || !out) return SGI_ERR_INVALID_ARG;
    *out = nullptr;
    if (checkDims(params->width, params->full_height, params->y_offset, params->rows) != SGI_OK) return SGI_ERR_INVALID_ARG;
    if (params->cube_res <= 0 || params->csm_res <= 0 || params->denoise_iters < 1 || params->denoise_iters > 16) return SGI_ERR_INVALID_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || params->device < 0 || params->device >= ndev) {
        cudaGetLastError();
        return SGI_ERR_CUDA;     // no usable CUDA device: there is deliberately no CPU path
    }
    SgiHandle h = new (std::nothrow) SgiHandle_();
    if (!h) return SGI_ERR_INVALID_ARG;
    h->cp = *params;
    std::memset(&h->buf, 0, sizeof(h->buf));
    auto bail = [&](int code) { sgi_destroy(h); return code; };
    if (cudaSetDevice(params->device) != cudaSuccess) return bail(SGI_ERR_CUDA);
    if (params->stream) { h->stream = (cudaStream_t)params->stream; h->ownStream = false; }
    else {
        if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(SGI_ERR_CUDA);
        h->ownStream = true;
    }
    {
        int lo = 0, hi = 0;
        if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) return bail(SGI_ERR_CUDA);      // lo = least priority
        if (cudaStreamCreateWithPriority(&h->tileStream, cudaStreamNonBlocking, lo) != cudaSuccess) return bail(SGI_ERR_CUDA);
        if (cudaEventCreateWithFlags(&h->evFork, cudaEventDisableTiming) != cudaSuccess) return bail(SGI_ERR_CUDA);
        if (cudaEventCreateWithFlags(&h->evJoin, cudaEventDisableTiming) != cudaSuccess) return bail(SGI_ERR_CUDA);
        if (cudaEventCreateWithFlags(&h->evTablesFree, cudaEventDisableTiming) != cudaSuccess) return bail(SGI_ERR_CUDA);
        if (cudaEventCreateWithFlags(&h->evEarly, cudaEventDisableTiming) != cudaSuccess) return bail(SGI_ERR_CUDA);
        if (cudaEventCreateWithFlags(&h->evTemporal, cudaEventDisableTiming) != cudaSuccess) return bail(SGI_ERR_CUDA);
        if (cudaStreamCreateWithFlags(&h->copyStream, cudaStreamNonBlocking) != cudaSuccess) return bail(SGI_ERR_CUDA);
        if (cudaStreamCreateWithPriority(&h->auxStream, cudaStreamNonBlocking, lo) != cudaSuccess) return bail(SGI_ERR_CUDA);
        if (cudaEventCreateWithFlags(&h->evFinal, cudaEventDisableTiming) != cudaSuccess) return bail(SGI_ERR_CUDA);
        if (cudaEventCreateWithFlags(&h->evCopyDone, cudaEventDisableTiming) != cudaSuccess) return bail(SGI_ERR_CUDA);
    }
    if (cudaMalloc(&h->cullBallots, 512 * sizeof(uint32_t)) != cudaSuccess) return bail(SGI_ERR_CUDA);
    if (cudaMalloc((void**)&h->posRadius, SGI_MAX_TOTAL_VPLS_PER_FRAME * sizeof(float4)) != cudaSuccess) return bail(SGI_ERR_CUDA);
    if (cudaMalloc((void**)&h->visibleCopy, (SGI_MAX_TOTAL_VPLS_PER_FRAME + 1) * sizeof(int32_t)) != cudaSuccess) return bail(SGI_ERR_CUDA);
    if (cudaEventCreateWithFlags(&h->evStage1, cudaEventDisableTiming) != cudaSuccess) return bail(SGI_ERR_CUDA);
    if (cudaMalloc(&h->cullTicket, sizeof(unsigned int)) != cudaSuccess) return bail(SGI_ERR_CUDA);
    if (cudaMalloc((void**)&h->gatherCounters, 2 * sizeof(int)) != cudaSuccess) return bail(SGI_ERR_CUDA);
    if (cudaMemsetAsync(h->gatherCounters, 0, 2 * sizeof(int), h->stream) != cudaSuccess) return bail(SGI_ERR_CUDA);
    if (cudaMemsetAsync(h->cullTicket, 0, sizeof(unsigned int), h->stream) != cudaSuccess) return bail(SGI_ERR_CUDA);
    if (cudaMallocHost(&h->pinnedCount, sizeof(int32_t)) != cudaSuccess) return bail(SGI_ERR_CUDA);
    if (cudaMalloc((void**)&h->prepassInside, sizeof(int)) != cudaSuccess) return bail(SGI_ERR_CUDA);
    if (!(preloadPrepassKernels() && preloadVplKernels() && preloadTileKernels() && preloadGatherKernels() && preloadDenoiseKernels() && preloadBandKernels() &&
          preloadDirectKernels() && preloadSsaoKernels() && preloadTaaKernels() && preloadCubemapKernels()
#ifdef __CUDACC__
          && preloadAtrousTmaKernels()            // (the per-kernel host-emulation builds of tests/cuda_on_host leave k_atrous_tma.cu out)
#endif
          )) return bail(SGI_ERR_CUDA);
    if (allocPlanes(h) != SGI_OK) return bail(SGI_ERR_CUDA);
    *out = h;
    return SGI_OK;
}

int sgi_destroy(SgiHandle h) {
    if (!h) return SGI_ERR_INVALID_ARG;
    cudaSetDevice(h->cp.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->tileStream) { cudaStreamSynchronize(h->tileStream); cudaStreamDestroy(h->tileStream); }
    if (h->copyStream) { cudaStreamSynchronize(h->copyStream); cudaStreamDestroy(h->copyStream); }
    if (h->auxStream) { cudaStreamSynchronize(h->auxStream); cudaStreamDestroy(h->auxStream); }
    if (h->evFinal) cudaEventDestroy(h->evFinal);
    if (h->evCopyDone) cudaEventDestroy(h->evCopyDone);
    if (h->evFork) cudaEventDestroy(h->evFork);
    if (h->evJoin) cudaEventDestroy(h->evJoin);
    if (h->evTablesFree) cudaEventDestroy(h->evTablesFree);
    if (h->evEarly) cudaEventDestroy(h->evEarly);
    if (h->evTemporal) cudaEventDestroy(h->evTemporal);
    sgi::detachPeers(h);
    freePlanes(h);
    if (h->cullBallots) cudaFree(h->cullBallots);
    if (h->posRadius) cudaFree(h->posRadius);
    if (h->visibleCopy) cudaFree(h->visibleCopy);
    if (h->evStage1) cudaEventDestroy(h->evStage1);
    if (h->cullTicket) cudaFree(h->cullTicket);
    if (h->gatherCounters) cudaFree(h->gatherCounters);
    if (h->taaTables) cudaFree(h->taaTables);
    if (h->prepassKeys) cudaFree(h->prepassKeys);
    if (h->prepassInside) cudaFree(h->prepassInside);
    if (h->pinnedCount) cudaFreeHost(h->pinnedCount);
    if (h->ownStream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return SGI_OK;
}

int sgi_resize(SgiHandle h, int32_t width, int32_t full_height, int32_t y_offset, int32_t rows) {
    if (!h) return SGI_ERR_INVALID_ARG;
    if (checkDims(width, full_height, y_offset, rows) != SGI_OK) return fail(h, SGI_ERR_INVALID_ARG, "sgi_resize: bad dimensions");
    SGI_CUDA(h, cudaSetDevice(h->cp.device));
    SGI_CUDA(h, cudaStreamSynchronize(h->stream));
    SGI_CUDA(h, cudaStreamSynchronize(h->tileStream));
    SGI_CUDA(h, cudaStreamSynchronize(h->copyStream));       // a band's rows may still be travelling to the owner's image
    SGI_CUDA(h, cudaStreamSynchronize(h->auxStream));
    sgi::detachPeers(h);
    freePlanes(h);
    h->banded = false;                      // a resized handle is a plain whole-range handle again; call sgi_band_setup to re-band
    h->cp.width = width; h->cp.full_height = full_height; h->cp.y_offset = y_offset; h->cp.rows = rows;
    h->bound = false;                       // per-pixel planes of the old size are no longer valid
    return allocPlanes(h);
}

int sgi_bind_device_buffers(SgiHandle h, const SgiBuffers* buffers) {
    if (!h || !buffers) return SGI_ERR_INVALID_ARG;
    h->buf = *buffers;
    h->bound = true;
    h->guideValid = false;
    h->pass0Fused = false;
    h->tileListsValid = false;
    h->temporalDone = false;
    return SGI_OK;
}

int sgi_register_gl(SgiHandle h, const uint32_t*, int32_t) {
    if (!h) return SGI_ERR_INVALID_ARG;
    return fail(h, SGI_ERR_UNSUPPORTED, "sgi_register_gl: library built without OpenGL headers (none in this image); use sgi_bind_device_buffers");
}

#define NEED(h, field)                                                                             \
    do { if (!(h)->bound || (h)->buf.field == nullptr) return fail(h, SGI_ERR_NOT_BOUND, "buffer not bound: " #field); } while (0)

static int postLaunch(SgiHandle h, int n) {
    h->launches += n;
    SGI_CUDA(h, cudaGetLastError());
    return SGI_OK;
}

static int cullImpl(SgiHandle h, const SgiFrameParams* fp, bool colorToo, cudaStream_t stream = nullptr) {
    if (!h || !fp) return SGI_ERR_INVALID_ARG;
    if (stream == nullptr) {
        stream = h->stream;
        if (h->earlyPending) {                   // an explicit cull supersedes a pending early one: order after it, then overwrite
            SGI_CUDA(h, cudaSetDevice(h->cp.device));
            SGI_CUDA(h, cudaStreamWaitEvent(h->stream, h->evEarly, 0));
            h->earlyPending = false;
        }
    }
    if (fp->totalNumLights < 0 || fp->totalNumLights > SGI_MAX_TOTAL_VPLS_BEFORE_CULLING) return fail(h, SGI_ERR_INVALID_ARG, "sgi_cull: totalNumLights out of range");
    NEED(h, vplData); NEED(h, vplUpdatedData); NEED(h, vplVisibleIndices); NEED(h, infiniteLightShadowMap);
    if (h->buf.atlasEntryByVpl && !h->buf.shadowDiffuseIndices) return fail(h, SGI_ERR_NOT_BOUND, "buffer not bound: shadowDiffuseIndices");
    if (colorToo) { NEED(h, shadowDiffuseIndices); NEED(h, atlasEntryByVpl); NEED(h, shadowCubeMaps[0]); NEED(h, diffuseCubeMaps[0]); }
    FrameConsts fc; fillFrameConsts(fp, &fc);
    SGI_CUDA(h, cudaSetDevice(h->cp.device));
    const AtlasTable atl = atlasTable(h);
    const int n = launchCull(stream, fc, bandDims(h), h->buf.vplData, h->buf.atlasEntryByVpl, (const float*)h->buf.infiniteLightShadowMap,
                             h->buf.vplUpdatedData, h->buf.vplVisibleIndices, h->buf.shadowDiffuseIndices, h->cullBallots, h->cullTicket,
                             colorToo ? &atl : nullptr, h->buf.vplCountDevice);
    h->tileListsValid = false;      // tilePos holds positions into the previous cull's compacted table
    return postLaunch(h, n);
}

int sgi_vpl_prepass(SgiHandle h, const SgiPrepassParams* pp) {
    if (!h || !pp) return SGI_ERR_INVALID_ARG;
    if (pp->numLights < 0 || pp->numLights > SGI_MAX_PREPASS_LIGHTS) return fail(h, SGI_ERR_INVALID_ARG, "sgi_vpl_prepass: numLights out of range");
    if (!pp->countOut || (pp->numLights > 0 && (!pp->lights || !pp->vplDataOut))) return fail(h, SGI_ERR_NOT_BOUND, "sgi_vpl_prepass: lights, vplDataOut and countOut are required");
    SGI_CUDA(h, cudaSetDevice(h->cp.device));
    if (pp->numLights > h->prepassCap) {                      // grow the key scratch (stream-ordered with respect to its last use)
        SGI_CUDA(h, cudaStreamSynchronize(h->stream));
        if (h->prepassKeys) cudaFree(h->prepassKeys);
        h->prepassKeys = nullptr; h->prepassCap = 0;
        const int cap = ((pp->numLights + 4095) / 4096) * 4096;
        SGI_CUDA(h, cudaMalloc((void**)&h->prepassKeys, (size_t)cap * sizeof(unsigned long long)));
        h->prepassCap = cap;
    }
    return postLaunch(h, launchPrepass(h->stream, *pp, h->prepassKeys, h->prepassInside));
}

int sgi_render_vpl_cubemaps(SgiHandle h, const SgiCubeRenderParams* rp) {
    if (!h || !rp) return SGI_ERR_INVALID_ARG;
    if (rp->numMaps < 0 || rp->numMaps > SGI_MAX_TOTAL_VPLS_BEFORE_CULLING || rp->numBoxes < 0 || rp->numBoxes > SGI_MAX_SCENE_BOXES)
        return fail(h, SGI_ERR_INVALID_ARG, "sgi_render_vpl_cubemaps: numMaps or numBoxes out of range");
    if (rp->numMaps == 0) return SGI_OK;
    if (!rp->vplIndices || (rp->numBoxes > 0 && !rp->boxes)) return fail(h, SGI_ERR_NOT_BOUND, "sgi_render_vpl_cubemaps: boxes and vplIndices are required");
    NEED(h, vplData); NEED(h, atlasEntryByVpl); NEED(h, shadowCubeMaps[0]); NEED(h, diffuseCubeMaps[0]);
    SGI_CUDA(h, cudaSetDevice(h->cp.device));
    return postLaunch(h, launchCubemaps(rp->stream ? (cudaStream_t)rp->stream : h->stream, *rp, h->buf.vplData, h->buf.atlasEntryByVpl, h->cp.cube_res, atlasTable(h)));
}

int sgi_ssao(SgiHandle h, const SgiSsaoParams* sp) {
    if (!h || !sp) return SGI_ERR_INVALID_ARG;
    if (!sp->structureBuffer || !sp->rotationLookup || !sp->occlusionScratch || !sp->ssaoOut) return fail(h, SGI_ERR_NOT_BOUND, "sgi_ssao: structureBuffer, rotationLookup, occlusionScratch and ssaoOut are required");
    if (h->banded || h->cp.rows != h->cp.full_height) return fail(h, SGI_ERR_UNSUPPORTED, "sgi_ssao: whole-frame handles only (the taps of a band would need halo rows)");
    SGI_CUDA(h, cudaSetDevice(h->cp.device));
    return postLaunch(h, launchSsao(h->stream, bandDims(h), *sp));
}

int sgi_taa(SgiHandle h, const SgiTaaParams* tp) {
    if (!h || !tp) return SGI_ERR_INVALID_ARG;
    if (!tp->screen || !tp->prevScreen || !tp->velocity || !tp->colorOut) return fail(h, SGI_ERR_NOT_BOUND, "sgi_taa: screen, prevScreen, velocity and colorOut are required");
    if (tp->historyOut && tp->historyOut == tp->prevScreen) return fail(h, SGI_ERR_INVALID_ARG, "sgi_taa: historyOut may not alias prevScreen (other pixels still read it)");
    if (h->banded || h->cp.rows != h->cp.full_height) return fail(h, SGI_ERR_UNSUPPORTED, "sgi_taa: whole-frame handles only");
    SGI_CUDA(h, cudaSetDevice(h->cp.device));
    if (h->taaTables == nullptr) SGI_CUDA(h, cudaMalloc((void**)&h->taaTables, sgi::kTaaTableFloats * sizeof(float)));
    return postLaunch(h, launchTaa(h->stream, bandDims(h), *tp, h->taaTables));
}

int sgi_direct_lighting(SgiHandle h, const SgiFrameParams* fp, const SgiDirectParams* dp) {
    if (!h || !fp || !dp) return SGI_ERR_INVALID_ARG;
    if (dp->numLights < 0 || dp->numShadowLights < 0 || !dp->screenOut) return fail(h, SGI_ERR_INVALID_ARG, "sgi_direct_lighting: bad light counts or no output plane");
    if ((dp->numLights > 0 && !dp->nonShadowCasters) || (dp->numShadowLights > 0 && (!dp->shadowCasters || !dp->shadowIndices)))
        return fail(h, SGI_ERR_NOT_BOUND, "sgi_direct_lighting: light tables missing");
    NEED(h, depth); NEED(h, normal); NEED(h, albedo); NEED(h, baseReflectivity); NEED(h, roughnessMetallicAmbient); NEED(h, ssao);
    if (dp->sunEnabled) NEED(h, infiniteLightShadowMap);
    if (dp->numShadowLights > 0) NEED(h, shadowCubeMaps[0]);
    FrameConsts fc; fillFrameConsts(fp, &fc);
    SGI_CUDA(h, cudaSetDevice(h->cp.device));
    const int n = launchDirectLighting(h->stream, fc, bandDims(h), (const float*)h->buf.depth, (const uint32_t*)h->buf.normal, (const uint32_t*)h->buf.albedo,
                                       (const uint16_t*)h->buf.baseReflectivity, (const uint32_t*)h->buf.roughnessMetallicAmbient,
                                       (const unsigned short*)h->buf.ssao, (const float*)h->buf.infiniteLightShadowMap, atlasTable(h), *dp, (uint2*)dp->screenOut);
    return postLaunch(h, n);
}

int sgi_cull(SgiHandle h, const SgiFrameParams* fp) { return cullImpl(h, fp, false); }

int sgi_early_begin(SgiHandle h, void* stream) {
    if (!h || !stream) return SGI_ERR_INVALID_ARG;
    SGI_CUDA(h, cudaSetDevice(h->cp.device));
    if (h->tablesFreeRecorded) SGI_CUDA(h, cudaStreamWaitEvent((cudaStream_t)stream, h->evTablesFree, 0));
    if (h->stage1Outstanding) SGI_CUDA(h, cudaStreamWaitEvent((cudaStream_t)stream, h->evStage1, 0));      // stage 2 reads stage 1's copies
    return SGI_OK;
}

int sgi_cull_early(SgiHandle h, const SgiFrameParams* fp, void* stream, int32_t with_color) {
    if (!h || !fp || !stream) return SGI_ERR_INVALID_ARG;
    if (h->earlyPending) return fail(h, SGI_ERR_INVALID_ARG, "sgi_cull_early: the previous early cull has not been consumed by sgi_frame / sgi_cull yet");
    const int rc = cullImpl(h, fp, with_color != 0, (cudaStream_t)stream);
    if (rc != SGI_OK) return rc;
    SGI_CUDA(h, cudaEventRecord(h->evEarly, (cudaStream_t)stream));
    h->earlyPending = true; h->earlyColored = with_color != 0; h->earlyParams = *fp;
    return SGI_OK;
}

int sgi_color(SgiHandle h, const SgiFrameParams* fp) {
    if (!h || !fp) return SGI_ERR_INVALID_ARG;
    NEED(h, vplUpdatedData); NEED(h, vplVisibleIndices); NEED(h, shadowDiffuseIndices); NEED(h, shadowCubeMaps[0]); NEED(h, diffuseCubeMaps[0]);
    FrameConsts fc; fillFrameConsts(fp, &fc);
    SGI_CUDA(h, cudaSetDevice(h->cp.device));
    const int n = launchColor(h->stream, fc, bandDims(h), h->buf.vplUpdatedData, h->buf.vplVisibleIndices, h->buf.shadowDiffuseIndices, atlasTable(h));
    return postLaunch(h, n);
}

static int tileListsOn(SgiHandle h, const SgiFrameParams* fp, cudaStream_t stream) {
    if (!h || !fp) return SGI_ERR_INVALID_ARG;
    NEED(h, depth); NEED(h, normal); NEED(h, vplUpdatedData); NEED(h, vplVisibleIndices); NEED(h, vplStage1Results); NEED(h, vplVisiblePerTile);
    if (h->cp.width % SGI_TILE_DIVISOR || h->cp.rows % SGI_TILE_DIVISOR || h->cp.y_offset % SGI_TILE_DIVISOR)
        return fail(h, SGI_ERR_INVALID_ARG, "sgi_tile_lists: width, rows and y_offset must be multiples of 5");
    FrameConsts fc; fillFrameConsts(fp, &fc);
    SGI_CUDA(h, cudaSetDevice(h->cp.device));
    const BandDims bd = bandDims(h);
    int n = launchStage1(stream, fc, bd, (const float*)h->buf.depth, (const uint32_t*)h->buf.normal, h->buf.vplStage1Results, h->patchScratch,
                         h->buf.vplUpdatedData, h->buf.vplVisibleIndices, h->posRadius, h->visibleCopy);
    if (stream != h->stream) {               // forked (sgi_frame): from here on the tile stages read library memory only
        SGI_CUDA(h, cudaEventRecord(h->evStage1, stream));
        h->stage1Outstanding = true;
    }
    // the compacted table (entry k = lightData[visible[k+1]], written by the cull) stands in for the indexed reads of stage2.cs:81-82
    n += launchStage2(stream, bd, h->buf.vplStage1Results, h->posRadius, h->visibleCopy, h->buf.vplVisiblePerTile, h->tilePos, h->patchScratch);
    h->tileListsValid = true;
    return postLaunch(h, n);
}

int sgi_tile_lists(SgiHandle h, const SgiFrameParams* fp) { return h ? tileListsOn(h, fp, h->stream) : SGI_ERR_INVALID_ARG; }

int sgi_gather(SgiHandle h, const SgiFrameParams* fp) {
    if (!h || !fp) return SGI_ERR_INVALID_ARG;
    if (fp->gatherMode != SGI_GATHER_STOCHASTIC && fp->gatherMode != SGI_GATHER_TILE_LIST) return fail(h, SGI_ERR_INVALID_ARG, "sgi_gather: unknown gatherMode");
    const bool tileMode = fp->gatherMode == SGI_GATHER_TILE_LIST;
    if (tileMode) {
        NEED(h, vplVisiblePerTile);
        if (!h->tileListsValid) return fail(h, SGI_ERR_NOT_BOUND, "sgi_gather: SGI_GATHER_TILE_LIST needs sgi_tile_lists first (since the last bind and the last cull)");
    }
    NEED(h, depth); NEED(h, normal); NEED(h, albedo); NEED(h, baseReflectivity); NEED(h, roughnessMetallicAmbient); NEED(h, ssao);
    NEED(h, vplUpdatedData); NEED(h, vplVisibleIndices); NEED(h, shadowDiffuseIndices); NEED(h, shadowCubeMaps[0]);
    FrameConsts fc; fillFrameConsts(fp, &fc);
    SGI_CUDA(h, cudaSetDevice(h->cp.device));
    int n = 0;
    const bool fuse = h->cp.denoise_iters > 1;
    BandLinks links = sgi::noLinks();
    if (h->banded) {
        // a new band frame starts here.  Neighbours must be done with the previous frame (their final pass read the halo rows
        // this gather is about to overwrite) before the boundary rows of pass 0 are pushed into their memory.
        ++h->frameNo;
        h->gatherDoneThisFrame = true;
        const int prevDone = h->frameNo > 1 ? sgi::bandEpoch(h->frameNo - 1, h->cp.denoise_iters) : 0;
        if (recomputesHalo(h)) {
            // nothing of this band's gather or non-final passes leaves the GPU, so there is nothing to wait for here: the
            // neighbours' "previous frame complete" flag is awaited by the temporal pre-pass, which the final pass (the only
            // kernel that stores into a neighbour: its GI rows) is ordered after
        } else if (fuse) {
            links = sgi::bandLinks(h, 0, false, prevDone);               // the CTAs that push wait inside the kernel
            sgi::armInKernelSignal(h, &links, sgi::bandEpoch(h->frameNo, 1));
        } else if (prevDone) n += sgi::waitNeighbours(h, prevDone);
    }
    bool signalled = false;
    n += launchGather(h->stream, fc, bandDimsForPass(h, 0), (const float*)h->buf.depth, (const uint32_t*)h->buf.normal, (const uint32_t*)h->buf.albedo,
                               (const uint16_t*)h->buf.baseReflectivity, (const uint32_t*)h->buf.roughnessMetallicAmbient,
                               (const unsigned short*)h->buf.ssao, h->buf.vplUpdatedData, h->buf.vplVisibleIndices, h->buf.shadowDiffuseIndices,
                               atlasTable(h), (uint2*)h->giColor, (uint2*)h->giReservoir, (float4*)h->guide,
                               fuse ? (uint2*)h->resA : nullptr, links, tileMode ? h->buf.vplVisiblePerTile : nullptr, h->tilePos,
                               h->gatherCounters, &h->gatherLaunches, &signalled, h->coordU, h->coordRow);
    if (h->banded && fuse && !signalled && !recomputesHalo(h)) n += sgi::signalNeighbours(h, sgi::bandEpoch(h->frameNo, 1), false);
    SGI_CUDA(h, cudaEventRecord(h->evTablesFree, h->stream));      // the compacted tables and the atlases have been read (sgi_early_begin)
    h->tablesFreeRecorded = true;
    h->guideValid = true;
    // pass 0 of the denoiser (centre-texel merge) is a per-pixel function of the reservoir: the gather has written it
    // into vplGIDenoisedFbo1's reservoir attachment already, valid for a pass-0 call with the same `time`
    h->pass0Fused = h->cp.denoise_iters > 1;
    h->pass0Time = fp->time;
    return postLaunch(h, n);
}

// The temporal pre-pass of the frame in flight (k_denoise.cu temporalKernel): reads the previous frame's GI / history / G-buffer
// planes and this frame's velocity / id / normal / depth, writes this frame's history and the (prevGi, 1 / history) plane the final
// pass blends with.  sgi_frame forks it onto the side stream at the start of the frame; a caller of the staged entry points gets it
// on the main stream right before the final pass.
static int temporalOn(SgiHandle h, const SgiFrameParams* fp, cudaStream_t stream) {
    NEED(h, depth); NEED(h, normal); NEED(h, velocity); NEED(h, id); NEED(h, prevNormal); NEED(h, prevId); NEED(h, prevDepth);
    FrameConsts fc; fillFrameConsts(fp, &fc);
    DenoiseIO io;
    std::memset(&io, 0, sizeof(io));
    io.velocity = (const uint32_t*)h->buf.velocity; io.normal = (const uint32_t*)h->buf.normal;
    io.ids = (const float*)h->buf.id; io.depth = (const float*)h->buf.depth;
    io.prevNormal = (const uint32_t*)h->buf.prevNormal; io.prevIds = (const float*)h->buf.prevId; io.prevDepth = (const float*)h->buf.prevDepth;
    io.prevGi = (const uint2*)h->giIllum[h->cur ^ 1];
    io.historyIn = (const unsigned short*)h->history[h->cur ^ 1];
    io.historyOut = (unsigned short*)h->history[h->cur];
    io.temporal = (float4*)h->temporal;
    io.links = sgi::noLinks();
    int n = 0;
    if (h->banded) {
        io.links.reprojMisses = h->flags + 3;
        // the reprojected taps may land in halo rows of the previous frame's GI plane, which the neighbours' final passes stored:
        // wait (in stream order) until both have completed that frame.  frameNo still counts the previous frame here when this
        // runs before the frame's gather.
        const int64_t prevFrame = h->gatherDoneThisFrame ? h->frameNo - 1 : h->frameNo;
        if (prevFrame > 0) {
            const int32_t* f0 = h->peer[SGI_PEER_UP].attached ? h->flags + 0 : nullptr;
            const int32_t* f1 = h->peer[SGI_PEER_DOWN].attached ? h->flags + 1 : nullptr;
            n += sgi::launchWaitFlags(stream, f0, f1, sgi::bandEpoch(prevFrame, h->cp.denoise_iters), h->flags + 2);
        }
    }
    n += launchTemporal(stream, fc, bandDims(h), io);
    h->temporalDone = true;
    return postLaunch(h, n);
}

int sgi_denoise_pass(SgiHandle h, const SgiFrameParams* fp, int32_t pass) {
    if (!h || !fp) return SGI_ERR_INVALID_ARG;
    const int iters = h->cp.denoise_iters;
    if (pass < 0 || pass >= iters) return fail(h, SGI_ERR_INVALID_ARG, "sgi_denoise_pass: pass out of range");
    NEED(h, depth); NEED(h, normal); NEED(h, vplVisibleIndices);
    const bool merge = pass < 1;                 // maxReservoirMergingPasses = 1 (BE:1913)
    const bool final_ = (pass + 1 == iters);     // BE:1929-1931
    if (final_) { NEED(h, screen); NEED(h, velocity); NEED(h, id); NEED(h, prevNormal); NEED(h, prevId); NEED(h, prevDepth); }
    FrameConsts fc; fillFrameConsts(fp, &fc);
    SGI_CUDA(h, cudaSetDevice(h->cp.device));
    DenoiseIO io;
    std::memset(&io, 0, sizeof(io));
    io.screen = (const uint2*)h->buf.screen; io.velocity = (const uint32_t*)h->buf.velocity; io.normal = (const uint32_t*)h->buf.normal;
    io.ids = (const float*)h->buf.id; io.depth = (const float*)h->buf.depth;
    io.prevNormal = (const uint32_t*)h->buf.prevNormal; io.prevIds = (const float*)h->buf.prevId; io.prevDepth = (const float*)h->buf.prevDepth;
    io.illum = (const uint2*)h->giColor;
    void* fbo[2] = {h->resA, h->resB};           // buffers[bufferIndex % 2] (BE:1921)
    io.shadowsIn = (const uint2*)(pass == 0 ? h->giReservoir : fbo[(pass - 1) % 2]);
    io.reservoirOut = (uint2*)fbo[pass % 2];
    io.prevGi = (const uint2*)h->giIllum[h->cur ^ 1];
    io.historyIn = (const unsigned short*)h->history[h->cur ^ 1];
    io.numVisible = h->buf.vplVisibleIndices;
    io.combinedOut = (uint2*)(h->buf.lightingOut ? h->buf.lightingOut : h->combined);
    io.giOut = (uint2*)h->giIllum[h->cur];
    io.historyOut = (unsigned short*)h->history[h->cur];
    io.multiplier = (1 << pass) - 1;             // std::pow(2, i) - 1 (BE:1920)
    io.guide = (const float4*)h->guide;
    int n = 0;
    if (h->banded && h->frameNo == 0) return fail(h, SGI_ERR_INVALID_ARG, "sgi_denoise_pass: a band handle needs sgi_gather first in every frame");
    if (pass == 0 && !final_ && h->pass0Fused && std::memcmp(&h->pass0Time, &fp->time, sizeof(float)) == 0) {
        h->pass0Fused = false;                   // resA already holds this pass's output (written by sgi_gather)
        return SGI_OK;
    }
    if (pass == 0) h->pass0Fused = false;
    if (!h->guideValid) {                        // denoise called without this library's gather in the same frame
        n += launchPrepareGuide(h->stream, bandDims(h), (const uint32_t*)h->buf.normal, (const float*)h->buf.depth, (float4*)h->guide);
        h->guideValid = true;
    }
    const bool recompute = recomputesHalo(h);
    if (h->banded && recompute) {
        // the halo rows this pass reads were computed by this band's previous pass; only the final pass stores into the
        // neighbours (GI rows of the next frame's reprojection) and releases their "frame complete" flags
        if (final_) {
            io.links = sgi::bandLinks(h, pass % 2, true, 0);
            sgi::armInKernelSignal(h, &io.links, sgi::bandEpoch(h->frameNo, pass + 1));
        }
    } else if (h->banded) {
        // pass p reads the neighbours' pass p-1 rows from this handle's halo: wait (in stream order) for their flags
        // (inside the kernel, by the CTAs whose taps reach the halo; interior CTAs start at once)
        io.links = sgi::bandLinks(h, pass % 2, final_, pass >= 1 ? sgi::bandEpoch(h->frameNo, pass) : 0);
        sgi::armInKernelSignal(h, &io.links, sgi::bandEpoch(h->frameNo, pass + 1));
    }
    if (final_) {
        if (!h->temporalDone) { const int rc = temporalOn(h, fp, h->stream); if (rc != SGI_OK) return rc; }
        else if (h->temporalForked) SGI_CUDA(h, cudaStreamWaitEvent(h->stream, h->evTemporal, 0));
        io.temporal = (float4*)h->temporal;
    }
    bool signalled = false;
    const bool copyOut = h->banded && final_ && !h->band.owns_frame_image && h->peer[SGI_PEER_ROOT].attached &&
                         h->peer[SGI_PEER_ROOT].info.off_image[(int)(h->frameNo & 1)] >= 0;
    if (copyOut && h->copyPending) SGI_CUDA(h, cudaStreamWaitEvent(h->stream, h->evCopyDone, 0));   // the previous frame's rows have left `combined`
    n += launchDenoisePass(h->stream, fc, bandDimsForPass(h, pass), io, merge, final_, &signalled);
    if (merge) SGI_CUDA(h, cudaEventRecord(h->evTablesFree, h->stream));      // an unfused pass 0 reads the visible count after the gather
    if (h->banded) {
        // the pass's boundary CTAs released the neighbours' flags themselves; otherwise a one-thread kernel does it after the pass
        if (recompute && !final_) {}             // nobody waits for a non-final pass of a band that recomputes its halo
        else if (!signalled) n += sgi::signalNeighbours(h, sgi::bandEpoch(h->frameNo, pass + 1), final_);
        else if (final_) n += sgi::signalOwnerRowsDone(h);
        if (copyOut) {
            // the band's lit rows -> the owner's frame image of this parity, by the copy engine, then this band's flag in the
            // owner's slab; both on the copy stream, so the next frame starts at once on the main stream
            const SgiHandle_::Peer& root = h->peer[SGI_PEER_ROOT];
            const size_t rowBytes = (size_t)h->cp.width * 8;
            const char* src = (const char*)io.combinedOut + (size_t)(h->band.compute_y0 - h->cp.y_offset) * rowBytes;
            char* dst = root.base + root.info.off_image[(int)(h->frameNo & 1)] + (size_t)h->band.compute_y0 * rowBytes;
            SGI_CUDA(h, cudaEventRecord(h->evFinal, h->stream));
            SGI_CUDA(h, cudaStreamWaitEvent(h->copyStream, h->evFinal, 0));
            SGI_CUDA(h, cudaMemcpyAsync(dst, src, (size_t)h->band.compute_rows * rowBytes, cudaMemcpyDeviceToDevice, h->copyStream));
            n += sgi::launchSignal(h->copyStream, nullptr, nullptr, 0, (int32_t*)(root.base + root.info.off_flags) + 8 + h->band.rank, (int)(uint32_t)h->frameNo);
            SGI_CUDA(h, cudaEventRecord(h->evCopyDone, h->copyStream));
            h->copyPending = true;
        }
        // the owner waits for the PREVIOUS frame's rows here (they travelled during this frame); sgi_frame_image waits for
        // the current frame's
        if (final_ && h->band.owns_frame_image && h->frameNo > 1 && h->imageWaitedFrame < h->frameNo - 1) {
            n += sgi::waitAllBands(h, h->frameNo - 1);
            h->imageWaitedFrame = h->frameNo - 1;
        }
    }
    if (final_) { h->cur ^= 1; h->guideValid = false; h->temporalDone = false; h->temporalForked = false; h->gatherDoneThisFrame = false; }   // swap with the previous-frame FBO (BE:1949-1951)
    return postLaunch(h, n);
}

int sgi_denoise(SgiHandle h, const SgiFrameParams* fp) {
    if (!h || !fp) return SGI_ERR_INVALID_ARG;
    for (int pass = 0; pass < h->cp.denoise_iters; ++pass) {
        const int rc = sgi_denoise_pass(h, fp, pass);
        if (rc != SGI_OK) return rc;
    }
    return SGI_OK;
}

int sgi_frame(SgiHandle h, const SgiFrameParams* fp) {
    if (!h || !fp) return SGI_ERR_INVALID_ARG;
    int rc;
    if (fp->totalNumLights == 0) return SGI_OK;      // `if (perVPLDistToViewer.size() == 0) return;` (BE:1582, 1675, 1817)
    if (h->earlyPending) {
        // the cull (and, if the caller said the atlases were final, the colouring) of this frame already ran on the early stream
        // (sgi_cull_early): the main stream only waits for it.  The frame must be the one the early cull was given.
        const SgiFrameParams& e = h->earlyParams;
        if (e.totalNumLights != fp->totalNumLights || std::memcmp(e.cascadeProjViews, fp->cascadeProjViews, sizeof(e.cascadeProjViews)) != 0 ||
            std::memcmp(e.cascadePlanes, fp->cascadePlanes, sizeof(e.cascadePlanes)) != 0 ||
            std::memcmp(e.infiniteLightDirection, fp->infiniteLightDirection, sizeof(e.infiniteLightDirection)) != 0 ||
            std::memcmp(e.infiniteLightColor, fp->infiniteLightColor, sizeof(e.infiniteLightColor)) != 0)
            return fail(h, SGI_ERR_INVALID_ARG, "sgi_frame: the pending sgi_cull_early ran with different lights / sun / cascades than this frame");
        SGI_CUDA(h, cudaSetDevice(h->cp.device));
        SGI_CUDA(h, cudaStreamWaitEvent(h->stream, h->evEarly, 0));
        h->earlyPending = false;
        if (!h->earlyColored && (rc = sgi_color(h, fp)) != SGI_OK) return rc;
    } else if (h->bound && h->buf.atlasEntryByVpl && h->buf.shadowDiffuseIndices && h->buf.diffuseCubeMaps[0] && h->buf.shadowCubeMaps[0]) {
        if ((rc = cullImpl(h, fp, true)) != SGI_OK) return rc;
    } else {
        if ((rc = sgi_cull(h, fp)) != SGI_OK) return rc;
        if ((rc = sgi_color(h, fp)) != SGI_OK) return rc;
    }
    const bool lists = h->buf.vplStage1Results && h->buf.vplVisiblePerTile;
    // The stochastic gather (the reference's live shader) and the denoiser do not read the per-tile lists, so the two tile
    // stages run on a low-priority side stream next to them: forked here, joined at the end of the frame.  Their long,
    // thinly populated tail (a few loose patches) then overlaps with the gather instead of idling the SMs.  The tile-list
    // gather consumes the lists, so that mode stays serial.
    const bool fork = lists && fp->gatherMode == SGI_GATHER_STOCHASTIC;
    if (lists && !fork) { if ((rc = sgi_tile_lists(h, fp)) != SGI_OK) return rc; }
    if (fork) {
        SGI_CUDA(h, cudaSetDevice(h->cp.device));
        SGI_CUDA(h, cudaEventRecord(h->evFork, h->stream));
        SGI_CUDA(h, cudaStreamWaitEvent(h->tileStream, h->evFork, 0));
        rc = tileListsOn(h, fp, h->tileStream);
        // whatever the side stream holds by now is joined before this call returns, on the error paths too: its kernels read
        // vplUpdatedData / the visible list, which the caller may rebind or the next cull overwrites
        SGI_CUDA(h, cudaEventRecord(h->evJoin, h->tileStream));
        if (rc != SGI_OK) { cudaStreamWaitEvent(h->stream, h->evJoin, 0); return rc; }
    }
    rc = sgi_gather(h, fp);
    // The temporal pre-pass (reprojection, history decision) streams 56 bytes per pixel and needs nothing of this frame's gather
    // or filter: it runs on its own stream under the two non-final à-trous passes, which are bound by MUFU / issue slots and leave
    // the DRAM idle; the final pass waits for it.  (Not under the gather: its 512-thread, 98 KB CTAs cannot become resident while a
    // stream of small CTAs keeps refilling the SMs — measured, profiles/r02d.)
    if (rc == SGI_OK && h->cp.denoise_iters > 2 && h->buf.velocity && h->buf.id && h->buf.prevNormal && h->buf.prevId && h->buf.prevDepth && !h->temporalDone) {
        SGI_CUDA(h, cudaStreamWaitEvent(h->auxStream, h->evTablesFree, 0));     // recorded by sgi_gather after its launch
        if ((rc = temporalOn(h, fp, h->auxStream)) == SGI_OK) {
            SGI_CUDA(h, cudaEventRecord(h->evTemporal, h->auxStream));
            h->temporalForked = true;
        }
    }
    if (rc == SGI_OK) rc = sgi_denoise(h, fp);
    if (rc != SGI_OK && h->temporalForked) {
        // the final pass, which joins the pre-pass, was not reached: join it here, it reads the caller's G-buffer planes
        cudaStreamWaitEvent(h->stream, h->evTemporal, 0);
        h->temporalForked = false; h->temporalDone = false;
    }
    if (fork) {
        const std::string keep = h->lastError;
        const cudaError_t e = cudaStreamWaitEvent(h->stream, h->evJoin, 0);
        if (rc == SGI_OK) SGI_CUDA(h, e); else h->lastError = keep;
    }
    return rc;
}

int sgi_visible_count(SgiHandle h, int32_t* out_count) {
    if (!h || !out_count) return SGI_ERR_INVALID_ARG;
    NEED(h, vplVisibleIndices);
    SGI_CUDA(h, cudaSetDevice(h->cp.device));
    SGI_CUDA(h, cudaMemcpyAsync(h->pinnedCount, h->buf.vplVisibleIndices, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    SGI_CUDA(h, cudaStreamSynchronize(h->stream));
    *out_count = *h->pinnedCount;
    return SGI_OK;
}

int sgi_get_plane(SgiHandle h, int32_t which, void** out_ptr, int64_t* out_bytes) {
    if (!h || !out_ptr) return SGI_ERR_INVALID_ARG;
    const int64_t n = (int64_t)pixels(h);
    void* p = nullptr; int64_t bytes = n * 8;
    const int prev = h->cur ^ 1;    // after a completed frame `cur` was swapped, so "this frame's" output sits in prev
    switch (which) {
        case SGI_PLANE_GI_COLOR: p = h->giColor; break;
        case SGI_PLANE_GI_RESERVOIR: p = h->giReservoir; break;
        case SGI_PLANE_DENOISE_RES_A: p = h->resA; break;
        case SGI_PLANE_DENOISE_RES_B: p = h->resB; break;
        case SGI_PLANE_GI_ILLUM: case SGI_PLANE_PREV_GI_ILLUM: p = h->giIllum[prev]; break;
        case SGI_PLANE_HISTORY: case SGI_PLANE_PREV_HISTORY: p = h->history[prev]; bytes = n * 2; break;
        case SGI_PLANE_COMBINED: p = h->buf.lightingOut ? h->buf.lightingOut : h->combined; break;
        default: return fail(h, SGI_ERR_INVALID_ARG, "sgi_get_plane: unknown plane");
    }
    *out_ptr = p;
    if (out_bytes) *out_bytes = bytes;
    return SGI_OK;
}

int sgi_download_plane(SgiHandle h, int32_t which, void* host_dst, int64_t bytes) {
    if (!h || !host_dst) return SGI_ERR_INVALID_ARG;
    void* p = nullptr; int64_t n = 0;
    const int rc = sgi_get_plane(h, which, &p, &n);
    if (rc != SGI_OK) return rc;
    if (bytes < 0 || bytes > n) return fail(h, SGI_ERR_INVALID_ARG, "sgi_download_plane: negative size or more bytes requested than the plane holds");
    SGI_CUDA(h, cudaSetDevice(h->cp.device));
    SGI_CUDA(h, cudaMemcpyAsync(host_dst, p, (size_t)bytes, cudaMemcpyDeviceToHost, h->stream));
    SGI_CUDA(h, cudaStreamSynchronize(h->stream));
    return SGI_OK;
}

int sgi_reset_history(SgiHandle h) {
    if (!h) return SGI_ERR_INVALID_ARG;
    SGI_CUDA(h, cudaSetDevice(h->cp.device));
    const size_t n = pixels(h);
    for (int i = 0; i < 2; ++i) {
        SGI_CUDA(h, cudaMemsetAsync(h->giIllum[i], 0, n * 8, h->stream));
        SGI_CUDA(h, cudaMemsetAsync(h->history[i], 0, n * 2, h->stream));
    }
    return SGI_OK;
}

int sgi_synchronize(SgiHandle h) {
    if (!h) return SGI_ERR_INVALID_ARG;
    SGI_CUDA(h, cudaSetDevice(h->cp.device));
    SGI_CUDA(h, cudaStreamSynchronize(h->stream));
    SGI_CUDA(h, cudaStreamSynchronize(h->tileStream));
    SGI_CUDA(h, cudaStreamSynchronize(h->copyStream));
    SGI_CUDA(h, cudaStreamSynchronize(h->auxStream));
    return SGI_OK;
}

void* sgi_stream(SgiHandle h) { return h ? (void*)h->stream : nullptr; }
int64_t sgi_launch_count(SgiHandle h) { return h ? h->launches : 0; }
const char* sgi_last_error(SgiHandle h) { return h ? h->lastError.c_str() : "null handle"; }

}  // extern "C"
