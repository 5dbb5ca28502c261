// bands_on_host.cpp — one band of a multi-GPU run, executed on the host: every band is a PROCESS (as on the GPU box), the slabs that
// neighbours store into are POSIX shared memory instead of CUDA IPC mappings, and the band's kernels run in lockstep mode.  The
// translated csrc/sgi_band.cu supplies the real bandLinks / armInKernelSignal / signalNeighbours; this file builds the SgiHandle_
// they read and drives sgi_gather + sgi_denoise_pass in the order of csrc/sgi_abi.cu (band branch).  What it checks: that the
// in-kernel waits and signals (producer counting, boundary-first order, epochs) neither dead-lock nor let a pass read a halo row
// before its neighbour stored it — with MIDDLE bands, which a 2-GPU run does not have.
#define SGI_HOST_LOCKSTEP
#include <unistd.h>
#include <vector>
#include "cuda_host_shim.h"
#include "sgi_band.cu"             // tests/cuda_on_host/_build/sgi_band.cu (namespace sgi part only)

int sgiAllocPlanesForBand(SgiHandle) { return SGI_ERR_UNSUPPORTED; }     // (referenced by nothing that runs here)

namespace sgi { void fillFrameConsts(const SgiFrameParams* fp, FrameConsts* fc); }

struct HostPeer { int32_t present, y_offset, rows, compute_y0, compute_rows, rank; char* base; };

namespace {
struct State {
    SgiHandle_ h;
    std::vector<float4> guide, temporal;
    std::vector<uint2> giColor, giReservoir;
    int gatherCounters[2] = {0, 0};
    bool kernelSignal = true;
    bool recompute = false;          // SGI_HALO_RECOMPUTE (include/stratusgi.h)
} *st = nullptr;
// test hooks (host_band_test_hooks): this band sleeps before every final pass; the temporal pre-pass skips its wait for the neighbours'
// previous frame.  Together they show that the wait is what keeps a late neighbour's GI rows out of this band's reprojected reads.
int gDelayFinalSeconds = 0;
bool gDropInterFrameWait = false;

// the slab layout of csrc/sgi_abi.cu allocPlanes for `rowsMax` held rows (the same in every band process of a test)
struct Layout { int64_t res[2], gi[2], hist[2], flags, total; };
Layout layoutOf(int W, int rowsMax) {
    Layout L; int64_t off = 0; const int64_t n = (int64_t)W * rowsMax;
    auto take = [&off](int64_t bytes) { const int64_t o = off; off = (off + bytes + 255) & ~int64_t(255); return o; };
    L.res[0] = take(n * 8); L.res[1] = take(n * 8); L.gi[0] = take(n * 8); L.gi[1] = take(n * 8); L.hist[0] = take(n * 2); L.hist[1] = take(n * 2);
    L.flags = take(sgi::kBandFlags * sizeof(int32_t)); L.total = off;
    return L;
}
sgi::BandDims bandDims(const SgiHandle_& h) {
    sgi::BandDims bd;
    bd.W = h.cp.width; bd.fullH = h.cp.full_height; bd.yOffset = h.cp.y_offset; bd.rows = h.cp.rows; bd.cubeRes = h.cp.cube_res; bd.csmRes = h.cp.csm_res;
    bd.c0 = h.banded ? h.band.compute_y0 - h.cp.y_offset : 0;
    bd.crows = h.banded ? h.band.compute_rows : h.cp.rows;
    return bd;
}
// bandDimsForPass of csrc/sgi_abi.cu: the rows pass `pass` of a band that recomputes its halo covers
sgi::BandDims bandDimsForPass(const SgiHandle_& h, int pass, bool recompute) {
    sgi::BandDims bd = bandDims(h);
    if (!h.banded || !recompute) return bd;
    int ext = 0;
    for (int j = pass + 1; j < h.cp.denoise_iters; ++j) ext += 2 << j;
    const int up = ext < bd.c0 ? ext : bd.c0;
    const int below = bd.rows - (bd.c0 + bd.crows);
    const int down = ext < below ? ext : below;
    bd.c0 -= up; bd.crows += up + down;
    return bd;
}
}  // namespace

extern "C" void host_band_test_hooks(int delayFinalSeconds, int dropInterFrameWait) { gDelayFinalSeconds = delayFinalSeconds; gDropInterFrameWait = dropInterFrameWait != 0; }

extern "C" int64_t host_band_slab_bytes(int W, int rowsMax) { return layoutOf(W, rowsMax).total; }

extern "C" int host_band_init(int W, int fullH, int cubeRes, int csmRes, int iters, int rank, int world, int y_offset, int rows, int compute_y0,
                              int compute_rows, char* slab, int rowsMax, const HostPeer* up, const HostPeer* down, int kernelSignal) {
    delete st; st = new State();
    SgiHandle_& h = st->h;
    h.cp.width = W; h.cp.full_height = fullH; h.cp.y_offset = y_offset; h.cp.rows = rows; h.cp.cube_res = cubeRes; h.cp.csm_res = csmRes;
    h.cp.denoise_iters = iters; h.cp.device = 0; h.cp.stream = nullptr;
    const Layout L = layoutOf(W, rowsMax);
    h.slab = slab; h.slabBytes = (size_t)L.total;
    h.resA = slab + L.res[0]; h.resB = slab + L.res[1];
    for (int i = 0; i < 2; ++i) { h.giIllum[i] = slab + L.gi[i]; h.history[i] = slab + L.hist[i]; h.offRes[i] = L.res[i]; h.offGi[i] = L.gi[i]; h.offHist[i] = L.hist[i]; }
    h.offFlags = L.flags; h.flags = (int32_t*)(slab + L.flags); h.bandDone = h.flags + 4;
    h.banded = world > 1;
    h.band.compute_y0 = compute_y0; h.band.compute_rows = compute_rows; h.band.rank = rank; h.band.world = world; h.band.owns_frame_image = 0;
    h.band.halo_mode = (kernelSignal & 2) ? SGI_HALO_RECOMPUTE : SGI_HALO_EXCHANGE;        // (bit 1 of the mode argument)
    const HostPeer* ps[2] = {up, down};
    for (int which = 0; which < 2; ++which) {
        if (!ps[which] || !ps[which]->present) continue;
        SgiHandle_::Peer& p = h.peer[which];
        p.attached = true; p.ipc = false; p.base = ps[which]->base;
        for (int i = 0; i < 2; ++i) { p.info.off_res[i] = L.res[i]; p.info.off_gi[i] = L.gi[i]; p.info.off_hist[i] = L.hist[i]; p.info.off_image[i] = -1; }
        p.info.off_flags = L.flags;
        p.info.y_offset = ps[which]->y_offset; p.info.rows = ps[which]->rows; p.info.compute_y0 = ps[which]->compute_y0;
        p.info.compute_rows = ps[which]->compute_rows; p.info.rank = ps[which]->rank;
    }
    const size_t n = (size_t)W * rows;
    st->guide.assign(n, make_float4(0.f, 0.f, 0.f, 0.f));
    st->temporal.assign(n, make_float4(0.f, 0.f, 0.f, 0.f));
    st->giColor.assign(n, make_uint2(0u, 0u)); st->giReservoir.assign(n, make_uint2(0u, 0u));
    st->kernelSignal = (kernelSignal & 1) != 0;
    st->recompute = (kernelSignal & 2) != 0;
    return 0;
}

// one frame: sgi_gather (fused pass 0) + sgi_denoise_pass 1 .. iters-1, csrc/sgi_abi.cu
extern "C" int host_band_frame(const SgiFrameParams* fp, const float* depth, const uint32_t* normal, const uint32_t* albedo, const uint16_t* refl,
                               const uint32_t* rma, const unsigned short* ssao, const uint32_t* velocity, const float* ids, const void* screenH4,
                               const uint32_t* prevNormal, const float* prevIds, const float* prevDepth, const SgiVplData* updated,
                               const int32_t* visible, const SgiAtlasEntry* entries, const void* const* shadow, int numAtlases, void* combinedOutH4) {
    using namespace sgi;
    SgiHandle_& h = st->h;
    const int iters = h.cp.denoise_iters;
    FrameConsts fc; fillFrameConsts(fp, &fc);
    const BandDims bd = bandDims(h);
    AtlasTable atl;
    for (int i = 0; i < SGI_MAX_TOTAL_SHADOW_ATLASES; ++i) { atl.shadow[i] = i < numAtlases ? (const unsigned short*)shadow[i] : nullptr; atl.diffuse[i] = nullptr; }
    // ---- sgi_gather ----
    const bool recompute = h.banded && st->recompute;
    BandLinks links = noLinks();
    if (h.banded) {
        ++h.frameNo;
        const int prevDone = h.frameNo > 1 ? bandEpoch(h.frameNo - 1, iters) : 0;
        if (!recompute) {          // (a band that recomputes its halo neither waits nor pushes here)
            links = bandLinks(&h, 0, false, prevDone);
            if (st->kernelSignal) armInKernelSignal(&h, &links, bandEpoch(h.frameNo, 1));
        }
    }
    bool signalled = false;
    std::vector<float> coordU((size_t)bd.W);
    std::vector<float2> coordRow((size_t)bd.rows);
    launchCoordTables(nullptr, bd, coordU.data(), coordRow.data());
    launchGather(nullptr, fc, bandDimsForPass(h, 0, recompute), depth, normal, albedo, refl, rma, ssao, updated, visible, entries, atl, st->giColor.data(), st->giReservoir.data(),
                 st->guide.data(), (uint2*)h.resA, links, nullptr, nullptr, st->gatherCounters, &h.gatherLaunches, &signalled, coordU.data(), coordRow.data());
    if (h.banded && !recompute && !signalled) signalNeighbours(&h, bandEpoch(h.frameNo, 1), false);
    // ---- sgi_denoise_pass 1 .. iters-1 (pass 0 was written by the gather) ----
    for (int pass = 1; pass < iters; ++pass) {
        const bool final_ = pass + 1 == iters;
        DenoiseIO io;
        std::memset(&io, 0, sizeof(io));
        io.screen = (const uint2*)screenH4; io.velocity = velocity; io.normal = normal; io.ids = ids; io.depth = depth;
        io.prevNormal = prevNormal; io.prevIds = prevIds; io.prevDepth = prevDepth;
        io.illum = st->giColor.data();
        void* fbo[2] = {h.resA, h.resB};
        io.shadowsIn = (const uint2*)fbo[(pass - 1) % 2]; io.reservoirOut = (uint2*)fbo[pass % 2];
        io.prevGi = (const uint2*)h.giIllum[h.cur ^ 1]; io.historyIn = (const unsigned short*)h.history[h.cur ^ 1];
        io.numVisible = visible;
        io.combinedOut = (uint2*)combinedOutH4; io.giOut = (uint2*)h.giIllum[h.cur]; io.historyOut = (unsigned short*)h.history[h.cur];
        io.multiplier = (1 << pass) - 1;
        io.guide = st->guide.data();
        if (recompute) {           // only the final pass stores into the neighbours (GI rows) and releases their flags; it waits for nobody
            if (final_) {
                io.links = bandLinks(&h, pass % 2, true, 0);
                if (st->kernelSignal) armInKernelSignal(&h, &io.links, bandEpoch(h.frameNo, pass + 1));
            }
        } else if (h.banded) {
            io.links = bandLinks(&h, pass % 2, final_, bandEpoch(h.frameNo, pass));
            if (st->kernelSignal) armInKernelSignal(&h, &io.links, bandEpoch(h.frameNo, pass + 1));
        }
        if (final_) {
            // the temporal pre-pass (temporalOn, csrc/sgi_abi.cu): after the neighbours have completed the previous frame, whose
            // final passes stored the GI halo rows the reprojected taps may read
            io.temporal = st->temporal.data();
            DenoiseIO tio = io;
            tio.links = noLinks();
            if (h.banded) {
                tio.links.reprojMisses = h.flags + 3;
                if (h.frameNo > 1 && !gDropInterFrameWait)
                    launchWaitFlags(nullptr, h.peer[SGI_PEER_UP].attached ? h.flags + 0 : nullptr, h.peer[SGI_PEER_DOWN].attached ? h.flags + 1 : nullptr,
                                    bandEpoch(h.frameNo - 1, iters), h.flags + 2);
            }
            launchTemporal(nullptr, fc, bd, tio);
        }
        if (final_ && gDelayFinalSeconds > 0) sleep((unsigned)gDelayFinalSeconds);
        signalled = false;
        launchDenoisePass(nullptr, fc, bandDimsForPass(h, pass, recompute), io, false, final_, &signalled);
        if (recompute && !final_) {}
        else if (h.banded && !signalled) signalNeighbours(&h, bandEpoch(h.frameNo, pass + 1), final_);
        if (final_) h.cur ^= 1;
    }
    return h.flags[2];      // time-outs (none can occur here: clock64() is constant on the host, a dead-lock shows as a hang)
}

// Several frames in ONE call (tools/race_check_bands_on_host.py): between two frames the band thread does not return to the Python
// interpreter, whose lock would otherwise order everything one band does in frame f before what another does in frame f + 1 and hide
// a missing inter-frame wait from ThreadSanitizer.  ptrs: nFrames x 16 pointers in host_band_frame's order (depth .. entries, shadow).
extern "C" int host_band_frames(int nFrames, const SgiFrameParams* fps, const void* const* ptrs, int numAtlases, void* const* combinedOut) {
    int rc = 0;
    for (int f = 0; f < nFrames; ++f) {
        const void* const* p = ptrs + (size_t)f * 16;
        rc |= host_band_frame(fps + f, (const float*)p[0], (const uint32_t*)p[1], (const uint32_t*)p[2], (const uint16_t*)p[3], (const uint32_t*)p[4],
                              (const unsigned short*)p[5], (const uint32_t*)p[6], (const float*)p[7], p[8], (const uint32_t*)p[9], (const float*)p[10],
                              (const float*)p[11], (const SgiVplData*)p[12], (const int32_t*)p[13], (const SgiAtlasEntry*)p[14],
                              (const void* const*)p[15], numAtlases, combinedOut[f]);
    }
    return rc;
}
