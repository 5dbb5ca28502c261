// sgi_band.cu — row bands across GPUs with peer-to-peer halos over NVLink.
//
// The reference is a single-GPU engine (StratusRendererBackend.cpp has one GL context), so nothing here replaces
// reference code; it is the multi-GPU plan of BASELINE.json's north_star: bands of rows, denoiser halos exchanged
// between neighbouring bands, bands gathered into the final image.
//
// Mechanism.  Each band handle keeps the planes other bands write into in one cudaMalloc slab.  Neighbours map that
// slab (CUDA IPC across processes, peer access inside one process) and the band's own kernels store their boundary
// rows into it while they compute (k_gather.cu / k_denoise.cu: pushRes, pushGiHistory, the frame-image store) —
// the transfer is part of the producing kernel, not a separate copy or collective.  Ordering is by epoch flags in
// the consumer's slab: after a pass the producer's stream runs a one-thread kernel that fences system-wide and
// stores the epoch into both neighbours' flag words; before the consuming pass the consumer's stream runs a
// one-thread kernel that spins (acquire loads, nanosleep) until both flags reached the epoch.  A spin that lasts
// longer than ~2 s gives up and bumps a time-out counter (sgi_band_timeouts) instead of hanging the GPU.
//
// Epoch = frameNo*32 + stage, stage k = "output of denoise pass k-1 is in your halo" (stage 1 comes from the gather,
// which applies pass 0 itself); stage denoise_iters = the band's frame is complete.  Write-after-read hazards are
// covered by the same flags: a band's pass p+2 overwrites the halo rows its neighbour's pass p+1 read, and it has
// waited for the neighbour's pass p+1 flag by then; the next frame's gather waits for the neighbours' frame-complete flag.
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include "sgi_internal.h"

namespace sgi {
namespace {

__global__ void waitFlagsKernel(const int32_t* f0, const int32_t* f1, int epoch, int32_t* timeouts) {
    const long long t0 = clock64();
    bool ok = true;
    if (f0 != nullptr) ok = spinUntil(f0, epoch, t0) && ok;
    if (f1 != nullptr) ok = spinUntil(f1, epoch, t0) && ok;
    if (!ok) atomicAdd(timeouts, 1);
}

__global__ void waitAllBandsKernel(const int32_t* flags, int world, int epoch, int32_t* timeouts) {
    const long long t0 = clock64();
    bool ok = true;
    for (int r = threadIdx.x; r < world; r += blockDim.x) ok = spinUntil(flags + r, epoch, t0) && ok;
    if (!ok) atomicAdd(timeouts, 1);
}

__global__ void signalKernel(int32_t* a, int32_t* b, int epoch, int32_t* c, int valueC) {
    __threadfence_system();          // everything this stream's earlier kernels stored (peer halo rows, image rows) is visible first
    if (a != nullptr) stReleaseSys(a, epoch);
    if (b != nullptr) stReleaseSys(b, epoch);
    if (c != nullptr) stReleaseSys(c, valueC);
}

}  // namespace

// CUDA loads kernels lazily on first launch, and that load can stall behind a kernel that is spinning on a band flag
// (sgi_band.cu) — so every kernel of this file is loaded when the handle is created.
bool preloadBandKernels() {
    cudaFuncAttributes a; bool ok = true;
    ok = ok && cudaFuncGetAttributes(&a, waitFlagsKernel) == cudaSuccess;
    ok = ok && cudaFuncGetAttributes(&a, waitAllBandsKernel) == cudaSuccess;
    ok = ok && cudaFuncGetAttributes(&a, signalKernel) == cudaSuccess;
    return ok;
}

int launchWaitFlags(cudaStream_t stream, const int32_t* f0, const int32_t* f1, int epoch, int32_t* timeouts) {
    if (f0 == nullptr && f1 == nullptr) return 0;
    waitFlagsKernel<<<1, 1, 0, stream>>>(f0, f1, epoch, timeouts);
    return 1;
}
int launchWaitAllBands(cudaStream_t stream, const int32_t* flags, int world, int epoch, int32_t* timeouts) {
    waitAllBandsKernel<<<1, 32, 0, stream>>>(flags, world, epoch, timeouts);
    return 1;
}
int launchSignal(cudaStream_t stream, int32_t* a, int32_t* b, int epoch, int32_t* c, int valueC) {
    if (a == nullptr && b == nullptr && c == nullptr) return 0;
    signalKernel<<<1, 1, 0, stream>>>(a, b, epoch, c, valueC);
    return 1;
}

BandLinks noLinks() {
    BandLinks L;
    std::memset(&L, 0, sizeof(L));
    return L;
}

static int32_t* peerFlags(SgiHandle_* h, int which) {
    SgiHandle_::Peer& p = h->peer[which];
    return p.attached ? (int32_t*)(p.base + p.info.off_flags) : nullptr;
}

// Where the pass that writes reservoir `resIndex` also stores its boundary rows.  `cur` parity is the same on every band
// (all bands run the same number of final passes), so the neighbour's GI/history of this frame is its plane [cur].
BandLinks bandLinks(SgiHandle_* h, int resIndex, bool final_, int waitEpoch) {
    BandLinks L = noLinks();
    const int myC0 = h->band.compute_y0, myC1 = h->band.compute_y0 + h->band.compute_rows;
    for (int which = 0; which < 2; ++which) {
        SgiHandle_::Peer& p = h->peer[which];
        if (!p.attached) continue;
        PushTarget& t = which == SGI_PEER_UP ? L.up : L.down;
        const SgiPeerInfo& pi = p.info;
        // SGI_HALO_RECOMPUTE: no reservoir rows travel (only the final pass has links at all, for its GI rows)
        t.res = h->band.halo_mode == SGI_HALO_RECOMPUTE ? nullptr : (uint2*)(p.base + pi.off_res[resIndex]);
        if (final_) t.gi = (uint2*)(p.base + pi.off_gi[h->cur]);
        t.yOffset = pi.y_offset;
        // the neighbour's halo on my side: its held rows minus its computed rows, intersected with the rows I compute
        const int lo = which == SGI_PEER_UP ? pi.compute_y0 + pi.compute_rows : pi.y_offset;
        const int hi = which == SGI_PEER_UP ? pi.y_offset + pi.rows : pi.compute_y0;
        t.firstRow = lo > myC0 ? lo : myC0;
        t.lastRow = hi < myC1 ? hi : myC1;
        if (t.firstRow >= t.lastRow) { t.res = nullptr; t.gi = nullptr; }
    }
    if (waitEpoch != 0) {          // my flags[0] is written by the band above, flags[1] by the band below
        L.waitUp = h->peer[SGI_PEER_UP].attached ? h->flags + 0 : nullptr;
        L.waitDown = h->peer[SGI_PEER_DOWN].attached ? h->flags + 1 : nullptr;
        L.waitEpoch = waitEpoch;
        L.timeouts = h->flags + 2;
    }
    if (final_) {
        L.reprojMisses = h->flags + 3;
        const int parity = (int)(h->frameNo & 1);
        if (h->band.owns_frame_image) L.image = (uint2*)h->frameImage[parity];      // the owner's own rows; the other bands' rows arrive by copy engine (sgi_abi.cu)
    }
    return L;
}

int waitNeighbours(SgiHandle_* h, int epoch) {
    // my flags[0] is written by the band above, flags[1] by the band below
    const int32_t* f0 = h->peer[SGI_PEER_UP].attached ? h->flags + 0 : nullptr;
    const int32_t* f1 = h->peer[SGI_PEER_DOWN].attached ? h->flags + 1 : nullptr;
    return launchWaitFlags(h->stream, f0, f1, epoch, h->flags + 2);
}

int signalNeighbours(SgiHandle_* h, int epoch, bool final_) {
    int32_t* up = peerFlags(h, SGI_PEER_UP);        // I am the band BELOW my upper neighbour: its flags[1]
    int32_t* down = peerFlags(h, SGI_PEER_DOWN);    // and the band ABOVE my lower neighbour: its flags[0]
    int32_t* a = up ? up + 1 : nullptr;
    int32_t* b = down ? down + 0 : nullptr;
    int32_t* c = nullptr;
    if (final_) {
        if (h->band.owns_frame_image) c = h->flags + 8 + h->band.rank;
    }
    return launchSignal(h->stream, a, b, epoch, c, (int)(uint32_t)h->frameNo);
}

void armInKernelSignal(SgiHandle_* h, BandLinks* L, int epoch) {
    int32_t* up = peerFlags(h, SGI_PEER_UP);
    int32_t* down = peerFlags(h, SGI_PEER_DOWN);
    if (h->bandDone == nullptr || (up == nullptr && down == nullptr)) return;
    L->doneCounter = h->bandDone;
    L->sigUp = up ? up + 1 : nullptr;
    L->sigDown = down ? down + 0 : nullptr;
    L->sigEpoch = epoch;
}

int signalOwnerRowsDone(SgiHandle_* h) {
    if (!h->band.owns_frame_image) return 0;
    return launchSignal(h->stream, nullptr, nullptr, 0, h->flags + 8 + h->band.rank, (int)(uint32_t)h->frameNo);
}

int waitAllBands(SgiHandle_* h, int64_t frameNo) {
    return launchWaitAllBands(h->stream, h->flags + 8, h->band.world, (int)(uint32_t)frameNo, h->flags + 2);
}

void detachPeers(SgiHandle_* h) {
    for (int which = 0; which < 3; ++which) {
        SgiHandle_::Peer& p = h->peer[which];
        if (p.attached && p.ipc && p.base) {
            bool shared = false;                  // UP/DOWN and ROOT may be the same mapping
            for (int o = 0; o < which; ++o) if (h->peer[o].attached && h->peer[o].base == p.base) shared = true;
            if (!shared) cudaIpcCloseMemHandle(p.base);
        }
        p = SgiHandle_::Peer();
    }
}

}  // namespace sgi

using namespace sgi;

#define SGI_CUDA_B(h, expr)                                                                        \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            (h)->lastError = std::string(#expr) + " failed: " + cudaGetErrorString(_e);           \
            return SGI_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

int sgiAllocPlanesForBand(SgiHandle h);     // sgi_abi.cu

extern "C" {

int sgi_band_setup(SgiHandle h, const SgiBandParams* bp) {
    if (!h || !bp) return SGI_ERR_INVALID_ARG;
    const int y0 = h->cp.y_offset, y1 = h->cp.y_offset + h->cp.rows;
    if (bp->world < 1 || bp->world > kBandFlags - 8 || bp->rank < 0 || bp->rank >= bp->world || bp->compute_rows <= 0 ||
        bp->compute_y0 < y0 || bp->compute_y0 + bp->compute_rows > y1) {
        h->lastError = "sgi_band_setup: computed rows must lie inside the held rows, 0 <= rank < world <= 56";
        return SGI_ERR_INVALID_ARG;
    }
    if (bp->compute_y0 % SGI_TILE_DIVISOR || bp->compute_rows % SGI_TILE_DIVISOR || y0 % SGI_TILE_DIVISOR) {
        h->lastError = "sgi_band_setup: y_offset, compute_y0 and compute_rows must be multiples of 5 (tiles never straddle bands)";
        return SGI_ERR_INVALID_ARG;
    }
    {
        // an inner edge must hold at least the reach of the widest à-trous pass, 2*(multiplier+1) = 2^iters rows (BE:1920): the
        // neighbour fills exactly the held rows beyond the computed ones, and taps beyond them would be clamped silently
        const int iters = h->cp.denoise_iters;
        const int reach = iters > 1 ? 2 * (1 << (iters - 1)) : 0;
        const int top = bp->compute_y0 - y0, bottom = y1 - (bp->compute_y0 + bp->compute_rows);
        if ((bp->rank > 0 && top < reach) || (bp->rank + 1 < bp->world && bottom < reach)) {
            h->lastError = "sgi_band_setup: an inner edge holds fewer halo rows than the à-trous reach 2^denoise_iters (SGI_HALO_ROWS covers 4 iterations)";
            return SGI_ERR_INVALID_ARG;
        }
        if (bp->halo_mode != SGI_HALO_EXCHANGE && bp->halo_mode != SGI_HALO_RECOMPUTE) { h->lastError = "sgi_band_setup: unknown halo_mode"; return SGI_ERR_INVALID_ARG; }
        // SGI_HALO_RECOMPUTE: pass 0 is recomputed on the reach of ALL later passes, sum_{p>=1} 2^(p+1) = 2^(iters+1) - 4 rows
        const int recompute = iters > 1 ? (1 << (iters + 1)) - 4 : 0;
        if (bp->halo_mode == SGI_HALO_RECOMPUTE && ((bp->rank > 0 && top < recompute) || (bp->rank + 1 < bp->world && bottom < recompute))) {
            h->lastError = "sgi_band_setup: SGI_HALO_RECOMPUTE needs 2^(denoise_iters+1) - 4 held rows beyond every inner edge (28 for 4 iterations)";
            return SGI_ERR_INVALID_ARG;
        }
    }
    SGI_CUDA_B(h, cudaSetDevice(h->cp.device));
    SGI_CUDA_B(h, cudaStreamSynchronize(h->stream));
    SGI_CUDA_B(h, cudaStreamSynchronize(h->copyStream));
    detachPeers(h);
    h->banded = true;
    h->band = *bp;
    const int rc = sgiAllocPlanesForBand(h);
    if (rc != SGI_OK) return rc;
    // the slab (flags and halo rows) is zeroed on the stream: that must have happened before a neighbour that attaches after this
    // call returns can store pass-0 rows or raise a flag in it
    SGI_CUDA_B(h, cudaStreamSynchronize(h->stream));
    return SGI_OK;
}

int sgi_band_export(SgiHandle h, SgiPeerInfo* out) {
    if (!h || !out) return SGI_ERR_INVALID_ARG;
    if (!h->banded) { h->lastError = "sgi_band_export: call sgi_band_setup first"; return SGI_ERR_INVALID_ARG; }
    std::memset(out, 0, sizeof(*out));
    SGI_CUDA_B(h, cudaSetDevice(h->cp.device));
    cudaIpcMemHandle_t mh;
    static_assert(sizeof(mh) == sizeof(out->mem_handle), "cudaIpcMemHandle_t is 64 bytes");
    SGI_CUDA_B(h, cudaIpcGetMemHandle(&mh, h->slab));
    std::memcpy(out->mem_handle, &mh, sizeof(mh));
    out->local_base = (uint64_t)(uintptr_t)h->slab;
    for (int i = 0; i < 2; ++i) { out->off_res[i] = h->offRes[i]; out->off_gi[i] = h->offGi[i]; out->off_hist[i] = h->offHist[i]; out->off_image[i] = h->offImage[i]; }
    out->off_flags = h->offFlags;
    out->y_offset = h->cp.y_offset; out->rows = h->cp.rows;
    out->compute_y0 = h->band.compute_y0; out->compute_rows = h->band.compute_rows;
    out->rank = h->band.rank; out->device = h->cp.device; out->pid = (int32_t)getpid();
    return SGI_OK;
}

int sgi_band_attach(SgiHandle h, int32_t which, const SgiPeerInfo* peer) {
    if (!h || !peer || which < 0 || which > 2) return SGI_ERR_INVALID_ARG;
    if (!h->banded) { h->lastError = "sgi_band_attach: call sgi_band_setup first"; return SGI_ERR_INVALID_ARG; }
    if (which == SGI_PEER_UP && peer->rank != h->band.rank - 1) { h->lastError = "sgi_band_attach: SGI_PEER_UP must be band rank-1"; return SGI_ERR_INVALID_ARG; }
    if (which == SGI_PEER_DOWN && peer->rank != h->band.rank + 1) { h->lastError = "sgi_band_attach: SGI_PEER_DOWN must be band rank+1"; return SGI_ERR_INVALID_ARG; }
    if (which != SGI_PEER_ROOT) {
        // the neighbour's halo on my side is filled by MY kernels from the rows I compute: it may not reach beyond them
        const int myC0 = h->band.compute_y0, myC1 = h->band.compute_y0 + h->band.compute_rows;
        const bool ok = which == SGI_PEER_UP ? (peer->compute_y0 + peer->compute_rows == myC0 && peer->y_offset + peer->rows <= myC1)
                                             : (peer->compute_y0 == myC1 && peer->y_offset >= myC0);
        if (!ok) { h->lastError = "sgi_band_attach: the neighbour's computed rows must adjoin mine and its halo may not be taller than my band"; return SGI_ERR_INVALID_ARG; }
    }
    SGI_CUDA_B(h, cudaSetDevice(h->cp.device));
    SgiHandle_::Peer& p = h->peer[which];
    if (p.attached) { h->lastError = "sgi_band_attach: peer already attached"; return SGI_ERR_INVALID_ARG; }
    char* base = nullptr; bool ipc = false;
    for (int o = 0; o < 3; ++o)      // the root may also be a neighbour: reuse its mapping
        if (h->peer[o].attached && h->peer[o].info.pid == peer->pid && h->peer[o].info.local_base == peer->local_base) { base = h->peer[o].base; ipc = h->peer[o].ipc; }
    if (base == nullptr) {
        if (peer->pid == (int32_t)getpid()) {
            if (peer->device != h->cp.device) {
                int can = 0;
                SGI_CUDA_B(h, cudaDeviceCanAccessPeer(&can, h->cp.device, peer->device));
                if (!can) { h->lastError = "sgi_band_attach: no peer access between the two devices"; return SGI_ERR_UNSUPPORTED; }
                const cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { h->lastError = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); return SGI_ERR_CUDA; }
                cudaGetLastError();
            }
            base = (char*)(uintptr_t)peer->local_base;
        } else {
            cudaIpcMemHandle_t mh;
            std::memcpy(&mh, peer->mem_handle, sizeof(mh));
            void* mapped = nullptr;
            SGI_CUDA_B(h, cudaIpcOpenMemHandle(&mapped, mh, cudaIpcMemLazyEnablePeerAccess));
            base = (char*)mapped; ipc = true;
        }
    }
    p.attached = true; p.ipc = ipc; p.base = base; p.info = *peer;
    return SGI_OK;
}

int sgi_frame_image(SgiHandle h, void** out_ptr, int64_t* out_bytes) {
    if (!h || !out_ptr) return SGI_ERR_INVALID_ARG;
    if (!h->banded || !h->band.owns_frame_image) { h->lastError = "sgi_frame_image: this handle does not own the frame image"; return SGI_ERR_INVALID_ARG; }
    // the owner's stream waits here for every band's rows of the last frame started on this handle (the wait at the end of
    // a frame is for the frame before it, so that the transfers overlap the next frame)
    if (h->frameNo > 0 && h->imageWaitedFrame < h->frameNo) {
        SGI_CUDA_B(h, cudaSetDevice(h->cp.device));
        h->launches += sgi::waitAllBands(h, h->frameNo);
        h->imageWaitedFrame = h->frameNo;
        SGI_CUDA_B(h, cudaGetLastError());
    }
    *out_ptr = h->frameImage[(int)(h->frameNo & 1)];
    if (out_bytes) *out_bytes = (int64_t)h->cp.width * h->cp.full_height * 8;
    return SGI_OK;
}

int sgi_band_timeouts(SgiHandle h, int32_t* out_count) {
    if (!h || !out_count) return SGI_ERR_INVALID_ARG;
    *out_count = 0;
    if (!h->banded) return SGI_OK;
    SGI_CUDA_B(h, cudaSetDevice(h->cp.device));
    SGI_CUDA_B(h, cudaMemcpyAsync(h->pinnedCount, h->flags + 2, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    SGI_CUDA_B(h, cudaStreamSynchronize(h->stream));
    *out_count = *h->pinnedCount;
    return SGI_OK;
}

int sgi_band_stats(SgiHandle h, int32_t out[8]) {
    if (!h || !out) return SGI_ERR_INVALID_ARG;
    for (int i = 0; i < 8; ++i) out[i] = 0;
    if (!h->banded) return SGI_OK;
    SGI_CUDA_B(h, cudaSetDevice(h->cp.device));
    SGI_CUDA_B(h, cudaStreamSynchronize(h->stream));
    SGI_CUDA_B(h, cudaMemcpy(out, h->flags, 8 * sizeof(int32_t), cudaMemcpyDeviceToHost));
    SGI_CUDA_B(h, cudaMemset(h->flags + 5, 0, 3 * sizeof(int32_t)));       // the spin account restarts
    return SGI_OK;
}

int sgi_band_reprojection_misses(SgiHandle h, int32_t* out_count) {
    if (!h || !out_count) return SGI_ERR_INVALID_ARG;
    *out_count = 0;
    if (!h->banded) return SGI_OK;
    SGI_CUDA_B(h, cudaSetDevice(h->cp.device));
    SGI_CUDA_B(h, cudaMemcpyAsync(h->pinnedCount, h->flags + 3, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    SGI_CUDA_B(h, cudaStreamSynchronize(h->stream));
    *out_count = *h->pinnedCount;
    return SGI_OK;
}

}  // extern "C"
