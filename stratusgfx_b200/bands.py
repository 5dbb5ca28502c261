"""Row-band split of one frame across the GPUs of a box — host-side plumbing.

One process per GPU (torch.distributed; NCCL on the GPU box, gloo in the CPU tests).  GPU g owns rows
[g*H/G, (g+1)*H/G) of every per-pixel plane.  Per frame:

  1. rank 0 broadcasts what the CPU side of the renderer produced for this frame: the VPL table,
     the per-VPL atlas entries and the <= 5 cube shadow maps re-rendered this frame
     (StratusRendererBackend.h:378) — < 0.5 MB in total;
  2. every rank runs the whole pass on its band.  Two ways to feed the à-trous passes at the band edges:
       peer-to-peer (default): the band holds halo_rows() extra rows per inner edge; the neighbours' kernels store
         their boundary rows of every pass straight into them over NVLink (csrc/sgi_band.cu), ordered by epoch flags;
         the final pass also stores the lit rows into rank 0's frame image, so there is no separate gather;
       recompute (SGI_HALO_RECOMPUTE): the band holds recompute_rows() extra rows per inner edge and computes the gather and the
         non-final passes on as many of them as its later passes read (28 / 24 / 16 rows): same bytes, no exchange inside the
         frame; only the final pass's GI rows (next frame's reprojection) and the lit rows travel;
       apron (comparison): the band holds APRON = 30 extra rows and recomputes ALL passes on them redundantly (reach 4 + 8 + 16 = 28),
         no exchange, then
  3. (apron mode only) the lit bands are gathered to rank 0 with NCCL.

Band heights and aprons are multiples of the 5-pixel tile so per-tile lists never straddle a band.
"""
import ctypes as C
from dataclasses import dataclass

APRON = 30   # >= 28 rows of à-trous reach, rounded up to the 5-pixel tile grid (redundant-compute mode)
HALO = 20    # >= 16 rows = reach of one à-trous pass (peer-to-peer mode, SGI_HALO_ROWS)
TILE = 5


@dataclass(frozen=True)
class Band:
    rank: int
    b0: int      # first owned row
    b1: int      # one past the last owned row
    y0: int      # first row held (owned rows + apron, clipped to the frame)
    y1: int

    @property
    def rows(self):
        return self.y1 - self.y0

    @property
    def owned(self):
        """slice of the held rows that this rank owns"""
        return slice(self.b0 - self.y0, self.b1 - self.y0)


def halo_rows(denoise_iters):
    """Rows a band must hold beyond its own for ONE à-trous pass of the widest spacing: 2*(multiplier+1) with
    multiplier = 2^i - 1 (StratusRendererBackend.cpp:1920), rounded up to the tile grid.  4 iterations -> 16 -> 20."""
    reach = 2 * (1 << max(denoise_iters - 1, 0)) if denoise_iters > 1 else 0
    return (reach + TILE - 1) // TILE * TILE


def recompute_rows(denoise_iters):
    """Rows a band must hold beyond its own in SGI_HALO_RECOMPUTE mode: the summed reach of all à-trous passes,
    2^(iters+1) - 4 (4 + 8 + 16 = 28 for 4 iterations), rounded up to the tile grid -> 30."""
    reach = (1 << (denoise_iters + 1)) - 4 if denoise_iters > 1 else 0
    return (reach + TILE - 1) // TILE * TILE


def plan_bands(height, world, apron=APRON):
    """Bands for `world` GPUs.  `apron` = extra rows held on each inner edge: APRON for the redundant-compute mode
    (every à-trous pass recomputed locally), halo_rows(iters) for the peer-to-peer mode (neighbours fill them)."""
    if height % (TILE * world):
        raise ValueError(f"height {height} does not split into {world} bands of whole {TILE}-pixel tiles")
    if apron % TILE:
        raise ValueError("apron must be a multiple of the tile size")
    h = height // world
    if world > 1 and apron > h:
        raise ValueError(f"bands of {h} rows are thinner than the {apron}-row halo")
    out = []
    for r in range(world):
        b0, b1 = r * h, (r + 1) * h
        y0, y1 = (max(0, b0 - apron), min(height, b1 + apron)) if world > 1 else (0, height)
        out.append(Band(r, b0, b1, y0, y1))
    return out


def bands_from_cuts(cuts, height, apron):
    """Bands owning rows [cuts[r], cuts[r+1]) plus `apron` held rows on each inner edge."""
    world = len(cuts) - 1
    if cuts[0] != 0 or cuts[-1] != height or any(c % TILE for c in cuts) or any(b <= a for a, b in zip(cuts, cuts[1:])):
        raise ValueError(f"bad band cuts {cuts}")
    if world > 1 and min(b - a for a, b in zip(cuts, cuts[1:])) < apron:
        raise ValueError("a band is thinner than the halo")
    out = []
    for r in range(world):
        b0, b1 = cuts[r], cuts[r + 1]
        y0, y1 = (max(0, b0 - apron), min(height, b1 + apron)) if world > 1 else (0, height)
        out.append(Band(r, b0, b1, y0, y1))
    return out


def balance_cuts(bands, times, height, min_rows):
    """Load balancing from measurements: band r took times[r] for its rows, so its rows cost times[r]/rows each
    (piecewise-constant density); move the cuts so every band gets the same share of the total, on the tile grid,
    never thinner than `min_rows`.  A VPL gather is content dependent (1-6 samples per pixel, re-draws, shadow tests),
    so equal-height bands are not equal work."""
    world = len(bands)
    if world == 1:
        return [0, height]
    min_rows = max(TILE, (min_rows + TILE - 1) // TILE * TILE)
    dens = []
    for b, t in zip(bands, times):
        dens += [float(t) / (b.b1 - b.b0)] * (b.b1 - b.b0)
    cum = [0.0]
    for d in dens:
        cum.append(cum[-1] + d)
    total = cum[-1]
    cuts = [0]
    y = 0
    for r in range(1, world):
        target = total * r / world
        while y < height and cum[y] < target:
            y += 1
        c = int(round(y / TILE)) * TILE
        c = max(c, cuts[-1] + min_rows)                     # not thinner than the halo
        c = min(c, height - (world - r) * min_rows)         # and leave room for the bands below
        cuts.append(c)
    cuts.append(height)
    return cuts


def connect_bands(dist, gp, band, world, root=0, halo_mode=0):
    """Peer-to-peer mode: band_setup on this rank's pass, all-gather the peer descriptors (plain bytes), attach the
    neighbours and the frame-image owner.  Works with NCCL (GPU box) and gloo (tests) alike: only host bytes travel."""
    import torch
    from ._abi import PEER_DOWN, PEER_ROOT, PEER_UP, PeerInfo
    gp.band_setup(band.b0, band.b1 - band.b0, band.rank, world, owns_frame_image=(band.rank == root), halo_mode=halo_mode)
    mine = gp.band_export()
    raw = bytes(mine)
    infos = [None] * world
    if world > 1:
        dist.all_gather_object(infos, raw)
    else:
        infos[0] = raw
    infos = [PeerInfo.from_buffer_copy(b) for b in infos]
    if band.rank > 0:
        gp.band_attach(PEER_UP, infos[band.rank - 1])
    if band.rank + 1 < world:
        gp.band_attach(PEER_DOWN, infos[band.rank + 1])
    if band.rank != root:
        gp.band_attach(PEER_ROOT, infos[root])
    return infos


def broadcast_frame_inputs(dist, tensors, src=0):
    """Step 1: NCCL/gloo broadcast of the small per-frame tables (list of tensors, in place)."""
    for t in tensors:
        dist.broadcast(t, src)


def gather_bands(dist, band_tensor, full_tensor, bands, rank, dst=0):
    """Step 3: gather each rank's owned rows into `full_tensor` ([H, ...]) on rank `dst`.  Bands of equal height use one
    collective gather; balanced (unequal) bands are received row range by row range (point-to-point)."""
    if len({b.b1 - b.b0 for b in bands}) == 1:
        if rank == dst:
            views = [full_tensor[b.b0:b.b1] for b in bands]
            dist.gather(band_tensor, views, dst=dst)
        else:
            dist.gather(band_tensor, None, dst=dst)
        return
    if rank == dst:
        me = bands[dst]
        full_tensor[me.b0:me.b1].copy_(band_tensor)
        reqs = [dist.irecv(full_tensor[b.b0:b.b1], src=b.rank) for b in bands if b.rank != dst]
        for r in reqs:
            r.wait()
    else:
        dist.send(band_tensor.contiguous(), dst)
