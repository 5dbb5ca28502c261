"""Row-band split of one frame across the GPUs of a box — host-side plumbing.

One process per GPU (torch.distributed; NCCL on the GPU box, gloo in the CPU tests).  GPU g owns rows
[g*H/G, (g+1)*H/G) of every per-pixel plane.  Per frame:

  1. rank 0 broadcasts what the CPU side of the renderer produced for this frame: the VPL table,
     the per-VPL atlas entries and the <= 5 cube shadow maps re-rendered this frame
     (StratusRendererBackend.h:378) — < 0.5 MB in total;
  2. every rank runs the whole pass on its band plus an apron of APRON rows on each inner edge, so the
     three à-trous passes (reach 4 + 8 + 16 = 28 rows) need no halo exchange; apron rows are
     recomputed redundantly and discarded;
  3. the lit bands are gathered to rank 0.

Band heights and aprons are multiples of the 5-pixel tile so per-tile lists never straddle a band.
"""
from dataclasses import dataclass

APRON = 30   # >= 28 rows of à-trous reach, rounded up to the 5-pixel tile grid
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


def plan_bands(height, world, apron=APRON):
    if height % (TILE * world):
        raise ValueError(f"height {height} does not split into {world} bands of whole {TILE}-pixel tiles")
    if apron % TILE:
        raise ValueError("apron must be a multiple of the tile size")
    h = height // world
    out = []
    for r in range(world):
        b0, b1 = r * h, (r + 1) * h
        y0, y1 = (max(0, b0 - apron), min(height, b1 + apron)) if world > 1 else (0, height)
        out.append(Band(r, b0, b1, y0, y1))
    return out


def broadcast_frame_inputs(dist, tensors, src=0):
    """Step 1: NCCL/gloo broadcast of the small per-frame tables (list of tensors, in place)."""
    for t in tensors:
        dist.broadcast(t, src)


def gather_bands(dist, band_tensor, full_tensor, bands, rank, dst=0):
    """Step 3: gather each rank's owned rows into `full_tensor` ([H, ...]) on rank `dst`."""
    if rank == dst:
        views = [full_tensor[b.b0:b.b1] for b in bands]
        dist.gather(band_tensor, views, dst=dst)
    else:
        dist.gather(band_tensor, None, dst=dst)
