"""Debug aid: run 2 in-process p2p bands for a few frames and dump the epoch flags of both handles."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from tests.parity_util import SceneData, band_of, make_pass
from stratusgfx_b200.bands import halo_rows, plan_bands
from stratusgfx_b200._abi import PEER_DOWN, PEER_ROOT, PEER_UP


class V:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (int(ptr), False), "version": 3, "strides": None}


d = SceneData("c1", frames=3, yaw_deg_per_frame=0.25)
world = 2
bands = plan_bands(d.H, world, halo_rows(4))
gps = [make_pass(d, b.y0, b.rows) for b in bands]
for gp, b in zip(gps, bands):
    gp.band_setup(b.b0, b.b1 - b.b0, b.rank, world, owns_frame_image=(b.rank == 0))
infos = [gp.band_export() for gp in gps]
for gp, b in zip(gps, bands):
    if b.rank > 0: gp.band_attach(PEER_UP, infos[b.rank - 1])
    if b.rank + 1 < world: gp.band_attach(PEER_DOWN, infos[b.rank + 1])
    if b.rank != 0: gp.band_attach(PEER_ROOT, infos[0])
flags = [torch.as_tensor(V(i.local_base + i.off_flags, 16), device="cuda:0") for i in infos]
for f in range(3):
    for gp, b in zip(gps, bands):
        gp.set_gbuffer(band_of(d.gb[f], b.y0, b.rows), band_of(d.gb[f - 1], b.y0, b.rows)); gp.set_csm(d.csm[f]); gp.bind()
    t = time.time()
    for gp in gps:
        gp.frame(d.fp[f])
    for gp in gps:
        gp.synchronize()
    print("frame", f, "took %.3f s" % (time.time() - t), "flags A", flags[0].cpu().tolist(), "flags B", flags[1].cpu().tolist(), flush=True)
