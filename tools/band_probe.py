#!/usr/bin/env python
"""Single-GPU probe of the band-split efficiency: every band of an N-way split of a workload's frame is run as a plain
handle (no neighbours, no pushes) on cuda:0; per-stage CUDA-event times and the back-to-back frame time are printed next to
the whole-frame numbers, so that kernels that lose efficiency on small bands show up without a multi-GPU box.

    python tools/band_probe.py [--workload c4] [--world 8] [--bands 0,3,7]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(scene, b0, rows, reps, vd, at, sh, df, csm, band_halo=0):
    """band_halo > 0: the same rows as a BAND handle (SGI_HALO_RECOMPUTE, band_halo held rows beyond each inner edge, no neighbour
    attached, owner of the frame image): what the band branches of the kernels and of the ABI cost without any peer traffic."""
    import torch
    from stratusgfx_b200.gipass import VplGiPass
    dev = torch.device("cuda", 0)
    W, H, iters = scene.width, scene.height, scene.denoise_iters
    y0, y1 = max(0, b0 - band_halo), min(H, b0 + rows + band_halo)
    gp = VplGiPass(W, H, y0, y1 - y0, scene.cube_res, scene.csm_res, iters, device=0)
    gp.set_gbuffer(scene.gbuffer(0, y0, y1 - y0), scene.gbuffer(-1, y0, y1 - y0))
    gp.set_vpls(vd, at); gp.set_atlases(sh, df); gp.set_csm(csm)
    if band_halo:
        gp.band_setup(b0, rows, 1 if b0 > 0 else 0, 3 if b0 > 0 else 2, owns_frame_image=True, halo_mode=1)
    gp.bind()
    fp = scene.frame_params(0)
    cs = torch.cuda.ExternalStream(gp.stream_ptr(), device=dev)
    order = [("cull", lambda: gp.cull(fp)), ("color", lambda: gp.color(fp)), ("tile_lists", lambda: gp.tile_lists(fp)), ("gather", lambda: gp.gather(fp))]
    order += [(f"pass{i}", (lambda i=i: gp.denoise_pass(fp, i))) for i in range(iters)]
    for _ in range(3):
        gp.frame(fp)
    gp.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(order) + 1)] for _ in range(reps)]
    for k in range(reps):
        evs[k][0].record(cs)
        for j, (_, fn) in enumerate(order):
            fn()
            evs[k][j + 1].record(cs)
    gp.synchronize()
    st = {name: float(np.median([evs[k][j].elapsed_time(evs[k][j + 1]) for k in range(reps)])) for j, (name, _) in enumerate(order)}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cs)
    for _ in range(reps):
        gp.frame(fp)
    e1.record(cs)
    gp.synchronize()
    st["frame"] = e0.elapsed_time(e1) / reps
    gp.close()
    return st


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--bands", default="")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--no-whole", action="store_true")
    ap.add_argument("--band-halo", type=int, default=0, help="also run every band as a band handle that recomputes this many halo rows (no neighbour attached)")
    a = ap.parse_args()
    from stratusgfx_b200.scene import Scene
    scene = Scene(a.workload)
    vd, at = scene.vpls(); sh, df = scene.cubemaps(); csm = scene.csm(0)
    H = scene.height
    h = H // a.world
    which = [int(x) for x in a.bands.split(",")] if a.bands else list(range(a.world))
    out = {} if a.no_whole else {"whole": run(scene, 0, H, a.reps, vd, at, sh, df, csm)}
    for r in which:
        out[f"band{r}"] = run(scene, r * h, h, a.reps, vd, at, sh, df, csm)
        if a.band_halo:
            out[f"band{r}_as_band_handle"] = run(scene, r * h, h, a.reps, vd, at, sh, df, csm, a.band_halo)
    keys = list(next(iter(out.values())).keys())
    print("| rows | " + " | ".join(keys) + " |")
    print("|---|" + "---|" * len(keys))
    for name, st in out.items():
        print(f"| {name} | " + " | ".join(f"{st[k] * 1e3:.1f}" for k in keys) + " |")
    tot = {k: sum(out[f"band{r}"][k] for r in which) for k in keys}
    if a.band_halo:
        print(json.dumps(out)); return
    if len(which) == a.world:
        print("| sum of bands | " + " | ".join(f"{tot[k] * 1e3:.1f}" for k in keys) + " |")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
