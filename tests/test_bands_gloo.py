"""The N>1 host path on CPU: world_size-2 gloo processes run the row-band plan of stratusgfx_b200/bands.py
(broadcast of the per-frame VPL tables, per-band processing with an apron, gather of the lit bands).
The per-band processor here is the CPU oracle in band mode — the GPU box runs the same plumbing with
the CUDA pass and NCCL (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest

from stratusgfx_b200.bands import APRON, plan_bands


def test_band_plan_covers_the_frame_on_the_tile_grid():
    for H, world in ((2160, 1), (2160, 2), (2160, 4), (2160, 8), (4320, 8), (1080, 8), (360, 2)):
        bands = plan_bands(H, world)
        assert [b.b0 for b in bands] == [r * H // world for r in range(world)]
        assert bands[-1].b1 == H and bands[0].y0 == 0 and bands[-1].y1 == H
        for b in bands:
            assert b.b0 % 5 == 0 and b.b1 % 5 == 0 and b.y0 % 5 == 0 and b.rows % 5 == 0
            assert b.y0 <= b.b0 < b.b1 <= b.y1
            if world > 1:
                assert b.b0 - b.y0 in (0, APRON) and b.y1 - b.b1 in (0, APRON)
            assert b.owned == slice(b.b0 - b.y0, b.b1 - b.y0)
    with pytest.raises(ValueError):
        plan_bands(1080, 16)        # 67.5 rows per band
    assert APRON >= 4 + 8 + 16      # reach of the three à-trous passes (BE:1920: multiplier 1, 3, 7 -> 2*(m+1))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, workload, out_path, cuts=None):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "4"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle_py import OracleFrame
    from stratusgfx_b200.bands import bands_from_cuts, broadcast_frame_inputs, gather_bands
    from stratusgfx_b200.scene import Scene
    scene = Scene(workload)
    W, H = scene.width, scene.height
    bands = plan_bands(H, world) if cuts is None else bands_from_cuts(list(cuts), H, APRON)
    me = bands[rank]
    fp = scene.frame_params(0)
    gb, pgb = scene.gbuffer(0, me.y0, me.rows), scene.gbuffer(-1, me.y0, me.rows)
    sh, df = scene.cubemaps()
    csm = scene.csm(0)
    # only rank 0 "knows" this frame's VPL tables; the others receive them
    vd, at = scene.vpls()
    t_vd, t_at = torch.from_numpy(vd.copy()), torch.from_numpy(at.copy())
    if rank != 0:
        t_vd.zero_(); t_at.zero_()
    broadcast_frame_inputs(dist, [t_vd, t_at])
    assert np.array_equal(t_vd.numpy(), vd) and np.array_equal(t_at.numpy(), at)
    o = OracleFrame(W, me.rows, scene.cube_res, scene.csm_res, scene.denoise_iters, y_offset=me.y0, full_height=H)
    o.frame(fp, gb, pgb, t_vd.numpy(), t_at.numpy(), csm, sh, df)
    band_out = torch.from_numpy(np.ascontiguousarray(o.combined[me.owned]))
    full = torch.zeros((H, W, 4), dtype=torch.float16) if rank == 0 else None
    gather_bands(dist, band_out, full, bands, rank)
    if rank == 0:
        np.save(out_path, full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cuts", [None, (0, 205, 360)], ids=["equal", "balanced"])
def test_two_rank_gloo_bands_reproduce_the_whole_frame(built, tmp_path, cuts):
    """Equal bands (one collective gather) and measured-balance bands of unequal height (point-to-point gather)."""
    import torch.multiprocessing as mp
    from tests.parity_util import SceneData, run_oracle
    workload = "c1"
    out_path = str(tmp_path / "assembled.npy")
    mp.spawn(_worker, args=(2, _free_port(), workload, out_path, cuts), nprocs=2, join=True)
    assembled = np.load(out_path)
    whole = run_oracle(SceneData(workload), keep_passes=False)[0]["combined"]
    # apron rows make every owned row see exactly the taps it would see in the whole frame
    assert np.array_equal(assembled.view(np.uint16), whole.view(np.uint16))


# ---- property tests of the host-side band planning (hypothesis) ------------------------------------------------------------
def test_balanced_cuts_are_always_a_valid_partition():
    """Whatever the measured times, balance_cuts must return cuts on the 5-row tile grid that start at 0, end at the frame
    height, increase strictly and leave every band at least as tall as the halo — i.e. something bands_from_cuts accepts."""
    from hypothesis import given, settings, strategies as st
    from stratusgfx_b200.bands import TILE, balance_cuts, bands_from_cuts, halo_rows, plan_bands, recompute_rows

    @settings(max_examples=200, deadline=None)
    @given(world=st.sampled_from([2, 3, 4, 8]), tiles_per_band=st.integers(5, 120), iters=st.integers(1, 5), recompute=st.booleans(),
           times=st.lists(st.floats(1e-3, 10.0, allow_nan=False, allow_infinity=False), min_size=8, max_size=8))
    def check(world, tiles_per_band, iters, recompute, times):
        halo = (recompute_rows if recompute else halo_rows)(iters)          # both halo modes of the band split (include/stratusgi.h)
        height = world * tiles_per_band * TILE
        if height // world < halo:
            return
        bands = plan_bands(height, world, halo)
        cuts = balance_cuts(bands, times[:world], height, halo)
        assert cuts[0] == 0 and cuts[-1] == height and len(cuts) == world + 1
        assert all(c % TILE == 0 for c in cuts) and all(b - a >= max(halo, TILE) for a, b in zip(cuts, cuts[1:]))
        out = bands_from_cuts(cuts, height, halo)
        assert [b.b0 for b in out] == cuts[:-1] and [b.b1 for b in out] == cuts[1:]
        for b in out:                                   # held rows = owned rows + halo, clipped to the frame
            assert b.y0 == max(0, b.b0 - halo) and b.y1 == min(height, b.b1 + halo)
        # equal times leave equal bands where they were
        same = balance_cuts(bands, [1.0] * world, height, halo)
        assert same == [r * (height // world) for r in range(world)] + [height]

    check()


def test_halo_rows_cover_the_reach_of_the_widest_pass():
    from stratusgfx_b200.bands import TILE, halo_rows
    for iters in range(1, 7):
        reach = 2 * (1 << (iters - 1)) if iters > 1 else 0          # 2 * (multiplier + 1), multiplier = 2^(iters-1) - 1 (BE:1920)
        h = halo_rows(iters)
        assert h >= reach and h % TILE == 0 and h - reach < TILE


def test_recompute_rows_cover_the_summed_reach_of_all_passes():
    """SGI_HALO_RECOMPUTE (include/stratusgi.h): pass p is recomputed on the band's rows +- sum_{j>p} 2^(j+1); pass 0 therefore on
    2^(iters+1) - 4 rows, which the halo must hold (28 for the reference's four iterations, 30 on the tile grid), and every pass's
    taps (2^(p+1) rows) stay inside what the pass before it computed."""
    from stratusgfx_b200.bands import TILE, plan_bands, recompute_rows
    for iters in range(1, 7):
        h = recompute_rows(iters)
        ext = [sum(2 << j for j in range(p + 1, iters)) for p in range(iters)]        # bandDimsForPass (csrc/sgi_abi.cu)
        assert h % TILE == 0 and (ext[0] if iters > 1 else 0) <= h < (ext[0] if iters > 1 else 0) + TILE
        assert ext[-1] == 0
        for p in range(1, iters):
            assert ext[p] + (2 << p) == ext[p - 1]          # rows pass p reads = rows pass p-1 wrote
    assert [recompute_rows(i) for i in (1, 4, 5)] == [0, 30, 60]
    bands = plan_bands(2160, 8, recompute_rows(4))
    assert all(b.b0 - b.y0 in (0, 30) and (b.y0 + b.rows) - b.b1 in (0, 30) for b in bands)


def test_both_bench_arms_print_the_same_config_object():
    """bench.py: `config` is a function of the workload and the command line only, so `--impl reference` launched with the same flags
    prints the same object as the CUDA arm (what a run measures or decides lives under `run`)."""
    import argparse
    import json
    import bench
    from stratusgfx_b200.scene import Scene
    scene = Scene("c4")
    for world in (1, 2, 8):
        for bands in ("recompute", "p2p", "apron"):
            for split in ("bands", "views"):
                args = argparse.Namespace(gpus=world, bands=bands, split=split, maps="broadcast", no_early=False, no_prefetch=False, balance=2)
                cfg = bench.bench_config(scene, world, args)
                assert cfg == bench.bench_config(scene, max(args.gpus, 1), args)
                assert cfg["workload"].startswith("c4: 3840x2160") and "parallelism" in cfg
                json.dumps(cfg)
                assert not any(k in cfg for k in ("visible_vpls", "band_rows", "sample", "image_check"))
    assert bench.bench_config(scene, 1, args)["parallelism"] == "single GPU"
    assert "28-row halos" not in bench.bench_config(scene, 8, argparse.Namespace(gpus=8, bands="p2p", split="bands", maps="device", no_early=True, no_prefetch=True, balance=0))["parallelism"]
