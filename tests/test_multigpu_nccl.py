"""Row-band split on real GPUs: one process per GPU, NCCL broadcast of the per-frame VPL tables, the CUDA
pass on each band (+apron), NCCL gather of the lit bands — the assembled image must equal the single-GPU
whole-frame image bit for bit, and the oracle within tolerance.  Needs >= 2 GPUs (gpurun --gpus 2)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _p2p_worker(rank, world, port, workload, out_path, share_gpu, frames, yaw=0.25, halo=None, halo_mode=0):
    """One process per band; peer descriptors travel as bytes (all_gather_object), memory is attached with CUDA IPC."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    device = 0 if share_gpu else rank
    torch.cuda.set_device(device)
    dev = torch.device("cuda", device)
    if share_gpu:
        dist.init_process_group("gloo", rank=rank, world_size=world)          # both ranks on one GPU: NCCL refuses that, gloo carries the bytes
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from stratusgfx_b200.bands import connect_bands, halo_rows, plan_bands, recompute_rows
    from stratusgfx_b200.gipass import VplGiPass
    from stratusgfx_b200.scene import Scene
    scene = Scene(workload, yaw_deg_per_frame=yaw)
    W, H = scene.width, scene.height
    bands = plan_bands(H, world, (recompute_rows if halo_mode else halo_rows)(scene.denoise_iters) if halo is None else halo)
    me = bands[rank]
    vd, at = scene.vpls()
    sh, df = scene.cubemaps()
    gp = VplGiPass(W, H, me.y0, me.rows, scene.cube_res, scene.csm_res, scene.denoise_iters, device=device)
    gp.set_vpls(vd, at)
    gp.set_atlases(sh, df)
    connect_bands(dist, gp, me, world, halo_mode=halo_mode)
    for f in range(frames):
        gp.set_gbuffer(scene.gbuffer(f, me.y0, me.rows), scene.gbuffer(f - 1, me.y0, me.rows))
        gp.set_csm(scene.csm(f))
        gp.bind()
        gp.frame(scene.frame_params(f))
        gp.synchronize()
        dist.barrier()                      # every band's final pass has completed: the assembled image is final
        if rank == 0:
            np.save(out_path % f, gp.frame_image().cpu().numpy())
    assert gp.band_timeouts() == 0
    with open((out_path % 0) + f".misses{rank}", "w") as fh:
        fh.write(str(gp.band_reprojection_misses()))
    dist.barrier()
    gp.close()
    dist.destroy_process_group()


def _check_p2p(tmp_path, world, share_gpu, workload="c1", frames=3, yaw=0.25, halo=None, expect_equal=True, halo_mode=0):
    import torch.multiprocessing as mp
    from tests.parity_util import SceneData, run_cuda_staged
    out_path = str(tmp_path / "p2p_frame%d.npy")
    mp.spawn(_p2p_worker, args=(world, _free_port(), workload, out_path, share_gpu, frames, yaw, halo, halo_mode), nprocs=world, join=True)
    misses = sum(int(open((out_path % 0) + f".misses{r}").read()) for r in range(world))
    single = run_cuda_staged(SceneData(workload, frames=frames, yaw_deg_per_frame=yaw))
    equal = all(np.array_equal(np.load(out_path % f).view(np.uint16), single[f]["combined"].view(np.uint16)) for f in range(frames))
    if expect_equal:
        assert misses == 0 and equal
    return misses, equal


def test_fast_camera_needs_a_halo_as_tall_as_the_motion(built, tmp_path):
    """The final pass reads the previous frame at uv - velocity (vpl_pbr_gi_denoise.fs:280,352-360).  A 12-degree yaw per frame
    moves pixels of c1 by up to ~50 rows away from the horizon line: with four bands of 90 rows and the default 20-row halo the
    bands count reprojection taps outside their held rows (sgi_band_reprojection_misses) and the assembled image differs from the
    one-GPU frame; with a 60-row halo — the halo height is the caller's choice — there is no miss and the image is the one-GPU
    frame byte for byte.  (Two bands would not do: a yaw never moves a pixel across the horizon line, which is their cut.)"""
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    misses, equal = _check_p2p(tmp_path / "a", 4, share_gpu=True, workload="c1", frames=2, yaw=12.0, halo=None, expect_equal=False)
    assert misses > 0 and not equal
    _check_p2p(tmp_path / "b", 4, share_gpu=True, workload="c1", frames=2, yaw=12.0, halo=60)



def test_p2p_bands_over_cuda_ipc_two_processes_on_one_gpu(built, tmp_path):
    """The cross-process path (CUDA IPC mapping of the neighbour's slab, epoch flags) exercised on a single GPU."""
    _check_p2p(tmp_path, 2, share_gpu=True)


def test_recompute_bands_over_cuda_ipc_three_processes_on_one_gpu(built, tmp_path):
    """SGI_HALO_RECOMPUTE across processes (a middle band included): the GI rows of the final pass and the lit rows are what
    still travels through the CUDA IPC mappings; the frames must be the one-GPU frames byte for byte."""
    _check_p2p(tmp_path, 3, share_gpu=True, halo_mode=1)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_p2p_bands_over_nvlink_equal_single_gpu_frames(built, tmp_path, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    _check_p2p(tmp_path, world, share_gpu=False, workload="c2", frames=2)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_recompute_bands_over_nvlink_equal_single_gpu_frames(built, tmp_path, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    _check_p2p(tmp_path, world, share_gpu=False, workload="c2", frames=3, halo_mode=1)


def _worker(rank, world, port, workload, out_path):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from stratusgfx_b200.bands import broadcast_frame_inputs, gather_bands, plan_bands
    from stratusgfx_b200.gipass import VplGiPass
    from stratusgfx_b200.scene import Scene
    scene = Scene(workload)
    W, H = scene.width, scene.height
    bands = plan_bands(H, world)
    me = bands[rank]
    fp = scene.frame_params(0)
    vd, at = scene.vpls()
    if rank != 0:                       # only rank 0 has this frame's tables before the broadcast
        vd = np.zeros_like(vd); at = np.zeros_like(at)
    sh, df = scene.cubemaps()
    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        gp = VplGiPass(W, H, me.y0, me.rows, scene.cube_res, scene.csm_res, scene.denoise_iters, device=rank, use_torch_stream=True)
        gp.set_gbuffer(scene.gbuffer(0, me.y0, me.rows), scene.gbuffer(-1, me.y0, me.rows))
        gp.set_vpls(vd, at)
        gp.set_atlases(sh, df)
        gp.set_csm(scene.csm(0))
        gp.bind()
        broadcast_frame_inputs(dist, [gp.t["vplData"], gp.t["atlasEntryByVpl"]])
        gp.frame(fp)
        full = torch.zeros((H, W, 4), dtype=torch.float16, device=dev) if rank == 0 else None
        gather_bands(dist, gp.t["lightingOut"][me.owned].contiguous(), full, bands, rank)
        torch.cuda.synchronize(dev)
        if rank == 0:
            np.save(out_path, full.cpu().numpy())
        gp.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_row_bands_equal_single_gpu_frame(built, tmp_path, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    from tests.parity_util import SceneData, compare_radiance, run_cuda_staged, run_oracle
    workload = "c2"      # 1080 rows: splits into 2, 4 and 8 bands of whole tiles
    out_path = str(tmp_path / f"assembled{world}.npy")
    mp.spawn(_worker, args=(world, _free_port(), workload, out_path), nprocs=world, join=True)
    assembled = np.load(out_path)
    d = SceneData(workload)
    single = run_cuda_staged(d)[0]["combined"]
    assert np.array_equal(assembled.view(np.uint16), single.view(np.uint16))
    assert compare_radiance(assembled, run_oracle(d, keep_passes=False)[0]["combined"])["ok"]
