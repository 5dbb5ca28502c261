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
