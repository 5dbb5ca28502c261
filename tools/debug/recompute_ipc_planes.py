"""Debug: band 0 of a 3-band SGI_HALO_RECOMPUTE split, staged calls, planes after every pass — once with the three bands in ONE
process, once with one process per band on the same GPU.  Prints where the planes of band 0 differ between the two."""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from stratusgfx_b200 import _abi

PLANES = {"resA": _abi.PLANE_DENOISE_RES_A, "resB": _abi.PLANE_DENOISE_RES_B, "reservoir": _abi.PLANE_GI_RESERVOIR, "combined": _abi.PLANE_COMBINED}


def staged_frame(gp, fp, iters, dump, tag):
    if os.environ.get("USE_FRAME"):
        gp.frame(fp); gp.synchronize()
        if dump:
            np.save(dump % (tag, "g_resA"), gp.plane(PLANES["resA"])); np.save(dump % (tag, "g_reservoir"), gp.plane(PLANES["reservoir"]))
            np.save(dump % (tag, "p1_res"), gp.plane(PLANES["resB"])); np.save(dump % (tag, "p2_res"), gp.plane(PLANES["resA"])); np.save(dump % (tag, "p3_res"), gp.plane(PLANES["resB"]))
            np.save(dump % (tag, "combined"), gp.plane(PLANES["combined"]))
            np.save(dump % (tag, "gi"), gp.plane(_abi.PLANE_GI_ILLUM)); np.save(dump % (tag, "hist"), gp.plane(_abi.PLANE_HISTORY))
            np.save(dump % (tag, "image"), gp.frame_image().cpu().numpy()[:gp.rows])
        return
    gp.cull(fp); gp.color(fp); gp.gather(fp)
    if dump: np.save(dump % (tag, "g_resA"), gp.plane(PLANES["resA"])); np.save(dump % (tag, "g_reservoir"), gp.plane(PLANES["reservoir"]))
    for p in range(iters):
        gp.denoise_pass(fp, p)
        if dump and p >= 1:
            np.save(dump % (tag, f"p{p}_res"), gp.plane(PLANES["resB" if p % 2 else "resA"]))
    if dump: np.save(dump % (tag, "combined"), gp.plane(PLANES["combined"]))


def worker(rank, world, port, dump, share):
    import torch, torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from stratusgfx_b200.bands import connect_bands, plan_bands, recompute_rows
    from stratusgfx_b200.gipass import VplGiPass
    from stratusgfx_b200.scene import Scene
    scene = Scene("c1", yaw_deg_per_frame=0.25)
    bands = plan_bands(scene.height, world, recompute_rows(scene.denoise_iters))
    me = bands[rank]
    vd, at = scene.vpls(); sh, df = scene.cubemaps()
    gp = VplGiPass(scene.width, scene.height, me.y0, me.rows, scene.cube_res, scene.csm_res, scene.denoise_iters, device=0)
    gp.set_vpls(vd, at); gp.set_atlases(sh, df)
    connect_bands(dist, gp, me, world, halo_mode=1)
    gp.set_gbuffer(scene.gbuffer(0, me.y0, me.rows), scene.gbuffer(-1, me.y0, me.rows)); gp.set_csm(scene.csm(0)); gp.bind()
    staged_frame(gp, scene.frame_params(0), scene.denoise_iters, dump if rank == 0 else None, "multi")
    gp.synchronize(); dist.barrier(); gp.close(); dist.destroy_process_group()


if __name__ == "__main__":
    import torch.multiprocessing as mp
    from tests.test_multigpu_nccl import _free_port
    tmp = tempfile.mkdtemp(); dump = os.path.join(tmp, "%s_%s.npy")
    mp.spawn(worker, args=(3, _free_port(), dump, True), nprocs=3, join=True)
    # the same three bands in one process
    from stratusgfx_b200.bands import plan_bands, recompute_rows
    from stratusgfx_b200._abi import PEER_DOWN, PEER_ROOT, PEER_UP
    from stratusgfx_b200.gipass import VplGiPass
    from stratusgfx_b200.scene import Scene
    scene = Scene("c1", yaw_deg_per_frame=0.25)
    bands = plan_bands(scene.height, 3, recompute_rows(scene.denoise_iters))
    vd, at = scene.vpls(); sh, df = scene.cubemaps()
    gps = []
    for b in bands:
        gp = VplGiPass(scene.width, scene.height, b.y0, b.rows, scene.cube_res, scene.csm_res, scene.denoise_iters, device=0)
        gp.set_vpls(vd, at); gp.set_atlases(sh, df)
        gp.band_setup(b.b0, b.b1 - b.b0, b.rank, 3, owns_frame_image=(b.rank == 0), halo_mode=1)
        gps.append(gp)
    infos = [gp.band_export() for gp in gps]
    for gp, b in zip(gps, bands):
        if b.rank > 0: gp.band_attach(PEER_UP, infos[b.rank - 1])
        if b.rank + 1 < 3: gp.band_attach(PEER_DOWN, infos[b.rank + 1])
        if b.rank != 0: gp.band_attach(PEER_ROOT, infos[0])
    for gp, b in zip(gps, bands):
        gp.set_gbuffer(scene.gbuffer(0, b.y0, b.rows), scene.gbuffer(-1, b.y0, b.rows)); gp.set_csm(scene.csm(0)); gp.bind()
    if os.environ.get("USE_FRAME"):
        for gp in gps[1:]: gp.frame(scene.frame_params(0))
    staged_frame(gps[0], scene.frame_params(0), scene.denoise_iters, dump, "single")
    names = ("g_resA", "g_reservoir", "p1_res", "p2_res", "p3_res", "combined") + (("gi", "hist", "image") if os.environ.get("USE_FRAME") else ())
    for name in names:
        a = np.load(dump % ("multi", name)); b = np.load(dump % ("single", name))
        a = a.reshape(bands[0].rows, scene.width, -1).view(np.uint16); b = b.reshape(a.shape).view(np.uint16)
        d = (a != b).any(axis=2)
        ys, xs = np.nonzero(d)
        print(name, "differing pixels", int(d.sum()), "rows", (int(ys.min()), int(ys.max())) if len(ys) else None, "cols", (int(xs.min()), int(xs.max())) if len(xs) else None)
