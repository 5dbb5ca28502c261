"""Per-CTA phase timing of stage 2 on two bands of c4.  Needs a library built with SGI_EXTRA_NVCC_FLAGS=-DSGI_S2_PROFILE."""
import sys, os, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from stratusgfx_b200.scene import Scene
from stratusgfx_b200.gipass import VplGiPass, load_library
scene = Scene("c4")
vd, at = scene.vpls(); sh, df = scene.cubemaps(); csm = scene.csm(0)
H = scene.height; h = H // 8
L = load_library()
for band in (5, 0):
    b0 = band * h
    gp = VplGiPass(scene.width, H, b0, h, scene.cube_res, scene.csm_res, scene.denoise_iters, device=0)
    gp.set_gbuffer(scene.gbuffer(0, b0, h), scene.gbuffer(-1, b0, h)); gp.set_vpls(vd, at); gp.set_atlases(sh, df); gp.set_csm(csm); gp.bind()
    fp = scene.frame_params(0)
    ncta = (scene.width // 5 + 7) // 8 * ((h // 5 + 7) // 8)
    buf = torch.zeros((ncta, 10), dtype=torch.float32, device="cuda")
    for _ in range(2):
        gp.frame(fp)
    gp.synchronize()
    L.sgi_debug_s2prof.argtypes = [C.c_void_p]
    assert L.sgi_debug_s2prof(buf.data_ptr()) == 0
    gp.cull(fp); gp.color(fp); gp.tile_lists(fp); gp.synchronize()
    L.sgi_debug_s2prof(None)
    a = buf.cpu().numpy()
    tot = a[:, 1] + a[:, 2] + a[:, 3] + a[:, 4]
    print("band", band, "ctas", ncta, "V", gp.visible_count())
    for name, col in (("setup", a[:, 1]), ("bin", a[:, 2]), ("scan", a[:, 3]), ("merge", a[:, 4]), ("total", tot), ("rounds", a[:, 5]), ("cands", a[:, 6]), ("maxins", a[:, 7]), ("R", a[:, 8])):
        print(f"  {name:6s} mean {col.mean():10.1f}  p50 {np.percentile(col,50):10.1f}  p90 {np.percentile(col,90):10.1f}  p99 {np.percentile(col,99):10.1f}  max {col.max():10.1f}")
    st = a[:, 9]; print("  start-clock spread (low 24 bits):", st.min(), st.max())
    order = np.argsort(-tot)[:6]
    for i in order:
        print("   slow cta", i, a[i])
    gp.close()
