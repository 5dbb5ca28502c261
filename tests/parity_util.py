"""Shared parity harness: run the CUDA path (through the C-ABI) and the CPU oracle on the same
seeded inputs and compare stage by stage.

Bars (BASELINE.json north_star): integer outputs bit-exact; radiance within 1e-3 relative per
pixel or PSNR >= 60 dB over the frame.  Tolerances are stated where they are applied.

    python -m tests.parity_util c1        # prints a JSON report (needs a GPU)
"""
import json
import sys

import numpy as np

from stratusgfx_b200 import _abi
from stratusgfx_b200.scene import Scene

REL_TOL = 1e-3          # north_star: radiance within 1e-3 relative per pixel ...
PSNR_MIN_DB = 60.0      # ... or PSNR >= 60 dB over the frame
ABS_FLOOR = 1e-4        # values below this are compared absolutely (relative error is meaningless at ~0)


class SceneData:
    """All inputs of one workload, generated once (numpy, host)."""

    def __init__(self, workload="c1", frames=1, yaw_deg_per_frame=0.0, **overrides):
        self.scene = Scene(workload, yaw_deg_per_frame=yaw_deg_per_frame, **overrides)
        s = self.scene
        self.W, self.H, self.N = s.width, s.height, s.num_vpls
        self.vpl_data, self.atlas_by_vpl = s.vpls()
        self.shadow_atlas, self.diffuse_atlas = s.cubemaps()
        self.frames = frames
        self.fp = [s.frame_params(f) for f in range(frames)]
        self.gb = {f: s.gbuffer(f) for f in range(-1, frames)}
        self.csm = [s.csm(f) for f in range(frames)] if yaw_deg_per_frame != 0.0 else [s.csm(0)] * frames


def psnr_db(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    mse = np.mean((a - b) ** 2)
    peak = max(float(np.max(np.abs(b))), 1e-12)
    return float("inf") if mse == 0 else 10.0 * np.log10(peak * peak / mse)


def compare_radiance(got_h, want_h, channels=3):
    """half planes -> dict(bit_exact fraction, within-tolerance fraction, max rel err, PSNR)."""
    g = got_h[..., :channels].astype(np.float32) if got_h.ndim == 3 else got_h.astype(np.float32)
    w = want_h[..., :channels].astype(np.float32) if want_h.ndim == 3 else want_h.astype(np.float32)
    exact = float(np.mean(got_h.view(np.uint16) == want_h.view(np.uint16)))
    diff = np.abs(g - w)
    rel = diff / np.maximum(np.abs(w), ABS_FLOOR)
    px_ok = (rel <= REL_TOL)
    if px_ok.ndim == 3:
        px_ok = px_ok.all(axis=-1)
    return dict(bit_exact=exact, within_tol=float(np.mean(px_ok)), max_rel=float(rel.max()), psnr_db=psnr_db(g, w),
                ok=bool(px_ok.all() or psnr_db(g, w) >= PSNR_MIN_DB))


def make_pass(data, y0=0, rows=None, denoise_iters=None):
    from stratusgfx_b200.gipass import VplGiPass
    s = data.scene
    rows = data.H - y0 if rows is None else rows
    gp = VplGiPass(data.W, data.H, y0, rows, s.cube_res, s.csm_res, denoise_iters or s.denoise_iters)
    gp.set_vpls(data.vpl_data, data.atlas_by_vpl)
    gp.set_atlases(data.shadow_atlas, data.diffuse_atlas)
    gp.set_csm(data.csm[0])
    return gp


def band_of(gb, y0, rows):
    """Slice a whole-frame scene.GBuffer into a dict of the band's rows."""
    return {k: getattr(gb, k)[y0:y0 + rows] for k in ("depth", "normal", "albedo", "baseReflectivity", "roughnessMetallicAmbient",
                                                       "velocity", "id", "ssao", "screen")}


def run_oracle(data, frames=None, denoise_iters=None, keep_passes=True):
    from oracle.oracle_py import OracleFrame
    s = data.scene
    o = OracleFrame(data.W, data.H, s.cube_res, s.csm_res, denoise_iters or s.denoise_iters)
    outs = []
    for f in range(frames or data.frames):
        o.frame(data.fp[f], data.gb[f], data.gb[f - 1], data.vpl_data, data.atlas_by_vpl, data.csm[f], data.shadow_atlas, data.diffuse_atlas,
                keep_passes=keep_passes)
        outs.append(dict(visible=o.visible.copy(), updated=o.updated.copy(), entries=o.entries.copy(), stage1=o.stage1.copy(),
                         stage2=o.stage2.copy(), gi_color=o.gi_color.copy(), gi_reservoir=o.gi_reservoir.copy(),
                         pass_reservoirs=[p.copy() for p in o.pass_reservoirs], combined=o.combined.copy(), gi_illum=o.gi_illum.copy(),
                         history=o.history.copy(), sample_stats=o.sample_stats.copy()))
    return outs


def run_cuda_staged(data, frames=None, denoise_iters=None):
    """Stage-by-stage run through the C-ABI, downloading every intermediate."""
    gp = make_pass(data, denoise_iters=denoise_iters)
    iters = gp.denoise_iters
    outs = []
    for f in range(frames or data.frames):
        fp = data.fp[f]
        gp.set_gbuffer(data.gb[f], data.gb[f - 1])
        gp.set_csm(data.csm[f])
        gp.bind()
        gp.PerformVirtualPointLightCullingStage1(fp)
        v = gp.visible_count()
        o = dict(V=v, visible=gp.output("vplVisibleIndices"))
        gp.color(fp)
        o["updated"] = gp.output("vplUpdatedData")
        o["entries"] = gp.output("shadowDiffuseIndices")
        gp.tile_lists(fp)
        o["stage1"] = gp.output("vplStage1Results")
        o["stage2"] = gp.output("vplVisiblePerTile")
        gp.gather(fp)
        o["gi_color"] = gp.plane(_abi.PLANE_GI_COLOR)
        o["gi_reservoir"] = gp.plane(_abi.PLANE_GI_RESERVOIR)
        o["pass_reservoirs"] = []
        for i in range(iters):
            gp.denoise_pass(fp, i)
            o["pass_reservoirs"].append(gp.plane(_abi.PLANE_DENOISE_RES_A if i % 2 == 0 else _abi.PLANE_DENOISE_RES_B))
        o["combined"] = gp.output("lightingOut")
        o["gi_illum"] = gp.plane(_abi.PLANE_GI_ILLUM)
        o["history"] = gp.plane(_abi.PLANE_HISTORY)
        o["launches"] = gp.launch_count()
        outs.append(o)
    gp.close()
    return outs


def compare_frame(cu, orc, tiles_count_only=False):
    """-> (report dict, list of failure strings)"""
    rep, bad = {}, []
    V = int(orc["visible"][0])
    rep["V"] = V
    # integer outputs: bit-exact
    vis_ok = int(cu["visible"][0]) == V and np.array_equal(cu["visible"][:V + 1], orc["visible"][:V + 1])
    rep["visible_list_exact"] = bool(vis_ok)
    ent_ok = np.array_equal(cu["entries"][:V], orc["entries"][:V])
    rep["atlas_entries_exact"] = bool(ent_ok)
    upd_ok = np.array_equal(cu["updated"][:V].view(np.uint32), orc["updated"][:V].view(np.uint32))
    rep["vpl_updated_bitexact"] = bool(upd_ok)
    s1_ok = np.array_equal(cu["stage1"].view(np.uint32), orc["stage1"].view(np.uint32))
    rep["stage1_bitexact"] = bool(s1_ok)
    cnt = orc["stage2"][:, 0]
    s2_ok = np.array_equal(cu["stage2"][:, 0], cnt)
    if s2_ok:
        mask = np.arange(12)[None, :] < cnt[:, None]
        s2_ok = np.array_equal(cu["stage2"][:, 1:][mask], orc["stage2"][:, 1:][mask])
    rep["stage2_lists_exact"] = bool(s2_ok)
    for name, ok in (("visible list", vis_ok), ("atlas entries", ent_ok), ("updated VPL data", upd_ok), ("stage1", s1_ok), ("stage2 lists", s2_ok)):
        if not ok:
            bad.append(name + " not bit-exact")
    # radiance planes
    for name in ("gi_color", "gi_reservoir", "combined", "gi_illum"):
        r = compare_radiance(cu[name], orc[name], 4 if name == "gi_reservoir" else 3)
        rep[name] = r
        if not r["ok"]:
            bad.append(f"{name}: within_tol={r['within_tol']:.6f} psnr={r['psnr_db']:.1f} dB")
    for i, (a, b) in enumerate(zip(cu["pass_reservoirs"], orc["pass_reservoirs"])):
        r = compare_radiance(a, b, 3)
        rep[f"denoise_pass{i}"] = r
        if not r["ok"]:
            bad.append(f"denoise pass {i}: within_tol={r['within_tol']:.6f} psnr={r['psnr_db']:.1f} dB")
    h_ok = np.array_equal(cu["history"].view(np.uint16), orc["history"].view(np.uint16))
    rep["history_exact"] = bool(h_ok)
    if not h_ok:
        bad.append("history plane differs")
    return rep, bad


def main(argv):
    wl = argv[1] if len(argv) > 1 else "c1"
    frames = int(argv[2]) if len(argv) > 2 else 1
    data = SceneData(wl, frames=frames)
    orc = run_oracle(data)
    cu = run_cuda_staged(data)
    allbad = []
    for f in range(frames):
        rep, bad = compare_frame(cu[f], orc[f])
        print(json.dumps(dict(workload=wl, frame=f, launches=cu[f]["launches"], **rep)))
        allbad += bad
    print("PARITY", "FAIL: " + "; ".join(allbad) if allbad else "OK")
    return 1 if allbad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
