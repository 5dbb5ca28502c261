#!/usr/bin/env python
"""Device times of the widening rows on one GPU: deferred direct lighting (pbr.fs), the two SSAO passes, TAA,
the VPL cube-map production (5 maps, the reference's per-frame budget), on the G-buffer of a bench workload.

    python tools/bench_widening.py [--workload c4] [--reps 20]

These kernels are exact-tier (every operation an IEEE one in the shader's order: their outputs are bit-identical to the compiled
shader text, tests/test_zz_*.py); this script gives them what the benchmarked path has — CUDA-event times, algorithmic bytes and the
fraction of the measured HBM peak — so that the next optimisation step starts from numbers.  Each entry point is called once through
the Python wrapper (which uploads its inputs) and then re-issued with the same device-resident arguments inside the timed loop.
Prints one JSON object.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    import torch
    from stratusgfx_b200.gipass import VplGiPass
    from stratusgfx_b200.scene import Scene
    from stratusgfx_b200.ssao import rotation_table, ssao_uniforms, structure_buffer
    scene = Scene(a.workload, yaw_deg_per_frame=0.25)
    W, H, N = scene.width, scene.height, scene.num_vpls
    P = W * H
    dev = torch.device("cuda", 0)
    gp = VplGiPass(W, H, 0, H, scene.cube_res, scene.csm_res, scene.denoise_iters, device=0)
    vd, at = scene.vpls(); sh, df = scene.cubemaps()
    gb, pgb = scene.gbuffer(0), scene.gbuffer(-1)
    gp.set_gbuffer(gb, pgb); gp.set_vpls(vd, at); gp.set_atlases(sh, df); gp.set_csm(scene.csm(0)); gp.bind()
    fp = scene.frame_params(0)
    cs = torch.cuda.ExternalStream(gp.stream_ptr(), device=dev)
    peak = 6650.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])

    def timed(entry, first_call, bytes_per_launch, launches_note, fix_up=None):
        """first_call() goes through the wrapper once (uploads + the call); the captured argument tuple is then re-issued.
        fix_up(args): re-point device pointers of the captured parameter struct at tensors this script keeps alive."""
        captured = {}
        orig = getattr(gp.L, entry)

        def capture(*args):
            captured["args"] = args
            return orig(*args)
        setattr(gp.L, entry, capture)
        try:
            first_call()
        finally:
            setattr(gp.L, entry, orig)
        gp.synchronize()
        args = captured["args"]
        if fix_up is not None:
            fix_up(args)
        for _ in range(3):
            orig(*args)
        gp.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.reps)]
        for e0, e1 in evs:
            e0.record(cs); rc = orig(*args); e1.record(cs)
            assert rc == 0, (entry, rc)
        gp.synchronize()
        ms = float(np.median([e0.elapsed_time(e1) for e0, e1 in evs]))
        gbs = bytes_per_launch / (ms * 1e-3) / 1e9
        return {"ms": round(ms, 4), "algorithmic_bytes": int(bytes_per_launch), "gbs": round(gbs, 1), "hbm_frac": round(gbs / peak, 4), "what": launches_note}

    out = {"workload": a.workload, "pixels": P, "hbm_peak_gbs": peak}
    # ---- pbr.fs: 3 point lights, 2 shadow-casting lights (8-tap PCF), the sun through the CSM, emissive ----
    rng = np.random.default_rng(5)
    lights = np.zeros((3, 12), np.float32)
    lights[:, 0:3] = vd[rng.integers(0, N, 3), 0:3]; lights[:, 3] = 1.0
    lights[:, 4:7] = rng.uniform(50, 400, (3, 3)); lights[:, 7] = 1.0; lights[:, 8] = 40.0; lights[:, 9] = 40.0
    pick = rng.integers(0, N, 2)
    shadow_lights = np.zeros((2, 12), np.float32)
    shadow_lights[:, 0:4] = vd[pick, 0:4]; shadow_lights[:, 4:7] = rng.uniform(50, 400, (2, 3)); shadow_lights[:, 7] = 1.0
    shadow_lights[:, 8] = 500.0; shadow_lights[:, 9] = vd[pick, 13]
    keep = [torch.from_numpy(x).to(dev) for x in (lights, shadow_lights, np.ascontiguousarray(at[pick], np.int32))]

    def repoint(args):          # the wrapper's uploads are locals of the wrapper: use tensors that stay alive
        dp = args[2]._obj
        dp.nonShadowCasters, dp.shadowCasters, dp.shadowIndices = keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr()
    out["direct_lighting"] = timed("sgi_direct_lighting",
                                   lambda: gp.direct_lighting(fp, lights=lights, shadow_lights=shadow_lights, shadow_entries=at[pick], emission_strength=0.5),
                                   28 * P, "pbr.fs: 20 B of G-buffer read + half4 written per pixel; 3 + 2 lights and the sun", repoint)
    # ---- ssao.fs + ssao_blur.fs ----
    structure, rot, u = structure_buffer(fp, gb.depth), rotation_table(), ssao_uniforms(W, H)
    out["ssao"] = timed("sgi_ssao", lambda: gp.ssao(structure, rot, **u), 22 * P, "ssao.fs + ssao_blur.fs: structure half4 read twice, R16F written twice, R16F read once")
    # ---- taa.fs ----
    screen = np.clip(gb.screen[..., :3].astype(np.float32), 0.0, 1.0)
    ldr = np.concatenate([np.rint(screen * 255.0).astype(np.uint8), np.full((H, W, 1), 255, np.uint8)], -1)
    out["taa"] = timed("sgi_taa", lambda: gp.taa(ldr, ldr, gb.velocity), 24 * P, "taa.fs: rgba8 current + previous, half2 velocity read; half4 + rgba8 written")
    # ---- cube shadow / diffuse maps of 5 VPLs ----
    gp.set_scene_boxes(scene.boxes())
    five = torch.arange(min(5, N), dtype=torch.int32, device=dev)
    out["cubemaps_5"] = timed("sgi_render_vpl_cubemaps", lambda: gp.render_vpl_cubemaps(five), 5 * 36864, "5 maps x (32x32x6 unorm16 + rgba8): launch-latency bound")
    gp.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
