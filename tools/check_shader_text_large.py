#!/usr/bin/env python
"""Offline (minutes, single-threaded shader text): the oracle against the reference's compiled shader text on the full-size
1080p workloads — c2 (Cornell, 256 VPLs, 4 temporal frames) and c3 (hall, 1024 VPLs, 5 denoise passes).  The CPU test suite does
the same on `tiny` and c1 (tests/test_reference_glsl.py); this script is the large-frame version of it.

    python tools/check_shader_text_large.py [c2 c3]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_py as O                     # noqa: E402
from tests.parity_util import SceneData               # noqa: E402


def run(workload, frames):
    ref = O.reference_glsl()
    d = SceneData(workload, frames=frames)
    s = d.scene
    o = O.OracleFrame(d.W, d.H, s.cube_res, s.csm_res, s.denoise_iters)
    den = O.ReferenceGlslDenoiser(ref, d.W, d.H, s.denoise_iters)
    ok = True
    for f in range(frames):
        t0 = time.time()
        fp, gb, pgb = d.fp[f], d.gb[f], d.gb[f - 1]
        V = o.cull(fp, d.vpl_data, d.atlas_by_vpl, d.csm[f])
        _, vis = O.reference_glsl_cull(ref, fp, d.vpl_data, d.csm[f], s.csm_res)
        c_ok = int(vis[0]) == V and sorted(vis[1:V + 1].tolist()) == o.visible[1:V + 1].tolist()
        before = o.updated.copy()
        o.color(fp, d.shadow_atlas, d.diffuse_atlas)
        col = O.reference_glsl_color(ref, fp, before, o.visible, o.entries, d.shadow_atlas, d.diffuse_atlas, s.cube_res)
        col_ok = np.array_equal(col[:V].view(np.uint32), o.updated[:V].view(np.uint32))
        o.tile_lists(fp, gb, d.vpl_data)
        s1_ok = np.array_equal(O.reference_glsl_stage1(ref, fp, gb).view(np.uint32), o.stage1.view(np.uint32))
        s2_ok = None
        if (d.W // 5) % 32 == 0 and (d.H // 5) % 2 == 0:
            s2 = O.reference_glsl_stage2(ref, fp, d.W, d.H, d.vpl_data, o.stage1, o.visible)
            cnt = o.stage2[:, 0]; mask = np.arange(12)[None, :] < cnt[:, None]
            s2_ok = np.array_equal(s2[:, 0], cnt) and np.array_equal(s2[:, 1:][mask], o.stage2[:, 1:][mask])
        o.gather(fp, gb)
        color, res = O.reference_glsl_gather(ref, fp, gb, o.updated, o.visible, o.entries, d.shadow_atlas, s.cube_res)
        g_ok = np.array_equal(O._to_half(res).view(np.uint16), o.gi_reservoir.view(np.uint16)) and \
            np.array_equal(O._to_half(color).view(np.uint16), o.gi_color[..., :3].view(np.uint16))
        passes = den.frame(fp, gb, pgb, V, o.gi_color, o.gi_reservoir)
        o.denoise(fp, gb, pgb, keep_passes=True)
        p_ok = all(np.array_equal(p["reservoir"].view(np.uint16), q.view(np.uint16)) for p, q in zip(passes, o.pass_reservoirs))
        f_ok = np.array_equal(passes[-1]["combined"][..., :3].view(np.uint16), o.combined[..., :3].view(np.uint16)) and \
            np.array_equal(passes[-1]["history"].view(np.uint16), o.history.view(np.uint16)) and \
            np.array_equal(passes[-1]["gi"][..., :3].view(np.uint16), o.gi_illum[..., :3].view(np.uint16))
        print(f"{workload} frame {f}: V={V} cull={c_ok} colour={col_ok} stage1={s1_ok} stage2={s2_ok} gather={g_ok} passes={p_ok} final={f_ok}  ({time.time() - t0:.0f} s)", flush=True)
        ok = ok and c_ok and col_ok and s1_ok and (s2_ok is not False) and g_ok and p_ok and f_ok
    return ok


if __name__ == "__main__":
    todo = sys.argv[1:] or ["c2", "c3"]
    good = all([run(w, 4 if w == "c2" else 1) for w in todo])
    print("SHADER TEXT vs ORACLE:", "ALL BIT-EXACT" if good else "MISMATCH")
    sys.exit(0 if good else 1)
