#!/usr/bin/env python
"""Golden fixture from the oracle built against the REFERENCE's vendored glm (oracle/_ref/liboracle_glm.so, which exists
only where /root/reference does): two frames of the `tiny` workload and the VPL pre-pass on a seeded light set.

    python tests/golden/make_glm_golden.py          # writes tests/golden/tiny_glm_golden.json

Integer outputs (visible list, atlas entries, per-tile lists, pre-pass order) are stored as they are; float planes as SHA-256 of
their bytes.  The CPU suite checks the oracle's own maths layer against this file (so the pin on glm's arithmetic travels to
boxes without the reference tree) and the GPU suite checks the CUDA path's integer outputs against it directly.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def frame_record(o, V):
    cnt = o.stage2[:, 0]
    mask = np.arange(12)[None, :] < cnt[:, None]
    return dict(V=int(V), visible=o.visible[:V + 1].tolist(), entries=o.entries[:V].reshape(-1).tolist(),
                stage2_counts=cnt.tolist(), stage2_indices=o.stage2[:, 1:][mask].tolist(),
                updated_sha=sha(o.updated[:V]), stage1_sha=sha(o.stage1), gi_color_sha=sha(o.gi_color), gi_reservoir_sha=sha(o.gi_reservoir),
                combined_sha=sha(o.combined), gi_illum_sha=sha(o.gi_illum), history_sha=sha(o.history))


def compute(variant):
    from oracle import oracle_py
    from oracle.oracle_py import OracleFrame
    from stratusgfx_b200.prepass import frustum_planes
    from tests.parity_util import SceneData
    from tests.test_oracle_selfchecks import _random_lights
    d = SceneData("tiny", frames=2)
    s = d.scene
    o = OracleFrame(d.W, d.H, s.cube_res, s.csm_res, s.denoise_iters, variant=variant)
    frames = []
    for f in range(2):
        o.frame(d.fp[f], d.gb[f], d.gb[f - 1], d.vpl_data, d.atlas_by_vpl, d.csm[f], d.shadow_atlas, d.diffuse_atlas)
        frames.append(frame_record(o, int(o.visible[0])))
    rec = _random_lights(3000, 7)
    planes = frustum_planes(list(d.fp[0].invProjectionView))
    vd, at, order = oracle_py.vpl_prepass(rec, list(d.fp[0].viewPosition), planes, variant=variant)
    return dict(workload="tiny", frames=frames, prepass=dict(n=3000, seed=7, order=order.tolist(), vpl_data_sha=sha(vd), atlas_sha=sha(at)))


if __name__ == "__main__":
    from oracle import oracle_py
    oracle_py.build()
    assert oracle_py.has_variant("glm"), "needs /root/reference/glm (oracle/_ref/liboracle_glm.so)"
    out = compute("glm")
    out["made_by"] = "tests/golden/make_glm_golden.py with oracle/_ref/liboracle_glm.so (reference glm 0.9.8.5, GLM_FORCE_PURE)"
    path = os.path.join(ROOT, "tests", "golden", "tiny_glm_golden.json")
    json.dump(out, open(path, "w"))
    print(path, os.path.getsize(path), "bytes")
