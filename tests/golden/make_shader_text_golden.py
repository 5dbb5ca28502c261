#!/usr/bin/env python
"""Golden fixture from the REFERENCE'S OWN SHADER TEXT: the six shaders of the GI path compiled from /root/reference/Source/Shaders
by oracle/make_ref_glsl.py (-> oracle/_ref/libref_glsl.so) and chained as the reference's backend chains them, on two frames of the
`tiny` workload.  No function of oracle/oracle.cpp takes part.

    python tests/golden/make_shader_text_golden.py        # writes tests/golden/tiny_shader_text_golden.json

Chain (StratusRendererBackend.cpp:1578-1952): viscull_vpls.cs -> (host: atlas entries of the visible list, BE:1690-1701) ->
vpl_light_color.cs -> tiled culling stage 1 -> stage 2 -> vpl_pbr_gi.fs -> vpl_pbr_gi_denoise.fs x 4 with the FBO ping-pong.
The cull's helper invocations race for the output slots; the list is put in ascending order, which is the order of the shader's
own serial variant (viscull_vpls.cs:108-119) and the canonical order of this repository.  Same record layout as
make_glm_golden.py, so the two fixtures can be compared field by field.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def compute():
    from make_glm_golden import sha
    from oracle import oracle_py as O
    from tests.parity_util import SceneData
    ref = O.reference_glsl()
    d = SceneData("tiny", frames=2)
    s = d.scene
    den = O.ReferenceGlslDenoiser(ref, d.W, d.H, s.denoise_iters)
    frames = []
    for f in range(2):
        fp, gb, pgb = d.fp[f], d.gb[f], d.gb[f - 1]
        _, vis = O.reference_glsl_cull(ref, fp, d.vpl_data, d.csm[f], s.csm_res)
        V = int(vis[0])
        idx = np.sort(vis[1:V + 1])                                    # canonical (serial-variant) order
        visible = np.concatenate([[V], idx]).astype(np.int32)
        updated = np.zeros((4480, 16), np.float32); updated[:V] = d.vpl_data[idx]
        entries = np.zeros((4480, 2), np.int32); entries[:V] = d.atlas_by_vpl[idx]      # the host's shadowDiffuseIndices (BE:1690-1701)
        updated = O.reference_glsl_color(ref, fp, updated, visible, entries, d.shadow_atlas, d.diffuse_atlas, s.cube_res)
        stage1 = O.reference_glsl_stage1(ref, fp, gb)
        stage2 = O.reference_glsl_stage2(ref, fp, d.W, d.H, d.vpl_data, stage1, visible)
        color, res = O.reference_glsl_gather(ref, fp, gb, updated, visible, entries, d.shadow_atlas, s.cube_res)
        one = np.ones((d.H, d.W, 1), np.float16)
        gi_color = np.concatenate([O._to_half(color), one], -1)        # RGB16F attachment read back as (r, g, b, 1)
        gi_reservoir = O._to_half(res)
        last = den.frame(fp, gb, pgb, V, gi_color, gi_reservoir)[-1]
        cnt = stage2[:, 0]
        mask = np.arange(12)[None, :] < cnt[:, None]
        frames.append(dict(V=V, visible=visible.tolist(), entries=entries[:V].reshape(-1).tolist(), stage2_counts=cnt.tolist(),
                           stage2_indices=stage2[:, 1:][mask].tolist(), updated_sha=sha(updated[:V]), stage1_sha=sha(stage1),
                           gi_color_sha=sha(gi_color), gi_reservoir_sha=sha(gi_reservoir), combined_sha=sha(last["combined"]),
                           gi_illum_sha=sha(last["gi"]), history_sha=sha(last["history"])))
    return dict(workload="tiny", frames=frames)


if __name__ == "__main__":
    from oracle import oracle_py
    oracle_py.build()
    assert oracle_py.reference_glsl() is not None, "needs /root/reference (oracle/_ref/libref_glsl.so)"
    out = compute()
    out["made_by"] = "tests/golden/make_shader_text_golden.py: the reference's shader text compiled by oracle/make_ref_glsl.py (glm 0.9.8.5, GL fixed function stubbed per the GL 4.6 specification)"
    path = os.path.join(ROOT, "tests", "golden", "tiny_shader_text_golden.json")
    json.dump(out, open(path, "w"))
    print(path, os.path.getsize(path), "bytes")
