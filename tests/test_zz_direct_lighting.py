"""Widening row f3 — deferred direct lighting (pbr.fs) on the device, against the reference's own shader text compiled by
oracle/make_ref_glsl.py.  Green on a B200 since round 2 (profiles/r02a); the same body runs over the host-built library in the CPU suite."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("workload", ["tiny", "c1"])
def test_direct_lighting_against_the_compiled_shader_text(built, workload):
    from oracle import oracle_py as O
    from tests.parity_util import SceneData, compare_radiance, make_pass
    ref = O.reference_glsl()
    if ref is None:
        pytest.skip("oracle/_ref/libref_glsl.so did not travel with the repository")
    d = SceneData(workload, frames=1)
    s, fp, gb = d.scene, d.fp[0], d.gb[0]
    rng = np.random.default_rng(5)
    lights = np.zeros((3, 12), np.float32)
    lights[:, 0:3] = d.vpl_data[rng.integers(0, d.N, 3), 0:3]; lights[:, 3] = 1.0
    lights[:, 4:7] = rng.uniform(50, 400, (3, 3)); lights[:, 7] = 1.0; lights[:, 8] = 40.0; lights[:, 9] = 40.0
    shadow_lights = np.zeros((2, 12), np.float32)
    pick = rng.integers(0, d.N, 2)
    shadow_lights[:, 0:4] = d.vpl_data[pick, 0:4]; shadow_lights[:, 4:7] = rng.uniform(50, 400, (2, 3)); shadow_lights[:, 7] = 1.0
    shadow_lights[:, 8] = 500.0; shadow_lights[:, 9] = d.vpl_data[pick, 13]
    entries = d.atlas_by_vpl[pick]
    want = O.reference_glsl_direct_lighting(ref, fp, gb, d.csm[0], s.csm_res, d.shadow_atlas, s.cube_res, lights=lights, shadow_lights=shadow_lights,
                                            shadow_entries=entries, emission_strength=0.5, depth_bias=0.0)
    gp = make_pass(d)
    gp.set_gbuffer(gb, d.gb[-1]); gp.bind()
    got = gp.direct_lighting(fp, lights=lights, shadow_lights=shadow_lights, shadow_entries=entries, emission_strength=0.5).cpu().numpy()
    gp.close()
    with np.errstate(over="ignore"):
        want_h = np.concatenate([want.astype(np.float16), np.ones(want.shape[:2] + (1,), np.float16)], -1)
    rep = compare_radiance(got, want_h)
    assert rep["ok"], rep
    assert rep["bit_exact"] > 0.99
