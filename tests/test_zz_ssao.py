"""Widening row f4, first half — the two SSAO passes (ssao.fs, ssao_blur.fs) on the device through sgi_ssao, against the reference's
own shader text compiled by oracle/make_ref_glsl.py.  The kernels' logic equals the
shader text bit for bit on the host too (tests/test_ssao_kernel_on_host.py); every operation is an IEEE one that CUDA rounds as the
CPU does, so the bar is both R16F planes bit-exact (green on a B200 since round 2, profiles/r02a)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("workload", ["tiny", "c1"])
def test_ssao_against_the_compiled_shader_text(built, workload):
    from oracle import oracle_py as O
    from stratusgfx_b200.ssao import rotation_table, ssao_uniforms, structure_buffer
    from tests.parity_util import SceneData, make_pass
    ref = O.reference_glsl()
    if ref is None:
        pytest.skip("oracle/_ref/libref_glsl.so did not travel with the repository")
    d = SceneData(workload, frames=1)
    structure, rot, u = structure_buffer(d.fp[0], d.gb[0].depth), rotation_table(), ssao_uniforms(d.W, d.H)
    want_occ, want = O.reference_glsl_ssao(ref, structure, rot, **u)
    gp = make_pass(d)
    got = gp.ssao(structure, rot, **u).cpu().numpy()
    occ = gp.t["ssaoOcclusion"].cpu().numpy()
    gp.close()
    assert np.array_equal(occ.view(np.uint16), want_occ.view(np.uint16))
    assert np.array_equal(got.view(np.uint16), want.view(np.uint16))
