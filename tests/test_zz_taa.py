"""Widening row f4, second half — temporal anti-aliasing (taa.fs) on the device through sgi_taa, against the reference's own shader
text compiled by oracle/make_ref_glsl.py.  Green on a B200 since round 2 (profiles/r02a).  The kernel's logic
equals the shader text bit for bit on the host too (tests/test_taa_kernel_on_host.py); every operation is an IEEE one that CUDA rounds
as the CPU does, so the bar here is the same."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("workload,yaw", [("tiny", 1.5), ("c1", 3.0)])
def test_taa_against_the_compiled_shader_text(built, workload, yaw):
    from oracle import oracle_py as O
    from tests.parity_util import SceneData, make_pass, run_oracle
    from tests.test_taa_kernel_on_host import taa_inputs
    ref = O.reference_glsl()
    if ref is None:
        pytest.skip("oracle/_ref/libref_glsl.so did not travel with the repository")
    d = SceneData(workload, frames=2, yaw_deg_per_frame=yaw)
    screen, prev, vel = taa_inputs(d, run_oracle(d, keep_passes=False))
    want = O.reference_glsl_taa(ref, screen, prev, vel)
    gp = make_pass(d)
    out, hist = gp.taa(screen, prev, vel)
    out, hist = out.cpu().numpy(), hist.cpu().numpy()
    gp.close()
    assert np.array_equal(out.view(np.uint16), want.view(np.uint16))
    expect = np.rint(np.clip(out[..., :3].astype(np.float32), 0.0, 1.0) * np.float32(255.0)).astype(np.uint8)
    assert np.array_equal(hist[..., :3], expect)
