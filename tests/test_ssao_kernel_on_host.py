"""csrc/k_ssao.cu (widening row f4, first half) compiled by g++ (tests/cuda_on_host/host_tu.py) and executed
thread by thread on the host, against the reference's own ssao.fs and ssao_blur.fs compiled by oracle/make_ref_glsl.py.  Both
passes are exact-tier arithmetic, so the two R16F planes must agree bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

from stratusgfx_b200._abi import SsaoParams


@pytest.fixture(scope="module")
def host_ssao(built):
    from oracle import oracle_py
    ref = oracle_py.reference_glsl()
    if ref is None:
        pytest.skip("oracle/_ref/libref_glsl.so not built (needs the reference tree)")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    from tests.cuda_on_host.host_tu import compile_wrapper
    L = C.CDLL(compile_wrapper("ssao_on_host.cpp", "k_ssao.cu", "libssao_on_host.so"))
    L.host_ssao.argtypes = [C.c_int, C.c_int, C.POINTER(SsaoParams)]
    return L, ref


def test_rotation_table_is_the_engine_table():
    from stratusgfx_b200.ssao import rotation_table
    t = rotation_table().astype(np.float32)
    ang = np.round(np.arctan2(t[..., 1], t[..., 0]) / (2 * np.pi / 16)).astype(int) % 16
    assert sorted(ang.ravel().tolist()) == list(range(16))          # a permutation of the 16 angles (BE:759-779)
    assert np.allclose(t[..., 0] ** 2 + t[..., 1] ** 2, 1.0, atol=2e-3) and not t[..., 2:].any()
    assert ang.ravel().tolist() != list(range(16))                   # shuffled


@pytest.mark.parametrize("workload,yaw", [("tiny", 0.0), ("c1", 2.0)])
def test_ssao_kernels_on_the_host_equal_the_shader_text(host_ssao, workload, yaw):
    from oracle import oracle_py as O
    from stratusgfx_b200.ssao import rotation_table, ssao_uniforms, structure_buffer
    from tests.parity_util import SceneData
    L, ref = host_ssao
    d = SceneData(workload, frames=1, yaw_deg_per_frame=yaw)
    structure = structure_buffer(d.fp[0], d.gb[0].depth)
    rot = rotation_table()
    u = ssao_uniforms(d.W, d.H)
    want_occ, want = O.reference_glsl_ssao(ref, structure, rot, **u)
    occ, out = np.zeros((d.H, d.W), np.float16), np.zeros((d.H, d.W), np.float16)
    sp = SsaoParams()
    sp.aspectRatio, sp.projPlaneZDist, sp.intensity = float(u["aspectRatio"]), float(u["projPlaneZDist"]), float(u["intensity"])
    sp.structureBuffer, sp.rotationLookup, sp.occlusionScratch, sp.ssaoOut = structure.ctypes.data, rot.ctypes.data, occ.ctypes.data, out.ctypes.data
    assert L.host_ssao(d.W, d.H, C.byref(sp)) == 2
    # the planes carry information: occluded creases and open surfaces
    w = want.astype(np.float32)
    assert np.isfinite(w).all() and w.max() <= 1.0 and (w < 0.9).mean() > 0.001 and (w > 0.99).mean() > 0.05
    assert np.array_equal(occ.view(np.uint16), want_occ.view(np.uint16))
    assert np.array_equal(out.view(np.uint16), want.view(np.uint16))
