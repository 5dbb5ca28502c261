"""csrc/k_taa.cu (widening row f4, second half) compiled by g++ (tests/cuda_on_host/host_tu.py) and executed
thread by thread on the host, against the reference's own taa.fs compiled by oracle/make_ref_glsl.py: the RGBA16F target must agree
bit for bit, and the history plane must be the GL float -> UNORM8 conversion of it."""
import ctypes as C
import os

import numpy as np
import pytest

from stratusgfx_b200._abi import TaaParams


def taa_inputs(d, orc):
    """Two consecutive LDR frames (what the FXAA target holds) from the lit frames of a turning camera, and the velocity plane."""
    def ldr(img):
        c = img[..., :3].astype(np.float32)
        c = c / (1.0 + c)                                              # any tone curve will do for test data
        out = np.full(img.shape[:2] + (4,), 255, np.uint8)
        out[..., :3] = np.clip(np.rint(c * 255.0), 0, 255).astype(np.uint8)
        return out
    return ldr(orc[1]["combined"]), ldr(orc[0]["combined"]), np.ascontiguousarray(d.gb[1].velocity)


@pytest.fixture(scope="module")
def host_taa(built):
    from oracle import oracle_py
    ref = oracle_py.reference_glsl()
    if ref is None:
        pytest.skip("oracle/_ref/libref_glsl.so not built (needs the reference tree)")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    from tests.cuda_on_host.host_tu import compile_wrapper
    L = C.CDLL(compile_wrapper("taa_on_host.cpp", "k_taa.cu", "libtaa_on_host.so"))
    L.host_taa.argtypes = [C.c_int, C.c_int, C.POINTER(TaaParams)]
    return L, ref


@pytest.mark.parametrize("workload,yaw", [("tiny", 1.5), ("c1", 3.0)])
def test_taa_kernel_on_the_host_equals_the_shader_text(host_taa, workload, yaw):
    from oracle import oracle_py as O
    from tests.parity_util import SceneData, run_oracle
    L, ref = host_taa
    d = SceneData(workload, frames=2, yaw_deg_per_frame=yaw)
    orc = run_oracle(d, keep_passes=False)
    screen, prev, vel = taa_inputs(d, orc)
    assert np.abs(vel.astype(np.float32)).max() > 0                   # the camera turns: the history read is a filtered one
    want = O.reference_glsl_taa(ref, screen, prev, vel)
    out, hist = np.zeros((d.H, d.W, 4), np.float16), np.zeros((d.H, d.W, 4), np.uint8)
    tp = TaaParams()
    tp.screen, tp.prevScreen, tp.velocity, tp.colorOut, tp.historyOut = screen.ctypes.data, prev.ctypes.data, vel.ctypes.data, out.ctypes.data, hist.ctypes.data
    assert L.host_taa(d.W, d.H, C.byref(tp)) == 2          # the table kernel + the resolve
    assert np.array_equal(out.view(np.uint16), want.view(np.uint16))
    assert (out[..., 3] == np.float16(0.0625)).all() and np.isfinite(out.astype(np.float32)).all()
    # the blit into the RGB8 history: clamp, scale, round to nearest
    expect = np.rint(np.clip(out[..., :3].astype(np.float32), 0.0, 1.0) * np.float32(255.0)).astype(np.uint8)
    assert np.array_equal(hist[..., :3], expect) and (hist[..., 3] == 255).all()
    # and it behaves like a temporal filter: the output lies between the current frame and the clamped history
    cur = screen[..., :3].astype(np.float32) / 255.0
    assert np.abs(out[..., :3].astype(np.float32) - cur).mean() < 0.05
