"""csrc/k_denoise.cu compiled by g++ (tests/cuda_on_host/host_tu.py: a translated copy, the source is untouched) and executed thread
by thread on the host — the shipped kernels' LOGIC (the clamp-free interior path and the clamped edge path of the à-trous
kernel, the final pass's temporal epilogue) against the CPU oracle without a GPU.

The MUFU approximations (lg2/ex2) are libm calls here, so radiance is compared under the parity bar of tests/parity_util.py
(1e-3 relative per pixel or 60 dB PSNR) like on the GPU; the history plane — the one decision of this shader — must be exact."""
import ctypes as C
import os

import numpy as np
import pytest

from stratusgfx_b200._abi import FrameParams


def _load(defines=(), name="libdenoise_on_host.so"):
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    from tests.cuda_on_host.host_tu import compile_wrapper
    L = C.CDLL(compile_wrapper("denoise_on_host.cpp", "k_denoise.cu", name, defines))
    L.host_denoise.argtypes = [C.POINTER(FrameParams)] + [C.c_int] * 3 + [C.c_void_p] * 17
    return L


@pytest.fixture(scope="module")
def host_denoiser(built):
    return _load()


def _run(L, data, orc, f, iters):
    gb, pgb = data.gb[f], data.gb[f - 1]
    H, W = data.H, data.W
    prev_gi = orc[f - 1]["gi_illum"] if f > 0 else np.zeros((H, W, 4), np.float16)
    prev_hist = orc[f - 1]["history"] if f > 0 else np.zeros((H, W), np.float16)
    passes = np.zeros((iters, H, W, 4), np.float16)
    combined, gi, hist = np.zeros((H, W, 4), np.float16), np.zeros((H, W, 4), np.float16), np.zeros((H, W), np.float16)
    p = lambda a: np.ascontiguousarray(a).ctypes.data
    keep = [np.ascontiguousarray(a) for a in (gb.normal, gb.depth, gb.screen, gb.velocity, gb.id, pgb.normal, pgb.id, pgb.depth,
                                              orc[f]["gi_color"], orc[f]["gi_reservoir"], prev_gi, prev_hist, orc[f]["visible"])]
    rc = L.host_denoise(C.byref(data.fp[f]), W, H, iters, *[a.ctypes.data for a in keep], p(passes), p(combined), p(gi), p(hist))
    assert rc == 0
    return passes, combined, gi, hist


@pytest.mark.parametrize("yaw", [0.0, 1.5])
def test_denoiser_kernels_on_the_host_match_the_oracle(host_denoiser, yaw):
    """Static camera (every reprojected tap on its own texel) and a turning one (bilinear reprojection, history resets)."""
    from tests.parity_util import SceneData, compare_radiance, run_oracle
    d = SceneData("tiny", frames=2, yaw_deg_per_frame=yaw)
    iters = d.scene.denoise_iters
    orc = run_oracle(d)
    for f in range(2):
        passes, combined, gi, hist = _run(host_denoiser, d, orc, f, iters)
        for i in range(iters):
            r = compare_radiance(passes[i], orc[f]["pass_reservoirs"][i], 3)
            assert r["ok"] and r["bit_exact"] > 0.99, (f, i, r)
        for got, want in ((combined, orc[f]["combined"]), (gi, orc[f]["gi_illum"])):
            r = compare_radiance(got, want, 3)
            assert r["ok"] and r["bit_exact"] > 0.99, (f, r)
        assert np.array_equal(hist.view(np.uint16), orc[f]["history"].view(np.uint16))
    if yaw:
        assert len(np.unique(orc[1]["history"].view(np.uint16))) > 1          # some pixels were reset, some accumulated
