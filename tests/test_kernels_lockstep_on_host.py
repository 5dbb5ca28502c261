"""The kernels of csrc/ that use barriers, shared memory and warp collectives — pre-pass ranking, visibility cull with ordered
compaction, colouring, both tile-list stages, the persistent stochastic gather — compiled by g++ and executed on the host in
LOCKSTEP mode (tests/cuda_on_host/cuda_host_shim.h: one OS thread per CUDA thread of a CTA, real barriers / shuffles / votes,
__shared__ as CTA-local storage), against the CPU oracle.  Same bars as the GPU parity tests: integer outputs bit-exact, radiance
within the bar of tests/parity_util.py.  This is a check of the kernels' LOGIC that needs no GPU; the GPU tests remain the parity
tests proper (hardware scheduling, MUFU approximations and memory ordering exist only there)."""
import ctypes as C
import os

import numpy as np
import pytest

from stratusgfx_b200._abi import FrameParams, PrepassParams


def _lib(name, wrappers, kernels):
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    from tests.cuda_on_host.host_tu import compile_wrapper
    return C.CDLL(compile_wrapper(wrappers, kernels, name))


@pytest.fixture(scope="module")
def host_prepass(built):
    L = _lib("libprepass_on_host.so", ("prepass_on_host.cpp",), ("k_prepass.cu",))
    L.host_vpl_prepass.argtypes = [C.POINTER(PrepassParams)]
    return L


def _run_prepass(L, rec, view, planes):
    n = int(rec.shape[0])
    cap = max(n, 1)
    vd, at, order, count = np.zeros((cap, 16), np.float32), np.zeros((cap, 2), np.int32), np.zeros(cap, np.int32), np.zeros(1, np.int32)
    rec = np.ascontiguousarray(rec)
    pp = PrepassParams()
    for i in range(3):
        pp.viewPosition[i] = float(view[i])
    fr = np.asarray(planes, np.float32).reshape(6, 4)
    for k in range(6):
        for i in range(4):
            pp.frustumPlanes[k][i] = float(fr[k, i])
    pp.numLights = n
    pp.lights, pp.vplDataOut, pp.atlasEntryOut, pp.orderOut, pp.countOut = rec.ctypes.data, vd.ctypes.data, at.ctypes.data, order.ctypes.data, count.ctypes.data
    L.host_vpl_prepass(C.byref(pp))
    c = int(count[0])
    return vd[:c], at[:c], order[:c]


@pytest.mark.parametrize("n,seed", [(0, 1), (1, 2), (257, 3), (2500, 5)])
def test_prepass_kernels_in_lockstep_equal_the_oracle(host_prepass, n, seed):
    from oracle import oracle_py as O
    from tests.test_parity_gpu import _prepass_inputs
    rec, view, planes = _prepass_inputs(n, seed)
    vd, at, order = O.vpl_prepass(rec, view, planes)
    gvd, gat, gorder = _run_prepass(host_prepass, rec, view, planes)
    assert len(gorder) == len(order) and (n < 257 or len(order) > 0)
    assert np.array_equal(gorder, order) and np.array_equal(gat, at)
    assert np.array_equal(gvd.view(np.uint32), vd.view(np.uint32))


def test_prepass_frustum_boundary_cases_in_lockstep(host_prepass):
    """tests/test_parity_gpu.py::test_vpl_prepass_frustum_boundary_cases on the host: a sphere touching a plane from outside is kept,
    one float ulp further out it is dropped."""
    from oracle import oracle_py as O
    from stratusgfx_b200.prepass import LIGHT_DTYPE
    planes = np.array([[1, 0, 0, 0], [-1, 0, 0, 100], [0, 1, 0, 0], [0, -1, 0, 100], [0, 0, 1, 0], [0, 0, -1, 100]], np.float32)
    rec = np.zeros(6, LIGHT_DTYPE)
    rec["position"] = [[-2, 50, 50], [-2.0000002, 50, 50], [50, 102, 50], [50, 50, -1.9999999], [50, 50, 50], [103, 50, 50]]
    rec["radius"] = 2.0; rec["farPlane"] = 2.0; rec["intensity"] = 1.0
    _, _, order = O.vpl_prepass(rec, [50, 50, 50], planes)
    assert sorted(order.tolist()) == [0, 2, 3, 4]
    _, _, got = _run_prepass(host_prepass, rec, [50, 50, 50], planes)
    assert np.array_equal(got, order)
