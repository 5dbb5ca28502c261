"""The reference code that compiles here (oracle/ref_shim.cpp: StratusMath.h and StratusGpuCommon.h of the reference tree,
unmodified, behind four empty third-party stubs) against the ABI and the oracle.  Skipped where oracle/_ref/libref_stratus.so
was never built (it is built wherever /root/reference exists and travels with the repository to the GPU box)."""
import ctypes as C

import numpy as np
import pytest

from stratusgfx_b200 import _abi


@pytest.fixture(scope="module")
def ref(built):
    from oracle import oracle_py
    L = oracle_py.reference_shim()
    if L is None:
        pytest.skip("oracle/_ref/libref_stratus.so not built (needs the reference tree)")
    return L


def test_wire_formats_and_limits_are_the_reference_headers_own(ref):
    """sizeof / offsetof / limits as the reference's compiler sees them (StratusGpuCommon.h:15-36, 192-232, 276-279), not restated numbers."""
    out = np.zeros(32, np.int32)
    k = ref.ref_layout(out.ctypes.data)
    got = out[:k].tolist()
    V, S1, S2, A = _abi.VplData, _abi.VplStage1PerTileOutputs, _abi.VplStage2PerTileOutputs, _abi.AtlasEntry
    want = [C.sizeof(V), V.position.offset, V.color.offset, V.specularPosition.offset, V.radius.offset, V.farPlane.offset, V.intensity.offset,
            C.sizeof(S1), S1.averageLocalNormal.offset, C.sizeof(S2), S2.indices.offset, C.sizeof(A), A.layer.offset,
            _abi.MAX_TOTAL_SHADOW_ATLASES, _abi.MAX_TOTAL_SHADOWS_PER_ATLAS, _abi.MAX_TOTAL_VPLS_BEFORE_CULLING, _abi.MAX_TOTAL_VPLS_PER_FRAME,
            _abi.MAX_VPLS_PER_TILE]
    assert got == want


def test_oracle_prepass_uses_the_reference_frustum_test_and_record_layout(ref):
    """Every light's in/out decision of the oracle's pre-pass equals stratus::IsSphereInFrustum called on the same floats, and the
    records it emits are the bytes InitVplFrameData_ produces with the reference's own GpuVplData / GpuVec."""
    from oracle import oracle_py as O
    from stratusgfx_b200.prepass import frustum_planes
    from stratusgfx_b200.scene import Scene
    from tests.test_oracle_selfchecks import _random_lights
    fp = Scene("tiny").frame_params(0)
    planes = np.ascontiguousarray(frustum_planes(list(fp.invProjectionView)), np.float32)
    rec = _random_lights(8000, 11)
    # push some spheres onto the planes: centres at exactly -radius from a plane are the boundary of `distance < -radius`
    vd, at, order = O.vpl_prepass(rec, list(fp.viewPosition), planes)
    inside_oracle = np.zeros(len(rec), bool)
    inside_oracle[order] = True
    pos = np.ascontiguousarray(rec["position"], np.float32)
    inside_ref = np.array([ref.ref_is_sphere_in_frustum(pos[i].ctypes.data, float(rec["radius"][i]), planes.ctypes.data) for i in range(len(rec))], bool)
    assert len(order) < 10000                      # the cap does not bite here, so membership is the frustum decision alone
    assert np.array_equal(inside_oracle, inside_ref) and 0 < inside_ref.sum() < len(rec)
    out = np.zeros(16, np.float32)
    for j in (0, 1, len(order) // 2, len(order) - 1):
        i = int(order[j])
        ref.ref_init_vpl_frame_data(pos[i].ctypes.data, float(rec["farPlane"][i]), float(rec["radius"][i]), float(rec["intensity"][i]), out.ctypes.data)
        assert np.array_equal(out.view(np.uint32), vd[j].view(np.uint32))


def test_reference_frustum_test_boundary_semantics(ref):
    """IsSphereInFrustum rejects only when distance < -radius: a sphere touching a plane from outside (distance == -radius) is kept.
    The oracle (and with it the CUDA kernel, tests/test_parity_gpu.py) must agree on exactly those cases."""
    from oracle import oracle_py as O
    from stratusgfx_b200.prepass import LIGHT_DTYPE
    planes = np.array([[1, 0, 0, 0], [-1, 0, 0, 100], [0, 1, 0, 0], [0, -1, 0, 100], [0, 0, 1, 0], [0, 0, -1, 100]], np.float32)   # the box [0,100]^3
    rec = np.zeros(6, LIGHT_DTYPE)
    rec["position"] = [[-2, 50, 50], [-2.0000002, 50, 50], [50, 102, 50], [50, 50, -1.9999999], [50, 50, 50], [103, 50, 50]]
    rec["radius"] = 2.0
    rec["farPlane"] = 2.0; rec["intensity"] = 1.0
    want = [ref.ref_is_sphere_in_frustum(np.ascontiguousarray(rec["position"][i]).ctypes.data, 2.0, planes.ctypes.data) for i in range(6)]
    assert want == [1, 0, 1, 1, 1, 0]
    vd, at, order = O.vpl_prepass(rec, [50, 50, 50], planes)
    assert sorted(order.tolist()) == [i for i in range(6) if want[i]]
