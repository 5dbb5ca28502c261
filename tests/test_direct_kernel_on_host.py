"""csrc/k_direct.cu (widening row f3) compiled by g++ (tests/cuda_on_host/host_tu.py: a translated copy, the source is untouched) and executed
thread by thread on the host, against the reference's own pbr.fs compiled by oracle/make_ref_glsl.py.  This checks the kernel's
logic — light loops, shadow paths by view distance, cascade choice, Burley BRDF, emissive, boundHDR — before any GPU time is spent;
what it cannot check is device-only behaviour (CUDA's powf / expf, launch geometry on real hardware)."""
import ctypes as C
import os

import numpy as np
import pytest

from stratusgfx_b200._abi import DirectParams, FrameParams

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host_kernel(built):
    from oracle import oracle_py
    ref = oracle_py.reference_glsl()
    if ref is None:
        pytest.skip("oracle/_ref/libref_glsl.so not built (needs the reference tree)")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    from tests.cuda_on_host.host_tu import compile_wrapper
    out = compile_wrapper("direct_on_host.cpp", "k_direct.cu", "libdirect_on_host.so")
    L = C.CDLL(out)
    L.host_direct_lighting.argtypes = [C.POINTER(FrameParams)] + [C.c_int] * 4 + [C.c_void_p] * 8 + [C.c_int, C.POINTER(DirectParams), C.c_void_p]
    return L, ref


@pytest.mark.parametrize("workload,sun,bias", [("tiny", True, 0.0), ("tiny", False, 0.0), ("c1", True, 0.002)])
def test_direct_kernel_logic_on_the_host_equals_the_shader_text(host_kernel, workload, sun, bias):
    from oracle import oracle_py as O
    from tests.parity_util import SceneData
    L, ref = host_kernel
    d = SceneData(workload, frames=1)
    s, fp, gb = d.scene, d.fp[0], d.gb[0]
    rng = np.random.default_rng(5)
    lights = np.zeros((3, 12), np.float32)
    lights[:, 0:3] = d.vpl_data[rng.integers(0, d.N, 3), 0:3]; lights[:, 3] = 1.0
    lights[:, 4:7] = rng.uniform(50, 400, (3, 3)); lights[:, 7] = 1.0; lights[:, 8] = 40.0; lights[:, 9] = 40.0
    pick = rng.integers(0, d.N, 2)
    shadow_lights = np.zeros((2, 12), np.float32)
    shadow_lights[:, 0:4] = d.vpl_data[pick, 0:4]; shadow_lights[:, 4:7] = rng.uniform(50, 400, (2, 3)); shadow_lights[:, 7] = 1.0
    shadow_lights[:, 8] = 500.0; shadow_lights[:, 9] = d.vpl_data[pick, 13]
    entries = np.ascontiguousarray(d.atlas_by_vpl[pick], np.int32)
    want = O.reference_glsl_direct_lighting(ref, fp, gb, d.csm[0], s.csm_res, d.shadow_atlas, s.cube_res, lights=lights, shadow_lights=shadow_lights,
                                            shadow_entries=entries, emission_strength=0.5, depth_bias=bias,
                                            sun_color=None if sun else [0.0, 0.0, 0.0])
    dp = DirectParams()
    dp.numLights, dp.numShadowLights, dp.sunEnabled = 3, 2, 1           # the compiled pbr.fs always has INFINITE_LIGHTING_ENABLED: "no sun" = black sun
    dp.nonShadowCasters, dp.shadowCasters, dp.shadowIndices = lights.ctypes.data, shadow_lights.ctypes.data, entries.ctypes.data
    dp.emissionStrength, dp.infiniteLightDepthBias, dp.infiniteLightZnear, dp.infiniteLightZfar, dp.pointLightAmbientIntensity = 0.5, bias, 1.0, 1000.0, 0.003
    out = np.zeros((d.H, d.W, 4), np.float16)
    dp.screenOut = out.ctypes.data
    fpk = fp.copy()
    if not sun:
        for i in range(3):
            fpk.infiniteLightColor[i] = 0.0
    tab = O._atlas_ptrs(d.shadow_atlas)
    planes = [np.ascontiguousarray(x) for x in (gb.depth, gb.normal, gb.albedo, gb.baseReflectivity, gb.roughnessMetallicAmbient, gb.ssao, d.csm[0])]
    n = L.host_direct_lighting(C.byref(fpk), d.W, d.H, s.cube_res, s.csm_res, *[x.ctypes.data for x in planes], C.addressof(tab), int(d.shadow_atlas.shape[0]),
                               C.byref(dp), out.ctypes.data)
    assert n == 1
    with np.errstate(over="ignore"):
        want_h = want.astype(np.float16)
    assert np.array_equal(out[..., :3].view(np.uint16), want_h.view(np.uint16))
    assert np.all(out[..., 3] == np.float16(1.0)) and float(want.max()) > 0.0
