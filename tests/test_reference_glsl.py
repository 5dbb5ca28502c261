"""The reference's own GLSL, compiled: oracle/make_ref_glsl.py turns the texture-free functions of pbr2.glsl, pbr.glsl, fog.glsl and
common.glsl that the VPL gather calls into C++ against the reference's vendored glm (oracle/_ref/libref_glsl.so).  The oracle's
restatement of the same functions must agree with that shader text bit for bit.  Skipped where the library was never built
(it is built wherever /root/reference exists and travels with the repository)."""
import ctypes as C

import numpy as np
import pytest

from stratusgfx_b200._abi import FrameParams


@pytest.fixture(scope="module")
def libs(built):
    from oracle import oracle_py
    ref = oracle_py.reference_glsl()
    if ref is None:
        pytest.skip("oracle/_ref/libref_glsl.so not built (needs the reference tree)")
    return oracle_py.lib(), ref


def _bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


def test_rng_of_the_shader_text(libs):
    """hash(uvec3) / floatConstruct / random(vec3) (common.glsl:183-214): pure integer arithmetic on the float bits."""
    orc, ref = libs
    rng = np.random.default_rng(1)
    for _ in range(2000):
        x, y, t = np.float32(rng.integers(0, 7680) + 0.5), np.float32(rng.integers(0, 4320) + 0.5), np.float32(1000.0 + 16.0 * rng.integers(0, 100000))
        assert orc.orc_hash3(x, y, t) == ref.refglsl_hash3(x, y, t)
        assert _bits(orc.orc_random3(x, y, t)) == _bits(ref.refglsl_random3(x, y, t))


def test_world_position_from_depth_of_the_shader_text(libs):
    """worldPositionFromDepth (common.glsl:143-155) on the scene's own inverse projection-view matrices."""
    from stratusgfx_b200.scene import Scene
    orc, ref = libs
    rng = np.random.default_rng(2)
    for wl in ("tiny", "c3"):
        fp = Scene(wl, yaw_deg_per_frame=0.25).frame_params(3)
        m = np.ascontiguousarray(list(fp.invProjectionView), np.float32)
        a, b = np.zeros(3, np.float32), np.zeros(3, np.float32)
        for _ in range(2000):
            u, v, d = np.float32(rng.random()), np.float32(rng.random()), np.float32(rng.random())
            orc.orc_test_world_pos(u, v, d, m.ctypes.data, a.ctypes.data)
            ref.refglsl_world_pos(u, v, d, m.ctypes.data, b.ctypes.data)
            assert np.array_equal(_bits(a), _bits(b))


def test_vpl_attenuation_of_the_shader_text(libs):
    orc, ref = libs
    rng = np.random.default_rng(3)
    for _ in range(4000):
        d = np.ascontiguousarray(rng.normal(0, 200, 3), np.float32)
        r = np.float32(rng.choice([5.0, 50.0, 500.0]))
        assert _bits(orc.orc_test_attenuation(d.ctypes.data, r)) == _bits(ref.refglsl_attenuation(d.ctypes.data, r))
    assert ref.refglsl_bound_hdr(1e6) == 65504.0 and ref.refglsl_bound_hdr(3.5) == 3.5


@pytest.mark.parametrize("perceptual,min_rough,fog_density", [(1, 0.08, 0.0), (0, 0.08, 0.0), (1, 0.3, 0.002)])
def test_virtual_point_lighting_of_the_shader_text(libs, perceptual, min_rough, fog_density):
    """calculateVirtualPointLighting2 (pbr2.glsl:613-657) and everything under it — calculateLighting_Lambert_SeparateDiffuseSpecular,
    BRDF_Lambert_SeparateDiffuseSpecular, BRDF_Base, BRDF_DotProducts, singleScatteringBRDF_Specular, NDF_GGX, Visibility_G_SmithGGX,
    Fr_Fresnel_Schlick, RemapRoughness, vplDiffuseAttenuation, applyFog — on random surfaces and lights: the oracle's restatement must
    return the bytes the shader text returns."""
    orc, ref = libs
    fp = FrameParams()
    fp.usePerceptualRoughness, fp.minRoughness, fp.fogDensity = perceptual, min_rough, fog_density
    fog = np.array([0.5, 0.6, 0.7], np.float32)
    for i in range(3):
        fp.fogColor[i] = float(fog[i])
    ref.refglsl_set_uniforms(perceptual, min_rough, fog.ctypes.data, fog_density)
    rng = np.random.default_rng(4)
    a, o1, o2 = np.zeros(30, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32)
    nonzero = 0
    for _ in range(5000):
        unit = lambda: (lambda v: v / np.linalg.norm(v))(rng.normal(0, 1, 3))
        a[0:3] = rng.uniform(-100, 100, 3)                     # fragPosition
        a[3:6] = np.maximum(rng.integers(0, 256, 3) / 255.0, 0.01)   # baseColor, floored as the gather does (vpl_pbr_gi.fs:118)
        a[6:9] = unit(); a[9:12] = unit()                       # normal, viewDir
        a[12:15] = a[0:3] + rng.normal(0, 60, 3); a[15:18] = a[12:15] + rng.normal(0, 3, 3)    # specular / diffuse light position
        a[18:21] = rng.uniform(0, 30, 3)                        # lightColor
        a[21] = rng.uniform(1, 900); a[22] = rng.choice([50.0, 500.0])       # viewDist, lightRadius
        a[23] = max(0.5, rng.integers(0, 256) / 255.0); a[24] = rng.choice([0.0, 0.0, rng.random()])   # roughness, metallic
        a[25] = np.clip(rng.random(), 0.35, 1.0); a[26] = rng.choice([0.0, 1.0, rng.random()])        # ambient occlusion, shadow factor
        a[27:30] = rng.integers(0, 256) / 255.0                 # baseReflectance (vec3(r))
        orc.orc_test_vpl_lighting(C.addressof(fp), a.ctypes.data, o1.ctypes.data)
        ref.refglsl_vpl_lighting(a.ctypes.data, o2.ctypes.data)
        assert np.array_equal(_bits(o1), _bits(o2)), (a, o1, o2)
        nonzero += int(np.any(o2 != 0))
    assert nonzero > 1000                                        # (lights behind the surface or fully shadowed give exactly 0 in both)


@pytest.mark.parametrize("workload,frame", [("tiny", 0), ("tiny", 1), ("c1", 0)])
def test_whole_gather_shader_text_against_the_oracle(libs, workload, frame):
    """vpl_pbr_gi.fs as a whole: `performLightingCalculations` — the sample budget, the draw / re-draw loop with `--i; continue;`,
    the cube-shadow test, the occlusion floor and re-weighting, the BRDF chain, `boundHDR` — compiled from the reference's shader
    text and run over a full frame.  After the RGB16F / RGBA16F store (IEEE round-to-nearest-even) its two outputs must be the
    oracle's planes bit for bit: every light choice, every shadow verdict and every radiance value of every pixel."""
    from oracle.oracle_py import OracleFrame, reference_glsl_gather
    from tests.parity_util import SceneData
    orc, ref = libs
    d = SceneData(workload, frames=frame + 1)
    s = d.scene
    o = OracleFrame(d.W, d.H, s.cube_res, s.csm_res, s.denoise_iters)
    fp, gb = d.fp[frame], d.gb[frame]
    o.cull(fp, d.vpl_data, d.atlas_by_vpl, d.csm[frame])
    o.color(fp, d.shadow_atlas, d.diffuse_atlas)
    o.gather(fp, gb)
    color, res = reference_glsl_gather(ref, fp, gb, o.updated, o.visible, o.entries, d.shadow_atlas, s.cube_res)
    with np.errstate(over="ignore"):
        color_h, res_h = color.astype(np.float16), res.astype(np.float16)
    assert np.array_equal(color_h.view(np.uint16), o.gi_color[..., :3].view(np.uint16))
    assert np.array_equal(res_h.view(np.uint16), o.gi_reservoir.view(np.uint16))
    assert float(res[..., :3].max()) > 0.0 and 1.0 <= float(res[..., 3].min()) and float(res[..., 3].max()) <= 6.0


@pytest.mark.parametrize("workload,frames,yaw", [("tiny", 3, 0.0), ("tiny", 3, 0.25), ("c1", 1, 0.0)])
def test_whole_denoiser_shader_text_against_the_oracle(libs, workload, frames, yaw):
    """vpl_pbr_gi_denoise.fs as a whole — computeMergedReservoir (pass 0), the 25-tap filter with filterInput's pow/exp weights
    (à-trous passes), the reprojection, the history reset decision, the temporal mix and the re-modulation (final pass) — compiled
    from the reference's shader text and driven through the reference's host loop (BE:1879-1952) over several frames, static and
    moving camera.  After the 16F stores, every pass's reservoir and the final pass's combined colour, GI and history planes must
    equal the oracle's bit for bit."""
    from oracle.oracle_py import OracleFrame, ReferenceGlslDenoiser
    from tests.parity_util import SceneData
    orc, ref = libs
    d = SceneData(workload, frames=frames, yaw_deg_per_frame=yaw)
    s = d.scene
    o = OracleFrame(d.W, d.H, s.cube_res, s.csm_res, s.denoise_iters)
    rd = ReferenceGlslDenoiser(ref, d.W, d.H, s.denoise_iters)
    resets = 0
    for f in range(frames):
        fp, gb, pgb = d.fp[f], d.gb[f], d.gb[f - 1]
        o.cull(fp, d.vpl_data, d.atlas_by_vpl, d.csm[f]); o.color(fp, d.shadow_atlas, d.diffuse_atlas); o.gather(fp, gb)
        passes = rd.frame(fp, gb, pgb, int(o.visible[0]), o.gi_color, o.gi_reservoir)      # same gather output into both denoisers
        o.denoise(fp, gb, pgb, keep_passes=True)
        for i, fbo in enumerate(passes):
            assert np.array_equal(fbo["reservoir"].view(np.uint16), o.pass_reservoirs[i].view(np.uint16)), (f, i)
        last = passes[-1]
        assert np.array_equal(last["history"].view(np.uint16), o.history.view(np.uint16)), f
        assert np.array_equal(last["gi"][..., :3].view(np.uint16), o.gi_illum[..., :3].view(np.uint16)), f
        assert np.array_equal(last["combined"][..., :3].view(np.uint16), o.combined[..., :3].view(np.uint16)), f
        resets += int(np.sum(last["history"] == np.float16(1.0))) if f > 0 else 0
    if yaw != 0.0:
        assert resets > 0                      # a moving camera resets the history somewhere (disocclusion): the decision path ran
    elif frames > 1:
        assert resets == 0 and float(o.history.max()) == float(frames)


# ---- the compute shaders of the path, compiled from the shader text ----------------------------------------------------------
@pytest.mark.parametrize("workload", ["tiny", "c1", "c3"])
def test_cull_shader_text_against_the_oracle(libs, workload):
    """viscull_vpls.cs main() run as one work group of 1024 real threads with a real barrier(), with calculateInfiniteShadowValue /
    sampleShadowTexture / computeTexelSize from pbr.glsl and common.glsl.  The visible SET and count must equal the oracle's; the
    shader's eight helper invocations race for the slots (atomicAdd), so its order is one interleaving and the oracle's ascending
    order another (the shader's own serial variant, :108-119); each slot must hold the VPL its index names."""
    from oracle.oracle_py import OracleFrame, reference_glsl_cull
    from tests.parity_util import SceneData
    orc, ref = libs
    d = SceneData(workload, frames=1)
    s = d.scene
    o = OracleFrame(d.W, d.H, s.cube_res, s.csm_res, s.denoise_iters)
    V = o.cull(d.fp[0], d.vpl_data, d.atlas_by_vpl, d.csm[0])
    updated, visible = reference_glsl_cull(ref, d.fp[0], d.vpl_data, d.csm[0], s.csm_res)
    assert int(visible[0]) == V and 0 < V < d.N
    assert sorted(visible[1:V + 1].tolist()) == o.visible[1:V + 1].tolist()
    assert np.array_equal(updated[:V].view(np.uint32), d.vpl_data[visible[1:V + 1]].view(np.uint32))


@pytest.mark.parametrize("workload", ["tiny", "c1"])
def test_color_and_tile_shader_texts_against_the_oracle(libs, workload):
    """vpl_light_color.cs, vpl_tiled_deferred_culling_stage1.cs (25 invocations per tile, shared arrays, invocation 0 sums in index
    order) and ..._stage2.cs (the insertion with SHUFFLE_DOWN) from the shader text: the coloured table, the per-tile means and the
    per-tile lists must be the oracle's bit for bit."""
    from oracle.oracle_py import OracleFrame, reference_glsl_color, reference_glsl_stage1, reference_glsl_stage2
    from tests.parity_util import SceneData
    orc, ref = libs
    d = SceneData(workload, frames=1)
    s = d.scene
    fp, gb = d.fp[0], d.gb[0]
    o = OracleFrame(d.W, d.H, s.cube_res, s.csm_res, s.denoise_iters)
    V = o.cull(fp, d.vpl_data, d.atlas_by_vpl, d.csm[0])
    before = o.updated.copy()
    o.color(fp, d.shadow_atlas, d.diffuse_atlas)
    colored = reference_glsl_color(ref, fp, before, o.visible, o.entries, d.shadow_atlas, d.diffuse_atlas, s.cube_res)
    assert np.array_equal(colored[:V].view(np.uint32), o.updated[:V].view(np.uint32))
    assert not np.array_equal(colored[:V, 4:12], before[:V, 4:12])                      # colour and specular position were written
    o.tile_lists(fp, gb, d.vpl_data)
    s1 = reference_glsl_stage1(ref, fp, gb)
    assert np.array_equal(s1.view(np.uint32), o.stage1.view(np.uint32))
    s2 = reference_glsl_stage2(ref, fp, d.W, d.H, d.vpl_data, o.stage1, o.visible)
    cnt = o.stage2[:, 0]
    assert np.array_equal(s2[:, 0], cnt) and cnt.max() <= 12 and cnt.max() > 0
    mask = np.arange(12)[None, :] < cnt[:, None]
    assert np.array_equal(s2[:, 1:][mask], o.stage2[:, 1:][mask])


@pytest.mark.parametrize("num_vpls", [1, 2, 5])
def test_gather_and_denoiser_shader_text_with_very_few_vpls(libs, num_vpls):
    """One and two VPLs: `maxRandomIndex = numVisible[0] - 1` is 0 or 1, so every draw picks VPL 0 (the last VPL is never drawn,
    vpl_pbr_gi.fs:146-151) — the degenerate ends of the draw loop must match the shader text as well."""
    from oracle.oracle_py import OracleFrame, ReferenceGlslDenoiser, reference_glsl_gather
    from tests.parity_util import SceneData
    orc, ref = libs
    d = SceneData("tiny", frames=1, num_vpls=num_vpls)
    s = d.scene
    o = OracleFrame(d.W, d.H, s.cube_res, s.csm_res, s.denoise_iters)
    fp, gb = d.fp[0], d.gb[0]
    V = o.cull(fp, d.vpl_data, d.atlas_by_vpl, d.csm[0])
    if V == 0:
        pytest.skip("no VPL of this tiny set is visible to the sun")
    o.color(fp, d.shadow_atlas, d.diffuse_atlas); o.gather(fp, gb)
    color, res = reference_glsl_gather(ref, fp, gb, o.updated, o.visible, o.entries, d.shadow_atlas, s.cube_res)
    with np.errstate(over="ignore"):
        assert np.array_equal(res.astype(np.float16).view(np.uint16), o.gi_reservoir.view(np.uint16))
        assert np.array_equal(color.astype(np.float16).view(np.uint16), o.gi_color[..., :3].view(np.uint16))
    passes = ReferenceGlslDenoiser(ref, d.W, d.H, s.denoise_iters).frame(fp, gb, d.gb[-1], V, o.gi_color, o.gi_reservoir)
    o.denoise(fp, gb, d.gb[-1], keep_passes=True)
    assert np.array_equal(passes[-1]["combined"][..., :3].view(np.uint16), o.combined[..., :3].view(np.uint16))


def test_direct_lighting_shader_text_runs_and_behaves(libs):
    """pbr.fs (widening row f3, the pass that produces the GI path's `screen` input) compiled from the shader text: no CUDA
    counterpart exists yet, so this only fixes its behaviour — emissive pass-through with every light off, linearity in the sun's
    colour, sun shadows darkening, and a point light adding energy around itself."""
    from oracle.oracle_py import reference_glsl_direct_lighting
    from tests.parity_util import SceneData
    orc, ref = libs
    d = SceneData("tiny", frames=1)
    s, fp, gb = d.scene, d.fp[0], d.gb[0]
    common = dict(csm=d.csm[0], csm_res=s.csm_res, shadow_atlas=d.shadow_atlas, cube_res=s.cube_res)
    # every light off: fsColor = boundHDR(emissive * emissionStrength), emissive = (albedo.a, baseReflectivity.g, rma.b)
    off = reference_glsl_direct_lighting(ref, fp, gb, sun_color=[0, 0, 0], emission_strength=2.0, **common)
    emissive = np.stack([gb.albedo[..., 3] / np.float32(255), gb.baseReflectivity[..., 1] / np.float32(255),
                         gb.roughnessMetallicAmbient[..., 2] / np.float32(255)], -1).astype(np.float32)
    assert np.array_equal(off, emissive * np.float32(2.0))
    sun1 = reference_glsl_direct_lighting(ref, fp, gb, **common)
    sun2 = reference_glsl_direct_lighting(ref, fp, gb, sun_color=[2 * c for c in fp.infiniteLightColor], **common)
    assert np.isfinite(sun1).all() and sun1.min() >= 0.0 and sun1.max() > 0.0
    assert np.allclose(sun2, 2.0 * sun1, rtol=1e-5, atol=1e-7)                       # the sun's colour enters linearly
    lit = sun1.sum(-1) > 0
    assert 0.02 < lit.mean() < 0.98                                                   # some pixels are lit by the sun, some are in its shadow
    light = np.zeros((1, 12), np.float32)
    light[0, 0:4] = [0.0, 15.0, 0.0, 1.0]; light[0, 4:8] = [400.0, 400.0, 400.0, 1.0]; light[0, 8] = 40.0; light[0, 9] = 40.0
    withp = reference_glsl_direct_lighting(ref, fp, gb, lights=light, **common)
    assert (withp >= sun1 - 1e-6).all() and float((withp - sun1).sum()) > 0.0


def test_sun_shadow_value_of_the_shader_text(libs):
    """calculateInfiniteShadowValue (pbr.glsl:158-227: cascade choice and blend, 3x3 PCF taps on two cascades) as a VALUE, not only
    as the cull's visible / hidden decision: random points of the scene, sun direction and a tilted normal."""
    from tests.parity_util import SceneData
    orc, ref = libs
    orc.orc_test_sun_shadow.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    orc.orc_test_sun_shadow.restype = C.c_float
    ref.refglsl_sun_shadow.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p]
    ref.refglsl_sun_shadow.restype = C.c_float
    rng = np.random.default_rng(9)
    for wl in ("c1", "c3"):
        d = SceneData(wl, frames=1)
        fp, csm = d.fp[0], np.ascontiguousarray(d.csm[0], np.float32)
        sun = np.ascontiguousarray(list(fp.infiniteLightDirection), np.float32)
        cpv = np.ascontiguousarray([list(m) for m in fp.cascadeProjViews], np.float32).reshape(-1)
        planes = np.ascontiguousarray([list(p) for p in fp.cascadePlanes], np.float32).reshape(-1)
        values = []
        for k in range(3000):
            p = np.ascontiguousarray(d.vpl_data[rng.integers(0, d.N), :3] + rng.normal(0, 6, 3), np.float32)
            n = sun if k % 2 == 0 else np.ascontiguousarray((lambda v: v / np.linalg.norm(v))(rng.normal(0, 1, 3)), np.float32)
            a = orc.orc_test_sun_shadow(C.addressof(fp), csm.ctypes.data, d.scene.csm_res, p.ctypes.data, n.ctypes.data)
            b = ref.refglsl_sun_shadow(sun.ctypes.data, cpv.ctypes.data, planes.ctypes.data, csm.ctypes.data, d.scene.csm_res, p.ctypes.data, n.ctypes.data)
            assert _bits(a) == _bits(b), (wl, k, a, b)
            values.append(b)
        values = np.array(values)
        assert (values == 0).any() and (values == 1).any() and ((values > 0) & (values < 1)).any()      # shadowed, lit and penumbra all occur


def test_cube_map_fragment_stage_equals_the_shader_text(libs):
    """Widening row f2: the fragment stage of the VPL cube-map pass — shadowVpl.fs main(), compiled from the shader text — against
    the oracle's restatement (oracle.cpp::shadowVplDepth), bit for bit on random fragments, the far-plane clamp included; and the
    colour output is the material's base colour unchanged."""
    from oracle import oracle_py as O
    _, ref = libs
    rng = np.random.default_rng(7)
    ref.refglsl_shadow_vpl.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    for k in range(4000):
        light = rng.uniform(-200, 200, 3).astype(np.float32)
        far = np.float32(rng.choice([50.0, 500.0, 123.456]))
        frag = (light + rng.normal(0, 1, 3) * rng.choice([0.01, 1.0, 40.0, 400.0, 2000.0])).astype(np.float32)
        base = rng.uniform(0, 1, 4).astype(np.float32)
        depth, colour = np.zeros(1, np.float32), np.zeros(3, np.float32)
        ref.refglsl_shadow_vpl(frag.ctypes.data, light.ctypes.data, float(far), base.ctypes.data, depth.ctypes.data, colour.ctypes.data)
        want = np.float32(O.shadow_vpl_depth(frag, light, float(far)))
        assert depth[0].view(np.uint32) == want.view(np.uint32), (k, depth[0], want)
        assert np.array_equal(colour, base[:3])
        assert 0.0 <= depth[0] <= 1.0
