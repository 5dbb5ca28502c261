"""Python front end of the CPU oracle (oracle/oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  It loads oracle/liboracle.so (own maths layer) or, with variant="glm",
oracle/_ref/liboracle_glm.so (same source compiled against the reference's vendored glm).

`OracleFrame` mirrors the pass sequencing of the reference backend
(StratusRendererBackend.cpp:1578-1952): cull -> colour -> [tile lists] -> gather -> denoise passes
with the Fbo1/Fbo2 ping-pong and the previous-frame swap.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from stratusgfx_b200._abi import (AtlasEntry, FrameParams, MAX_TOTAL_SHADOW_ATLASES, MAX_TOTAL_VPLS_PER_FRAME,
                                   TILE_DIVISOR)

_HERE = os.path.dirname(os.path.abspath(__file__))
_REFERENCE_ROOT = "/root/reference"
_FLAGS = ["-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared"]


def build(force=False):
    """Compile the oracle; also the glm-backed cross-check when the reference tree is present."""
    src = os.path.join(_HERE, "oracle.cpp")
    deps = [src, os.path.join(_HERE, "vm.h"), os.path.join(_HERE, "..", "include", "stratusgi.h")]
    newest = max(os.path.getmtime(d) for d in deps)
    out = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(out) or os.path.getmtime(out) < newest:
        subprocess.check_call(["g++"] + _FLAGS + ["-o", out, src])
    if os.path.isdir(os.path.join(_REFERENCE_ROOT, "glm")):
        ref_dir = os.path.join(_HERE, "_ref")
        os.makedirs(ref_dir, exist_ok=True)
        out_glm = os.path.join(ref_dir, "liboracle_glm.so")
        if force or not os.path.exists(out_glm) or os.path.getmtime(out_glm) < newest:
            subprocess.check_call(["g++"] + _FLAGS + ["-DORACLE_USE_GLM", "-I" + _REFERENCE_ROOT, "-o", out_glm, src])
        # the reference headers that compile here (StratusMath.h, StratusGpuCommon.h), unmodified, behind a shim (oracle/ref_shim.cpp)
        shim = os.path.join(_HERE, "ref_shim.cpp")
        engine = os.path.join(_REFERENCE_ROOT, "Source", "Engine")
        out_ref = os.path.join(ref_dir, "libref_stratus.so")
        if os.path.isdir(engine) and (force or not os.path.exists(out_ref) or os.path.getmtime(out_ref) < os.path.getmtime(shim)):
            subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-DGLM_FORCE_PURE",
                                   "-I" + os.path.join(_HERE, "ref_stubs"), "-I" + engine, "-I" + _REFERENCE_ROOT,
                                   "-I" + os.path.join(_REFERENCE_ROOT, "glm"), "-o", out_ref, shim])
        # the reference's own GLSL text for the gather's pure-maths functions, compiled as C++ against its glm (oracle/make_ref_glsl.py)
        try:
            from oracle import make_ref_glsl
        except ImportError:
            import make_ref_glsl
        make_ref_glsl.build(force)
    return out


def reference_glsl():
    """ctypes handle of oracle/_ref/libref_glsl.so (the reference's GLSL functions compiled here), or None where it was never built."""
    path = os.path.join(_HERE, "_ref", "libref_glsl.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    L.refglsl_set_uniforms.argtypes = [C.c_int, C.c_float, C.c_void_p, C.c_float]
    L.refglsl_vpl_lighting.argtypes = [C.c_void_p, C.c_void_p]
    L.refglsl_world_pos.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    L.refglsl_random3.argtypes = [C.c_float] * 3
    L.refglsl_random3.restype = C.c_float
    L.refglsl_hash3.argtypes = [C.c_float] * 3
    L.refglsl_hash3.restype = C.c_uint32
    L.refglsl_attenuation.argtypes = [C.c_void_p, C.c_float]
    L.refglsl_attenuation.restype = C.c_float
    L.refglsl_bound_hdr.argtypes = [C.c_float]
    L.refglsl_bound_hdr.restype = C.c_float
    L.refglsl_gather.argtypes = [C.c_void_p]
    L.refglsl_denoise_pass.argtypes = [C.c_void_p]
    for name in ("refglsl_color", "refglsl_stage1", "refglsl_stage2", "refglsl_cull", "refglsl_direct_lighting", "refglsl_ssao", "refglsl_ssao_blur", "refglsl_taa"):
        getattr(L, name).argtypes = [C.c_void_p]
    return L


class RefColorArgs(C.Structure):
    _fields_ = [("cubeRes", C.c_int), ("numAtlases", C.c_int)] + [(n, C.c_void_p) for n in ("sunDir", "sunColor", "lightData", "numVisible", "diffuseIndices",
                                                                                        "shadowAtlases", "diffuseAtlases")]


class RefStage1Args(C.Structure):
    _fields_ = [("W", C.c_int), ("H", C.c_int)] + [(n, C.c_void_p) for n in ("depth", "normal4", "invProjectionView", "out")] + \
               [("tileRowBegin", C.c_int), ("tileRowEnd", C.c_int)]


class RefStage2Args(C.Structure):
    _fields_ = [("tilesX", C.c_int), ("tilesY", C.c_int), ("viewPosition", C.c_void_p), ("lightData", C.c_void_p), ("stage1Data", C.c_void_p),
                ("numVisible", C.c_int), ("vplVisibleIndex", C.c_void_p), ("out", C.c_void_p), ("tileRowBegin", C.c_int), ("tileRowEnd", C.c_int)]


class RefCullArgs(C.Structure):
    _fields_ = [("totalNumLights", C.c_int), ("csmRes", C.c_int)] + [(n, C.c_void_p) for n in ("sunDir", "cascadeProjViews", "cascadePlanes", "csm",
                                                                                              "lightData", "updatedLightData", "vplVisibleIndex")]


class RefDirectArgs(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("W", "H", "cubeRes", "numAtlases", "csmRes", "numLights", "numShadowLights", "rowBegin", "rowEnd")] + \
               [(n, C.c_void_p) for n in ("depth", "normal4", "albedo4", "refl2", "rma4", "ssaoH", "viewPosition", "invProjectionView", "sunDir", "sunColor",
                                          "cascadeProjViews", "cascadePlanes", "csm", "nonShadowCasters", "shadowCasters", "shadowIndices", "shadowAtlases")] + \
               [(n, C.c_float) for n in ("emissionStrength", "depthBias", "zNear", "zFar")] + [("colorOut3", C.c_void_p)]


def reference_glsl_direct_lighting(ref, fp, gb, csm, csm_res, shadow_atlas, cube_res, lights=None, shadow_lights=None, shadow_entries=None,
                                   sun_color=None, emission_strength=0.0, depth_bias=0.0, z_near=1.0, z_far=1000.0, rows=None):
    """pbr.fs (the reference's deferred direct-lighting pass, compiled from its shader text): float RGB of
    `fsColor` before the RGB16F store — what the GI path receives as `screen`.  lights / shadow_lights: float32 [n, 12] PointLight
    records (position xyzw, colour xyzw, radius, farPlane, 2 pad); shadow_entries: int32 [n, 2] atlas entries of the shadow casters."""
    H, W = gb.depth.shape
    out = np.zeros((H, W, 3), np.float32)
    z = lambda n, dt: np.zeros((max(n, 1), 12 if dt == np.float32 else 2), dt)
    lights = z(0, np.float32) if lights is None else np.ascontiguousarray(lights, np.float32)
    shadow_lights = z(0, np.float32) if shadow_lights is None else np.ascontiguousarray(shadow_lights, np.float32)
    shadow_entries = z(0, np.int32) if shadow_entries is None else np.ascontiguousarray(shadow_entries, np.int32)
    a = RefDirectArgs()
    a.W, a.H, a.cubeRes, a.numAtlases, a.csmRes = W, H, cube_res, int(shadow_atlas.shape[0]), csm_res
    a.numLights = 0 if lights is None or lights.shape[1] != 12 or not lights.any() else int(lights.shape[0])
    a.numShadowLights = 0 if not shadow_lights.any() else int(shadow_lights.shape[0])
    a.rowBegin, a.rowEnd = (0, H) if rows is None else rows
    tab = _atlas_ptrs(shadow_atlas)
    sun = list(fp.infiniteLightColor) if sun_color is None else sun_color
    keep = [np.ascontiguousarray(x) for x in (gb.depth, gb.normal, gb.albedo, gb.baseReflectivity, gb.roughnessMetallicAmbient, gb.ssao)] + \
           [_f32(list(fp.viewPosition)), _f32(list(fp.invProjectionView)), _f32(list(fp.infiniteLightDirection)), _f32(sun),
            _f32([list(m) for m in fp.cascadeProjViews]), _f32([list(p) for p in fp.cascadePlanes]), np.ascontiguousarray(csm, np.float32),
            lights, shadow_lights, shadow_entries]
    (a.depth, a.normal4, a.albedo4, a.refl2, a.rma4, a.ssaoH, a.viewPosition, a.invProjectionView, a.sunDir, a.sunColor, a.cascadeProjViews,
     a.cascadePlanes, a.csm, a.nonShadowCasters, a.shadowCasters, a.shadowIndices) = [x.ctypes.data for x in keep]
    a.shadowAtlases = C.addressof(tab)
    a.emissionStrength, a.depthBias, a.zNear, a.zFar = float(emission_strength), float(depth_bias), float(z_near), float(z_far)
    a.colorOut3 = out.ctypes.data
    ref.refglsl_direct_lighting(C.addressof(a))
    return out


def _f32(x):
    return np.ascontiguousarray(np.array(x, np.float32).reshape(-1))


class RefTaaArgs(C.Structure):
    _fields_ = [("W", C.c_int), ("H", C.c_int), ("screen4", C.c_void_p), ("prevScreen4", C.c_void_p), ("velocityH2", C.c_void_p), ("out4", C.c_void_p)]


def reference_glsl_taa(ref, screen, prev_screen, velocity):
    """taa.fs (the reference's shader text, compiled): the RGBA16F target as float16 [H, W, 4].
    screen / prev_screen: uint8 [H, W, 4]; velocity: float16 [H, W, 2]."""
    H, W = screen.shape[:2]
    keep = [np.ascontiguousarray(screen, np.uint8), np.ascontiguousarray(prev_screen, np.uint8), np.ascontiguousarray(velocity, np.float16)]
    out = np.zeros((H, W, 4), np.float32)
    a = RefTaaArgs()
    a.W, a.H = W, H
    a.screen4, a.prevScreen4, a.velocityH2, a.out4 = keep[0].ctypes.data, keep[1].ctypes.data, keep[2].ctypes.data, out.ctypes.data
    ref.refglsl_taa(C.addressof(a))
    with np.errstate(over="ignore"):
        return out.astype(np.float16)


class RefSsaoArgs(C.Structure):
    _fields_ = [("W", C.c_int), ("H", C.c_int), ("structureH4", C.c_void_p), ("rotationH4", C.c_void_p), ("occlusionH", C.c_void_p),
                ("aspectRatio", C.c_float), ("projPlaneZDist", C.c_float), ("intensity", C.c_float), ("out", C.c_void_p),
                ("rowBegin", C.c_int), ("rowEnd", C.c_int)]


def reference_glsl_ssao(ref, structure, rotation, aspectRatio, projPlaneZDist, intensity=5.0):
    """ssao.fs then ssao_blur.fs (the reference's shader text, compiled): (occlusion, blurred) as the
    R16F render targets hold them (float16 [H, W]).  structure: half4 [H, W, 4]; rotation: half4 [4, 4, 4]."""
    H, W = structure.shape[:2]
    structure, rotation = np.ascontiguousarray(structure, np.float16), np.ascontiguousarray(rotation, np.float16)
    a = RefSsaoArgs()
    a.W, a.H, a.rowBegin, a.rowEnd = W, H, 0, H
    a.structureH4, a.rotationH4 = structure.ctypes.data, rotation.ctypes.data
    a.aspectRatio, a.projPlaneZDist, a.intensity = float(aspectRatio), float(projPlaneZDist), float(intensity)
    tmp = np.zeros((H, W), np.float32)
    a.out = tmp.ctypes.data
    ref.refglsl_ssao(C.addressof(a))
    with np.errstate(over="ignore"):
        occlusion = tmp.astype(np.float16)               # the store into the R16F target (round to nearest even)
    a.occlusionH = occlusion.ctypes.data
    tmp2 = np.zeros((H, W), np.float32)
    a.out = tmp2.ctypes.data
    ref.refglsl_ssao_blur(C.addressof(a))
    with np.errstate(over="ignore"):
        return occlusion, tmp2.astype(np.float16)


def reference_glsl_cull(ref, fp, vpl_data, csm, csm_res):
    """viscull_vpls.cs (shader text, one work group of 1024 real threads).  Returns (updatedLightData [4480,16], vplVisibleIndex [N+1]).
    The eight helper invocations race for the output slots as they do on a GPU, so the ORDER of the list is not defined."""
    n = int(fp.totalNumLights)
    updated = np.zeros((MAX_TOTAL_VPLS_PER_FRAME + 1, 16), np.float32)
    visible = np.zeros(n + 2, np.int32)
    a = RefCullArgs()
    a.totalNumLights, a.csmRes = n, csm_res
    keep = [_f32(list(fp.infiniteLightDirection)), _f32([list(m) for m in fp.cascadeProjViews]), _f32([list(p) for p in fp.cascadePlanes]),
            np.ascontiguousarray(csm, np.float32), np.ascontiguousarray(vpl_data, np.float32)]
    a.sunDir, a.cascadeProjViews, a.cascadePlanes, a.csm, a.lightData = [x.ctypes.data for x in keep]
    a.updatedLightData, a.vplVisibleIndex = updated.ctypes.data, visible.ctypes.data
    ref.refglsl_cull(C.addressof(a))
    return updated, visible


def reference_glsl_color(ref, fp, updated, visible, entries, shadow_atlas, diffuse_atlas, cube_res):
    """vpl_light_color.cs (shader text) in place on a copy of the compacted table; returns the coloured table."""
    out = np.ascontiguousarray(updated, np.float32).copy()
    a = RefColorArgs()
    a.cubeRes, a.numAtlases = cube_res, int(shadow_atlas.shape[0])
    sh, df = _atlas_ptrs(shadow_atlas), _atlas_ptrs(diffuse_atlas)
    keep = [_f32(list(fp.infiniteLightDirection)), _f32(list(fp.infiniteLightColor)), np.ascontiguousarray(visible, np.int32), np.ascontiguousarray(entries, np.int32)]
    a.sunDir, a.sunColor, a.numVisible, a.diffuseIndices = [x.ctypes.data for x in keep]
    a.lightData = out.ctypes.data
    a.shadowAtlases, a.diffuseAtlases = C.addressof(sh), C.addressof(df)
    ref.refglsl_color(C.addressof(a))
    return out


def reference_glsl_stage1(ref, fp, gb, tile_rows=None):
    """vpl_tiled_deferred_culling_stage1.cs (shader text): [tiles, 8] float32."""
    H, W = gb.depth.shape
    out = np.zeros(((W // 5) * (H // 5), 8), np.float32)
    a = RefStage1Args()
    a.W, a.H = W, H
    a.tileRowBegin, a.tileRowEnd = (0, H // 5) if tile_rows is None else tile_rows
    keep = [np.ascontiguousarray(gb.depth), np.ascontiguousarray(gb.normal), _f32(list(fp.invProjectionView))]
    a.depth, a.normal4, a.invProjectionView = [x.ctypes.data for x in keep]
    a.out = out.ctypes.data
    ref.refglsl_stage1(C.addressof(a))
    return out


def reference_glsl_stage2(ref, fp, W, H, vpl_data, stage1, visible, tile_rows=None):
    """vpl_tiled_deferred_culling_stage2.cs (shader text): [tiles, 13] int32.  The shader's `numVisible` / `vplVisibleIndex[]` are the
    count and the list that follows it in the cull's output (slot 0 holds the count)."""
    tx, ty = W // 5, H // 5
    if tx % 32 or ty % 2:
        raise ValueError("the reference dispatches whole 32x2-tile work groups (BE:1774-1778): tile grid must be a multiple of 32 x 2")
    out = np.zeros((tx * ty, 13), np.int32)
    v = int(visible[0])
    idx = np.ascontiguousarray(visible[1:1 + v], np.int32)
    a = RefStage2Args()
    a.tilesX, a.tilesY, a.numVisible = tx, ty, v
    a.tileRowBegin, a.tileRowEnd = (0, ty) if tile_rows is None else tile_rows
    keep = [_f32(list(fp.viewPosition)), np.ascontiguousarray(vpl_data, np.float32), np.ascontiguousarray(stage1, np.float32)]
    a.viewPosition, a.lightData, a.stage1Data = [x.ctypes.data for x in keep]
    a.vplVisibleIndex, a.out = idx.ctypes.data, out.ctypes.data
    ref.refglsl_stage2(C.addressof(a))
    return out


class RefDenoiseArgs(C.Structure):
    _fields_ = [("W", C.c_int), ("H", C.c_int), ("multiplier", C.c_int), ("mergeReservoirs", C.c_int), ("final_", C.c_int),
                ("time", C.c_float), ("framesPerSecond", C.c_float)] + \
               [(n, C.c_void_p) for n in ("numVisible", "screenH4", "velocityH2", "normal4", "ids", "depth", "prevNormal4", "prevIds", "prevDepth",
                                          "indirectIlluminationH4", "indirectShadowsH4", "prevIndirectIlluminationH4", "historyDepthH",
                                          "combinedOut3", "giOut3", "reservoirOut4", "historyOut")] + \
               [("rowBegin", C.c_int), ("rowEnd", C.c_int)]


def _to_half(a):
    with np.errstate(over="ignore"):
        return a.astype(np.float16)          # the 16F framebuffer store: IEEE round-to-nearest-even, overflow to infinity


class ReferenceGlslDenoiser:
    """The host loop of StratusRendererBackend.cpp:1879-1952 around vpl_pbr_gi_denoise.fs compiled from the reference's shader text
    (refglsl_denoise_pass): Fbo1/Fbo2 ping-pong, multiplier 2^i - 1, merge pass first, final pass last, swap with the previous-frame
    FBO.  Attachments are stored as binary16 between passes, as the reference's RGB16F / RGBA16F / R16F targets do."""

    def __init__(self, ref, W, H, denoise_iters, rows=None):
        self.ref, self.W, self.H, self.iters = ref, W, H, denoise_iters
        self.rows = (0, H) if rows is None else rows                  # rows of the full-height planes to process (bounded samples)
        self.prev_gi = np.zeros((H, W, 4), np.float16)
        self.prev_history = np.zeros((H, W), np.float16)

    def frame(self, fp, gb, prev_gb, num_visible, gi_color_h4, gi_reservoir_h4):
        H, W = self.H, self.W
        nv = np.array([num_visible], np.int32)
        illum, shadows = np.ascontiguousarray(gi_color_h4), np.ascontiguousarray(gi_reservoir_h4)
        passes = []
        for i in range(self.iters):
            a = RefDenoiseArgs()
            a.W, a.H, a.multiplier = W, H, (1 << i) - 1
            a.rowBegin, a.rowEnd = self.rows
            a.mergeReservoirs, a.final_ = (1 if i < 1 else 0), (1 if i + 1 == self.iters else 0)
            a.time, a.framesPerSecond = float(fp.time), float(np.float32(1.0 / fp.deltaSeconds))
            combined, gi = np.zeros((H, W, 3), np.float32), np.zeros((H, W, 3), np.float32)
            res, hist = np.zeros((H, W, 4), np.float32), np.zeros((H, W), np.float32)
            keep = [np.ascontiguousarray(x) for x in (nv, gb.screen, gb.velocity, gb.normal, gb.id, gb.depth, prev_gb.normal, prev_gb.id, prev_gb.depth,
                                                      illum, shadows, self.prev_gi, self.prev_history)]
            (a.numVisible, a.screenH4, a.velocityH2, a.normal4, a.ids, a.depth, a.prevNormal4, a.prevIds, a.prevDepth,
             a.indirectIlluminationH4, a.indirectShadowsH4, a.prevIndirectIlluminationH4, a.historyDepthH) = [x.ctypes.data for x in keep]
            a.combinedOut3, a.giOut3, a.reservoirOut4, a.historyOut = combined.ctypes.data, gi.ctypes.data, res.ctypes.data, hist.ctypes.data
            self.ref.refglsl_denoise_pass(C.addressof(a))
            pad = lambda v3: np.concatenate([_to_half(v3), np.ones((H, W, 1), np.float16)], axis=-1)      # RGB16F read back as (r, g, b, 1)
            fbo = dict(combined=pad(combined), gi=pad(gi), reservoir=_to_half(res), history=_to_half(hist))
            illum, shadows = fbo["gi"], fbo["reservoir"]
            passes.append(fbo)
        last = passes[-1]
        self.prev_gi, self.prev_history = last["gi"].copy(), last["history"].copy()
        return passes


class RefGatherArgs(C.Structure):
    _fields_ = [("W", C.c_int), ("H", C.c_int), ("cubeRes", C.c_int), ("numAtlases", C.c_int)] + \
               [(n, C.c_void_p) for n in ("depth", "normal4", "albedo4", "refl2", "rma4", "ssaoH", "screenH4", "viewPosition", "invProjectionView")] + \
               [("time", C.c_float), ("minGiOcclusionFactor", C.c_float)] + \
               [(n, C.c_void_p) for n in ("lightData", "numVisible", "shadowIndices", "shadowAtlases", "colorOut3", "reservoirOut4")] + \
               [("rowBegin", C.c_int), ("rowEnd", C.c_int), ("usePerceptualRoughness", C.c_int), ("minRoughness", C.c_float),
                ("fogColor", C.c_float * 3), ("fogDensity", C.c_float)]


def reference_glsl_gather(ref, fp, gb, updated, visible, entries, shadow_atlas, cube_res, rows=None):
    """vpl_pbr_gi.fs (the reference's shader text compiled as C++) over one frame: float `color` [H,W,3] and `reservoir` [H,W,4]
    as written by the shader, before the 16F framebuffer store.  `updated`/`visible`/`entries` are the coloured compacted VPL
    table, the visible list and shadowDiffuseIndices (the outputs of the two stages before the gather)."""
    H, W = gb.depth.shape
    color = np.zeros((H, W, 3), np.float32)
    res = np.zeros((H, W, 4), np.float32)
    a = RefGatherArgs()
    a.W, a.H, a.cubeRes, a.numAtlases = W, H, cube_res, int(shadow_atlas.shape[0])
    tab = _atlas_ptrs(shadow_atlas)
    vp = np.ascontiguousarray(list(fp.viewPosition), np.float32)
    ipv = np.ascontiguousarray(list(fp.invProjectionView), np.float32)
    a.usePerceptualRoughness, a.minRoughness, a.fogDensity = int(fp.usePerceptualRoughness), float(fp.minRoughness), float(fp.fogDensity)
    for i in range(3):
        a.fogColor[i] = float(fp.fogColor[i])
    a.rowBegin, a.rowEnd = (0, H) if rows is None else rows          # rows: (begin, end) of the full-height planes to shade (bounded samples)
    keep = [np.ascontiguousarray(x) for x in (gb.depth, gb.normal, gb.albedo, gb.baseReflectivity, gb.roughnessMetallicAmbient, gb.ssao, gb.screen,
                                              updated, visible, entries)]
    (a.depth, a.normal4, a.albedo4, a.refl2, a.rma4, a.ssaoH, a.screenH4, a.lightData, a.numVisible, a.shadowIndices) = [x.ctypes.data for x in keep]
    a.viewPosition, a.invProjectionView = vp.ctypes.data, ipv.ctypes.data
    a.time, a.minGiOcclusionFactor = float(fp.time), float(fp.minGiOcclusionFactor)
    a.shadowAtlases = C.addressof(tab)
    a.colorOut3, a.reservoirOut4 = color.ctypes.data, res.ctypes.data
    ref.refglsl_gather(C.addressof(a))
    return color, res


def reference_shim():
    """ctypes handle of oracle/_ref/libref_stratus.so (reference code compiled here), or None where it was never built."""
    path = os.path.join(_HERE, "_ref", "libref_stratus.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    L.ref_is_sphere_in_frustum.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
    L.ref_layout.argtypes = [C.c_void_p]
    L.ref_init_vpl_frame_data.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p]
    return L


class OrcDims(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("cube_res", C.c_int32), ("csm_res", C.c_int32),
                ("y_offset", C.c_int32), ("full_height", C.c_int32)]


class OrcDenoisePass(C.Structure):
    _fields_ = [("multiplier", C.c_int32), ("mergeReservoirs", C.c_int32), ("final_", C.c_int32), ("numVisible", C.c_int32)] + \
               [(n, C.c_void_p) for n in ("screenH4", "velocityH2", "normal4", "ids", "depth", "prevNormal4", "prevIds", "prevDepth",
                                          "indirectIlluminationH4", "indirectShadowsH4", "prevIndirectIlluminationH4", "historyDepthH",
                                          "combinedOutH4", "giOutH4", "reservoirOutH4", "historyOutH")]


_libs = {}


def has_variant(variant):
    path = os.path.join(_HERE, "liboracle.so") if variant == "own" else os.path.join(_HERE, "_ref", "liboracle_glm.so")
    return os.path.exists(path)


def lib(variant="own"):
    if variant not in _libs:
        if variant == "own":
            path = os.path.join(_HERE, "liboracle.so")
            if not os.path.exists(path):
                build()
        else:
            path = os.path.join(_HERE, "_ref", "liboracle_glm.so")
        L = C.CDLL(path)
        L.orc_f2h.restype = C.c_uint16
        L.orc_f2h.argtypes = [C.c_float]
        L.orc_h2f.restype = C.c_float
        L.orc_h2f.argtypes = [C.c_uint16]
        L.orc_hash3.restype = C.c_uint32
        L.orc_hash3.argtypes = [C.c_float] * 3
        L.orc_random3.restype = C.c_float
        L.orc_random3.argtypes = [C.c_float] * 3
        for name, n in (("orc_cull", 8), ("orc_color", 7), ("orc_stage1", 5), ("orc_gather", 15)):
            getattr(L, name).argtypes = [C.c_void_p] * n
        L.orc_gather_tiles.argtypes = [C.c_void_p] * 16
        L.orc_stage2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.orc_denoise_pass.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_vpl_prepass.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_test_vpl_lighting.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_test_world_pos.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        L.orc_test_attenuation.argtypes = [C.c_void_p, C.c_float]
        L.orc_test_attenuation.restype = C.c_float
        _libs[variant] = L
    return _libs[variant]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _atlas_ptrs(arr):
    """arr: [A, 320, 6, R, R(,4)] contiguous -> (c_void_p * 14) of per-atlas base pointers."""
    tab = (C.c_void_p * MAX_TOTAL_SHADOW_ATLASES)()
    stride = arr[0].nbytes
    for a in range(arr.shape[0]):
        tab[a] = arr.ctypes.data + a * stride
    return tab


def vpl_prepass(lights, view_position, frustum_planes, variant="own"):
    """UpdatePointLights_ (virtual point lights) + InitVplFrameData_ on the CPU (orc_vpl_prepass).
    lights: numpy array of stratusgfx_b200.prepass.LIGHT_DTYPE.  Returns (vplData [N,16] f32, atlas [N,2] i32, order [N] i32)."""
    n = int(lights.shape[0])
    cap = max(1, min(n, 10000))
    vd = np.zeros((cap, 16), np.float32)
    at = np.zeros((cap, 2), np.int32)
    order = np.zeros(cap, np.int32)
    vp = np.ascontiguousarray(view_position, np.float32)
    fr = np.ascontiguousarray(frustum_planes, np.float32).reshape(6, 4)
    lights = np.ascontiguousarray(lights)
    count = lib(variant).orc_vpl_prepass(_p(lights), n, _p(vp), _p(fr), _p(vd), _p(at), _p(order))
    return vd[:count], at[:count], order[:count]


def render_vpl_cubemaps(boxes_u8, vpl_data, atlas_by_vpl, vpl_indices, res, shadow_atlas, diffuse_atlas, variant="own"):
    """Widening row f2 on the CPU (orc_render_vpl_cubemaps): renders the cube depth + diffuse maps of the listed VPLs in place into
    `shadow_atlas` (uint16 [A,320,6,R,R]) and `diffuse_atlas` (uint8 [A,320,6,R,R,4]).  boxes_u8: uint8 [n,32] SgiSceneBox records."""
    L = lib(variant)
    boxes = np.ascontiguousarray(boxes_u8, np.uint8)
    idx = np.ascontiguousarray(vpl_indices, np.int32)
    vd = np.ascontiguousarray(vpl_data, np.float32); at = np.ascontiguousarray(atlas_by_vpl, np.int32)
    L.orc_render_vpl_cubemaps(_p(boxes), int(boxes.shape[0]), _p(vd), _p(at), _p(idx), int(idx.shape[0]), int(res),
                              _atlas_ptrs(shadow_atlas), _atlas_ptrs(diffuse_atlas))


def shadow_vpl_depth(fs_position, light_pos, far_plane, variant="own"):
    L = lib(variant)
    L.orc_shadow_vpl_depth.restype = C.c_float
    L.orc_shadow_vpl_depth.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
    a = np.ascontiguousarray(fs_position, np.float32); b = np.ascontiguousarray(light_pos, np.float32)
    return float(L.orc_shadow_vpl_depth(_p(a), _p(b), float(far_plane)))


class OracleFrame:
    """Whole-frame CPU pipeline over numpy buffers, keeping temporal state between frames."""

    def __init__(self, width, height, cube_res=32, csm_res=1024, denoise_iters=4, variant="own", y_offset=0, full_height=0):
        self.L = lib(variant)
        self.W, self.H = width, height
        self.dims = OrcDims(width, height, cube_res, csm_res, y_offset, full_height)
        self.denoise_iters = denoise_iters
        self.tiles = (width // TILE_DIVISOR) * (height // TILE_DIVISOR)
        z4 = lambda: np.zeros((height, width, 4), np.float16)
        self.prev_gi = z4()                                   # vplGIDenoisedPrevFrameFbo attachment 1
        self.prev_history = np.zeros((height, width), np.float16)   # attachment 3

    # ---- a1 ----
    def cull(self, fp, vpl_data, atlas_by_vpl, csm):
        n = fp.totalNumLights
        self.updated = np.zeros((MAX_TOTAL_VPLS_PER_FRAME, 16), np.float32)
        self.visible = np.zeros(max(n, 0) + 1, np.int32)
        self.entries = np.zeros((MAX_TOTAL_VPLS_PER_FRAME, 2), np.int32)
        self.L.orc_cull(C.addressof(fp), C.addressof(self.dims), _p(vpl_data), _p(atlas_by_vpl), _p(csm),
                        _p(self.updated), _p(self.visible), _p(self.entries))
        return int(self.visible[0])

    # ---- a2 ----
    def color(self, fp, shadow_atlas, diffuse_atlas):
        self._sh_ptrs = _atlas_ptrs(shadow_atlas)
        self._df_ptrs = _atlas_ptrs(diffuse_atlas)
        self._keep = (shadow_atlas, diffuse_atlas)
        self.L.orc_color(C.addressof(fp), C.addressof(self.dims), _p(self.updated), _p(self.visible), _p(self.entries),
                         C.addressof(self._sh_ptrs), C.addressof(self._df_ptrs))

    # ---- a3 + a4 ----
    def tile_lists(self, fp, gb, vpl_data):
        self.stage1 = np.zeros((self.tiles, 8), np.float32)
        self.stage2 = np.zeros((self.tiles, 13), np.int32)
        self.L.orc_stage1(C.addressof(fp), C.addressof(self.dims), _p(gb.depth), _p(gb.normal), _p(self.stage1))
        v = int(self.visible[0])
        idx = np.ascontiguousarray(self.visible[1:1 + v])
        self.L.orc_stage2(C.addressof(self.dims), _p(self.stage1), _p(vpl_data), v, _p(idx), _p(self.stage2))

    # ---- a5 ----
    def gather(self, fp, gb, tile_mode=False):
        H, W = self.H, self.W
        self.gi_color = np.zeros((H, W, 4), np.float16)
        self.gi_reservoir = np.zeros((H, W, 4), np.float16)
        self.sample_stats = np.zeros(3, np.int64)
        if tile_mode:          # SGI_GATHER_TILE_LIST: needs tile_lists() first
            self.L.orc_gather_tiles(C.addressof(fp), C.addressof(self.dims), _p(gb.depth), _p(gb.normal), _p(gb.albedo), _p(gb.baseReflectivity),
                                    _p(gb.roughnessMetallicAmbient), _p(gb.ssao), _p(self.updated), _p(self.visible), _p(self.entries),
                                    C.addressof(self._sh_ptrs), _p(self.gi_color), _p(self.gi_reservoir), _p(self.sample_stats), _p(self.stage2))
            return
        self.L.orc_gather(C.addressof(fp), C.addressof(self.dims), _p(gb.depth), _p(gb.normal), _p(gb.albedo), _p(gb.baseReflectivity),
                          _p(gb.roughnessMetallicAmbient), _p(gb.ssao), _p(self.updated), _p(self.visible), _p(self.entries),
                          C.addressof(self._sh_ptrs), _p(self.gi_color), _p(self.gi_reservoir), _p(self.sample_stats))

    # ---- a6: the host loop of BE:1879-1952 ----
    def denoise(self, fp, gb, prev_gb, keep_passes=False):
        H, W = self.H, self.W
        mk4 = lambda: np.zeros((H, W, 4), np.float16)
        fbos = [dict(combined=mk4(), gi=mk4(), reservoir=mk4(), history=np.zeros((H, W), np.float16)) for _ in range(2)]
        illum, shadows = self.gi_color, self.gi_reservoir
        self.pass_reservoirs = []
        iters = self.denoise_iters
        last = None
        for i in range(iters):
            fbo = fbos[i % 2]
            ps = OrcDenoisePass()
            ps.multiplier = (1 << i) - 1
            ps.mergeReservoirs = 1 if i < 1 else 0
            ps.final_ = 1 if i + 1 == iters else 0
            ps.numVisible = int(self.visible[0])
            ps.screenH4, ps.velocityH2, ps.normal4, ps.ids, ps.depth = _p(gb.screen), _p(gb.velocity), _p(gb.normal), _p(gb.id), _p(gb.depth)
            ps.prevNormal4, ps.prevIds, ps.prevDepth = _p(prev_gb.normal), _p(prev_gb.id), _p(prev_gb.depth)
            ps.indirectIlluminationH4, ps.indirectShadowsH4 = _p(illum), _p(shadows)
            ps.prevIndirectIlluminationH4, ps.historyDepthH = _p(self.prev_gi), _p(self.prev_history)
            ps.combinedOutH4, ps.giOutH4, ps.reservoirOutH4, ps.historyOutH = _p(fbo["combined"]), _p(fbo["gi"]), _p(fbo["reservoir"]), _p(fbo["history"])
            self.L.orc_denoise_pass(C.addressof(fp), C.addressof(self.dims), C.addressof(ps))
            illum, shadows = fbo["gi"], fbo["reservoir"]
            if keep_passes:
                self.pass_reservoirs.append(fbo["reservoir"].copy())
            last = fbo
        self.combined = last["combined"]        # blitted into lightingFbo (BE:1946)
        self.gi_illum = last["gi"]
        self.history = last["history"]
        # swap with the previous-frame FBO (BE:1949-1951)
        self.prev_gi, self.prev_history = last["gi"].copy(), last["history"].copy()

    def frame(self, fp, gb, prev_gb, vpl_data, atlas_by_vpl, csm, shadow_atlas, diffuse_atlas, tile_lists=True, keep_passes=False):
        self.cull(fp, vpl_data, atlas_by_vpl, csm)
        self.color(fp, shadow_atlas, diffuse_atlas)
        if tile_lists:
            self.tile_lists(fp, gb, vpl_data)
        self.gather(fp, gb, tile_mode=(fp.gatherMode == 1))
        self.denoise(fp, gb, prev_gb, keep_passes=keep_passes)
