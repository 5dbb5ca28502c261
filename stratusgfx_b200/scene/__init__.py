"""Procedural inputs for the VPL GI pass — numpy front end of scene_gen.cpp.

Host-side only.  Produces, for a named workload (c1..c5), the G-buffer planes, VPL table,
cube shadow/diffuse atlases, sun CSM and the per-frame uniforms, in the wire formats of
include/stratusgi.h.
"""
import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

from .._abi import (AtlasEntry, FrameParams, VplData, MAX_TOTAL_SHADOWS_PER_ATLAS,
                    MAX_TOTAL_SHADOW_ATLASES)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsgi_scene.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "scene_gen.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-o", _LIB_PATH, src])
    return _LIB_PATH


class _Desc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("num_vpls", C.c_int32),
                ("csm_res", C.c_int32), ("cube_res", C.c_int32), ("seed", C.c_uint32), ("yaw_deg_per_frame", C.c_float)]


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.sgs_create.restype = C.c_void_p
        _lib.sgs_create.argtypes = [C.POINTER(_Desc)]
        _lib.sgs_destroy.argtypes = [C.c_void_p]
        _lib.sgs_frame_params.argtypes = [C.c_void_p, C.c_int, C.POINTER(FrameParams)]
        _lib.sgs_gbuffer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 9
        _lib.sgs_vpls.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.sgs_boxes.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _lib.sgs_cubemaps.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib.sgs_csm.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.sgs_sun_visibility.argtypes = [C.c_void_p, C.c_void_p]
        _lib.sgs_ssao_rotation_table.argtypes = [C.c_void_p]
        _lib.sgs_fovy_degrees.restype = C.c_double
        _lib.sgs_float_to_half.restype = C.c_uint16
        _lib.sgs_float_to_half.argtypes = [C.c_float]
    return _lib


# BASELINE.json configs -> concrete workloads (SURVEY.md §8 "Config shorthands")
WORKLOADS = {
    #        kind, W,    H,    N,    denoise iterations
    "tiny": (0, 160, 90, 32, 4),       # test-sized
    "c1": (0, 640, 360, 64, 4),
    "c2": (0, 1920, 1080, 256, 4),
    "c3": (1, 1920, 1080, 1024, 5),
    "c4": (1, 3840, 2160, 2048, 4),
    "c5": (1, 7680, 4320, 4096, 4),
}
_CONFIG_INDEX = {"tiny": 0, "c1": 1, "c2": 2, "c3": 3, "c4": 4, "c5": 5}


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class GBuffer:
    """One frame's per-pixel planes for rows [y0, y0+rows)."""
    y0: int
    rows: int
    depth: np.ndarray      # float32 [rows, W]
    normal: np.ndarray     # uint8  [rows, W, 4]
    albedo: np.ndarray     # uint8  [rows, W, 4]
    baseReflectivity: np.ndarray  # uint8 [rows, W, 2]
    roughnessMetallicAmbient: np.ndarray  # uint8 [rows, W, 4]
    ssao: np.ndarray       # float16 [rows, W]
    velocity: np.ndarray   # float16 [rows, W, 2]
    id: np.ndarray         # float32 [rows, W]
    screen: np.ndarray     # float16 [rows, W, 4]


class Scene:
    def __init__(self, workload="c1", width=None, height=None, num_vpls=None, kind=None, csm_res=1024, cube_res=32,
                 yaw_deg_per_frame=0.0, seed=None):
        k, w, h, n, iters = WORKLOADS[workload]
        self.workload = workload
        self.kind = k if kind is None else kind
        self.width = w if width is None else width
        self.height = h if height is None else height
        self.num_vpls = n if num_vpls is None else num_vpls
        self.denoise_iters = iters
        self.csm_res = csm_res
        self.cube_res = cube_res
        self.seed = (0x57A7 + _CONFIG_INDEX[workload]) if seed is None else seed
        lib = _load()
        d = _Desc(self.kind, self.width, self.height, self.num_vpls, csm_res, cube_res, self.seed, yaw_deg_per_frame)
        self._h = lib.sgs_create(C.byref(d))
        if not self._h:
            raise ValueError("sgs_create rejected the scene description")

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.sgs_destroy(self._h)
            self._h = None

    # uniforms -----------------------------------------------------------------------------------
    def frame_params(self, frame=0) -> FrameParams:
        fp = FrameParams()
        _load().sgs_frame_params(self._h, frame, C.byref(fp))
        return fp

    # per-pixel planes ---------------------------------------------------------------------------
    def gbuffer(self, frame=0, y0=0, rows=None) -> GBuffer:
        rows = self.height - y0 if rows is None else rows
        W = self.width
        g = GBuffer(y0, rows,
                    np.empty((rows, W), np.float32), np.empty((rows, W, 4), np.uint8), np.empty((rows, W, 4), np.uint8),
                    np.empty((rows, W, 2), np.uint8), np.empty((rows, W, 4), np.uint8), np.empty((rows, W), np.float16),
                    np.empty((rows, W, 2), np.float16), np.empty((rows, W), np.float32), np.empty((rows, W, 4), np.float16))
        _load().sgs_gbuffer(self._h, frame, y0, rows, _ptr(g.depth), _ptr(g.normal), _ptr(g.albedo), _ptr(g.baseReflectivity),
                            _ptr(g.roughnessMetallicAmbient), _ptr(g.ssao), _ptr(g.velocity), _ptr(g.id), _ptr(g.screen))
        return g

    # VPL table + atlas assignment ------------------------------------------------------------------
    def vpls(self):
        n = self.num_vpls
        data = np.zeros((n, 16), np.float32)
        atlas = np.zeros((n, 2), np.int32)
        _load().sgs_vpls(self._h, _ptr(data), _ptr(atlas))
        return data, atlas

    def boxes(self):
        """The scene's geometry as SgiSceneBox records (numpy structured array viewable as bytes): what
        sgi_render_vpl_cubemaps casts its texel rays against."""
        from .._abi import SceneBox
        n = _load().sgs_boxes(self._h, None, 0)
        arr = (SceneBox * n)()
        _load().sgs_boxes(self._h, C.cast(arr, C.c_void_p), n)
        return np.frombuffer(bytes(arr), np.uint8).reshape(n, C.sizeof(SceneBox)).copy()

    def num_atlases(self):
        return min(MAX_TOTAL_SHADOW_ATLASES, (self.num_vpls + MAX_TOTAL_SHADOWS_PER_ATLAS - 1) // MAX_TOTAL_SHADOWS_PER_ATLAS)

    def cubemaps(self):
        """(depth uint16 [A,320,6,R,R], diffuse uint8 [A,320,6,R,R,4]) — contiguous concatenation of the atlases."""
        A, R = max(1, self.num_atlases()), self.cube_res
        depth = np.full((A, MAX_TOTAL_SHADOWS_PER_ATLAS, 6, R, R), 65535, np.uint16)
        diffuse = np.zeros((A, MAX_TOTAL_SHADOWS_PER_ATLAS, 6, R, R, 4), np.uint8)
        _load().sgs_cubemaps(self._h, 0, self.num_vpls, _ptr(depth), _ptr(diffuse))
        return depth, diffuse

    def sun_visibility(self):
        """Ray-traced (exact geometry) sun visibility per VPL, uint8 [N]."""
        out = np.zeros(self.num_vpls, np.uint8)
        _load().sgs_sun_visibility(self._h, _ptr(out))
        return out

    def csm(self, frame=0):
        out = np.empty((4, self.csm_res, self.csm_res), np.float32)
        _load().sgs_csm(self._h, frame, _ptr(out))
        return out
