"""Host-side mirror of the reference's GI pass interface over the C-ABI (include/stratusgi.h).

`VplGiPass` exposes the three RendererBackend methods the CUDA path replaces, under their
reference names (Source/Engine/StratusRendererBackend.h:666-669):

    PerformVirtualPointLightCullingStage1   -> sgi_cull
    PerformVirtualPointLightCullingStage2   -> sgi_color (+ sgi_tile_lists, dormant in the reference)
    ComputeVirtualPointLightGlobalIllumination -> sgi_gather + sgi_denoise

PyTorch is used only for device memory (torch.cuda tensors own the buffers that are handed
over as raw pointers) and streams.  There is no CPU fallback: a missing library or a missing
GPU raises.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import (BandParams, Buffers, CreateParams, DirectParams, FrameParams, PeerInfo, PrepassParams, SsaoParams, TaaParams, MAX_TOTAL_SHADOW_ATLASES, MAX_TOTAL_VPLS_PER_FRAME,
                   TILE_DIVISOR)

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libstratusgi.so")
_lib = None


class StratusGiError(RuntimeError):
    pass


def library_path():
    return _LIB_PATH


def load_library():
    """Load libstratusgi.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise StratusGiError(f"{_LIB_PATH} is missing: run `python -m stratusgfx_b200.build` (nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(_LIB_PATH)
    H = C.c_void_p
    L.sgi_create.argtypes = [C.POINTER(CreateParams), C.POINTER(H)]
    L.sgi_destroy.argtypes = [H]
    L.sgi_resize.argtypes = [H, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    L.sgi_bind_device_buffers.argtypes = [H, C.POINTER(Buffers)]
    L.sgi_register_gl.argtypes = [H, C.c_void_p, C.c_int32]
    L.sgi_vpl_prepass.argtypes = [H, C.POINTER(PrepassParams)]
    L.sgi_render_vpl_cubemaps.argtypes = [H, C.c_void_p]
    L.sgi_direct_lighting.argtypes = [H, C.POINTER(FrameParams), C.POINTER(DirectParams)]
    L.sgi_ssao.argtypes = [H, C.POINTER(SsaoParams)]
    L.sgi_taa.argtypes = [H, C.POINTER(TaaParams)]
    for name in ("sgi_cull", "sgi_color", "sgi_tile_lists", "sgi_gather", "sgi_denoise", "sgi_frame"):
        getattr(L, name).argtypes = [H, C.POINTER(FrameParams)]
    L.sgi_denoise_pass.argtypes = [H, C.POINTER(FrameParams), C.c_int32]
    L.sgi_early_begin.argtypes = [H, C.c_void_p]
    L.sgi_cull_early.argtypes = [H, C.POINTER(FrameParams), C.c_void_p, C.c_int32]
    L.sgi_band_setup.argtypes = [H, C.POINTER(BandParams)]
    L.sgi_band_export.argtypes = [H, C.POINTER(PeerInfo)]
    L.sgi_band_attach.argtypes = [H, C.c_int32, C.POINTER(PeerInfo)]
    L.sgi_frame_image.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.sgi_band_timeouts.argtypes = [H, C.POINTER(C.c_int32)]
    L.sgi_band_reprojection_misses.argtypes = [H, C.POINTER(C.c_int32)]
    L.sgi_band_stats.argtypes = [H, C.POINTER(C.c_int32)]
    L.sgi_visible_count.argtypes = [H, C.POINTER(C.c_int32)]
    L.sgi_get_plane.argtypes = [H, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.sgi_download_plane.argtypes = [H, C.c_int32, C.c_void_p, C.c_int64]
    L.sgi_reset_history.argtypes = [H]
    L.sgi_synchronize.argtypes = [H]
    L.sgi_stream.argtypes = [H]
    L.sgi_stream.restype = C.c_void_p
    L.sgi_launch_count.argtypes = [H]
    L.sgi_launch_count.restype = C.c_int64
    L.sgi_last_error.argtypes = [H]
    L.sgi_last_error.restype = C.c_char_p
    L.sgi_version.restype = C.c_char_p
    _lib = L
    return L


_PER_PIXEL = ("depth", "normal", "albedo", "baseReflectivity", "roughnessMetallicAmbient", "velocity", "id",
              "prevNormal", "prevId", "prevDepth", "ssao", "screen")


class _DeviceView:
    """A torch tensor over library-owned device memory (no copy, no ownership) via __cuda_array_interface__."""

    def __init__(self, torch, ptr, shape, device):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f2", "data": (int(ptr), False), "version": 3, "strides": None}
        self.tensor = torch.as_tensor(self, device=device)


class VplGiPass:
    """One GI pass instance for a frame (or a row band of it) on one GPU."""

    def __init__(self, width, full_height, y_offset=0, rows=None, cube_res=32, csm_res=1024, denoise_iters=4, device=0,
                 use_torch_stream=False):
        import torch
        if not torch.cuda.is_available():
            raise StratusGiError("no CUDA device: the GI pass has no CPU path")
        self.torch = torch
        self.L = load_library()
        self.device = torch.device("cuda", device)
        rows = full_height - y_offset if rows is None else rows
        self.width, self.full_height, self.y_offset, self.rows = width, full_height, y_offset, rows
        self.cube_res, self.csm_res, self.denoise_iters = cube_res, csm_res, denoise_iters
        stream = None
        if use_torch_stream:
            # run on torch's current stream.  Its default stream has handle 0, which the ABI reads as
            # "create your own", so name the legacy default stream explicitly (cudaStreamLegacy == 0x1).
            stream = torch.cuda.current_stream(self.device).cuda_stream or 1
        cp = CreateParams(width, full_height, y_offset, rows, cube_res, csm_res, denoise_iters, device, stream)
        self._h = C.c_void_p()
        rc = self.L.sgi_create(C.byref(cp), C.byref(self._h))
        if rc != _abi.SGI_OK:
            raise StratusGiError(f"sgi_create failed with status {rc}")
        self.t = {}            # name -> torch tensor keeping the device memory alive
        self.buffers = Buffers()
        tiles = (width // TILE_DIVISOR) * (rows // TILE_DIVISOR)
        dev = self.device
        # outputs the caller owns in the reference (state_.vpls.*, StratusRendererBackend.h:352-357)
        self.t["vplUpdatedData"] = torch.zeros((MAX_TOTAL_VPLS_PER_FRAME, 16), dtype=torch.float32, device=dev)
        self.t["shadowDiffuseIndices"] = torch.zeros((MAX_TOTAL_VPLS_PER_FRAME, 2), dtype=torch.int32, device=dev)
        self.t["vplStage1Results"] = torch.zeros((max(tiles, 1), 8), dtype=torch.float32, device=dev)
        self.t["vplVisiblePerTile"] = torch.zeros((max(tiles, 1), 13), dtype=torch.int32, device=dev)
        self.t["lightingOut"] = torch.zeros((rows, width, 4), dtype=torch.float16, device=dev)
        self.tiles = tiles

    # ---- buffer hand-over --------------------------------------------------------------------------
    def _check(self, rc):
        if rc != _abi.SGI_OK:
            raise StratusGiError(f"status {rc}: {self.L.sgi_last_error(self._h).decode()}")

    def _up(self, a):
        torch = self.torch
        if isinstance(a, np.ndarray):
            return torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        return a.to(self.device).contiguous()

    def set_gbuffer(self, gb, prev_gb):
        """gb / prev_gb: scene.GBuffer of the band's rows (numpy) or dicts of cuda tensors."""
        get = (lambda o, k: getattr(o, k)) if not isinstance(gb, dict) else (lambda o, k: o[k])
        for k in ("depth", "normal", "albedo", "baseReflectivity", "roughnessMetallicAmbient", "velocity", "id", "ssao", "screen"):
            self.t[k] = self._up(get(gb, k))
        self.t["prevNormal"] = self._up(get(prev_gb, "normal"))
        self.t["prevId"] = self._up(get(prev_gb, "id"))
        self.t["prevDepth"] = self._up(get(prev_gb, "depth"))

    def set_vpls(self, vpl_data, atlas_by_vpl):
        self.t["vplData"] = self._up(vpl_data)
        self.t["atlasEntryByVpl"] = self._up(atlas_by_vpl)
        n = int(self.t["vplData"].shape[0])
        self.t["vplVisibleIndices"] = self.torch.zeros(n + 1, dtype=self.torch.int32, device=self.device)

    def set_atlases(self, shadow_atlas, diffuse_atlas):
        """[A,320,6,R,R] uint16 and [A,320,6,R,R,4] uint8, contiguous concatenation of the atlases."""
        torch = self.torch
        if isinstance(shadow_atlas, np.ndarray):
            shadow_atlas = torch.from_numpy(shadow_atlas.view(np.int16))     # torch has no uint16 arithmetic; bytes are what matter
        self.t["shadowAtlas"] = shadow_atlas.to(self.device).contiguous()
        self.t["diffuseAtlas"] = self._up(diffuse_atlas)

    def set_csm(self, csm):
        self.t["infiniteLightShadowMap"] = self._up(csm)

    def bind(self):
        b = Buffers()
        for k in _PER_PIXEL + ("lightingOut", "vplData", "vplUpdatedData", "vplVisibleIndices", "shadowDiffuseIndices",
                               "vplStage1Results", "vplVisiblePerTile", "atlasEntryByVpl", "infiniteLightShadowMap", "vplCountDevice"):
            if k in self.t:
                setattr(b, k, self.t[k].data_ptr())
        if "shadowAtlas" in self.t:
            sa, da = self.t["shadowAtlas"], self.t["diffuseAtlas"]
            for a in range(min(sa.shape[0], MAX_TOTAL_SHADOW_ATLASES)):
                b.shadowCubeMaps[a] = sa[a].data_ptr()
                b.diffuseCubeMaps[a] = da[a].data_ptr()
        self.buffers = b
        self.torch.cuda.synchronize(self.device)     # uploads/fills ran on torch's stream; the pass runs on its own
        self._check(self.L.sgi_bind_device_buffers(self._h, C.byref(b)))

    # ---- the step before the cull: UpdatePointLights_ for virtual point lights (BE:1359-1392) + InitVplFrameData_ ----
    def set_lights(self, records):
        """records: numpy array of prepass.LIGHT_DTYPE (every virtual point light of the scene).  Allocates the cull's
        input table for min(len, 10000) VPLs; vpl_prepass() fills it on the device."""
        torch = self.torch
        n = int(records.shape[0])
        cap = max(1, min(n, _abi.MAX_TOTAL_VPLS_BEFORE_CULLING))
        self.num_lights = n
        self.t["lightRecords"] = torch.from_numpy(np.ascontiguousarray(records).view(np.uint8).reshape(n, 32)).to(self.device)
        self.t["vplData"] = torch.zeros((cap, 16), dtype=torch.float32, device=self.device)
        self.t["atlasEntryByVpl"] = torch.zeros((cap, 2), dtype=torch.int32, device=self.device)
        self.t["vplVisibleIndices"] = torch.zeros(cap + 1, dtype=torch.int32, device=self.device)
        self.t["prepassOrder"] = torch.zeros(cap, dtype=torch.int32, device=self.device)
        self.t["vplCountDevice"] = torch.zeros(1, dtype=torch.int32, device=self.device)
        return cap            # the upper bound to pass as FrameParams.totalNumLights

    def vpl_prepass(self, view_position, frustum_planes):
        """Frustum test, distance order and the 10000 cap on the device (stream-ordered, no read-back)."""
        pp = PrepassParams()
        for i in range(3):
            pp.viewPosition[i] = float(view_position[i])
        fr = np.asarray(frustum_planes, np.float32).reshape(6, 4)
        for k in range(6):
            for i in range(4):
                pp.frustumPlanes[k][i] = float(fr[k, i])
        pp.numLights = self.num_lights
        pp.lights = self.t["lightRecords"].data_ptr()
        pp.vplDataOut = self.t["vplData"].data_ptr()
        pp.atlasEntryOut = self.t["atlasEntryByVpl"].data_ptr()
        pp.orderOut = self.t["prepassOrder"].data_ptr()
        pp.countOut = self.t["vplCountDevice"].data_ptr()
        self._check(self.L.sgi_vpl_prepass(self._h, C.byref(pp)))

    # ---- VPL cube shadow / diffuse map production (include/stratusgi.h sgi_render_vpl_cubemaps) ----
    def set_scene_boxes(self, boxes_u8):
        """boxes_u8: uint8 [n, 32] SgiSceneBox records (Scene.boxes())."""
        self.t["sceneBoxes"] = self.torch.from_numpy(np.ascontiguousarray(boxes_u8, np.uint8)).to(self.device)

    def render_vpl_cubemaps(self, vpl_indices, stream=None):
        """Renders the cube depth + diffuse maps of the listed VPLs (indices into the bound vplData table) into their atlas layers
        (on `stream`, a raw CUDA stream pointer, when given: see early_begin)."""
        from ._abi import CubeRenderParams
        idx = vpl_indices if hasattr(vpl_indices, "data_ptr") else self.torch.as_tensor(np.asarray(vpl_indices, np.int32), device=self.device)
        self.t["cubeRenderIdx"] = idx                      # kept alive until the stream has consumed it
        rp = CubeRenderParams(int(self.t["sceneBoxes"].shape[0]), int(idx.numel()), self.t["sceneBoxes"].data_ptr(), idx.data_ptr(), stream)
        self._check(self.L.sgi_render_vpl_cubemaps(self._h, C.byref(rp)))

    # ---- the passes that produce the `ssao` input plane: ssao.fs + ssao_blur.fs (csrc/k_ssao.cu) ----
    def ssao(self, structure, rotation, aspectRatio, projPlaneZDist, intensity=5.0):
        """RenderSsaoOcclude_ + RenderSsaoBlur_ over a structure buffer (half4 [rows, W, 4]) -> the half [rows, W] plane, kept as
        self.t["ssao"] (the plane bind() hands to the gather) and returned as a torch tensor."""
        torch = self.torch
        self.t["structure"] = torch.from_numpy(np.ascontiguousarray(structure, np.float16)).to(self.device)
        self.t["ssaoRotation"] = torch.from_numpy(np.ascontiguousarray(rotation, np.float16)).to(self.device)
        for k in ("ssaoOcclusion", "ssao"):
            if k not in self.t or tuple(self.t[k].shape) != (self.rows, self.width) or self.t[k].dtype != torch.float16:
                self.t[k] = torch.zeros((self.rows, self.width), dtype=torch.float16, device=self.device)
        sp = SsaoParams()
        sp.aspectRatio, sp.projPlaneZDist, sp.intensity = float(aspectRatio), float(projPlaneZDist), float(intensity)
        sp.structureBuffer, sp.rotationLookup = self.t["structure"].data_ptr(), self.t["ssaoRotation"].data_ptr()
        sp.occlusionScratch, sp.ssaoOut = self.t["ssaoOcclusion"].data_ptr(), self.t["ssao"].data_ptr()
        self._check(self.L.sgi_ssao(self._h, C.byref(sp)))
        return self.t["ssao"]

    # ---- the resolve at the end of the frame: taa.fs (csrc/k_taa.cu) ----
    def taa(self, screen, prev_screen, velocity):
        """PerformTaaPostFx_ over uint8 [rows, W, 4] current / previous LDR frames and the half2 velocity plane ->
        (half4 [rows, W, 4] resolved colour, uint8 [rows, W, 4] next history), torch tensors kept in self.t."""
        torch = self.torch
        self.t["taaScreen"] = torch.from_numpy(np.ascontiguousarray(screen, np.uint8)).to(self.device)
        self.t["taaPrev"] = torch.from_numpy(np.ascontiguousarray(prev_screen, np.uint8)).to(self.device)
        self.t["taaVelocity"] = torch.from_numpy(np.ascontiguousarray(velocity, np.float16)).to(self.device)
        self.t["taaOut"] = torch.zeros((self.rows, self.width, 4), dtype=torch.float16, device=self.device)
        self.t["taaHistory"] = torch.zeros((self.rows, self.width, 4), dtype=torch.uint8, device=self.device)
        tp = TaaParams()
        tp.screen, tp.prevScreen, tp.velocity = self.t["taaScreen"].data_ptr(), self.t["taaPrev"].data_ptr(), self.t["taaVelocity"].data_ptr()
        tp.colorOut, tp.historyOut = self.t["taaOut"].data_ptr(), self.t["taaHistory"].data_ptr()
        self._check(self.L.sgi_taa(self._h, C.byref(tp)))
        return self.t["taaOut"], self.t["taaHistory"]

    # ---- the pass before this path: deferred direct lighting, pbr.fs (csrc/k_direct.cu) ----
    def direct_lighting(self, fp, lights=None, shadow_lights=None, shadow_entries=None, sun=True, emission_strength=0.0,
                        depth_bias=0.0, z_near=1.0, z_far=1000.0, point_light_ambient=0.003):
        """pbr.fs over the bound G-buffer -> a half4 [rows, W] plane (kept as self.t["directOut"]): the `screen` input of the GI path.
        lights / shadow_lights: float32 [n, 12] PointLight records; shadow_entries: int32 [n, 2]."""
        torch = self.torch
        dp = DirectParams()
        keep = []
        def up(a, dt):
            if a is None or len(a) == 0:
                return None, 0
            t = torch.from_numpy(np.ascontiguousarray(a, dt)).to(self.device); keep.append(t)
            return t.data_ptr(), int(t.shape[0])
        dp.nonShadowCasters, dp.numLights = up(lights, np.float32)
        dp.shadowCasters, dp.numShadowLights = up(shadow_lights, np.float32)
        dp.shadowIndices, _ = up(shadow_entries, np.int32)
        dp.sunEnabled = 1 if sun else 0
        dp.emissionStrength, dp.infiniteLightDepthBias = float(emission_strength), float(depth_bias)
        dp.infiniteLightZnear, dp.infiniteLightZfar, dp.pointLightAmbientIntensity = float(z_near), float(z_far), float(point_light_ambient)
        if "directOut" not in self.t:
            self.t["directOut"] = torch.zeros((self.rows, self.width, 4), dtype=torch.float16, device=self.device)
        dp.screenOut = self.t["directOut"].data_ptr()
        torch.cuda.synchronize(self.device)
        self._check(self.L.sgi_direct_lighting(self._h, C.byref(fp), C.byref(dp)))
        self.synchronize()
        return self.t["directOut"]

    # ---- the reference's pass interface --------------------------------------------------------------
    def PerformVirtualPointLightCullingStage1(self, fp):
        """viscull_vpls.cs; fills vplUpdatedData / vplVisibleIndices / shadowDiffuseIndices (BE:1578)."""
        self._check(self.L.sgi_cull(self._h, C.byref(fp)))

    def PerformVirtualPointLightCullingStage2(self, fp, tile_lists=True):
        """vpl_light_color.cs, then the (reference-dormant) per-tile list stages (BE:1666)."""
        self._check(self.L.sgi_color(self._h, C.byref(fp)))
        if tile_lists:
            self._check(self.L.sgi_tile_lists(self._h, C.byref(fp)))

    def ComputeVirtualPointLightGlobalIllumination(self, fp):
        """vpl_pbr_gi.fs + the vpl_pbr_gi_denoise.fs passes, blit and history swap (BE:1816)."""
        self._check(self.L.sgi_gather(self._h, C.byref(fp)))
        self._check(self.L.sgi_denoise(self._h, C.byref(fp)))

    # ---- finer-grained entry points --------------------------------------------------------------------
    def cull(self, fp): self._check(self.L.sgi_cull(self._h, C.byref(fp)))
    def color(self, fp): self._check(self.L.sgi_color(self._h, C.byref(fp)))

    def early_begin(self, stream):
        """`stream` (raw CUDA stream pointer) waits until the frame in flight has finished reading the tables and the atlases."""
        self._check(self.L.sgi_early_begin(self._h, stream))

    def cull_early(self, fp, stream, with_color=True):
        """The next frame's cull (+ colouring) on `stream`; the next frame() waits for it instead of running its own."""
        self._check(self.L.sgi_cull_early(self._h, C.byref(fp), stream, 1 if with_color else 0))

    def tile_lists(self, fp): self._check(self.L.sgi_tile_lists(self._h, C.byref(fp)))
    def gather(self, fp): self._check(self.L.sgi_gather(self._h, C.byref(fp)))
    def denoise_pass(self, fp, i): self._check(self.L.sgi_denoise_pass(self._h, C.byref(fp), i))
    def denoise(self, fp): self._check(self.L.sgi_denoise(self._h, C.byref(fp)))
    def frame(self, fp): self._check(self.L.sgi_frame(self._h, C.byref(fp)))
    def synchronize(self): self._check(self.L.sgi_synchronize(self._h))
    def reset_history(self): self._check(self.L.sgi_reset_history(self._h))

    # ---- row bands with peer-to-peer halos (include/stratusgi.h "multi-GPU row bands") -------------------------
    def band_setup(self, compute_y0, compute_rows, rank, world, owns_frame_image=False, halo_mode=0):
        bp = BandParams(compute_y0, compute_rows, rank, world, 1 if owns_frame_image else 0, int(halo_mode))
        self._check(self.L.sgi_band_setup(self._h, C.byref(bp)))
        self.band = bp

    def band_export(self):
        info = PeerInfo()
        self._check(self.L.sgi_band_export(self._h, C.byref(info)))
        return info

    def band_attach(self, which, info):
        self._check(self.L.sgi_band_attach(self._h, which, C.byref(info)))

    def band_timeouts(self):
        n = C.c_int32()
        self._check(self.L.sgi_band_timeouts(self._h, C.byref(n)))
        return n.value

    def band_stats(self):
        """Raw protocol counters (include/stratusgi.h sgi_band_stats); the spin account restarts with every call."""
        out = (C.c_int32 * 8)()
        self._check(self.L.sgi_band_stats(self._h, out))
        return {"epoch_up": out[0], "epoch_down": out[1], "timeouts": out[2], "reprojection_misses": out[3],
                "waits_that_spun": out[5], "spin_sum_kcycles": out[6], "spin_max_kcycles": out[7]}

    def band_reprojection_misses(self):
        """Reprojected previous-frame taps of the final passes so far that fell outside the held rows (halo shorter than the motion)."""
        n = C.c_int32()
        self._check(self.L.sgi_band_reprojection_misses(self._h, C.byref(n)))
        return n.value

    def frame_image(self):
        """Owner band only: the assembled frame image of the last frame as a torch view (half [H, W, 4]) of library memory."""
        ptr, nbytes = C.c_void_p(), C.c_int64()
        self._check(self.L.sgi_frame_image(self._h, C.byref(ptr), C.byref(nbytes)))
        self.synchronize()          # the view is read on torch's stream: the pass's stream has now waited for every band's rows
        return _DeviceView(self.torch, ptr.value, (self.full_height, self.width, 4), self.device).tensor

    def frame_image_wait(self):
        """Owner band only: make the pass's stream wait for every band's rows of the last frame (no host synchronisation)."""
        ptr, nbytes = C.c_void_p(), C.c_int64()
        self._check(self.L.sgi_frame_image(self._h, C.byref(ptr), C.byref(nbytes)))

    def visible_count(self):
        n = C.c_int32()
        self._check(self.L.sgi_visible_count(self._h, C.byref(n)))
        return n.value

    def launch_count(self):
        return int(self.L.sgi_launch_count(self._h))

    def stream_ptr(self):
        return self.L.sgi_stream(self._h)

    def plane(self, which):
        """Blocking download of a library-owned plane as a numpy array."""
        if which in (_abi.PLANE_HISTORY, _abi.PLANE_PREV_HISTORY):
            out = np.empty((self.rows, self.width), np.float16)
        else:
            out = np.empty((self.rows, self.width, 4), np.float16)
        self._check(self.L.sgi_download_plane(self._h, which, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def output(self, name):
        """Blocking download of a caller-owned output buffer as numpy."""
        self.synchronize()
        return self.t[name].cpu().numpy()

    # ---- host-buffer entry point (what a caller without device-resident buffers uses) -----------------
    def stage_host_inputs(self, gb, prev_gb, vpl_data, atlas_by_vpl, csm, refreshed_maps=None):
        """Register one frame's HOST inputs (numpy).  They are kept in pinned memory; `frame_host`
        copies them to the device every call.  `refreshed_maps` = (slot indices, depth [k,6,R,R] uint16,
        diffuse [k,6,R,R,4] uint8): the <= 5 cube maps re-rendered this frame (StratusRendererBackend.h:378)."""
        torch = self.torch
        get = (lambda o, k: getattr(o, k)) if not isinstance(gb, dict) else (lambda o, k: o[k])
        host = {}
        for k in ("depth", "normal", "albedo", "baseReflectivity", "roughnessMetallicAmbient", "velocity", "id", "ssao", "screen"):
            host[k] = torch.from_numpy(np.ascontiguousarray(get(gb, k))).pin_memory()
        host["prevNormal"] = torch.from_numpy(np.ascontiguousarray(get(prev_gb, "normal"))).pin_memory()
        host["prevId"] = torch.from_numpy(np.ascontiguousarray(get(prev_gb, "id"))).pin_memory()
        host["prevDepth"] = torch.from_numpy(np.ascontiguousarray(get(prev_gb, "depth"))).pin_memory()
        host["vplData"] = torch.from_numpy(np.ascontiguousarray(vpl_data)).pin_memory()
        host["atlasEntryByVpl"] = torch.from_numpy(np.ascontiguousarray(atlas_by_vpl)).pin_memory()
        host["infiniteLightShadowMap"] = torch.from_numpy(np.ascontiguousarray(csm)).pin_memory()
        self._host = host
        self._host_maps = None
        if refreshed_maps is not None:
            slots, dep, dif = refreshed_maps
            self._host_maps = (list(slots), torch.from_numpy(np.ascontiguousarray(dep).view(np.int16)).pin_memory(),
                               torch.from_numpy(np.ascontiguousarray(dif)).pin_memory())
        for k, v in host.items():
            if k not in self.t or self.t[k].shape != v.shape:
                self.t[k] = torch.empty(v.shape, dtype=v.dtype, device=self.device)
        if "vplVisibleIndices" not in self.t or self.t["vplVisibleIndices"].numel() != host["vplData"].shape[0] + 1:
            self.t["vplVisibleIndices"] = torch.zeros(host["vplData"].shape[0] + 1, dtype=torch.int32, device=self.device)
        self._host_out = torch.empty((self.rows, self.width, 4), dtype=torch.float16).pin_memory()
        self.bind()
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
        if self._host_maps:
            self.h2d_bytes += self._host_maps[1].numel() * 2 + self._host_maps[2].numel()
        self.d2h_bytes = self._host_out.numel() * 2

    def frame_host(self, fp):
        """One frame from HOST buffers: H2D of every per-frame input, the whole pass, D2H of the lit image.
        Requires use_torch_stream=True so copies and kernels share one stream.  Returns the pinned result."""
        with self.torch.cuda.device(self.device):
            for k, v in self._host.items():
                self.t[k].copy_(v, non_blocking=True)
            if self._host_maps:
                slots, dep, dif = self._host_maps
                sa = self.t["shadowAtlas"].view(-1, 6, self.cube_res, self.cube_res)
                da = self.t["diffuseAtlas"].view(-1, 6, self.cube_res, self.cube_res, 4)
                for j, s in enumerate(slots):
                    sa[s].copy_(dep[j], non_blocking=True)
                    da[s].copy_(dif[j], non_blocking=True)
            self.frame(fp)
            self._host_out.copy_(self.t["lightingOut"], non_blocking=True)
        return self._host_out

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.L.sgi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
