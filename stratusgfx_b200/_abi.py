"""ctypes mirror of include/stratusgi.h (wire formats, params and buffer tables).

Field order and sizes must match the header exactly; tests/test_abi_layout.py checks
the sizes against the reference's static_asserts (StratusGpuCommon.h:297-306).
"""
import ctypes as C

MAX_TOTAL_SHADOW_ATLASES = 14
MAX_TOTAL_SHADOWS_PER_ATLAS = 320
MAX_TOTAL_SHADOW_MAPS = MAX_TOTAL_SHADOW_ATLASES * MAX_TOTAL_SHADOWS_PER_ATLAS
MAX_TOTAL_VPLS_BEFORE_CULLING = 10000
MAX_TOTAL_VPLS_PER_FRAME = MAX_TOTAL_SHADOW_MAPS
MAX_VPLS_PER_TILE = 12
TILE_DIVISOR = 5
NUM_CASCADES = 4

SGI_OK, SGI_ERR_INVALID_ARG, SGI_ERR_CUDA, SGI_ERR_NOT_BOUND, SGI_ERR_UNSUPPORTED = range(5)
GATHER_STOCHASTIC, GATHER_TILE_LIST = 0, 1

(PLANE_GI_COLOR, PLANE_GI_RESERVOIR, PLANE_DENOISE_RES_A, PLANE_DENOISE_RES_B, PLANE_GI_ILLUM,
 PLANE_HISTORY, PLANE_PREV_GI_ILLUM, PLANE_PREV_HISTORY, PLANE_COMBINED) = range(9)


class VplData(C.Structure):
    _fields_ = [("position", C.c_float * 4), ("color", C.c_float * 4), ("specularPosition", C.c_float * 4),
                ("radius", C.c_float), ("farPlane", C.c_float), ("intensity", C.c_float), ("_pad", C.c_float)]


class VplStage1PerTileOutputs(C.Structure):
    _fields_ = [("averageLocalPosition", C.c_float * 4), ("averageLocalNormal", C.c_float * 4)]


class VplStage2PerTileOutputs(C.Structure):
    _fields_ = [("numVisible", C.c_int32), ("indices", C.c_int32 * MAX_VPLS_PER_TILE)]


class AtlasEntry(C.Structure):
    _fields_ = [("index", C.c_int32), ("layer", C.c_int32)]


class CreateParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("full_height", C.c_int32), ("y_offset", C.c_int32), ("rows", C.c_int32),
                ("cube_res", C.c_int32), ("csm_res", C.c_int32), ("denoise_iters", C.c_int32), ("device", C.c_int32),
                ("stream", C.c_void_p)]


class FrameParams(C.Structure):
    _fields_ = [("invProjectionView", C.c_float * 16), ("prevInvProjectionView", C.c_float * 16),
                ("viewPosition", C.c_float * 3), ("time", C.c_float), ("deltaSeconds", C.c_double),
                ("minGiOcclusionFactor", C.c_float), ("minRoughness", C.c_float), ("usePerceptualRoughness", C.c_int32),
                ("fogDensity", C.c_float), ("fogColor", C.c_float * 3), ("infiniteLightDirection", C.c_float * 3),
                ("infiniteLightColor", C.c_float * 3), ("cascadeProjViews", (C.c_float * 16) * NUM_CASCADES),
                ("cascadePlanes", (C.c_float * 4) * 3), ("totalNumLights", C.c_int32), ("gatherMode", C.c_int32)]

    def copy(self):
        out = FrameParams()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(FrameParams))
        return out


class Buffers(C.Structure):
    _fields_ = [("depth", C.c_void_p), ("normal", C.c_void_p), ("albedo", C.c_void_p), ("baseReflectivity", C.c_void_p),
                ("roughnessMetallicAmbient", C.c_void_p), ("velocity", C.c_void_p), ("id", C.c_void_p),
                ("prevNormal", C.c_void_p), ("prevId", C.c_void_p), ("prevDepth", C.c_void_p),
                ("ssao", C.c_void_p), ("screen", C.c_void_p), ("lightingOut", C.c_void_p),
                ("vplData", C.c_void_p), ("vplUpdatedData", C.c_void_p), ("vplVisibleIndices", C.c_void_p),
                ("shadowDiffuseIndices", C.c_void_p), ("vplStage1Results", C.c_void_p), ("vplVisiblePerTile", C.c_void_p),
                ("atlasEntryByVpl", C.c_void_p),
                ("shadowCubeMaps", C.c_void_p * MAX_TOTAL_SHADOW_ATLASES),
                ("diffuseCubeMaps", C.c_void_p * MAX_TOTAL_SHADOW_ATLASES),
                ("infiniteLightShadowMap", C.c_void_p), ("vplCountDevice", C.c_void_p)]


MAX_PREPASS_LIGHTS = 262144


class LightRecord(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("radius", C.c_float), ("farPlane", C.c_float), ("intensity", C.c_float),
                ("atlasIndex", C.c_int32), ("atlasLayer", C.c_int32)]


class PointLight(C.Structure):
    _fields_ = [("position", C.c_float * 4), ("color", C.c_float * 4), ("radius", C.c_float), ("farPlane", C.c_float), ("_pad", C.c_float * 2)]


class DirectParams(C.Structure):
    _fields_ = [("numLights", C.c_int32), ("numShadowLights", C.c_int32), ("sunEnabled", C.c_int32), ("_pad", C.c_int32),
                ("nonShadowCasters", C.c_void_p), ("shadowCasters", C.c_void_p), ("shadowIndices", C.c_void_p),
                ("emissionStrength", C.c_float), ("infiniteLightDepthBias", C.c_float), ("infiniteLightZnear", C.c_float), ("infiniteLightZfar", C.c_float),
                ("pointLightAmbientIntensity", C.c_float), ("_pad2", C.c_float * 3), ("screenOut", C.c_void_p)]


class SsaoParams(C.Structure):
    _fields_ = [("aspectRatio", C.c_float), ("projPlaneZDist", C.c_float), ("intensity", C.c_float), ("_pad", C.c_float),
                ("structureBuffer", C.c_void_p), ("rotationLookup", C.c_void_p), ("occlusionScratch", C.c_void_p), ("ssaoOut", C.c_void_p)]


class TaaParams(C.Structure):
    _fields_ = [("screen", C.c_void_p), ("prevScreen", C.c_void_p), ("velocity", C.c_void_p), ("colorOut", C.c_void_p), ("historyOut", C.c_void_p)]


MAX_SCENE_BOXES = 256


class SceneBox(C.Structure):
    _fields_ = [("lo", C.c_float * 3), ("hi", C.c_float * 3), ("albedo", C.c_uint8 * 4), ("insideOut", C.c_int32)]


class CubeRenderParams(C.Structure):
    _fields_ = [("numBoxes", C.c_int32), ("numMaps", C.c_int32), ("boxes", C.c_void_p), ("vplIndices", C.c_void_p), ("stream", C.c_void_p)]


class PrepassParams(C.Structure):
    _fields_ = [("viewPosition", C.c_float * 3), ("numLights", C.c_int32), ("frustumPlanes", (C.c_float * 4) * 6),
                ("lights", C.c_void_p), ("vplDataOut", C.c_void_p), ("atlasEntryOut", C.c_void_p), ("orderOut", C.c_void_p),
                ("countOut", C.c_void_p)]


HALO_ROWS = 20
PEER_UP, PEER_DOWN, PEER_ROOT = 0, 1, 2


class BandParams(C.Structure):
    _fields_ = [("compute_y0", C.c_int32), ("compute_rows", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
                ("owns_frame_image", C.c_int32), ("halo_mode", C.c_int32)]


HALO_EXCHANGE, HALO_RECOMPUTE = 0, 1      # SgiBandParams.halo_mode


class PeerInfo(C.Structure):
    _fields_ = [("mem_handle", C.c_uint8 * 64), ("local_base", C.c_uint64), ("off_res", C.c_int64 * 2), ("off_gi", C.c_int64 * 2),
                ("off_hist", C.c_int64 * 2), ("off_flags", C.c_int64), ("off_image", C.c_int64 * 2),
                ("y_offset", C.c_int32), ("rows", C.c_int32), ("compute_y0", C.c_int32), ("compute_rows", C.c_int32),
                ("rank", C.c_int32), ("device", C.c_int32), ("pid", C.c_int32), ("_pad", C.c_int32)]


ABI_SYMBOLS = [
    "sgi_create", "sgi_destroy", "sgi_resize", "sgi_bind_device_buffers", "sgi_register_gl",
    "sgi_band_setup", "sgi_band_export", "sgi_band_attach", "sgi_frame_image", "sgi_band_timeouts", "sgi_band_reprojection_misses", "sgi_band_stats",
    "sgi_vpl_prepass", "sgi_render_vpl_cubemaps", "sgi_direct_lighting", "sgi_ssao", "sgi_taa", "sgi_cull", "sgi_early_begin", "sgi_cull_early", "sgi_color", "sgi_tile_lists", "sgi_gather", "sgi_denoise_pass", "sgi_denoise", "sgi_frame",
    "sgi_visible_count", "sgi_get_plane", "sgi_download_plane", "sgi_reset_history", "sgi_synchronize", "sgi_stream",
    "sgi_launch_count", "sgi_last_error", "sgi_version",
]
