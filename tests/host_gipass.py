"""TEST INFRASTRUCTURE: the Python host class (stratusgfx_b200.gipass.VplGiPass) over tests/cuda_on_host/_build_abi/libstratusgi_host.so —
every translation unit of the product library, the C ABI included, compiled for the host (lockstep kernels, CUDA-runtime stand-ins).
With it the call sequences of the GPU tests run where there is no GPU: the state machine of csrc/sgi_abi.cu (fused pass 0, guide
plane, history swap, bind resets) and the test code itself are exercised before a GPU minute is spent.  Nothing in the product
imports this module; the product class refuses to start without a CUDA device."""
import contextlib
import ctypes as C
import os
import types

import numpy as np

from stratusgfx_b200 import _abi, gipass
from stratusgfx_b200._abi import Buffers, CreateParams, MAX_TOTAL_VPLS_PER_FRAME, TILE_DIVISOR

_host_lib = None


def host_library():
    global _host_lib
    if _host_lib is None:
        from tests.cuda_on_host.host_tu import compile_whole_library
        path = os.environ.get("SGI_HOST_LIB_OVERRIDE") or compile_whole_library()
        saved = gipass._LIB_PATH, gipass._lib
        try:
            gipass._LIB_PATH, gipass._lib = path, None
            _host_lib = gipass.load_library()              # the same argtypes / restypes as the product library
        finally:
            gipass._LIB_PATH, gipass._lib = saved
    return _host_lib


class _TorchOnHost:
    """torch with the two `torch.cuda.synchronize` calls of the host class turned into no-ops (there is no second stream here)."""

    def __init__(self, torch):
        self._torch = torch
        self.cuda = types.SimpleNamespace(synchronize=lambda *a, **k: None)

    def __getattr__(self, name):
        return getattr(self._torch, name)


class HostGiPass(gipass.VplGiPass):
    def __init__(self, width, full_height, y_offset=0, rows=None, cube_res=32, csm_res=1024, denoise_iters=4, device=0, use_torch_stream=False):
        import torch
        self.torch = _TorchOnHost(torch)
        self.L = host_library()
        self.device = torch.device("cpu")
        rows = full_height - y_offset if rows is None else rows
        self.width, self.full_height, self.y_offset, self.rows = width, full_height, y_offset, rows
        self.cube_res, self.csm_res, self.denoise_iters = cube_res, csm_res, denoise_iters
        cp = CreateParams(width, full_height, y_offset, rows, cube_res, csm_res, denoise_iters, 0, None)
        self._h = C.c_void_p()
        rc = self.L.sgi_create(C.byref(cp), C.byref(self._h))
        if rc != _abi.SGI_OK:
            raise gipass.StratusGiError(f"sgi_create failed with status {rc}")
        self.t = {}
        self.buffers = Buffers()
        tiles = (width // TILE_DIVISOR) * (rows // TILE_DIVISOR)
        self.t["vplUpdatedData"] = torch.zeros((MAX_TOTAL_VPLS_PER_FRAME, 16), dtype=torch.float32)
        self.t["shadowDiffuseIndices"] = torch.zeros((MAX_TOTAL_VPLS_PER_FRAME, 2), dtype=torch.int32)
        self.t["vplStage1Results"] = torch.zeros((max(tiles, 1), 8), dtype=torch.float32)
        self.t["vplVisiblePerTile"] = torch.zeros((max(tiles, 1), 13), dtype=torch.int32)
        self.t["lightingOut"] = torch.zeros((rows, width, 4), dtype=torch.float16)
        self.tiles = tiles

    def frame_image(self):
        ptr, nbytes = C.c_void_p(), C.c_int64()
        self._check(self.L.sgi_frame_image(self._h, C.byref(ptr), C.byref(nbytes)))
        raw = np.ctypeslib.as_array((C.c_uint16 * (nbytes.value // 2)).from_address(ptr.value))
        return self.torch.from_numpy(raw.view(np.float16).reshape(self.full_height, self.width, 4).copy())

    def output(self, name):
        return super().output(name).copy()          # (.cpu().numpy() of a CPU tensor is a view of the live buffer, not a download)


@contextlib.contextmanager
def gpu_tests_on_the_host():
    """Inside this context `VplGiPass` is the host-backed class, so tests.parity_util.make_pass and the bodies of the GPU tests run on the CPU."""
    saved = gipass.VplGiPass
    gipass.VplGiPass = HostGiPass
    try:
        yield
    finally:
        gipass.VplGiPass = saved
