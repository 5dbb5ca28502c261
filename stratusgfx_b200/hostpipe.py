"""Frames from HOST buffers, pipelined: what a caller whose G-buffer lives in host memory uses.

The reference's G-buffer never leaves the GPU, so it has no counterpart of this module; it exists
because the path's end-to-end figure (bench.py `e2e`) is defined with host buffers.  Per frame i:

    copy-in stream : pinned host -> device slot i%3  (G-buffer planes, VPL table, atlas entries, CSM,
                     the <= 5 cube maps re-rendered this frame, StratusRendererBackend.h:378)
    pass stream    : refreshed maps -> atlas, bind slot, sgi_frame (cull..denoise)   [the handle's stream]
    copy-out stream: lit image (lightingOut of slot i%3) -> pinned host

Three device slots rotate so the copy-in of frame i+1 overlaps the pass of frame i.  The previous
frame's normal/id/depth planes are never uploaded again: slot (i-1)%3 still holds them, which is the
reference's current/previous G-buffer swap (StratusRendererBackend.cpp:818-820).
"""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import Buffers, MAX_TOTAL_SHADOW_ATLASES

_GBUFFER = ("depth", "normal", "albedo", "baseReflectivity", "roughnessMetallicAmbient", "velocity", "id", "ssao", "screen")
_TABLES = ("vplData", "atlasEntryByVpl", "infiniteLightShadowMap")
_PREV = (("prevNormal", "normal"), ("prevId", "id"), ("prevDepth", "depth"))
SLOTS = 3


def pin_frame(torch, gb, vpl_data, atlas_by_vpl, csm):
    """One frame's host inputs (numpy / scene.GBuffer) as a dict of pinned tensors."""
    get = (lambda k: getattr(gb, k)) if not isinstance(gb, dict) else (lambda k: gb[k])
    out = {k: torch.from_numpy(np.ascontiguousarray(get(k))).pin_memory() for k in _GBUFFER}
    out["vplData"] = torch.from_numpy(np.ascontiguousarray(vpl_data)).pin_memory()
    out["atlasEntryByVpl"] = torch.from_numpy(np.ascontiguousarray(atlas_by_vpl)).pin_memory()
    out["infiniteLightShadowMap"] = torch.from_numpy(np.ascontiguousarray(csm)).pin_memory()
    return out


def pin_maps(torch, slots, depth_maps, diffuse_maps):
    """Refreshed cube maps: slot indices (atlas*320+layer), depth [k,6,R,R] uint16, diffuse [k,6,R,R,4] uint8."""
    return (list(slots), torch.from_numpy(np.ascontiguousarray(depth_maps).view(np.int16)).pin_memory(),
            torch.from_numpy(np.ascontiguousarray(diffuse_maps)).pin_memory())


class HostFramePipeline:
    def __init__(self, gp, max_refreshed_maps=5):
        """gp: a VplGiPass created with use_torch_stream=True whose atlases are already on the device (set_atlases)."""
        torch = gp.torch
        self.gp, self.torch = gp, torch
        dev = gp.device
        ptr = gp.stream_ptr()
        if ptr in (None, 0, 1, 2):
            raise ValueError("HostFramePipeline needs a pass that runs on a real stream (not the legacy/per-thread default stream)")
        self.pass_stream = torch.cuda.ExternalStream(ptr, device=dev)
        self.in_stream = torch.cuda.Stream(device=dev)
        self.out_stream = torch.cuda.Stream(device=dev)
        self.slots = [None] * SLOTS            # device tensors per slot, allocated on first submit
        self.buffers = [None] * SLOTS
        R = gp.cube_res
        self.map_stage = [(torch.empty((max_refreshed_maps, 6, R, R), dtype=torch.int16, device=dev),
                           torch.empty((max_refreshed_maps, 6, R, R, 4), dtype=torch.uint8, device=dev)) for _ in range(SLOTS)]
        self.out_dev = [torch.zeros((gp.rows, gp.width, 4), dtype=torch.float16, device=dev) for _ in range(SLOTS)]
        self.out_host = [torch.empty((gp.rows, gp.width, 4), dtype=torch.float16).pin_memory() for _ in range(SLOTS)]
        self.h2d_done, self.pass_done, self.d2h_done = {}, {}, {}
        self.count = 0
        self.h2d_bytes = 0
        self.d2h_bytes = self.out_host[0].numel() * 2

    def _alloc_slot(self, s, host):
        torch, dev = self.torch, self.gp.device
        self.slots[s] = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in host.items()}

    def _bind_struct(self, s):
        gp = self.gp
        cur, prev = self.slots[s], self.slots[(s - 1) % SLOTS]
        b = Buffers()
        for k in _GBUFFER + _TABLES:
            setattr(b, k, cur[k].data_ptr())
        for dst, src in _PREV:
            setattr(b, dst, prev[src].data_ptr())
        b.lightingOut = self.out_dev[s].data_ptr()
        for k in ("vplUpdatedData", "shadowDiffuseIndices", "vplStage1Results", "vplVisiblePerTile", "vplVisibleIndices"):
            setattr(b, k, gp.t[k].data_ptr())
        sa, da = gp.t["shadowAtlas"], gp.t["diffuseAtlas"]
        for a in range(min(sa.shape[0], MAX_TOTAL_SHADOW_ATLASES)):
            b.shadowCubeMaps[a] = sa[a].data_ptr()
            b.diffuseCubeMaps[a] = da[a].data_ptr()
        return b

    def submit(self, fp, host, refreshed_maps=None, first_prev=None):
        """Enqueue frame `self.count`.  `host`: pin_frame(...) dict; `refreshed_maps`: pin_maps(...) or None;
        `first_prev`: pinned dict with normal/id/depth of the frame before the first one (defaults to `host`)."""
        torch, gp = self.torch, self.gp
        i = self.count
        s = i % SLOTS
        if self.slots[s] is None:
            for k in range(SLOTS):
                self._alloc_slot(k, host)
            n = host["vplData"].shape[0]
            if "vplVisibleIndices" not in gp.t or gp.t["vplVisibleIndices"].numel() != n + 1:
                gp.t["vplVisibleIndices"] = torch.zeros(n + 1, dtype=torch.int32, device=gp.device)
            torch.cuda.synchronize(gp.device)
            self.buffers = [self._bind_struct(k) for k in range(SLOTS)]
        with torch.cuda.stream(self.in_stream):
            if i - 2 in self.pass_done:                       # slot s was `cur` of frame i-3 and `prev` of frame i-2
                self.in_stream.wait_event(self.pass_done[i - 2])
            nbytes = 0
            if i == 0:
                src = first_prev if first_prev is not None else host
                for _, k in _PREV:
                    self.slots[(s - 1) % SLOTS][k].copy_(src[k], non_blocking=True)
            for k, v in host.items():
                self.slots[s][k].copy_(v, non_blocking=True)
                nbytes += v.numel() * v.element_size()
            nmaps = 0
            if refreshed_maps is not None:
                idx, dep, dif = refreshed_maps
                nmaps = len(idx)
                self.map_stage[s][0][:nmaps].copy_(dep, non_blocking=True)
                self.map_stage[s][1][:nmaps].copy_(dif, non_blocking=True)
                nbytes += dep.numel() * 2 + dif.numel()
            ev = torch.cuda.Event()
            ev.record(self.in_stream)
            self.h2d_done[i] = ev
            self.h2d_bytes = nbytes
        with torch.cuda.stream(self.pass_stream):
            self.pass_stream.wait_event(self.h2d_done[i])
            if i - SLOTS in self.d2h_done:                    # lightingOut of this slot must have left the device
                self.pass_stream.wait_event(self.d2h_done[i - SLOTS])
            if nmaps:
                R = gp.cube_res
                sa = gp.t["shadowAtlas"].view(-1, 6, R, R)
                da = gp.t["diffuseAtlas"].view(-1, 6, R, R, 4)
                for j, m in enumerate(refreshed_maps[0]):
                    sa[m].copy_(self.map_stage[s][0][j], non_blocking=True)
                    da[m].copy_(self.map_stage[s][1][j], non_blocking=True)
            gp._check(gp.L.sgi_bind_device_buffers(gp._h, C.byref(self.buffers[s])))
            gp.frame(fp)
            ev = torch.cuda.Event()
            ev.record(self.pass_stream)
            self.pass_done[i] = ev
        with torch.cuda.stream(self.out_stream):
            self.out_stream.wait_event(self.pass_done[i])
            self.out_host[s].copy_(self.out_dev[s], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.out_stream)
            self.d2h_done[i] = ev
        for d in (self.h2d_done, self.pass_done, self.d2h_done):
            d.pop(i - 2 * SLOTS, None)
        self.count += 1
        return i

    def result(self, ticket):
        """Block until frame `ticket`'s lit image is in host memory; valid until frame ticket+3 is submitted."""
        self.d2h_done[ticket].synchronize()
        return self.out_host[ticket % SLOTS]
