"""Host side of the SSAO passes (first half): the inputs RenderSsaoOcclude_ / RenderSsaoBlur_
(StratusRendererBackend.cpp:1160-1204) hand to ssao.fs and ssao_blur.fs.

  rotation_table()        the 4x4 RGB16F texture of InitSSAO_ (BE:759-779), as half4 [4, 4, 4]
  ssao_uniforms(W, H)     aspectRatio, projPlaneZDist = 1 / tan(fovy / 2), intensity = 5 (BE:1165-1170)
  structure_buffer(...)   a structure buffer for the procedural scenes: what pbr_geometry_pass.fs:61-75 writes per fragment —
                          (dFdx z, dFdy z, z with the low 13 mantissa bits cleared, the remainder) in RGBA16F — derived here from
                          the depth plane (z = distance of the reconstructed position from the viewer; derivatives per 2x2 quad,
                          as fragment hardware takes them)
"""
import math

import numpy as np


def rotation_table():
    from .scene import _load
    t = np.zeros((16, 3), np.float32)
    _load().sgs_ssao_rotation_table(t.ctypes.data)
    out = np.zeros((4, 4, 4), np.float16)
    out[..., :3] = t.reshape(4, 4, 3)
    return out


def ssao_uniforms(width, height):
    from .scene import _load
    fovy = math.radians(_load().sgs_fovy_degrees())
    return dict(aspectRatio=np.float32(width) / np.float32(height), projPlaneZDist=np.float32(1.0 / math.tan(fovy / 2.0)), intensity=np.float32(5.0))


def structure_buffer(fp, depth):
    """half4 [H, W, 4] from the frame's uniforms and its depth plane (float32 [H, W], window-space depth)."""
    H, W = depth.shape
    inv = np.array(list(fp.invProjectionView), np.float64).reshape(4, 4).T          # column-major in the ABI
    u = (np.arange(W) + 0.5) / W * 2.0 - 1.0
    v = (np.arange(H) + 0.5) / H * 2.0 - 1.0
    ndc = np.stack([np.broadcast_to(u[None, :], (H, W)), np.broadcast_to(v[:, None], (H, W)), depth.astype(np.float64) * 2.0 - 1.0,
                    np.ones((H, W))], axis=-1)
    wp = ndc @ inv.T
    pos = wp[..., :3] / wp[..., 3:4]
    z = np.linalg.norm(pos - np.array(list(fp.viewPosition), np.float64)[None, None, :], axis=-1).astype(np.float32)
    hi = (z.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    # derivatives per 2x2 quad: the difference inside the quad, shared by its pixels
    dx, dy = np.zeros_like(z), np.zeros_like(z)
    we, he = W - W % 2, H - H % 2
    d = z[:, 1:we:2] - z[:, 0:we:2]
    dx[:, 0:we:2] = d; dx[:, 1:we:2] = d
    d = z[1:he:2, :] - z[0:he:2, :]
    dy[0:he:2, :] = d; dy[1:he:2, :] = d
    out = np.zeros((H, W, 4), np.float16)
    with np.errstate(over="ignore"):
        out[..., 0], out[..., 1], out[..., 2], out[..., 3] = dx, dy, hi, z - hi
    return out
