"""Host side of the VPL pre-pass (include/stratusgi.h: sgi_vpl_prepass) — SURVEY.md section 8f, rank 1.

Mirrors what RendererBackend::UpdatePointLights_ does for virtual point lights before the cull
(Source/Engine/StratusRendererBackend.cpp:1359-1392) and the table upload of InitVplFrameData_
(:1281-1292): frustum test per light sphere, order by distance to the viewer, keep the nearest
10000.  The work runs on the device; this module only lays out the inputs.
"""
import numpy as np

# SgiLightRecord: what the engine keeps per VirtualPointLight plus its slot in the shadow-map cache
LIGHT_DTYPE = np.dtype([("position", np.float32, 3), ("radius", np.float32), ("farPlane", np.float32), ("intensity", np.float32),
                        ("atlasIndex", np.int32), ("atlasLayer", np.int32)])
assert LIGHT_DTYPE.itemsize == 32


def light_records(vpl_data, atlas_by_vpl):
    """SgiLightRecord array from a VplData table ([N,16] f32) and its atlas entries ([N,2] i32)."""
    n = vpl_data.shape[0]
    out = np.zeros(n, LIGHT_DTYPE)
    out["position"] = vpl_data[:, 0:3]
    out["radius"] = vpl_data[:, 12]
    out["farPlane"] = vpl_data[:, 13]
    out["intensity"] = vpl_data[:, 14]
    out["atlasIndex"] = atlas_by_vpl[:, 0]
    out["atlasLayer"] = atlas_by_vpl[:, 1]
    return out


def frustum_planes(inv_projection_view):
    """The six planes the frontend hands over as frame_->viewFrustumPlanes (StratusRendererFrontend.cpp:1075-1089):
    rows of the projection-view matrix combined as left, right, bottom, top, near, far; not normalised."""
    inv_pv = np.asarray(inv_projection_view, np.float64).reshape(4, 4).T          # column-major storage -> matrix
    pv = np.linalg.inv(inv_pv)
    r = [pv[i, :] for i in range(4)]                                               # vpt[i] = column i of transpose(pv) = row i of pv
    planes = [r[3] + r[0], r[3] - r[0], r[3] + r[1], r[3] - r[1], r[3] + r[2], r[3] - r[2]]
    return np.asarray(planes, np.float32)
