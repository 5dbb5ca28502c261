"""Widening row f2 — VPL cube shadow / diffuse map production (BE:1426-1575, shadowVpl.fs) on the device through
sgi_render_vpl_cubemaps, against oracle.cpp::orc_render_vpl_cubemaps: every texel is an integer output, so the bar is bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def check_cubemaps(workload, which):
    from oracle import oracle_py as O
    from tests.parity_util import SceneData, make_pass
    d = SceneData(workload, frames=1)
    s = d.scene
    boxes = s.boxes()
    N = s.num_vpls
    idx = np.arange(N, dtype=np.int32) if which == "all" else np.array([0, N // 3, N // 2, N - 2, N - 1][:min(5, N)], np.int32)
    want_d = np.full_like(d.shadow_atlas, 0x1234); want_c = np.full_like(d.diffuse_atlas, 0x56)
    O.render_vpl_cubemaps(boxes, d.vpl_data, d.atlas_by_vpl, idx, s.cube_res, want_d, want_c)
    gp = make_pass(d)
    gp.t["shadowAtlas"].fill_(0x1234); gp.t["diffuseAtlas"].fill_(0x56)      # only the listed layers may change
    gp.set_scene_boxes(boxes)
    gp.bind()
    gp.render_vpl_cubemaps(idx)
    gp.synchronize()
    got_d = gp.t["shadowAtlas"].cpu().numpy().view(np.uint16).reshape(want_d.shape)
    got_c = gp.t["diffuseAtlas"].cpu().numpy().reshape(want_c.shape)
    gp.close()
    assert np.array_equal(got_d, want_d)
    assert np.array_equal(got_c, want_c)
    # and the maps are the scene generator's (double precision, the atlases every other test reads) up to one unorm16 step on the
    # few texels where float and double round apart; the surface hit is the same except on box silhouettes
    slots = d.atlas_by_vpl[idx, 0] * 320 + d.atlas_by_vpl[idx, 1]
    a = got_d.reshape(-1, 6, s.cube_res, s.cube_res)[slots].astype(np.int32)
    b = d.shadow_atlas.reshape(-1, 6, s.cube_res, s.cube_res)[slots].astype(np.int32)
    assert np.mean(a == b) > 0.99 and np.mean(np.abs(a - b) <= 1) > 0.999
    ca = got_c.reshape(-1, 6, s.cube_res, s.cube_res, 4)[slots]; cb = d.diffuse_atlas.reshape(-1, 6, s.cube_res, s.cube_res, 4)[slots]
    assert np.mean(np.all(ca == cb, axis=-1)) > 0.999


@pytest.mark.parametrize("workload,which", [("tiny", "all"), ("c1", "all"), ("c3", "five"), ("c3", "all")])
def test_cube_maps_rendered_on_the_device_equal_the_oracle(built, workload, which):
    check_cubemaps(workload, which)


def test_rendered_maps_feed_the_frame(built):
    """A frame whose refreshed maps were rendered on the device: the path downstream (colouring, gather, denoise) still meets the
    oracle run on the same (device-rendered = oracle-rendered) atlases."""
    from oracle import oracle_py as O
    from tests.parity_util import SceneData, compare_frame, make_pass, run_oracle
    d = SceneData("tiny", frames=1)
    s = d.scene
    idx = np.arange(s.num_vpls, dtype=np.int32)
    O.render_vpl_cubemaps(s.boxes(), d.vpl_data, d.atlas_by_vpl, idx, s.cube_res, d.shadow_atlas, d.diffuse_atlas)      # the oracle's atlases
    orc = run_oracle(d)[0]
    gp = make_pass(d)
    gp.t["shadowAtlas"].zero_(); gp.t["diffuseAtlas"].zero_()
    gp.set_scene_boxes(s.boxes())
    gp.set_gbuffer(d.gb[0], d.gb[-1]); gp.bind()
    gp.render_vpl_cubemaps(idx)
    gp.frame(d.fp[0])
    got = dict(visible=gp.output("vplVisibleIndices"), combined=gp.output("lightingOut"))
    gp.close()
    assert np.array_equal(got["visible"][: orc["visible"][0] + 1], orc["visible"][: orc["visible"][0] + 1])
    from tests.parity_util import compare_radiance
    r = compare_radiance(got["combined"], orc["combined"], 3)
    assert r["ok"] and r["bit_exact"] > 0.99, r
