"""GPU parity tests: the CUDA path, called through the C-ABI, against the CPU oracle on the same
seeded inputs.  Integer outputs bit-exact; radiance within 1e-3 relative per pixel or PSNR >= 60 dB
(BASELINE.json north_star; tolerances applied in tests/parity_util.py)."""
import ctypes as C

import numpy as np
import pytest

from stratusgfx_b200 import _abi
from tests.parity_util import (SceneData, band_of, compare_frame, compare_radiance, make_pass, run_cuda_staged, run_oracle)

pytestmark = pytest.mark.gpu


def _assert_parity(cu, orc):
    for f, (a, b) in enumerate(zip(cu, orc)):
        rep, bad = compare_frame(a, b)
        assert not bad, f"frame {f}: {bad}"
        # beyond the tolerance bar: the two paths should agree to the last half bit almost everywhere
        assert rep["combined"]["bit_exact"] > 0.99


def test_native_library_is_the_one_running():
    import torch
    assert torch.cuda.is_available()
    from stratusgfx_b200.gipass import load_library, library_path
    L = load_library()
    assert b"sm_100a" in L.sgi_version()
    maps = open("/proc/self/maps").read()
    assert library_path() in maps


def test_tiny_two_frames_stage_by_stage(tiny, tiny_oracle):
    _assert_parity(run_cuda_staged(tiny), tiny_oracle)


def test_c1_cornell_640x360_64vpl(built):
    d = SceneData("c1", frames=1)
    _assert_parity(run_cuda_staged(d), run_oracle(d))


def test_c2_cornell_1080p_256vpl_four_temporal_frames(built):
    d = SceneData("c2", frames=4)
    cu, orc = run_cuda_staged(d), run_oracle(d)
    _assert_parity(cu, orc)
    for t, o in enumerate(cu, start=1):
        assert np.all(o["history"] == np.float16(t))           # static camera: history counts frames


def test_moving_camera_reprojection(built):
    d = SceneData("c1", frames=3, yaw_deg_per_frame=0.25)
    cu, orc = run_cuda_staged(d), run_oracle(d)
    _assert_parity(cu, orc)
    h = orc[2]["history"].astype(np.float32)
    assert h.min() == 1.0 and h.max() == 3.0                    # some pixels were disoccluded, some kept their history


def test_c3_hall_1080p_1024vpl_five_denoise_iterations(built):
    d = SceneData("c3", frames=1)
    assert d.scene.denoise_iters == 5
    _assert_parity(run_cuda_staged(d), run_oracle(d))


@pytest.mark.parametrize("iters", [1, 2, 3])
def test_denoise_iteration_count_is_a_parameter(tiny, iters):
    _assert_parity(run_cuda_staged(tiny, frames=1, denoise_iters=iters), run_oracle(tiny, frames=1, denoise_iters=iters))


def test_tile_list_gather_mode(built):
    """SGI_GATHER_TILE_LIST: every VPL of the pixel's 5x5-tile list instead of random draws (the tile lookup that is
    commented out at vpl_pbr_gi.fs:86-97, composed with the live per-light body); deterministic, lists staged through
    shared memory.  Checked against the oracle's restatement of the same composition, through the denoiser."""
    for wl, frames in (("tiny", 2), ("c1", 1)):
        d = SceneData(wl, frames=frames)
        for fp in d.fp:
            fp.gatherMode = _abi.GATHER_TILE_LIST
        cu, orc = run_cuda_staged(d), run_oracle(d)
        _assert_parity(cu, orc)
        assert orc[0]["sample_stats"][1] > 0                      # lights were shaded
        w = cu[0]["gi_reservoir"][..., 3].astype(np.float32)
        assert w.min() >= 1.0 and w.max() <= 12.0                  # at most 12 VPLs per tile, weights re-scaled into [minGi, 1] each
    # the mode needs the tile lists of this frame
    gp = make_pass(d)
    gp.set_gbuffer(d.gb[0], d.gb[-1]); gp.bind()
    gp.cull(d.fp[0]); gp.color(d.fp[0])
    with pytest.raises(Exception, match="sgi_tile_lists"):
        gp.gather(d.fp[0])
    gp.close()


def check_tile_list_band(workload, y0, rows):
    """One band of `workload` in SGI_GATHER_TILE_LIST mode, stage by stage against the oracle on the same band (the tile
    lists are per tile and the gather per pixel, so a band is the whole frame's rows).  -> (report, stage-2 counts)"""
    from oracle.oracle_py import OracleFrame
    from stratusgfx_b200.gipass import VplGiPass
    from stratusgfx_b200.scene import Scene
    s = Scene(workload)
    W, H = s.width, s.height
    vd, at = s.vpls(); sh, df = s.cubemaps(); csm = s.csm(0); fp = s.frame_params(0)
    fp.gatherMode = _abi.GATHER_TILE_LIST
    gb, pgb = s.gbuffer(0, y0, rows), s.gbuffer(-1, y0, rows)
    gp = VplGiPass(W, H, y0, rows, s.cube_res, s.csm_res, s.denoise_iters)
    gp.set_vpls(vd, at); gp.set_atlases(sh, df); gp.set_csm(csm); gp.set_gbuffer(gb, pgb); gp.bind()
    gp.cull(fp); V = gp.visible_count(); gp.color(fp); gp.tile_lists(fp); gp.gather(fp)
    res = gp.plane(_abi.PLANE_GI_RESERVOIR)
    gp.denoise(fp)
    cu = dict(V=V, visible=gp.output("vplVisibleIndices"), updated=gp.output("vplUpdatedData"), entries=gp.output("shadowDiffuseIndices"),
              stage1=gp.output("vplStage1Results"), stage2=gp.output("vplVisiblePerTile"), gi_color=gp.plane(_abi.PLANE_GI_COLOR),
              gi_reservoir=res, combined=gp.output("lightingOut"), gi_illum=gp.plane(_abi.PLANE_GI_ILLUM),
              history=gp.plane(_abi.PLANE_HISTORY), pass_reservoirs=[])
    gp.close()
    o = OracleFrame(W, rows, s.cube_res, s.csm_res, s.denoise_iters, y_offset=y0, full_height=H)
    o.frame(fp, gb, pgb, vd, at, csm, sh, df)
    orc = dict(visible=o.visible, updated=o.updated, entries=o.entries, stage1=o.stage1, stage2=o.stage2, gi_color=o.gi_color,
               gi_reservoir=o.gi_reservoir, pass_reservoirs=[], combined=o.combined, gi_illum=o.gi_illum, history=o.history)
    rep, bad = compare_frame(cu, orc)
    assert not bad, bad
    assert rep["stage2_lists_exact"]
    w = res[..., 3].astype(np.float32)
    assert w.min() >= 1.0 and w.max() <= 12.0
    return rep, cu["stage2"][:, 0]


def test_tile_list_gather_mode_with_full_lists(built):
    """The tile-list mode where the lists are long: c3 (1920x1080, 1024 VPLs) and c4 (3840x2160, 2048 VPLs), a band of the
    floor / column rows each.  Every tile holds the full 12 entries there, so the shared-memory staging of a CTA's 8 x 2 tile
    lists and the 12-light shading loop see the lengths that tiny and c1 never reach (27 tile rows of the c3 band: a partial CTA row)."""
    for wl, y0, rows in (("c3", 700, 135), ("c4", 1400, 90)):
        rep, counts = check_tile_list_band(wl, y0, rows)
        assert counts.max() == 12 and np.mean(counts == 12) > 0.5, (wl, np.bincount(counts, minlength=13))


@pytest.mark.parametrize("width,height,tile_list_mode", [(170, 95, False), (165, 95, False), (1370, 770, False), (170, 95, True)])
def test_ragged_frame_sizes(built, width, height, tile_list_mode):
    """Sizes that are multiples of the 5-pixel tile only: 170 and 1370 columns leave 2 columns beyond the last whole multiple of
    the à-trous spacings 4 and 8 (they and the 2 leftover rows of 770 go to the border CTAs next to the TMA tiles), an odd width
    has no 16-byte row pitch for a tensor map (every pass through the clamped global-memory kernel), and none of the sizes is a
    multiple of the 32-pixel segments of the gather or of the 8-row CTAs of the temporal pre-pass.  Two frames of a turning camera.
    tile_list_mode: 34 x 19 tiles are 4.25 x 9.5 CTAs of the tile-list gather (8 x 2 tiles each): partial CTAs at both edges."""
    d = SceneData("tiny", frames=2, yaw_deg_per_frame=1.0, width=width, height=height)
    if tile_list_mode:
        for fp in d.fp:
            fp.gatherMode = _abi.GATHER_TILE_LIST
    for f, (cu, orc) in enumerate(zip(run_cuda_staged(d), run_oracle(d))):
        rep, bad = compare_frame(cu, orc)                  # integer outputs and history bit-exact, radiance within the bar
        assert not bad, f"frame {f}: {bad}"


def test_fused_frame_call_equals_staged_calls(tiny):
    staged = run_cuda_staged(tiny)
    gp = make_pass(tiny)
    for f in range(2):
        gp.set_gbuffer(tiny.gb[f], tiny.gb[f - 1]); gp.bind()
        gp.frame(tiny.fp[f])
        assert np.array_equal(gp.output("lightingOut").view(np.uint16), staged[f]["combined"].view(np.uint16))
        assert np.array_equal(gp.output("vplVisiblePerTile"), staged[f]["stage2"])
        assert np.array_equal(gp.plane(_abi.PLANE_HISTORY).view(np.uint16), staged[f]["history"].view(np.uint16))
    gp.close()


def _second_stream(gp):
    """A second CUDA stream for the early-cull calls: a real one on a GPU; on the host-built library (tests/host_gipass.py) streams
    are tokens and launches are synchronous, so any non-null value does."""
    if gp.device.type != "cuda":
        return None, 0x2
    st = gp.torch.cuda.Stream(device=gp.device, priority=-1)
    return st, st.cuda_stream


def test_early_cull_gives_the_frames_of_the_serial_order(tiny):
    """sgi_early_begin / sgi_cull_early: the cull + colouring of frame f+1 enqueued on a second stream right after frame f (the
    reference's call order, BE:1395 before BE:2043) — the frames must be the bytes of the serial order, the visible lists and the
    per-tile lists too; a frame whose parameters differ from the early cull's is refused; an explicit cull supersedes a pending
    early one."""
    from stratusgfx_b200.gipass import StratusGiError
    staged = run_cuda_staged(tiny)
    gp = make_pass(tiny)
    st, ptr = _second_stream(gp)
    gp.set_gbuffer(tiny.gb[0], tiny.gb[-1]); gp.set_csm(tiny.csm[0]); gp.bind()
    for f in range(3):                                     # frame 0 culls itself; 1 and 2 consume an early cull (with, without colouring)
        gp.frame(tiny.fp[f % 2])
        if f < 2:                                          # the staged run's frames (same history)
            assert np.array_equal(gp.output("lightingOut").view(np.uint16), staged[f]["combined"].view(np.uint16)), f
            nv = gp.output("vplVisibleIndices")
            assert np.array_equal(nv, staged[f]["visible"][: nv.shape[0]]), f
            assert np.array_equal(gp.output("vplVisiblePerTile"), staged[f]["stage2"]), f
        else:                                              # frame 0's inputs again: the same lists as frame 0
            assert np.array_equal(gp.output("vplVisiblePerTile"), staged[0]["stage2"])
        nf = (f + 1) % 2
        gp.synchronize()                                   # (the test swaps whole input planes between frames; a renderer's are per-frame resources)
        gp.set_gbuffer(tiny.gb[nf], tiny.gb[nf - 1]); gp.set_csm(tiny.csm[nf]); gp.bind()
        gp.early_begin(ptr)
        gp.cull_early(tiny.fp[nf], ptr, with_color=(f % 2 == 0))
    # refusals (an early cull for frame 1's inputs is pending here)
    with pytest.raises(StratusGiError):
        gp.cull_early(tiny.fp[1], ptr)                     # not consumed yet
    other = tiny.fp[1].copy(); other.totalNumLights -= 1
    with pytest.raises(StratusGiError):
        gp.frame(other)                                    # a different frame than the early cull was given
    gp.cull(tiny.fp[1])                                    # explicit cull: supersedes the pending one
    gp.synchronize()
    gp.close()


def test_host_pipeline_equals_device_resident_frames(built):
    """The pipelined host-buffer API (hostpipe.py: rotating device slots, previous-frame planes kept on the device,
    copies on side streams) must produce the bytes of the plain device-resident path, frame after frame."""
    import torch
    from stratusgfx_b200.hostpipe import HostFramePipeline, pin_frame
    d = SceneData("tiny", frames=5, yaw_deg_per_frame=0.25)
    staged = run_cuda_staged(d)
    gp = make_pass(d)
    pipe = HostFramePipeline(gp)
    first_prev = {k: torch.from_numpy(getattr(d.gb[-1], k)).pin_memory() for k in ("normal", "id", "depth")}
    for f in range(5):
        host = pin_frame(torch, d.gb[f], d.vpl_data, d.atlas_by_vpl, d.csm[f])
        tk = pipe.submit(d.fp[f], host, None, first_prev if f == 0 else None)
        out = pipe.result(tk).numpy()
        assert np.array_equal(out.view(np.uint16), staged[f]["combined"].view(np.uint16)), f
    assert pipe.h2d_bytes > 0 and pipe.d2h_bytes == d.W * d.H * 8
    gp.close()


def test_reference_method_names_drive_the_same_path(tiny):
    staged = run_cuda_staged(tiny, frames=1)
    gp = make_pass(tiny)
    gp.set_gbuffer(tiny.gb[0], tiny.gb[-1]); gp.bind()
    fp = tiny.fp[0]
    gp.PerformVirtualPointLightCullingStage1(fp)
    gp.PerformVirtualPointLightCullingStage2(fp)
    gp.ComputeVirtualPointLightGlobalIllumination(fp)
    assert np.array_equal(gp.output("lightingOut").view(np.uint16), staged[0]["combined"].view(np.uint16))
    gp.close()


def test_row_bands_with_apron_reproduce_the_whole_frame(built):
    """Multi-GPU split: a band computed with a 28-row apron (4+8+16 rows of à-trous reach)
    must equal the same rows of the whole-frame result bit for bit."""
    d = SceneData("c1", frames=1)
    whole = run_cuda_staged(d)[0]
    H, apron = d.H, 30
    for (b0, b1) in ((0, 180), (180, 360), (90, 270)):
        y0, y1 = max(0, b0 - apron), min(H, b1 + apron)
        gp = make_pass(d, y0, y1 - y0)
        gp.set_gbuffer(band_of(d.gb[0], y0, y1 - y0), band_of(d.gb[-1], y0, y1 - y0)); gp.bind()
        gp.frame(d.fp[0])
        out = gp.output("lightingOut")[b0 - y0:b1 - y0]
        assert np.array_equal(out.view(np.uint16), whole["combined"][b0:b1].view(np.uint16)), (b0, b1)
        t0, t1 = (b0 - y0) // 5 * (d.W // 5), (b1 - y0) // 5 * (d.W // 5)
        assert np.array_equal(gp.output("vplVisiblePerTile")[t0:t1], whole["stage2"][b0 // 5 * (d.W // 5):b1 // 5 * (d.W // 5)])
        gp.close()


def _run_p2p_bands_in_one_process(d, world, frames, halo_mode=0):
    """`world` band handles on one GPU in one process, linked through sgi_band_attach (same-process peer pointers)."""
    from stratusgfx_b200.bands import halo_rows, plan_bands, recompute_rows
    from stratusgfx_b200._abi import PEER_DOWN, PEER_ROOT, PEER_UP
    bands = plan_bands(d.H, world, (recompute_rows if halo_mode else halo_rows)(d.scene.denoise_iters))
    gps = [make_pass(d, b.y0, b.rows) for b in bands]
    for gp, b in zip(gps, bands):
        gp.band_setup(b.b0, b.b1 - b.b0, b.rank, world, owns_frame_image=(b.rank == 0), halo_mode=halo_mode)
    infos = [gp.band_export() for gp in gps]
    for gp, b in zip(gps, bands):
        if b.rank > 0: gp.band_attach(PEER_UP, infos[b.rank - 1])
        if b.rank + 1 < world: gp.band_attach(PEER_DOWN, infos[b.rank + 1])
        if b.rank != 0: gp.band_attach(PEER_ROOT, infos[0])
    images = []
    for f in range(frames):
        for gp, b in zip(gps, bands):
            gp.set_gbuffer(band_of(d.gb[f], b.y0, b.rows), band_of(d.gb[f - 1], b.y0, b.rows)); gp.set_csm(d.csm[f]); gp.bind()
        order = gps if f % 2 == 0 else gps[::-1]          # the enqueue order of the bands must not matter
        for gp in order:
            gp.frame(d.fp[f])
        for gp in gps:
            gp.synchronize()
        images.append(gps[0].frame_image().cpu().numpy())
    timeouts = [gp.band_timeouts() for gp in gps]
    tiles = [gp.output("vplVisiblePerTile") for gp in gps]
    for gp in gps:
        gp.close()
    return bands, images, timeouts, tiles


@pytest.mark.parametrize("world", [2, 3])
def test_p2p_row_bands_reproduce_the_whole_frame(built, world):
    """Peer-to-peer halo mode (csrc/sgi_band.cu): each band computes only its own rows; the neighbours' kernels store the
    boundary rows of every denoise pass (and of the GI/history planes) into its halo, the final pass stores the lit rows
    into band 0's frame image.  Three frames of a moving camera: the assembled image must equal the whole-frame bytes."""
    d = SceneData("c1", frames=3, yaw_deg_per_frame=0.25)
    whole = run_cuda_staged(d)
    bands, images, timeouts, tiles = _run_p2p_bands_in_one_process(d, world, 3)
    assert timeouts == [0] * world
    for f in range(3):
        assert np.array_equal(images[f].view(np.uint16), whole[f]["combined"].view(np.uint16)), f
    tx = d.W // 5
    for b, t in zip(bands, tiles):          # tile lists of the band's own tile rows
        r0, r1 = (b.b0 - b.y0) // 5, (b.b1 - b.y0) // 5
        assert np.array_equal(t[r0 * tx:r1 * tx], whole[2]["stage2"][b.b0 // 5 * tx:b.b1 // 5 * tx])


@pytest.mark.parametrize("world", [2, 3, 4])
def test_row_bands_that_recompute_their_halo_reproduce_the_whole_frame(built, world):
    """SGI_HALO_RECOMPUTE (include/stratusgi.h): a band computes the gather and the non-final passes on the halo rows its later
    passes read (28 / 24 / 16 rows) instead of receiving them; only the final pass's GI rows and the lit rows travel.  Same
    arithmetic on the same inputs, so three frames of a moving camera must give the whole-frame bytes; a halo shorter than
    28 rows is refused."""
    from stratusgfx_b200.bands import recompute_rows
    from stratusgfx_b200.gipass import StratusGiError
    assert recompute_rows(4) == 30 and recompute_rows(5) == 60 and recompute_rows(1) == 0
    d = SceneData("c1", frames=3, yaw_deg_per_frame=0.25)
    whole = run_cuda_staged(d)
    bands, images, timeouts, tiles = _run_p2p_bands_in_one_process(d, world, 3, halo_mode=1)
    assert timeouts == [0] * world
    for f in range(3):
        assert np.array_equal(images[f].view(np.uint16), whole[f]["combined"].view(np.uint16)), f
    tx = d.W // 5
    for b, t in zip(bands, tiles):
        r0, r1 = (b.b0 - b.y0) // 5, (b.b1 - b.y0) // 5
        assert np.array_equal(t[r0 * tx:r1 * tx], whole[2]["stage2"][b.b0 // 5 * tx:b.b1 // 5 * tx])
    gp = make_pass(d, 0, d.H // 2 + 20)                  # 20 halo rows: enough to exchange, not to recompute
    with pytest.raises(StratusGiError):
        gp.band_setup(0, d.H // 2, 0, 2, owns_frame_image=True, halo_mode=1)
    gp.close()


def test_row_bands_that_recompute_their_halo_five_iterations(built):
    d = SceneData("c1", frames=2, yaw_deg_per_frame=0.25)
    d.scene.denoise_iters = 5
    whole = run_cuda_staged(d)
    bands, images, timeouts, _ = _run_p2p_bands_in_one_process(d, 2, 2, halo_mode=1)
    assert timeouts == [0, 0]
    for f in range(2):
        assert np.array_equal(images[f].view(np.uint16), whole[f]["combined"].view(np.uint16)), f


def test_p2p_row_bands_five_iterations_need_a_wider_halo(built):
    from stratusgfx_b200.bands import halo_rows
    assert halo_rows(4) == 20 and halo_rows(5) == 35 and halo_rows(1) == 0
    d = SceneData("c1", frames=1)
    d.scene.denoise_iters = 5               # make_pass / run_cuda_staged read the iteration count from the scene
    whole = run_cuda_staged(d)
    bands, images, timeouts, _ = _run_p2p_bands_in_one_process(d, 2, 1)
    assert timeouts == [0, 0]
    assert np.array_equal(images[0].view(np.uint16), whole[0]["combined"].view(np.uint16))


# ---- edge cases ------------------------------------------------------------------------------------------
def test_single_vpl_and_two_vpls(built):
    for n in (1, 2):
        d = SceneData("tiny", frames=1, num_vpls=n)
        _assert_parity(run_cuda_staged(d), run_oracle(d))


def test_no_visible_vpl(built):
    d = SceneData("tiny", frames=1, num_vpls=8)
    d.vpl_data[:, 0:3] = (0.0, -30.0, 0.0)          # under the floor slab: the sun never reaches them
    cu, orc = run_cuda_staged(d), run_oracle(d)
    assert int(orc[0]["visible"][0]) == 0 and cu[0]["V"] == 0
    _assert_parity(cu, orc)
    assert np.all(cu[0]["gi_reservoir"][..., :3] == 0) and np.all(cu[0]["stage2"][:, 0] == 0)


def test_more_visible_vpls_than_the_per_frame_cap(built):
    d = SceneData("tiny", frames=1, num_vpls=6000)
    d.vpl_data[:, 0:3] = d.vpl_data[:, 0:3] * 0 + np.array([0.0, 200.0, -100.0], np.float32)   # high above everything: all lit
    d.vpl_data[:, 0] += np.linspace(-5, 5, 6000, dtype=np.float32)
    cu, orc = run_cuda_staged(d), run_oracle(d)
    assert int(orc[0]["visible"][0]) == _abi.MAX_TOTAL_VPLS_PER_FRAME == cu[0]["V"]
    _assert_parity(cu, orc)


def test_zero_lights_is_a_no_op(tiny):
    gp = make_pass(tiny)
    gp.set_gbuffer(tiny.gb[0], tiny.gb[-1]); gp.bind()
    fp = tiny.fp[0].copy(); fp.totalNumLights = 0
    before = gp.launch_count()
    gp.frame(fp)
    assert gp.launch_count() == before
    gp.close()


def test_unbound_buffers_and_bad_passes_are_errors(tiny):
    from stratusgfx_b200.gipass import StratusGiError, VplGiPass
    gp = VplGiPass(tiny.W, tiny.H)
    with pytest.raises(StratusGiError, match="not bound"):
        gp.frame(tiny.fp[0])
    gp2 = make_pass(tiny)
    gp2.set_gbuffer(tiny.gb[0], tiny.gb[-1]); gp2.bind()
    with pytest.raises(StratusGiError, match="out of range"):
        gp2.denoise_pass(tiny.fp[0], 9)
    fp = tiny.fp[0].copy(); fp.gatherMode = _abi.GATHER_TILE_LIST
    with pytest.raises(StratusGiError):
        gp2.gather(fp)
    assert gp2.L.sgi_register_gl(gp2._h, None, 0) == _abi.SGI_ERR_UNSUPPORTED
    gp.close(); gp2.close()


def test_frame_that_fails_in_its_last_pass_leaves_a_usable_handle(tiny):
    """sgi_frame without a `screen` plane fails in the final denoise pass, after the tile stages and the temporal pre-pass were
    forked onto their side streams: the error comes back, the side streams are joined (they read the caller's planes), and the
    handle gives the frame's bytes once the plane is bound."""
    from stratusgfx_b200.gipass import StratusGiError
    want = run_cuda_staged(tiny, frames=1)[0]
    gp = make_pass(tiny)
    gp.set_gbuffer(tiny.gb[0], tiny.gb[-1])
    screen = gp.t.pop("screen"); gp.bind()
    with pytest.raises(StratusGiError, match="screen"):
        gp.frame(tiny.fp[0])
    gp.t["screen"] = screen; gp.reset_history(); gp.bind()
    gp.frame(tiny.fp[0])
    assert np.array_equal(gp.output("lightingOut").view(np.uint16), want["combined"].view(np.uint16))
    assert np.array_equal(gp.plane(_abi.PLANE_HISTORY).view(np.uint16), want["history"].view(np.uint16))
    gp.close()


def test_resize_then_run(built):
    d = SceneData("tiny", frames=1)
    gp = make_pass(d)
    gp._check(gp.L.sgi_resize(gp._h, d.W, d.H, 0, d.H))
    gp.set_gbuffer(d.gb[0], d.gb[-1]); gp.bind()
    gp.frame(d.fp[0])
    orc = run_oracle(d)[0]
    assert compare_radiance(gp.output("lightingOut"), orc["combined"])["ok"]
    gp.close()


# ---- BASELINE.json full size: size-independent properties + full oracle parity ----------------------------------
def test_c4_4k_2048vpl_properties_and_parity(built):
    d = SceneData("c4", frames=1)
    cu = run_cuda_staged(d)[0]
    V = cu["V"]
    idx = cu["visible"][1:V + 1]
    assert 0 < V <= d.N and np.all(np.diff(idx) > 0)                                   # ordered compaction
    assert np.array_equal(cu["entries"][:V], d.atlas_by_vpl[idx])
    assert np.array_equal(cu["updated"][:V, 0:4], d.vpl_data[idx, 0:4])
    # pass 0 == rgb / a, exactly
    res = cu["gi_reservoir"].astype(np.float32)
    want0 = (res[..., :3] / np.maximum(np.float32(1e-5), res[..., 3:4])).astype(np.float16)
    assert np.array_equal(cu["pass_reservoirs"][0][..., :3].view(np.uint16), want0.view(np.uint16))
    assert np.all(cu["history"] == np.float16(1.0))
    # tile lists: counts bounded, indices come from the visible list, sorted by ratio
    s2 = cu["stage2"]
    assert s2[:, 0].min() >= 0 and s2[:, 0].max() <= 12
    mask = np.arange(12)[None, :] < s2[:, 0:1]
    assert np.isin(s2[:, 1:][mask], idx).all()
    # the whole frame against the oracle
    orc = run_oracle(d)[0]
    rep, bad = compare_frame(cu, orc)
    assert not bad, bad
    # gather is deterministic: a second run gives the same bytes
    cu2 = run_cuda_staged(d)[0]
    assert np.array_equal(cu2["combined"].view(np.uint16), cu["combined"].view(np.uint16))


# ---- widening row f1: the VPL pre-pass (UpdatePointLights_ + InitVplFrameData_, BE:1359-1392, :1281-1292) ---------------
def _prepass_inputs(n, seed):
    from stratusgfx_b200.prepass import frustum_planes
    from stratusgfx_b200.scene import Scene
    from tests.test_oracle_selfchecks import _random_lights
    fp = Scene("tiny").frame_params(0)
    return _random_lights(n, seed), list(fp.viewPosition), frustum_planes(list(fp.invProjectionView))


@pytest.mark.parametrize("n,seed", [(0, 1), (1, 2), (257, 3), (2048, 5), (30000, 4)])
def test_vpl_prepass_is_bit_exact(tiny, n, seed):
    """Frustum decisions, distance order (ties keep input order) and the 10000 cap: the device table must equal the CPU
    restatement of the reference's multiset loop byte for byte."""
    from oracle import oracle_py as O
    rec, view, planes = _prepass_inputs(n, seed)
    vd, at, order = O.vpl_prepass(rec, view, planes)
    gp = make_pass(tiny)
    gp.set_lights(rec)
    gp.vpl_prepass(view, planes)
    gp.synchronize()
    count = int(gp.t["vplCountDevice"].cpu().numpy()[0])
    assert count == len(order)
    assert n < 30000 or count == 10000
    if count:
        assert np.array_equal(gp.t["prepassOrder"].cpu().numpy()[:count], order)
        assert np.array_equal(gp.t["vplData"].cpu().numpy()[:count].view(np.uint32), vd.view(np.uint32))
        assert np.array_equal(gp.t["atlasEntryByVpl"].cpu().numpy()[:count], at)
    gp.close()


def test_frame_from_device_built_vpl_table_equals_host_built(built):
    """The whole chain without a read-back: pre-pass -> cull (N clamped on the device by vplCountDevice) -> colour -> tile
    lists -> gather -> denoise must give the bytes of the same frame fed with the CPU-built table and its exact N."""
    from oracle import oracle_py as O
    from stratusgfx_b200.prepass import frustum_planes, light_records
    d = SceneData("c1", frames=1)
    fp = d.fp[0]
    rec = light_records(d.vpl_data, d.atlas_by_vpl)
    # shrink the radii so that the frustum test actually rejects some of the lights behind the camera plane
    rec["radius"] = 2.0
    planes = frustum_planes(list(fp.invProjectionView))
    view = list(fp.viewPosition)
    vd, at, order = O.vpl_prepass(rec, view, planes)
    # host-built reference run: exact N, the oracle's table (radius restored to what the gather expects is irrelevant: both runs share it)
    ref = make_pass(d)
    ref.set_vpls(vd, at); ref.set_gbuffer(d.gb[0], d.gb[-1]); ref.bind()
    fp_ref = fp.copy(); fp_ref.totalNumLights = len(order)
    ref.frame(fp_ref)
    want = ref.output("lightingOut").view(np.uint16).copy()
    want_vis = ref.output("vplVisibleIndices")[:len(order) + 1].copy()
    ref.close()
    gp = make_pass(d)
    upper = gp.set_lights(rec)
    gp.set_gbuffer(d.gb[0], d.gb[-1]); gp.bind()
    gp.vpl_prepass(view, planes)
    fp_dev = fp.copy(); fp_dev.totalNumLights = upper            # an upper bound: the cull reads the real N on the device
    gp.frame(fp_dev)
    got = gp.output("lightingOut").view(np.uint16)
    got_vis = gp.output("vplVisibleIndices")
    assert 0 < len(order) <= upper
    assert np.array_equal(got_vis[:want_vis[0] + 1], want_vis[:want_vis[0] + 1])
    assert np.array_equal(got, want)
    gp.close()


# ---- BASELINE.json configs[4]: one 7680x4320 view with 4096 VPLs (the largest size of the path) -------------------------
def test_c5_8k_4096vpl_band_parity_and_whole_view_properties(built):
    """The 8K view is too slow for the CPU oracle as a whole, so: (1) a 270-row band of it runs as a plain band handle and
    must match the oracle on the same band stage by stage (integer lists bit-exact, radiance within the bar); (2) the whole
    view runs once through sgi_frame and is checked through size-independent properties — ordered compaction, list
    bounds, history, determinism — and its gather output on the band's rows must be the band handle's bytes (the gather
    is per pixel, so a band and the whole view agree exactly)."""
    from oracle.oracle_py import OracleFrame
    from stratusgfx_b200.gipass import VplGiPass
    from stratusgfx_b200.scene import Scene
    s = Scene("c5")
    W, H, N = s.width, s.height, s.num_vpls
    assert (W, H, N) == (7680, 4320, 4096)
    vd, at = s.vpls(); sh, df = s.cubemaps(); csm = s.csm(0); fp = s.frame_params(0)
    y0, rows = 2700, 270                                   # a band of the lower half (floor, column bases: the expensive rows)
    gb, pgb = s.gbuffer(0, y0, rows), s.gbuffer(-1, y0, rows)
    # (1) band handle vs the oracle on the same band
    gp = VplGiPass(W, H, y0, rows, s.cube_res, s.csm_res, s.denoise_iters)
    gp.set_vpls(vd, at); gp.set_atlases(sh, df); gp.set_csm(csm); gp.set_gbuffer(gb, pgb); gp.bind()
    gp.cull(fp); V = gp.visible_count(); gp.color(fp); gp.tile_lists(fp); gp.gather(fp)
    band_res = gp.plane(_abi.PLANE_GI_RESERVOIR)
    gp.denoise(fp)
    cu = dict(V=V, visible=gp.output("vplVisibleIndices"), updated=gp.output("vplUpdatedData"), entries=gp.output("shadowDiffuseIndices"),
              stage1=gp.output("vplStage1Results"), stage2=gp.output("vplVisiblePerTile"), gi_color=gp.plane(_abi.PLANE_GI_COLOR),
              gi_reservoir=band_res, combined=gp.output("lightingOut"), gi_illum=gp.plane(_abi.PLANE_GI_ILLUM),
              history=gp.plane(_abi.PLANE_HISTORY), pass_reservoirs=[])
    gp.close()
    o = OracleFrame(W, rows, s.cube_res, s.csm_res, s.denoise_iters, y_offset=y0, full_height=H)
    o.frame(fp, gb, pgb, vd, at, csm, sh, df)
    orc = dict(visible=o.visible, updated=o.updated, entries=o.entries, stage1=o.stage1, stage2=o.stage2, gi_color=o.gi_color,
               gi_reservoir=o.gi_reservoir, pass_reservoirs=[], combined=o.combined, gi_illum=o.gi_illum, history=o.history)
    rep, bad = compare_frame(cu, orc)
    assert not bad, bad
    assert 0 < V <= N and rep["stage2_lists_exact"]
    del o, gb, pgb
    # (2) the whole view, once, through sgi_frame
    gbw, pgbw = s.gbuffer(0), s.gbuffer(-1)
    gp = VplGiPass(W, H, 0, H, s.cube_res, s.csm_res, s.denoise_iters)
    gp.set_vpls(vd, at); gp.set_atlases(sh, df); gp.set_csm(csm); gp.set_gbuffer(gbw, pgbw); gp.bind()
    gp.frame(fp)
    vis = gp.output("vplVisibleIndices")
    assert vis[0] == V and np.array_equal(vis[:V + 1], cu["visible"][:V + 1]) and np.all(np.diff(vis[1:V + 1]) > 0)
    s2 = gp.output("vplVisiblePerTile")
    assert s2[:, 0].min() >= 0 and s2[:, 0].max() <= 12
    tx = W // 5
    assert np.array_equal(s2[(y0 // 5) * tx:((y0 + rows) // 5) * tx], cu["stage2"])          # tile lists are per tile
    assert np.array_equal(gp.plane(_abi.PLANE_GI_RESERVOIR)[y0:y0 + rows].view(np.uint16), band_res.view(np.uint16))
    assert np.all(gp.plane(_abi.PLANE_HISTORY) == np.float16(1.0))
    first = gp.output("lightingOut").view(np.uint16).copy()
    assert np.isfinite(gp.output("lightingOut").astype(np.float32)).all()
    gp.reset_history()
    gp.set_gbuffer(gbw, pgbw); gp.bind()
    gp.frame(fp)
    assert np.array_equal(gp.output("lightingOut").view(np.uint16), first)                    # deterministic
    gp.close()


def test_gather_modes_can_alternate_on_one_handle(tiny):
    """Stochastic -> tile-list -> stochastic on the same handle: the persistent gather's two work counters alternate per
    persistent launch, so a tile-list launch in between must not leave the next stochastic launch without work."""
    gp = make_pass(tiny)
    gp.set_gbuffer(tiny.gb[0], tiny.gb[-1]); gp.bind()
    fp = tiny.fp[0].copy()
    outs = []
    for mode in (_abi.GATHER_STOCHASTIC, _abi.GATHER_TILE_LIST, _abi.GATHER_STOCHASTIC, _abi.GATHER_STOCHASTIC, _abi.GATHER_TILE_LIST, _abi.GATHER_STOCHASTIC):
        fp.gatherMode = mode
        gp.cull(fp); gp.color(fp); gp.tile_lists(fp); gp.gather(fp)
        outs.append((mode, gp.plane(_abi.PLANE_GI_RESERVOIR).view(np.uint16).copy()))
    sto = [o for m, o in outs if m == _abi.GATHER_STOCHASTIC]
    til = [o for m, o in outs if m == _abi.GATHER_TILE_LIST]
    assert sto[0].any() and all(np.array_equal(sto[0], o) for o in sto[1:])
    assert all(np.array_equal(til[0], o) for o in til[1:]) and not np.array_equal(til[0], sto[0])
    gp.close()


def test_cuda_integer_outputs_match_the_glm_golden_fixture(tiny):
    """Committed fixtures (tests/golden/tiny_shader_text_golden.json: the reference's shader text compiled and chained;
    tiny_glm_golden.json: the oracle built against the reference's glm; identical field by field): the CUDA path's visible list,
    atlas entries, per-tile lists and pre-pass order must equal them exactly, with no oracle in the loop."""
    import json
    import os
    from stratusgfx_b200.prepass import frustum_planes
    from tests.test_oracle_selfchecks import _random_lights
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_glm_golden.json")))
    # the same frames computed from the reference's shader text alone (tests/golden/make_shader_text_golden.py)
    text = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_shader_text_golden.json")))
    assert text["frames"] == gold["frames"]
    cu = run_cuda_staged(tiny)
    for f in range(2):
        g = gold["frames"][f]
        V = g["V"]
        assert cu[f]["V"] == V and cu[f]["visible"][:V + 1].tolist() == g["visible"]
        assert cu[f]["entries"][:V].reshape(-1).tolist() == g["entries"]
        cnt = cu[f]["stage2"][:, 0]
        assert cnt.tolist() == g["stage2_counts"]
        mask = np.arange(12)[None, :] < cnt[:, None]
        assert cu[f]["stage2"][:, 1:][mask].tolist() == g["stage2_indices"]
    p = gold["prepass"]
    rec = _random_lights(p["n"], p["seed"])
    gp = make_pass(tiny)
    gp.set_lights(rec)
    gp.vpl_prepass(list(tiny.fp[0].viewPosition), frustum_planes(list(tiny.fp[0].invProjectionView)))
    gp.synchronize()
    count = int(gp.t["vplCountDevice"].cpu().numpy()[0])
    assert gp.t["prepassOrder"].cpu().numpy()[:count].tolist() == p["order"]
    gp.close()


def test_vpl_prepass_frustum_boundary_cases(tiny):
    """Spheres that touch a frustum plane from outside (distance == -radius) are kept, a float ulp further out they are dropped:
    the cases of tests/test_reference_shim.py::test_reference_frustum_test_boundary_semantics, on the device."""
    from oracle import oracle_py as O
    from stratusgfx_b200.prepass import LIGHT_DTYPE
    planes = np.array([[1, 0, 0, 0], [-1, 0, 0, 100], [0, 1, 0, 0], [0, -1, 0, 100], [0, 0, 1, 0], [0, 0, -1, 100]], np.float32)
    rec = np.zeros(6, LIGHT_DTYPE)
    rec["position"] = [[-2, 50, 50], [-2.0000002, 50, 50], [50, 102, 50], [50, 50, -1.9999999], [50, 50, 50], [103, 50, 50]]
    rec["radius"] = 2.0; rec["farPlane"] = 2.0; rec["intensity"] = 1.0
    vd, at, order = O.vpl_prepass(rec, [50, 50, 50], planes)
    assert sorted(order.tolist()) == [0, 2, 3, 4]
    gp = make_pass(tiny)
    gp.set_lights(rec)
    gp.vpl_prepass([50, 50, 50], planes)
    gp.synchronize()
    count = int(gp.t["vplCountDevice"].cpu().numpy()[0])
    assert count == 4 and np.array_equal(gp.t["prepassOrder"].cpu().numpy()[:4], order)
    gp.close()
