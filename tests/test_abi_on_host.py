"""The C ABI itself without a GPU: every translation unit of libstratusgi.so — the TMA-staged à-trous kernel included, over a stand-in
for tensor maps and cp.async.bulk.tensor — compiled for the host (tests/cuda_on_host/host_tu.py compile_whole_library: lockstep kernels,
CUDA-runtime stand-ins) behind the product's own Python host class (tests/host_gipass.py).
What this adds to the per-kernel host tests is the host-side state machine of csrc/sgi_abi.cu — bind resets, the pass 0 the gather
fuses, the guide plane, history ping-pong, sgi_frame's stage order — and a dry run of the GPU tests' own code, including the
tests of the widening kernels (direct lighting, SSAO, TAA)."""
import os

import numpy as np
import pytest

from tests.host_gipass import gpu_tests_on_the_host


@pytest.fixture(scope="module", autouse=True)
def _need_cuda_headers(built):
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")


def _tma_counters():
    """(boxes copied by the stand-in for cp.async.bulk.tensor, tensor-map encodes refused) of the host-built library so far"""
    import ctypes
    from tests.host_gipass import host_library
    L = host_library()
    L.sgihost_tma_loads.restype = L.sgihost_tma_rejected.restype = ctypes.c_longlong
    return int(L.sgihost_tma_loads()), int(L.sgihost_tma_rejected())


def test_staged_calls_and_sgi_frame_through_the_host_built_abi(tiny, tiny_oracle):
    """Two frames stage by stage against the oracle (the GPU suite's central test), then the same two frames through sgi_frame on a
    second handle: same bytes.  The à-trous passes with spacing 2 and 4 go through csrc/k_atrous_tma.cu as on a GPU — lattice tiles
    from the emulated tensor maps next to the border CTAs in one launch (160 x 90 has too few lattice rows for spacing 8)."""
    from tests.parity_util import compare_frame, make_pass, run_cuda_staged
    boxes0, _ = _tma_counters()
    with gpu_tests_on_the_host():
        cu = run_cuda_staged(tiny)
        assert _tma_counters()[0] - boxes0 == 2 * 2 * (18 + 24), "the TMA-staged à-trous kernel did not run (2 frames x 2 planes x the tiles of spacing 2 and 4)"
        for f in range(2):
            rep, bad = compare_frame(cu[f], tiny_oracle[f])
            assert not bad, (f, bad)
        gp = make_pass(tiny)
        for f in range(2):
            gp.set_gbuffer(tiny.gb[f], tiny.gb[f - 1]); gp.set_csm(tiny.csm[f]); gp.bind()
            gp.frame(tiny.fp[f])
            assert np.array_equal(gp.output("lightingOut").view(np.uint16), cu[f]["combined"].view(np.uint16)), f
            assert np.array_equal(gp.output("vplVisiblePerTile")[:, 0], cu[f]["stage2"][:, 0])
        gp.close()


def test_prepass_boundary_test_of_the_gpu_suite_runs_on_the_host(tiny):
    from tests import test_parity_gpu as G
    with gpu_tests_on_the_host():
        G.test_vpl_prepass_frustum_boundary_cases(tiny)
        G.test_vpl_prepass_is_bit_exact(tiny, 257, 3)


def test_gpu_tests_of_the_widening_kernels_run_on_the_host(built):
    """tests/test_zz_direct_lighting.py, test_zz_ssao.py, test_zz_taa.py: the same code, the same bars."""
    from oracle import oracle_py
    if oracle_py.reference_glsl() is None:
        pytest.skip("oracle/_ref/libref_glsl.so not built (needs the reference tree)")
    from tests import test_zz_direct_lighting as D, test_zz_ssao as S, test_zz_taa as T
    with gpu_tests_on_the_host():
        D.test_direct_lighting_against_the_compiled_shader_text(built, "tiny")
        S.test_ssao_against_the_compiled_shader_text(built, "tiny")
        T.test_taa_against_the_compiled_shader_text(built, "tiny", 1.5)


def test_cube_map_production_gpu_test_body_runs_on_the_host(built):
    """tests/test_zz_cubemaps.py over the host-built library: csrc/k_cubemaps.cu in lockstep mode against the oracle, bit for bit."""
    from tests import test_zz_cubemaps as Cm
    with gpu_tests_on_the_host():
        Cm.check_cubemaps("tiny", "all")
        if os.environ.get("SGI_HOST_SUITE", "") == "full":      # (a whole frame downstream of the rendered maps: 25 s on the host)
            Cm.test_rendered_maps_feed_the_frame(built)


# ---- the bodies of GPU-suite tests (tests/test_parity_gpu.py), unchanged, over the host-built library --------------------------
_FAST = ["test_reference_method_names_drive_the_same_path", "test_zero_lights_is_a_no_op", "test_unbound_buffers_and_bad_passes_are_errors"]
# (45-50 s each on the host since the gather's per-pixel work grew tables and prefetches: sgi_frame against staged calls is also
# covered by test_staged_calls_and_sgi_frame_through_the_host_built_abi, the early cull by the GPU suite)
_SLOW = ["test_fused_frame_call_equals_staged_calls", "test_early_cull_gives_the_frames_of_the_serial_order",
         "test_gather_modes_can_alternate_on_one_handle", "test_cuda_integer_outputs_match_the_glm_golden_fixture",
         "test_frame_that_fails_in_its_last_pass_leaves_a_usable_handle"]
_FAST_BUILT = ["test_no_visible_vpl", "test_resize_then_run"]
_SLOW_BUILT = ["test_single_vpl_and_two_vpls", "test_more_visible_vpls_than_the_per_frame_cap", "test_tile_list_gather_mode",
               "test_row_bands_with_apron_reproduce_the_whole_frame"]


@pytest.mark.parametrize("name", _FAST + _SLOW)
def test_gpu_suite_body_on_the_host(tiny, name):
    if name in _SLOW and os.environ.get("SGI_HOST_SUITE", "") != "full":
        pytest.skip("10-50 s on the host: set SGI_HOST_SUITE=full")
    from tests import test_parity_gpu as G
    with gpu_tests_on_the_host():
        getattr(G, name)(tiny)


@pytest.mark.parametrize("name", _FAST_BUILT + _SLOW_BUILT)
def test_gpu_suite_edge_cases_on_the_host(built, name):
    """No visible VPL and resize on every run; one / two VPLs, the per-frame cap, the tile-list gather mode and apron bands (3 minutes
    together) with SGI_HOST_SUITE=full (tools/sanitize_on_host.py runs them all)."""
    if name in _SLOW_BUILT and os.environ.get("SGI_HOST_SUITE", "") != "full":
        pytest.skip("slow on the host: set SGI_HOST_SUITE=full")
    from tests import test_parity_gpu as G
    with gpu_tests_on_the_host():
        getattr(G, name)(built)


def test_tile_list_gather_with_full_lists_on_the_host(built):
    """tests/test_parity_gpu.py::test_tile_list_gather_mode_with_full_lists, ten rows of its c3 band (every tile holds 12 VPLs)."""
    if os.environ.get("SGI_HOST_SUITE", "") != "full":
        pytest.skip("27 s on the host: set SGI_HOST_SUITE=full")
    from tests import test_parity_gpu as G
    with gpu_tests_on_the_host():
        rep, counts = G.check_tile_list_band("c3", 700, 10)
    assert counts.min() == 12


@pytest.mark.parametrize("width,tma,tile_list_mode", [(170, True, False), (165, False, False), (170, True, True)])
def test_ragged_frame_size_on_the_host(built, width, tma, tile_list_mode):
    """tests/test_parity_gpu.py::test_ragged_frame_sizes at 170 x 95 (no size is a multiple of a warp segment, a CTA or a spacing:
    lattice tiles with leftover columns and rows next to them) and 165 x 95 (an odd width: the reservoir plane's row pitch is no
    multiple of 16 bytes, the tensor-map encode is refused as by the driver and every pass takes the clamped kernel)."""
    if os.environ.get("SGI_HOST_SUITE", "") != "full":
        pytest.skip("15 s each on the host: set SGI_HOST_SUITE=full")
    from tests import test_parity_gpu as G
    boxes0, refused0 = _tma_counters()
    with gpu_tests_on_the_host():
        G.test_ragged_frame_sizes(built, width, 95, tile_list_mode)
    boxes, refused = _tma_counters()
    assert (boxes > boxes0) == tma and (refused > refused0) == (not tma)


@pytest.mark.parametrize("iters", [1, 3])
def test_denoise_iteration_count_on_the_host(tiny, iters):
    from tests import test_parity_gpu as G
    with gpu_tests_on_the_host():
        G.test_denoise_iteration_count_is_a_parameter(tiny, iters)


@pytest.mark.parametrize("halo_mode", [0, 1], ids=["exchange", "recompute"])
def test_p2p_bands_through_the_abi_with_a_middle_band(built, halo_mode):
    """sgi_band_setup / export / attach (same-process peers), the band branches of sgi_gather and sgi_denoise_pass, the image
    transfer into band 0 and sgi_frame_image, on 3 bands over two frames of a turning camera.  A host launch is synchronous, so
    the bands advance stage by stage (every in-kernel wait is already satisfied when it is reached); on a GPU the same calls run
    concurrently on one stream per band (tests/test_parity_gpu.py::test_p2p_row_bands_reproduce_the_whole_frame)."""
    from stratusgfx_b200._abi import PEER_DOWN, PEER_ROOT, PEER_UP
    from stratusgfx_b200.bands import bands_from_cuts, halo_rows, recompute_rows
    from tests.parity_util import SceneData, band_of, make_pass, run_cuda_staged
    d = SceneData("tiny", frames=2, yaw_deg_per_frame=1.5, height=180)
    iters = d.scene.denoise_iters
    if halo_mode == 0 and os.environ.get("SGI_HOST_SUITE", "") != "full":
        pytest.skip("50 s on the host; the exchange protocol itself runs in tests/test_bands_lockstep_on_host.py on every run: set SGI_HOST_SUITE=full")
    with gpu_tests_on_the_host():
        whole = run_cuda_staged(d)
        # halo_mode 1 = SGI_HALO_RECOMPUTE: the bands compute the non-final passes on their halo rows themselves (28 / 24 / 16 rows)
        bands = bands_from_cuts([0, 60, 120, 180], d.H, (recompute_rows if halo_mode else halo_rows)(iters))
        gps = [make_pass(d, b.y0, b.rows) for b in bands]
        for gp, b in zip(gps, bands):
            gp.band_setup(b.b0, b.b1 - b.b0, b.rank, 3, owns_frame_image=(b.rank == 0), halo_mode=halo_mode)
        infos = [gp.band_export() for gp in gps]
        for gp, b in zip(gps, bands):
            if b.rank > 0: gp.band_attach(PEER_UP, infos[b.rank - 1])
            if b.rank + 1 < 3: gp.band_attach(PEER_DOWN, infos[b.rank + 1])
            if b.rank != 0: gp.band_attach(PEER_ROOT, infos[0])
        for f in range(2):
            fp = d.fp[f]
            order = gps if f % 2 == 0 else gps[::-1]
            for gp, b in zip(gps, bands):
                gp.set_gbuffer(band_of(d.gb[f], b.y0, b.rows), band_of(d.gb[f - 1], b.y0, b.rows)); gp.set_csm(d.csm[f]); gp.bind()
            for gp in order:
                gp.cull(fp); gp.color(fp)
            for gp in order:
                gp.gather(fp)
            for p in range(iters):
                for gp in order:
                    gp.denoise_pass(fp, p)
            image = gps[0].frame_image().cpu().numpy()
            assert np.array_equal(image.view(np.uint16), whole[f]["combined"].view(np.uint16)), f
        assert [gp.band_timeouts() for gp in gps] == [0, 0, 0]
        for gp in gps:
            gp.close()
