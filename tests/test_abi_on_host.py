"""The C ABI itself without a GPU: every translation unit of libstratusgi.so compiled for the host (tests/cuda_on_host/host_tu.py
compile_whole_library: lockstep kernels, CUDA-runtime stand-ins) behind the product's own Python host class (tests/host_gipass.py).
What this adds to the per-kernel host tests is the host-side state machine of csrc/sgi_abi.cu — bind resets, the pass 0 the gather
fuses, the guide plane, history ping-pong, sgi_frame's stage order — and a dry run of the GPU tests' own code, including the opt-in
tests of kernels that have not met a GPU yet (direct lighting, SSAO, TAA)."""
import os

import numpy as np
import pytest

from tests.host_gipass import gpu_tests_on_the_host


@pytest.fixture(scope="module", autouse=True)
def _need_cuda_headers(built):
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")


def test_staged_calls_and_sgi_frame_through_the_host_built_abi(tiny, tiny_oracle):
    """Two frames stage by stage against the oracle (the GPU suite's central test), then the same two frames through sgi_frame on a
    second handle: same bytes."""
    from tests.parity_util import compare_frame, make_pass, run_cuda_staged
    with gpu_tests_on_the_host():
        cu = run_cuda_staged(tiny)
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


def test_opt_in_gpu_tests_of_the_unvalidated_kernels_run_on_the_host(built):
    """tests/test_zz_direct_lighting.py, test_zz_ssao.py, test_zz_taa.py: the same code, the same bars."""
    from oracle import oracle_py
    if oracle_py.reference_glsl() is None:
        pytest.skip("oracle/_ref/libref_glsl.so not built (needs the reference tree)")
    from tests import test_zz_direct_lighting as D, test_zz_ssao as S, test_zz_taa as T
    with gpu_tests_on_the_host():
        D.test_direct_lighting_against_the_compiled_shader_text(built, "tiny")
        S.test_ssao_against_the_compiled_shader_text(built, "tiny")
        T.test_taa_against_the_compiled_shader_text(built, "tiny", 1.5)


# ---- the bodies of GPU-suite tests (tests/test_parity_gpu.py), unchanged, over the host-built library --------------------------
_FAST = ["test_fused_frame_call_equals_staged_calls", "test_reference_method_names_drive_the_same_path", "test_zero_lights_is_a_no_op",
         "test_unbound_buffers_and_bad_passes_are_errors", "test_gather_modes_can_alternate_on_one_handle",
         "test_cuda_integer_outputs_match_the_glm_golden_fixture"]
_FAST_BUILT = ["test_single_vpl_and_two_vpls", "test_no_visible_vpl", "test_resize_then_run"]
_SLOW_BUILT = ["test_more_visible_vpls_than_the_per_frame_cap", "test_tile_list_gather_mode", "test_row_bands_with_apron_reproduce_the_whole_frame"]


@pytest.mark.parametrize("name", _FAST)
def test_gpu_suite_body_on_the_host(tiny, name):
    from tests import test_parity_gpu as G
    with gpu_tests_on_the_host():
        getattr(G, name)(tiny)


@pytest.mark.parametrize("name", _FAST_BUILT + _SLOW_BUILT)
def test_gpu_suite_edge_cases_on_the_host(built, name):
    """Single VPL, no visible VPL, resize on every run; the per-frame cap, the tile-list gather mode and apron bands (3 minutes
    together) with SGI_HOST_SUITE=full."""
    if name in _SLOW_BUILT and os.environ.get("SGI_HOST_SUITE", "") != "full":
        pytest.skip("slow on the host: set SGI_HOST_SUITE=full")
    from tests import test_parity_gpu as G
    with gpu_tests_on_the_host():
        getattr(G, name)(built)


@pytest.mark.parametrize("iters", [1, 3])
def test_denoise_iteration_count_on_the_host(tiny, iters):
    from tests import test_parity_gpu as G
    with gpu_tests_on_the_host():
        G.test_denoise_iteration_count_is_a_parameter(tiny, iters)
