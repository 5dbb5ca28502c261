"""Memory check of the kernels and the C ABI without a GPU: the whole library built for the host (tests/cuda_on_host/host_tu.py) with
AddressSanitizer, driven by the GPU suite's own test bodies.  Every "device" allocation is its own malloc block and the caller's
buffers are ordinary heap tensors, so a kernel that reads or writes past the end of a plane — which a GPU forgives inside an
allocation granule — is reported with the source line.

    python tools/asan_on_host.py            # re-executes itself under LD_PRELOAD=libasan.so

Covers: two staged frames of a turning camera against the oracle, sgi_frame, the tile-list gather mode, 1-3 denoise passes, the
golden fixture, one / two / no visible VPL, more visible VPLs than the per-frame cap, resize, the pre-pass (0 .. 2048 lights and the
frustum boundary cases) and the three kernels that have not met a GPU yet (direct lighting, SSAO, TAA)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    if os.environ.get("SGI_ASAN_CHILD") != "1":
        from tests.cuda_on_host.host_tu import compile_whole_library
        lib = compile_whole_library(sanitize=True)
        asan = subprocess.check_output(["gcc", "-print-file-name=libasan.so"], text=True).strip()
        env = dict(os.environ, SGI_ASAN_CHILD="1", LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1", SGI_HOST_LIB_OVERRIDE=lib)
        return subprocess.call([sys.executable, os.path.abspath(__file__)], env=env)
    from tests import test_parity_gpu as G, test_zz_direct_lighting as D, test_zz_ssao as S, test_zz_taa as T
    from tests.host_gipass import gpu_tests_on_the_host
    from tests.parity_util import SceneData, compare_frame, make_pass, run_cuda_staged, run_oracle
    from stratusgfx_b200 import _abi
    tiny = SceneData("tiny", frames=2)
    turning = SceneData("tiny", frames=2, yaw_deg_per_frame=1.5)
    orc = run_oracle(turning)
    with gpu_tests_on_the_host():
        cu = run_cuda_staged(turning)
        for f in range(2):
            _, bad = compare_frame(cu[f], orc[f])
            assert not bad, bad
        gp = make_pass(turning)
        for f in range(2):
            gp.set_gbuffer(turning.gb[f], turning.gb[f - 1]); gp.set_csm(turning.csm[f]); gp.bind(); gp.frame(turning.fp[f])
        gp.close()
        print("staged frames, sgi_frame: clean", flush=True)
        for name, args in (("test_tile_list_gather_mode", (None,)), ("test_denoise_iteration_count_is_a_parameter", (tiny, 1)),
                           ("test_denoise_iteration_count_is_a_parameter", (tiny, 3)), ("test_cuda_integer_outputs_match_the_glm_golden_fixture", (tiny,)),
                           ("test_single_vpl_and_two_vpls", (None,)), ("test_no_visible_vpl", (None,)),
                           ("test_more_visible_vpls_than_the_per_frame_cap", (None,)), ("test_resize_then_run", (None,)),
                           ("test_vpl_prepass_frustum_boundary_cases", (tiny,))):
            getattr(G, name)(*args)
            print(name, args[1:] if len(args) > 1 else "", "clean", flush=True)
        for n, seed in ((0, 1), (1, 2), (257, 3), (2048, 5)):
            G.test_vpl_prepass_is_bit_exact(tiny, n, seed)
        D.test_direct_lighting_against_the_compiled_shader_text(None, "tiny")
        S.test_ssao_against_the_compiled_shader_text(None, "tiny")
        T.test_taa_against_the_compiled_shader_text(None, "tiny", 1.5)
        print("pre-pass, direct lighting, SSAO, TAA: clean", flush=True)
    print("ADDRESS SANITIZER: NO REPORT")
    return 0


if __name__ == "__main__":
    sys.exit(main())
