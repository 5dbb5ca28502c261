"""Memory check and race check of the kernels and the C ABI without a GPU: the whole library built for the host
(tests/cuda_on_host/host_tu.py) with AddressSanitizer or ThreadSanitizer, driven by the GPU suite's own test bodies.

    python tools/sanitize_on_host.py address      # re-executes itself under LD_PRELOAD=libasan.so   (4 minutes on 8 cores; thread: 6, undefined: 4)
    python tools/sanitize_on_host.py thread       # ... libtsan.so
    python tools/sanitize_on_host.py undefined    # ... libubsan.so (misaligned vector accesses, signed overflow, bad shifts, float -> int overflow)
    python tools/sanitize_on_host.py address tma  # the short run: two staged frames, sgi_frame and the ragged frame sizes (what reaches k_atrous_tma.cu)

address: every "device" allocation is its own malloc block and the caller's buffers are ordinary heap tensors, so a kernel that
reads or writes past the end of a plane — which a GPU forgives inside an allocation granule — is reported with the source line.
thread: in lockstep mode every CUDA thread of a CTA is an OS thread and barriers / warp collectives are the only synchronisation,
so two threads of a CTA touching the same shared-memory or global word without one in between are reported (what compute-sanitizer's
racecheck looks for).  CTAs run one after the other, so races between CTAs are not visible to it.

Covers: two staged frames of a turning camera against the oracle, sgi_frame, the tile-list gather mode, 1 and 3 denoise passes, the
golden fixture, one / two / no visible VPL, more visible VPLs than the per-frame cap, resize, the pre-pass (0 .. 2048 lights and the
frustum boundary cases) and the widening kernels (direct lighting, SSAO, TAA)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "address"
    if kind not in ("address", "thread", "undefined"):
        print(__doc__); return 2
    if os.environ.get("SGI_ASAN_CHILD") != "1":
        from tests.cuda_on_host.host_tu import compile_whole_library
        lib = compile_whole_library(sanitize=kind)
        rt = subprocess.check_output(["gcc", "-print-file-name=" + {"address": "libasan.so", "thread": "libtsan.so", "undefined": "libubsan.so"}[kind]], text=True).strip()
        env = dict(os.environ, SGI_ASAN_CHILD="1", LD_PRELOAD=rt, SGI_HOST_LIB_OVERRIDE=lib, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1",
                   TSAN_OPTIONS="halt_on_error=1:report_signal_unsafe=0:history_size=2", UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")
        if kind == "thread":
            env["OMP_NUM_THREADS"] = "1"      # libgomp is not instrumented: its own barriers would look like races (oracle, scene generator)
        rc = subprocess.call([sys.executable, os.path.abspath(__file__), kind] + sys.argv[2:], env=env)
        print("exit code", rc, "(a sanitizer report ends the run with a non-zero code)")
        return rc
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
        if "tma" in sys.argv[2:]:
            # the short run: only what goes through the TMA-staged à-trous kernel (csrc/k_atrous_tma.cu over the shim's tensor maps:
            # a tap read past a tile's shared-memory box is a global-buffer-overflow on the kernel's static arrays) — the frames
            # above plus the ragged sizes (leftover columns and rows, a partial last tile row; an odd width without a tensor map)
            G.test_ragged_frame_sizes(None, 170, 95, False)
            G.test_ragged_frame_sizes(None, 165, 95, False)
            G.test_ragged_frame_sizes(None, 170, 95, True)       # tile-list gather: partial CTAs of 8 x 2 tiles at both edges
            print("ragged frame sizes: clean", flush=True)
            print(kind.upper(), "SANITIZER: NO REPORT")
            return 0
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
    print(kind.upper(), "SANITIZER: NO REPORT")
    return 0


if __name__ == "__main__":
    sys.exit(main())
