"""Race check of the peer-to-peer band protocol without a GPU.

The bands of tests/test_bands_lockstep_on_host.py, but as THREADS of one process under ThreadSanitizer: every band drives its own copy
of the band library (tests/cuda_on_host/bands_on_host.cpp; one DSO per band, so each has its own lockstep runtime and shared-memory
statics), the halo slabs are ordinary arrays, and the epoch flags are what they are on the device — release stores and acquire loads
(cuda_host_shim.h: st.release.sys / ld.acquire.sys as __atomic builtins; atomicAdd as a sequentially consistent read-modify-write).
A neighbour's kernel storing a halo row and this band's kernel reading it are then ordered ONLY by the protocol of csrc/sgi_band.cu
and csrc/sgi_internal.h (fence, producer count, last producer releases the flag; consumer acquires it, then a barrier), and
ThreadSanitizer reports any pair of accesses the protocol does not order — a missing wait, a flag released before the last producer's
stores, a halo overwritten while the neighbour may still read it.

    python tools/race_check_bands_on_host.py [bands=4] [kernel|host] [exchange|recompute] [nowait]      # re-executes itself under LD_PRELOAD=libtsan.so
    (SGI_TSAN_EXTRA=":verbosity=2" is appended to the child's TSAN_OPTIONS)

`recompute` = SGI_HALO_RECOMPUTE: the bands compute the non-final passes on their halo rows themselves; what is left to order is the
final pass's GI rows (stored into the neighbours, read by their temporal pre-pass of the next frame, overwritten two frames later),
which is why the check runs three frames, all of them inside ONE library call per band (the interpreter lock would otherwise order
everything across frames).  Limit, observed: with the pre-pass's wait for the neighbours' previous frame REMOVED (`nowait` as a
fourth argument), ThreadSanitizer still reports nothing in this mode.  The cause is not established.  It is not the number of emulation
threads between the stores and the reads, a whole frame apart: the lockstep runtime used to create one OS thread per CUDA thread of
every CTA (hundreds of thousands per frame) and now creates them once per launch (a few thousand), and the outcome is the same.
So this mode shows that nothing ELSE is shared between the bands, not that the inter-frame wait is needed.
The intra-pass waits of the exchange mode are checked by this tool (a removed wait is reported); the inter-frame GI rows rest on
the ordering argument of DESIGN.md section 6 and on the byte comparisons of tests/test_bands_lockstep_on_host.py and the GPU suite.
"""
import ctypes as C
import os
import shutil
import subprocess
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
W, H, FRAMES = 160, 240, 3


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    kernel_signal = (sys.argv[2] if len(sys.argv) > 2 else "kernel") == "kernel"
    recompute = (sys.argv[3] if len(sys.argv) > 3 else "exchange") == "recompute"
    drop_inter_frame_wait = "nowait" in sys.argv[4:]      # sensitivity check: the temporal pre-pass skips its wait for the neighbours' previous frame
    from tests.cuda_on_host.host_tu import BUILD, compile_wrapper
    if os.environ.get("SGI_TSAN_CHILD") != "1":
        so = compile_wrapper(("vpl_on_host.cpp", "gather_on_host.cpp", "denoise_lockstep_on_host.cpp", "bands_on_host.cpp"),
                             ("k_vpl.cu", "k_gather.cu", "k_denoise.cu", "sgi_band.cu"), "libbands_on_host_tsan.so",
                             extra=("-O1", "-g", "-fno-omit-frame-pointer", "-fsanitize=thread", "-fno-gnu-unique"))
        rt = subprocess.check_output(["gcc", "-print-file-name=libtsan.so"], text=True).strip()
        env = dict(os.environ, SGI_TSAN_CHILD="1", LD_PRELOAD=rt, OMP_NUM_THREADS="1", SGI_BANDS_TSAN_LIB=so,
                   TSAN_OPTIONS="halt_on_error=1:report_signal_unsafe=0:history_size=4" + os.environ.get("SGI_TSAN_EXTRA", ""))
        rc = subprocess.call([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env)
        print("exit code", rc, "(a ThreadSanitizer report ends the run with a non-zero code)")
        return rc

    import numpy as np
    from oracle.oracle_py import _atlas_ptrs
    from stratusgfx_b200._abi import FrameParams
    from stratusgfx_b200.bands import bands_from_cuts, halo_rows, recompute_rows
    from stratusgfx_b200.scene import Scene
    from tests.parity_util import SceneData, run_oracle
    from tests.test_bands_lockstep_on_host import HostPeer

    d = SceneData("tiny", frames=FRAMES, yaw_deg_per_frame=1.5, height=H)
    orc = run_oracle(d, keep_passes=False)
    tables = [tuple(np.ascontiguousarray(o[k]) for k in ("updated", "visible", "entries")) for o in orc]
    cuts = {3: [0, 80, 160, H], 4: [0, 60, 120, 180, H], 5: [0, 40, 95, 140, 200, H], 8: list(range(0, H + 1, 30))}[world]
    bands = bands_from_cuts(cuts, H, (recompute_rows if recompute else halo_rows)(4))
    rows_max = max(b.rows for b in bands)
    libs = []
    for r in range(world):                                   # one DSO per band: its own runtime state and __shared__ statics
        path = os.path.join(BUILD, f"libbands_on_host_tsan_band{r}.so")
        shutil.copyfile(os.environ["SGI_BANDS_TSAN_LIB"], path)
        L = C.CDLL(path)
        L.host_band_slab_bytes.restype = C.c_int64
        L.host_band_init.argtypes = [C.c_int] * 11 + [C.c_void_p, C.c_int, C.POINTER(HostPeer), C.POINTER(HostPeer), C.c_int]
        L.host_band_frame.argtypes = [C.POINTER(FrameParams)] + [C.c_void_p] * 16 + [C.c_int, C.c_void_p]
        L.host_band_frames.argtypes = [C.c_int, C.POINTER(FrameParams), C.c_void_p, C.c_int, C.c_void_p]
        if drop_inter_frame_wait:
            L.host_band_test_hooks(0, 1)
        libs.append(L)
    slab_bytes = int(libs[0].host_band_slab_bytes(W, rows_max))
    slabs = [np.zeros(slab_bytes, np.uint8) for _ in range(world)]
    scene = d.scene
    sh, _ = scene.cubemaps()
    shp = _atlas_ptrs(sh)
    lit = np.zeros((FRAMES, H, W, 4), np.float16)
    inputs = []
    for r, b in enumerate(bands):                            # everything a band reads is prepared before the threads start
        per_frame = []
        for f in range(FRAMES):
            gb, pgb = scene.gbuffer(f, b.y0, b.rows), scene.gbuffer(f - 1, b.y0, b.rows)
            per_frame.append([np.ascontiguousarray(a) for a in (gb.depth, gb.normal, gb.albedo, gb.baseReflectivity, gb.roughnessMetallicAmbient,
                                                                gb.ssao, gb.velocity, gb.id, gb.screen, pgb.normal, pgb.id, pgb.depth)])
        inputs.append(per_frame)
    fps = []
    for f in range(FRAMES):
        fp = scene.frame_params(f); fp.totalNumLights = int(len(tables[f][1]) - 1); fps.append(fp)

    def peer(r):
        if r < 0 or r >= world:
            return HostPeer(0, 0, 0, 0, 0, 0, None)
        b = bands[r]
        return HostPeer(1, b.y0, b.rows, b.b0, b.b1 - b.b0, r, slabs[r].ctypes.data)

    errors = []

    def run_band(r):
        try:
            b, L = bands[r], libs[r]
            up, down = peer(r - 1), peer(r + 1)
            L.host_band_init(W, H, scene.cube_res, scene.csm_res, scene.denoise_iters, r, world, b.y0, b.rows, b.b0, b.b1 - b.b0,
                             slabs[r].ctypes.data, rows_max, C.byref(up), C.byref(down), (1 if kernel_signal else 0) | (2 if recompute else 0))
            # all frames in ONE library call: the interpreter lock is not touched between two frames, so the only thing that orders
            # a neighbour's stores of frame f against this band's reads of frame f + 1 (and its stores of frame f + 2) is the protocol
            combined = [np.zeros((b.rows, W, 4), np.float16) for _ in range(FRAMES)]
            ptrs = (C.c_void_p * (16 * FRAMES))()
            for f in range(FRAMES):
                updated, visible, entries = tables[f]
                vals = [a.ctypes.data for a in inputs[r][f]] + [updated.ctypes.data, visible.ctypes.data, entries.ctypes.data, C.addressof(shp)]
                for k, v in enumerate(vals):
                    ptrs[16 * f + k] = v
            outs = (C.c_void_p * FRAMES)(*[c.ctypes.data for c in combined])
            fp_arr = (FrameParams * FRAMES)(*fps)
            rc = L.host_band_frames(FRAMES, fp_arr, ptrs, sh.shape[0], outs)
            assert rc == 0
            for f in range(FRAMES):
                lit[f, b.b0:b.b1] = combined[f][b.owned]
        except Exception as e:          # noqa: BLE001
            errors.append((r, repr(e)))

    threads = [threading.Thread(target=run_band, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    # and the bytes are the one-handle frame's (the oracle within the bar is enough here; tests/test_bands_lockstep_on_host.py holds the bytes)
    from tests.parity_util import compare_radiance
    for f in range(FRAMES):
        rep = compare_radiance(lit[f], orc[f]["combined"], 3)
        assert rep["ok"] and rep["bit_exact"] > 0.99, (f, rep)
    print(f"{world} bands as threads, {'recomputed halos' if recompute else 'exchanged halos'}, {'in-kernel' if kernel_signal else 'host-launched'} signals, {FRAMES} frames: THREAD SANITIZER: NO REPORT")
    return 0


if __name__ == "__main__":
    sys.exit(main())
