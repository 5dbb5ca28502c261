"""Row bands with peer-to-peer halos and IN-KERNEL waits / signals, with MIDDLE bands, without a GPU.

Every band is a process (as on the GPU box); the slabs neighbours store into are POSIX shared memory instead of CUDA IPC mappings;
the band's gather and denoise kernels are the shipped ones, compiled for the host in lockstep mode (tests/cuda_on_host/), and the
links / epochs / signals come from the translated csrc/sgi_band.cu.  Two frames (the second one starts with the wait for the
neighbours' frame-complete flag and reads the history rows they pushed) on 3 and 4 bands must give, byte for byte, what ONE handle
computes for the whole frame — a pass that read a halo row before its neighbour stored it, or a producer count that is off by one
(dead-lock: the test times out), cannot pass.  The 2-GPU runs on hardware cover band 0 and band N-1 only; this covers the bands
in between, with the in-kernel signal and with the host-launched signal kernel."""
import ctypes as C
import multiprocessing as mp
import os

import numpy as np
import pytest

W, H, FRAMES = 160, 240, 3      # (three frames: the GI ping-pong plane a final pass stores into is the one the neighbour's temporal pre-pass read two frames earlier)
WORKLOAD = dict(workload="tiny", frames=FRAMES, yaw_deg_per_frame=1.5, height=H)


class HostPeer(C.Structure):
    _fields_ = [("present", C.c_int32), ("y_offset", C.c_int32), ("rows", C.c_int32), ("compute_y0", C.c_int32), ("compute_rows", C.c_int32),
                ("rank", C.c_int32), ("base", C.c_void_p)]


def _load():
    from stratusgfx_b200._abi import FrameParams
    from tests.cuda_on_host.host_tu import compile_wrapper
    defines = ()
    L = C.CDLL(compile_wrapper(("vpl_on_host.cpp", "gather_on_host.cpp", "denoise_lockstep_on_host.cpp", "bands_on_host.cpp"),
                               ("k_vpl.cu", "k_gather.cu", "k_denoise.cu", "sgi_band.cu"),
                               "libbands_on_host" + "".join("_" + x[-2:].lower() for x in defines) + ".so", defines))
    L.host_band_slab_bytes.restype = C.c_int64
    L.host_band_init.argtypes = [C.c_int] * 11 + [C.c_void_p, C.c_int, C.POINTER(HostPeer), C.POINTER(HostPeer), C.c_int]
    L.host_band_frame.argtypes = [C.POINTER(FrameParams)] + [C.c_void_p] * 16 + [C.c_int, C.c_void_p]
    return L


def _band_worker(rank, bands, shm_names, slab_bytes, rows_max, tables, kernel_signal, out_name, hooks=None):
    """One band process: attach every slab, run FRAMES frames, leave the lit owned rows in the output block."""
    from multiprocessing import shared_memory
    from oracle.oracle_py import _atlas_ptrs
    from stratusgfx_b200.scene import Scene
    L = _load()
    world = len(bands)
    shms = [shared_memory.SharedMemory(name=n) for n in shm_names]
    out = shared_memory.SharedMemory(name=out_name)
    base = [C.addressof(C.c_char.from_buffer(s.buf)) for s in shms]
    me = bands[rank]
    def peer(r):
        if r < 0 or r >= world:
            return HostPeer(0, 0, 0, 0, 0, 0, None)
        b = bands[r]
        return HostPeer(1, b.y0, b.rows, b.b0, b.b1 - b.b0, r, base[r])
    up, down = peer(rank - 1), peer(rank + 1)
    scene = Scene("tiny", height=H, yaw_deg_per_frame=1.5)
    L.host_band_init(W, H, scene.cube_res, scene.csm_res, scene.denoise_iters, rank, world, me.y0, me.rows, me.b0, me.b1 - me.b0,
                     base[rank], rows_max, C.byref(up), C.byref(down), int(kernel_signal))          # bit 0: in-kernel signal, bit 1: SGI_HALO_RECOMPUTE
    if hooks is not None:            # (delay of this rank's final passes in seconds, drop the inter-frame wait) per rank
        L.host_band_test_hooks(int(hooks[rank][0]), int(hooks[rank][1]))
    sh, _ = scene.cubemaps()
    shp = _atlas_ptrs(sh)
    lit = np.ndarray((FRAMES, H, W, 4), np.float16, buffer=out.buf)
    for f in range(FRAMES):
        fp = scene.frame_params(f)
        gb, pgb = scene.gbuffer(f, me.y0, me.rows), scene.gbuffer(f - 1, me.y0, me.rows)
        updated, visible, entries = (np.ascontiguousarray(a) for a in tables[f])
        fp.totalNumLights = int(len(visible) - 1)
        combined = np.zeros((me.rows, W, 4), np.float16)
        keep = [np.ascontiguousarray(a) for a in (gb.depth, gb.normal, gb.albedo, gb.baseReflectivity, gb.roughnessMetallicAmbient, gb.ssao,
                                                  gb.velocity, gb.id, gb.screen, pgb.normal, pgb.id, pgb.depth)]
        rc = L.host_band_frame(C.byref(fp), *[a.ctypes.data for a in keep], updated.ctypes.data, visible.ctypes.data, entries.ctypes.data,
                               C.addressof(shp), sh.shape[0], combined.ctypes.data)
        assert rc == 0
        lit[f, me.b0:me.b1] = combined[me.owned]
    del lit
    for s in shms + [out]:
        s.close()


def _run(world, tables, kernel_signal, timeout=240, recompute=False, hooks=None):
    from multiprocessing import shared_memory
    from stratusgfx_b200.bands import bands_from_cuts, halo_rows, recompute_rows
    L = _load()
    cuts = {1: [0, H], 3: [0, 80, 160, H], 4: [0, 60, 120, 180, H], 5: [0, 40, 95, 140, 200, H], 8: list(range(0, H + 1, 30))}[world]      # (5: unequal bands, the balanced plan)
    bands = bands_from_cuts(cuts, H, (recompute_rows if recompute else halo_rows)(4))
    rows_max = max(b.rows for b in bands)
    slab_bytes = int(L.host_band_slab_bytes(W, rows_max))
    shms = [shared_memory.SharedMemory(create=True, size=slab_bytes) for _ in range(world)]
    out = shared_memory.SharedMemory(create=True, size=FRAMES * H * W * 8)
    try:
        for s in shms:
            np.ndarray(slab_bytes, np.uint8, buffer=s.buf)[:] = 0
        ctx = mp.get_context("spawn")
        procs = [ctx.Process(target=_band_worker, args=(r, bands, [s.name for s in shms], slab_bytes, rows_max, tables, (1 if kernel_signal else 0) | (2 if recompute else 0), out.name, hooks))
                 for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout)
        hung = [p for p in procs if p.is_alive()]
        for p in hung:
            p.kill(); p.join()
        assert not hung, f"{len(hung)} band process(es) still waiting after {timeout} s: dead-lock in the band flags"
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        return np.ndarray((FRAMES, H, W, 4), np.float16, buffer=out.buf).copy()
    finally:
        for s in shms + [out]:
            s.close(); s.unlink()


@pytest.fixture(scope="module")
def whole_frame(built):
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    from tests.parity_util import SceneData, compare_radiance, run_oracle
    d = SceneData(**WORKLOAD)
    orc = run_oracle(d, keep_passes=False)
    tables = [(o["updated"], o["visible"], o["entries"]) for o in orc]
    lit = _run(1, tables, True)
    for f in range(FRAMES):          # the one-handle host run is itself the oracle's frame (radiance bar)
        r = compare_radiance(lit[f], orc[f]["combined"], 3)
        assert r["ok"] and r["bit_exact"] > 0.99, (f, r)
    return tables, lit


@pytest.mark.parametrize("world,kernel_signal", [(3, True), (4, True), (4, False), (5, True), (8, True)])
def test_bands_with_middle_ranks_reproduce_the_whole_frame_bytes(whole_frame, world, kernel_signal):
    if world == 8 and os.environ.get("SGI_HOST_SUITE", "") != "full":
        pytest.skip("8 bands of 30 rows (halo 20): set SGI_HOST_SUITE=full")
    if world == 5 and os.environ.get("SGI_HOST_SUITE", "") != "full":
        pytest.skip("5 unequal bands (25 s; 3 and 4 bands run every time): set SGI_HOST_SUITE=full")
    if not kernel_signal and os.environ.get("SGI_HOST_SUITE", "") != "full":
        pytest.skip("host-launched signals (the path the 4- and 8-GPU hardware runs used): set SGI_HOST_SUITE=full")
    tables, want = whole_frame
    got = _run(world, tables, kernel_signal)
    for f in range(FRAMES):
        assert np.array_equal(got[f].view(np.uint16), want[f].view(np.uint16)), f


@pytest.mark.parametrize("world", [3, 4, 5])
def test_bands_that_recompute_their_halo_reproduce_the_whole_frame_bytes(whole_frame, world):
    """SGI_HALO_RECOMPUTE with middle bands, one process per band: the gather and the non-final passes run on the halo rows the later
    passes read, nothing is exchanged inside the frame; the final pass stores its GI rows into the neighbours and releases their
    flags, the temporal pre-pass waits for the neighbours' previous frame.  Three frames, so that a final pass stores into the GI
    plane the neighbour's pre-pass read two frames before (the wait above is what orders the two)."""
    if world == 5 and os.environ.get("SGI_HOST_SUITE", "") != "full":
        pytest.skip("5 unequal bands of 40-60 rows with 30-row halos: set SGI_HOST_SUITE=full")
    tables, want = whole_frame
    got = _run(world, tables, True, recompute=True)
    for f in range(FRAMES):
        assert np.array_equal(got[f].view(np.uint16), want[f].view(np.uint16)), f


@pytest.mark.parametrize("drop_wait", [False, True], ids=["wait", "no-wait"])
def test_the_inter_frame_wait_is_what_orders_a_late_neighbours_gi_rows(whole_frame, drop_wait):
    """SGI_HALO_RECOMPUTE leaves ONE dependency between bands: the temporal pre-pass of frame f + 1 reads GI halo rows the neighbours'
    final passes of frame f stored, and waits for their "frame complete" flag first.  Band 1 of 3 is made late (it sleeps before each
    of its final passes).  With the wait the frames are still the one-handle bytes; with the wait dropped (test hook) bands 0 and 2
    run ahead, read the rows before they are there, and the bytes differ — so the comparison above is sensitive to that wait."""
    if os.environ.get("SGI_HOST_SUITE", "") != "full":
        pytest.skip("a minute per case (band 1 sleeps): set SGI_HOST_SUITE=full")
    tables, want = whole_frame
    hooks = [(0, drop_wait), (4, drop_wait), (0, drop_wait)]
    got = _run(3, tables, True, recompute=True, hooks=hooks)
    same = all(np.array_equal(got[f].view(np.uint16), want[f].view(np.uint16)) for f in range(FRAMES))
    assert same != drop_wait
