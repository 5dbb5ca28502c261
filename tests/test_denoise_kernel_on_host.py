"""csrc/k_denoise.cu compiled by g++ (tests/cuda_on_host/host_tu.py: a translated copy, the source is untouched) and executed thread
by thread on the host — the shipped kernels' LOGIC (the clamp-free interior path and the clamped edge path of the à-trous
kernel, the final pass's temporal epilogue) against the CPU oracle without a GPU.

The MUFU approximations (lg2/ex2) are libm calls here, so radiance is compared under the parity bar of tests/parity_util.py
(1e-3 relative per pixel or 60 dB PSNR) like on the GPU; the history plane — the one decision of this shader — must be exact."""
import ctypes as C
import os

import numpy as np
import pytest

from stratusgfx_b200._abi import FrameParams


def _load(defines=(), name="libdenoise_on_host.so"):
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    from tests.cuda_on_host.host_tu import compile_wrapper
    L = C.CDLL(compile_wrapper("denoise_on_host.cpp", "k_denoise.cu", name, defines))
    L.host_denoise.argtypes = [C.POINTER(FrameParams)] + [C.c_int] * 3 + [C.c_void_p] * 17
    return L


@pytest.fixture(scope="module")
def host_denoiser(built):
    return _load()


def _run(L, data, orc, f, iters):
    gb, pgb = data.gb[f], data.gb[f - 1]
    H, W = data.H, data.W
    prev_gi = orc[f - 1]["gi_illum"] if f > 0 else np.zeros((H, W, 4), np.float16)
    prev_hist = orc[f - 1]["history"] if f > 0 else np.zeros((H, W), np.float16)
    passes = np.zeros((iters, H, W, 4), np.float16)
    combined, gi, hist = np.zeros((H, W, 4), np.float16), np.zeros((H, W, 4), np.float16), np.zeros((H, W), np.float16)
    p = lambda a: np.ascontiguousarray(a).ctypes.data
    keep = [np.ascontiguousarray(a) for a in (gb.normal, gb.depth, gb.screen, gb.velocity, gb.id, pgb.normal, pgb.id, pgb.depth,
                                              orc[f]["gi_color"], orc[f]["gi_reservoir"], prev_gi, prev_hist, orc[f]["visible"])]
    rc = L.host_denoise(C.byref(data.fp[f]), W, H, iters, *[a.ctypes.data for a in keep], p(passes), p(combined), p(gi), p(hist))
    assert rc == 0
    return passes, combined, gi, hist


@pytest.mark.parametrize("yaw", [0.0, 1.5])
def test_denoiser_kernels_on_the_host_match_the_oracle(host_denoiser, yaw):
    """Static camera (every reprojected tap on its own texel) and a turning one (bilinear reprojection, history resets)."""
    from tests.parity_util import SceneData, compare_radiance, run_oracle
    d = SceneData("tiny", frames=2, yaw_deg_per_frame=yaw)
    iters = d.scene.denoise_iters
    orc = run_oracle(d)
    for f in range(2):
        passes, combined, gi, hist = _run(host_denoiser, d, orc, f, iters)
        for i in range(iters):
            r = compare_radiance(passes[i], orc[f]["pass_reservoirs"][i], 3)
            assert r["ok"] and r["bit_exact"] > 0.99, (f, i, r)
        for got, want in ((combined, orc[f]["combined"]), (gi, orc[f]["gi_illum"])):
            r = compare_radiance(got, want, 3)
            assert r["ok"] and r["bit_exact"] > 0.99, (f, r)
        assert np.array_equal(hist.view(np.uint16), orc[f]["history"].view(np.uint16))
    if yaw:
        assert len(np.unique(orc[1]["history"].view(np.uint16))) > 1          # some pixels were reset, some accumulated


@pytest.mark.parametrize("W,rows,c0,crows,s", [(3840, 2160, 0, 2160, 2), (3840, 2160, 0, 2160, 8), (3840, 330, 30, 270, 8), (3840, 330, 6, 318, 2),
                                               (3840, 300, 0, 298, 4), (1920, 420, 30, 360, 4), (640, 360, 0, 360, 8), (100, 77, 5, 60, 2),
                                               (170, 95, 0, 95, 2), (170, 95, 0, 95, 4), (1370, 770, 0, 770, 4), (1370, 770, 0, 770, 8), (1370, 330, 30, 270, 8)])
def test_border_ctas_are_enumerated_exactly_once(host_denoiser, W, rows, c0, crows, s):
    """A pass whose interior goes through the TMA tiles launches, ahead of the tiles, the CTAs of the clamped grid that reach outside the
    tiles' rectangle (csrc/sgi_denoise.cuh: borderGeometry / borderCtaCount / borderCta — arithmetic, no device list).  For whole
    frames, thin bands with and without halo rows above, and a plane smaller than any tile: the enumeration must be exactly the set of
    CTAs that atrousCta's own early exit does NOT send home, each once.  The rectangle is the one k_atrous_tma.cu's planTiles would give
    (whole lattice rows inside the computed rows, two lattice points from every plane edge)."""
    L = host_denoiser
    L.host_border_enumeration.argtypes = [C.c_int] * 9 + [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.host_border_enumeration.restype = C.c_int
    LX, LY = W // s, rows // s
    ix0, ix1 = 2, LX - 2
    iy0, iy1 = max(2, -(-c0 // s)), min(LY - 2, (c0 + crows) // s)
    rects = [(ix0 * s, max(ix0, ix1) * s, iy0 * s, max(iy0, iy1) * s), (0, 0, 0, 0), (0, W, c0, c0 + crows)]
    for x0, x1, y0, y1 in rects:
        cap = 1 << 16
        lst = np.full((cap, 2), -1, np.int32); grid = np.zeros(2, np.int32); inside = np.zeros(1 << 16, np.uint8)
        n = L.host_border_enumeration(W, rows, c0, crows, s, x0, x1, y0, y1, lst.ctypes.data, cap, grid.ctypes.data, inside.ctypes.data)
        gx, gy = int(grid[0]), int(grid[1])
        assert gx * gy <= inside.size and 0 <= n <= cap
        inside2d = inside[:gx * gy].reshape(gy, gx).astype(bool)
        want = {(bx, by) for by in range(gy) for bx in range(gx) if not inside2d[by, bx]}
        got = [tuple(int(v) for v in lst[i]) for i in range(n)]
        assert len(set(got)) == len(got) == n, "a border CTA is enumerated twice"
        assert set(got) == want, (len(got), len(want))


# ---- the TMA-staged à-trous kernel against the clamped kernel, pass by pass, on the host ---------------------------------------
_GEOMETRIES_FAST = [(200, 176, 0, 176), (170, 95, 0, 95), (256, 120, 28, 60), (164, 210, 30, 150)]
_GEOMETRIES_FULL = [(160, 90, 0, 90), (165, 95, 0, 95), (136, 72, 0, 72), (134, 70, 0, 70), (272, 330, 30, 270), (272, 330, 6, 318), (200, 300, 0, 298),
                    (320, 116, 28, 60), (320, 88, 0, 60), (320, 88, 28, 60), (1370, 64, 0, 64), (72, 400, 0, 400), (178, 181, 5, 171)]


def _atrous_pass_libs():
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    from tests.cuda_on_host.host_tu import compile_wrapper
    tma = C.CDLL(compile_wrapper(("atrous_pass_on_host.cpp", "atrous_tma_tu_on_host.cpp"), ("k_denoise.cu", "k_atrous_tma.cu"), "libatrous_tma_on_host.so", ("SGI_HOST_TMA",)))
    plain = C.CDLL(compile_wrapper("atrous_pass_on_host.cpp", "k_denoise.cu", "libatrous_plain_on_host.so", ()))
    for L in (tma, plain):
        L.host_atrous_pass.argtypes = [C.c_int] * 5 + [C.c_void_p] * 8
        L.host_tma_boxes.restype = C.c_longlong
    return tma, plain


def _random_planes(W, rows, seed):
    """A guide plane with structure (a few normal directions and depth steps, so that weights range from 0 to 1) and a random reservoir."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:rows, 0:W]
    region = ((xx // 13) + 2 * (yy // 9)) % 5
    normals = rng.normal(size=(5, 3)).astype(np.float32)
    normals /= np.linalg.norm(normals, axis=1, keepdims=True)
    n = normals[region] + rng.normal(scale=0.02, size=(rows, W, 3)).astype(np.float32)
    n /= np.linalg.norm(n, axis=2, keepdims=True)
    depth = (5.0 + 3.0 * region + rng.random((rows, W)) * 0.5).astype(np.float32)
    guide = np.concatenate([n.astype(np.float32), depth[..., None]], axis=2).astype(np.float32)
    res = rng.random((rows, W, 4)).astype(np.float16)
    res[..., 3] = 1.0
    return np.ascontiguousarray(guide), np.ascontiguousarray(res)


@pytest.mark.parametrize("W,rows,c0,crows", _GEOMETRIES_FAST + _GEOMETRIES_FULL)
def test_tma_tiles_and_border_ctas_write_the_clamped_kernels_bytes(built, W, rows, c0, crows):
    """csrc/k_atrous_tma.cu (lattice tiles from the shim's tensor maps + the border CTAs at the head of the same launch) against
    csrc/k_denoise.cu's atrousKernel on the whole plane, one pass per spacing — and the final pass with its blend epilogue at spacings
    8 and 2 — on the same random planes: whole frames and row bands
    (halo rows above / below the computed ones), leftover columns and rows beyond the last whole multiple of the spacing, a last tile
    shifted back or left partial, planes that only just hold one tile, planes that hold none (then both runs are the clamped kernel),
    an odd width (encode refused).  Same bytes in every computed pixel, and no pixel outside the computed rows touched by either."""
    if (W, rows, c0, crows) in _GEOMETRIES_FULL and os.environ.get("SGI_HOST_SUITE", "") != "full":
        pytest.skip("the long list of geometries: set SGI_HOST_SUITE=full")
    tma, plain = _atrous_pass_libs()
    guide, res = _random_planes(W, rows, seed=W * 1000 + rows)
    rng = np.random.default_rng(rows * 1000 + W)
    temporal = np.ascontiguousarray(rng.random((rows, W, 4)).astype(np.float32))          # (reprojected GI, 1 / history in (0, 1])
    screen, illum = (np.ascontiguousarray(rng.random((rows, W, 4)).astype(np.float16)) for _ in range(2))
    computed = np.zeros(rows, bool); computed[c0:c0 + crows] = True
    took_tiles = []
    for multiplier, final_ in ((1, False), (3, False), (7, False), (7, True), (1, True)):
        outs = []
        for L in (tma, plain):
            planes = [np.full((rows, W, 4), np.float16(-7.0)) for _ in range(3)]        # sentinel: no pass writes a negative value
            extra = [temporal.ctypes.data, screen.ctypes.data, illum.ctypes.data, planes[1].ctypes.data, planes[2].ctypes.data] if final_ else [None] * 5
            before = L.host_tma_boxes()
            assert L.host_atrous_pass(W, rows, c0, crows, multiplier, guide.ctypes.data, res.ctypes.data, planes[0].ctypes.data, *extra) >= 1
            if L is tma and not final_:
                took_tiles.append(L.host_tma_boxes() > before)
            outs.append(planes)
        for k, (pa, pb) in enumerate(zip(*outs)):                                           # reservoir; final pass: lit image, GI
            a, b = pa.view(np.uint16), pb.view(np.uint16)
            assert np.array_equal(a, b), (multiplier, final_, k, int((a != b).any(axis=2).sum()))
            written = k == 0 or final_
            assert np.all(pa[~computed] == np.float16(-7.0)) and np.all((pa[computed][..., 3] == np.float16(1.0)) == written)
    expect = {(200, 176, 0, 176): [True, True, True], (170, 95, 0, 95): [True, True, False], (165, 95, 0, 95): [False, False, False]}
    if (W, rows, c0, crows) in expect:
        assert took_tiles == expect[(W, rows, c0, crows)], took_tiles
