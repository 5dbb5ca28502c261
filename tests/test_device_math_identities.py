"""Identities the CUDA kernels rely on to stay bit-exact while avoiding IEEE division sequences.

sgi_math.cuh computes b/255.0f and u/65535.0f as   q = x*r;  q' = fma(fma(-q, D, x), r, q)   with
r = fl(1/D).  Here every input is checked against the correctly rounded quotient using exact
rational arithmetic (no floating-point evaluation of the fma), so the claim does not depend on
any hardware."""
import math
from fractions import Fraction

import numpy as np


def round_f32(fr):
    """Fraction -> nearest-even binary32 (normal range), returned as a Python float."""
    if fr == 0:
        return 0.0
    sign = -1 if fr < 0 else 1
    fr = abs(fr)
    e = math.floor(math.log2(fr))
    while Fraction(2) ** e > fr:
        e -= 1
    while Fraction(2) ** (e + 1) <= fr:
        e += 1
    scale = Fraction(2) ** (e - 23)
    m = fr / scale
    lo = m.numerator // m.denominator
    rem = m - lo
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and lo % 2 == 1):
        lo += 1
    return sign * float(lo * scale)


def _check(D, n):
    r = Fraction(float(np.float32(1.0) / np.float32(D)))
    for b in range(n):
        q = Fraction(round_f32(Fraction(b) * r))
        rem = Fraction(round_f32(Fraction(b) - q * D))        # fma(-q, D, x): one rounding
        q2 = round_f32(rem * r + q)                            # fma(rem, r, q): one rounding
        assert q2 == float(np.float32(b) / np.float32(D)), (D, b)


def test_unorm8_identity_is_exact_for_all_256_inputs():
    _check(255, 256)


def test_unorm16_identity_is_exact_for_all_65536_inputs():
    _check(65535, 65536)


def test_plain_reciprocal_multiply_would_not_be_exact():
    # why the correction step exists
    b = np.arange(256, dtype=np.float32)
    assert np.any(b * (np.float32(1.0) / np.float32(255.0)) != b / np.float32(255.0))


# ---- two-tier decisions of the gather (csrc/k_gather.cu, csrc/sgi_math.cuh): the cheap route may only answer when its answer is
#      the exact route's.  Both properties are checked here against every estimate the stated error bound allows, in IEEE binary32
#      arithmetic (numpy), so the claim does not depend on how a particular GPU rounds rcp.approx. -------------------------------

def _ulp_step(x, k):
    """x (float32 array, positive) moved by k units in the last place"""
    return (x.view(np.int32) + np.int32(k)).view(np.float32)


def test_pass0_epilogue_estimate_rounds_like_the_division_whenever_it_is_sure():
    """mergedChannel (k_gather.cu): f2h(a / denom) for binary16 a, denom >= 1 is taken from q = a * rcp(denom) unless q is within
    8 ulp of a binary16 rounding boundary or below binary16's normal range.  The stated bound: |q - fl(a / denom)| <= 4 ulp.  For
    every binary16 a, a spread of denominators and every q within 4 ulp of the IEEE quotient: sure(q) implies f2h(q) == f2h(a / denom)."""
    a = np.arange(1, 0x7c00, dtype=np.uint16).view(np.float16).astype(np.float32)                       # every positive finite binary16
    rng = np.random.default_rng(3)
    denoms = np.unique(np.concatenate([np.float16([1.0, 1.5, 2.0, 3.0, 5.999, 6.0, 65504.0]),
                                       rng.uniform(1.0, 6.0, 60).astype(np.float16), rng.uniform(1.0, 60000.0, 20).astype(np.float16)])).astype(np.float32)
    for d in denoms:
        exact = a / d                                                                                      # IEEE binary32 division
        want = exact.astype(np.float16).view(np.uint16)
        for k in range(-4, 5):
            q = _ulp_step(exact, k)
            b = q.view(np.uint32)
            sure = (((b + np.uint32(8) - np.uint32(0x1000)) & np.uint32(0x1fff)) > 16) & (q >= np.float32(6.2e-5))
            got = q.astype(np.float16).view(np.uint16)
            assert np.array_equal(got[sure], want[sure]), (float(d), k)
        assert sure.mean() > 0.75 or d > 100         # (and the cheap route is the common one: the rest are quotients below binary16's normal range)


def test_cube_texel_estimate_agrees_with_the_exact_chain_whenever_it_is_sure():
    """cubeTexelDecided (sgi_math.cuh): texel coordinate floor(fl(fl(fl(sc / ma) + 1) * 0.5) * res) is taken from the estimate
    se = fma(sc * r, res / 2, res / 2), r = rcp.approx(ma) (1 ulp), unless frac(se) is within eps = res 2^-22 (|q| + 1) of 0 or 1.
    Random directions and the planes of the cube (sc / ma tiny), every reciprocal within 1 ulp: sure implies the same texel."""
    rng = np.random.default_rng(11)
    n = 400000
    ma = rng.uniform(0.05, 400.0, n).astype(np.float32)
    ratio = rng.uniform(-1.0, 1.0, n).astype(np.float32)
    ratio[: n // 4] = (rng.uniform(-1.0, 1.0, n // 4) * 1e-5).astype(np.float32)                          # a pixel in the plane of its light
    ratio[n // 4: n // 2] = (np.round(rng.uniform(-16, 16, n // 4)) / 16.0 + rng.uniform(-1, 1, n // 4) * 3e-6).astype(np.float32)   # next to texel edges
    sc = (ratio * ma).astype(np.float32)
    sc = np.clip(sc, -ma, ma)
    for res in (32, 64, 17):
        fres, half, epsk = np.float32(res), np.float32(0.5 * res), np.float32(res * 2.0 ** -22)
        q = sc / ma
        exact = np.floor((((q + np.float32(1.0)).astype(np.float32) * np.float32(0.5)).astype(np.float32) * fres).astype(np.float32))
        exact = np.clip(exact, 0, res - 1)
        r0 = (np.float32(1.0) / ma).astype(np.float32)
        for k in (-1, 0, 1):
            r = _ulp_step(r0, k)
            qe = (sc * r).astype(np.float32)
            se = (qe.astype(np.float64) * np.float64(half) + np.float64(half)).astype(np.float32)         # fma: one rounding
            sf = np.floor(se)
            eps = (np.abs(qe).astype(np.float64) * np.float64(epsk) + np.float64(epsk)).astype(np.float32)
            sure = np.abs((se - sf) - np.float32(0.5)) < (np.float32(0.5) - eps)
            assert np.array_equal(sf[sure], exact[sure]), (res, k, int((sf[sure] != exact[sure]).sum()))
            assert sure.mean() > 0.7


def test_redraw_estimate_agrees_with_the_exact_chain_whenever_it_is_sure():
    """behindOrOutOfRange (k_gather.cu): `dot(normal, normalize(l)) < 0 || length(l) > radius` (vpl_pbr_gi.fs:164-166) is answered from
    s = n.l and |l|^2 (FMA) when s^2 > 1e-10 |l|^2 and |l|^2 is outside radius^2 (1 +- 4e-5).  Adversarial inputs — light directions
    almost in the surface plane, distances within 1e-4 of the radius — and every estimate within 2 ulp of the FMA values: a sure
    estimate gives the decision of the exact chain in glm's operation order."""
    rng = np.random.default_rng(23)
    n = 300000
    f32 = np.float32
    nrm = rng.normal(size=(n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm = nrm.astype(f32)
    l = rng.normal(size=(n, 3)) * rng.uniform(0.1, 80.0, (n, 1))
    # a third of the directions nearly perpendicular to the normal
    k = n // 3
    perp = l[:k] - (l[:k] * nrm[:k]).sum(1, keepdims=True) * nrm[:k]
    l[:k] = perp + nrm[:k] * (rng.uniform(-1, 1, (k, 1)) * 1e-5 * np.linalg.norm(perp, axis=1, keepdims=True))
    l = l.astype(f32)
    radius = np.linalg.norm(l.astype(np.float64), axis=1)
    radius[k:2 * k] *= 1.0 + rng.uniform(-1, 1, k) * 1e-4                                                 # distance next to the radius
    radius[2 * k:] *= rng.uniform(0.5, 2.0, n - 2 * k)
    radius = radius.astype(f32)
    lx, ly, lz = l[:, 0], l[:, 1], l[:, 2]
    nx, ny, nz = nrm[:, 0], nrm[:, 1], nrm[:, 2]
    # exact chain, glm order: dot = (x + y) + z, length = sqrt(dot), normalize = v * (1 / sqrt(dot))
    d = ((lx * lx + ly * ly).astype(f32) + lz * lz).astype(f32)
    dist = np.sqrt(d).astype(f32)
    inv = (f32(1.0) / dist).astype(f32)
    side = (((nx * (lx * inv).astype(f32)).astype(f32) + (ny * (ly * inv).astype(f32)).astype(f32)).astype(f32) + (nz * (lz * inv).astype(f32)).astype(f32)).astype(f32)
    exact = (side < 0) | (dist > radius)
    # estimates: FMA chains (one rounding each), then every neighbour within 2 ulp
    fma = lambda a, b, c: (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)
    s0 = fma(nz, lz, fma(ny, ly, (nx * lx).astype(f32)))
    d20 = fma(lz, lz, fma(ly, ly, (lx * lx).astype(f32)))
    r2 = (radius * radius).astype(f32)
    m = f32(1e-5)
    for ks in (-2, 0, 2):
        s = np.where(s0 != 0, (s0.view(np.int32) + np.int32(ks)).view(f32), s0)
        for kd in (-2, 0, 2):
            d2 = _ulp_step(d20, kd)
            side_sure = (s * s).astype(f32) > ((m * m).astype(f32) * d2).astype(f32)
            far = d2 > (r2 * (f32(1.0) + f32(4.0) * m)).astype(f32)
            near = d2 < (r2 * (f32(1.0) - f32(4.0) * m)).astype(f32)
            sure = side_sure & (far | near)
            got = (s < 0) | far
            assert np.array_equal(got[sure], exact[sure]), (ks, kd, int((got[sure] != exact[sure]).sum()))
    assert sure.mean() > 0.5


def test_shadow_verdict_estimate_agrees_with_the_exact_chain_whenever_it_is_sure():
    """shadeAcceptedLight (k_gather.cu): lit / shadowed is `currentDepth - currentDepth * bias > texelDepth * farPlane`
    (calculateShadowValue1Sample, pbr.glsl:116-136).  The estimate uses |l| = d2 * rsqrt(d2) (<= 3 ulp), texel * (1 / 65535) (1 ulp)
    and distance * (1 - bias), and answers only when the two sides differ by more than 8e-5 of the larger.  Texel depths placed
    within 2e-4 of the decision, every estimate inside those bounds: a sure estimate gives the exact verdict."""
    rng = np.random.default_rng(31)
    f32 = np.float32
    n = 300000
    frag_dir = (rng.normal(size=(n, 3)) * rng.uniform(0.5, 200.0, (n, 1))).astype(f32)
    lnd = (rng.uniform(-5.0, 40.0, n)).astype(f32)                       # dot(lightMinusFrag, normal): both sides of the 0.1 / 0.03 switch
    far_plane = rng.uniform(50.0, 600.0, n).astype(f32)
    x, y, z = frag_dir[:, 0], frag_dir[:, 1], frag_dir[:, 2]
    d = ((x * x + y * y).astype(f32) + z * z).astype(f32)
    cur = np.sqrt(d).astype(f32)                                          # length(fragDir), glm order
    sat = np.clip(lnd, f32(0.0), f32(1.0))
    bias_f = np.maximum(f32(0.1), (sat * f32(0.03)).astype(f32))
    lhs_exact = (cur - (cur * bias_f).astype(f32)).astype(f32)
    # a 16-bit texel whose depth * farPlane lands next to lhs (or anywhere, for a third of the samples)
    target = lhs_exact.astype(np.float64) / far_plane * (1.0 + rng.uniform(-1, 1, n) * 2e-4)
    target[: n // 3] = rng.uniform(0, 1, n // 3)
    texel = np.clip(np.round(target * 65535.0), 0, 65535).astype(np.uint32)
    depth_exact = (texel.astype(f32) / f32(65535.0)).astype(f32)
    exact = lhs_exact > (depth_exact * far_plane).astype(f32)
    d2 = (z.astype(np.float64) * z + (y.astype(np.float64) * y + (x * x).astype(f32))).astype(f32)       # FMA chain (one rounding each is finer than this)
    dist0 = np.sqrt(d2.astype(np.float64)).astype(f32)
    depth0 = (texel.astype(f32) * f32(1.0 / 65535.0)).astype(f32)
    for kd in (-3, 0, 3):
        dist = _ulp_step(dist0, kd)
        for kt in (-1, 0, 1):
            dep = np.where(texel > 0, _ulp_step(np.maximum(depth0, f32(1e-30)), kt), depth0)
            lhs = (dist * (f32(1.0) - bias_f).astype(f32)).astype(f32)
            rhs = (dep * far_plane).astype(f32)
            gap = (lhs - rhs).astype(f32)
            sure = np.abs(gap) > (f32(8.0) * f32(1e-5) * np.maximum(lhs, rhs)).astype(f32)
            got = gap > 0
            assert np.array_equal(got[sure], exact[sure]), (kd, kt, int((got[sure] != exact[sure]).sum()))
    assert 0.3 < sure.mean() < 0.95                                       # (the adversarial two thirds really sit inside the margin)


def test_stage2_pretests_only_drop_lights_the_exact_chain_drops():
    """stage2Kernel (k_tiles.cu): before the exact chain of stage2.cs:86-94 a light is dropped when its FMA distance estimate says it is
    surely beyond the tile's current threshold (d2e > radius^2 (1 + 4e-5) w^2 (1 + 1e-6)) or surely behind the tile
    (n.l < 0 with (n.l)^2 > 1e-10 d2e).  Both must be conservative: a dropped light has exact ratio = length(l) / radius > w, or an
    exact side check < 0.  Thresholds within 1e-4 of the true ratio and directions within 1e-5 of the tile plane."""
    rng = np.random.default_rng(41)
    f32 = np.float32
    n = 300000
    l = (rng.normal(size=(n, 3)) * rng.uniform(0.1, 60.0, (n, 1))).astype(f32)
    nrm = rng.normal(size=(n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    k = n // 2
    perp = l[:k].astype(np.float64) - (l[:k] * nrm[:k]).sum(1, keepdims=True) * nrm[:k]
    l[:k] = (perp + nrm[:k] * (rng.uniform(-1, 1, (k, 1)) * 1e-5 * np.linalg.norm(perp, axis=1, keepdims=True))).astype(f32)
    nrm = nrm.astype(f32)
    lx, ly, lz = l[:, 0], l[:, 1], l[:, 2]
    nx, ny, nz = nrm[:, 0], nrm[:, 1], nrm[:, 2]
    d = ((lx * lx + ly * ly).astype(f32) + lz * lz).astype(f32)
    radius = (np.sqrt(d.astype(np.float64)) * rng.uniform(0.8, 3.0, n)).astype(f32)
    ratio = (np.sqrt(d).astype(f32) / radius).astype(f32)                                                  # fdiv(length(l), radius)
    side = (((nx * lx).astype(f32) + (ny * ly).astype(f32)).astype(f32) + (nz * lz).astype(f32)).astype(f32)   # dot(normal, l), glm order
    w = np.minimum(f32(1.0), (ratio.astype(np.float64) * (1.0 + rng.uniform(-1, 1, n) * 1e-4)).astype(f32))
    w[::3] = rng.uniform(0.0, 1.0, len(w[::3])).astype(f32)
    w2 = ((w * w).astype(f32) * (f32(1.0) + f32(1e-6))).astype(f32)
    fma = lambda a, b, c: (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)
    d2e0 = fma(lz, lz, fma(ly, ly, (lx * lx).astype(f32)))
    se0 = fma(nz, lz, fma(ny, ly, (nx * lx).astype(f32)))
    for kd in (-2, 0, 2):
        d2e = _ulp_step(d2e0, kd)
        far = d2e > (((radius * radius).astype(f32) * (f32(1.0) + f32(4e-5))).astype(f32) * w2).astype(f32)
        assert (ratio[far] > w[far]).all(), kd
        behind = (se0 < 0) & ((se0 * se0).astype(f32) > (f32(1e-10) * d2e).astype(f32))
        assert (side[behind] < 0).all(), kd
    assert far.mean() > 0.1 and behind.mean() > 0.1
