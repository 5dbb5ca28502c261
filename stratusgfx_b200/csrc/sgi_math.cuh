// sgi_math.cuh — device maths for the VPL GI kernels.
//
// The integer outputs of this path (visible-VPL list, per-tile lists, the gather's
// light choice) are decided by float comparisons, so every operation that feeds a
// decision is written with explicit round-to-nearest intrinsics (__fmul_rn/__fadd_rn
// are never contracted into FMA by nvcc, __fdiv_rn/__fsqrt_rn are IEEE-exact) in the
// evaluation order of glm 0.9.8.5 (GLM_FORCE_PURE), the maths library of the
// reference's C++ side:
//     dot3 = (x+y)+z        dot4 = (x+y)+(z+w)       length = sqrt(dot)
//     normalize = v * (1/sqrt(dot))                  mat4*vec4 = (m0v0+m1v1)+(m2v2+m3v3)
//     mix(x,y,a) = x + a*(y-x)    clamp = min(max(x,lo),hi)   min(x,y) = x<y?x:y
// The .cu files are additionally compiled with -fmad=false so plain scalar
// expressions are not contracted either.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace sgi {

#define SGI_DEV __device__ __forceinline__

struct f2 { float x, y; };
struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };
struct m4 { f4 c[4]; };   // column-major

SGI_DEV float fmul(float a, float b) { return __fmul_rn(a, b); }
SGI_DEV float fadd(float a, float b) { return __fadd_rn(a, b); }
SGI_DEV float fsub(float a, float b) { return __fadd_rn(a, -b); }
SGI_DEV float fdiv(float a, float b) { return __fdiv_rn(a, b); }
SGI_DEV float fsqrt(float a) { return __fsqrt_rn(a); }

SGI_DEV f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
SGI_DEV f3 mk3(float s) { return mk3(s, s, s); }
SGI_DEV f4 mk4(float x, float y, float z, float w) { f4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

SGI_DEV f3 operator+(f3 a, f3 b) { return mk3(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)); }
SGI_DEV f3 operator-(f3 a, f3 b) { return mk3(fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)); }
SGI_DEV f3 operator*(f3 a, f3 b) { return mk3(fmul(a.x, b.x), fmul(a.y, b.y), fmul(a.z, b.z)); }
SGI_DEV f3 operator/(f3 a, f3 b) { return mk3(fdiv(a.x, b.x), fdiv(a.y, b.y), fdiv(a.z, b.z)); }
SGI_DEV f3 operator*(f3 a, float s) { return mk3(fmul(a.x, s), fmul(a.y, s), fmul(a.z, s)); }
SGI_DEV f3 operator*(float s, f3 a) { return mk3(fmul(s, a.x), fmul(s, a.y), fmul(s, a.z)); }
SGI_DEV f3 operator/(f3 a, float s) { return mk3(fdiv(a.x, s), fdiv(a.y, s), fdiv(a.z, s)); }
SGI_DEV f3 operator/(float s, f3 a) { return mk3(fdiv(s, a.x), fdiv(s, a.y), fdiv(s, a.z)); }
SGI_DEV f3 operator+(f3 a, float s) { return mk3(fadd(a.x, s), fadd(a.y, s), fadd(a.z, s)); }
SGI_DEV f3 operator+(float s, f3 a) { return mk3(fadd(s, a.x), fadd(s, a.y), fadd(s, a.z)); }
SGI_DEV f3 operator-(f3 a, float s) { return mk3(fsub(a.x, s), fsub(a.y, s), fsub(a.z, s)); }
SGI_DEV f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }

SGI_DEV f4 operator+(f4 a, f4 b) { return mk4(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z), fadd(a.w, b.w)); }
SGI_DEV f4 operator*(f4 a, float s) { return mk4(fmul(a.x, s), fmul(a.y, s), fmul(a.z, s), fmul(a.w, s)); }
SGI_DEV f4 operator*(float s, f4 a) { return mk4(fmul(s, a.x), fmul(s, a.y), fmul(s, a.z), fmul(s, a.w)); }

SGI_DEV float vmin(float x, float y) { return x < y ? x : y; }
SGI_DEV float vmax(float x, float y) { return x > y ? x : y; }
SGI_DEV f3 vmin(f3 a, float s) { return mk3(vmin(a.x, s), vmin(a.y, s), vmin(a.z, s)); }
SGI_DEV float vclamp(float x, float lo, float hi) { return vmin(vmax(x, lo), hi); }
SGI_DEV float saturate(float x) { return vclamp(x, 0.0f, 1.0f); }
SGI_DEV int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

SGI_DEV float dot(f3 a, f3 b) { return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z)); }
SGI_DEV float dot(f4 a, f4 b) { return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fadd(fmul(a.z, b.z), fmul(a.w, b.w))); }
SGI_DEV float length(f3 v) { return fsqrt(dot(v, v)); }
SGI_DEV f3 normalize(f3 v) { return v * fdiv(1.0f, fsqrt(dot(v, v))); }
SGI_DEV float mix(float x, float y, float a) { return fadd(x, fmul(a, fsub(y, x))); }
SGI_DEV f3 mix(f3 x, f3 y, float a) { return x + a * (y - x); }

SGI_DEV f4 mul(const m4& m, f4 v) {
    const f4 add0 = (m.c[0] * v.x) + (m.c[1] * v.y);
    const f4 add1 = (m.c[2] * v.z) + (m.c[3] * v.w);
    return add0 + add1;
}
SGI_DEV m4 loadMat(const float* p) {
    m4 m;
    #pragma unroll
    for (int c = 0; c < 4; ++c) m.c[c] = mk4(p[c * 4], p[c * 4 + 1], p[c * 4 + 2], p[c * 4 + 3]);
    return m;
}

// ---- RNG: Source/Shaders/common.glsl:183-214 (pure integer arithmetic) ----------------------
SGI_DEV uint32_t hash1(uint32_t x) {
    x += (x << 10u); x ^= (x >> 6u); x += (x << 3u); x ^= (x >> 11u); x += (x << 15u);
    return x;
}
SGI_DEV float floatConstruct(uint32_t m) {
    m &= 0x007FFFFFu; m |= 0x3F800000u;
    return fsub(__uint_as_float(m), 1.0f);
}
SGI_DEV float random3(float x, float y, float z) {
    return floatConstruct(hash1(__float_as_uint(x) ^ hash1(__float_as_uint(y)) ^ hash1(__float_as_uint(z))));
}

// ---- common.glsl:143-155 ---------------------------------------------------------------------
SGI_DEV f3 worldPositionFromDepth(float u, float v, float depth, const m4& invProjectionView) {
    const float z = fsub(fmul(depth, 2.0f), 1.0f);
    const f4 ndc = mk4(fsub(fmul(u, 2.0f), 1.0f), fsub(fmul(v, 2.0f), 1.0f), z, 1.0f);
    const f4 wp = mul(invProjectionView, ndc);
    return mk3(fdiv(wp.x, wp.w), fdiv(wp.y, wp.w), fdiv(wp.z, wp.w));
}

// b / 255.0f and u / 65535.0f, bit-identical to the IEEE division for every 8-/16-bit input, in three
// instructions: q = x*r, one FMA residual, one FMA correction (r = fl(1/D)).  Verified exhaustively with
// exact rational arithmetic in tests/test_device_math_identities.py.
SGI_DEV float unorm8(uint32_t b) {
    const float x = (float)b, r = 1.0f / 255.0f;
    const float q = __fmul_rn(x, r);
    return __fmaf_rn(__fmaf_rn(-q, 255.0f, x), r, q);
}
SGI_DEV float unorm16(uint32_t u) {
    const float x = (float)u, r = 1.0f / 65535.0f;
    const float q = __fmul_rn(x, r);
    return __fmaf_rn(__fmaf_rn(-q, 65535.0f, x), r, q);
}

// normalize(texel.rgb * 2.0 - vec3(1.0)) from a packed rgba8 texel (common.glsl:216-220)
SGI_DEV f3 decodeNormal(uint32_t rgba) {
    const f3 t = mk3(unorm8(rgba & 0xffu), unorm8((rgba >> 8) & 0xffu), unorm8((rgba >> 16) & 0xffu));
    return normalize(t * 2.0f - 1.0f);
}

// ---- cube-map array lookup, NEAREST (GL 4.6 table 8.19) ---------------
SGI_DEV int cubeTexel(f3 d, int res) {
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int face; float sc, tc, ma;
    if (ax >= ay && ax >= az) { ma = ax; if (d.x >= 0.0f) { face = 0; sc = -d.z; tc = -d.y; } else { face = 1; sc = d.z; tc = -d.y; } }
    else if (ay >= az)        { ma = ay; if (d.y >= 0.0f) { face = 2; sc = d.x; tc = d.z; } else { face = 3; sc = d.x; tc = -d.z; } }
    else                      { ma = az; if (d.z >= 0.0f) { face = 4; sc = d.x; tc = -d.y; } else { face = 5; sc = -d.x; tc = -d.y; } }
    const float s = fmul(fadd(fdiv(sc, ma), 1.0f), 0.5f), t = fmul(fadd(fdiv(tc, ma), 1.0f), 0.5f);
    const int i = iclamp((int)floorf(fmul(s, (float)res)), 0, res - 1);
    const int j = iclamp((int)floorf(fmul(t, (float)res)), 0, res - 1);
    return (face * res + j) * res + i;
}

// ---- the same lookup as a two-tier decision (k_gather.cu, k_direct.cu) ------------------------------------------------------
SGI_DEV float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
SGI_DEV float rcpFast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
static __device__ __noinline__ int exactCubeTexel(f3 fragDir, int res) { return cubeTexel(fragDir, res); }      // (rarely executed)
// Cube texel with exact semantics.  The face is chosen by exact comparisons of the inputs (GL 4.6 table 8.19: signs folded in with
// bit operations, -0 as a major component only when the direction is all zeros, which the NaN of the estimate sends to the exact
// chain).  The texel coordinates are estimated as se = fma(sc * rcp(ma), res/2, res/2) and only a coordinate within `eps` of a texel
// edge needs the exact division chain of cubeTexel (sgi_math.cuh).  Bound: q = fl(sc/ma), qe = sc * rcp.approx(ma) differ by at most
// |q| 2^-22 (1 ulp of the reciprocal, one product, one division); the exact chain fl(fl(fl(q + 1) / 2) res) is within
// res (2^-25 + 2^-24) of (q + 1) res / 2 and the estimate within res 2^-24 of (qe + 1) res / 2, so |se - exact| <=
// res (|q| 2^-23 + 2.5 * 2^-24) and eps = res 2^-22 (|qe| + 1) keeps a factor 1.6 to 2 over it.  (Round 1 used a flat 1e-4 texels,
// round 2 first 3e-5: pixels that lie in the plane of a VPL — a floor pixel under a floor VPL — have |q| ~ 1e-6, that is se within
// 1e-5 .. 1e-4 of the middle edge, and took the exact chain on 70 % of the warp iterations; profiles/r02k-q.)
SGI_DEV int cubeTexelDecided(f3 d, int res, float halfRes, float epsK) {
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    const bool xMajor = ax >= ay && ax >= az, yMajor = !xMajor && ay >= az;
    const uint32_t bx = __float_as_uint(d.x), by = __float_as_uint(d.y), bz = __float_as_uint(d.z), kSign = 0x80000000u;
    const float ma = xMajor ? ax : (yMajor ? ay : az);
    const uint32_t major = xMajor ? bx : (yMajor ? by : bz);                                  // its sign bit: the negative face
    const int face = (xMajor ? 0 : (yMajor ? 2 : 4)) + (int)(major >> 31);
    // sc: +x: -z, -x: z | y: x | +z: x, -z: -x          tc: +y: z, -y: -z | others: -y
    const float sc = __uint_as_float(xMajor ? (bz ^ (~bx & kSign)) : (yMajor ? bx : (bx ^ (bz & kSign))));
    const float tc = __uint_as_float(yMajor ? (bz ^ (by & kSign)) : (by ^ kSign));
    const float inv = rcpFast(ma);
    const float qs = sc * inv, qt = tc * inv;
    const float se = fma_(qs, halfRes, halfRes), te = fma_(qt, halfRes, halfRes);
    const float sf = floorf(se), tf = floorf(te);
    const float es = fma_(fabsf(qs), epsK, epsK), et = fma_(fabsf(qt), epsK, epsK);
    // eps < frac < 1 - eps  <=>  |frac - 0.5| < 0.5 - eps   (false for a NaN estimate)
    const bool sure = fabsf((se - sf) - 0.5f) < 0.5f - es && fabsf((te - tf) - 0.5f) < 0.5f - et;
    if (!sure) return exactCubeTexel(d, res);
    // a sure coordinate lies strictly inside [0, res): no clamp
    return (face * res + (int)tf) * res + (int)sf;
}


// ---- binary16 planes --------------------------------------------------------------------------
SGI_DEV float h2f(unsigned short h) { return __half2float(__ushort_as_half(h)); }
SGI_DEV unsigned short f2h(float f) { return __half_as_ushort(__float2half_rn(f)); }
SGI_DEV f4 loadHalf4(const uint2* plane, size_t i) {
    const uint2 v = __ldg(plane + i);
    return mk4(h2f((unsigned short)(v.x & 0xffffu)), h2f((unsigned short)(v.x >> 16)),
               h2f((unsigned short)(v.y & 0xffffu)), h2f((unsigned short)(v.y >> 16)));
}
SGI_DEV uint2 packHalf4(float r, float g, float b, float a) {
    uint2 v;
    v.x = (uint32_t)f2h(r) | ((uint32_t)f2h(g) << 16);
    v.y = (uint32_t)f2h(b) | ((uint32_t)f2h(a) << 16);
    return v;
}
SGI_DEV void storeHalf4(uint2* plane, size_t i, float r, float g, float b, float a) { plane[i] = packHalf4(r, g, b, a); }

}  // namespace sgi
