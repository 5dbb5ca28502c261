// k_taa.cu — temporal anti-aliasing (second half): the reprojection-and-clamp resolve at the end of the frame,
// which shares its velocity reprojection with the denoiser's final pass.
//
// Replaces taa.fs:24-100 (reproject the previous frame by the velocity plane, clamp it to the tonemapped min/max of the current
// pixel's cross neighbourhood, blend 0.9 / 0.1 in tonemapped-inverse space; tonemap / inverseTonemap of aa_common.glsl:17-25) as
// driven by PerformTaaPostFx_ (StratusRendererBackend.cpp:2324-2354), including the blit that turns the result into the next
// frame's history.  Textures as the host creates them: `screen` RGBA8 LINEAR (:694-696; sampled at texel centres and whole-texel
// offsets, i.e. single texels), `prevScreen` RGB8 LINEAR, CLAMP_TO_EDGE (:450-452; the one filtered read), velocity RG16F NEAREST
// (:412-414), the target RGBA16F (:708-710).  Exact tier throughout (csrc/sgi_math.cuh): the output equals the compiled shader text
// bit for bit; the filtered read is the 8-bit sub-texel weight model of SURVEY.md Appendix A1.  The history store follows the GL
// float -> UNORM8 conversion (clamp to [0,1], scale by 255, round to nearest) applied to the binary16 value the blit reads.
//
// Parity: bit-identical on a B200 to the reference's taa.fs compiled by oracle/make_ref_glsl.py (tests/test_zz_taa.py, profiles/r02a),
// and on the host (tests/test_taa_kernel_on_host.py).
#include "sgi_internal.h"
#include "sgi_math.cuh"

namespace sgi {
namespace {

// One RGB8 texel as three floats (b / 255.0f: unorm8 of sgi_math.cuh is that division bit for bit, in three instructions and
// without the slow path a zero numerator sends __fdiv_rn down)
SGI_DEV uint32_t texel8(const uint32_t* __restrict__ plane, int W, int H, int i, int j) {
    return __ldg(plane + (size_t)iclamp(j, 0, H - 1) * W + iclamp(i, 0, W - 1));
}
SGI_DEV f3 rgbOf(uint32_t t) { return mk3(unorm8(t & 0xffu), unorm8((t >> 8) & 0xffu), unorm8((t >> 16) & 0xffu)); }
// a / b for b > 0, with a zero numerator (black history texels are common) returned as it is instead of taking __fdiv_rn's slow path
SGI_DEV float fdivPos(float a, float b) { return a == 0.0f ? a : fdiv(a, b); }
SGI_DEV f3 tonemap(f3 c) { return mk3(fdivPos(c.x, fadd(c.x, 1.0f)), fdivPos(c.y, fadd(c.y, 1.0f)), fdivPos(c.z, fadd(c.z, 1.0f))); }      // aa_common.glsl:17-20: color / (color + vec3(1.0))
SGI_DEV f3 inverseTonemap(f3 c) { return mk3(fdivPos(c.x, fsub(1.0f, c.x)), fdivPos(c.y, fsub(1.0f, c.y)), fdivPos(c.z, fsub(1.0f, c.z))); }      // aa_common.glsl:22-24: color / (vec3(1.0) - color)
SGI_DEV f3 min3(f3 a, f3 b) { return mk3(b.x < a.x ? b.x : a.x, b.y < a.y ? b.y : a.y, b.z < a.z ? b.z : a.z); }       // glm::min(x, y) = y < x ? y : x
SGI_DEV f3 max3(f3 a, f3 b) { return mk3(a.x < b.x ? b.x : a.x, a.y < b.y ? b.y : a.y, a.z < b.z ? b.z : a.z); }       // glm::max(x, y) = x < y ? y : x

// Every value the shader tonemaps from the CURRENT frame is a texel component, i.e. one of 256 floats k / 255: the centre texel and
// the component-wise minimum / maximum of the five texels of its cross (b / 255 is monotonic in b, so min and max commute with it).
// tonemap(k / 255) and inverseTonemap(tonemap(k / 255)) therefore come from two 256-entry tables, built by the same exact operations
// right before the resolve (taaTableKernel, one CTA); only the filtered history read keeps its divisions.  39 of the 47 IEEE divisions
// per pixel go away (0.559 -> see profiles/r02v for c4).
constexpr int kTaaTable = 256;            // floats per table; tables[0 .. 256) = tonemap, [256 .. 512) = inverseTonemap(tonemap)
__global__ void __launch_bounds__(kTaaTable)
taaTableKernel(float* __restrict__ tables) {
    const int k = threadIdx.x;
    const float c = unorm8((uint32_t)k);
    const float t = fdiv(c, fadd(c, 1.0f));
    tables[k] = t;
    tables[kTaaTable + k] = fdiv(t, fsub(1.0f, t));
}

struct TaaArgs { int W, H; const uint32_t* screen; const uint32_t* prevScreen; const uint32_t* velocity; uint2* colorOut; uint32_t* historyOut; const float* tables; };

SGI_DEV uint32_t min8(uint32_t a, uint32_t b) { return a < b ? a : b; }
SGI_DEV uint32_t max8(uint32_t a, uint32_t b) { return a > b ? a : b; }

__global__ void __launch_bounds__(256)
taaKernel(const __grid_constant__ TaaArgs a) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= a.W || y >= a.H) return;
    const int W = a.W, H = a.H;
    const size_t p = (size_t)y * W + x;
    const float u = fdiv(fadd((float)x, 0.5f), (float)W), v = fdiv(fadd((float)y, 0.5f), (float)H);
    const uint32_t vel = __ldg(a.velocity + p);
    const float pu = fsub(u, h2f((unsigned short)(vel & 0xffffu))), pv = fsub(v, h2f((unsigned short)(vel >> 16)));

    // texture(prevScreen, prevTexCoords): LINEAR, CLAMP_TO_EDGE
    f3 prevColor;
    {
        const float fu = fsub(fmul(pu, (float)W), 0.5f), fv = fsub(fmul(pv, (float)H), 0.5f);
        const float i0f = floorf(fu), j0f = floorf(fv);
        const float wa = fmul(floorf(fadd(fmul(fsub(fu, i0f), 256.0f), 0.5f)), 1.0f / 256.0f);
        const float wb = fmul(floorf(fadd(fmul(fsub(fv, j0f), 256.0f), 0.5f)), 1.0f / 256.0f);
        const int i0 = (int)i0f, j0 = (int)j0f;
        const f3 t00 = rgbOf(texel8(a.prevScreen, W, H, i0, j0)), t10 = rgbOf(texel8(a.prevScreen, W, H, i0 + 1, j0));
        const f3 t01 = rgbOf(texel8(a.prevScreen, W, H, i0, j0 + 1)), t11 = rgbOf(texel8(a.prevScreen, W, H, i0 + 1, j0 + 1));
        const float na = fsub(1.0f, wa), nb = fsub(1.0f, wb);
        const float w00 = fmul(na, nb), w10 = fmul(wa, nb), w01 = fmul(na, wb), w11 = fmul(wa, wb);
        prevColor.x = fadd(fadd(fadd(fmul(w00, t00.x), fmul(w10, t10.x)), fmul(w01, t01.x)), fmul(w11, t11.x));
        prevColor.y = fadd(fadd(fadd(fmul(w00, t00.y), fmul(w10, t10.y)), fmul(w01, t01.y)), fmul(w11, t11.y));
        prevColor.z = fadd(fadd(fadd(fmul(w00, t00.z), fmul(w10, t10.z)), fmul(w01, t01.z)), fmul(w11, t11.z));
    }
    prevColor = tonemap(prevColor);
    // textureOffset(screen, fsTexCoords, ivec2(..)): whole-texel offsets from a texel centre are single texels (CLAMP_TO_EDGE at the
    // border).  min / max of the cross per component on the 8-bit values, then the tables
    const uint32_t t0 = texel8(a.screen, W, H, x, y), t1 = texel8(a.screen, W, H, x, y + 1), t2 = texel8(a.screen, W, H, x, y - 1),
                   t3 = texel8(a.screen, W, H, x + 1, y), t4 = texel8(a.screen, W, H, x - 1, y);
    const float* __restrict__ tm = a.tables;
    const float* __restrict__ itm = a.tables + kTaaTable;
    f3 minColor, maxColor, currentColor;
    {
        uint32_t lo[3], hi[3];
        #pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int sh = 8 * c;
            const uint32_t k0 = (t0 >> sh) & 0xffu, k1 = (t1 >> sh) & 0xffu, k2 = (t2 >> sh) & 0xffu, k3 = (t3 >> sh) & 0xffu, k4 = (t4 >> sh) & 0xffu;
            lo[c] = min8(k0, min8(k1, min8(k2, min8(k3, k4))));
            hi[c] = max8(k0, max8(k1, max8(k2, max8(k3, k4))));
        }
        minColor = mk3(__ldg(tm + lo[0]), __ldg(tm + lo[1]), __ldg(tm + lo[2]));
        maxColor = mk3(__ldg(tm + hi[0]), __ldg(tm + hi[1]), __ldg(tm + hi[2]));
        // currentColor = inverseTonemap(tonemap(texel))
        currentColor = mk3(__ldg(itm + (t0 & 0xffu)), __ldg(itm + ((t0 >> 8) & 0xffu)), __ldg(itm + ((t0 >> 16) & 0xffu)));
    }
    prevColor = min3(max3(prevColor, minColor), maxColor);                 // glm::clamp(x, lo, hi) = min(max(x, lo), hi)
    prevColor = inverseTonemap(prevColor);
    const f3 outc = mix(currentColor, prevColor, 0.9f);
    const uint2 hv = packHalf4(outc.x, outc.y, outc.z, 0.0625f);
    a.colorOut[p] = hv;
    if (a.historyOut != nullptr) {
        // previousFrameBuffer.CopyFrom(finalScreenBuffer, ...) (BE:2347-2353): RGBA16F -> RGB8
        const float r = h2f((unsigned short)(hv.x & 0xffffu)), g = h2f((unsigned short)(hv.x >> 16)), b = h2f((unsigned short)(hv.y & 0xffffu));
        const uint32_t ri = (uint32_t)rintf(fmul(saturate(r), 255.0f)), gi = (uint32_t)rintf(fmul(saturate(g), 255.0f)), bi = (uint32_t)rintf(fmul(saturate(b), 255.0f));
        a.historyOut[p] = ri | (gi << 8) | (bi << 16) | 0xff000000u;
    }
}

}  // namespace

bool preloadTaaKernels() {
    cudaFuncAttributes a; bool ok = true;
    ok = ok && cudaFuncGetAttributes(&a, taaKernel) == cudaSuccess;
    ok = ok && cudaFuncGetAttributes(&a, taaTableKernel) == cudaSuccess;
    return ok;
}

int launchTaa(cudaStream_t stream, const BandDims& bd, const SgiTaaParams& tp, float* tables /* kTaaTableFloats floats, library-owned */) {
    taaTableKernel<<<1, kTaaTable, 0, stream>>>(tables);
    TaaArgs a;
    a.tables = tables;
    a.W = bd.W; a.H = bd.rows;
    a.screen = (const uint32_t*)tp.screen; a.prevScreen = (const uint32_t*)tp.prevScreen; a.velocity = (const uint32_t*)tp.velocity;
    a.colorOut = (uint2*)tp.colorOut; a.historyOut = (uint32_t*)tp.historyOut;
    dim3 block(32, 8), grid((bd.W + 31) / 32, (bd.rows + 7) / 8);
    taaKernel<<<grid, block, 0, stream>>>(a);
    return 2;
}

}  // namespace sgi
