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

SGI_DEV f3 rgb8(const uint32_t* __restrict__ plane, int W, int H, int i, int j) {
    const uint32_t t = __ldg(plane + (size_t)iclamp(j, 0, H - 1) * W + iclamp(i, 0, W - 1));
    return mk3(fdiv((float)(t & 0xffu), 255.0f), fdiv((float)((t >> 8) & 0xffu), 255.0f), fdiv((float)((t >> 16) & 0xffu), 255.0f));
}
SGI_DEV f3 tonemap(f3 c) { return c / (c + 1.0f); }                      // aa_common.glsl:17-20: color / (color + vec3(1.0))
SGI_DEV f3 inverseTonemap(f3 c) { return c / mk3(fsub(1.0f, c.x), fsub(1.0f, c.y), fsub(1.0f, c.z)); }      // aa_common.glsl:22-24: color / (vec3(1.0) - color)
SGI_DEV f3 min3(f3 a, f3 b) { return mk3(b.x < a.x ? b.x : a.x, b.y < a.y ? b.y : a.y, b.z < a.z ? b.z : a.z); }       // glm::min(x, y) = y < x ? y : x
SGI_DEV f3 max3(f3 a, f3 b) { return mk3(a.x < b.x ? b.x : a.x, a.y < b.y ? b.y : a.y, a.z < b.z ? b.z : a.z); }       // glm::max(x, y) = x < y ? y : x

struct TaaArgs { int W, H; const uint32_t* screen; const uint32_t* prevScreen; const uint32_t* velocity; uint2* colorOut; uint32_t* historyOut; };

__global__ void __launch_bounds__(256)
taaKernel(const __grid_constant__ TaaArgs a) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= a.W || y >= a.H) return;
    const int W = a.W, H = a.H;
    const size_t p = (size_t)y * W + x;
    const float u = fdiv(fadd((float)x, 0.5f), (float)W), v = fdiv(fadd((float)y, 0.5f), (float)H);
    const uint32_t vel = __ldg(a.velocity + p);
    const float pu = fsub(u, h2f((unsigned short)(vel & 0xffffu))), pv = fsub(v, h2f((unsigned short)(vel >> 16)));

    f3 currentColor = rgb8(a.screen, W, H, x, y);
    // texture(prevScreen, prevTexCoords): LINEAR, CLAMP_TO_EDGE
    f3 prevColor;
    {
        const float fu = fsub(fmul(pu, (float)W), 0.5f), fv = fsub(fmul(pv, (float)H), 0.5f);
        const float i0f = floorf(fu), j0f = floorf(fv);
        const float wa = fmul(floorf(fadd(fmul(fsub(fu, i0f), 256.0f), 0.5f)), 1.0f / 256.0f);
        const float wb = fmul(floorf(fadd(fmul(fsub(fv, j0f), 256.0f), 0.5f)), 1.0f / 256.0f);
        const int i0 = (int)i0f, j0 = (int)j0f;
        const f3 t00 = rgb8(a.prevScreen, W, H, i0, j0), t10 = rgb8(a.prevScreen, W, H, i0 + 1, j0);
        const f3 t01 = rgb8(a.prevScreen, W, H, i0, j0 + 1), t11 = rgb8(a.prevScreen, W, H, i0 + 1, j0 + 1);
        const float na = fsub(1.0f, wa), nb = fsub(1.0f, wb);
        const float w00 = fmul(na, nb), w10 = fmul(wa, nb), w01 = fmul(na, wb), w11 = fmul(wa, wb);
        prevColor.x = fadd(fadd(fadd(fmul(w00, t00.x), fmul(w10, t10.x)), fmul(w01, t01.x)), fmul(w11, t11.x));
        prevColor.y = fadd(fadd(fadd(fmul(w00, t00.y), fmul(w10, t10.y)), fmul(w01, t01.y)), fmul(w11, t11.y));
        prevColor.z = fadd(fadd(fadd(fmul(w00, t00.z), fmul(w10, t10.z)), fmul(w01, t01.z)), fmul(w11, t11.z));
    }
    prevColor = tonemap(prevColor);
    // textureOffset(screen, fsTexCoords, ivec2(..)): whole-texel offsets from a texel centre are single texels (CLAMP_TO_EDGE at the border)
    const f3 c1 = rgb8(a.screen, W, H, x, y + 1), c2 = rgb8(a.screen, W, H, x, y - 1), c3 = rgb8(a.screen, W, H, x + 1, y), c4 = rgb8(a.screen, W, H, x - 1, y);
    const f3 minColor = tonemap(min3(currentColor, min3(c1, min3(c2, min3(c3, c4)))));
    const f3 maxColor = tonemap(max3(currentColor, max3(c1, max3(c2, max3(c3, c4)))));
    currentColor = tonemap(currentColor);
    prevColor = min3(max3(prevColor, minColor), maxColor);                 // glm::clamp(x, lo, hi) = min(max(x, lo), hi)
    currentColor = inverseTonemap(currentColor);
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
    cudaFuncAttributes a;
    return cudaFuncGetAttributes(&a, taaKernel) == cudaSuccess;
}

int launchTaa(cudaStream_t stream, const BandDims& bd, const SgiTaaParams& tp) {
    TaaArgs a;
    a.W = bd.W; a.H = bd.rows;
    a.screen = (const uint32_t*)tp.screen; a.prevScreen = (const uint32_t*)tp.prevScreen; a.velocity = (const uint32_t*)tp.velocity;
    a.colorOut = (uint2*)tp.colorOut; a.historyOut = (uint32_t*)tp.historyOut;
    dim3 block(32, 8), grid((bd.W + 31) / 32, (bd.rows + 7) / 8);
    taaKernel<<<grid, block, 0, stream>>>(a);
    return 1;
}

}  // namespace sgi
