// k_ssao.cu — screen-space ambient occlusion, the two passes that produce the `ssao` plane the gather and the direct-lighting
// pass read (first half).
//
// Replaces ssao.fs:28-88 (calculateAmbientOcclusion: 4 rotated offset vectors around the pixel in the structure buffer) as driven by
// RenderSsaoOcclude_ (StratusRendererBackend.cpp:1160-1187), and ssao_blur.fs:18-65 (blurAmbientOcclusion: depth-aware average of
// 4 neighbours) as driven by RenderSsaoBlur_ (:1189-1204).  Textures as the host creates them: the structure buffer is a
// RGBA16F rectangle texture, LINEAR, CLAMP_TO_EDGE (:406-408; x, y = dFdx, dFdy of the camera-space depth, z + w = that depth in
// two halves, pbr_geometry_pass.fs:61-75), the rotation table a 4x4 RGB16F texture, NEAREST, REPEAT (:759-779), both occlusion
// targets R16F, NEAREST (:491-503).
//
// The whole of both shaders is a dozen dependent operations per sample, so everything runs in the exact tier (csrc/sgi_math.cuh:
// IEEE operations in the shader's / glm's evaluation order) and the output equals the compiled shader text bit for bit; the
// LINEAR fetch is the filter model of the denoiser's reprojected reads (8-bit sub-texel weights).
//
// Parity: bit-identical on a B200 to the reference's ssao.fs / ssao_blur.fs compiled by oracle/make_ref_glsl.py (tests/test_zz_ssao.py,
// profiles/r02a), and on the host (tests/test_ssao_kernel_on_host.py).  Whole-frame handles only (no band halos yet).
#include "sgi_internal.h"
#include "sgi_math.cuh"

namespace sgi {
namespace {

// texture(sampler2DRect, pixelCoords), LINEAR + CLAMP_TO_EDGE, on a half4 plane
SGI_DEV f4 rectLinearHalf4(const uint2* __restrict__ plane, int W, int H, float px, float py) {
    const float fu = fsub(px, 0.5f), fv = fsub(py, 0.5f);
    const float i0f = floorf(fu), j0f = floorf(fv);
    const float a = fmul(floorf(fadd(fmul(fsub(fu, i0f), 256.0f), 0.5f)), 1.0f / 256.0f);
    const float b = fmul(floorf(fadd(fmul(fsub(fv, j0f), 256.0f), 0.5f)), 1.0f / 256.0f);
    const int i0 = iclamp((int)i0f, 0, W - 1), i1 = iclamp((int)i0f + 1, 0, W - 1), j0 = iclamp((int)j0f, 0, H - 1), j1 = iclamp((int)j0f + 1, 0, H - 1);
    const f4 t00 = loadHalf4(plane, (size_t)j0 * W + i0), t10 = loadHalf4(plane, (size_t)j0 * W + i1);
    const f4 t01 = loadHalf4(plane, (size_t)j1 * W + i0), t11 = loadHalf4(plane, (size_t)j1 * W + i1);
    const float na = fsub(1.0f, a), nb = fsub(1.0f, b);
    const float w00 = fmul(na, nb), w10 = fmul(a, nb), w01 = fmul(na, b), w11 = fmul(a, b);
    f4 r;
    r.x = fadd(fadd(fadd(fmul(w00, t00.x), fmul(w10, t10.x)), fmul(w01, t01.x)), fmul(w11, t11.x));
    r.y = fadd(fadd(fadd(fmul(w00, t00.y), fmul(w10, t10.y)), fmul(w01, t01.y)), fmul(w11, t11.y));
    r.z = fadd(fadd(fadd(fmul(w00, t00.z), fmul(w10, t10.z)), fmul(w01, t01.z)), fmul(w11, t11.z));
    r.w = fadd(fadd(fadd(fmul(w00, t00.w), fmul(w10, t10.w)), fmul(w01, t01.w)), fmul(w11, t11.w));
    return r;
}

struct SsaoArgs {
    int W, H;
    float aspectRatio, projPlaneZDist, intensity;
    float vectorScale;                 // 2 aspectRatio / (W projPlaneZDist) (ssao.fs:34): uniform, computed once on the host with the same IEEE operations
    const uint2* structure;            // half4 [H][W]
    const uint2* rotation;             // half4 [4][4]
    const unsigned short* occlusion;   // half [H][W] (blur pass input)
    unsigned short* out;               // half [H][W]
};

// ---- ssao.fs ----
__global__ void __launch_bounds__(256)
ssaoOccludeKernel(const __grid_constant__ SsaoArgs a) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= a.W || y >= a.H) return;
    const float w = (float)a.W, h = (float)a.H;
    // gl: fsTexCoords * vec2(windowWidth, windowHeight), fsTexCoords = (x + 0.5) / W (ssao.vs)
    const float pcx = fmul(fdiv(fadd((float)x, 0.5f), w), w), pcy = fmul(fdiv(fadd((float)y, 0.5f), h), h);
    const float tau = fdiv(1.0f, 32.0f), r = 0.4f;          // (dmax = 2.0: the division by it below is a multiplication by 0.5, the same rounding)
    const float vectorScale = a.vectorScale;
    const float dx[4] = {fmul(0.25f, r), 0.0f, fmul(-0.75f, r), 0.0f};
    const float dy[4] = {0.0f, fmul(0.5f, r), 0.0f, -r};

    const f4 structure = rectLinearHalf4(a.structure, a.W, a.H, pcx, pcy);
    const float z0 = fadd(structure.z, structure.w);
    float scale = fmul(z0, vectorScale);
    const f3 normal = normalize(mk3(structure.x, structure.y, -scale));
    scale = fdiv(1.0f, scale);

    // texture(rotationLookup, 0.25 * pixelCoords).xy: NEAREST, REPEAT on a 4x4 table
    const float ru = fmul(0.25f, pcx), rv = fmul(0.25f, pcy);
    int ri = (int)floorf(fmul(ru, 4.0f)) % 4, rj = (int)floorf(fmul(rv, 4.0f)) % 4;
    if (ri < 0) ri += 4;
    if (rj < 0) rj += 4;
    const f4 rot4 = loadHalf4(a.rotation, (size_t)rj * 4 + ri);
    const float rotX = rot4.x, rotY = rot4.y;
    float occlusion = 0.0f, weight = 0.0f;
    #pragma unroll
    for (int i = 0; i < 4; ++i) {
        f3 v;
        v.x = fsub(fmul(rotX, dx[i]), fmul(rotY, dy[i]));
        v.y = fadd(fmul(rotY, dx[i]), fmul(rotX, dy[i]));
        const f4 d = rectLinearHalf4(a.structure, a.W, a.H, fadd(pcx, fmul(v.x, scale)), fadd(pcy, fmul(v.y, scale)));
        const float z = fadd(d.z, d.w);
        v.z = fsub(z, z0);
        const float c = vmax(fsub(dot(normal, normalize(v)), tau), 0.0f);
        const float fv = fsub(1.0f, fsqrt(fsub(1.0f, fmul(c, c))));
        const float wv = saturate(fsub(1.0f, fmul(dot(normal, v), 0.5f)));
        occlusion = fadd(occlusion, fmul(wv, fv));
        weight = fadd(weight, wv);
    }
    a.out[(size_t)y * a.W + x] = f2h(fsub(1.0f, fdiv(fmul(a.intensity, occlusion), vmax(weight, 0.0001f))));
}

// ---- ssao_blur.fs ----
__global__ void __launch_bounds__(256)
ssaoBlurKernel(const __grid_constant__ SsaoArgs a) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= a.W || y >= a.H) return;
    const float w = (float)a.W, h = (float)a.H;
    const float pcx = fmul(fdiv(fadd((float)x, 0.5f), w), w), pcy = fmul(fdiv(fadd((float)y, 0.5f), h), h);
    const float deltaDepth = fdiv(1.0f, 128.0f), deltaP = 1.5f;
    const f4 structure = rectLinearHalf4(a.structure, a.W, a.H, pcx, pcy);
    const float range = fmul(deltaP, fadd(vmax(fabsf(structure.x), fabsf(structure.y)), deltaDepth));
    const float z0 = fadd(structure.z, structure.w);
    float occX = 0.0f, occY = 0.0f, occZ = 0.0f;
    #pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float oy = fsub(fmul((float)j, 2.0f), 0.5f);
        #pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float ox = fsub(fmul((float)i, 2.0f), 0.5f);
            const float sx = fadd(pcx, ox), sy = fadd(pcy, oy);
            // occlusionBuffer: R16F rectangle texture, NEAREST, CLAMP_TO_EDGE
            const int ti = iclamp((int)floorf(sx), 0, a.W - 1), tj = iclamp((int)floorf(sy), 0, a.H - 1);
            const float sample = h2f(__ldg(a.occlusion + (size_t)tj * a.W + ti));
            occZ = fadd(occZ, sample);
            const f4 d = rectLinearHalf4(a.structure, a.W, a.H, sx, sy);
            const float z = fadd(d.z, d.w);
            if (fabsf(fsub(z, z0)) < range) { occX = fadd(occX, sample); occY = fadd(occY, 1.0f); }
        }
    }
    a.out[(size_t)y * a.W + x] = f2h(occY > 0.0f ? fdiv(occX, occY) : fmul(0.25f, occZ));
}

}  // namespace

bool preloadSsaoKernels() {
    cudaFuncAttributes a; bool ok = true;
    ok = ok && cudaFuncGetAttributes(&a, ssaoOccludeKernel) == cudaSuccess;
    ok = ok && cudaFuncGetAttributes(&a, ssaoBlurKernel) == cudaSuccess;
    return ok;
}

int launchSsao(cudaStream_t stream, const BandDims& bd, const SgiSsaoParams& sp) {
    SsaoArgs a;
    a.W = bd.W; a.H = bd.rows;
    a.aspectRatio = sp.aspectRatio; a.projPlaneZDist = sp.projPlaneZDist; a.intensity = sp.intensity;
    {
        // (volatile: every intermediate is rounded to binary32, whatever the host compiler would like to keep in a wider register)
        volatile float num = 2.0f * sp.aspectRatio, den = (float)bd.W * sp.projPlaneZDist;
        volatile float q = num / den;
        a.vectorScale = q;
    }
    a.structure = (const uint2*)sp.structureBuffer; a.rotation = (const uint2*)sp.rotationLookup;
    dim3 block(32, 8), grid((bd.W + 31) / 32, (bd.rows + 7) / 8);
    a.occlusion = nullptr; a.out = (unsigned short*)sp.occlusionScratch;
    ssaoOccludeKernel<<<grid, block, 0, stream>>>(a);
    a.occlusion = (const unsigned short*)sp.occlusionScratch; a.out = (unsigned short*)sp.ssaoOut;
    ssaoBlurKernel<<<grid, block, 0, stream>>>(a);
    return 2;
}

}  // namespace sgi
