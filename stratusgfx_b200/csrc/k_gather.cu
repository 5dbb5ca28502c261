// k_gather.cu — per-pixel stochastic VPL gather over the deferred G-buffer (a5).
//
// Replaces vpl_pbr_gi.fs:81-207 and the functions it calls:
//   random/hash/floatConstruct           common.glsl:183-214
//   worldPositionFromDepth               common.glsl:143-155
//   calculateShadowValue1Sample          pbr.glsl:116-136
//   vplDiffuseAttenuation                pbr.glsl:260-267
//   calculateVirtualPointLighting2       pbr2.glsl:613-657
//   calculateLighting_Lambert_SeparateDiffuseSpecular  pbr2.glsl:541-574
//   BRDF_Lambert_SeparateDiffuseSpecular pbr2.glsl:382-452  (BRDF_Base :146-168, BRDF_DotProducts :170-188,
//   singleScatteringBRDF_Specular :102-115, NDF_GGX :51-55, Visibility_G_SmithGGX :59-64, Fr_Fresnel_Schlick :73-76)
//   applyFog                             fog.glsl:37-41
//
// One thread per pixel.  The stochastic mode (the reference's live shader) runs as a PERSISTENT kernel:
// one or two CTAs per SM stage the frame's compacted VPL table (48 B per VPL: position+radius,
// colour+farPlane, specular position+atlas entry) in shared memory once, then their warps stride
// over 32-pixel row segments.  A pixel's draws are a fixed sequence (draw k hashes (x, y, time+10000k)),
// so the kernel first runs the accept/re-draw decisions for the whole sequence (phase 1: one hash,
// one 16 B shared-memory read and the side/range test per draw, every lane busy until its own
// sample budget is met) and then shades the accepted VPLs in order (phase 2).  The shader's single
// loop does the same work in the same order; splitting it keeps the lanes of a warp together in the
// expensive part and replaces scattered L1 reads (32 tag look-ups per request) by shared-memory reads.
// Two arithmetic tiers:
//
//   DECISIONS (which VPL is drawn, re-draw or keep, sample budget, cube texel, shadowed or lit) are
//   evaluated so that they come out exactly as the IEEE / glm-order evaluation of the shader text
//   would: the light index is pure integer hashing of (x+0.5, y+0.5, time); every float comparison is
//   first tried with a cheap FMA estimate and a conservative margin (1e-5 relative, >50x the rounding
//   error of either evaluation) and falls back to the exact operation chain only when the estimate
//   is inside that margin.  Both routes give the same answer by construction.
//
//   RADIANCE (attenuation, GGX, Fresnel, Lambert, accumulation) uses FMA, rsqrt/rcp/ex2 approximations
//   and (1-u)^5 by multiplication.  It differs from the CPU transcription by a few float ulps, far
//   inside the 1e-3 relative / 60 dB PSNR bar once stored as binary16.
//
// The kernel also writes the frame's "guide" plane (decoded normal + depth, float4) that the
// denoiser passes read instead of re-decoding 25 normals per pixel per pass, and (when asked) the
// output of denoise pass 0, which is a per-pixel function of the reservoir it has just computed.
#include <cstdlib>
#include "sgi_internal.h"
#include "sgi_math.cuh"

namespace sgi {
namespace {

constexpr float PI_F = 3.14159265359f;
constexpr float PREVENT_DIV_BY_ZERO = 0.00001f;
constexpr float HALF_FLOAT_MAX = 65504.0f;
constexpr float kMargin = 1e-5f;

// radiance tier and margin estimates: one MUFU each (the non-ftz forms nvcc emits for the fast-division, reciprocal-square-root
// and exponential intrinsics carry a denormal-range rescue of three to four more instructions per call; no operand of this
// file gets near that range)
SGI_DEV float rsqrtFast(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
SGI_DEV float ex2Fast(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
SGI_DEV float sqrtFast(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
SGI_DEV void prefetchL2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
constexpr float kLog2e = 1.4426950408889634f;
SGI_DEV float expFast(float x) { return ex2Fast(x * kLog2e); }
SGI_DEV float dotFast(f3 a, f3 b) { return fma_(a.z, b.z, fma_(a.y, b.y, a.x * b.x)); }
SGI_DEV float satFast(float x) { return __saturatef(x); }

// ---- exact fallbacks (rarely executed) ----------------------------------------------------------------------
__device__ __noinline__ bool exactRedraw(f3 normal, f3 lightMinusFrag, float lightRadius) {
    const float distance = length(lightMinusFrag);
    const float sideCheck = dot(normal, normalize(lightMinusFrag));
    return sideCheck < 0.0f || distance > lightRadius;
}
__device__ __noinline__ bool exactShadowed(f3 fragDir, float lightNormalDot, float texelDepth, float farPlane) {
    const float currentDepth = length(fragDir);
    const float bias = fmul(currentDepth, vmax(0.1f, fmul(saturate(lightNormalDot), 0.03f)));
    return fsub(currentDepth, bias) > fmul(texelDepth, farPlane);
}

struct PixelConsts {   // per-pixel values that do not depend on the light
    f3 fragPos, normal, viewDir, diffuseOverPi, f0;
    float NdotV, ggxvTerm, r2, fogAo;
};

// radiance tier: ambientOcclusion * calculateVirtualPointLighting2(...) with the constant arguments of the
// call site (vpl_pbr_gi.fs:195-197: shadowFactor 0, ambientIntensity 0, fogIntensity 0)
SGI_DEV f3 shadeFast(const PixelConsts& px, f3 specularLightPos, f3 diffuseLightPos, f3 lightColor, float lightRadius, float invRadius) {
    const f3 Ls = mk3(specularLightPos.x - px.fragPos.x, specularLightPos.y - px.fragPos.y, specularLightPos.z - px.fragPos.z);
    const f3 Ld = mk3(diffuseLightPos.x - px.fragPos.x, diffuseLightPos.y - px.fragPos.y, diffuseLightPos.z - px.fragPos.z);
    const float ds2 = dotFast(Ls, Ls), dd2 = dotFast(Ld, Ld);
    const float is = rsqrtFast(ds2), id = rsqrtFast(dd2);
    // vplDiffuseAttenuation for both directions (pbr.glsl:260-267): 1 / (mix(10,1,exp(-3*ratio))*radius + dist^2)
    const float rs = fminf(ds2 * is * invRadius, 1.0f), rd = fminf(dd2 * id * invRadius, 1.0f);
    const float es = ex2Fast(rs * (-3.0f * kLog2e)), ed = ex2Fast(rd * (-3.0f * kLog2e));
    const float attS = rcpFast(fma_(fma_(es, -9.0f, 10.0f), lightRadius, ds2));
    const float attD = rcpFast(fma_(fma_(ed, -9.0f, 10.0f), lightRadius, dd2));
    // specular lobe
    const f3 L = mk3(Ls.x * is, Ls.y * is, Ls.z * is);
    f3 H = mk3(px.viewDir.x + L.x, px.viewDir.y + L.y, px.viewDir.z + L.z);
    const float ih = rsqrtFast(dotFast(H, H));
    const float NdotL = satFast(dotFast(px.normal, L));
    const float NdotH = satFast(dotFast(px.normal, H) * ih);
    const float LdotH = satFast(dotFast(L, H) * ih);
    const float f = fma_(fma_(NdotH, px.r2, -NdotH), NdotH, 1.0f);
    const float ndfDen = fmaxf(PI_F * f * f, PREVENT_DIV_BY_ZERO);
    const float ggxl = px.NdotV * sqrtFast(fma_(fma_(-px.r2, NdotL, NdotL), NdotL, px.r2));
    const float visDen = fmaxf(fma_(NdotL, px.ggxvTerm, ggxl), PREVENT_DIV_BY_ZERO);
    const float ndfVis = (px.r2 * 0.5f) * rcpFast(ndfDen * visDen);
    const float x1 = 1.0f - LdotH, x2 = x1 * x1, fr = x2 * x2 * x1;
    const float sw = ndfVis * NdotL * attS, omf = 1.0f - fr;
    // diffuse lobe
    const float NdotL2 = satFast(dotFast(px.normal, Ld) * id);
    const float dw = NdotL2 * attD;
    f3 brdf;
    brdf.x = fma_(sw, fma_(px.f0.x, omf, fr), dw * px.diffuseOverPi.x);
    brdf.y = fma_(sw, fma_(px.f0.y, omf, fr), dw * px.diffuseOverPi.y);
    brdf.z = fma_(sw, fma_(px.f0.z, omf, fr), dw * px.diffuseOverPi.z);
    return mk3(brdf.x * lightColor.x * px.fogAo, brdf.y * lightColor.y * px.fogAo, brdf.z * lightColor.z * px.fogAo);
}


// decision: `if (dot(normal, normalize(l)) < 0 || length(l) > radius)` (vpl_pbr_gi.fs:164-166), exact by the two-tier rule
SGI_DEV bool behindOrOutOfRange(f3 normal, f3 lightMinusFrag, float d2, float lightRadius) {
    const float s = dotFast(normal, lightMinusFrag);      // same sign as the normalised dot product
    const float r2 = lightRadius * lightRadius;
    const bool sideSure = s * s > (kMargin * kMargin) * d2;
    const bool farSure = d2 > r2 * (1.0f + 4.0f * kMargin), nearSure = d2 < r2 * (1.0f - 4.0f * kMargin);
    if (sideSure && (farSure || nearSure)) return (s < 0.0f) || farSure;
    return exactRedraw(normal, lightMinusFrag, lightRadius);
}

// what does not change from light to light in the shadow lookup
struct CubeConsts {
    int res, texelsPerLayer;
    float halfRes, epsK;       // res / 2, res * 2^-22 (cubeTexelDecided)
};
SGI_DEV CubeConsts cubeConsts(int cubeRes) {
    CubeConsts c;
    c.res = cubeRes; c.texelsPerLayer = 6 * cubeRes * cubeRes;
    c.halfRes = 0.5f * (float)cubeRes; c.epsK = (float)cubeRes * (1.0f / 4194304.0f);
    return c;
}

// the shadow-map texel a fragment sees of one light (calculateShadowValue1Sample, pbr.glsl:116-136: the texel choice is a decision)
SGI_DEV uint32_t loadShadowTexel(f3 fragPos, float4 posRadius, int entryBits /* atlas index << 16 | layer */, const CubeConsts& cc, const AtlasTable& atlases) {
    const f3 fragDir = fragPos - mk3(posRadius.x, posRadius.y, posRadius.z);
    const uint32_t offset = (uint32_t)(entryBits & 0xffff) * (uint32_t)cc.texelsPerLayer + (uint32_t)cubeTexelDecided(fragDir, cc.res, cc.halfRes, cc.epsK);
    return __ldg(atlases.shadow[entryBits >> 16] + offset);
}

// one accepted light: shadow verdict on the texel loaded above, occlusion floor, re-weighting and shading (vpl_pbr_gi.fs:173-198)
SGI_DEV void shadeAcceptedLight(const PixelConsts& px, f3 lightPosition, f3 lightMinusFrag, float d2, float lightRadius, float farPlane,
                                f3 lightColor, f3 specularPosition, uint32_t texel, bool shadowsOn, float minGi, f3& vplColor, float& validSamples) {
    const float invRadius = rcpFast(lightRadius);
    const float distance = d2 * rsqrtFast(d2);
    const float distanceRatio = satFast(2.0f * distance * invRadius);
    float shadowFactor = 0.0f;
    if (shadowsOn) {
        // minBias 0.1; the lit / shadowed verdict is a decision
        const float texelDepthEst = (float)texel * (1.0f / 65535.0f);      // within an ulp of the exact unorm16 (margin below: 8e-5)
        const float lightNormalDot = dotFast(lightMinusFrag, px.normal);
        const float biasFactor = fmaxf(0.1f, satFast(lightNormalDot) * 0.03f);
        const float lhs = distance * (1.0f - biasFactor), rhs = texelDepthEst * farPlane;
        const float gap = lhs - rhs;
        // lightNormalDot only matters through saturate(.)*0.03 vs 0.1, i.e. when it exceeds 3.33: never borderline for the verdict's margin
        bool shadowed;
        if (fabsf(gap) > 8.0f * kMargin * fmaxf(lhs, rhs)) shadowed = gap > 0.0f;
        else shadowed = exactShadowed(px.fragPos - lightPosition, dot(lightPosition - px.fragPos, px.normal), unorm16(texel), farPlane);
        shadowFactor = shadowed ? 1.0f : 0.0f;
    }
    shadowFactor = fminf(shadowFactor, fma_(distanceRatio, 1.0f - minGi, minGi));
    float reweightingFactor = 1.0f;
    if (shadowFactor > 0.0f) reweightingFactor = fma_(1.0f - distanceRatio, minGi, distanceRatio);
    validSamples += reweightingFactor;
    const f3 tmpColor = shadeFast(px, specularPosition, lightPosition, lightColor, lightRadius, invRadius);
    const float lit = 1.0f - shadowFactor;
    vplColor = mk3(fma_(lit, tmpColor.x, vplColor.x), fma_(lit, tmpColor.y, vplColor.y), fma_(lit, tmpColor.z, vplColor.z));
}

// Denoise pass 0 of one channel on the binary16 reservoir just stored: f2h(a / denom), a and denom binary16 values, denom >= 1.
// The quotient is needed to 11 bits only, so it is estimated as a * rcp(denom) — within 4 ulp (binary32) of the IEEE quotient — and
// only an estimate within 8 ulp of a binary16 rounding boundary (a midpoint of two neighbours: low 13 mantissa bits 0x1000), or
// below binary16's normal range (boundaries are spaced differently there), takes the exact division.  a = 0 gives 0 on both routes.
SGI_DEV unsigned short mergedChannel(float a, float denom, float invDenom) {
    const float q = a * invDenom;
    const uint32_t b = __float_as_uint(q);
    const bool sure = (((b + 8u - 0x1000u) & 0x1fffu) > 16u && q >= 6.2e-5f) || a == 0.0f;
    return f2h(sure ? q : fdiv(a, denom));
}

// tile-list mode: the <= 12 VPLs of a tile as the kernel stages them in shared memory
struct StagedLight { float4 posRadius, colorFar, specEntry; };       // specEntry.w = bits of (atlas index << 16 | layer)
constexpr int kTableBytesPerVpl = 48;                                // three float4 per staged VPL
constexpr int kTableBytesTwoCtas = 110 * 1024;                       // two CTAs per SM while 2 x (table + 1 KB reserve) <= 227 KB
constexpr int kTileCtaX = 8, kTileCtaY = 2;                          // tiles per CTA in tile-list mode: 40 x 10 pixels, one thread per pixel

// ---- tile-list mode (SGI_GATHER_TILE_LIST): one thread per pixel, one CTA per 8 x 2 tiles ----------------------------------------
__global__ void __launch_bounds__(kTileCtaX * kTileCtaY * 25, 2)
gatherTileListKernel(const __grid_constant__ FrameConsts fc, BandDims bd, const float* __restrict__ gDepth, const uint32_t* __restrict__ gNormal,
             const uint32_t* __restrict__ gAlbedo, const uint16_t* __restrict__ gRefl, const uint32_t* __restrict__ gRma,
             const unsigned short* __restrict__ ssao, const SgiVplData* __restrict__ lightData, const int32_t* __restrict__ numVisible,
             const SgiAtlasEntry* __restrict__ shadowIndices, const __grid_constant__ AtlasTable atlases,
             uint2* __restrict__ colorOut, uint2* __restrict__ reservoirOut, float4* __restrict__ guideOut, uint2* __restrict__ mergedOut,
             const __grid_constant__ BandLinks links, const SgiVplStage2PerTileOutputs* __restrict__ tileLists, const unsigned short* __restrict__ tilePos) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int ly = blockIdx.y * blockDim.y + threadIdx.y;
    // ---- tile-list mode: stage this CTA's 8x2 tiles' VPL lists through shared memory ----
    __shared__ StagedLight sLight[kTileCtaX * kTileCtaY * SGI_MAX_VPLS_PER_TILE];
    __shared__ int sCount[kTileCtaX * kTileCtaY];
    {
        const int tilesX = bd.W / SGI_TILE_DIVISOR, tilesY = bd.rows / SGI_TILE_DIVISOR;
        const int t = threadIdx.y * blockDim.x + threadIdx.x;
        if (t < kTileCtaX * kTileCtaY * SGI_MAX_VPLS_PER_TILE) {
            const int lt = t / SGI_MAX_VPLS_PER_TILE, l = t % SGI_MAX_VPLS_PER_TILE;
            const int tx = blockIdx.x * kTileCtaX + lt % kTileCtaX, ty = blockIdx.y * kTileCtaY + lt / kTileCtaX;
            int count = 0;
            if (tx < tilesX && ty < tilesY) {
                const int tile = ty * tilesX + tx;
                count = min(__ldg(&tileLists[tile].numVisible), SGI_MAX_VPLS_PER_TILE);
                if (l < count) {
                    const int k = min((int)__ldg(tilePos + (size_t)tile * SGI_MAX_VPLS_PER_TILE + l), SGI_MAX_TOTAL_VPLS_PER_FRAME - 1);   // position in the compacted (coloured) table
                    const float4* L = reinterpret_cast<const float4*>(lightData + k);
                    const float4 lp = __ldg(L), lc = __ldg(L + 1), sp = __ldg(L + 2), rfi = __ldg(L + 3);
                    const SgiAtlasEntry e = shadowIndices[k];
                    StagedLight sl;
                    sl.posRadius = make_float4(lp.x, lp.y, lp.z, rfi.x);
                    sl.colorFar = make_float4(lc.x, lc.y, lc.z, rfi.y);
                    sl.specEntry = make_float4(sp.x, sp.y, sp.z, __int_as_float((e.index << 16) | (e.layer & 0xffff)));
                    sLight[t] = sl;
                }
            }
            if (l == 0) sCount[lt] = count;
        }
        __syncthreads();
    }
    const int y = ly + bd.yOffset;
    // row bands: the CTAs that will store pass-0 rows into a neighbour's halo first wait until that neighbour has finished
    // the previous frame (its final pass read those halo rows)
    bandWaitCta(links, links.up.res != nullptr && y >= links.up.firstRow && y < links.up.lastRow,
                links.down.res != nullptr && y >= links.down.firstRow && y < links.down.lastRow);
    if (x >= bd.W || ly >= bd.rows) return;
    const size_t p = (size_t)ly * bd.W + x;
    const CubeConsts cc = cubeConsts(bd.cubeRes);

    // ---- per-pixel set-up: exact (feeds the sample budget, the 500-unit shadow cut-off and every re-draw test) ----
    const float depth = __ldg(gDepth + p);
    PixelConsts px;
    px.normal = decodeNormal(__ldg(gNormal + p));
    guideOut[p] = make_float4(px.normal.x, px.normal.y, px.normal.z, depth);
    if (ly < bd.c0 || ly >= bd.c0 + bd.crows) return;      // halo row of a band: only the guide plane is needed here
    const m4 invPV = loadMat(fc.invPV);
    const f3 viewPosition = mk3(fc.viewPos[0], fc.viewPos[1], fc.viewPos[2]);
    const float pcx = fadd((float)x, 0.5f), pcy = fadd((float)y, 0.5f);          // gl_FragCoord.xy
    const float u = fdiv(pcx, (float)bd.W), v = fdiv(pcy, (float)bd.fullH);       // fsTexCoords
    px.fragPos = worldPositionFromDepth(u, v, depth, invPV);
    const f3 toView = viewPosition - px.fragPos;
    const float distToCamera = length(toView);
    const uint32_t alb = __ldg(gAlbedo + p);
    const f3 baseColor = mk3(vmax(unorm8(alb & 0xffu), 0.01f), vmax(unorm8((alb >> 8) & 0xffu), 0.01f), vmax(unorm8((alb >> 16) & 0xffu), 0.01f));
    const uint32_t rma = __ldg(gRma + p);
    const float roughness = vmax(0.5f, unorm8(rma & 0xffu));
    const float roughnessWeight = fsub(1.0f, roughness);
    const float distRatioToCamera = vmin(fsub(1.0f, fdiv(distToCamera, 1000.0f)), 1.0f);
    const int maxSamplesPerPixel = (int)mix(2.0f, 10.0f, roughnessWeight);
    const int sampleCount = max(1, (int)fmul((float)maxSamplesPerPixel, distRatioToCamera));
    const bool shadowsOn = distToCamera < 500.0f;

    // ---- per-pixel set-up: radiance tier (BRDF_Base pbr2.glsl:146-168 and everything light-independent) ----
    {
        const float metallic = unorm8((rma >> 8) & 0xffu);
        const float ambientOcclusion = vclamp(h2f(__ldg(ssao + p)), 0.35f, 1.0f);
        const float refl = unorm8((uint32_t)__ldg(gRefl + p) & 0xffu);
        float rr = roughness;
        if (fc.usePerceptualRoughness) rr = rr * rr;
        rr = fmaxf(fmaxf(0.0f, fc.minRoughness), rr);
        px.r2 = rr * rr;
        const float iv = rsqrtFast(dotFast(toView, toView));
        px.viewDir = mk3(toView.x * iv, toView.y * iv, toView.z * iv);
        px.NdotV = satFast(dotFast(px.normal, px.viewDir));
        px.ggxvTerm = sqrtFast(fma_(fma_(-px.r2, px.NdotV, px.NdotV), px.NdotV, px.r2));
        const float om = 1.0f - metallic, f0s = 0.16f * refl * refl * om;
        const f3 inv = mk3(rcpFast(baseColor.x + PREVENT_DIV_BY_ZERO), rcpFast(baseColor.y + PREVENT_DIV_BY_ZERO), rcpFast(baseColor.z + PREVENT_DIV_BY_ZERO));
        // diffuseColor / (diffuseDivisor + eps) / PI   and   f0 * specularMultiplier
        px.diffuseOverPi = mk3(om * baseColor.x * inv.x * (1.0f / PI_F), om * baseColor.y * inv.y * (1.0f / PI_F), om * baseColor.z * inv.z * (1.0f / PI_F));
        px.f0 = mk3(fma_(baseColor.x, metallic, f0s) * inv.x, fma_(baseColor.y, metallic, f0s) * inv.y, fma_(baseColor.z, metallic, f0s) * inv.z);
        px.fogAo = expFast(-fc.fogDensity * distToCamera) * ambientOcclusion;
    }

    f3 vplColor = mk3(0.0f);
    float validSamples = 0.0f;
    const int V = __ldg(numVisible);
    const float minGi = fc.minGiOcclusionFactor;

    {
        // Tile-list mode (no live reference counterpart: the tile lookup of vpl_pbr_gi.fs:86-97 is commented out): every VPL of
        // the pixel's 5x5 tile list, in list order, through the same per-light body; a VPL behind the pixel or out of range is
        // skipped (the stochastic shader would re-draw).  Deterministic: no random numbers.
        const int lt = ((int)threadIdx.y / SGI_TILE_DIVISOR) * kTileCtaX + (int)threadIdx.x / SGI_TILE_DIVISOR;
        const int count = sCount[lt];
        for (int l = 0; l < count; ++l) {
            const StagedLight sl = sLight[lt * SGI_MAX_VPLS_PER_TILE + l];
            const f3 lightPosition = mk3(sl.posRadius.x, sl.posRadius.y, sl.posRadius.z);
            const f3 lightMinusFrag = lightPosition - px.fragPos;
            const float lightRadius = sl.posRadius.w;
            const float d2 = dotFast(lightMinusFrag, lightMinusFrag);
            if (behindOrOutOfRange(px.normal, lightMinusFrag, d2, lightRadius)) continue;
            const uint32_t texel = shadowsOn ? loadShadowTexel(px.fragPos, sl.posRadius, __float_as_int(sl.specEntry.w), cc, atlases) : 0u;
            shadeAcceptedLight(px, lightPosition, lightMinusFrag, d2, lightRadius, sl.colorFar.w, mk3(sl.colorFar.x, sl.colorFar.y, sl.colorFar.z),
                               mk3(sl.specEntry.x, sl.specEntry.y, sl.specEntry.z), texel, shadowsOn, minGi, vplColor, validSamples);
        }
    }
    validSamples = fmaxf(validSamples, 1.0f);
    const f3 color = baseColor + PREVENT_DIV_BY_ZERO;
    storeHalf4(colorOut, p, color.x, color.y, color.z, 1.0f);
    const unsigned short hr = f2h(fminf(vplColor.x, HALF_FLOAT_MAX)), hg = f2h(fminf(vplColor.y, HALF_FLOAT_MAX)), hb = f2h(fminf(vplColor.z, HALF_FLOAT_MAX)),
                         ha = f2h(validSamples);
    uint2 packed;
    packed.x = (uint32_t)hr | ((uint32_t)hg << 16);
    packed.y = (uint32_t)hb | ((uint32_t)ha << 16);
    reservoirOut[p] = packed;
    if (mergedOut != nullptr) {
        // Denoise pass 0 on the binary16 reservoir just stored.  mergeCentreTexel (sgi_merge.cuh) reduces to rgb / max(1e-5, a)
        // here: the normal is unit length, |z - z| = 0, and with a >= 1 and 1 <= V <= 4480 the single candidate's acceptance
        // ratio probability / cw is x / x = 1 exactly, which every draw in [0,1) passes.  V <= 0 leaves the reservoir empty.
        const float denom = vmax(PREVENT_DIV_BY_ZERO, h2f(ha)), invDenom = rcpFast(denom);
        uint2 mv = make_uint2(0u, 0x3c000000u);                      // (0, 0, 0, 1)
        if (V > 0) {
            mv.x = (uint32_t)mergedChannel(h2f(hr), denom, invDenom) | ((uint32_t)mergedChannel(h2f(hg), denom, invDenom) << 16);
            mv.y |= (uint32_t)mergedChannel(h2f(hb), denom, invDenom);
        }
        mergedOut[p] = mv;
        pushRes(links, y, x, bd.W, mv);
    }
}


// ---- stochastic mode, persistent two-phase kernel ------------------------------------------------------------------
constexpr int kMaxSamples = 6;                // int(mix(2, 10, 1 - max(roughness, 0.5))) <= 6 (vpl_pbr_gi.fs:138-143)
constexpr int kMaxResamples = 5;              // vpl_pbr_gi.fs:164-171
constexpr int kMaxDraws = kMaxSamples + kMaxResamples;
constexpr int kUnitChunk = 2;                 // 32-pixel row segments a warp takes per fetch from the work counter
static_assert(kMaxDraws <= kGatherMaxDraws, "FrameConsts::hashZ is too short");

SGI_DEV void bandWaitWarp(const BandLinks& L, bool needUp, bool needDown) {
    if (L.waitUp == nullptr && L.waitDown == nullptr) return;          // uniform across the launch
    const bool up = __any_sync(0xffffffffu, needUp && L.waitUp != nullptr);
    const bool down = __any_sync(0xffffffffu, needDown && L.waitDown != nullptr);
    if (up || down) {
        if ((threadIdx.x & 31) == 0) {
            const long long t0 = clock64();
            bool ok = true;
            if (up) ok = spinUntil(L.waitUp, L.waitEpoch, t0) && ok;
            if (down) ok = spinUntil(L.waitDown, L.waitEpoch, t0) && ok;
            if (!ok) atomicAdd(L.timeouts, 1);
            bandSpinAccount(L, t0);
        }
        __syncwarp();
    }
}

// per-column / per-row values of the pixel grid that the shader recomputes for every fragment: built once per handle
// geometry by coordTableKernel with the exact operations of the shader (fsTexCoords = gl_FragCoord.xy / resolution)
struct CoordTables {
    const float* u;            // [W]     (x + 0.5) / W
    const float2* row;         // [rows]  ((y + 0.5) / fullH, bits of hash(floatBitsToUint(y + 0.5))) for frame row y = yOffset + local row
};
__global__ void coordTableKernel(BandDims bd, float* __restrict__ u, float2* __restrict__ row) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < bd.W) u[i] = fdiv(fadd((float)i, 0.5f), (float)bd.W);
    if (i < bd.rows) {
        const float pcy = fadd((float)(i + bd.yOffset), 0.5f);
        row[i] = make_float2(fdiv(pcy, (float)bd.fullH), __uint_as_float(hash1(__float_as_uint(pcy))));
    }
}

// BANDS = false: a whole-frame handle (no neighbour, no halo rows): the band branches are not compiled in
template <int THREADS, bool BANDS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
gatherStochasticKernel(const __grid_constant__ FrameConsts fc, BandDims bd, const float* __restrict__ gDepth, const uint32_t* __restrict__ gNormal,
                       const uint32_t* __restrict__ gAlbedo, const uint16_t* __restrict__ gRefl, const uint32_t* __restrict__ gRma,
                       const unsigned short* __restrict__ ssao, const SgiVplData* __restrict__ lightData, const int32_t* __restrict__ numVisible,
                       const SgiAtlasEntry* __restrict__ shadowIndices, const __grid_constant__ AtlasTable atlases,
                       uint2* __restrict__ colorOut, uint2* __restrict__ reservoirOut, float4* __restrict__ guideOut, uint2* __restrict__ mergedOut,
                       const __grid_constant__ BandLinks links, int tableCap, int* __restrict__ unitCounters, int counterSel,
                       CoordTables tabs, uint32_t unitsXMagic /* ceil(2^32 / unitsX), or 0: divide */, int unitChunk /* units per counter fetch */) {
    extern __shared__ float4 sTab[];
    float4* __restrict__ sPos = sTab;                       // (position, radius)
    float4* __restrict__ sCol = sTab + tableCap;            // (colour, farPlane)
    float4* __restrict__ sSpec = sTab + 2 * tableCap;       // (specular position, atlas index << 16 | layer)
    const int V = __ldg(numVisible);
    const int staged = min(max(V, 0), tableCap);
    for (int i = threadIdx.x; i < staged; i += THREADS) {
        const float4* L = reinterpret_cast<const float4*>(lightData + i);
        const float4 lp = __ldg(L), lc = __ldg(L + 1), sp = __ldg(L + 2), rfi = __ldg(L + 3);
        const SgiAtlasEntry e = shadowIndices[i];
        sPos[i] = make_float4(lp.x, lp.y, lp.z, rfi.x);
        sCol[i] = make_float4(lc.x, lc.y, lc.z, rfi.y);
        sSpec[i] = make_float4(sp.x, sp.y, sp.z, __int_as_float((e.index << 16) | (e.layer & 0xffff)));
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int unitsX = (bd.W + 31) >> 5;
    const int units = unitsX * bd.rows;
    const CubeConsts cc = cubeConsts(bd.cubeRes);
    const m4 invPV = loadMat(fc.invPV);
    const f3 viewPosition = mk3(fc.viewPos[0], fc.viewPos[1], fc.viewPos[2]);
    const float maxRandomIndexF = (float)(V - 1);
    const float minGi = fc.minGiOcclusionFactor;
    const int lastStaged = max(staged - 1, 0);

    // Warps fetch 32-pixel row segments from a global counter, kUnitChunk at a time, so the kernel finishes on time even when not
    // all of its CTAs are resident from the start (sgi_frame runs the tile-list stages next to it on a side stream).  Counter
    // [frame & 1] serves this launch; the other one is re-armed for the next launch on this stream.
    int topLo = 0, nTop = 0, botLo = 0, nBot = 0;               // band-local row ranges pushed to the upper / lower neighbour
    if (BANDS && links.doneCounter != nullptr) {
        if (links.up.res != nullptr) { topLo = links.up.firstRow - bd.yOffset; nTop = links.up.lastRow - links.up.firstRow; }
        if (links.down.res != nullptr) { botLo = links.down.firstRow - bd.yOffset; nBot = links.down.lastRow - links.down.firstRow; }
        if (nTop > 0 && nBot > 0 && topLo + nTop > botLo) { nTop = 0; nBot = 0; }      // overlapping ranges (a very thin band): plain order
    }
    int* __restrict__ counter = unitCounters + (counterSel & 1);
    if (blockIdx.x == 0 && threadIdx.x == 0) unitCounters[(counterSel & 1) ^ 1] = 0;
    // unit -> (band-local row, first column).  Row order: the rows whose pass-0 output goes to a neighbour's halo first (top range,
    // bottom range), then the rest
    auto unitRow = [&](int u, int* ux) -> int {
        int ly = unitsXMagic != 0u ? (int)__umulhi((uint32_t)u, unitsXMagic) : u / unitsX;
        *ux = (u - ly * unitsX) * 32;
        if (BANDS) {
            if (ly < nTop) ly = topLo + ly;
            else if (ly < nTop + nBot) ly = botLo + (ly - nTop);
            else {
                ly -= nTop + nBot;
                if (nTop > 0 && ly >= topLo) ly += nTop;
                if (nBot > 0 && ly >= botLo) ly += nBot;
            }
        }
        return ly;
    };
    // The counter is read one chunk ahead (the atomic's round trip, ~700 cycles, is over before its value is needed), and while a
    // unit is processed the G-buffer lines of the warp's next unit are on their way into L2.
    int nextBase = 0;
    if (lane == 0) nextBase = atomicAdd(counter, unitChunk);
    int unit = 0, unitEnd = 0;
    while (true) {
        if (unit >= unitEnd) {
            const int base = __shfl_sync(0xffffffffu, nextBase, 0);
            if (base >= units) break;
            if (lane == 0) nextBase = atomicAdd(counter, unitChunk);
            unit = base; unitEnd = min(base + unitChunk, units);
        }
        const int thisUnit = unit++;
        int ux;
        const int ly = unitRow(thisUnit, &ux);
        const int x = ux + lane;
        const int y = ly + bd.yOffset;
        do {
        // row bands: a warp that will store pass-0 rows into a neighbour's halo first waits until that neighbour has finished
        // the previous frame (its final pass read those halo rows)
        if (BANDS) bandWaitWarp(links, links.up.res != nullptr && y >= links.up.firstRow && y < links.up.lastRow,
                                links.down.res != nullptr && y >= links.down.firstRow && y < links.down.lastRow);
        const int nb = __shfl_sync(0xffffffffu, nextBase, 0);        // (the warp is converged here)
        if (x >= bd.W) continue;
        const size_t p = (size_t)ly * bd.W + x;

        // every G-buffer value of the pixel is requested before the first one is used (one exposed memory latency per unit instead
        // of one per plane: the exact divisions below are calls on their slow path, which the compiler does not move loads across)
        const float depth = __ldg(gDepth + p);
        const uint32_t nrm = __ldg(gNormal + p);
        const uint32_t alb = __ldg(gAlbedo + p);
        const uint32_t rma = __ldg(gRma + p);
        const uint32_t aoBits = __ldg(ssao + p), reflBits = __ldg(gRefl + p);
        const float2 rowV = __ldg(tabs.row + ly);                                    // fsTexCoords.y and hash(gl_FragCoord.y) of this row
        const float texU = __ldg(tabs.u + x);
        if (unit < unitEnd) {                      // the chunk's next unit: the next 32 pixels of the row-major planes (W % 32 == 0), else close by
            const size_t pn = min(p + 32, (size_t)bd.W * bd.rows - 1);
            prefetchL2(gDepth + pn); prefetchL2(gNormal + pn); prefetchL2(gAlbedo + pn); prefetchL2(gRma + pn); prefetchL2(ssao + pn); prefetchL2(gRefl + pn);
        } else {
            if (nb < units) {
                int nx;
                const int nly = unitRow(nb, &nx);
                const size_t pn = (size_t)nly * bd.W + min(nx + lane, bd.W - 1);
                prefetchL2(gDepth + pn); prefetchL2(gNormal + pn); prefetchL2(gAlbedo + pn); prefetchL2(gRma + pn); prefetchL2(ssao + pn); prefetchL2(gRefl + pn);
            }
        }

        // ---- per-pixel set-up: exact (feeds the sample budget, the 500-unit shadow cut-off and every re-draw test) ----
        PixelConsts px;
        px.normal = decodeNormal(nrm);
        guideOut[p] = make_float4(px.normal.x, px.normal.y, px.normal.z, depth);
        if (BANDS && (ly < bd.c0 || ly >= bd.c0 + bd.crows)) continue;      // halo row of a band: only the guide plane is needed here
        const float pcx = fadd((float)x, 0.5f);                                      // gl_FragCoord.x
        px.fragPos = worldPositionFromDepth(texU, rowV.x, depth, invPV);
        const f3 toView = viewPosition - px.fragPos;
        const float distToCamera = length(toView);
        const f3 baseColor = mk3(vmax(unorm8(alb & 0xffu), 0.01f), vmax(unorm8((alb >> 8) & 0xffu), 0.01f), vmax(unorm8((alb >> 16) & 0xffu), 0.01f));
        const float roughness = vmax(0.5f, unorm8(rma & 0xffu));
        const float roughnessWeight = fsub(1.0f, roughness);
        const float distRatioToCamera = vmin(fsub(1.0f, fdiv(distToCamera, 1000.0f)), 1.0f);
        const int maxSamplesPerPixel = (int)mix(2.0f, 10.0f, roughnessWeight);
        // V <= 0: undefined in the reference (stale lights); defined here as "no light"
        const int sampleCount = V > 0 ? min(max(1, (int)fmul((float)maxSamplesPerPixel, distRatioToCamera)), kMaxSamples) : 0;
        const bool shadowsOn = distToCamera < 500.0f;

        // ---- phase 1: the accept / re-draw decisions of the pixel's draw sequence (vpl_pbr_gi.fs:148-171) ----
        // draw k: rand = random(vec3(gl_FragCoord.xy, time + 10000 k)); the hash of the third component is the same for every
        // pixel and comes precomputed in fc.hashZ (common.glsl:199-214).  `--i; continue;` re-draws without consuming a sample,
        // at most 5 times per pixel; after that every draw is accepted.
        const uint32_t seedXY = __float_as_uint(pcx) ^ __float_as_uint(rowV.y);
        uint32_t q0 = 0u, q1 = 0u, q2 = 0u;          // accepted VPL positions, six 16-bit slots in sample order
        int accepted = 0, resamples = 0;
        #pragma unroll
        for (int k = 0; k < kMaxDraws; ++k) {
            if (accepted >= sampleCount) break;
            const float rnd = floatConstruct(hash1(seedXY ^ fc.hashZ[k]));
            const int lightIndex = min((int)fmul(maxRandomIndexF, rnd), lastStaged);
            bool reject = false;
            if (resamples < kMaxResamples) {
                const float4 pr = sPos[lightIndex];
                const f3 lightMinusFrag = mk3(pr.x, pr.y, pr.z) - px.fragPos;        // exact (single rounding per component)
                reject = behindOrOutOfRange(px.normal, lightMinusFrag, dotFast(lightMinusFrag, lightMinusFrag), pr.w);
            }
            if (reject) ++resamples;
            else {
                const uint32_t sv = (uint32_t)lightIndex << ((accepted & 1) * 16);
                if (accepted < 2) q0 |= sv; else if (accepted < 4) q1 |= sv; else q2 |= sv;
                ++accepted;
            }
        }

        // ---- per-pixel set-up: radiance tier (BRDF_Base pbr2.glsl:146-168 and everything light-independent) ----
        {
            const float metallic = unorm8((rma >> 8) & 0xffu);
            const float ambientOcclusion = vclamp(h2f((unsigned short)aoBits), 0.35f, 1.0f);
            const float refl = unorm8(reflBits & 0xffu);
            float rr = roughness;
            if (fc.usePerceptualRoughness) rr = rr * rr;
            rr = fmaxf(fmaxf(0.0f, fc.minRoughness), rr);
            px.r2 = rr * rr;
            const float iv = rsqrtFast(dotFast(toView, toView));
            px.viewDir = mk3(toView.x * iv, toView.y * iv, toView.z * iv);
            px.NdotV = satFast(dotFast(px.normal, px.viewDir));
            px.ggxvTerm = sqrtFast(fma_(fma_(-px.r2, px.NdotV, px.NdotV), px.NdotV, px.r2));
            const float om = 1.0f - metallic, f0s = 0.16f * refl * refl * om;
            const f3 inv = mk3(rcpFast(baseColor.x + PREVENT_DIV_BY_ZERO), rcpFast(baseColor.y + PREVENT_DIV_BY_ZERO), rcpFast(baseColor.z + PREVENT_DIV_BY_ZERO));
            // diffuseColor / (diffuseDivisor + eps) / PI   and   f0 * specularMultiplier
            px.diffuseOverPi = mk3(om * baseColor.x * inv.x * (1.0f / PI_F), om * baseColor.y * inv.y * (1.0f / PI_F), om * baseColor.z * inv.z * (1.0f / PI_F));
            px.f0 = mk3(fma_(baseColor.x, metallic, f0s) * inv.x, fma_(baseColor.y, metallic, f0s) * inv.y, fma_(baseColor.z, metallic, f0s) * inv.z);
            px.fogAo = expFast(-fc.fogDensity * distToCamera) * ambientOcclusion;
        }

        // ---- phase 2: shade the accepted VPLs in sample order (vpl_pbr_gi.fs:173-198) ----
        // The shadow-map texel of sample s + 1 is requested before sample s is shaded: its latency (an L2 hit at best: the maps of
        // ~1000 visible VPLs do not fit the L1 left next to the table) passes under ~150 instructions of shading.
        f3 vplColor = mk3(0.0f);
        float validSamples = 0.0f;
        auto sampleIndex = [&](int s) -> int { return (int)(((s < 2 ? q0 : (s < 4 ? q1 : q2)) >> ((s & 1) * 16)) & 0xffffu); };
        uint32_t texelNext = 0u;
        if (shadowsOn && accepted > 0) { const int i0 = sampleIndex(0); texelNext = loadShadowTexel(px.fragPos, sPos[i0], __float_as_int(sSpec[i0].w), cc, atlases); }
        for (int s = 0; s < accepted; ++s) {
            const int lightIndex = sampleIndex(s);
            const uint32_t texel = texelNext;
            if (shadowsOn && s + 1 < accepted) { const int i1 = sampleIndex(s + 1); texelNext = loadShadowTexel(px.fragPos, sPos[i1], __float_as_int(sSpec[i1].w), cc, atlases); }
            const float4 pr = sPos[lightIndex], cf = sCol[lightIndex], se = sSpec[lightIndex];
            const f3 lightPosition = mk3(pr.x, pr.y, pr.z);
            const f3 lightMinusFrag = lightPosition - px.fragPos;
            shadeAcceptedLight(px, lightPosition, lightMinusFrag, dotFast(lightMinusFrag, lightMinusFrag), pr.w, cf.w, mk3(cf.x, cf.y, cf.z), mk3(se.x, se.y, se.z),
                               texel, shadowsOn, minGi, vplColor, validSamples);
        }
        validSamples = fmaxf(validSamples, 1.0f);
        const f3 color = baseColor + PREVENT_DIV_BY_ZERO;
        storeHalf4(colorOut, p, color.x, color.y, color.z, 1.0f);
        const unsigned short hr = f2h(fminf(vplColor.x, HALF_FLOAT_MAX)), hg = f2h(fminf(vplColor.y, HALF_FLOAT_MAX)), hb = f2h(fminf(vplColor.z, HALF_FLOAT_MAX)),
                             ha = f2h(validSamples);
        uint2 packed;
        packed.x = (uint32_t)hr | ((uint32_t)hg << 16);
        packed.y = (uint32_t)hb | ((uint32_t)ha << 16);
        reservoirOut[p] = packed;
        if (mergedOut != nullptr) {
            // Denoise pass 0 on the binary16 reservoir just stored.  mergeCentreTexel (sgi_merge.cuh) reduces to rgb / max(1e-5, a)
            // here: the normal is unit length, |z - z| = 0, and with a >= 1 and 1 <= V <= 4480 the single candidate's acceptance
            // ratio probability / cw is x / x = 1 exactly, which every draw in [0,1) passes.  V <= 0 leaves the reservoir empty.
            const float denom = vmax(PREVENT_DIV_BY_ZERO, h2f(ha)), invDenom = rcpFast(denom);
            uint2 mv = make_uint2(0u, 0x3c000000u);                      // (0, 0, 0, 1)
            if (V > 0) {
                mv.x = (uint32_t)mergedChannel(h2f(hr), denom, invDenom) | ((uint32_t)mergedChannel(h2f(hg), denom, invDenom) << 16);
                mv.y |= (uint32_t)mergedChannel(h2f(hb), denom, invDenom);
            }
            mergedOut[p] = mv;
            if (BANDS) pushRes(links, y, x, bd.W, mv);
        }
        } while (false);                                           // (`continue` above leaves the pixel, not the segment)
        if (BANDS && links.doneCounter != nullptr && pushesRow(links, y)) { // warp-uniform: this segment stored rows into a neighbour's halo
            __syncwarp();
            if (lane == 0) producerDone(links);
        }
    }
}

}  // namespace

// CUDA loads kernels lazily on first launch, and that load can stall behind a kernel that is spinning on a band flag
// (sgi_band.cu) — so every kernel of this file is loaded when the handle is created.
using GatherFn = void (*)(const FrameConsts, BandDims, const float*, const uint32_t*, const uint32_t*, const uint16_t*, const uint32_t*, const unsigned short*,
                          const SgiVplData*, const int32_t*, const SgiAtlasEntry*, const AtlasTable, uint2*, uint2*, float4*, uint2*, const BandLinks, int, int*, int,
                          CoordTables, uint32_t, int);
static GatherFn gatherFn(bool wide, bool bands) {
    if (wide) return bands ? gatherStochasticKernel<1024, true> : gatherStochasticKernel<1024, false>;
    return bands ? gatherStochasticKernel<512, true> : gatherStochasticKernel<512, false>;
}

bool preloadGatherKernels() {
    cudaFuncAttributes a; bool ok = true;
    ok = ok && cudaFuncGetAttributes(&a, gatherTileListKernel) == cudaSuccess;
    ok = ok && cudaFuncGetAttributes(&a, coordTableKernel) == cudaSuccess;
    // the persistent kernel keeps the VPL table in dynamic shared memory: up to 4480 x 48 B = 210 KB of the SM's 227 KB
    for (int wide = 0; wide < 2; ++wide) for (int bands = 0; bands < 2; ++bands) {
        GatherFn f = gatherFn(wide != 0, bands != 0);
        ok = ok && cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, wide ? SGI_MAX_TOTAL_VPLS_PER_FRAME * kTableBytesPerVpl : kTableBytesTwoCtas) == cudaSuccess;
        ok = ok && cudaFuncGetAttributes(&a, f) == cudaSuccess;
    }
    return ok;
}

// (x + 0.5) / W per column and ((y + 0.5) / fullH, hash(y + 0.5)) per held row: once per handle geometry (allocPlanes)
int launchCoordTables(cudaStream_t stream, const BandDims& bd, float* u, float2* row) {
    const int n = max(bd.W, bd.rows);
    coordTableKernel<<<(n + 255) / 256, 256, 0, stream>>>(bd, u, row);
    return 1;
}

static int smCount() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        cached[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
    }
    return cached[dev];
}

int launchGather(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, const float* depth, const uint32_t* normal,
                 const uint32_t* albedo, const uint16_t* refl, const uint32_t* rma, const unsigned short* ssao,
                 const SgiVplData* updated, const int32_t* visible, const SgiAtlasEntry* entries, const AtlasTable& atlases,
                 uint2* colorOut, uint2* reservoirOut, float4* guideOut, uint2* mergedOut, const BandLinks& links,
                 const SgiVplStage2PerTileOutputs* tileLists, const unsigned short* tilePos, int* unitCounters, int64_t* persistentLaunches, bool* signalled,
                 const float* coordU, const float2* coordRow) {
    if (signalled) *signalled = false;
    BandLinks plain = links;                 // the kernels without the in-kernel signal
    plain.doneCounter = nullptr;
    if (tileLists != nullptr) {
        dim3 block(kTileCtaX * SGI_TILE_DIVISOR, kTileCtaY * SGI_TILE_DIVISOR);
        dim3 grid((bd.W + block.x - 1) / block.x, (bd.rows + block.y - 1) / block.y);
        gatherTileListKernel<<<grid, block, 0, stream>>>(fc, bd, depth, normal, albedo, refl, rma, ssao, updated, visible, entries, atlases,
                                                                       colorOut, reservoirOut, guideOut, mergedOut, plain, tileLists, tilePos);
    } else {
        // persistent: every SM holds the table once (two 512-thread CTAs while two copies fit, else one 1024-thread CTA)
        const int cap = max(1, min(fc.totalNumLights, SGI_MAX_TOTAL_VPLS_PER_FRAME));
        const int bytes = cap * kTableBytesPerVpl;
        const int units = ((bd.W + 31) / 32) * bd.rows;
        // the two work counters alternate between PERSISTENT launches (each launch re-arms the other one), so the parity is
        // counted here and not per sgi_gather call: a tile-list launch in between must not flip it
        const int counterSel = (int)((*persistentLaunches)++ & 1);
        // producers of the in-kernel signal: the 32-pixel segments of the rows pushed to a neighbour (pass 0 is pushed only when fused)
        BandLinks armed = links;
        if (armed.doneCounter != nullptr) {
            int rows = 0;
            if (mergedOut != nullptr) for (int y = bd.yOffset + bd.c0; y < bd.yOffset + bd.c0 + bd.crows; ++y) {
                const bool up = armed.up.res != nullptr && y >= armed.up.firstRow && y < armed.up.lastRow;
                const bool down = armed.down.res != nullptr && y >= armed.down.firstRow && y < armed.down.lastRow;
                rows += (up || down) ? 1 : 0;
            }
            armed.producers = rows * ((bd.W + 31) / 32);
            if (armed.producers == 0) armed.doneCounter = nullptr;
            else if (signalled) *signalled = true;
        }
        const bool bands = bd.c0 != 0 || bd.crows != bd.rows || links.up.res != nullptr || links.down.res != nullptr || links.waitUp != nullptr ||
                           links.waitDown != nullptr || links.doneCounter != nullptr;
        const bool wide = bytes > kTableBytesTwoCtas;
        const uint32_t unitsX = (uint32_t)((bd.W + 31) / 32);
        // row of a unit = unit / unitsX as one multiply-high: exact while unit * unitsX < 2^32 (8K: 2.5e8)
        const uint32_t magic = (unitsX > 1u && (uint64_t)units * unitsX < (1ull << 32)) ? (uint32_t)((1ull << 32) / unitsX) + 1u : 0u;
        CoordTables tabs; tabs.u = coordU; tabs.row = coordRow;
        const int threads = wide ? 1024 : 512;
        const int grid = wide ? max(1, min(smCount(), (units + 31) / 32)) : max(1, min(2 * smCount(), (units + 15) / 16));
        // two units per fetch halve the counter traffic; a short launch (a thin row band: < 24 units per warp) takes them one by one,
        // so that its warps finish within one unit (6 us) of each other instead of two
        const int unitChunk = units >= 24 * grid * (threads / 32) ? kUnitChunk : 1;
        GatherFn kf = gatherFn(wide, bands);
        kf<<<dim3(grid), dim3(threads), bytes, stream>>>(fc, bd, depth, normal, albedo, refl, rma, ssao, updated, visible, entries, atlases,
                               colorOut, reservoirOut, guideOut, mergedOut, armed, cap, unitCounters, counterSel, tabs, magic, unitChunk);
    }
    return 1;
}

}  // namespace sgi
