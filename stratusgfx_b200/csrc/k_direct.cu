// k_direct.cu — deferred direct lighting (widening row f3 of SURVEY.md §8f): the pass that produces the `screen` term the GI path
// adds its result to.
//
// Replaces pbr.fs:82-144 (main) and what it calls: calculatePointLighting2 / calculateDirectionalLighting (pbr2.glsl:576-611) ->
// calculateLighting_Burley (:454-476) -> BRDF_Burley (:194-250) [BRDF_Base :146-168, BRDF_DotProducts :170-188,
// singleScatteringBRDF_Specular :102-115, singleScatteringBRDF_Diffuse_Burley :117-126, Fd_Burley :91-96], applyFog (fog.glsl:37-41),
// quadraticAttenuation / calculateShadowValue8Samples / calculateShadowValue1Sample / calculateInfiniteShadowValue (pbr.glsl).
//
// Parity: bit-identical on a B200 to the reference's own pbr.fs compiled by oracle/make_ref_glsl.py (tests/test_zz_direct_lighting.py,
// profiles/r02a), and on the host, where tests/test_direct_kernel_on_host.py compiles this very file with g++ and runs it thread by thread.
//
// One thread per pixel, every operation in the exact IEEE chain of csrc/sgi_math.cuh in the shader's order (only
// pow/exp come from the CUDA maths library), so that the first comparison with the shader text is about logic, not rounding.
#include "sgi_internal.h"
#include "sgi_math.cuh"

namespace sgi {
namespace {

constexpr float PI_F = 3.14159265359f;
constexpr float PREVENT_DIV_BY_ZERO = 0.00001f;
constexpr float HALF_FLOAT_MAX = 65504.0f;

struct DirectArgs {
    int numLights, numShadowLights, sunEnabled;
    const SgiPointLight* nonShadowCasters; const SgiPointLight* shadowCasters; const SgiAtlasEntry* shadowIndices;
    float emissionStrength, depthBias, zNear, zFar, pointLightAmbientIntensity;
};

// hardware PCF of a sampler2DArrayShadow (LINEAR, compare LEQUAL): SURVEY.md Appendix A4, as k_vpl.cu / oracle.cpp
SGI_DEV float shadowArrayPcf(const float* __restrict__ csm, int res, float u, float v, int layer, float ref) {
    ref = vclamp(ref, 0.0f, 1.0f);
    const float fu = fsub(fmul(u, (float)res), 0.5f), fv = fsub(fmul(v, (float)res), 0.5f);
    const float i0f = floorf(fu), j0f = floorf(fv);
    const float a = fsub(fu, i0f), b = fsub(fv, j0f);
    const int i0 = iclamp((int)i0f, 0, res - 1), i1 = iclamp((int)i0f + 1, 0, res - 1);
    const int j0 = iclamp((int)j0f, 0, res - 1), j1 = iclamp((int)j0f + 1, 0, res - 1);
    const float* L = csm + (size_t)layer * res * res;
    const float t00 = ref <= __ldg(L + (size_t)j0 * res + i0) ? 1.0f : 0.0f, t10 = ref <= __ldg(L + (size_t)j0 * res + i1) ? 1.0f : 0.0f;
    const float t01 = ref <= __ldg(L + (size_t)j1 * res + i0) ? 1.0f : 0.0f, t11 = ref <= __ldg(L + (size_t)j1 * res + i1) ? 1.0f : 0.0f;
    const float na = fsub(1.0f, a), nb = fsub(1.0f, b);
    float r = fmul(fmul(na, nb), t00);
    r = fadd(r, fmul(fmul(a, nb), t10));
    r = fadd(r, fmul(fmul(na, b), t01));
    r = fadd(r, fmul(fmul(a, b), t11));
    return r;
}

// pbr.glsl:158-227, one thread: lit fraction of the sun at fragPos (cascade choice, blend, 3x3 taps on two cascades)
SGI_DEV float calculateInfiniteShadowValue(const FrameConsts& fc, const float* __restrict__ csm, int csmRes, f3 fragPos, f3 cascadeBlends, f3 normal, float bias) {
    const f3 sun = mk3(fc.sunDir[0], fc.sunDir[1], fc.sunDir[2]);
    const f3 off = (normal * fsub(1.0f, saturate(dot(normal, sun)))) * 1.0f;
    const f4 position = mk4(fadd(fragPos.x, off.x), fadd(fragPos.y, off.y), fadd(fragPos.z, off.z), 1.0f);
    const bool beyondCascade2 = cascadeBlends.y >= 0.0f, beyondCascade3 = cascadeBlends.z >= 0.0f;
    const int index1 = beyondCascade2 ? 2 : 0, index2 = beyondCascade3 ? 3 : 1;
    const f4 c1 = mul(loadMat(fc.cascadeProjViews[index1]), position), c2 = mul(loadMat(fc.cascadeProjViews[index2]), position);
    const f3 cc1 = (mk3(c1.x, c1.y, c1.z) / c1.w) * 0.5f + 0.5f, cc2 = (mk3(c2.x, c2.y, c2.z) / c2.w) * 0.5f + 0.5f;
    const float depth1 = cc1.z, depth2 = beyondCascade3 ? saturate(cc2.z) : cc2.z;
    const float weight = beyondCascade2 ? fsub(saturate(cascadeBlends.y), saturate(cascadeBlends.z)) : fsub(1.0f, saturate(cascadeBlends.x));
    const float wh = fdiv(1.0f, (float)csmRes);
    float light1 = 0.0f, light2 = 0.0f, samples = 0.0f;
    for (int iy = -1; iy <= 1; ++iy)
        for (int ix = -1; ix <= 1; ++ix) {
            const float ox = fmul((float)ix, wh), oy = fmul((float)iy, wh);
            light1 = fadd(light1, shadowArrayPcf(csm, csmRes, fadd(cc1.x, ox), fadd(cc1.y, oy), index1, fsub(depth1, bias)));
            light2 = fadd(light2, shadowArrayPcf(csm, csmRes, fadd(cc2.x, ox), fadd(cc2.y, oy), index2, fsub(depth2, bias)));
            samples = fadd(samples, 1.0f);
        }
    return fmul(mix(light2, light1, weight), fdiv(1.0f, samples));
}

// texture(samplerCubeArray, vec4(dir, layer)).r of a unorm16 depth map.  The texel is the exact one (cubeTexelDecided, sgi_math.cuh:
// an estimate with a margin, the exact division chain of cubeTexel only for a coordinate next to a texel edge) and texel / 65535.0f
// is unorm16's three instructions, bit for bit the division (tests/test_device_math_identities.py).
SGI_DEV float cubeDepth(const AtlasTable& atlases, SgiAtlasEntry e, f3 dir, int res) {
    const float halfRes = 0.5f * (float)res, epsK = (float)res * (1.0f / 4194304.0f);
    return unorm16(__ldg(atlases.shadow[e.index] + (size_t)e.layer * 6 * res * res + cubeTexelDecided(dir, res, halfRes, epsK)));
}
// pbr.glsl:116-136 and :79-112
SGI_DEV float calculateShadowValue1Sample(const AtlasTable& atlases, SgiAtlasEntry e, int res, float farPlane, f3 fragPos, f3 lightPos, float lightNormalDot, float minBias) {
    const f3 fragDir = fragPos - lightPos;
    const float currentDepth = length(fragDir);
    const float bias = fmul(currentDepth, vmax(minBias, fmul(saturate(lightNormalDot), 0.03f)));
    return fsub(currentDepth, bias) > fmul(cubeDepth(atlases, e, fragDir, res), farPlane) ? 1.0f : 0.0f;
}
SGI_DEV float calculateShadowValue8Samples(const AtlasTable& atlases, SgiAtlasEntry e, int res, float farPlane, f3 fragPos, f3 lightPos, float lightNormalDot, float minBias) {
    const f3 fragDir = fragPos - lightPos;
    const float currentDepth = length(fragDir);
    const float bias = fmul(currentDepth, vmax(minBias, fmul(saturate(lightNormalDot), 0.03f)));
    float shadow = 0.0f;
    const float offsets[2] = {-0.3f, 0.3f};
    for (int x = 0; x < 2; ++x) for (int y = 0; y < 2; ++y) for (int z = 0; z < 2; ++z) {
        const float depth = fmul(cubeDepth(atlases, e, fragDir + mk3(offsets[x], offsets[y], offsets[z]), res), farPlane);
        if (fsub(currentDepth, bias) > depth) shadow = fadd(shadow, 1.0f);
    }
    return fmul(shadow, 0.125f);                  // shadow / 8.0: a power of two, the same rounding
}

struct Surface { f3 fragPos, baseColor, normal, viewDir, baseReflectance; float viewDist, roughness, metallic, ambient; };

// What calculateLighting_Burley computes from the surface alone, once per pixel instead of once per light (each value by the same
// operations in the same order, so nothing changes but the count): BRDF_Base (pbr2.glsl:146-168), the view half of
// BRDF_DotProducts and of the Smith / Burley terms, the fog factor (fog.glsl:37-41).
struct SurfaceTerms { f3 diffuseColor, f0; float remappedRoughness, r2, NdotV, ggxlRoot /* sqrt((NdotV - r2 NdotV) NdotV + r2) */, viewPow5, fog; };
SGI_DEV SurfaceTerms surfaceTerms(const FrameConsts& fc, const Surface& s) {
    SurfaceTerms t;
    float rr = s.roughness;
    if (fc.usePerceptualRoughness) rr = fmul(rr, rr);
    t.remappedRoughness = vmax(vmax(0.0f, fc.minRoughness), rr);
    f3 diffuseColor = fsub(1.0f, s.metallic) * s.baseColor;
    t.diffuseColor = diffuseColor / (mk3(1.0f) + PREVENT_DIV_BY_ZERO);
    f3 f0 = ((0.16f * s.baseReflectance) * s.baseReflectance) * fsub(1.0f, s.metallic) + s.baseColor * s.metallic;
    t.f0 = f0 * mk3(1.0f);
    t.NdotV = saturate(dot(s.normal, s.viewDir));
    t.r2 = fmul(t.remappedRoughness, t.remappedRoughness);
    t.ggxlRoot = fsqrt(fadd(fmul(fsub(t.NdotV, fmul(t.r2, t.NdotV)), t.NdotV), t.r2));
    t.viewPow5 = powf(fsub(1.0f, t.NdotV), 5.0f);
    t.fog = expf(fmul(-fc.fogDensity, s.viewDist));
    return t;
}

// calculateLighting_Burley (pbr2.glsl:454-476) with diffuseDivisor = specularMultiplier = vec3(1.0)
SGI_DEV f3 calculateLightingBurley(const FrameConsts& fc, const Surface& s, const SurfaceTerms& st, f3 lightColor, f3 lightDir, float fogIntensity, float shadowFactor,
                                   float attenuationFactor, float ambientIntensity) {
    const float remappedRoughness = st.remappedRoughness;
    // BRDF_DotProducts (:170-188)
    const f3 V = s.viewDir, L = normalize(lightDir), H = normalize(V + L), N = s.normal;
    const float NdotV = st.NdotV, NdotL = saturate(dot(N, L)), NdotH = saturate(dot(N, H)), LdotH = saturate(dot(L, H));
    // singleScatteringBRDF_Specular (:102-115)
    const float r2 = st.r2;
    const float fN = fadd(fmul(fsub(fmul(NdotH, r2), NdotH), NdotH), 1.0f);
    const float NDF = fdiv(r2, vmax(fmul(fmul(PI_F, fN), fN), PREVENT_DIV_BY_ZERO));
    const float fF = powf(fsub(1.0f, LdotH), 5.0f);
    const f3 Fresnel = fF + st.f0 * fsub(1.0f, fF);
    const float ggxv = fmul(NdotL, st.ggxlRoot);
    const float ggxl = fmul(NdotV, fsqrt(fadd(fmul(fsub(NdotL, fmul(r2, NdotL)), NdotL), r2)));
    const float Visibility = fdiv(0.5f, vmax(fadd(ggxv, ggxl), PREVENT_DIV_BY_ZERO));
    const f3 Fr = fmul(NDF, Visibility) * Fresnel;
    // singleScatteringBRDF_Diffuse_Burley (:117-126) -> Fd_Burley (:91-96) -> Fd_Fresnel_Schlick (:85-88)
    const float f90 = fadd(0.5f, fmul(fmul(fmul(2.0f, remappedRoughness), LdotH), LdotH));
    const float lightScatter = fadd(1.0f, fmul(fsub(f90, 1.0f), powf(fsub(1.0f, NdotL), 5.0f)));
    const float viewScatter = fadd(1.0f, fmul(fsub(f90, 1.0f), st.viewPow5));
    const f3 Fd = st.diffuseColor * fmul(fmul(lightScatter, viewScatter), fdiv(1.0f, PI_F));
    const f3 brdf = (Fd + Fr) * NdotL;
    const f3 ambient = ((brdf * s.ambient) * lightColor) * ambientIntensity;
    const f3 finalBrightness = brdf * lightColor;
    // applyFog (fog.glsl:37-41)
    const float f = st.fog;
    const f3 fogC = mk3(fc.fogColor[0], fc.fogColor[1], fc.fogColor[2]) * fogIntensity;
    const f3 fogged = finalBrightness * f + fogC * fsub(1.0f, f);
    return attenuationFactor * (ambient + shadowFactor * fogged);
}

__global__ void __launch_bounds__(256)
directKernel(const __grid_constant__ FrameConsts fc, BandDims bd, const float* __restrict__ gDepth, const uint32_t* __restrict__ gNormal,
             const uint32_t* __restrict__ gAlbedo, const uint16_t* __restrict__ gRefl, const uint32_t* __restrict__ gRma,
             const unsigned short* __restrict__ ssao, const float* __restrict__ csm, const __grid_constant__ AtlasTable atlases,
             const __grid_constant__ DirectArgs da, uint2* __restrict__ out) {
    const int x = blockIdx.x * 32 + threadIdx.x, ly = blockIdx.y * 8 + threadIdx.y;
    if (x >= bd.W || ly >= bd.rows) return;
    const int y = ly + bd.yOffset;
    const size_t p = (size_t)ly * bd.W + x;
    const float u = fdiv(fadd((float)x, 0.5f), (float)bd.W), v = fdiv(fadd((float)y, 0.5f), (float)bd.fullH);
    Surface s;
    s.fragPos = worldPositionFromDepth(u, v, __ldg(gDepth + p), loadMat(fc.invPV));
    const f3 viewMinusFrag = mk3(fc.viewPos[0], fc.viewPos[1], fc.viewPos[2]) - s.fragPos;
    s.viewDir = normalize(viewMinusFrag);
    s.viewDist = length(viewMinusFrag);
    const uint32_t alb = __ldg(gAlbedo + p), rma = __ldg(gRma + p), refl = __ldg(gRefl + p);
    s.baseColor = mk3(unorm8(alb & 0xffu), unorm8((alb >> 8) & 0xffu), unorm8((alb >> 16) & 0xffu));
    s.normal = decodeNormal(__ldg(gNormal + p));
    s.roughness = unorm8(rma & 0xffu); s.metallic = unorm8((rma >> 8) & 0xffu);
    s.ambient = h2f(__ldg(ssao + p));                        // texture(ssao, texCoords * vec2(windowWidth, windowHeight)).r: the pixel's own texel
    s.baseReflectance = mk3(unorm8(refl & 0xffu));
    const f3 emissive = mk3(unorm8(alb >> 24), unorm8((refl >> 8) & 0xffu), unorm8((rma >> 16) & 0xffu));

    const SurfaceTerms st = surfaceTerms(fc, s);
    f3 color = mk3(0.0f);
    for (int i = 0; i < da.numLights; ++i) {
        const float4 lp = __ldg(reinterpret_cast<const float4*>(da.nonShadowCasters + i)), lc = __ldg(reinterpret_cast<const float4*>(da.nonShadowCasters + i) + 1);
        const float radius = __ldg(&da.nonShadowCasters[i].radius);
        const f3 lightPos = mk3(lp.x, lp.y, lp.z), lightColor = mk3(lc.x, lc.y, lc.z), lightDir = lightPos - s.fragPos;
        if (length(lightDir) < radius) {
            const float dist = length(lightDir);
            color = color + calculateLightingBurley(fc, s, st, lightColor, lightDir, fdiv(length(lightColor), 12.0f), fsub(1.0f, 0.0f),
                                                    fdiv(1.0f, fadd(1.0f, fmul(dist, dist))), da.pointLightAmbientIntensity);
        }
    }
    for (int i = 0; i < da.numShadowLights; ++i) {
        const float4 lp = __ldg(reinterpret_cast<const float4*>(da.shadowCasters + i)), lc = __ldg(reinterpret_cast<const float4*>(da.shadowCasters + i) + 1);
        const float radius = __ldg(&da.shadowCasters[i].radius), farPlane = __ldg(&da.shadowCasters[i].farPlane);
        const f3 lightPos = mk3(lp.x, lp.y, lp.z), lightColor = mk3(lc.x, lc.y, lc.z), lightDir = lightPos - s.fragPos;
        if (length(lightDir) < radius) {
            float shadowFactor = 0.0f;
            const SgiAtlasEntry e = da.shadowIndices[i];
            if (s.viewDist < 150.0f) shadowFactor = calculateShadowValue8Samples(atlases, e, bd.cubeRes, farPlane, s.fragPos, lightPos, dot(lightDir, s.normal), 0.03f);
            else if (s.viewDist < 750.0f) shadowFactor = calculateShadowValue1Sample(atlases, e, bd.cubeRes, farPlane, s.fragPos, lightPos, dot(lightDir, s.normal), 0.03f);
            const float dist = length(lightDir);
            color = color + calculateLightingBurley(fc, s, st, lightColor, lightDir, fdiv(length(lightColor), 12.0f), fsub(1.0f, shadowFactor),
                                                    fdiv(1.0f, fadd(1.0f, fmul(dist, dist))), da.pointLightAmbientIntensity);
        }
    }
    if (da.sunEnabled) {                                     // #ifdef INFINITE_LIGHTING_ENABLED (pbr.fs:130-141)
        const f4 fp4 = mk4(s.fragPos.x, s.fragPos.y, s.fragPos.z, 1.0f);
        f3 blends;
        blends.x = dot(mk4(fc.cascadePlanes[0][0], fc.cascadePlanes[0][1], fc.cascadePlanes[0][2], fc.cascadePlanes[0][3]), fp4);
        blends.y = dot(mk4(fc.cascadePlanes[1][0], fc.cascadePlanes[1][1], fc.cascadePlanes[1][2], fc.cascadePlanes[1][3]), fp4);
        blends.z = dot(mk4(fc.cascadePlanes[2][0], fc.cascadePlanes[2][1], fc.cascadePlanes[2][2], fc.cascadePlanes[2][3]), fp4);
        const float bias = fdiv(da.depthBias, fsub(da.zFar, da.zNear));          // useDepthBias = true (pbr.fs:139)
        const float litFraction = calculateInfiniteShadowValue(fc, csm, bd.csmRes, s.fragPos, blends, s.normal, bias);
        // calculateDirectionalLighting(..., 1.0 - shadowFactor, ..., 0.0) passes 1.0 - (1.0 - shadowFactor) on to calculateLighting_Burley
        const f3 sunDir = mk3(fc.sunDir[0], fc.sunDir[1], fc.sunDir[2]), sunColor = mk3(fc.sunColor[0], fc.sunColor[1], fc.sunColor[2]);
        color = color + calculateLightingBurley(fc, s, st, sunColor, sunDir, 0.0f, fsub(1.0f, fsub(1.0f, litFraction)), 1.0f, 0.0f);
    }
    const f3 result = vmin(color + emissive * da.emissionStrength, HALF_FLOAT_MAX);          // boundHDR (common.glsl:89-92)
    storeHalf4(out, p, result.x, result.y, result.z, 1.0f);
}

}  // namespace

bool preloadDirectKernels() {
    cudaFuncAttributes a;
    return cudaFuncGetAttributes(&a, directKernel) == cudaSuccess;
}

int launchDirectLighting(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, const float* depth, const uint32_t* normal, const uint32_t* albedo,
                         const uint16_t* refl, const uint32_t* rma, const unsigned short* ssao, const float* csm, const AtlasTable& atlases,
                         const SgiDirectParams& dp, uint2* out) {
    DirectArgs da;
    da.numLights = dp.numLights; da.numShadowLights = dp.numShadowLights; da.sunEnabled = dp.sunEnabled;
    da.nonShadowCasters = dp.nonShadowCasters; da.shadowCasters = dp.shadowCasters; da.shadowIndices = dp.shadowIndices;
    da.emissionStrength = dp.emissionStrength; da.depthBias = dp.infiniteLightDepthBias; da.zNear = dp.infiniteLightZnear; da.zFar = dp.infiniteLightZfar;
    da.pointLightAmbientIntensity = dp.pointLightAmbientIntensity;
    dim3 block(32, 8), grid((bd.W + 31) / 32, (bd.rows + 7) / 8);
    directKernel<<<grid, block, 0, stream>>>(fc, bd, depth, normal, albedo, refl, rma, ssao, csm, atlases, da, out);
    return 1;
}

}  // namespace sgi
