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
// One thread per pixel.  Which VPL a pixel draws is a pure integer function of (x+0.5, y+0.5, time)
// and the re-draw decision is IEEE-exact arithmetic in glm's evaluation order, so the light choice
// is identical to the CPU transcription; only exp/pow differ (by ulps).
#include "sgi_internal.h"
#include "sgi_math.cuh"

namespace sgi {
namespace {

constexpr float PI_F = 3.14159265359f;
constexpr float PREVENT_DIV_BY_ZERO = 0.00001f;
constexpr float HALF_FLOAT_MAX = 65504.0f;
constexpr int kGatherThreadsX = 32, kGatherThreadsY = 8;

SGI_DEV float vplDiffuseAttenuation(f3 lightDir, float lightRadius) {
    const float lightDist = length(lightDir);
    const float ratio = vmin(fdiv(lightDist, lightRadius), 1.0f);
    const float minDist = fmul(mix(10.0f, 1.0f, expf(fmul(-3.0f, ratio))), lightRadius);
    return fdiv(1.0f, fadd(minDist, fmul(fmul(1.0f, lightDist), lightDist)));
}

SGI_DEV float NDF_GGX(float NdotH, float roughness) {
    const float r2 = fmul(roughness, roughness);
    const float f = fadd(fmul(fsub(fmul(NdotH, r2), NdotH), NdotH), 1.0f);
    return fdiv(r2, vmax(fmul(fmul(PI_F, f), f), PREVENT_DIV_BY_ZERO));
}

SGI_DEV float Visibility_G_SmithGGX(float NdotV, float NdotL, float roughness) {
    const float a2 = fmul(roughness, roughness);
    const float ggxv = fmul(NdotL, fsqrt(fadd(fmul(fsub(NdotV, fmul(a2, NdotV)), NdotV), a2)));
    const float ggxl = fmul(NdotV, fsqrt(fadd(fmul(fsub(NdotL, fmul(a2, NdotL)), NdotL), a2)));
    return fdiv(0.5f, vmax(fadd(ggxv, ggxl), PREVENT_DIV_BY_ZERO));
}

struct Dots { float NdotV, NdotL, NdotH, LdotH; };
SGI_DEV Dots BRDF_DotProducts(f3 V, f3 lightDir, f3 N) {
    const f3 L = normalize(lightDir), H = normalize(V + L);
    Dots d;
    d.NdotV = saturate(dot(N, V)); d.NdotL = saturate(dot(N, L));
    d.NdotH = saturate(dot(N, H)); d.LdotH = saturate(dot(L, H));
    return d;
}

struct PixelConsts {   // per-pixel values that do not depend on the light
    f3 fragPos, baseColor, normal, viewDir, diffuseColor, f0base;
    float distToCamera, roughness, remappedRoughness, metallic, ambientOcclusion, ambient, fogF;
};

// calculateVirtualPointLighting2 with the constant arguments of the call site (vpl_pbr_gi.fs:195-197)
SGI_DEV f3 calculateVirtualPointLighting2(const FrameConsts& fc, const PixelConsts& px, f3 specularLightPos, f3 diffuseLightPos, f3 lightColor,
                                          float lightRadius) {
    const f3 specularLightDir = specularLightPos - px.fragPos;
    const f3 diffuseLightDir = diffuseLightPos - px.fragPos;
    const float adjustedShadowFactor = fsub(1.0f, 0.0f);
    const float specularAttenuationFactor = vplDiffuseAttenuation(specularLightDir, lightRadius);
    const float diffuseAttenuationFactor = vplDiffuseAttenuation(diffuseLightDir, lightRadius);

    Dots d = BRDF_DotProducts(px.viewDir, specularLightDir, px.normal);
    const float NDF = NDF_GGX(d.NdotH, px.remappedRoughness);
    const float fr = powf(fsub(1.0f, d.LdotH), 5.0f);
    const f3 Fresnel = fr + px.f0base * fsub(1.0f, fr);
    const float Visibility = Visibility_G_SmithGGX(d.NdotV, d.NdotL, px.remappedRoughness);
    const f3 Fr = (fmul(NDF, Visibility) * Fresnel) * d.NdotL;

    d = BRDF_DotProducts(px.viewDir, diffuseLightDir, px.normal);
    const f3 Fd = (px.diffuseColor * fdiv(1.0f, PI_F)) * d.NdotL;

    const f3 brdfSpecular = specularAttenuationFactor * Fr;
    const f3 brdfDiffuse = diffuseAttenuationFactor * Fd;
    const f3 brdf = brdfSpecular + brdfDiffuse;
    const f3 ambient = ((brdf * px.ambient) * lightColor) * 0.0f;      // ambientIntensity = 0
    const f3 finalBrightness = brdf * lightColor;
    const f3 fc3 = mk3(fc.fogColor[0], fc.fogColor[1], fc.fogColor[2]) * 0.0f;   // fogIntensity = 0
    const f3 fogged = finalBrightness * px.fogF + fc3 * fsub(1.0f, px.fogF);
    return ambient + adjustedShadowFactor * fogged;
}

__global__ void __launch_bounds__(kGatherThreadsX * kGatherThreadsY)
gatherKernel(const __grid_constant__ FrameConsts fc, BandDims bd, const float* __restrict__ gDepth, const uint32_t* __restrict__ gNormal,
             const uint32_t* __restrict__ gAlbedo, const uint16_t* __restrict__ gRefl, const uint32_t* __restrict__ gRma,
             const unsigned short* __restrict__ ssao, const SgiVplData* __restrict__ lightData, const int32_t* __restrict__ numVisible,
             const SgiAtlasEntry* __restrict__ shadowIndices, const __grid_constant__ AtlasTable atlases,
             uint2* __restrict__ colorOut, uint2* __restrict__ reservoirOut) {
    const int x = blockIdx.x * kGatherThreadsX + threadIdx.x;
    const int ly = blockIdx.y * kGatherThreadsY + threadIdx.y;
    if (x >= bd.W || ly >= bd.rows) return;
    const int y = ly + bd.yOffset;
    const size_t p = (size_t)ly * bd.W + x;
    const int cubeRes = bd.cubeRes;
    const int texelsPerLayer = 6 * cubeRes * cubeRes;

    const m4 invPV = loadMat(fc.invPV);
    const f3 viewPosition = mk3(fc.viewPos[0], fc.viewPos[1], fc.viewPos[2]);
    const float pcx = fadd((float)x, 0.5f), pcy = fadd((float)y, 0.5f);          // gl_FragCoord.xy
    const float u = fdiv(pcx, (float)bd.W), v = fdiv(pcy, (float)bd.fullH);       // fsTexCoords

    PixelConsts px;
    px.fragPos = worldPositionFromDepth(u, v, __ldg(gDepth + p), invPV);
    const f3 toView = viewPosition - px.fragPos;
    px.viewDir = normalize(toView);
    px.distToCamera = length(toView);
    const uint32_t alb = __ldg(gAlbedo + p);
    px.baseColor = mk3(vmax(unorm8(alb & 0xffu), 0.01f), vmax(unorm8((alb >> 8) & 0xffu), 0.01f), vmax(unorm8((alb >> 16) & 0xffu), 0.01f));
    px.normal = decodeNormal(__ldg(gNormal + p));
    const uint32_t rma = __ldg(gRma + p);
    px.roughness = vmax(0.5f, unorm8(rma & 0xffu));
    px.metallic = unorm8((rma >> 8) & 0xffu);
    px.ambientOcclusion = vclamp(h2f(__ldg(ssao + p)), 0.35f, 1.0f);
    px.ambient = unorm8((rma >> 16) & 0xffu);
    const f3 baseReflectivity = mk3(unorm8((uint32_t)__ldg(gRefl + p) & 0xffu));
    const float roughnessWeight = fsub(1.0f, px.roughness);

    // BRDF_Base (pbr2.glsl:146-168): light-independent, hoisted out of the sample loop
    {
        float rr = px.roughness;
        if (fc.usePerceptualRoughness) rr = fmul(rr, rr);
        px.remappedRoughness = vmax(vmax(0.0f, fc.minRoughness), rr);
        const f3 diffuseDivisor = px.baseColor;
        const f3 specularMultiplier = 1.0f / (px.baseColor + PREVENT_DIV_BY_ZERO);
        f3 diffuseColor = fsub(1.0f, px.metallic) * px.baseColor;
        px.diffuseColor = diffuseColor / (diffuseDivisor + PREVENT_DIV_BY_ZERO);
        f3 f0 = ((0.16f * baseReflectivity) * baseReflectivity) * fsub(1.0f, px.metallic) + px.baseColor * px.metallic;
        px.f0base = f0 * specularMultiplier;
        px.fogF = expf(fmul(-fc.fogDensity, px.distToCamera));
    }

    f3 vplColor = mk3(0.0f);
    float seedZ = fc.time;
    float validSamples = 0.0f;
    const float distRatioToCamera = vmin(fsub(1.0f, fdiv(px.distToCamera, 1000.0f)), 1.0f);
    const int maxSamplesPerPixel = (int)mix(2.0f, 10.0f, roughnessWeight);
    const int sampleCount = max(1, (int)fmul((float)maxSamplesPerPixel, distRatioToCamera));
    const int V = __ldg(numVisible);
    const int maxRandomIndex = V - 1;
    const float minGi = fc.minGiOcclusionFactor;

    // V <= 0: undefined in the reference (stale lights); defined here as "no light"
    for (int i = 0, resamples = 0; V > 0 && i < sampleCount; ++i) {
        const float rnd = random3(pcx, pcy, seedZ);
        seedZ = fadd(seedZ, 10000.0f);
        const int lightIndex = (int)fmul((float)maxRandomIndex, rnd);
        const float4* L = reinterpret_cast<const float4*>(lightData + lightIndex);
        const float4 lp = __ldg(L);
        const float4 rfi = __ldg(L + 3);                    // radius, farPlane, intensity, pad
        const f3 lightPosition = mk3(lp.x, lp.y, lp.z);
        const f3 lightMinusFrag = lightPosition - px.fragPos;
        const float lightRadius = rfi.x;
        const float distance = length(lightMinusFrag);
        if (resamples < 5) {
            const float sideCheck = dot(px.normal, normalize(lightMinusFrag));
            if (sideCheck < 0.0f || distance > lightRadius) { ++resamples; --i; continue; }
        }
        const float4 lc = __ldg(L + 1);
        const float4 sp = __ldg(L + 2);
        const float distanceRatio = vclamp(fdiv(fmul(2.0f, distance), lightRadius), 0.0f, 1.0f);
        float shadowFactor = 0.0f;
        if (px.distToCamera < 500.0f) {
            // calculateShadowValue1Sample (pbr.glsl:116-136), minBias 0.1
            const SgiAtlasEntry entry = shadowIndices[lightIndex];
            const f3 fragDir = px.fragPos - lightPosition;
            const float currentDepth = length(fragDir);
            const float lightNormalDot = dot(lightPosition - px.fragPos, px.normal);
            const float bias = fmul(currentDepth, vmax(0.1f, fmul(saturate(lightNormalDot), 0.03f)));
            const unsigned short texel = __ldg(atlases.shadow[entry.index] + (size_t)entry.layer * texelsPerLayer + cubeTexel(fragDir, cubeRes));
            const float depth = fmul(fdiv((float)texel, 65535.0f), rfi.y);
            if (fsub(currentDepth, bias) > depth) shadowFactor = 1.0f;
        }
        shadowFactor = vmin(shadowFactor, mix(minGi, 1.0f, distanceRatio));
        float reweightingFactor = 1.0f;
        if (shadowFactor > 0.0f) reweightingFactor = fadd(fmul(fsub(1.0f, distanceRatio), minGi), distanceRatio);
        validSamples = fadd(validSamples, reweightingFactor);
        const f3 tmpColor = px.ambientOcclusion * calculateVirtualPointLighting2(fc, px, mk3(sp.x, sp.y, sp.z), lightPosition, mk3(lc.x, lc.y, lc.z), lightRadius);
        vplColor = vplColor + fsub(1.0f, shadowFactor) * tmpColor;
    }
    validSamples = vmax(validSamples, 1.0f);
    const f3 color = px.baseColor + PREVENT_DIV_BY_ZERO;
    const f3 bounded = vmin(vplColor, HALF_FLOAT_MAX);
    storeHalf4(colorOut, p, color.x, color.y, color.z, 1.0f);
    storeHalf4(reservoirOut, p, bounded.x, bounded.y, bounded.z, validSamples);
}

}  // namespace

int launchGather(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, const float* depth, const uint32_t* normal,
                 const uint32_t* albedo, const uint16_t* refl, const uint32_t* rma, const unsigned short* ssao,
                 const SgiVplData* updated, const int32_t* visible, const SgiAtlasEntry* entries, const AtlasTable& atlases,
                 uint2* colorOut, uint2* reservoirOut) {
    dim3 block(kGatherThreadsX, kGatherThreadsY);
    dim3 grid((bd.W + kGatherThreadsX - 1) / kGatherThreadsX, (bd.rows + kGatherThreadsY - 1) / kGatherThreadsY);
    gatherKernel<<<grid, block, 0, stream>>>(fc, bd, depth, normal, albedo, refl, rma, ssao, updated, visible, entries, atlases, colorOut, reservoirOut);
    return 1;
}

}  // namespace sgi
