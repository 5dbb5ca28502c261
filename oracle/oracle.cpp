// oracle.cpp — CPU restatement of StratusGFX's VPL global-illumination shaders.
//
// *** TEST INFRASTRUCTURE, NOT PRODUCT. ***  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library; the CUDA
// product (stratusgfx_b200/csrc) never links, calls or falls back to it.
//
// HOW PARITY IS PINNED.  StratusGFX ships no golden vectors, known-answer tests or fixtures for this path,
// and its GLSL cannot run here as GLSL (no OpenGL).  What this file is held against instead:
// (0) THE REFERENCE'S OWN SHADER TEXT, compiled: oracle/make_ref_glsl.py turns the six shaders of the path (viscull_vpls.cs,
//     vpl_light_color.cs, tiled culling stage 1 / stage 2, vpl_pbr_gi.fs, vpl_pbr_gi_denoise.fs) and every pbr.glsl / pbr2.glsl /
//     fog.glsl / common.glsl function they call into C++ against the reference's vendored glm by a few lexical rules, with the GL
//     fixed function (texel fetch, cube face selection, hardware PCF, 16F stores, work-group scheduling) stubbed from the GL 4.6
//     specification.  tests/test_reference_glsl.py: every stage of this oracle equals that shader text bit for bit on whole frames
//     (the cull as a set); tests/golden/tiny_shader_text_golden.json is a whole two-frame chain computed without this file;
// (1) the maths layer is cross-checked bit for bit against the reference's own vendored glm (oracle/_ref build, -DORACLE_USE_GLM);
// (2) the self-checks of SURVEY.md §8(c) in tests/test_oracle_selfchecks.py;
// (3) frozen RNG known answers and a fixture made with the glm-backed build (tests/golden/);
// (4) for the pre-pass and the wire formats: the reference's own StratusMath.h / StratusGpuCommon.h compiled behind
//     oracle/ref_shim.cpp (tests/test_reference_shim.py).
// What stays unpinned: the GL fixed-function stubs themselves (this file's and the shim's follow the same reading of the
// specification, Appendix A of SURVEY.md) — no OpenGL driver exists in this image to check them against.
//
// Each function follows one shader entry point; citations are relative to the
// StratusGFX tree.  GL fixed-function semantics (texel addressing, cube-face
// selection, hardware PCF, 16F stores) follow SURVEY.md Appendix A.  Parity tests run
// it on whole frames (the product's band logic is checked against rows of the
// whole-frame result); OrcDims::y_offset/full_height exist so the CPU-baseline timing
// can take a bounded row-band sample of a 4K/8K frame with the right coordinates.
//
// Build: g++ -O2 -std=c++17 -fopenmp -ffp-contract=off  (no FMA contraction, no fast-math)
#include <cstdint>
#include <cstring>
#include <cmath>
#include <set>
#include <vector>
#include <omp.h>
#include "vm.h"
#include "../include/stratusgi.h"

namespace {
using namespace vm;

// ---- constants: Source/Shaders/common.glsl:3-4 -----------------------------------------------
const float PI_F = 3.14159265359f;
const float PREVENT_DIV_BY_ZERO = 0.00001f;
const float HALF_FLOAT_MAX = 65504.0f;
const float FLOAT_MAX_ = 3.402823466e+38f;

// ---- binary16 <-> binary32 (GL 16F render-target store = round to nearest even) -------------
inline uint16_t f2h(float f) {
    uint32_t x; std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const uint32_t em = x & 0x7fffffffu;
    if (em > 0x7f800000u) return (uint16_t)(sign | 0x7e00u);            // NaN
    if (em >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);           // >= 65520 -> inf
    if (em < 0x33000001u) return (uint16_t)sign;                        // <= 2^-25 -> 0
    const int e = (int)(em >> 23) - 127;
    uint32_t m = em & 0x7fffffu;
    int shift; uint32_t base;
    if (e < -14) { m |= 0x800000u; shift = 13 + (-14 - e); base = 0; }
    else { shift = 13; base = (uint32_t)(e + 15) << 10; }
    uint32_t q = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1), halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (q & 1u))) ++q;
    return (uint16_t)(sign | (base + q));
}
inline float h2f(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    const uint32_t e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    uint32_t x;
    if (e == 0) {
        if (m == 0) x = sign;
        else { float f = (float)m * (1.0f / 16777216.0f); std::memcpy(&x, &f, 4); x |= sign; }   // m * 2^-24
    } else if (e == 31) x = sign | 0x7f800000u | (m << 13);
    else x = sign | ((e + 112u) << 23) | (m << 13);
    float f; std::memcpy(&f, &x, 4); return f;
}

inline mat4 loadMat(const float* p) {
    mat4 m;
    for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) m[c][r] = p[c * 4 + r];
    return m;
}
inline float saturate(float v) { return clamp(v, 0.0f, 1.0f); }
inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline float unorm8(uint8_t b) { return (float)b / 255.0f; }

// ---- RNG: common.glsl:183-214 ----------------------------------------------------------------
inline uint32_t hash1(uint32_t x) {
    x += (x << 10u); x ^= (x >> 6u); x += (x << 3u); x ^= (x >> 11u); x += (x << 15u);
    return x;
}
inline uint32_t fbits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float floatConstruct(uint32_t m) {
    m &= 0x007FFFFFu; m |= 0x3F800000u;
    float f; std::memcpy(&f, &m, 4);
    return f - 1.0f;
}
inline float random3(vec3 v) {   // random(vec3) -> hash(uvec3)
    return floatConstruct(hash1(fbits(v.x) ^ hash1(fbits(v.y)) ^ hash1(fbits(v.z))));
}

// ---- common.glsl:143-155 ---------------------------------------------------------------------
inline vec3 worldPositionFromDepth(vec2 uv, float depth, const mat4& invProjectionView) {
    const float z = depth * 2.0f - 1.0f;
    const vec4 ndc = vec4(uv * 2.0f - 1.0f, z, 1.0f);
    const vec4 worldPosition = invProjectionView * ndc;
    return vec3(worldPosition) / worldPosition.w;
}

inline vec3 decodeNormal(const uint8_t* n4) {   // normalize(texel.rgb * 2.0 - vec3(1.0)), common.glsl:216-220
    return normalize(vec3(unorm8(n4[0]), unorm8(n4[1]), unorm8(n4[2])) * 2.0f - vec3(1.0f));
}

// ---- cube-map array lookup, NEAREST (GL 4.6 spec table 8.19) ----------
// Returns the texel index inside one cube layer: (face*res + t)*res + s.
inline int cubeTexel(vec3 d, int res) {
    const float ax = abs(d.x), ay = abs(d.y), az = abs(d.z);
    int face; float sc, tc, ma;
    if (ax >= ay && ax >= az) { ma = ax; if (d.x >= 0.0f) { face = 0; sc = -d.z; tc = -d.y; } else { face = 1; sc = d.z; tc = -d.y; } }
    else if (ay >= az)        { ma = ay; if (d.y >= 0.0f) { face = 2; sc = d.x; tc = d.z; } else { face = 3; sc = d.x; tc = -d.z; } }
    else                      { ma = az; if (d.z >= 0.0f) { face = 4; sc = d.x; tc = -d.y; } else { face = 5; sc = -d.x; tc = -d.y; } }
    const float s = (sc / ma + 1.0f) * 0.5f, t = (tc / ma + 1.0f) * 0.5f;
    const int i = clampi((int)floor(s * (float)res), 0, res - 1);
    const int j = clampi((int)floor(t * (float)res), 0, res - 1);
    return (face * res + j) * res + i;
}
inline float shadowCubeFetch(const void* const* atlases, SgiAtlasEntry e, vec3 dir, int res) {
    const uint16_t* a = (const uint16_t*)atlases[e.index];
    return (float)a[(size_t)e.layer * 6 * res * res + cubeTexel(dir, res)] / 65535.0f;
}

// ---- hardware PCF: sampler2DArrayShadow, LINEAR, compare LEQUAL ------
inline float shadowArrayPcf(const float* csm, int res, float u, float v, float layerF, float ref) {
    const int layer = clampi((int)floor(layerF + 0.5f), 0, SGI_NUM_CASCADES - 1);
    ref = clamp(ref, 0.0f, 1.0f);                     // D_ref is clamped for fixed-point depth formats
    const float fu = u * (float)res - 0.5f, fv = v * (float)res - 0.5f;
    const float i0f = floor(fu), j0f = floor(fv);
    const float a = fu - i0f, b = fv - j0f;
    const int i0 = clampi((int)i0f, 0, res - 1), i1 = clampi((int)i0f + 1, 0, res - 1);
    const int j0 = clampi((int)j0f, 0, res - 1), j1 = clampi((int)j0f + 1, 0, res - 1);
    const float* L = csm + (size_t)layer * res * res;
    const float t00 = ref <= L[(size_t)j0 * res + i0] ? 1.0f : 0.0f, t10 = ref <= L[(size_t)j0 * res + i1] ? 1.0f : 0.0f;
    const float t01 = ref <= L[(size_t)j1 * res + i0] ? 1.0f : 0.0f, t11 = ref <= L[(size_t)j1 * res + i1] ? 1.0f : 0.0f;
    return (1.0f - a) * (1.0f - b) * t00 + a * (1.0f - b) * t10 + (1.0f - a) * b * t01 + a * b * t11;
}

// ---- pbr.glsl:158-227 calculateInfiniteShadowValue (useDepthBias = false on this path) -------
float calculateInfiniteShadowValue(const SgiFrameParams& fp, const mat4 cascadeProjViews[4], const float* csm, int csmRes,
                                   vec4 fragPos, vec3 cascadeBlends, vec3 normal) {
    const vec3 infiniteLightDirection(fp.infiniteLightDirection[0], fp.infiniteLightDirection[1], fp.infiniteLightDirection[2]);
    const float bias = 0.0f;
    vec4 position = fragPos;
    {
        const vec3 off = normal * (1.0f - saturate(dot(normal, infiniteLightDirection))) * 1.0f;
        position.x += off.x; position.y += off.y; position.z += off.z;
    }
    vec3 cascadeCoords[4];
    for (int i = 0; i < 4; ++i) {
        const vec4 coords = cascadeProjViews[i] * position;
        cascadeCoords[i] = vec3(coords) / coords.w;
        cascadeCoords[i] = cascadeCoords[i] * 0.5f + vec3(0.5f);
    }
    const bool beyondCascade2 = cascadeBlends.y >= 0.0f;
    const bool beyondCascade3 = cascadeBlends.z >= 0.0f;
    const int index1 = beyondCascade2 ? 2 : 0;
    const int index2 = beyondCascade3 ? 3 : 1;
    const vec2 shadowCoord1(cascadeCoords[index1].x, cascadeCoords[index1].y);
    const vec2 shadowCoord2(cascadeCoords[index2].x, cascadeCoords[index2].y);
    const float depth1 = cascadeCoords[index1].z;
    float depth2 = cascadeCoords[index2].z;
    depth2 = beyondCascade3 ? saturate(depth2) : depth2;
    const float weight = beyondCascade2 ? saturate(cascadeBlends.y) - saturate(cascadeBlends.z) : 1.0f - saturate(cascadeBlends.x);
    const vec2 wh = 1.0f / vec2((float)csmRes, (float)csmRes);     // computeTexelSize, common.glsl:115-119
    float light1 = 0.0f, light2 = 0.0f, samples = 0.0f;
    const float bound = 1.0f;
    for (float y = -bound; y <= bound; y += 1.0f) {
        for (float x = -bound; x <= bound; x += 1.0f) {
            const vec2 o = vec2(x, y) * wh;                        // sampleShadowTexture, pbr.glsl:139-147
            light1 += shadowArrayPcf(csm, csmRes, shadowCoord1.x + o.x, shadowCoord1.y + o.y, (float)index1, depth1 - bias);
            light2 += shadowArrayPcf(csm, csmRes, shadowCoord2.x + o.x, shadowCoord2.y + o.y, (float)index2, depth2 - bias);
            ++samples;
        }
    }
    return mix(light2, light1, weight) * (1.0f / samples);
}

// ---- pbr.glsl:116-136 ------------------------------------------------------------------------
inline float calculateShadowValue1Sample(const void* const* atlases, SgiAtlasEntry entry, int cubeRes, float lightFarPlane,
                                         vec3 fragPos, vec3 lightPos, float lightNormalDotProduct, float minBias) {
    const vec3 fragDir = fragPos - lightPos;
    const float currentDepth = length(fragDir);
    const float bias = currentDepth * max(minBias, saturate(lightNormalDotProduct) * 0.03f);
    float shadow = 0.0f;
    float depth = shadowCubeFetch(atlases, entry, fragDir, cubeRes);
    depth = depth * lightFarPlane;
    if ((currentDepth - bias) > depth) shadow = 1.0f;
    return shadow;
}

// ---- pbr.glsl:260-267 ------------------------------------------------------------------------
inline float vplDiffuseAttenuation(vec3 lightDir, float lightRadius) {
    const float lightDist = length(lightDir);
    const float ratio = min(lightDist / lightRadius, 1.0f);
    const float minDist = mix(10.0f, 1.0f, exp(-3.0f * ratio)) * lightRadius;
    return 1.0f / (minDist + 1.0f * lightDist * lightDist);
}

// ---- pbr2.glsl --------------------------------------------------------------------------------
inline float NDF_GGX(float NdotH, float roughness) {                           // :51-55
    const float r2 = roughness * roughness;
    const float f = (NdotH * r2 - NdotH) * NdotH + 1.0f;
    return r2 / max(PI_F * f * f, PREVENT_DIV_BY_ZERO);
}
inline float Visibility_G_SmithGGX(float NdotV, float NdotL, float roughness) {   // :59-64
    const float a2 = roughness * roughness;
    const float ggxv = NdotL * sqrt((NdotV - a2 * NdotV) * NdotV + a2);
    const float ggxl = NdotV * sqrt((NdotL - a2 * NdotL) * NdotL + a2);
    return 0.5f / max(ggxv + ggxl, PREVENT_DIV_BY_ZERO);
}
inline vec3 Fr_Fresnel_Schlick(float u, vec3 f0) {                             // :73-76
    const float f = pow(1.0f - u, 5.0f);
    return f + f0 * (1.0f - f);
}
inline float RemapRoughness(const SgiFrameParams& fp, float roughness) {       // :139-144
    if (fp.usePerceptualRoughness) roughness = roughness * roughness;
    return max(max(0.0f, fp.minRoughness), roughness);
}
struct Dots { float NdotV, NdotL, NdotH, LdotH; };
inline Dots BRDF_DotProducts(vec3 viewDir, vec3 lightDir, vec3 normal) {       // :170-188
    const vec3 V = viewDir, L = normalize(lightDir), H = normalize(V + L), N = normal;
    Dots d;
    d.NdotV = saturate(dot(N, V)); d.NdotL = saturate(dot(N, L));
    d.NdotH = saturate(dot(N, H)); d.LdotH = saturate(dot(L, H));
    return d;
}

// calculateVirtualPointLighting2 (:613-657) -> calculateLighting_Lambert_SeparateDiffuseSpecular (:541-574)
// -> BRDF_Lambert_SeparateDiffuseSpecular (:382-452), with the constant arguments of the gather call site
// (vpl_pbr_gi.fs:195-197): shadowFactor 0, ambientIntensity 0, fogIntensity 0.
vec3 calculateVirtualPointLighting2(const SgiFrameParams& fp, vec3 fragPosition, vec3 baseColor, vec3 normal, vec3 viewDir,
                                    vec3 specularLightPos, vec3 diffuseLightPos, vec3 lightColor, float viewDist, float lightRadius,
                                    float roughness, float metallic, float ambientOcclusion, float shadowFactor, vec3 baseReflectance) {
    const vec3 specularLightDir = specularLightPos - fragPosition;
    const vec3 diffuseLightDir = diffuseLightPos - fragPosition;
    const float adjustedShadowFactor = 1.0f - shadowFactor;
    const float specularAttenuationFactor = vplDiffuseAttenuation(specularLightDir, lightRadius);   // vplSpecularAttenuation, pbr.glsl:269-271
    const float diffuseAttenuationFactor = vplDiffuseAttenuation(diffuseLightDir, lightRadius);
    const float ambientIntensity = 0.0f, fogIntensity = 0.0f;
    const vec3 diffuseDivisor = baseColor;
    const vec3 specularMultiplier = 1.0f / (baseColor + PREVENT_DIV_BY_ZERO);

    // BRDF_Base (:146-168)
    const float remappedRoughness = RemapRoughness(fp, roughness);
    vec3 diffuseColor = (1.0f - metallic) * baseColor;
    diffuseColor = diffuseColor / (diffuseDivisor + PREVENT_DIV_BY_ZERO);
    vec3 f0 = 0.16f * baseReflectance * baseReflectance * (1.0f - metallic) + baseColor * metallic;
    f0 = f0 * specularMultiplier;

    Dots d = BRDF_DotProducts(viewDir, specularLightDir, normal);
    // singleScatteringBRDF_Specular (:102-115)
    const float NDF = NDF_GGX(d.NdotH, remappedRoughness);
    const vec3 Fresnel = Fr_Fresnel_Schlick(d.LdotH, f0);
    const float Visibility = Visibility_G_SmithGGX(d.NdotV, d.NdotL, remappedRoughness);
    const vec3 Fr = ((NDF * Visibility) * Fresnel) * d.NdotL;

    d = BRDF_DotProducts(viewDir, diffuseLightDir, normal);
    const vec3 Fd = (diffuseColor * (1.0f / PI_F)) * d.NdotL;   // Fd_Lambert, :98-100

    vec3 brdfSpecular = specularAttenuationFactor * Fr;
    vec3 brdfDiffuse = diffuseAttenuationFactor * Fd;
    const vec3 brdf = brdfSpecular + brdfDiffuse;
    const vec3 ambient = brdf * ambientOcclusion * lightColor * ambientIntensity;
    const vec3 finalBrightness = brdf * lightColor;
    // applyFog, fog.glsl:37-41
    const float f = exp(-fp.fogDensity * viewDist);
    const vec3 fc = vec3(fp.fogColor[0], fp.fogColor[1], fp.fogColor[2]) * fogIntensity;
    const vec3 fogged = finalBrightness * f + fc * (1.0f - f);
    return ambient + adjustedShadowFactor * fogged;
}

inline void storeHalf4(uint16_t* dst, float r, float g, float b, float a) {
    dst[0] = f2h(r); dst[1] = f2h(g); dst[2] = f2h(b); dst[3] = f2h(a);
}

// bilinear weights with the 8 fractional bits GPUs use for filtering (keeps a tap that lands on a
// texel centre exact)
struct Bilin { int i0, i1, j0, j1; float a, b; };
inline Bilin bilinSetup(float u, float v, int W, int H) {
    const float fu = u * (float)W - 0.5f, fv = v * (float)H - 0.5f;
    const float i0f = floor(fu), j0f = floor(fv);
    Bilin r;
    r.a = floor((fu - i0f) * 256.0f + 0.5f) * (1.0f / 256.0f);
    r.b = floor((fv - j0f) * 256.0f + 0.5f) * (1.0f / 256.0f);
    r.i0 = clampi((int)i0f, 0, W - 1); r.i1 = clampi((int)i0f + 1, 0, W - 1);
    r.j0 = clampi((int)j0f, 0, H - 1); r.j1 = clampi((int)j0f + 1, 0, H - 1);
    return r;
}
inline float bilinMix(const Bilin& k, float t00, float t10, float t01, float t11) {
    return (1.0f - k.a) * (1.0f - k.b) * t00 + k.a * (1.0f - k.b) * t10 + (1.0f - k.a) * k.b * t01 + k.a * k.b * t11;
}

}  // namespace

extern "C" {

typedef struct OrcDims {
    int32_t width, height;     // rows held by the per-pixel planes (== the whole frame unless y_offset/full_height say otherwise)
    int32_t cube_res, csm_res;
    int32_t y_offset;          // first frame row of the planes; used for the bounded CPU-baseline samples of large frames
    int32_t full_height;       // 0 = height
} OrcDims;
static inline int orcFullH(const OrcDims* dm) { return dm->full_height > 0 ? dm->full_height : dm->height; }

int orc_num_threads(void) { return omp_get_max_threads(); }
uint16_t orc_f2h(float f) { return f2h(f); }
float orc_h2f(uint16_t h) { return h2f(h); }
uint32_t orc_hash3(float x, float y, float z) { return hash1(fbits(x) ^ hash1(fbits(y)) ^ hash1(fbits(z))); }
float orc_random3(float x, float y, float z) { return random3(vec3(x, y, z)); }
int orc_uses_glm(void) {
#ifdef ORACLE_USE_GLM
    return 1;
#else
    return 0;
#endif
}

// ---- test hooks: the restated pure-maths functions, callable one at a time (tests/test_reference_glsl.py holds them against the
// reference's own GLSL text compiled as C++, oracle/make_ref_glsl.py) ----
// a: fragPosition, baseColor, normal, viewDir, specularLightPos, diffuseLightPos, lightColor (7 x vec3), then viewDist, lightRadius,
//    roughness, metallic, ambientOcclusion, shadowFactor (6 floats), then baseReflectance (vec3): 30 floats
void orc_test_vpl_lighting(const SgiFrameParams* fp, const float* a, float* out3) {
    auto v = [&](int i) { return vec3(a[3 * i], a[3 * i + 1], a[3 * i + 2]); };
    const vec3 r = calculateVirtualPointLighting2(*fp, v(0), v(1), v(2), v(3), v(4), v(5), v(6), a[21], a[22], a[23], a[24], a[25], a[26],
                                                  vec3(a[27], a[28], a[29]));
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
void orc_test_world_pos(float u, float v, float depth, const float* invPV16, float* out3) {
    const vec3 r = worldPositionFromDepth(vec2(u, v), depth, loadMat(invPV16));
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
float orc_test_shadow1(const void* atlas, int res, int layer, float farPlane, const float* fragPos, const float* lightPos, float lightNormalDot, float minBias) {
    const void* atlases[SGI_MAX_TOTAL_SHADOW_ATLASES] = {atlas};
    SgiAtlasEntry e; e.index = 0; e.layer = layer;
    return calculateShadowValue1Sample(atlases, e, res, farPlane, vec3(fragPos[0], fragPos[1], fragPos[2]), vec3(lightPos[0], lightPos[1], lightPos[2]), lightNormalDot, minBias);
}
// calculateInfiniteShadowValue(vec4(pos, 1), cascadeBlends(pos), normal, false) as the cull calls it (viscull_vpls.cs:64-68)
float orc_test_sun_shadow(const SgiFrameParams* fp, const float* csm, int csmRes, const float* pos3, const float* normal3) {
    mat4 cpv[4]; for (int i = 0; i < 4; ++i) cpv[i] = loadMat(fp->cascadeProjViews[i]);
    const vec3 p(pos3[0], pos3[1], pos3[2]);
    vec3 blends;
    for (int i = 0; i < 3; ++i) {
        const vec4 plane(fp->cascadePlanes[i][0], fp->cascadePlanes[i][1], fp->cascadePlanes[i][2], fp->cascadePlanes[i][3]);
        const float b = dot(plane, vec4(p, 1.0f));
        if (i == 0) blends.x = b; else if (i == 1) blends.y = b; else blends.z = b;
    }
    return calculateInfiniteShadowValue(*fp, cpv, csm, csmRes, vec4(p, 1.0f), blends, vec3(normal3[0], normal3[1], normal3[2]));
}
float orc_test_attenuation(const float* dir3, float radius) { return vplDiffuseAttenuation(vec3(dir3[0], dir3[1], dir3[2]), radius); }

// a1: viscull_vpls.cs:47-106.  Canonical order = ascending original index, which is one legal
// interleaving of the reference's 8 racing helper threads and what its serial variant (:108-119) did.
int orc_cull(const SgiFrameParams* fp, const OrcDims* dm, const SgiVplData* lightData, const SgiAtlasEntry* atlasByVpl,
             const float* csm, SgiVplData* updatedLightData, int32_t* vplVisibleIndex, SgiAtlasEntry* shadowDiffuseIndices) {
    const int N = fp->totalNumLights;
    mat4 cpv[4]; for (int i = 0; i < 4; ++i) cpv[i] = loadMat(fp->cascadeProjViews[i]);
    vec4 planes[3]; for (int i = 0; i < 3; ++i) planes[i] = vec4(fp->cascadePlanes[i][0], fp->cascadePlanes[i][1], fp->cascadePlanes[i][2], fp->cascadePlanes[i][3]);
    const vec3 sun(fp->infiniteLightDirection[0], fp->infiniteLightDirection[1], fp->infiniteLightDirection[2]);
    std::vector<uint8_t> lightVisible(N > 0 ? N : 1, 0);
    #pragma omp parallel for schedule(static)
    for (int index = 0; index < N; ++index) {
        const vec3 lightPos(lightData[index].position[0], lightData[index].position[1], lightData[index].position[2]);
        const vec3 cascadeBlends(dot(planes[0], vec4(lightPos, 1.0f)), dot(planes[1], vec4(lightPos, 1.0f)), dot(planes[2], vec4(lightPos, 1.0f)));
        const float shadowFactor = 1.0f - calculateInfiniteShadowValue(*fp, cpv, csm, dm->csm_res, vec4(lightPos, 1.0f), cascadeBlends, sun);
        if (shadowFactor < 1.0f) lightVisible[index] = 1;
    }
    int count = 0;
    for (int index = 0; index < N; ++index) {
        if (!lightVisible[index]) continue;
        const int localIndex = count++;
        if (localIndex >= SGI_MAX_TOTAL_VPLS_PER_FRAME) continue;     // keep counting, stop writing (see note in DESIGN.md)
        updatedLightData[localIndex] = lightData[index];
        vplVisibleIndex[localIndex + 1] = index;
        if (shadowDiffuseIndices && atlasByVpl) shadowDiffuseIndices[localIndex] = atlasByVpl[index];
    }
    vplVisibleIndex[0] = count > SGI_MAX_TOTAL_VPLS_PER_FRAME ? SGI_MAX_TOTAL_VPLS_PER_FRAME : count;
    return 0;
}

// f1: the virtual-point-light branch of UpdatePointLights_ (StratusRendererBackend.cpp:1359-1392)
// and InitVplFrameData_ (:1281-1292).  The reference inserts VplDistKey_(light, distance) into a std::multiset ordered by the
// distance alone (StratusRendererBackend.h:602-623) and copies its first MAX_TOTAL_VPLS_BEFORE_CULLING entries.
int orc_vpl_prepass(const SgiLightRecord* lights, int n, const float* viewPosition, const float* frustumPlanes /* [6][4] */,
                    SgiVplData* vplDataOut, SgiAtlasEntry* atlasEntryOut, int32_t* orderOut) {
    struct Key { int light; double distance; };
    struct Less { bool operator()(const Key& l, const Key& r) const { return l.distance < r.distance; } };
    std::multiset<Key, Less> perVPLDistToViewerSet;
    const vec3 c(viewPosition[0], viewPosition[1], viewPosition[2]);
    for (int i = 0; i < n; ++i) {
        const vec3 position(lights[i].position[0], lights[i].position[1], lights[i].position[2]);
        const double distance = length(position - c);                        // glm::distance(p0, p1) = length(p1 - p0)
        bool inFrustum = true;                                               // IsSphereInFrustum, StratusMath.h:401-412
        for (int k = 0; k < 6; ++k) {
            const vec4 g(frustumPlanes[4 * k], frustumPlanes[4 * k + 1], frustumPlanes[4 * k + 2], frustumPlanes[4 * k + 3]);
            const float d = dot(g, vec4(position, 1.0f));
            if (d < -lights[i].radius) { inFrustum = false; break; }
        }
        if (inFrustum) perVPLDistToViewerSet.insert(Key{i, distance});
    }
    int count = 0;
    for (const Key& entry : perVPLDistToViewerSet) {
        const SgiLightRecord& L = lights[entry.light];
        SgiVplData data;
        std::memset(&data, 0, sizeof(data));
        data.position[0] = L.position[0]; data.position[1] = L.position[1]; data.position[2] = L.position[2]; data.position[3] = 1.0f;
        data.farPlane = L.farPlane; data.radius = L.radius; data.intensity = L.intensity;
        vplDataOut[count] = data;
        if (atlasEntryOut) { atlasEntryOut[count].index = L.atlasIndex; atlasEntryOut[count].layer = L.atlasLayer; }
        if (orderOut) orderOut[count] = entry.light;
        if (++count >= SGI_MAX_TOTAL_VPLS_BEFORE_CULLING) break;              // `if (size() >= capacity()) break;` (:1387)
    }
    return count;
}

// a2: vpl_light_color.cs:41-76
int orc_color(const SgiFrameParams* fp, const OrcDims* dm, SgiVplData* lightData, const int32_t* numVisible,
              const SgiAtlasEntry* diffuseIndices, const void* const* shadowCubeMaps, const void* const* diffuseCubeMaps) {
    const float colorMultiplier = 50000.0f;
    const int visibleVpls = numVisible[0];
    const vec3 infiniteLightDirection(fp->infiniteLightDirection[0], fp->infiniteLightDirection[1], fp->infiniteLightDirection[2]);
    const vec3 infiniteLightColor(fp->infiniteLightColor[0], fp->infiniteLightColor[1], fp->infiniteLightColor[2]);
    const int res = dm->cube_res;
    for (int index = 0; index < visibleVpls; ++index) {
        SgiVplData data = lightData[index];
        const SgiAtlasEntry entry = diffuseIndices[index];
        const int texel = cubeTexel(-infiniteLightDirection, res);
        const uint8_t* dtex = (const uint8_t*)diffuseCubeMaps[entry.index] + ((size_t)entry.layer * 6 * res * res + texel) * 4;
        const vec3 color = vec3(unorm8(dtex[0]), unorm8(dtex[1]), unorm8(dtex[2])) * infiniteLightColor;
        const float magnitude = data.radius * shadowCubeFetch(shadowCubeMaps, entry, -infiniteLightDirection, res);
        const float totalColors = 1.0f;
        const vec3 pos(data.position[0], data.position[1], data.position[2]);
        const vec3 specularPos = pos - 1.1f * magnitude * infiniteLightDirection;
        const vec3 c = color / totalColors * data.intensity * colorMultiplier;
        lightData[index].color[0] = c.x; lightData[index].color[1] = c.y; lightData[index].color[2] = c.z; lightData[index].color[3] = 1.0f;
        lightData[index].specularPosition[0] = specularPos.x; lightData[index].specularPosition[1] = specularPos.y;
        lightData[index].specularPosition[2] = specularPos.z; lightData[index].specularPosition[3] = 1.0f;
    }
    return 0;
}

// a3: vpl_tiled_deferred_culling_stage1.cs:50-90
int orc_stage1(const SgiFrameParams* fp, const OrcDims* dm, const float* gDepth, const uint8_t* gNormal4,
               SgiVplStage1PerTileOutputs* out) {
    const int W = dm->width, H = dm->height;
    const int tilesX = W / SGI_TILE_DIVISOR, tilesY = H / SGI_TILE_DIVISOR;
    const int fullTilesY = orcFullH(dm) / SGI_TILE_DIVISOR, yOff = dm->y_offset;
    const mat4 invPV = loadMat(fp->invProjectionView);
    const float localWorkGroupSize = 25.0f;
    #pragma omp parallel for schedule(static)
    for (int ty = 0; ty < tilesY; ++ty) for (int tx = 0; tx < tilesX; ++tx) {
        vec3 averageLocalPosition(0.0f), averageLocalNormal(0.0f);
        for (int i = 0; i < 25; ++i) {                // gl_LocalInvocationIndex = ly*5 + lx
            const int px = tx * 5 + (i % 5), py = ty * 5 + (i / 5);
            const vec2 texCoords = (vec2((float)px, (float)(py + yOff)) + vec2(0.5f)) / vec2((float)(tilesX * 5), (float)(fullTilesY * 5));
            const float depth = gDepth[(size_t)py * W + px];
            const vec3 fragPos = worldPositionFromDepth(texCoords, depth, invPV);
            const vec3 normal = decodeNormal(gNormal4 + ((size_t)py * W + px) * 4);
            averageLocalPosition += fragPos;
            averageLocalNormal += normal;
        }
        averageLocalPosition /= localWorkGroupSize;
        averageLocalNormal /= localWorkGroupSize;
        averageLocalNormal = normalize(averageLocalNormal);
        SgiVplStage1PerTileOutputs& o = out[tx + ty * tilesX];
        o.averageLocalPosition[0] = averageLocalPosition.x; o.averageLocalPosition[1] = averageLocalPosition.y;
        o.averageLocalPosition[2] = averageLocalPosition.z; o.averageLocalPosition[3] = 1.0f;
        o.averageLocalNormal[0] = averageLocalNormal.x; o.averageLocalNormal[1] = averageLocalNormal.y;
        o.averageLocalNormal[2] = averageLocalNormal.z; o.averageLocalNormal[3] = 1.0f;
    }
    return 0;
}

// a4: vpl_tiled_deferred_culling_stage2.cs:61-129.  `lightData` is indexed THROUGH the visible list
// (:81-82), i.e. it is the un-compacted table; `vplVisibleIndex` here points at the first index
// (the shader predates the count-in-slot-0 layout, so the caller passes list+1).
int orc_stage2(const OrcDims* dm, const SgiVplStage1PerTileOutputs* stage1Data, const SgiVplData* lightData,
               int32_t numVisible, const int32_t* vplVisibleIndex, SgiVplStage2PerTileOutputs* out) {
    const int numTiles = (dm->width / SGI_TILE_DIVISOR) * (dm->height / SGI_TILE_DIVISOR);
    #pragma omp parallel for schedule(static)
    for (int tileIndex = 0; tileIndex < numTiles; ++tileIndex) {
        const vec3 fragPos(stage1Data[tileIndex].averageLocalPosition[0], stage1Data[tileIndex].averageLocalPosition[1], stage1Data[tileIndex].averageLocalPosition[2]);
        const vec3 normal(stage1Data[tileIndex].averageLocalNormal[0], stage1Data[tileIndex].averageLocalNormal[1], stage1Data[tileIndex].averageLocalNormal[2]);
        int numVisibleThisTile = 0;
        int indicesVisibleThisTile[SGI_MAX_VPLS_PER_TILE];
        float distancesVisibleThisTile[SGI_MAX_VPLS_PER_TILE];
        for (int i = 0; i < SGI_MAX_VPLS_PER_TILE; ++i) { indicesVisibleThisTile[i] = i; distancesVisibleThisTile[i] = FLOAT_MAX_; }
        for (int i = 0; i < numVisible; ++i) {
            const int lightIndex = vplVisibleIndex[i];
            const vec3 lightPosition(lightData[lightIndex].position[0], lightData[lightIndex].position[1], lightData[lightIndex].position[2]);
            const vec3 lightMinusFrag = lightPosition - fragPos;
            const float sideCheck = dot(normal, lightMinusFrag);
            if (sideCheck < 0) continue;
            const float radius = lightData[lightIndex].radius;
            float distance = length(lightMinusFrag);
            const float ratio = distance / radius;
            if (ratio > 1.0f) continue;
            distance = ratio;
            for (int ii = 0; ii < SGI_MAX_VPLS_PER_TILE; ++ii) {
                if (distance < distancesVisibleThisTile[ii]) {
                    if (distancesVisibleThisTile[ii] != FLOAT_MAX_) {
                        // SHUFFLE_DOWN (:55-59)
                        for (int index_ = SGI_MAX_VPLS_PER_TILE - 2; index_ >= ii; --index_) {
                            indicesVisibleThisTile[index_ + 1] = indicesVisibleThisTile[index_];
                            distancesVisibleThisTile[index_ + 1] = distancesVisibleThisTile[index_];
                        }
                    }
                    indicesVisibleThisTile[ii] = lightIndex;
                    distancesVisibleThisTile[ii] = distance;
                    if (numVisibleThisTile < SGI_MAX_VPLS_PER_TILE) ++numVisibleThisTile;
                    break;
                }
            }
        }
        out[tileIndex].numVisible = numVisibleThisTile;
        for (int i = 0; i < numVisibleThisTile; ++i) out[tileIndex].indices[i] = indicesVisibleThisTile[i];
    }
    return 0;
}

// a5: vpl_pbr_gi.fs:81-207.  Outputs: color (attachment 0, RGB16F) and reservoir (attachment 1, RGBA16F),
// both as half4 planes.  `sampleStats`, if given, receives [0] = executed loop iterations (S·P),
// [1] = shaded samples, [2] = cube-shadow fetches.
static int gatherImpl(const SgiFrameParams* fp, const OrcDims* dm, const float* gDepth, const uint8_t* gNormal4, const uint8_t* gAlbedo4,
               const uint8_t* gBaseReflectivity2, const uint8_t* gRoughnessMetallicAmbient4, const uint16_t* ssaoH,
               const SgiVplData* lightData, const int32_t* numVisible, const SgiAtlasEntry* shadowIndices,
               const void* const* shadowCubeMaps, uint16_t* colorOutH4, uint16_t* reservoirOutH4, int64_t* sampleStats,
               const SgiVplStage2PerTileOutputs* tileLists, const int32_t* visibleList) {
    const int W = dm->width, H = dm->height, cubeRes = dm->cube_res;
    const int fullH = orcFullH(dm), yOff = dm->y_offset;
    const mat4 invPV = loadMat(fp->invProjectionView);
    const vec3 viewPosition(fp->viewPosition[0], fp->viewPosition[1], fp->viewPosition[2]);
    const float time = fp->time;
    const float minGiOcclusionFactor = fp->minGiOcclusionFactor;
    const int V = numVisible[0];
    int64_t itersTotal = 0, shadedTotal = 0, fetchTotal = 0;
    #pragma omp parallel for schedule(dynamic, 4) reduction(+ : itersTotal, shadedTotal, fetchTotal)
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        const size_t p = (size_t)y * W + x;
        const vec2 pixelCoords((float)x + 0.5f, (float)(y + yOff) + 0.5f);            // gl_FragCoord.xy
        const vec2 texCoords(((float)x + 0.5f) / (float)W, ((float)(y + yOff) + 0.5f) / (float)fullH);   // fsTexCoords (Appendix A1)
        const float depth = gDepth[p];
        const vec3 fragPos = worldPositionFromDepth(texCoords, depth, invPV);
        const vec3 viewDir = normalize(viewPosition - fragPos);
        const float distToCamera = length(viewPosition - fragPos);
        vec3 baseColor(unorm8(gAlbedo4[p * 4]), unorm8(gAlbedo4[p * 4 + 1]), unorm8(gAlbedo4[p * 4 + 2]));
        baseColor = vec3(max(baseColor.x, 0.01f), max(baseColor.y, 0.01f), max(baseColor.z, 0.01f));
        const vec3 normal = decodeNormal(gNormal4 + p * 4);
        float roughness = unorm8(gRoughnessMetallicAmbient4[p * 4]);
        roughness = max(0.5f, roughness);
        const float metallic = unorm8(gRoughnessMetallicAmbient4[p * 4 + 1]);
        const float ambientOcclusion = clamp(h2f(ssaoH[p]), 0.35f, 1.0f);
        const float ambient = unorm8(gRoughnessMetallicAmbient4[p * 4 + 2]);
        const vec3 baseReflectivity = vec3(unorm8(gBaseReflectivity2[p * 2]));
        const float roughnessWeight = 1.0f - roughness;
        vec3 vplColor(0.0f);
        const float seedZOffset = 10000.0f;
        vec3 seed(pixelCoords.x, pixelCoords.y, time);
        float validSamples = 0.0f;
        const float distRatioToCamera = min(1.0f - distToCamera / 1000.0f, 1.0f);
        const int maxSamplesPerPixel = (int)mix(2.0f, 10.0f, roughnessWeight);
        int samplesMax = maxSamplesPerPixel;
        samplesMax = max(1, (int)((float)samplesMax * distRatioToCamera));
        const int sampleCount = samplesMax;
        const int maxRandomIndex = V - 1;
        if (tileLists != nullptr) {
            // SGI_GATHER_TILE_LIST (include/stratusgi.h): the dormant tile lookup of vpl_pbr_gi.fs:86-97 composed with the live
            // per-light body (:153-198).  Every VPL of the pixel's tile list in list order; a VPL behind the pixel or out of
            // range is skipped.  The list holds indices into the ORIGINAL table (stage2.cs:81-82); the coloured, compacted
            // table is indexed by position in the visible list, so the index is mapped back through that list.
            const int tilesX = W / SGI_TILE_DIVISOR;
            const SgiVplStage2PerTileOutputs& tl = tileLists[(size_t)(y / SGI_TILE_DIVISOR) * tilesX + x / SGI_TILE_DIVISOR];
            const int count = tl.numVisible < SGI_MAX_VPLS_PER_TILE ? tl.numVisible : SGI_MAX_VPLS_PER_TILE;
            for (int l = 0; l < count; ++l) {
                int lightIndex = -1;
                for (int k = 0; k < V; ++k) if (visibleList[1 + k] == tl.indices[l]) { lightIndex = k; break; }
                if (lightIndex < 0) continue;
                ++itersTotal;
                const SgiAtlasEntry entry = shadowIndices[lightIndex];
                const vec3 lightPosition(lightData[lightIndex].position[0], lightData[lightIndex].position[1], lightData[lightIndex].position[2]);
                const vec3 specularLightPosition(lightData[lightIndex].specularPosition[0], lightData[lightIndex].specularPosition[1], lightData[lightIndex].specularPosition[2]);
                const vec3 lightMinusFrag = lightPosition - fragPos;
                const float lightRadius = lightData[lightIndex].radius;
                const float distance = length(lightMinusFrag);
                const float sideCheck = dot(normal, normalize(lightMinusFrag));
                if (sideCheck < 0.0f || distance > lightRadius) continue;
                const float distanceRatio = clamp((2.0f * distance) / lightRadius, 0.0f, 1.0f);
                const float distAttenuation = distanceRatio;
                const vec3 lightColor(lightData[lightIndex].color[0], lightData[lightIndex].color[1], lightData[lightIndex].color[2]);
                float shadowFactor = 0.0f;
                if (distToCamera < 500.0f) {
                    ++fetchTotal;
                    shadowFactor = calculateShadowValue1Sample(shadowCubeMaps, entry, cubeRes, lightData[lightIndex].farPlane, fragPos, lightPosition,
                                                               dot(lightPosition - fragPos, normal), 0.1f);
                }
                shadowFactor = min(shadowFactor, mix(minGiOcclusionFactor, 1.0f, distanceRatio));
                float reweightingFactor = 1.0f;
                if (shadowFactor > 0.0f) reweightingFactor = (1.0f - distAttenuation) * minGiOcclusionFactor + distAttenuation;
                validSamples += reweightingFactor;
                ++shadedTotal;
                const vec3 tmpColor = ambientOcclusion * calculateVirtualPointLighting2(*fp, fragPos, baseColor, normal, viewDir, specularLightPosition,
                    lightPosition, lightColor, distToCamera, lightRadius, roughness, metallic, ambient, 0.0f, baseReflectivity);
                vplColor = vplColor + (1.0f - shadowFactor) * tmpColor;
            }
        } else
        // V <= 0: the reference would read stale lights (undefined); defined here as "no light".
        for (int i = 0, resamples = 0, count = 0; V > 0 && i < sampleCount; i += 1, count += 1) {
            ++itersTotal;
            const float rand = random3(seed);
            seed.z += seedZOffset;
            const int lightIndex = (int)((float)maxRandomIndex * rand);
            const SgiAtlasEntry entry = shadowIndices[lightIndex];
            const vec3 lightPosition(lightData[lightIndex].position[0], lightData[lightIndex].position[1], lightData[lightIndex].position[2]);
            const vec3 specularLightPosition(lightData[lightIndex].specularPosition[0], lightData[lightIndex].specularPosition[1], lightData[lightIndex].specularPosition[2]);
            const vec3 lightMinusFrag = lightPosition - fragPos;
            const float lightRadius = lightData[lightIndex].radius;
            const float distance = length(lightMinusFrag);
            if (resamples < 5) {
                const float sideCheck = dot(normal, normalize(lightMinusFrag));
                if (sideCheck < 0.0f || distance > lightRadius) { ++resamples; --i; continue; }
            }
            const float distanceRatio = clamp((2.0f * distance) / lightRadius, 0.0f, 1.0f);
            const float distAttenuation = distanceRatio;
            const vec3 lightColor(lightData[lightIndex].color[0], lightData[lightIndex].color[1], lightData[lightIndex].color[2]);
            float shadowFactor = 0.0f;
            if (distToCamera < 500.0f) {
                ++fetchTotal;
                shadowFactor = calculateShadowValue1Sample(shadowCubeMaps, entry, cubeRes, lightData[lightIndex].farPlane, fragPos, lightPosition,
                                                           dot(lightPosition - fragPos, normal), 0.1f);
            }
            shadowFactor = min(shadowFactor, mix(minGiOcclusionFactor, 1.0f, distanceRatio));
            float reweightingFactor = 1.0f;
            if (shadowFactor > 0.0f) reweightingFactor = (1.0f - distAttenuation) * minGiOcclusionFactor + distAttenuation;
            validSamples += reweightingFactor;
            ++shadedTotal;
            const vec3 tmpColor = ambientOcclusion * calculateVirtualPointLighting2(*fp, fragPos, baseColor, normal, viewDir, specularLightPosition,
                lightPosition, lightColor, distToCamera, lightRadius, roughness, metallic, ambient, 0.0f, baseReflectivity);
            vplColor = vplColor + (1.0f - shadowFactor) * tmpColor;
        }
        validSamples = max(validSamples, 1.0f);
        const vec3 color = baseColor + PREVENT_DIV_BY_ZERO;
        const vec3 bounded = min(vplColor, HALF_FLOAT_MAX);     // boundHDR, common.glsl:89-92
        storeHalf4(colorOutH4 + p * 4, color.x, color.y, color.z, 1.0f);
        storeHalf4(reservoirOutH4 + p * 4, bounded.x, bounded.y, bounded.z, validSamples);
    }
    if (sampleStats) { sampleStats[0] = itersTotal; sampleStats[1] = shadedTotal; sampleStats[2] = fetchTotal; }
    return 0;
}

int orc_gather(const SgiFrameParams* fp, const OrcDims* dm, const float* gDepth, const uint8_t* gNormal4, const uint8_t* gAlbedo4,
               const uint8_t* gBaseReflectivity2, const uint8_t* gRoughnessMetallicAmbient4, const uint16_t* ssaoH,
               const SgiVplData* lightData, const int32_t* numVisible, const SgiAtlasEntry* shadowIndices,
               const void* const* shadowCubeMaps, uint16_t* colorOutH4, uint16_t* reservoirOutH4, int64_t* sampleStats) {
    return gatherImpl(fp, dm, gDepth, gNormal4, gAlbedo4, gBaseReflectivity2, gRoughnessMetallicAmbient4, ssaoH, lightData, numVisible, shadowIndices,
                      shadowCubeMaps, colorOutH4, reservoirOutH4, sampleStats, nullptr, nullptr);
}

// tile-list mode of the gather: `numVisible` is the whole visible list ([0] = count), `tileLists` the stage-2 output
int orc_gather_tiles(const SgiFrameParams* fp, const OrcDims* dm, const float* gDepth, const uint8_t* gNormal4, const uint8_t* gAlbedo4,
                     const uint8_t* gBaseReflectivity2, const uint8_t* gRoughnessMetallicAmbient4, const uint16_t* ssaoH,
                     const SgiVplData* lightData, const int32_t* numVisible, const SgiAtlasEntry* shadowIndices,
                     const void* const* shadowCubeMaps, uint16_t* colorOutH4, uint16_t* reservoirOutH4, int64_t* sampleStats,
                     const SgiVplStage2PerTileOutputs* tileLists) {
    return gatherImpl(fp, dm, gDepth, gNormal4, gAlbedo4, gBaseReflectivity2, gRoughnessMetallicAmbient4, ssaoH, lightData, numVisible, shadowIndices,
                      shadowCubeMaps, colorOutH4, reservoirOutH4, sampleStats, tileLists, numVisible);
}

// a6: one pass of vpl_pbr_gi_denoise.fs:275-435.
typedef struct OrcDenoisePass {
    int32_t multiplier;        // uniform multiplier = 2^i - 1 (BE:1920)
    int32_t mergeReservoirs;   // pass < maxReservoirMergingPasses (BE:1927)
    int32_t final_;            // last pass (BE:1929-1931)
    int32_t numVisible;        // numVisible[0]
    const uint16_t* screenH4; const uint16_t* velocityH2; const uint8_t* normal4; const float* ids; const float* depth;
    const uint8_t* prevNormal4; const float* prevIds; const float* prevDepth;
    const uint16_t* indirectIlluminationH4; const uint16_t* indirectShadowsH4; const uint16_t* prevIndirectIlluminationH4;
    const uint16_t* historyDepthH;
    uint16_t* combinedOutH4; uint16_t* giOutH4; uint16_t* reservoirOutH4; uint16_t* historyOutH;
} OrcDenoisePass;

int orc_denoise_pass(const SgiFrameParams* fp, const OrcDims* dm, const OrcDenoisePass* ps) {
    const int W = dm->width, H = dm->height;
    const int fullH = orcFullH(dm), yOff = dm->y_offset;
    const float time = fp->time;
    const float framesPerSecond = (float)(1.0 / fp->deltaSeconds);
    const int multiplier = ps->multiplier;
    const float sigmaZ = 1.0f, sigmaN = 128.0f;
    const int dminmax = 2;
    #pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        const size_t p = (size_t)y * W + x;
        const vec2 fsTexCoords(((float)x + 0.5f) / (float)W, ((float)(y + yOff) + 0.5f) / (float)fullH);
        const vec3 screenColor(h2f(ps->screenH4[p * 4]), h2f(ps->screenH4[p * 4 + 1]), h2f(ps->screenH4[p * 4 + 2]));
        const vec2 velocityVal(h2f(ps->velocityH2[p * 2]), h2f(ps->velocityH2[p * 2 + 1]));
        const vec2 prevTexCoords = fsTexCoords - velocityVal;
        const vec3 centerIllum(h2f(ps->indirectIlluminationH4[p * 4]), h2f(ps->indirectIlluminationH4[p * 4 + 1]), h2f(ps->indirectIlluminationH4[p * 4 + 2]));
        const vec3 centerNormal = decodeNormal(ps->normal4 + p * 4);
        const float centerDepth = ps->depth[p];
        float historyAccum = h2f(ps->historyDepthH[p]);

        vec4 reservoirFiltered(0.0f);
        if (ps->mergeReservoirs) {
            // computeMergedReservoir (:192-273) with minmaxNearest = 0: the centre texel only
            vec3 seed((float)x + 0.5f, (float)(y + yOff) + 0.5f, time);
            vec4 centerReservoir(0.0f);
            const float depthCutoff = 0.1f * centerDepth;
            float runningSum = 0.0f;
            const float probabilisticWeight = 1.0f / (float)ps->numVisible;
            do {
                const vec3 currNormal = centerNormal;                                   // offset (0,0)
                if (dot(centerNormal, currNormal) < 0.906f) break;
                const float currDepth = centerDepth;
                if (abs(currDepth - centerDepth) > depthCutoff) break;
                const vec4 currReservoir(h2f(ps->indirectShadowsH4[p * 4]), h2f(ps->indirectShadowsH4[p * 4 + 1]),
                                         h2f(ps->indirectShadowsH4[p * 4 + 2]), h2f(ps->indirectShadowsH4[p * 4 + 3]));
                const float randUpdate = random3(seed);
                seed.z += 10000.0f;
                const float probability_ = currReservoir.w * probabilisticWeight;
                centerReservoir.w += probability_;
                if (randUpdate < (probability_ / centerReservoir.w)) {
                    centerReservoir.x += currReservoir.x; centerReservoir.y += currReservoir.y; centerReservoir.z += currReservoir.z;
                    runningSum += currReservoir.w;
                }
            } while (false);
            centerReservoir.w = runningSum;
            reservoirFiltered = centerReservoir;
        } else {
            for (int dx = -dminmax; dx <= dminmax; ++dx) {
                for (int dy = -dminmax; dy <= dminmax; ++dy) {
                    const int tx = clampi(x + dx + dx * multiplier, 0, W - 1), ty = clampi(y + dy + dy * multiplier, 0, H - 1);   // CLAMP_TO_EDGE
                    const size_t q = (size_t)ty * W + tx;
                    const vec4 reservoir(h2f(ps->indirectShadowsH4[q * 4]), h2f(ps->indirectShadowsH4[q * 4 + 1]),
                                         h2f(ps->indirectShadowsH4[q * 4 + 2]), h2f(ps->indirectShadowsH4[q * 4 + 3]));
                    float filtered;                                                     // filterInput (:142-190)
                    if (dx == 0 && dy == 0) filtered = 1.0f;
                    else {
                        const float currDepth = ps->depth[q];
                        const float wz = exp(-abs(centerDepth - currDepth)) / sigmaZ;
                        const vec3 currNormal = decodeNormal(ps->normal4 + q * 4);
                        float wn = max(0.0f, dot(centerNormal, currNormal));
                        wn = pow(wn, sigmaN);
                        filtered = wn * wz * 1.0f;
                    }
                    reservoirFiltered += filtered * reservoir;
                }
            }
        }
        const vec3 shadowFactor = vec3(reservoirFiltered) / max(PREVENT_DIV_BY_ZERO, reservoirFiltered.w);
        const vec3 gi = centerIllum;
        vec3 illumAvg = gi;
        if (ps->final_) {
            float accumMultiplier = 1.0f;
            const vec3 currGi = shadowFactor;
            Bilin k = bilinSetup(prevTexCoords.x, prevTexCoords.y, W, fullH);
            k.j0 = clampi(k.j0 - yOff, 0, H - 1); k.j1 = clampi(k.j1 - yOff, 0, H - 1);
            const size_t q00 = (size_t)k.j0 * W + k.i0, q10 = (size_t)k.j0 * W + k.i1, q01 = (size_t)k.j1 * W + k.i0, q11 = (size_t)k.j1 * W + k.i1;
            const float prevCenterDepth = bilinMix(k, ps->prevDepth[q00], ps->prevDepth[q10], ps->prevDepth[q01], ps->prevDepth[q11]);
            vec3 pn;
            for (int c = 0; c < 3; ++c)
                pn[c] = bilinMix(k, unorm8(ps->prevNormal4[q00 * 4 + c]), unorm8(ps->prevNormal4[q10 * 4 + c]), unorm8(ps->prevNormal4[q01 * 4 + c]), unorm8(ps->prevNormal4[q11 * 4 + c]));
            const vec3 prevCenterNormal = normalize(pn * 2.0f - vec3(1.0f));
            vec3 prevGi;
            for (int c = 0; c < 3; ++c)
                prevGi[c] = bilinMix(k, h2f(ps->prevIndirectIlluminationH4[q00 * 4 + c]), h2f(ps->prevIndirectIlluminationH4[q10 * 4 + c]),
                                     h2f(ps->prevIndirectIlluminationH4[q01 * 4 + c]), h2f(ps->prevIndirectIlluminationH4[q11 * 4 + c]));
            const float currId = ps->ids[p];
            // prevIds is NEAREST (BE:418)
            const int ni = clampi((int)floor(prevTexCoords.x * (float)W), 0, W - 1);
            const int nj = clampi(clampi((int)floor(prevTexCoords.y * (float)fullH), 0, fullH - 1) - yOff, 0, H - 1);
            const float prevId = ps->prevIds[(size_t)nj * W + ni];
            const float wn = max(0.0f, dot(centerNormal, prevCenterNormal));
            const float wz = exp(-abs(centerDepth - prevCenterDepth));
            const float wid = currId != prevId ? 0.0f : 1.0f;
            float similarity = wn * wz * wid;
            if (similarity < 0.99f) { similarity = 0.0f; accumMultiplier = 0.0f; }
            historyAccum = min(1.0f + historyAccum * accumMultiplier, framesPerSecond);
            const float maxAccumulationFactor = 1.0f / historyAccum;
            illumAvg = mix(prevGi, currGi, maxAccumulationFactor);
        }
        const vec3 combined = screenColor + gi * illumAvg;
        storeHalf4(ps->combinedOutH4 + p * 4, combined.x, combined.y, combined.z, 1.0f);
        storeHalf4(ps->giOutH4 + p * 4, illumAvg.x, illumAvg.y, illumAvg.z, 1.0f);
        storeHalf4(ps->reservoirOutH4 + p * 4, shadowFactor.x, shadowFactor.y, shadowFactor.z, 1.0f);
        ps->historyOutH[p] = f2h(historyAccum);
    }
    return 0;
}

// ---- widening row f2: VPL cube shadow / diffuse map production (BE:1426-1575, shadowVpl.fs:16-36) ------------------------------
// The reference rasterises the scene's meshes six times per light; here the scene is a list of axis-aligned boxes and the fragment
// of a texel is found by casting the ray through the texel centre (inverse of the cube-face table).  That
// part has no reference text (a rasteriser is fixed function); everything from the fragment on is shadowVpl.fs:
//     lightDistance = length(fsPosition.xyz - lightPos);  lightDistance = saturate(lightDistance / farPlane);  (:27-30)
//     gl_FragDepth = lightDistance;  color = baseColor.rgb;                                                      (:33-35)
// stored as unorm16 depth (round to nearest) and unorm8 colour.  Cleared texels (no hit): depth 1, colour 0 (BE:1452-1456).
inline vec3 cubeTexelDirection(int face, float sc, float tc) {
    switch (face) {       // inverse of GL 4.6 table 8.19
        case 0: return vec3(1.0f, -tc, -sc);
        case 1: return vec3(-1.0f, -tc, sc);
        case 2: return vec3(sc, 1.0f, tc);
        case 3: return vec3(sc, -1.0f, -tc);
        case 4: return vec3(sc, -tc, 1.0f);
        default: return vec3(-sc, -tc, -1.0f);
    }
}
inline float shadowVplDepth(vec3 fsPosition, vec3 lightPos, float farPlane) {      // shadowVpl.fs:27-30
    float lightDistance = length(fsPosition - lightPos);
    lightDistance = saturate(lightDistance / farPlane);
    return lightDistance;
}
float orc_shadow_vpl_depth(const float* fsPosition3, const float* lightPos3, float farPlane) {
    return shadowVplDepth(vec3(fsPosition3[0], fsPosition3[1], fsPosition3[2]), vec3(lightPos3[0], lightPos3[1], lightPos3[2]), farPlane);
}
int orc_render_vpl_cubemaps(const SgiSceneBox* boxes, int numBoxes, const SgiVplData* lightData, const SgiAtlasEntry* atlasByVpl,
                            const int32_t* vplIndices, int numMaps, int res, void* const* shadowAtlases, void* const* diffuseAtlases) {
    const float fres = (float)res;
    #pragma omp parallel for schedule(dynamic, 1)
    for (int m = 0; m < numMaps * 6; ++m) {
        const int k = vplIndices[m / 6], face = m % 6;
        const SgiVplData& v = lightData[k];
        const SgiAtlasEntry e = atlasByVpl[k];
        const vec3 o(v.position[0], v.position[1], v.position[2]);
        uint16_t* depth = (uint16_t*)shadowAtlases[e.index] + ((size_t)e.layer * 6 + face) * res * res;
        uint8_t* diffuse = (uint8_t*)diffuseAtlases[e.index] + ((size_t)e.layer * 6 + face) * res * res * 4;
        for (int j = 0; j < res; ++j) for (int i = 0; i < res; ++i) {
            const float sc = 2.0f * (((float)i + 0.5f) / fres) - 1.0f, tc = 2.0f * (((float)j + 0.5f) / fres) - 1.0f;
            const vec3 d = cubeTexelDirection(face, sc, tc);
            float best = 3.402823466e+38f; int hit = -1;
            for (int b = 0; b < numBoxes; ++b) {
                const SgiSceneBox& bx = boxes[b];
                float t0 = -3.402823466e+38f, t1 = 3.402823466e+38f; bool miss = false;
                for (int a = 0; a < 3; ++a) {
                    if (d[a] == 0.0f) { if (o[a] < bx.lo[a] || o[a] > bx.hi[a]) miss = true; continue; }
                    const float inv = 1.0f / d[a];
                    float ta = (bx.lo[a] - o[a]) * inv, tb = (bx.hi[a] - o[a]) * inv;
                    if (ta > tb) { const float tmp = ta; ta = tb; tb = tmp; }
                    if (ta > t0) t0 = ta;
                    if (tb < t1) t1 = tb;
                }
                if (miss || t0 > t1) continue;
                const float t = bx.insideOut ? t1 : t0;
                if (t > 1e-6f && t < best) { best = t; hit = b; }
            }
            const size_t q = (size_t)j * res + i;
            if (hit < 0) { depth[q] = 65535; diffuse[q * 4 + 0] = diffuse[q * 4 + 1] = diffuse[q * 4 + 2] = diffuse[q * 4 + 3] = 0; continue; }
            const vec3 fsPosition = o + d * best;
            const float z = shadowVplDepth(fsPosition, o, v.farPlane);
            depth[q] = (uint16_t)floor(z * 65535.0f + 0.5f);
            for (int c = 0; c < 3; ++c) diffuse[q * 4 + c] = boxes[hit].albedo[c];
            diffuse[q * 4 + 3] = 255;
        }
    }
    return 0;
}

}  // extern "C"
