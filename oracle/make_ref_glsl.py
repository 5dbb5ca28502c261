#!/usr/bin/env python
"""Compile the REFERENCE's own GLSL text for the GI path as C++ against the reference's vendored glm.  TEST INFRASTRUCTURE
(the checker's checker).

The shaders of StratusGFX need an OpenGL driver to run as GLSL, but GLSL of this kind is C++ once glm supplies the vector types
(glm exists to mirror GLSL), and the reference vendors glm.  This script reads the six shaders of the path and the library files
they include from where they lie under /root/reference/Source/Shaders, extracts by name the functions the path executes — main()
of viscull_vpls.cs, vpl_light_color.cs, vpl_tiled_deferred_culling_stage1.cs, ..._stage2.cs and vpl_pbr_gi_denoise.fs,
performLightingCalculations of vpl_pbr_gi.fs, and everything they call in pbr.glsl, pbr2.glsl, fog.glsl and common.glsl, with the
macros they use — and applies purely lexical adaptations:

    float literals get an `f` suffix                       (GLSL literals are 32-bit floats, C++ literals are doubles)
    `in T x` -> `T x`,  `out|inout T x` -> `T& x`          (parameter qualifiers)
    `.xyz` -> `.xyz()`                                     (multi-component swizzles are member functions in glm 0.9.8 on GCC)
    `v.xyz += e;` / `v.xy = e;` -> swizzleAdd3(v, e) ...   (swizzles as l-values: glm's swizzle functions return copies)
    `float a[2] = float[](x, y);` -> `= {x, y};`           (array constructors)
    `void main(` -> `void <shader>_main(`                  (one translation unit holds several shaders)

What an OpenGL driver would supply is written out in the prelude from the GL 4.6 specification and is NOT
reference code: samplers (texel fetch at texel centres, bilinear filter with 8-bit sub-texel weights, cube-map face selection,
hardware PCF), interface blocks and uniforms as globals, compute built-ins, barrier() (real: the cull runs as 1024 threads),
`shared` arrays.  The result goes to oracle/_ref/ref_glsl_gen.cpp (generated, git-ignored, never committed, deleted after the compile unless
SGI_KEEP_GENERATED=1) and is compiled to oracle/_ref/libref_glsl.so.  tests/test_reference_glsl.py holds every stage of the oracle against it bit for bit, and
tests/golden/make_shader_text_golden.py chains the compiled shaders into a fixture that no oracle code touched.
Nothing of the reference is copied into the repository.

    python oracle/make_ref_glsl.py          (also run by oracle/oracle_py.py::build when the reference tree is present)
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
SHADERS = os.path.join(REF, "Source", "Shaders")
OUT_DIR = os.path.join(HERE, "_ref")

# (file, function names in the order they are needed; every overload of a name is taken)
WANTED = [
    ("common.glsl", ["boundHDR", "saturate", "computeTexelSize", "worldPositionFromDepth", "hash", "floatConstruct", "random", "sampleNormalWithOffset", "sampleNormal"]),
    ("fog.glsl", ["applyFog"]),
    ("pbr.glsl", ["calculateShadowValue8Samples", "calculateShadowValue1Sample", "quadraticAttenuation", "sampleShadowTexture", "calculateInfiniteShadowValue", "vplDiffuseAttenuation", "vplSpecularAttenuation"]),
    ("pbr2.glsl", ["NDF_GGX", "Visibility_G_SmithGGX", "Fr_Fresnel_Schlick", "Fd_Lambert", "singleScatteringBRDF_Specular",
                   "singleScatteringBRDF_Diffuse_Lambert", "RemapRoughness", "BRDF_Base", "BRDF_DotProducts",
                   "BRDF_Lambert_SeparateDiffuseSpecular", "calculateLighting_Lambert_SeparateDiffuseSpecular",
                   "calculateVirtualPointLighting2",
                   # the deferred direct-lighting pass: Burley diffuse + the same specular lobe
                   "Fd_Fresnel_Schlick", "Fd_Burley", "singleScatteringBRDF_Diffuse_Burley", "BRDF_Burley", "calculateLighting_Burley",
                   "calculateDirectionalLighting", "calculatePointLighting2"]),
    ("aa_common.glsl", ["tonemap", "inverseTonemap"]),          # Reinhard pair used by taa.fs (widening row f4, second half)
    ("vpl_pbr_gi.fs", ["performLightingCalculations"]),          # the whole body of the gather shader (main() only forwards to it)
    ("vpl_pbr_gi_denoise.fs", ["filterInput", "computeMergedReservoir", "main"]),     # the denoiser: main() is emitted as denoise_main()
]
SHADER_CONSTS = [("vpl_pbr_gi_denoise.fs", ["sigmaZ", "sigmaN", "dminmax"])]        # file-scope `const T name = value;` the functions read
SHADER_DEFINES = [("vpl_pbr_gi.fs", ["STANDARD_MAX_SAMPLES_PER_PIXEL", "ABSOLUTE_MAX_SAMPLES_PER_PIXEL", "MAX_RESAMPLES_PER_PIXEL"])]
DEFINES = ["PI", "PREVENT_DIV_BY_ZERO", "HALF_FLOAT_MAX", "FLOAT_MAX"]       # taken from common.glsl
# the compute shaders: (file, nested namespace, name main() is emitted under, macros to take along)
COMPUTE = [
    ("vpl_light_color.cs", "cs_color", "color_main", []),
    ("vpl_tiled_deferred_culling_stage1.cs", "cs_stage1", "stage1_main", []),
    ("vpl_tiled_deferred_culling_stage2.cs", "cs_stage2", "stage2_main", ["SHUFFLE_DOWN"]),
    ("viscull_vpls.cs", "cs_cull", "cull_main", []),
    ("pbr.fs", "fs_direct", "direct_main", []),           # widening row f3: the deferred direct-lighting pass that produces `screen`
    # widening row f4 (first half): the two SSAO passes that produce the `ssao` plane the gather reads (BE:1160-1204)
    ("ssao.fs", "fs_ssao", "ssao_main", [], ["saturate", "calculateAmbientOcclusion"]),
    ("ssao_blur.fs", "fs_ssao_blur", "ssao_blur_main", [], ["blurAmbientOcclusion"]),
    ("taa.fs", "fs_taa", "taa_main", [], []),              # widening row f4 (second half): temporal anti-aliasing (BE:2324-2354)
    ("shadowVpl.fs", "fs_shadow_vpl", "shadow_vpl_main", [], []),      # widening row f2: the fragment stage of the VPL cube-map pass (BE:1426-1575)
]
TYPES = r"(?:float|int|uint|bool|vec2|vec3|vec4|uvec2|uvec3|uvec4|ivec2|mat4|void|sampler2D|sampler2DRect|samplerCubeArray|sampler2DArrayShadow)"


def strip_comments(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def extract(text, name):
    """Every definition `T name(...) { ... }` in `text` (brace matched), in file order."""
    out = []
    for m in re.finditer(r"^[ \t]*" + TYPES + r"\s+" + re.escape(name) + r"\s*\(", text, flags=re.M):
        i = text.index("{", m.end())
        if ";" in text[m.end():i]:
            continue                                    # a prototype
        depth, j = 0, i
        while True:
            c = text[j]
            depth += (c == "{") - (c == "}")
            j += 1
            if depth == 0:
                break
        if re.search(r"\bsamplerCube\b|\bsampler2DArray\b(?!Shadow)", text[m.start():j]):
            continue                                    # an overload for a sampler type these shaders never use
        out.append(text[m.start():j])
    return out


def extract_macro(text, name):
    """A function-like `#define name(...)` with its continuation lines."""
    m = re.search(r"^#define\s+" + re.escape(name) + r"\b.*?(?<!\\)\n", text, flags=re.M | re.S)
    return m.group(0)


def adapt(code):
    # 32-bit float literals
    code = re.sub(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])", r"\1f", code)
    # parameter qualifiers
    code = re.sub(r"\b(?:out|inout)\s+(" + TYPES + r")\s+(\w+)", r"\1& \2", code)
    code = re.sub(r"\bin\s+(" + TYPES + r")\s+(\w+)", r"\1 \2", code)
    # swizzles as l-values (`v.rgb += e;`, `c[i].xyz = e;`, `p.xy = e;`): glm's swizzle functions return copies, so spell them out
    lhs = r"((?:\b\w+)(?:\[[^\]]+\])?)"
    code = re.sub(lhs + r"\.(?:rgb|xyz)\s*\+=\s*([^;]+);", r"swizzleAdd3(\1, \2);", code)
    code = re.sub(lhs + r"\.(?:rgb|xyz)\s*=(?!=)\s*([^;]+);", r"swizzleSet3(\1, \2);", code)
    code = re.sub(lhs + r"\.(?:rg|xy)\s*\+=\s*([^;]+);", r"swizzleAdd2(\1, \2);", code)
    code = re.sub(lhs + r"\.(?:rg|xy)\s*=(?!=)\s*([^;]+);", r"swizzleSet2(\1, \2);", code)
    # GLSL array constructors: `float a[2] = float[](x, y);`
    code = re.sub(r"=\s*float\[\]\(([^;]*)\);", r"= {\1};", code)
    # multi-component swizzles used as values
    code = re.sub(r"\.((?:[xyzw]{2,4})|(?:[rgba]{2,4}))\b(?!\s*\()", r".\1()", code)
    return code


PRELUDE = r'''// GENERATED by oracle/make_ref_glsl.py from the reference's GLSL (see that file).  Do not edit, do not commit.
#define GLM_FORCE_PURE
#define GLM_FORCE_SWIZZLE
#include <glm/glm.hpp>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <omp.h>
#include <ucontext.h>
#include <memory>
#include <vector>
namespace refglsl {
using namespace glm;
// the uniforms these functions read (pbr2.glsl: usePerceptualRoughness, minRoughness; fog.glsl: fogColor, fogDensity)
static thread_local bool usePerceptualRoughness = true;
static thread_local float minRoughness = 0.08f;
static thread_local vec3 fogColor = vec3(0.5f);
static thread_local float fogDensity = 0.0f;
// GLSL converts int arguments to float implicitly (pbr.glsl: `mix(10, 1, ...)`)
// (the using-declaration keeps glm::mix in the overload set: a function declared in this namespace would otherwise hide it)
using glm::mix;
static inline float mix(int a, int b, float t) { return glm::mix((float)a, (float)b, t); }

// ---- what the GL driver supplies to vpl_pbr_gi.fs, restated from the GL 4.6 specification (SURVEY.md Appendix A), NOT reference code:
//      texel fetches at texel centres (NEAREST == LINEAR there), unorm / half decoding, cube-map face selection ----
static inline float halfToFloat(unsigned short h) {
    const unsigned s = (h >> 15) & 1u, e = (h >> 10) & 31u, m = h & 1023u;
    unsigned bits;
    if (e == 0) {
        if (m == 0) bits = s << 31;
        else { int ee = -1; unsigned mm = m; do { ++ee; mm <<= 1; } while (!(mm & 1024u)); bits = (s << 31) | ((unsigned)(112 - ee) << 23) | ((mm & 1023u) << 13); }
    } else if (e == 31) bits = (s << 31) | 0x7f800000u | (m << 13);
    else bits = (s << 31) | ((e + 112u) << 23) | (m << 13);
    float f; std::memcpy(&f, &bits, 4); return f;
}
enum TexFormat { R32F, RGBA8, RG8, R16F, RG16F, RGBA16F };
enum TexFilter { NEAREST, LINEAR };
struct sampler2D { const void* data = nullptr; int W = 0, H = 0; TexFormat fmt = R32F; TexFilter filter = NEAREST; bool repeat = false; };
struct sampler2DRect { const void* data = nullptr; int W = 0, H = 0; TexFormat fmt = R16F; TexFilter filter = NEAREST; };
static inline vec4 fetchTexel(const void* data, TexFormat fmt, int W, int i, int j) {
    const size_t p = (size_t)j * W + i;
    switch (fmt) {
        case R32F: return vec4(((const float*)data)[p], 0.0f, 0.0f, 1.0f);
        case RGBA8: { const unsigned char* t = (const unsigned char*)data + 4 * p; return vec4(t[0] / 255.0f, t[1] / 255.0f, t[2] / 255.0f, t[3] / 255.0f); }
        case RG8: { const unsigned char* t = (const unsigned char*)data + 2 * p; return vec4(t[0] / 255.0f, t[1] / 255.0f, 0.0f, 1.0f); }
        case R16F: return vec4(halfToFloat(((const unsigned short*)data)[p]), 0.0f, 0.0f, 1.0f);
        case RG16F: { const unsigned short* t = (const unsigned short*)data + 2 * p; return vec4(halfToFloat(t[0]), halfToFloat(t[1]), 0.0f, 1.0f); }
        case RGBA16F: { const unsigned short* t = (const unsigned short*)data + 4 * p; return vec4(halfToFloat(t[0]), halfToFloat(t[1]), halfToFloat(t[2]), halfToFloat(t[3])); }
    }
    return vec4(0.0f);
}
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int wrapi(int v, int n) { const int m = v % n; return m < 0 ? m + n : m; }
static inline vec4 textureNearest(const sampler2D& s, vec2 uv) {             // CLAMP_TO_EDGE, or REPEAT (the 4x4 SSAO rotation table, BE:776-778)
    if (s.repeat) return fetchTexel(s.data, s.fmt, s.W, wrapi((int)std::floor(uv.x * (float)s.W), s.W), wrapi((int)std::floor(uv.y * (float)s.H), s.H));
    return fetchTexel(s.data, s.fmt, s.W, clampi((int)std::floor(uv.x * (float)s.W), 0, s.W - 1), clampi((int)std::floor(uv.y * (float)s.H), 0, s.H - 1));
}
// texture(): NEAREST picks the texel that contains the coordinate; LINEAR blends the four neighbours with 8-bit sub-texel weights
// (the precision of the hardware filter), in the operation order of oracle.cpp's bilinMix — the same model of the GL filter on both sides
static inline vec4 texture(const sampler2D& s, vec2 uv) {
    if (s.filter == NEAREST) return textureNearest(s, uv);
    const float fu = uv.x * (float)s.W - 0.5f, fv = uv.y * (float)s.H - 0.5f;
    const float i0f = std::floor(fu), j0f = std::floor(fv);
    const float a = std::floor((fu - i0f) * 256.0f + 0.5f) * (1.0f / 256.0f), b = std::floor((fv - j0f) * 256.0f + 0.5f) * (1.0f / 256.0f);
    const int i0 = clampi((int)i0f, 0, s.W - 1), i1 = clampi((int)i0f + 1, 0, s.W - 1), j0 = clampi((int)j0f, 0, s.H - 1), j1 = clampi((int)j0f + 1, 0, s.H - 1);
    const vec4 t00 = fetchTexel(s.data, s.fmt, s.W, i0, j0), t10 = fetchTexel(s.data, s.fmt, s.W, i1, j0), t01 = fetchTexel(s.data, s.fmt, s.W, i0, j1), t11 = fetchTexel(s.data, s.fmt, s.W, i1, j1);
    vec4 r;
    for (int c = 0; c < 4; ++c) r[c] = (1.0f - a) * (1.0f - b) * t00[c] + a * (1.0f - b) * t10[c] + (1.0f - a) * b * t01[c] + a * b * t11[c];
    return r;
}
static inline vec4 textureLod(const sampler2D& s, vec2 uv, float) { return texture(s, uv); }
static inline vec4 textureLod(const sampler2D& s, vec2 uv, int) { return texture(s, uv); }
static inline vec4 textureOffset(const sampler2D& s, vec2 uv, ivec2 offset) { return texture(s, uv + vec2(offset) / vec2((float)s.W, (float)s.H)); }
// textureSize returns ivec2 in GLSL; every use in these shaders converts it to float at once (`1.0 / textureSize(t, 0).xy`, `vec2 wh = textureSize(..)`)
static inline vec2 textureSize(const sampler2D& s, int) { return vec2((float)s.W, (float)s.H); }
template <class V, class S> static inline void swizzleAdd3(V& v, const S& s) { v.x += s.x; v.y += s.y; v.z += s.z; }
template <class V, class S> static inline void swizzleSet3(V& v, const S& s) { const vec3 t(s); v.x = t.x; v.y = t.y; v.z = t.z; }
template <class V, class S> static inline void swizzleAdd2(V& v, const S& s) { v.x += s.x; v.y += s.y; }
template <class V, class S> static inline void swizzleSet2(V& v, const S& s) { const vec2 t(s); v.x = t.x; v.y = t.y; }
static inline vec4 texture(const sampler2DRect& s, vec2 pixelCoords) {       // rectangle textures are addressed in pixels; CLAMP_TO_EDGE
    if (s.filter == NEAREST)
        return fetchTexel(s.data, s.fmt, s.W, clampi((int)std::floor(pixelCoords.x), 0, s.W - 1), clampi((int)std::floor(pixelCoords.y), 0, s.H - 1));
    // LINEAR (the structure buffer, BE:406-408): the same filter model as sampler2D above, in pixel units
    const float fu = pixelCoords.x - 0.5f, fv = pixelCoords.y - 0.5f;
    const float i0f = std::floor(fu), j0f = std::floor(fv);
    const float a = std::floor((fu - i0f) * 256.0f + 0.5f) * (1.0f / 256.0f), b = std::floor((fv - j0f) * 256.0f + 0.5f) * (1.0f / 256.0f);
    const int i0 = clampi((int)i0f, 0, s.W - 1), i1 = clampi((int)i0f + 1, 0, s.W - 1), j0 = clampi((int)j0f, 0, s.H - 1), j1 = clampi((int)j0f + 1, 0, s.H - 1);
    const vec4 t00 = fetchTexel(s.data, s.fmt, s.W, i0, j0), t10 = fetchTexel(s.data, s.fmt, s.W, i1, j0), t01 = fetchTexel(s.data, s.fmt, s.W, i0, j1), t11 = fetchTexel(s.data, s.fmt, s.W, i1, j1);
    vec4 r;
    for (int c = 0; c < 4; ++c) r[c] = (1.0f - a) * (1.0f - b) * t00[c] + a * (1.0f - b) * t10[c] + (1.0f - a) * b * t01[c] + a * b * t11[c];
    return r;
}
struct samplerCubeArray { const unsigned short* data = nullptr; int res = 0; const unsigned char* rgba8 = nullptr; };   // DEPTH_COMPONENT16 (or RGBA8), [layer][face][res][res], NEAREST
static inline vec4 texture(const samplerCubeArray& s, vec4 dirLayer) {      // GL 4.6 table 8.19 (major axis, sc, tc), then the texel of the face
    const float x = dirLayer.x, y = dirLayer.y, z = dirLayer.z, ax = std::fabs(x), ay = std::fabs(y), az = std::fabs(z);
    int face; float sc, tc, ma;
    if (ax >= ay && ax >= az) { ma = ax; if (x >= 0.0f) { face = 0; sc = -z; tc = -y; } else { face = 1; sc = z; tc = -y; } }
    else if (ay >= az) { ma = ay; if (y >= 0.0f) { face = 2; sc = x; tc = z; } else { face = 3; sc = x; tc = -z; } }
    else { ma = az; if (z >= 0.0f) { face = 4; sc = x; tc = -y; } else { face = 5; sc = -x; tc = -y; } }
    const float sCoord = (sc / ma + 1.0f) / 2.0f, tCoord = (tc / ma + 1.0f) / 2.0f;
    const int i = clampi((int)std::floor(sCoord * (float)s.res), 0, s.res - 1), j = clampi((int)std::floor(tCoord * (float)s.res), 0, s.res - 1);
    const int layer = (int)dirLayer.w;
    const size_t texel = (((size_t)layer * 6 + face) * s.res + j) * s.res + i;
    if (s.rgba8 != nullptr) { const unsigned char* c = s.rgba8 + 4 * texel; return vec4(c[0] / 255.0f, c[1] / 255.0f, c[2] / 255.0f, c[3] / 255.0f); }
    const unsigned short t = s.data[texel];
    return vec4((float)t / 65535.0f, 0.0f, 0.0f, 1.0f);
}
static inline vec4 textureLod(const samplerCubeArray& s, vec4 dirLayer, int) { return texture(s, dirLayer); }
// sampler2DArrayShadow, LINEAR + compare LEQUAL (hardware PCF, SURVEY.md Appendix A4): the reference value (clamped to [0, 1] as for
// a fixed-point depth format) is compared with the four bilinear neighbours and the 0/1 results are blended with the bilinear
// weights, in the operation order of oracle.cpp's shadowArrayPcf — the same reading of the GL filter on both sides
struct sampler2DArrayShadow { const float* data = nullptr; int res = 0; int layers = 4; };
static inline float texture(const sampler2DArrayShadow& s, vec4 c) {         // c = (u, v, layer, reference depth)
    const int layer = clampi((int)std::floor(c.z + 0.5f), 0, s.layers - 1);
    const float* m = s.data + (size_t)layer * s.res * s.res;
    const float fu = c.x * (float)s.res - 0.5f, fv = c.y * (float)s.res - 0.5f;
    const float i0f = std::floor(fu), j0f = std::floor(fv);
    const float a = fu - i0f, b = fv - j0f;
    const float ref = glm::clamp(c.w, 0.0f, 1.0f);
    const int i0 = clampi((int)i0f, 0, s.res - 1), i1 = clampi((int)i0f + 1, 0, s.res - 1), j0 = clampi((int)j0f, 0, s.res - 1), j1 = clampi((int)j0f + 1, 0, s.res - 1);
    auto lit = [&](int i, int j) { return ref <= m[(size_t)j * s.res + i] ? 1.0f : 0.0f; };
    return (1.0f - a) * (1.0f - b) * lit(i0, j0) + a * (1.0f - b) * lit(i1, j0) + (1.0f - a) * b * lit(i0, j1) + a * b * lit(i1, j1);
}
static inline vec3 textureSize(const sampler2DArrayShadow& s, int) { return vec3((float)s.res, (float)s.res, (float)s.layers); }
// ---- compute-shader built-ins (per invocation) and the uniforms pbr.glsl declares for the sun's cascaded shadow map ----
static thread_local uvec3 gl_NumWorkGroups, gl_WorkGroupSize, gl_WorkGroupID, gl_LocalInvocationID, gl_GlobalInvocationID;
static thread_local uint gl_LocalInvocationIndex;
static void (*barrierHook)() = nullptr;
static inline void barrier() { if (barrierHook) barrierHook(); }
static inline int atomicAdd(int& mem, int v) { return __atomic_fetch_add(&mem, v, __ATOMIC_SEQ_CST); }
static thread_local vec3 infiniteLightDirection; static thread_local float infiniteLightZnear = 0.0f, infiniteLightZfar = 1.0f, infiniteLightDepthBias = 0.0f;
static thread_local float pointLightAmbientIntensity = 0.003f;            // pbr.glsl:66
static thread_local sampler2DArrayShadow infiniteLightShadowMap; static thread_local vec4 cascadePlanes[3]; static thread_local mat4 cascadeProjViews[4];
struct VplStage1PerTileOutputs { vec4 averageLocalPosition; vec4 averageLocalNormal; };        // vpl_common.glsl
#define MAX_VPLS_PER_TILE 12
#define MAX_TOTAL_VPLS_BEFORE_CULLING 10000
#define MAX_TOTAL_VPLS_PER_FRAME 4480
struct VplStage2PerTileOutputs { int numVisible; int indices[MAX_VPLS_PER_TILE]; };
// ---- the interface blocks and uniforms of vpl_pbr_gi.fs (declarations restated; values set by refglsl_gather) ----
#define MAX_TOTAL_SHADOW_ATLASES 14
struct VplData { vec4 position; vec4 color; vec4 specularPosition; float radius; float farPlane; float intensity; float _1; };      // vpl_common.glsl:48-56
struct AtlasEntry { int index; int layer; };
// (thread_local: every host thread that runs fragments owns a copy of the shader's globals, so the full-screen passes can be
//  spread over the host's cores; compiled with -ftls-model=local-dynamic: one TLS base lookup per function; initial-exec does not fit the static TLS block of a dlopen-ed library)
#define PERINV static thread_local
PERINV vec2 fsTexCoords; PERINV vec4 gl_FragCoord;
PERINV vec3 color; PERINV vec4 reservoir;                                            // the two fragment outputs
PERINV sampler2D gDepth, gNormal, gAlbedo, gBaseReflectivity, gRoughnessMetallicAmbient, screen, historyDepth;
PERINV sampler2DRect ssao;
PERINV vec3 viewPosition; PERINV mat4 invProjectionView; PERINV float time; PERINV float minGiOcclusionFactor = 0.95f;
PERINV const VplData* lightData; PERINV const int* numVisible; PERINV const AtlasEntry* shadowIndices;
PERINV samplerCubeArray shadowCubeMaps[MAX_TOTAL_SHADOW_ATLASES];
// ---- the same for vpl_pbr_gi_denoise.fs ----
PERINV vec3 combinedColor, giColor; PERINV vec4 reservoirValue; PERINV float newHistoryDepth;       // the four fragment outputs
PERINV sampler2D albedo, velocity, normal, ids, depth, prevNormal, prevIds, prevDepth, indirectIllumination, indirectShadows, prevIndirectIllumination;
PERINV int multiplier = 0, passNumber = 0, numReservoirNeighbors = 10; PERINV bool final = false, mergeReservoirs = false;
PERINV float framesPerSecond = 60.0f;
'''

EPILOGUE = r'''
}  // namespace refglsl

extern "C" {
void refglsl_set_uniforms(int usePerceptualRoughness, float minRoughness, const float* fogColor, float fogDensity) {
    refglsl::usePerceptualRoughness = usePerceptualRoughness != 0; refglsl::minRoughness = minRoughness;
    refglsl::fogColor = glm::vec3(fogColor[0], fogColor[1], fogColor[2]); refglsl::fogDensity = fogDensity;
}
// a: fragPosition, baseColor, normal, viewDir, specularLightPos, diffuseLightPos, lightColor (7 x vec3), then viewDist, lightRadius,
//    roughness, metallic, ambientOcclusion, shadowFactor (6 floats), then baseReflectance (vec3): 30 floats
void refglsl_vpl_lighting(const float* a, float* out3) {
    using glm::vec3;
    auto v = [&](int i) { return vec3(a[3 * i], a[3 * i + 1], a[3 * i + 2]); };
    const vec3 r = refglsl::calculateVirtualPointLighting2(v(0), v(1), v(2), v(3), v(4), v(5), v(6), a[21], a[22], a[23], a[24], a[25], a[26],
                                                           vec3(a[27], a[28], a[29]));
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
void refglsl_world_pos(float u, float v, float depth, const float* invPV16, float* out3) {
    glm::mat4 m; std::memcpy(&m[0][0], invPV16, 64);
    const glm::vec3 r = refglsl::worldPositionFromDepth(glm::vec2(u, v), depth, m);
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
float refglsl_random3(float x, float y, float z) { return refglsl::random(glm::vec3(x, y, z)); }
unsigned refglsl_hash3(float x, float y, float z) { return refglsl::hash(glm::floatBitsToUint(glm::vec3(x, y, z))); }
float refglsl_attenuation(const float* dir3, float radius) { return refglsl::vplDiffuseAttenuation(glm::vec3(dir3[0], dir3[1], dir3[2]), radius); }
float refglsl_bound_hdr(float v) { return refglsl::boundHDR(v); }
float refglsl_shadow1(const unsigned short* atlas, int res, int layer, float farPlane, const float* fragPos, const float* lightPos, float lightNormalDot, float minBias) {
    refglsl::samplerCubeArray s; s.data = atlas; s.res = res;
    return refglsl::calculateShadowValue1Sample(s, layer, farPlane, glm::vec3(fragPos[0], fragPos[1], fragPos[2]), glm::vec3(lightPos[0], lightPos[1], lightPos[2]), lightNormalDot, minBias);
}

float refglsl_sun_shadow(const float* sunDir, const float* cascadeProjViews16x4, const float* cascadePlanes4x3, const float* csm, int csmRes,
                         const float* pos3, const float* normal3) {
    using namespace refglsl;
    infiniteLightDirection = glm::vec3(sunDir[0], sunDir[1], sunDir[2]);
    for (int c = 0; c < 4; ++c) std::memcpy(&cascadeProjViews[c][0][0], cascadeProjViews16x4 + 16 * c, 64);
    for (int c = 0; c < 3; ++c) cascadePlanes[c] = glm::vec4(cascadePlanes4x3[4 * c], cascadePlanes4x3[4 * c + 1], cascadePlanes4x3[4 * c + 2], cascadePlanes4x3[4 * c + 3]);
    infiniteLightShadowMap.data = csm; infiniteLightShadowMap.res = csmRes; infiniteLightShadowMap.layers = 4;
    const glm::vec3 p(pos3[0], pos3[1], pos3[2]);
    const glm::vec3 blends(glm::dot(cascadePlanes[0], glm::vec4(p, 1.0f)), glm::dot(cascadePlanes[1], glm::vec4(p, 1.0f)), glm::dot(cascadePlanes[2], glm::vec4(p, 1.0f)));
    return calculateInfiniteShadowValue(glm::vec4(p, 1.0f), blends, glm::vec3(normal3[0], normal3[1], normal3[2]), false);
}

// One full-screen pass of vpl_pbr_gi.fs: main() is `performLightingCalculations(textureLod(screen, fsTexCoords, 0).rgb, gl_FragCoord.xy, fsTexCoords)`
// for every fragment of a W x H quad (fragment centres at (x + 0.5, y + 0.5), texture coordinates (x + 0.5) / W, (y + 0.5) / H).
// Outputs are the float values written to `color` and `reservoir` before the RGB16F / RGBA16F framebuffer store.  Rows
// [rowBegin, rowEnd) only (0, H = the whole quad; a sub-range is bench.py's bounded CPU sample); fragments are spread over the
// host's threads, each with its own copy of the shader's globals.
typedef struct RefGatherArgs {
    int W, H, cubeRes, numAtlases;
    const float* depth; const unsigned char* normal4; const unsigned char* albedo4; const unsigned char* refl2; const unsigned char* rma4;
    const unsigned short* ssaoH; const unsigned short* screenH4;
    const float* viewPosition; const float* invProjectionView; float time; float minGiOcclusionFactor;
    const void* lightData; const int* numVisible; const void* shadowIndices; const unsigned short* const* shadowAtlases;
    float* colorOut3; float* reservoirOut4;
    int rowBegin, rowEnd;
    int usePerceptualRoughness; float minRoughness; float fogColor[3]; float fogDensity;
} RefGatherArgs;
void refglsl_gather(const RefGatherArgs* a) {
    using namespace refglsl;
    #pragma omp parallel
    {
        auto tex = [&](const void* d, TexFormat f) { sampler2D s; s.data = d; s.W = a->W; s.H = a->H; s.fmt = f; return s; };
        gDepth = tex(a->depth, R32F); gNormal = tex(a->normal4, RGBA8); gAlbedo = tex(a->albedo4, RGBA8); gBaseReflectivity = tex(a->refl2, RG8);
        gRoughnessMetallicAmbient = tex(a->rma4, RGBA8);
        screen = tex(a->depth, R32F); historyDepth = tex(a->depth, R32F);    // bound by the reference but without effect on the outputs (any texture serves)
        ssao.data = a->ssaoH; ssao.W = a->W; ssao.H = a->H; ssao.fmt = R16F;
        viewPosition = glm::vec3(a->viewPosition[0], a->viewPosition[1], a->viewPosition[2]);
        std::memcpy(&invProjectionView[0][0], a->invProjectionView, 64);
        refglsl::time = a->time; minGiOcclusionFactor = a->minGiOcclusionFactor;
        usePerceptualRoughness = a->usePerceptualRoughness != 0; minRoughness = a->minRoughness;
        fogColor = glm::vec3(a->fogColor[0], a->fogColor[1], a->fogColor[2]); fogDensity = a->fogDensity;
        lightData = (const VplData*)a->lightData; numVisible = a->numVisible; shadowIndices = (const AtlasEntry*)a->shadowIndices;
        for (int i = 0; i < a->numAtlases && i < MAX_TOTAL_SHADOW_ATLASES; ++i) { shadowCubeMaps[i].data = a->shadowAtlases[i]; shadowCubeMaps[i].res = a->cubeRes; }
        #pragma omp for schedule(dynamic, 2)
        for (int y = a->rowBegin; y < a->rowEnd; ++y)
            for (int x = 0; x < a->W; ++x) {
                gl_FragCoord = glm::vec4((float)x + 0.5f, (float)y + 0.5f, 0.0f, 1.0f);
                fsTexCoords = glm::vec2(((float)x + 0.5f) / (float)a->W, ((float)y + 0.5f) / (float)a->H);
                performLightingCalculations(glm::vec3(textureLod(screen, fsTexCoords, 0)), glm::vec2(gl_FragCoord), fsTexCoords);
                const size_t p = (size_t)y * a->W + x;
                a->colorOut3[3 * p] = color.x; a->colorOut3[3 * p + 1] = color.y; a->colorOut3[3 * p + 2] = color.z;
                a->reservoirOut4[4 * p] = reservoir.x; a->reservoirOut4[4 * p + 1] = reservoir.y; a->reservoirOut4[4 * p + 2] = reservoir.z; a->reservoirOut4[4 * p + 3] = reservoir.w;
            }
    }
}

// One full-screen pass of vpl_pbr_gi_denoise.fs (its main(), emitted as denoise_main()).  Float outputs before the framebuffer store.
typedef struct RefDenoiseArgs {
    int W, H, multiplier, mergeReservoirs, final_;
    float time, framesPerSecond;
    const int* numVisible;
    const unsigned short* screenH4; const unsigned short* velocityH2; const unsigned char* normal4; const float* ids; const float* depth;
    const unsigned char* prevNormal4; const float* prevIds; const float* prevDepth;
    const unsigned short* indirectIlluminationH4; const unsigned short* indirectShadowsH4; const unsigned short* prevIndirectIlluminationH4;
    const unsigned short* historyDepthH;
    float* combinedOut3; float* giOut3; float* reservoirOut4; float* historyOut;
    int rowBegin, rowEnd;
} RefDenoiseArgs;
void refglsl_denoise_pass(const RefDenoiseArgs* a) {
    using namespace refglsl;
    #pragma omp parallel
    {
        auto tex = [&](const void* d, TexFormat f, TexFilter filt) { sampler2D s; s.data = d; s.W = a->W; s.H = a->H; s.fmt = f; s.filter = filt; return s; };
        // filters as the reference creates these textures (SURVEY.md Appendix A1): only the reprojected previous-frame reads are really bilinear
        screen = tex(a->screenH4, RGBA16F, NEAREST); velocity = tex(a->velocityH2, RG16F, NEAREST); normal = tex(a->normal4, RGBA8, NEAREST);
        ids = tex(a->ids, R32F, NEAREST); depth = tex(a->depth, R32F, NEAREST);
        prevNormal = tex(a->prevNormal4, RGBA8, LINEAR); prevIds = tex(a->prevIds, R32F, NEAREST); prevDepth = tex(a->prevDepth, R32F, LINEAR);
        indirectIllumination = tex(a->indirectIlluminationH4, RGBA16F, NEAREST); indirectShadows = tex(a->indirectShadowsH4, RGBA16F, NEAREST);
        prevIndirectIllumination = tex(a->prevIndirectIlluminationH4, RGBA16F, LINEAR); historyDepth = tex(a->historyDepthH, R16F, NEAREST);
        multiplier = a->multiplier; mergeReservoirs = a->mergeReservoirs != 0; final = a->final_ != 0;
        refglsl::time = a->time; framesPerSecond = a->framesPerSecond; numVisible = a->numVisible;
        #pragma omp for schedule(dynamic, 2)
        for (int y = a->rowBegin; y < a->rowEnd; ++y)
            for (int x = 0; x < a->W; ++x) {
                gl_FragCoord = glm::vec4((float)x + 0.5f, (float)y + 0.5f, 0.0f, 1.0f);
                fsTexCoords = glm::vec2(((float)x + 0.5f) / (float)a->W, ((float)y + 0.5f) / (float)a->H);
                denoise_main();
                const size_t p = (size_t)y * a->W + x;
                for (int c = 0; c < 3; ++c) { a->combinedOut3[3 * p + c] = combinedColor[c]; a->giOut3[3 * p + c] = giColor[c]; }
                for (int c = 0; c < 4; ++c) a->reservoirOut4[4 * p + c] = reservoirValue[c];
                a->historyOut[p] = newHistoryDepth;
            }
    }
}

// ---- the four compute shaders.  A dispatch is emulated invocation by invocation with the built-ins set as GL defines them;
//      invocations of a work group run one after the other where only invocation 0 reads `shared` data after the barrier
//      (stage 1: it runs last), and as 1024 real threads with a real barrier where the phases depend on each other (the cull).
//      Work groups are spread over the host's threads ----
// One full-screen pass of pbr.fs (deferred direct lighting: point lights without and with cube shadows, the sun with its cascaded shadow map).
typedef struct RefDirectArgs {
    int W, H, cubeRes, numAtlases, csmRes, numLights, numShadowLights, rowBegin, rowEnd;
    const float* depth; const unsigned char* normal4; const unsigned char* albedo4; const unsigned char* refl2; const unsigned char* rma4; const unsigned short* ssaoH;
    const float* viewPosition; const float* invProjectionView; const float* sunDir; const float* sunColor; const float* cascadeProjViews; const float* cascadePlanes; const float* csm;
    const void* nonShadowCasters; const void* shadowCasters; const void* shadowIndices; const unsigned short* const* shadowAtlases;
    float emissionStrength; float depthBias, zNear, zFar;
    float* colorOut3;
} RefDirectArgs;
void refglsl_direct_lighting(const RefDirectArgs* a) {
    using namespace refglsl;
    #pragma omp parallel
    {
        auto tex = [&](const void* d, TexFormat f) { sampler2D s; s.data = d; s.W = a->W; s.H = a->H; s.fmt = f; return s; };
        fs_direct::gDepth = tex(a->depth, R32F); fs_direct::gNormal = tex(a->normal4, RGBA8); fs_direct::gAlbedo = tex(a->albedo4, RGBA8);
        fs_direct::gBaseReflectivity = tex(a->refl2, RG8); fs_direct::gRoughnessMetallicAmbient = tex(a->rma4, RGBA8);
        fs_direct::ssao.data = a->ssaoH; fs_direct::ssao.W = a->W; fs_direct::ssao.H = a->H; fs_direct::ssao.fmt = R16F;
        std::memcpy(&fs_direct::invProjectionView[0][0], a->invProjectionView, 64);
        fs_direct::windowWidth = (float)a->W; fs_direct::windowHeight = (float)a->H;
        fs_direct::viewPosition = glm::vec3(a->viewPosition[0], a->viewPosition[1], a->viewPosition[2]);
        fs_direct::emissionStrength = a->emissionStrength;
        fs_direct::numLights = a->numLights; fs_direct::numShadowLights = a->numShadowLights;
        fs_direct::nonShadowCasters = (const fs_direct::PointLight*)a->nonShadowCasters; fs_direct::shadowCasters = (const fs_direct::PointLight*)a->shadowCasters;
        fs_direct::shadowIndices = (const AtlasEntry*)a->shadowIndices;
        for (int i = 0; i < a->numAtlases && i < MAX_TOTAL_SHADOW_ATLASES; ++i) { fs_direct::shadowCubeMaps[i].data = a->shadowAtlases[i]; fs_direct::shadowCubeMaps[i].res = a->cubeRes; }
        fs_direct::infiniteLightColor = glm::vec3(a->sunColor[0], a->sunColor[1], a->sunColor[2]);
        infiniteLightDirection = glm::vec3(a->sunDir[0], a->sunDir[1], a->sunDir[2]);
        infiniteLightDepthBias = a->depthBias; infiniteLightZnear = a->zNear; infiniteLightZfar = a->zFar;
        for (int c = 0; c < 4; ++c) std::memcpy(&cascadeProjViews[c][0][0], a->cascadeProjViews + 16 * c, 64);
        for (int c = 0; c < 3; ++c) cascadePlanes[c] = glm::vec4(a->cascadePlanes[4 * c], a->cascadePlanes[4 * c + 1], a->cascadePlanes[4 * c + 2], a->cascadePlanes[4 * c + 3]);
        infiniteLightShadowMap.data = a->csm; infiniteLightShadowMap.res = a->csmRes; infiniteLightShadowMap.layers = 4;
        #pragma omp for schedule(dynamic, 2)
        for (int y = a->rowBegin; y < a->rowEnd; ++y)
            for (int x = 0; x < a->W; ++x) {
                fs_direct::fsTexCoords = glm::vec2(((float)x + 0.5f) / (float)a->W, ((float)y + 0.5f) / (float)a->H);
                fs_direct::direct_main();
                const size_t p = (size_t)y * a->W + x;
                for (int c = 0; c < 3; ++c) a->colorOut3[3 * p + c] = fs_direct::fsColor[c];
            }
    }
}
typedef struct RefSsaoArgs {
    int W, H; const unsigned short* structureH4; const unsigned short* rotationH4 /* 4x4 RGBA16F */; const unsigned short* occlusionH /* blur only */;
    float aspectRatio, projPlaneZDist, intensity; float* out; int rowBegin, rowEnd;
} RefSsaoArgs;
void refglsl_ssao(const RefSsaoArgs* a) {                                   // RenderSsaoOcclude_ (BE:1160-1187)
    using namespace refglsl;
    #pragma omp parallel
    {
        fs_ssao::structureBuffer.data = a->structureH4; fs_ssao::structureBuffer.W = a->W; fs_ssao::structureBuffer.H = a->H;
        fs_ssao::structureBuffer.fmt = RGBA16F; fs_ssao::structureBuffer.filter = LINEAR;                       // BE:406-408
        fs_ssao::rotationLookup.data = a->rotationH4; fs_ssao::rotationLookup.W = 4; fs_ssao::rotationLookup.H = 4;
        fs_ssao::rotationLookup.fmt = RGBA16F; fs_ssao::rotationLookup.filter = NEAREST; fs_ssao::rotationLookup.repeat = true;   // BE:776-778
        fs_ssao::aspectRatio = a->aspectRatio; fs_ssao::projPlaneZDist = a->projPlaneZDist; fs_ssao::intensity = a->intensity;
        fs_ssao::windowWidth = (float)a->W; fs_ssao::windowHeight = (float)a->H;
        #pragma omp for schedule(dynamic, 4)
        for (int y = a->rowBegin; y < a->rowEnd; ++y)
            for (int x = 0; x < a->W; ++x) {
                fs_ssao::fsTexCoords = glm::vec2(((float)x + 0.5f) / (float)a->W, ((float)y + 0.5f) / (float)a->H);
                fs_ssao::ssao_main();
                a->out[(size_t)y * a->W + x] = fs_ssao::gLightFactor;
            }
    }
}
void refglsl_ssao_blur(const RefSsaoArgs* a) {                              // RenderSsaoBlur_ (BE:1189-1204)
    using namespace refglsl;
    #pragma omp parallel
    {
        fs_ssao_blur::structureBuffer.data = a->structureH4; fs_ssao_blur::structureBuffer.W = a->W; fs_ssao_blur::structureBuffer.H = a->H;
        fs_ssao_blur::structureBuffer.fmt = RGBA16F; fs_ssao_blur::structureBuffer.filter = LINEAR;
        fs_ssao_blur::occlusionBuffer.data = a->occlusionH; fs_ssao_blur::occlusionBuffer.W = a->W; fs_ssao_blur::occlusionBuffer.H = a->H;
        fs_ssao_blur::occlusionBuffer.fmt = R16F; fs_ssao_blur::occlusionBuffer.filter = NEAREST;              // BE:491-493
        fs_ssao_blur::windowWidth = (float)a->W; fs_ssao_blur::windowHeight = (float)a->H;
        #pragma omp for schedule(dynamic, 4)
        for (int y = a->rowBegin; y < a->rowEnd; ++y)
            for (int x = 0; x < a->W; ++x) {
                fs_ssao_blur::fsTexCoords = glm::vec2(((float)x + 0.5f) / (float)a->W, ((float)y + 0.5f) / (float)a->H);
                fs_ssao_blur::ssao_blur_main();
                a->out[(size_t)y * a->W + x] = fs_ssao_blur::gLightFactor;
            }
    }
}
typedef struct RefTaaArgs { int W, H; const unsigned char* screen4; const unsigned char* prevScreen4; const unsigned short* velocityH2; float* out4; } RefTaaArgs;
void refglsl_taa(const RefTaaArgs* a) {                                     // PerformTaaPostFx_ (BE:2324-2344), the shader part
    using namespace refglsl;
    #pragma omp parallel
    {
        auto tex = [&](const void* d, TexFormat f, TexFilter filt) { sampler2D s; s.data = d; s.W = a->W; s.H = a->H; s.fmt = f; s.filter = filt; return s; };
        fs_taa::screen = tex(a->screen4, RGBA8, LINEAR);                     // fxaaFbo2: RGBA8, LINEAR (BE:694-696)
        fs_taa::prevScreen = tex(a->prevScreen4, RGBA8, LINEAR);             // previousFrameBuffer: RGB8, LINEAR (BE:450-452); alpha is never read
        fs_taa::velocity = tex(a->velocityH2, RG16F, NEAREST);               // BE:412-414
        #pragma omp for schedule(dynamic, 4)
        for (int y = 0; y < a->H; ++y)
            for (int x = 0; x < a->W; ++x) {
                fs_taa::fsTexCoords = glm::vec2(((float)x + 0.5f) / (float)a->W, ((float)y + 0.5f) / (float)a->H);
                fs_taa::taa_main();
                for (int c = 0; c < 4; ++c) a->out4[4 * ((size_t)y * a->W + x) + c] = fs_taa::color[c];
            }
    }
}
// shadowVpl.fs main() for one fragment: world position, light, far plane, base colour -> gl_FragDepth and the colour output
void refglsl_shadow_vpl(const float* fsPosition3, const float* lightPos3, float farPlane, const float* baseColor4, float* depthOut, float* color3Out) {
    using namespace refglsl;
    fs_shadow_vpl::materials[0].diffuseColor = glm::vec4(baseColor4[0], baseColor4[1], baseColor4[2], baseColor4[3]);
    fs_shadow_vpl::materials[0].flags = 0u; fs_shadow_vpl::materialIndices[0] = 0; fs_shadow_vpl::fsDrawID = 0;
    fs_shadow_vpl::fsPosition = glm::vec4(fsPosition3[0], fsPosition3[1], fsPosition3[2], 1.0f);
    fs_shadow_vpl::lightPos = glm::vec3(lightPos3[0], lightPos3[1], lightPos3[2]); fs_shadow_vpl::farPlane = farPlane;
    fs_shadow_vpl::shadow_vpl_main();
    *depthOut = fs_shadow_vpl::gl_FragDepth;
    for (int c = 0; c < 3; ++c) color3Out[c] = fs_shadow_vpl::color[c];
}
typedef struct RefColorArgs {
    int cubeRes, numAtlases; const float* sunDir; const float* sunColor; void* lightData; const int* numVisible; const void* diffuseIndices;
    const unsigned short* const* shadowAtlases; const unsigned char* const* diffuseAtlases;
} RefColorArgs;
void refglsl_color(const RefColorArgs* a) {
    using namespace refglsl;
    #pragma omp parallel
    {
        infiniteLightDirection = glm::vec3(a->sunDir[0], a->sunDir[1], a->sunDir[2]);
        cs_color::infiniteLightColor = glm::vec3(a->sunColor[0], a->sunColor[1], a->sunColor[2]);
        cs_color::lightData = (VplData*)a->lightData; cs_color::numVisible = a->numVisible; cs_color::diffuseIndices = (const AtlasEntry*)a->diffuseIndices;
        for (int i = 0; i < a->numAtlases && i < MAX_TOTAL_SHADOW_ATLASES; ++i) {
            cs_color::shadowCubeMaps[i].data = a->shadowAtlases[i]; cs_color::shadowCubeMaps[i].res = a->cubeRes; cs_color::shadowCubeMaps[i].rgba8 = nullptr;
            cs_color::diffuseCubeMaps[i].data = nullptr; cs_color::diffuseCubeMaps[i].res = a->cubeRes; cs_color::diffuseCubeMaps[i].rgba8 = a->diffuseAtlases[i];
        }
        gl_NumWorkGroups = glm::uvec3(1, 1, 1); gl_WorkGroupSize = glm::uvec3(1024, 1, 1); gl_WorkGroupID = glm::uvec3(0);
        #pragma omp for schedule(static)
        for (int g = 0; g < 1024; ++g) {
            gl_GlobalInvocationID = gl_LocalInvocationID = glm::uvec3((unsigned)g, 0, 0); gl_LocalInvocationIndex = (unsigned)g;
            cs_color::color_main();
        }
    }
}
typedef struct RefStage1Args { int W, H; const float* depth; const unsigned char* normal4; const float* invProjectionView; void* out; int tileRowBegin, tileRowEnd; } RefStage1Args;
void refglsl_stage1(const RefStage1Args* a) {
    using namespace refglsl;
    const unsigned tilesX = a->W / 5, tilesY = a->H / 5;                    // DispatchCompute(viewportWidth / tileXDivisor, viewportHeight / tileYDivisor, 1)
    #pragma omp parallel
    {
        auto tex = [&](const void* d, TexFormat f) { sampler2D s; s.data = d; s.W = a->W; s.H = a->H; s.fmt = f; return s; };
        cs_stage1::gDepth = tex(a->depth, R32F); cs_stage1::gNormal = tex(a->normal4, RGBA8);
        std::memcpy(&cs_stage1::invProjectionView[0][0], a->invProjectionView, 64);
        cs_stage1::vplNumVisiblePerTile = (VplStage1PerTileOutputs*)a->out;
        gl_NumWorkGroups = glm::uvec3(tilesX, tilesY, 1); gl_WorkGroupSize = glm::uvec3(5, 5, 1);
        #pragma omp for schedule(dynamic, 1)
        for (int ty = a->tileRowBegin; ty < a->tileRowEnd; ++ty)
            for (unsigned tx = 0; tx < tilesX; ++tx) {
                gl_WorkGroupID = glm::uvec3(tx, (unsigned)ty, 0);
                for (unsigned k = 1; k <= 25; ++k) {                         // invocations 1..24, then invocation 0 (the one that reads the shared arrays)
                    const unsigned idx = k % 25, lx = idx % 5, ly = idx / 5;
                    gl_LocalInvocationID = glm::uvec3(lx, ly, 0); gl_LocalInvocationIndex = idx;
                    gl_GlobalInvocationID = glm::uvec3(tx * 5 + lx, (unsigned)ty * 5 + ly, 0);
                    cs_stage1::stage1_main();
                }
            }
    }
}
typedef struct RefStage2Args {
    int tilesX, tilesY; const float* viewPosition; const void* lightData; const void* stage1Data; int numVisible; const int* vplVisibleIndex; void* out;
    int tileRowBegin, tileRowEnd;
} RefStage2Args;
void refglsl_stage2(const RefStage2Args* a) {
    using namespace refglsl;
    #pragma omp parallel
    {
        cs_stage2::viewPosition = glm::vec3(a->viewPosition[0], a->viewPosition[1], a->viewPosition[2]);
        cs_stage2::lightData = (const VplData*)a->lightData; cs_stage2::stage1Data = (const VplStage1PerTileOutputs*)a->stage1Data;
        cs_stage2::numVisible = a->numVisible; cs_stage2::vplVisibleIndex = a->vplVisibleIndex; cs_stage2::vplNumVisiblePerTile = (VplStage2PerTileOutputs*)a->out;
        gl_WorkGroupSize = glm::uvec3(32, 2, 1); gl_NumWorkGroups = glm::uvec3(a->tilesX / 32, a->tilesY / 2, 1);      // BE:1774-1778
        #pragma omp for schedule(dynamic, 1)
        for (int gy = a->tileRowBegin; gy < a->tileRowEnd; ++gy)
            for (unsigned gx = 0; gx < gl_NumWorkGroups.x * 32; ++gx) {
                gl_GlobalInvocationID = glm::uvec3(gx, (unsigned)gy, 0); gl_WorkGroupID = glm::uvec3(gx / 32, (unsigned)gy / 2, 0);
                gl_LocalInvocationID = glm::uvec3(gx % 32, (unsigned)gy % 2, 0); gl_LocalInvocationIndex = ((unsigned)gy % 2) * 32 + gx % 32;
                cs_stage2::stage2_main();
            }
    }
}
typedef struct RefCullArgs {
    int totalNumLights, csmRes; const float* sunDir; const float* cascadeProjViews; const float* cascadePlanes; const float* csm;
    const void* lightData; void* updatedLightData; int* vplVisibleIndex;
} RefCullArgs;
// viscull_vpls.cs is ONE work group of 1024 invocations with four barrier()s (BE:1603).  The invocations run as coroutines
// (ucontext) spread over the OpenMP pool: barrier() yields to the pool thread's scheduler, which resumes its next invocation and,
// once all of its invocations have arrived, meets the other pool threads at a real barrier.  (An earlier version created 1024 OS
// threads per frame, which cost more than the shader's own work.)
struct CullCo { ucontext_t ctx; unsigned t; bool done; std::unique_ptr<char[]> stack; };
static const size_t kCullStack = 128 * 1024;
static thread_local ucontext_t* g_coSched = nullptr;
static thread_local CullCo* g_coCur = nullptr;
static int g_cullAlive = 0;
static void cullCoYield() { swapcontext(&g_coCur->ctx, g_coSched); }
static void cullCoEntry() {
    refglsl::cs_cull::cull_main();
    g_coCur->done = true;
    swapcontext(&g_coCur->ctx, g_coSched);
}
void refglsl_cull(const RefCullArgs* a) {
    using namespace refglsl;
    cs_cull::totalNumLights = a->totalNumLights; cs_cull::lightData = (const VplData*)a->lightData;
    cs_cull::updatedLightData = (VplData*)a->updatedLightData; cs_cull::vplVisibleIndex = a->vplVisibleIndex;
    const unsigned n = 1024;                                                 // one work group of local_size_x = 1024 (BE:1603)
    barrierHook = cullCoYield;
    #pragma omp parallel
    {
        // the uniforms are per thread (see PERINV): every pool thread sets its own copy
        infiniteLightDirection = glm::vec3(a->sunDir[0], a->sunDir[1], a->sunDir[2]);
        for (int c = 0; c < 4; ++c) std::memcpy(&cascadeProjViews[c][0][0], a->cascadeProjViews + 16 * c, 64);
        for (int c = 0; c < 3; ++c) cascadePlanes[c] = glm::vec4(a->cascadePlanes[4 * c], a->cascadePlanes[4 * c + 1], a->cascadePlanes[4 * c + 2], a->cascadePlanes[4 * c + 3]);
        infiniteLightShadowMap.data = a->csm; infiniteLightShadowMap.res = a->csmRes; infiniteLightShadowMap.layers = 4;
        gl_NumWorkGroups = glm::uvec3(1, 1, 1); gl_WorkGroupSize = glm::uvec3(1024, 1, 1); gl_WorkGroupID = glm::uvec3(0);
        const unsigned nt = (unsigned)omp_get_num_threads(), id = (unsigned)omp_get_thread_num();
        std::vector<CullCo> cos((n + nt - 1 - id) / nt);
        ucontext_t sched;
        g_coSched = &sched;
        for (size_t k = 0; k < cos.size(); ++k) {
            CullCo& co = cos[k];
            co.t = id + (unsigned)k * nt; co.done = false; co.stack.reset(new char[kCullStack]);             // uninitialised on purpose
            getcontext(&co.ctx);
            co.ctx.uc_stack.ss_sp = co.stack.get(); co.ctx.uc_stack.ss_size = kCullStack; co.ctx.uc_link = &sched;
            makecontext(&co.ctx, cullCoEntry, 0);
        }
        for (;;) {
            #pragma omp single
            g_cullAlive = 0;
            int alive = 0;
            for (CullCo& co : cos) {
                if (co.done) continue;
                g_coCur = &co;
                gl_LocalInvocationID = gl_GlobalInvocationID = glm::uvec3(co.t, 0, 0); gl_LocalInvocationIndex = co.t;
                swapcontext(&sched, &co.ctx);                  // runs until the invocation's next barrier() or its end
                if (!co.done) ++alive;
            }
            #pragma omp atomic
            g_cullAlive += alive;
            #pragma omp barrier
            const bool more = g_cullAlive != 0;                // the same on every pool thread (barriers sit in uniform control flow)
            #pragma omp barrier
            if (!more) break;
        }
        g_coSched = nullptr; g_coCur = nullptr;
    }
    barrierHook = nullptr;
}
}
'''


CS_PRE = {'fs_shadow_vpl': ('// what the material system and alpha_test.glsl supply, reduced to the untextured, opaque case the box scenes are\n'
                            'struct Material { vec4 diffuseColor; unsigned flags; sampler2D diffuseMap; };\n'
                            'static const unsigned GPU_DIFFUSE_MAPPED = 2u;\n'
                            'static inline bool bitwiseAndBool(unsigned a, unsigned b) { return (a & b) != 0u; }\n'
                            'static inline vec4 FLOAT4_TO_VEC4(vec4 v) { return v; }\n'
                            'static inline void runAlphaTest(float) {}\n'
                            'PERINV Material materials[1]; PERINV int materialIndices[1]; PERINV int fsDrawID;\n'
                            'PERINV vec4 fsPosition; PERINV vec2 fsTexCoords; PERINV vec3 lightPos; PERINV float farPlane; PERINV float gl_FragDepth; PERINV vec3 color;\n'),
          'fs_ssao': 'PERINV sampler2DRect structureBuffer; PERINV sampler2D rotationLookup;\nPERINV float aspectRatio, projPlaneZDist, windowHeight, windowWidth, intensity; PERINV vec2 fsTexCoords; PERINV float gLightFactor;\n',
          'fs_taa': 'PERINV sampler2D screen, prevScreen, velocity, previousVelocity; PERINV vec2 fsTexCoords; PERINV vec4 color;\n',
          'fs_ssao_blur': 'PERINV sampler2DRect structureBuffer, occlusionBuffer; PERINV float windowWidth, windowHeight; PERINV vec2 fsTexCoords; PERINV float gLightFactor;\n',
          'cs_color': 'PERINV vec3 infiniteLightColor;\nPERINV VplData* lightData; PERINV const int* numVisible; PERINV const AtlasEntry* diffuseIndices;\nPERINV samplerCubeArray diffuseCubeMaps[MAX_TOTAL_SHADOW_ATLASES], shadowCubeMaps[MAX_TOTAL_SHADOW_ATLASES];\n', 'cs_stage1': 'PERINV sampler2D gDepth, gNormal; PERINV mat4 invProjectionView;\nPERINV VplStage1PerTileOutputs* vplNumVisiblePerTile;\nPERINV vec3 localPositions[25], localNormals[25];          // `shared`, sized gl_WorkGroupSize.x * gl_WorkGroupSize.y = 5 * 5\n', 'cs_stage2': 'PERINV vec3 viewPosition;\nPERINV const VplData* lightData; PERINV const VplStage1PerTileOutputs* stage1Data; PERINV int numVisible; PERINV const int* vplVisibleIndex;\nPERINV VplStage2PerTileOutputs* vplNumVisiblePerTile;\n', 'cs_cull': 'static int totalNumLights;\nstatic const VplData* lightData; static VplData* updatedLightData; static int* vplVisibleIndex;\nstatic bool lightVisible[MAX_TOTAL_VPLS_BEFORE_CULLING]; static int lightVisibleIndex;      // `shared`\n', 'fs_direct': '#define INFINITE_LIGHTING_ENABLED\nstruct PointLight { vec4 position; vec4 color; float radius; float farPlane; float _1[2]; };        // pbr.fs:15-21\nPERINV sampler2D gDepth, gNormal, gAlbedo, gBaseReflectivity, gRoughnessMetallicAmbient; PERINV sampler2DRect ssao;\nPERINV mat4 invProjectionView; PERINV float windowWidth, windowHeight; PERINV vec3 viewPosition; PERINV float emissionStrength = 0.0f;\nPERINV int numLights = 0, numShadowLights = 0; PERINV const PointLight* nonShadowCasters; PERINV const PointLight* shadowCasters;\nPERINV samplerCubeArray shadowCubeMaps[MAX_TOTAL_SHADOW_ATLASES]; PERINV const AtlasEntry* shadowIndices;\nPERINV vec3 infiniteLightColor; PERINV vec2 fsTexCoords; PERINV vec3 fsColor;\n'}


def generate():
    parts = [PRELUDE]
    common = strip_comments(open(os.path.join(SHADERS, "common.glsl")).read())
    for d in DEFINES:
        m = re.search(r"^#define\s+" + d + r"\s+(\S+)\s*$", common, flags=re.M)
        parts.append(adapt(f"static const float {d} = {m.group(1)};\n"))
    for fname, names in SHADER_DEFINES:
        text = strip_comments(open(os.path.join(SHADERS, fname)).read())
        for d in names:
            m = re.search(r"^#define\s+" + d + r"\s+(\S+)\s*$", text, flags=re.M)
            parts.append(f"#define {d} {m.group(1)}\n")
    for fname, names in SHADER_CONSTS:
        text = strip_comments(open(os.path.join(SHADERS, fname)).read())
        for c in names:
            m = re.search(r"^const\s+(float|int)\s+" + c + r"\s*=\s*([^;]+);", text, flags=re.M)
            parts.append(adapt(f"static const {m.group(1)} {c} = {m.group(2)};\n"))
    for fname, names in WANTED:
        text = strip_comments(open(os.path.join(SHADERS, fname)).read())
        parts.append(f"\n// ---- from Source/Shaders/{fname} ----\n")
        for n in names:
            defs = extract(text, n)
            if not defs:
                raise SystemExit(f"{n} not found in {fname}")
            if n == "main":
                defs = [re.sub(r"\bvoid\s+main\s*\(", "void denoise_main(", d) for d in defs]
            parts += [adapt(d) + "\n" for d in defs]
    for fname, ns, main_name, macros, *more in COMPUTE:
        text = strip_comments(open(os.path.join(SHADERS, fname)).read())
        parts.append(f"\n// ---- from Source/Shaders/{fname} ----\nnamespace {ns} {{\n" + CS_PRE[ns])
        for mac in macros:
            parts.append(extract_macro(text, mac) + "\n")
        for n in (more[0] if more else []):               # the shader's own functions main() calls
            defs = extract(text, n)
            if not defs:
                raise SystemExit(f"{n} not found in {fname}")
            parts += [adapt(d) + "\n" for d in defs]
        (d,) = extract(text, "main")
        parts.append(adapt(re.sub(r"\bvoid\s+main\s*\(", f"void {main_name}(", d)) + f"\n}}  // namespace {ns}\n")
    parts.append(EPILOGUE)
    return "".join(parts)


def build(force=False):
    if not os.path.isdir(SHADERS):
        return None
    os.makedirs(OUT_DIR, exist_ok=True)
    src, lib = os.path.join(OUT_DIR, "ref_glsl_gen.cpp"), os.path.join(OUT_DIR, "libref_glsl.so")
    if not force and os.path.exists(lib) and os.path.getmtime(lib) >= os.path.getmtime(os.path.abspath(__file__)):
        return lib
    open(src, "w").write(generate())
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-fopenmp", "-ftls-model=local-dynamic", "-ffp-contract=off",
                           "-I" + REF, "-o", lib, src])
    if os.environ.get("SGI_KEEP_GENERATED", "") != "1":
        os.remove(src)                  # the adapted shader text is reference material: only the compiled library stays on disk
    return lib


if __name__ == "__main__":
    print(build(force=True))
