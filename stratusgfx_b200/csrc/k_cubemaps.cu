// k_cubemaps.cu — VPL cube shadow / diffuse map production.
//
// Replaces the per-light loop of UpdatePointLights_ (StratusRendererBackend.cpp:1426-1575): clear the light's cube layer
// (depth 1, colour 0, :1452-1456), draw the scene six times through the 90-degree views of GenerateLightViewTransforms with
// shadow.vs / shadowVpl.fs, whose fragment stage is
//     lightDistance = saturate(length(fsPosition.xyz - lightPos) / farPlane);  gl_FragDepth = lightDistance;  color = baseColor.rgb
// (shadowVpl.fs:27-35), into a 16-bit depth + RGB8 cube-map array layer (BE:2413-2421).
//
// The scenes of this repository are lists of axis-aligned boxes, so the fragment of a texel is found by casting the ray through
// the texel centre — the inverse of the cube-face table the gather's lookup uses (sgi_math.cuh cubeTexel),
// so a direction that the gather maps to texel (face, j, i) is answered by the surface that texel's own ray hit — instead of by
// rasterising triangles.  Four CTAs render one face of one light: the box list sits in shared memory, a thread owns one texel of
// the CTA's quarter (more at larger faces); a warp's 32 texels are one texel row, so both stores are coalesced (64 B + 128 B).
//
// All arithmetic is the exact tier (sgi_math.cuh: explicit round-to-nearest multiplies / adds / IEEE division and square root in
// the oracle's order, no FMA): every texel is an integer output, compared bit for bit with oracle.cpp::orc_render_vpl_cubemaps.
// The work is tiny — 5 lights x 6 faces x 1024 texels x <= 64 boxes per frame in the reference's budget — so the kernel is
// launch-latency bound; refreshing all 2048 lights of workload c4 is 12.6 M rays.
#include "sgi_internal.h"
#include "sgi_math.cuh"

namespace sgi {
namespace {

constexpr float kFloatMax = 3.402823466e+38f;
constexpr int kCubeThreads = 256;
constexpr int kCubeSlices = 4;          // CTAs per face: one texel per thread at the reference's 32 x 32 faces (the ray loop is the kernel's
                                        // latency), and CTAs small enough to find a slot next to the frame's own kernels (sgi_early_begin)

SGI_DEV f3 cubeTexelDirection(int face, float sc, float tc) {     // inverse of GL 4.6 table 8.19
    switch (face) {
        case 0: return mk3(1.0f, -tc, -sc);
        case 1: return mk3(-1.0f, -tc, sc);
        case 2: return mk3(sc, 1.0f, tc);
        case 3: return mk3(sc, -1.0f, -tc);
        case 4: return mk3(sc, -tc, 1.0f);
        default: return mk3(-sc, -tc, -1.0f);
    }
}

struct SBox { float lo[3], hi[3]; uint32_t albedo; int insideOut; };
static_assert(sizeof(SBox) == sizeof(SgiSceneBox), "SgiSceneBox is 32 bytes");

__global__ void __launch_bounds__(kCubeThreads)
cubemapKernel(const SgiSceneBox* __restrict__ boxes, int numBoxes, const SgiVplData* __restrict__ lightData, const SgiAtlasEntry* __restrict__ atlasByVpl,
              const int32_t* __restrict__ vplIndices, int res, const __grid_constant__ AtlasTable atlases) {
    __shared__ SBox sBox[SGI_MAX_SCENE_BOXES];
    for (int b = threadIdx.x; b < numBoxes; b += kCubeThreads) {
        const uint4* src = reinterpret_cast<const uint4*>(boxes + b);
        const uint4 a = __ldg(src), c = __ldg(src + 1);
        SBox s;
        s.lo[0] = __uint_as_float(a.x); s.lo[1] = __uint_as_float(a.y); s.lo[2] = __uint_as_float(a.z);
        s.hi[0] = __uint_as_float(a.w); s.hi[1] = __uint_as_float(c.x); s.hi[2] = __uint_as_float(c.y);
        s.albedo = c.z; s.insideOut = (int)c.w;
        sBox[b] = s;
    }
    __syncthreads();
    const int slice = blockIdx.x % kCubeSlices, mf = blockIdx.x / kCubeSlices;
    const int k = __ldg(vplIndices + mf / 6), face = mf % 6;
    const float4* L = reinterpret_cast<const float4*>(lightData + k);
    const float4 lp = __ldg(L), rfi = __ldg(L + 3);
    const float o[3] = {lp.x, lp.y, lp.z};
    const float farPlane = rfi.y;
    const SgiAtlasEntry e = atlasByVpl[k];
    const size_t base = ((size_t)e.layer * 6 + face) * res * res;
    unsigned short* __restrict__ depth = const_cast<unsigned short*>(atlases.shadow[e.index]) + base;
    uchar4* __restrict__ diffuse = const_cast<uchar4*>(atlases.diffuse[e.index]) + base;
    const float fres = (float)res;
    const int perSlice = (res * res + kCubeSlices - 1) / kCubeSlices;
    const int qEnd = min(res * res, (slice + 1) * perSlice);
    for (int q = slice * perSlice + threadIdx.x; q < qEnd; q += kCubeThreads) {
        const int j = q / res, i = q - j * res;
        const float sc = fsub(fmul(2.0f, fdiv(fadd((float)i, 0.5f), fres)), 1.0f), tc = fsub(fmul(2.0f, fdiv(fadd((float)j, 0.5f), fres)), 1.0f);
        const f3 dv = cubeTexelDirection(face, sc, tc);
        const float d[3] = {dv.x, dv.y, dv.z};
        float inv[3];
        #pragma unroll
        for (int a = 0; a < 3; ++a) inv[a] = d[a] == 0.0f ? 0.0f : fdiv(1.0f, d[a]);
        float best = kFloatMax; int hit = -1;
        for (int b = 0; b < numBoxes; ++b) {
            const SBox& bx = sBox[b];
            float t0 = -kFloatMax, t1 = kFloatMax; bool miss = false;
            #pragma unroll
            for (int a = 0; a < 3; ++a) {
                if (d[a] == 0.0f) { if (o[a] < bx.lo[a] || o[a] > bx.hi[a]) miss = true; continue; }
                float ta = fmul(fsub(bx.lo[a], o[a]), inv[a]), tb = fmul(fsub(bx.hi[a], o[a]), inv[a]);
                if (ta > tb) { const float tmp = ta; ta = tb; tb = tmp; }
                if (ta > t0) t0 = ta;
                if (tb < t1) t1 = tb;
            }
            if (miss || t0 > t1) continue;
            const float t = bx.insideOut ? t1 : t0;
            if (t > 1e-6f && t < best) { best = t; hit = b; }
        }
        if (hit < 0) {                      // nothing drawn: the cleared texel (BE:1452-1456)
            depth[q] = 65535; diffuse[q] = make_uchar4(0, 0, 0, 0);
            continue;
        }
        const f3 origin = mk3(o[0], o[1], o[2]);
        const f3 fsPosition = origin + dv * best;
        // shadowVpl.fs:27-30
        float lightDistance = length(fsPosition - origin);
        lightDistance = saturate(fdiv(lightDistance, farPlane));
        depth[q] = (unsigned short)floorf(fadd(fmul(lightDistance, 65535.0f), 0.5f));          // unorm16, round to nearest
        const uint32_t alb = sBox[hit].albedo;
        diffuse[q] = make_uchar4((unsigned char)(alb & 0xffu), (unsigned char)((alb >> 8) & 0xffu), (unsigned char)((alb >> 16) & 0xffu), 255);
    }
}

}  // namespace

bool preloadCubemapKernels() {
    cudaFuncAttributes a;
    return cudaFuncGetAttributes(&a, cubemapKernel) == cudaSuccess;
}

int launchCubemaps(cudaStream_t stream, const SgiCubeRenderParams& rp, const SgiVplData* lightData, const SgiAtlasEntry* atlasByVpl, int res, const AtlasTable& atlases) {
    if (rp.numMaps <= 0) return 0;
    cubemapKernel<<<rp.numMaps * 6 * kCubeSlices, kCubeThreads, 0, stream>>>(rp.boxes, rp.numBoxes, lightData, atlasByVpl, rp.vplIndices, res, atlases);
    return 1;
}

}  // namespace sgi
