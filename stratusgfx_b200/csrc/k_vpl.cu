// k_vpl.cu — VPL sun-visibility cull + ordered compaction (a1) and VPL colouring (a2).
//
// a1 replaces viscull_vpls.cs:47-106 (+ pbr.glsl:139-147, 158-227).  The reference runs ONE
// 1024-thread work group and compacts with 8 helper threads racing on a shared atomic, then the
// CPU blocks on a fence to read the list back (StratusRendererBackend.cpp:1604-1610, 1675-1701).
// Here: one thread per VPL across as many CTAs as needed, warp ballots written to global memory,
// and the last CTA to finish (atomic ticket) turns the ballots into an exclusive prefix sum and
// writes the compacted VplData / index list / atlas entries in ascending original index — one
// legal interleaving of the reference's race, and the order of its serial variant (:108-119).
// No host round trip: the atlas entry of each survivor is compacted on the device too.
//
// a2 replaces vpl_light_color.cs:41-76.
#include <cstring>
#include "sgi_internal.h"
#include "sgi_math.cuh"

namespace sgi {

namespace {

constexpr int kCullThreads = 1024;         // 32 warps = 32 VPLs = one ballot word per CTA

// hardware PCF of a sampler2DArrayShadow (LINEAR, compare LEQUAL): SURVEY.md Appendix A4
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

// pbr.glsl:158-227 with useDepthBias = false and normal = infiniteLightDirection (viscull_vpls.cs:68), one WARP per VPL:
// lane l < 18 evaluates ONE of the 18 PCF taps (lanes 0-8: the 3x3 taps of the first cascade, 9-17: of the second), so the
// 72 dependent-latency texel fetches of a VPL are in flight together; the taps are then summed in the shader's order
// (y outer, x inner, sequential adds) so the result is bit-identical to the serial loop.
SGI_DEV float calculateInfiniteShadowValueWarp(const FrameConsts& fc, const float* __restrict__ csm, int csmRes, f3 lightPos, f3 cascadeBlends, int lane) {
    const f3 sun = mk3(fc.sunDir[0], fc.sunDir[1], fc.sunDir[2]);
    const f3 normal = sun;
    const f3 off = (normal * fsub(1.0f, saturate(dot(normal, sun)))) * 1.0f;
    const f4 position = mk4(fadd(lightPos.x, off.x), fadd(lightPos.y, off.y), fadd(lightPos.z, off.z), 1.0f);
    const bool beyondCascade2 = cascadeBlends.y >= 0.0f;
    const bool beyondCascade3 = cascadeBlends.z >= 0.0f;
    const int index1 = beyondCascade2 ? 2 : 0;
    const int index2 = beyondCascade3 ? 3 : 1;
    const bool second = lane >= 9;
    const int idx = second ? index2 : index1;
    const m4 M = loadMat(fc.cascadeProjViews[idx]);
    const f4 coords = mul(M, position);
    const f3 cc = (mk3(coords.x, coords.y, coords.z) / coords.w) * 0.5f + 0.5f;
    const float depth = (second && beyondCascade3) ? saturate(cc.z) : cc.z;
    const float weight = beyondCascade2 ? fsub(saturate(cascadeBlends.y), saturate(cascadeBlends.z)) : fsub(1.0f, saturate(cascadeBlends.x));
    const float wh = fdiv(1.0f, (float)csmRes);
    const int t = second ? lane - 9 : lane;                 // tap index: (iy+1)*3 + (ix+1)
    float tap = 0.0f;
    if (lane < 18) {
        const float ox = fmul((float)(t % 3 - 1), wh), oy = fmul((float)(t / 3 - 1), wh);
        tap = shadowArrayPcf(csm, csmRes, fadd(cc.x, ox), fadd(cc.y, oy), idx, depth);
    }
    float light1 = 0.0f, light2 = 0.0f;
    #pragma unroll
    for (int k = 0; k < 9; ++k) {
        light1 = fadd(light1, __shfl_sync(0xffffffffu, tap, k));
        light2 = fadd(light2, __shfl_sync(0xffffffffu, tap, 9 + k));
    }
    return fmul(mix(light2, light1, weight), fdiv(1.0f, 9.0f));
}

// vpl_light_color.cs:41-76 for one visible VPL
SGI_DEV void colorOneVpl(const FrameConsts& fc, int cubeRes, int texel, SgiVplData* __restrict__ lightData, int index,
                         const SgiAtlasEntry* __restrict__ diffuseIndices, const AtlasTable& atlases) {
    const float colorMultiplier = 50000.0f;
    const f3 sun = mk3(fc.sunDir[0], fc.sunDir[1], fc.sunDir[2]);
    const f3 sunColor = mk3(fc.sunColor[0], fc.sunColor[1], fc.sunColor[2]);
    const float4 pos = reinterpret_cast<const float4*>(lightData + index)[0];
    const float4 rfi = reinterpret_cast<const float4*>(lightData + index)[3];   // radius, farPlane, intensity, pad
    const SgiAtlasEntry entry = diffuseIndices[index];
    const size_t t = (size_t)entry.layer * 6 * cubeRes * cubeRes + texel;
    const uchar4 d = __ldg(atlases.diffuse[entry.index] + t);
    const f3 color = mk3(unorm8(d.x), unorm8(d.y), unorm8(d.z)) * sunColor;
    const float depth = fdiv((float)__ldg(atlases.shadow[entry.index] + t), 65535.0f);
    const float magnitude = fmul(rfi.x, depth);
    const f3 specularPos = mk3(pos.x, pos.y, pos.z) - (fmul(1.1f, magnitude) * sun);
    const f3 c = ((color / 1.0f) * rfi.z) * colorMultiplier;
    reinterpret_cast<float4*>(lightData + index)[1] = make_float4(c.x, c.y, c.z, 1.0f);
    reinterpret_cast<float4*>(lightData + index)[2] = make_float4(specularPos.x, specularPos.y, specularPos.z, 1.0f);
}

// `colorToo`: the last CTA also colours the compacted VPLs (what the separate colorKernel launch would do), for sgi_frame
__global__ void __launch_bounds__(kCullThreads)
cullKernel(const __grid_constant__ FrameConsts fc, int csmRes, int cubeRes, const SgiVplData* __restrict__ lightData,
           const SgiAtlasEntry* __restrict__ atlasByVpl, const float* __restrict__ csm, SgiVplData* __restrict__ updated,
           int32_t* __restrict__ visible, SgiAtlasEntry* __restrict__ entries, uint32_t* __restrict__ ballots,
           unsigned int* __restrict__ ticket, const __grid_constant__ AtlasTable atlases, int colorToo, const int32_t* __restrict__ countDevice) {
    // the host's N may be an upper bound when the table was built on the device (sgi_vpl_prepass)
    const int N = countDevice != nullptr ? max(0, min(fc.totalNumLights, __ldg(countDevice))) : fc.totalNumLights;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int index = blockIdx.x * 32 + warp;                  // one warp per VPL
    __shared__ int sVis[32];
    bool vis = false;
    if (index < N) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(lightData + index));
        const f3 lightPos = mk3(p.x, p.y, p.z);
        const f4 lp4 = mk4(p.x, p.y, p.z, 1.0f);
        f3 blends;
        blends.x = dot(mk4(fc.cascadePlanes[0][0], fc.cascadePlanes[0][1], fc.cascadePlanes[0][2], fc.cascadePlanes[0][3]), lp4);
        blends.y = dot(mk4(fc.cascadePlanes[1][0], fc.cascadePlanes[1][1], fc.cascadePlanes[1][2], fc.cascadePlanes[1][3]), lp4);
        blends.z = dot(mk4(fc.cascadePlanes[2][0], fc.cascadePlanes[2][1], fc.cascadePlanes[2][2], fc.cascadePlanes[2][3]), lp4);
        const float shadowFactor = fsub(1.0f, calculateInfiniteShadowValueWarp(fc, csm, csmRes, lightPos, blends, lane));
        vis = shadowFactor < 1.0f;
    }
    if (lane == 0) sVis[warp] = vis ? 1 : 0;
    __syncthreads();
    if (warp == 0) {
        const uint32_t bal = __ballot_sync(0xffffffffu, sVis[lane] != 0);
        if (lane == 0) ballots[blockIdx.x] = bal;              // bit w of word b <-> VPL 32*b + w
    }

    // ---- last CTA to arrive does the ordered compaction for the whole list ----
    __shared__ bool isLast;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        isLast = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!isLast) return;
    __threadfence();

    __shared__ int prefix[(SGI_MAX_TOTAL_VPLS_BEFORE_CULLING + 31) / 32 + 1];
    __shared__ int warpSums[kCullThreads / 32];
    const int numWords = (N + 31) >> 5;
    // exclusive scan of popcounts over <= 313 words
    int carry = 0;
    for (int base = 0; base < numWords; base += kCullThreads) {
        const int w = base + threadIdx.x;
        const int c = w < numWords ? __popc(__ldcg(ballots + w)) : 0;
        int incl = c;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) warpSums[warp] = incl;
        __syncthreads();
        int warpBase = 0;
        for (int k = 0; k < warp; ++k) warpBase += warpSums[k];
        if (w < numWords) prefix[w] = carry + warpBase + incl - c;
        int total = 0;
        for (int k = 0; k < kCullThreads / 32; ++k) total += warpSums[k];
        carry += total;
        __syncthreads();
    }
    const int total = carry;
    for (int i = threadIdx.x; i < N; i += kCullThreads) {
        const uint32_t word = __ldcg(ballots + (i >> 5));
        if (!((word >> (i & 31)) & 1u)) continue;
        const int k = prefix[i >> 5] + __popc(word & ((1u << (i & 31)) - 1u));
        if (k >= SGI_MAX_TOTAL_VPLS_PER_FRAME) continue;
        const float4* src = reinterpret_cast<const float4*>(lightData + i);
        float4* dst = reinterpret_cast<float4*>(updated + k);
        dst[0] = __ldg(src); dst[1] = __ldg(src + 1); dst[2] = __ldg(src + 2); dst[3] = __ldg(src + 3);
        visible[k + 1] = i;
        if (atlasByVpl != nullptr) entries[k] = atlasByVpl[i];
    }
    const int V = total > SGI_MAX_TOTAL_VPLS_PER_FRAME ? SGI_MAX_TOTAL_VPLS_PER_FRAME : total;
    if (threadIdx.x == 0) {
        visible[0] = V;
        *ticket = 0u;   // re-arm for the next frame
    }
    if (colorToo) {
        __syncthreads();                                         // the compacted table was written by this CTA
        const int texel = cubeTexel(-mk3(fc.sunDir[0], fc.sunDir[1], fc.sunDir[2]), cubeRes);
        for (int k = threadIdx.x; k < V; k += kCullThreads) colorOneVpl(fc, cubeRes, texel, updated, k, entries, atlases);
    }
}

__global__ void __launch_bounds__(256)
colorKernel(const __grid_constant__ FrameConsts fc, int cubeRes, SgiVplData* __restrict__ lightData, const int32_t* __restrict__ numVisible,
            const SgiAtlasEntry* __restrict__ diffuseIndices, const __grid_constant__ AtlasTable atlases) {
    const int visibleVpls = __ldg(numVisible);
    const int texel = cubeTexel(-mk3(fc.sunDir[0], fc.sunDir[1], fc.sunDir[2]), cubeRes);
    for (int index = blockIdx.x * blockDim.x + threadIdx.x; index < visibleVpls; index += gridDim.x * blockDim.x)
        colorOneVpl(fc, cubeRes, texel, lightData, index, diffuseIndices, atlases);
}

}  // namespace

// CUDA loads kernels lazily on first launch, and that load can stall behind a kernel that is spinning on a band flag
// (sgi_band.cu) — so every kernel of this file is loaded when the handle is created.
bool preloadVplKernels() {
    cudaFuncAttributes a; bool ok = true;
    ok = ok && cudaFuncGetAttributes(&a, cullKernel) == cudaSuccess;
    ok = ok && cudaFuncGetAttributes(&a, colorKernel) == cudaSuccess;
    return ok;
}

void fillFrameConsts(const SgiFrameParams* fp, FrameConsts* fc) {
    for (int i = 0; i < 16; ++i) fc->invPV[i] = fp->invProjectionView[i];
    for (int i = 0; i < 3; ++i) {
        fc->viewPos[i] = fp->viewPosition[i]; fc->fogColor[i] = fp->fogColor[i];
        fc->sunDir[i] = fp->infiniteLightDirection[i]; fc->sunColor[i] = fp->infiniteLightColor[i];
    }
    fc->time = fp->time;
    fc->framesPerSecond = (float)(1.0 / fp->deltaSeconds);
    fc->minGiOcclusionFactor = fp->minGiOcclusionFactor;
    fc->minRoughness = fp->minRoughness;
    fc->usePerceptualRoughness = fp->usePerceptualRoughness;
    fc->fogDensity = fp->fogDensity;
    for (int c = 0; c < 4; ++c) for (int i = 0; i < 16; ++i) fc->cascadeProjViews[c][i] = fp->cascadeProjViews[c][i];
    for (int c = 0; c < 3; ++c) for (int i = 0; i < 4; ++i) fc->cascadePlanes[c][i] = fp->cascadePlanes[c][i];
    fc->totalNumLights = fp->totalNumLights;
    // seed.z of the gather's draw sequence: time, then `seed.z += 10000.0` in binary32 per draw (vpl_pbr_gi.fs:150)
    volatile float seedZ = fp->time;
    for (int k = 0; k < kGatherMaxDraws; ++k) {
        const float z = seedZ;
        uint32_t bits; std::memcpy(&bits, &z, sizeof(bits));
        bits += (bits << 10u); bits ^= (bits >> 6u); bits += (bits << 3u); bits ^= (bits >> 11u); bits += (bits << 15u);     // hash(uint), common.glsl:183-190
        fc->hashZ[k] = bits;
        seedZ = z + 10000.0f;
    }
}

int launchCull(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, const SgiVplData* lightData, const SgiAtlasEntry* atlasByVpl,
               const float* csm, SgiVplData* updated, int32_t* visible, SgiAtlasEntry* entries, uint32_t* ballots, unsigned int* ticket,
               const AtlasTable* colorAtlases, const int32_t* countDevice) {
    const int N = fc.totalNumLights;
    const int blocks = N > 0 ? (N + 31) / 32 : 1;
    AtlasTable none;
    for (int i = 0; i < SGI_MAX_TOTAL_SHADOW_ATLASES; ++i) { none.shadow[i] = nullptr; none.diffuse[i] = nullptr; }
    cullKernel<<<blocks, kCullThreads, 0, stream>>>(fc, bd.csmRes, bd.cubeRes, lightData, atlasByVpl, csm, updated, visible, entries, ballots, ticket,
                                                    colorAtlases ? *colorAtlases : none, colorAtlases ? 1 : 0, countDevice);
    return 1;
}

int launchColor(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, SgiVplData* updated, const int32_t* visible,
                const SgiAtlasEntry* entries, const AtlasTable& atlases) {
    colorKernel<<<(SGI_MAX_TOTAL_VPLS_PER_FRAME + 255) / 256, 256, 0, stream>>>(fc, bd.cubeRes, updated, visible, entries, atlases);
    return 1;
}

}  // namespace sgi
