// k_prepass.cu — the VPL pre-pass on the device.
//
// Replaces the CPU loop of RendererBackend::UpdatePointLights_ (StratusRendererBackend.cpp:1359-1392) for virtual point
// lights and InitVplFrameData_ (:1281-1292):
//   for every light:  distance = glm::distance(cameraPosition, lightPosition)                       (:1360)
//                     keep it if IsSphereInFrustum(position, radius, viewFrustumPlanes)             (:1363, StratusMath.h:401-412)
//                     perVPLDistToViewerSet.insert(VplDistKey_(light, distance))                    (:1365) — a std::multiset
//                       ordered by distance alone (StratusRendererBackend.h:609-611), so equal distances keep insertion order
//   the first MAX_TOTAL_VPLS_BEFORE_CULLING = 10000 entries become vplData[i] = {position, farPlane, radius, intensity}  (:1385-1390, :1281-1292)
//
// Device form.  Distances are non-negative floats, whose bit patterns order like the values, so (distance bits << 32 | input
// index) is a unique 64-bit key whose ascending order is exactly the multiset's iteration order; lights outside the frustum
// get the all-ones key.  The rank of a key (how many keys are smaller) is its position in the output: every CTA ranks 256
// keys against all keys, staged through shared memory in tiles — no sort network, no inter-CTA dependency, and one pass over
// n keys per 256 outputs is cheap at the reference's sizes (n <= ~10^4: 40 CTAs, < 10 us).  Every float operation that feeds
// the order or the frustum decision is the exact IEEE chain in glm's evaluation order (csrc/sgi_math.cuh), so the output
// table is bit-identical to the CPU loop's.
#include "sgi_internal.h"
#include "sgi_math.cuh"

namespace sgi {
namespace {

constexpr int kPpThreads = 256;
constexpr int kPpTile = 2048;                 // keys staged per tile (16 KB)
constexpr unsigned long long kOutside = ~0ull;

__global__ void __launch_bounds__(kPpThreads)
prepassKeysKernel(int n, const __grid_constant__ SgiPrepassParams pp, unsigned long long* __restrict__ keys, int* __restrict__ inside) {
    const int i = blockIdx.x * kPpThreads + threadIdx.x;
    bool in = false;
    if (i < n) {
        const SgiLightRecord L = pp.lights[i];
        const f3 p = mk3(L.position[0], L.position[1], L.position[2]);
        const f3 viewPos = mk3(pp.viewPosition[0], pp.viewPosition[1], pp.viewPosition[2]);
        const float distance = length(p - viewPos);            // glm::distance(p0, p1) = length(p1 - p0)
        in = true;
        #pragma unroll
        for (int k = 0; k < 6; ++k) {
            const f4 g = mk4(pp.frustumPlanes[k][0], pp.frustumPlanes[k][1], pp.frustumPlanes[k][2], pp.frustumPlanes[k][3]);
            if (dot(g, mk4(p.x, p.y, p.z, 1.0f)) < -L.radius) in = false;
        }
        keys[i] = in ? (((unsigned long long)__float_as_uint(distance) << 32) | (unsigned int)i) : kOutside;
    }
    const unsigned int bal = __ballot_sync(0xffffffffu, in);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(inside, __popc(bal));
}

__global__ void __launch_bounds__(kPpThreads)
prepassRankKernel(int n, const __grid_constant__ SgiPrepassParams pp, const unsigned long long* __restrict__ keys, const int* __restrict__ inside) {
    __shared__ unsigned long long sKeys[kPpTile];
    const int i = blockIdx.x * kPpThreads + threadIdx.x;
    const unsigned long long mine = i < n ? keys[i] : kOutside;
    int rank = 0;
    for (int base = 0; base < n; base += kPpTile) {
        const int m = min(kPpTile, n - base);
        __syncthreads();
        for (int t = threadIdx.x; t < m; t += kPpThreads) sKeys[t] = keys[base + t];
        __syncthreads();
        #pragma unroll 8
        for (int t = 0; t < m; ++t) rank += sKeys[t] < mine ? 1 : 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *pp.countOut = min(*inside, SGI_MAX_TOTAL_VPLS_BEFORE_CULLING);
    if (mine == kOutside || rank >= SGI_MAX_TOTAL_VPLS_BEFORE_CULLING) return;
    const SgiLightRecord L = pp.lights[i];
    float4* o = reinterpret_cast<float4*>(pp.vplDataOut + rank);
    o[0] = make_float4(L.position[0], L.position[1], L.position[2], 1.0f);       // GpuVec(glm::vec4(position, 1.0f))
    o[1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);                                   // color and specularPosition are filled by vpl_light_color.cs
    o[2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    o[3] = make_float4(L.radius, L.farPlane, L.intensity, 0.0f);
    if (pp.atlasEntryOut != nullptr) { SgiAtlasEntry e; e.index = L.atlasIndex; e.layer = L.atlasLayer; pp.atlasEntryOut[rank] = e; }
    if (pp.orderOut != nullptr) pp.orderOut[rank] = i;
}

}  // namespace

bool preloadPrepassKernels() {
    cudaFuncAttributes a; bool ok = true;
    ok = ok && cudaFuncGetAttributes(&a, prepassKeysKernel) == cudaSuccess;
    ok = ok && cudaFuncGetAttributes(&a, prepassRankKernel) == cudaSuccess;
    return ok;
}

int launchPrepass(cudaStream_t stream, const SgiPrepassParams& pp, unsigned long long* keys, int* inside) {
    const int n = pp.numLights;
    cudaMemsetAsync(inside, 0, sizeof(int), stream);
    if (n <= 0) {
        cudaMemsetAsync(pp.countOut, 0, sizeof(int32_t), stream);
        return 0;
    }
    const int blocks = (n + kPpThreads - 1) / kPpThreads;
    prepassKeysKernel<<<blocks, kPpThreads, 0, stream>>>(n, pp, keys, inside);
    prepassRankKernel<<<blocks, kPpThreads, 0, stream>>>(n, pp, keys, inside);
    return 2;
}

}  // namespace sgi
