// host_frame.cpp — drives the GI path from C++ through stratus::VplGiBackend exactly as the reference's RenderScene does
// (BE:1395, 2043-2044): lights -> stage 1 -> stage 2 -> GI, on the procedural scene of stratusgfx_b200/scene (libsgi_scene.so).
//
//   host_frame <kind> <width> <height> <num_vpls> <frames> <seed> <denoise_iters> <prepass 0|1> <out_file>
//
// Writes, per frame, the visible count (int32), the visible list (int32 x count) and the lit image (half4 x W x H) to <out_file>,
// and prints one summary line.  Exit codes: 0 ok, 2 usage, 3 no CUDA device / backend error, 4 CUDA runtime error.
// tests/test_cpp_host.py compares the file with the Python host path byte for byte.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "stratus_gi_backend.h"

extern "C" {
typedef struct SgsDesc { int32_t kind, width, height, num_vpls, csm_res, cube_res; uint32_t seed; float yaw_deg_per_frame; } SgsDesc;
void* sgs_create(const SgsDesc* d);
void sgs_destroy(void* p);
int sgs_frame_params(void* p, int frame, SgiFrameParams* out);
int sgs_gbuffer(void* p, int frame, int y0, int rows, float* depth, uint8_t* normal4, uint8_t* albedo4, uint8_t* refl2, uint8_t* rma4,
                uint16_t* ssaoH, uint16_t* velocityH2, float* id, uint16_t* screenH4);
int sgs_vpls(void* p, float* data16, int32_t* atlas2);
int sgs_cubemaps(void* p, int first, int count, uint16_t* depth, uint8_t* diffuse4);
int sgs_csm(void* p, int frame, float* out);
}

#define CU(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #expr, cudaGetErrorString(e_)); return 4; } } while (0)

template <typename T> static T* devAlloc(size_t n) { void* p = nullptr; return cudaMalloc(&p, (n ? n : 1) * sizeof(T)) == cudaSuccess ? (T*)p : nullptr; }

int main(int argc, char** argv) {
    if (argc != 10) { std::fprintf(stderr, "usage: host_frame kind width height num_vpls frames seed denoise_iters prepass out_file\n"); return 2; }
    const int kind = std::atoi(argv[1]), W = std::atoi(argv[2]), H = std::atoi(argv[3]), N = std::atoi(argv[4]), frames = std::atoi(argv[5]);
    const unsigned seed = (unsigned)std::strtoul(argv[6], nullptr, 0);
    const int iters = std::atoi(argv[7]), prepass = std::atoi(argv[8]);
    const int csmRes = 1024, cubeRes = 32;
    try {
        stratus::VplGiBackend backend(W, H, cubeRes, csmRes, iters, 0);          // throws without a CUDA device
        SgsDesc d{kind, W, H, N, csmRes, cubeRes, seed, 0.0f};
        void* scene = sgs_create(&d);
        if (!scene) { std::fprintf(stderr, "scene description rejected\n"); return 2; }
        const size_t P = (size_t)W * H, tiles = (size_t)(W / 5) * (H / 5);
        const int atlases = (N + SGI_MAX_TOTAL_SHADOWS_PER_ATLAS - 1) / SGI_MAX_TOTAL_SHADOWS_PER_ATLAS > 0 ? (N + SGI_MAX_TOTAL_SHADOWS_PER_ATLAS - 1) / SGI_MAX_TOTAL_SHADOWS_PER_ATLAS : 1;
        const size_t layerTexels = (size_t)6 * cubeRes * cubeRes, atlasTexels = (size_t)SGI_MAX_TOTAL_SHADOWS_PER_ATLAS * layerTexels;
        // host copies of the static inputs
        std::vector<float> vpl((size_t)N * 16); std::vector<int32_t> atlas((size_t)N * 2);
        sgs_vpls(scene, vpl.data(), atlas.data());
        std::vector<uint16_t> cubeDepth(atlases * atlasTexels, 65535); std::vector<uint8_t> cubeDiffuse(atlases * atlasTexels * 4, 0);
        sgs_cubemaps(scene, 0, N, cubeDepth.data(), cubeDiffuse.data());
        // device buffers (the caller owns them, as the reference owns its GL objects)
        SgiBuffers b{};
        float *dDepth = devAlloc<float>(P), *dId = devAlloc<float>(P), *dPrevDepth = devAlloc<float>(P), *dPrevId = devAlloc<float>(P), *dCsm = devAlloc<float>((size_t)4 * csmRes * csmRes);
        uint8_t *dNormal = devAlloc<uint8_t>(P * 4), *dAlbedo = devAlloc<uint8_t>(P * 4), *dRefl = devAlloc<uint8_t>(P * 2), *dRma = devAlloc<uint8_t>(P * 4), *dPrevNormal = devAlloc<uint8_t>(P * 4);
        uint16_t *dSsao = devAlloc<uint16_t>(P), *dVel = devAlloc<uint16_t>(P * 2), *dScreen = devAlloc<uint16_t>(P * 4), *dOut = devAlloc<uint16_t>(P * 4);
        SgiVplData *dVpl = devAlloc<SgiVplData>(N), *dUpdated = devAlloc<SgiVplData>(SGI_MAX_TOTAL_VPLS_PER_FRAME);
        int32_t *dVisible = devAlloc<int32_t>((size_t)N + 1), *dCount = devAlloc<int32_t>(1), *dOrder = devAlloc<int32_t>(N);
        SgiAtlasEntry *dEntries = devAlloc<SgiAtlasEntry>(SGI_MAX_TOTAL_VPLS_PER_FRAME), *dAtlasByVpl = devAlloc<SgiAtlasEntry>(N);
        SgiVplStage1PerTileOutputs* dS1 = devAlloc<SgiVplStage1PerTileOutputs>(tiles);
        SgiVplStage2PerTileOutputs* dS2 = devAlloc<SgiVplStage2PerTileOutputs>(tiles);
        uint16_t* dCubeDepth = devAlloc<uint16_t>(cubeDepth.size()); uint8_t* dCubeDiffuse = devAlloc<uint8_t>(cubeDiffuse.size());
        SgiLightRecord* dLights = devAlloc<SgiLightRecord>(N);
        if (!dDepth || !dOut || !dCubeDiffuse || !dLights) { std::fprintf(stderr, "cudaMalloc failed\n"); return 4; }
        CU(cudaMemset(dVisible, 0, ((size_t)N + 1) * sizeof(int32_t)));
        CU(cudaMemset(dUpdated, 0, SGI_MAX_TOTAL_VPLS_PER_FRAME * sizeof(SgiVplData)));
        CU(cudaMemset(dEntries, 0, SGI_MAX_TOTAL_VPLS_PER_FRAME * sizeof(SgiAtlasEntry)));
        CU(cudaMemset(dOut, 0, P * 8));
        CU(cudaMemcpy(dCubeDepth, cubeDepth.data(), cubeDepth.size() * 2, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(dCubeDiffuse, cubeDiffuse.data(), cubeDiffuse.size(), cudaMemcpyHostToDevice));
        if (prepass) {
            // the engine's per-light records; the device pre-pass builds vplData / atlasEntryByVpl / N from them
            std::vector<SgiLightRecord> rec(N);
            for (int i = 0; i < N; ++i) {
                std::memcpy(rec[i].position, &vpl[(size_t)i * 16], 12);
                rec[i].radius = vpl[(size_t)i * 16 + 12]; rec[i].farPlane = vpl[(size_t)i * 16 + 13]; rec[i].intensity = vpl[(size_t)i * 16 + 14];
                rec[i].atlasIndex = atlas[(size_t)i * 2]; rec[i].atlasLayer = atlas[(size_t)i * 2 + 1];
            }
            CU(cudaMemcpy(dLights, rec.data(), (size_t)N * sizeof(SgiLightRecord), cudaMemcpyHostToDevice));
            b.vplCountDevice = dCount;
        } else {
            CU(cudaMemcpy(dVpl, vpl.data(), (size_t)N * 64, cudaMemcpyHostToDevice));
            CU(cudaMemcpy(dAtlasByVpl, atlas.data(), (size_t)N * 8, cudaMemcpyHostToDevice));
        }
        b.depth = dDepth; b.normal = dNormal; b.albedo = dAlbedo; b.baseReflectivity = dRefl; b.roughnessMetallicAmbient = dRma;
        b.velocity = dVel; b.id = dId; b.prevNormal = dPrevNormal; b.prevId = dPrevId; b.prevDepth = dPrevDepth; b.ssao = dSsao; b.screen = dScreen;
        b.lightingOut = dOut; b.vplData = dVpl; b.vplUpdatedData = dUpdated; b.vplVisibleIndices = dVisible; b.shadowDiffuseIndices = dEntries;
        b.vplStage1Results = dS1; b.vplVisiblePerTile = dS2; b.atlasEntryByVpl = dAtlasByVpl; b.infiniteLightShadowMap = dCsm;
        for (int a = 0; a < atlases && a < SGI_MAX_TOTAL_SHADOW_ATLASES; ++a) { b.shadowCubeMaps[a] = dCubeDepth + a * atlasTexels; b.diffuseCubeMaps[a] = dCubeDiffuse + a * atlasTexels * 4; }
        backend.BindBuffers(b);

        FILE* out = std::fopen(argv[9], "wb");
        if (!out) { std::fprintf(stderr, "cannot open %s\n", argv[9]); return 2; }
        std::vector<float> hDepth(P), hId(P), hCsm((size_t)4 * csmRes * csmRes);
        std::vector<uint8_t> hNormal(P * 4), hAlbedo(P * 4), hRefl(P * 2), hRma(P * 4);
        std::vector<uint16_t> hSsao(P), hVel(P * 2), hScreen(P * 4), hOut(P * 4);
        std::vector<int32_t> hVisible((size_t)N + 1);
        int lastV = 0;
        for (int f = 0; f < frames; ++f) {
            SgiFrameParams fp; sgs_frame_params(scene, f, &fp);
            // previous frame's planes, then this frame's (a renderer keeps the former from the last frame; the generator recomputes them)
            sgs_gbuffer(scene, f - 1, 0, H, hDepth.data(), hNormal.data(), nullptr, nullptr, nullptr, nullptr, nullptr, hId.data(), nullptr);
            CU(cudaMemcpy(dPrevDepth, hDepth.data(), P * 4, cudaMemcpyHostToDevice)); CU(cudaMemcpy(dPrevNormal, hNormal.data(), P * 4, cudaMemcpyHostToDevice));
            CU(cudaMemcpy(dPrevId, hId.data(), P * 4, cudaMemcpyHostToDevice));
            sgs_gbuffer(scene, f, 0, H, hDepth.data(), hNormal.data(), hAlbedo.data(), hRefl.data(), hRma.data(), hSsao.data(), hVel.data(), hId.data(), hScreen.data());
            sgs_csm(scene, 0, hCsm.data());                                      // static camera: the sun's cascades of frame 0 serve every frame (as tests/parity_util.py does)
            CU(cudaMemcpy(dDepth, hDepth.data(), P * 4, cudaMemcpyHostToDevice)); CU(cudaMemcpy(dNormal, hNormal.data(), P * 4, cudaMemcpyHostToDevice));
            CU(cudaMemcpy(dAlbedo, hAlbedo.data(), P * 4, cudaMemcpyHostToDevice)); CU(cudaMemcpy(dRefl, hRefl.data(), P * 2, cudaMemcpyHostToDevice));
            CU(cudaMemcpy(dRma, hRma.data(), P * 4, cudaMemcpyHostToDevice)); CU(cudaMemcpy(dSsao, hSsao.data(), P * 2, cudaMemcpyHostToDevice));
            CU(cudaMemcpy(dVel, hVel.data(), P * 4, cudaMemcpyHostToDevice)); CU(cudaMemcpy(dId, hId.data(), P * 4, cudaMemcpyHostToDevice));
            CU(cudaMemcpy(dScreen, hScreen.data(), P * 8, cudaMemcpyHostToDevice)); CU(cudaMemcpy(dCsm, hCsm.data(), hCsm.size() * 4, cudaMemcpyHostToDevice));
            backend.BindBuffers(b);                                             // a new frame's textures are bound (resets per-frame state as Python's bind() does)
            if (prepass) {
                // a frustum that contains everything (six planes 0*x + 0*y + 0*z + 1 >= -radius): the pre-pass then only orders by distance
                SgiPrepassParams pp{};
                std::memcpy(pp.viewPosition, fp.viewPosition, 12);
                pp.numLights = N;
                for (int k = 0; k < 6; ++k) { pp.frustumPlanes[k][0] = pp.frustumPlanes[k][1] = pp.frustumPlanes[k][2] = 0.0f; pp.frustumPlanes[k][3] = 1.0f; }
                pp.lights = dLights; pp.vplDataOut = dVpl; pp.atlasEntryOut = dAtlasByVpl; pp.orderOut = dOrder; pp.countOut = dCount;
                backend.UpdatePointLights_(pp);
            }
            // RenderScene: BE:1395 (stage 1 inside UpdatePointLights_), BE:2043-2044
            backend.PerformVirtualPointLightCullingStage1_(fp);
            backend.PerformVirtualPointLightCullingStage2_(fp);
            backend.ComputeVirtualPointLightGlobalIllumination_(fp, fp.deltaSeconds);
            lastV = backend.VisibleVplCount();
            CU(cudaMemcpy(hVisible.data(), dVisible, ((size_t)N + 1) * 4, cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(hOut.data(), dOut, P * 8, cudaMemcpyDeviceToHost));
            const int32_t v32 = lastV;
            std::fwrite(&v32, 4, 1, out);
            std::fwrite(hVisible.data() + 1, 4, (size_t)lastV, out);
            std::fwrite(hOut.data(), 2, P * 4, out);
        }
        std::fclose(out);
        std::printf("host_frame ok: %dx%d N=%d frames=%d V=%d launches=%lld\n", W, H, N, frames, lastV, backend.LaunchCount());
        sgs_destroy(scene);
        return 0;
    } catch (const std::runtime_error& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 3;
    }
}
