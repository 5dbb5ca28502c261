// ref_shim.cpp — the parts of the REFERENCE that compile here, exposed for the oracle's tests.  TEST INFRASTRUCTURE.
//
// StratusGFX as a whole needs OpenGL, SDL2, assimp and a generated gl3w, so it cannot be built in this image.  Two of its
// headers on the GI path do compile from where they lie under /root/reference once the four third-party includes they drag in
// are satisfied by empty stubs (oracle/ref_stubs/): Source/Engine/StratusMath.h (IsSphereInFrustum, :401-412) and
// Source/Engine/StratusGpuCommon.h (the std430 wire structs and limits, :15-36, :192-232, :276-279).  This file includes
// them unmodified and exports
//   - the reference's own IsSphereInFrustum,
//   - sizeof / offsetof of its wire structs and its limits,
//   - a GpuVplData filled exactly as InitVplFrameData_ fills it (StratusRendererBackend.cpp:1281-1292),
// so tests/test_reference_shim.py can hold the oracle and the ABI against reference code instead of a restatement.
// Built by oracle/oracle_py.py::build into oracle/_ref/libref_stratus.so when /root/reference exists.  Nothing is copied.
#include "StratusMath.h"
#include <cstddef>
#include <cstring>
#include <vector>

extern "C" {

int ref_is_sphere_in_frustum(const float* center, float radius, const float* planes /* [6][4] */) {
    std::vector<glm::vec4> p(6);
    for (int i = 0; i < 6; ++i) p[i] = glm::vec4(planes[4 * i], planes[4 * i + 1], planes[4 * i + 2], planes[4 * i + 3]);
    return stratus::IsSphereInFrustum(glm::vec3(center[0], center[1], center[2]), radius, p) ? 1 : 0;
}

// out[0..] = sizeof(GpuVplData), offsets of position, color, specularPosition, radius, farPlane, intensity;
//            sizeof(GpuVplStage1PerTileOutputs), offset of averageLocalNormal; sizeof(GpuVplStage2PerTileOutputs), offset of indices;
//            sizeof(GpuAtlasEntry), offset of layer; the five limits
int ref_layout(int* out) {
    using namespace stratus;
    int k = 0;
    out[k++] = (int)sizeof(GpuVplData);
    out[k++] = (int)offsetof(GpuVplData, position); out[k++] = (int)offsetof(GpuVplData, color); out[k++] = (int)offsetof(GpuVplData, specularPosition);
    out[k++] = (int)offsetof(GpuVplData, radius); out[k++] = (int)offsetof(GpuVplData, farPlane); out[k++] = (int)offsetof(GpuVplData, intensity);
    out[k++] = (int)sizeof(GpuVplStage1PerTileOutputs); out[k++] = (int)offsetof(GpuVplStage1PerTileOutputs, averageLocalNormal);
    out[k++] = (int)sizeof(GpuVplStage2PerTileOutputs); out[k++] = (int)offsetof(GpuVplStage2PerTileOutputs, indices);
    out[k++] = (int)sizeof(GpuAtlasEntry); out[k++] = (int)offsetof(GpuAtlasEntry, layer);
    out[k++] = MAX_TOTAL_SHADOW_ATLASES; out[k++] = MAX_TOTAL_SHADOWS_PER_ATLAS; out[k++] = MAX_TOTAL_VPLS_BEFORE_CULLING;
    out[k++] = MAX_TOTAL_VPLS_PER_FRAME; out[k++] = MAX_VPLS_PER_TILE;
    return k;
}

// the loop body of InitVplFrameData_ (StratusRendererBackend.cpp:1284-1290) on the reference's own types; the one field neither
// the constructor nor the loop writes (placeholder2_) is zeroed so that the bytes are comparable
void ref_init_vpl_frame_data(const float* position, float farPlane, float radius, float intensity, void* out64) {
    stratus::GpuVplData data;
    data.placeholder2_ = 0.0f;
    data.position = stratus::GpuVec(glm::vec4(glm::vec3(position[0], position[1], position[2]), 1.0f));
    data.farPlane = farPlane;
    data.radius = radius;
    data.intensity = intensity;
    std::memcpy(out64, &data, sizeof(data));
}

}  // extern "C"
