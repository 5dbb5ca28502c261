// stratus_gi_backend.h — C++ host side above the C-ABI (include/stratusgi.h), in the reference's own language.
//
// The reference is a C++ engine; its GI hot path is three private methods of stratus::RendererBackend plus the light loop that
// feeds them (Source/Engine/StratusRendererBackend.h:657-669).  This header mirrors that interface — same method names,
// argument meaning and error behaviour (the reference throws std::runtime_error on fatal errors, StratusRendererBackend.cpp:70-72,
// 429-432) — over libstratusgi.so, so a maintainer can swap the bodies of the reference methods for these calls.
// It owns nothing but the handle: device buffers belong to the caller, as the GL objects do in the reference.
// There is no CPU fallback: constructing a backend on a box without a CUDA device throws.
#pragma once
#include <stdexcept>
#include <string>
#include "../../include/stratusgi.h"

namespace stratus {

class VplGiBackend {
public:
    // UpdateWindowDimensions_ (BE:436-445, 510-589) + InitializeVplData_ (BE:297-310)
    VplGiBackend(int width, int height, int cubeRes = 32, int csmRes = 1024, int denoiseIterations = 4, int device = 0, void* stream = nullptr) {
        SgiCreateParams cp{};
        cp.width = width; cp.full_height = height; cp.y_offset = 0; cp.rows = height;
        cp.cube_res = cubeRes; cp.csm_res = csmRes; cp.denoise_iters = denoiseIterations; cp.device = device; cp.stream = stream;
        const int rc = sgi_create(&cp, &handle_);
        if (rc != SGI_OK) throw std::runtime_error(rc == SGI_ERR_CUDA ? "VplGiBackend: no usable CUDA device (there is no CPU path)" : "VplGiBackend: sgi_create rejected the parameters");
    }
    ~VplGiBackend() { if (handle_) sgi_destroy(handle_); }
    VplGiBackend(const VplGiBackend&) = delete;
    VplGiBackend& operator=(const VplGiBackend&) = delete;

    // the BindBase / BindTexture calls of BE:1599-1601, 1700-1722, 1840-1861, 1888-1900
    void BindBuffers(const SgiBuffers& buffers) { Check_(sgi_bind_device_buffers(handle_, &buffers)); }

    // UpdatePointLights_ (BE:1359-1392), virtual-point-light branch, + InitVplFrameData_ (BE:1281-1292): on the device
    void UpdatePointLights_(const SgiPrepassParams& lights) { Check_(sgi_vpl_prepass(handle_, &lights)); }
    // the second half of UpdatePointLights_ (BE:1426-1575): the cube shadow / diffuse maps of the lights in frame_->lightsToUpdate,
    // at most maxShadowUpdatesPerFrame of them per frame (StratusRendererBackend.h:378), rendered into their atlas layers
    void UpdatePointLightShadowMaps_(const SgiCubeRenderParams& lightsToUpdate) { Check_(sgi_render_vpl_cubemaps(handle_, &lightsToUpdate)); }

    // RenderSsaoOcclude_ + RenderSsaoBlur_ (BE:1160-1204): ssao.fs and ssao_blur.fs, producing the `ssao` plane of BindBuffers.
    // (csrc/k_ssao.cu is prepared but has not run on a GPU yet, see include/stratusgi.h)
    void RenderSsao_(const SgiSsaoParams& ssao) { Check_(sgi_ssao(handle_, &ssao)); }

    // PerformTaaPostFx_ (BE:2324-2354): taa.fs and the blit into the history (csrc/k_taa.cu)
    void PerformTaaPostFx_(const SgiTaaParams& taa) { Check_(sgi_taa(handle_, &taa)); }

    // PerformVirtualPointLightCullingStage1_ (BE:1578-1612): viscull_vpls.cs, without the fence
    void PerformVirtualPointLightCullingStage1_(const SgiFrameParams& frame) {
        if (frame.totalNumLights == 0) return;                                       // BE:1582
        Check_(sgi_cull(handle_, &frame));
    }
    // PerformVirtualPointLightCullingStage2_ (BE:1666-1731, dormant tile stages 1733-1783)
    void PerformVirtualPointLightCullingStage2_(const SgiFrameParams& frame, bool tileLists = true) {
        if (frame.totalNumLights == 0) return;                                       // BE:1675
        Check_(sgi_color(handle_, &frame));
        if (tileLists) Check_(sgi_tile_lists(handle_, &frame));
    }
    // ComputeVirtualPointLightGlobalIllumination_ (BE:1816-1952): vpl_pbr_gi.fs, the denoise passes, blit, history swap
    void ComputeVirtualPointLightGlobalIllumination_(const SgiFrameParams& frame, const double deltaSeconds) {
        if (frame.totalNumLights == 0) return;                                       // BE:1817
        SgiFrameParams fp = frame;
        fp.deltaSeconds = deltaSeconds;                                              // framesPerSecond = float(1.0 / deltaSeconds), BE:1908
        Check_(sgi_gather(handle_, &fp));
        Check_(sgi_denoise(handle_, &fp));
    }
    // all three on one stream (tile lists next to the gather)
    void RenderGi(const SgiFrameParams& frame) { Check_(sgi_frame(handle_, &frame)); }

    int VisibleVplCount() { int32_t n = 0; Check_(sgi_visible_count(handle_, &n)); return n; }     // vplVisibleIndices.MapMemory()[0], BE:1675-1678
    void Synchronize() { Check_(sgi_synchronize(handle_)); }
    long long LaunchCount() const { return sgi_launch_count(handle_); }
    SgiHandle Handle() const { return handle_; }

private:
    void Check_(int rc) {
        if (rc != SGI_OK) throw std::runtime_error(std::string("VplGiBackend: status ") + std::to_string(rc) + ": " + sgi_last_error(handle_));
    }
    SgiHandle handle_ = nullptr;
};

}  // namespace stratus
