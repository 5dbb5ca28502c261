// sgi_internal.h — handle layout and stage launchers shared by the .cu files of libstratusgi.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include "../../include/stratusgi.h"

namespace sgi {

// per-launch geometry of a row band (see include/stratusgi.h "Row bands")
struct BandDims {
    int W;            // frame width
    int fullH;        // frame height
    int yOffset;      // first frame row held by the planes
    int rows;         // rows held by the planes
    int cubeRes;
    int csmRes;
};

// uniforms in device-friendly form; passed by value as a __grid_constant__ kernel parameter
struct FrameConsts {
    float invPV[16];
    float viewPos[3];
    float time;
    float framesPerSecond;
    float minGiOcclusionFactor;
    float minRoughness;
    int   usePerceptualRoughness;
    float fogDensity;
    float fogColor[3];
    float sunDir[3];
    float sunColor[3];
    float cascadeProjViews[4][16];
    float cascadePlanes[3][4];
    int   totalNumLights;
};

struct AtlasTable {
    const unsigned short* shadow[SGI_MAX_TOTAL_SHADOW_ATLASES];
    const uchar4* diffuse[SGI_MAX_TOTAL_SHADOW_ATLASES];
};

struct DenoiseIO {
    const uint2* screen;          // half4
    const uint32_t* velocity;     // half2
    const uint32_t* normal;       // rgba8
    const float* ids;
    const float* depth;
    const uint32_t* prevNormal;
    const float* prevIds;
    const float* prevDepth;
    const float4* guide;          // library-owned: (decoded normal, depth) of every pixel of this frame
    const uint2* illum;           // half4: the demodulation plane (vplGIFbo attachment 0)
    const uint2* shadowsIn;       // half4: reservoir input of this pass
    const uint2* prevGi;          // half4
    const unsigned short* historyIn;   // half
    const int32_t* numVisible;    // vplVisibleIndices (count in slot 0)
    uint2* reservoirOut;          // half4
    uint2* combinedOut;           // half4 (final pass)
    uint2* giOut;                 // half4 (final pass)
    unsigned short* historyOut;   // half  (final pass)
    int multiplier;
};

}  // namespace sgi

struct SgiHandle_ {
    SgiCreateParams cp;
    SgiBuffers buf;
    bool bound = false;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    int64_t launches = 0;
    std::string lastError;
    // library-owned planes
    void* giColor = nullptr;        // half4
    void* giReservoir = nullptr;    // half4
    void* resA = nullptr;           // half4 (vplGIDenoisedFbo1 attachment 2)
    void* resB = nullptr;           // half4 (vplGIDenoisedFbo2 attachment 2)
    void* combined = nullptr;       // half4 (used when lightingOut is not bound)
    void* giIllum[2] = {nullptr, nullptr};    // half4, [cur] / [cur^1] = previous frame
    void* history[2] = {nullptr, nullptr};    // half
    int cur = 0;
    void* guide = nullptr;          // float4 (normalize(normal*2-1), depth), rebuilt once per frame
    bool guideValid = false;        // set by the gather (or the first denoise pass), cleared by bind and the final pass
    bool pass0Fused = false;        // the gather of this frame already wrote denoise pass 0 into resA, for `pass0Time`
    float pass0Time = 0.0f;
    // scratch for the cull
    uint32_t* cullBallots = nullptr;
    unsigned int* cullTicket = nullptr;
    int32_t* pinnedCount = nullptr;
};

namespace sgi {

void fillFrameConsts(const SgiFrameParams* fp, FrameConsts* fc);

// each launcher enqueues its kernels on `stream` and returns the number of launches made
int launchCull(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, const SgiVplData* lightData, const SgiAtlasEntry* atlasByVpl,
               const float* csm, SgiVplData* updated, int32_t* visible, SgiAtlasEntry* entries, uint32_t* ballots, unsigned int* ticket);
int launchColor(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, SgiVplData* updated, const int32_t* visible,
                const SgiAtlasEntry* entries, const AtlasTable& atlases);
int launchStage1(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, const float* depth, const uint32_t* normal,
                 SgiVplStage1PerTileOutputs* out);
int launchStage2(cudaStream_t stream, const BandDims& bd, const SgiVplStage1PerTileOutputs* stage1, const SgiVplData* lightData,
                 const int32_t* visible, SgiVplStage2PerTileOutputs* out);
int launchGather(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, const float* depth, const uint32_t* normal,
                 const uint32_t* albedo, const uint16_t* refl, const uint32_t* rma, const unsigned short* ssao,
                 const SgiVplData* updated, const int32_t* visible, const SgiAtlasEntry* entries, const AtlasTable& atlases,
                 uint2* colorOut, uint2* reservoirOut, float4* guideOut, uint2* mergedOut);
int launchPrepareGuide(cudaStream_t stream, const BandDims& bd, const uint32_t* normal, const float* depth, float4* guide);
int launchDenoisePass(cudaStream_t stream, const FrameConsts& fc, const BandDims& bd, const DenoiseIO& io, bool merge, bool final_);

}  // namespace sgi
