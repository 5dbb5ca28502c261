/*
 * stratusgi.h — C-ABI of the B200-native VPL global-illumination pass.
 *
 * This is the drop-in boundary for the three private RendererBackend methods of
 * StratusGFX that make up its real-time GI hot path (reference paths relative to
 * the StratusGFX tree):
 *
 *   PerformVirtualPointLightCullingStage1_      Source/Engine/StratusRendererBackend.cpp:1578-1612
 *   PerformVirtualPointLightCullingStage2_      Source/Engine/StratusRendererBackend.cpp:1666-1731 (+ dormant 1733-1783)
 *   ComputeVirtualPointLightGlobalIllumination_ Source/Engine/StratusRendererBackend.cpp:1816-1952
 *
 * Everything is `extern "C"`, plain pointers and sizes, status returned as int
 * (0 = OK).  No C++ exception crosses the boundary and there is NO CPU fallback:
 * if no CUDA device is usable every entry point returns SGI_ERR_CUDA.
 *
 * Threading: one thread per handle, stream-ordered, non-reentrant — the same
 * contract as the reference, which calls these passes from the single thread
 * that owns the GL context (StratusRendererBackend.cpp:810, :1955).
 *
 * Memory formats ("planes") are the reference's G-buffer formats
 * (StratusRendererBackend.cpp:384-424, 511-585, 2413-2421) with 3-component
 * texels padded to 4 components — the same padding CUDA-GL interop forces
 * (cudaGraphicsGLRegisterImage accepts only 1/2/4-component formats):
 *
 *   depth, prevDepth            float32          [rows][W]   GL window depth in [0,1]
 *   normal, prevNormal          unorm8 x4 (rgb_) [rows][W]   n*0.5+0.5
 *   albedo                      unorm8 x4 (rgb_) [rows][W]
 *   baseReflectivity            unorm8 x2 (r_)   [rows][W]
 *   roughnessMetallicAmbient    unorm8 x4 (rgb_) [rows][W]
 *   ssao                        binary16         [rows][W]
 *   velocity                    binary16 x2      [rows][W]
 *   id, prevId                  float32          [rows][W]
 *   screen (lightingColorBuffer) binary16 x4 (rgb_) [rows][W]
 *   GI targets                  binary16 x4      [rows][W]   (RGB16F padded, RGBA16F as is)
 *   history                     binary16         [rows][W]
 *   shadowCubeMaps[a]           unorm16          [layers][6][res][res]   face order +X,-X,+Y,-Y,+Z,-Z
 *   diffuseCubeMaps[a]          unorm8 x4 (rgb_) [layers][6][res][res]
 *   infiniteLightShadowMap      float32          [cascades][csm][csm]    depth in [0,1]
 *
 * Row bands: a handle may own only rows [y_offset, y_offset+rows) of a
 * full_height-row frame (multi-GPU split).  Every per-pixel
 * plane handed to that handle then holds exactly `rows` rows.
 */
#ifndef STRATUSGI_H
#define STRATUSGI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- limits: Source/Engine/StratusGpuCommon.h:15-17, :34-36 ------------------------------ */
#define SGI_MAX_TOTAL_SHADOW_ATLASES        14
#define SGI_MAX_TOTAL_SHADOWS_PER_ATLAS     320
#define SGI_MAX_TOTAL_SHADOW_MAPS           (SGI_MAX_TOTAL_SHADOW_ATLASES * SGI_MAX_TOTAL_SHADOWS_PER_ATLAS)
#define SGI_MAX_TOTAL_VPLS_BEFORE_CULLING   10000
#define SGI_MAX_TOTAL_VPLS_PER_FRAME        SGI_MAX_TOTAL_SHADOW_MAPS   /* 4480 */
#define SGI_MAX_VPLS_PER_TILE               12
#define SGI_TILE_DIVISOR                    5   /* tileXDivisor/tileYDivisor, StratusRendererBackend.h:346-347 */
#define SGI_NUM_CASCADES                    4

/* ---- wire formats (std430, tightly packed) ------------------------------------------------ */
/* GpuVplData, Source/Engine/StratusGpuCommon.h:217-232; GLSL twin vpl_common.glsl:48-56 */
typedef struct SgiVplData {
    float position[4];
    float color[4];
    float specularPosition[4];
    float radius;
    float farPlane;
    float intensity;
    float _pad;
} SgiVplData;

/* GpuVplStage1PerTileOutputs, StratusGpuCommon.h:192-195 */
typedef struct SgiVplStage1PerTileOutputs {
    float averageLocalPosition[4];
    float averageLocalNormal[4];
} SgiVplStage1PerTileOutputs;

/* GpuVplStage2PerTileOutputs, StratusGpuCommon.h:203-209 */
typedef struct SgiVplStage2PerTileOutputs {
    int32_t numVisible;
    int32_t indices[SGI_MAX_VPLS_PER_TILE];
} SgiVplStage2PerTileOutputs;

/* GpuAtlasEntry, StratusGpuCommon.h:276-279 */
typedef struct SgiAtlasEntry {
    int32_t index;   /* which atlas (0..13) */
    int32_t layer;   /* cube layer inside that atlas */
} SgiAtlasEntry;

/* ---- status codes ------------------------------------------------------------------------- */
enum {
    SGI_OK              = 0,
    SGI_ERR_INVALID_ARG = 1,
    SGI_ERR_CUDA        = 2,   /* CUDA runtime/driver error, or no device: there is no CPU path */
    SGI_ERR_NOT_BOUND   = 3,   /* a buffer the stage needs was never bound */
    SGI_ERR_UNSUPPORTED = 4
};

enum { SGI_GATHER_STOCHASTIC = 0, SGI_GATHER_TILE_LIST = 1 };

/* which library-owned plane sgi_get_plane returns */
enum {
    SGI_PLANE_GI_COLOR       = 0,  /* vplGIFbo attachment 0: albedo + 1e-5            half4 */
    SGI_PLANE_GI_RESERVOIR   = 1,  /* vplGIFbo attachment 1: (radiance, sum weights)  half4 */
    SGI_PLANE_DENOISE_RES_A  = 2,  /* vplGIDenoisedFbo1 attachment 2 (reservoir)      half4 */
    SGI_PLANE_DENOISE_RES_B  = 3,  /* vplGIDenoisedFbo2 attachment 2                  half4 */
    SGI_PLANE_GI_ILLUM       = 4,  /* this frame's `giColor` after the final pass     half4 */
    SGI_PLANE_HISTORY        = 5,  /* this frame's `newHistoryDepth`                  half  */
    SGI_PLANE_PREV_GI_ILLUM  = 6,  /* vplGIDenoisedPrevFrameFbo attachment 1          half4 */
    SGI_PLANE_PREV_HISTORY   = 7,  /* vplGIDenoisedPrevFrameFbo attachment 3          half  */
    SGI_PLANE_COMBINED       = 8   /* `combinedColor` of the final pass               half4 */
};

/* ---- create-time parameters (mirrors UpdateWindowDimensions_/InitializeVplData_,
 *      StratusRendererBackend.cpp:297-310, 436-445, 510-589) -------------------------------- */
typedef struct SgiCreateParams {
    int32_t width;               /* frame width  (multiple of 5 for the tile stages)                */
    int32_t full_height;         /* frame height                                                    */
    int32_t y_offset;            /* first frame row owned by this handle (0 for a whole frame)      */
    int32_t rows;                /* rows owned (== full_height for a whole frame); multiple of 5    */
    int32_t cube_res;            /* vplShadowCubeMapX/Y, reference 32 (StratusRendererBackend.h:349) */
    int32_t csm_res;             /* cascadeResolutionXY                                             */
    int32_t denoise_iters;       /* maxIterations, reference 4 (StratusRendererBackend.cpp:1914)    */
    int32_t device;              /* CUDA device ordinal                                             */
    void*   stream;              /* cudaStream_t to run on; NULL = library creates its own          */
} SgiCreateParams;

/* ---- every uniform the five shaders read ------------------------------ */
typedef struct SgiFrameParams {
    float   invProjectionView[16];      /* column-major, as glm stores it                           */
    float   prevInvProjectionView[16];  /* bound by the reference, unused by the live shader text   */
    float   viewPosition[3];
    float   time;                       /* `time` uniform: milliseconds as float (BE:1823-1824)     */
    double  deltaSeconds;               /* framesPerSecond = float(1.0/deltaSeconds) (BE:1908)      */
    float   minGiOcclusionFactor;       /* RendererSettings default 0.95                             */
    float   minRoughness;               /* 0.08                                                      */
    int32_t usePerceptualRoughness;     /* 1                                                         */
    float   fogDensity;
    float   fogColor[3];
    float   infiniteLightDirection[3];
    float   infiniteLightColor[3];      /* GetLuminance() = colour * intensity                       */
    float   cascadeProjViews[SGI_NUM_CASCADES][16];   /* projectionViewSample, column-major          */
    float   cascadePlanes[3][4];
    int32_t totalNumLights;             /* N handed to the cull (<= 10000)                           */
    int32_t gatherMode;                 /* SGI_GATHER_*                                              */
} SgiFrameParams;

/* ---- caller-owned device buffers (the SSBOs and textures the reference binds) -------------- */
typedef struct SgiBuffers {
    /* current-frame G-buffer, state_.currentFrame.* (BE:384-424) */
    const void* depth;
    const void* normal;
    const void* albedo;
    const void* baseReflectivity;
    const void* roughnessMetallicAmbient;
    const void* velocity;
    const void* id;
    /* previous-frame G-buffer, state_.previousFrame.{normals,id,depth} */
    const void* prevNormal;
    const void* prevId;
    const void* prevDepth;
    /* ssaoOcclusionBlurredTexture (BE:501) and lightingColorBuffer (BE:462) */
    const void* ssao;
    const void* screen;
    /* where `combinedColor` goes (the blit into lightingFbo, BE:1946); may alias `screen` */
    void*       lightingOut;
    /* VPL SSBOs, state_.vpls.* (StratusRendererBackend.h:352-357) */
    const SgiVplData*            vplData;               /* [N]            in  */
    SgiVplData*                  vplUpdatedData;        /* [4480]         out */
    int32_t*                     vplVisibleIndices;     /* [N+1]          out: [0]=V, [1..V]=indices */
    SgiAtlasEntry*               shadowDiffuseIndices;  /* [4480]         out (built on device)      */
    SgiVplStage1PerTileOutputs*  vplStage1Results;      /* [tiles]        out */
    SgiVplStage2PerTileOutputs*  vplVisiblePerTile;     /* [tiles]        out */
    /* shadow-map cache assignment per *input* VPL.  The reference builds
     * shadowDiffuseIndices on the CPU from a blocking read-back of the visible
     * list (BE:1609-1610, 1675-1701); here the table is indexed by original VPL
     * index and compacted on the device by the cull, so no fence is needed. */
    const SgiAtlasEntry*         atlasEntryByVpl;       /* [N]            in  */
    /* vplSmapCache_.buffers[0..13] (BE:2413-2421) */
    const void* shadowCubeMaps[SGI_MAX_TOTAL_SHADOW_ATLASES];
    const void* diffuseCubeMaps[SGI_MAX_TOTAL_SHADOW_ATLASES];
    /* frame_->csc.fbo depth attachment (BE:348-352) */
    const void* infiniteLightShadowMap;
    /* optional: device int32 holding the number of valid entries of vplData (written by sgi_vpl_prepass).  When bound,
     * the cull uses min(totalNumLights, *vplCountDevice), so the host passes an upper bound and never reads N back. */
    const int32_t* vplCountDevice;
} SgiBuffers;

/* ---- VPL pre-pass: the CPU loop of UpdatePointLights_ (BE:1359-1392) and InitVplFrameData_
 *      (BE:1281-1292) on the device ----------------------------------------------------------------------------------------
 * What the engine keeps per VirtualPointLight (position, radius, far plane, intensity: Source/Engine/StratusLight.h) plus
 * the slot its cube maps occupy in the shadow-map cache. */
typedef struct SgiLightRecord {
    float   position[3];
    float   radius;
    float   farPlane;
    float   intensity;
    int32_t atlasIndex;        /* GpuAtlasEntry of the light's cube maps (StratusGpuCommon.h:276-279) */
    int32_t atlasLayer;
} SgiLightRecord;

#define SGI_MAX_PREPASS_LIGHTS 262144

typedef struct SgiPrepassParams {
    float   viewPosition[3];           /* c.GetPosition() (BE:1360)                                                   */
    int32_t numLights;                 /* virtual point lights of the frame, <= SGI_MAX_PREPASS_LIGHTS                */
    float   frustumPlanes[6][4];       /* frame_->viewFrustumPlanes, IsSphereInFrustum (StratusMath.h:401-412)        */
    const SgiLightRecord* lights;      /* device [numLights]                                                          */
    SgiVplData*    vplDataOut;         /* device [min(numLights, 10000)]: the table InitVplFrameData_ uploads         */
    SgiAtlasEntry* atlasEntryOut;      /* device, same length, or NULL: atlasEntryByVpl for the cull                  */
    int32_t*       orderOut;           /* device, same length, or NULL: index into `lights` of output i               */
    int32_t*       countOut;           /* device int32: N = min(lights in the frustum, 10000)                         */
} SgiPrepassParams;

/* ---- VPL cube shadow / diffuse map production: the producer of the atlases the gather and the
 *      colouring stage read.  Replaces the per-light loop of UpdatePointLights_ (BE:1426-1575: clear the light's cube layer to
 *      depth 1 / colour 0, six 90-degree views through shadow.vs, fragments through shadowVpl.fs:16-36) for scenes handed over
 *      as a list of axis-aligned boxes — the geometry of this repository's procedural workloads.  A texel's fragment is found by
 *      casting the ray through the texel centre (the inverse of the cube-face table) instead of by
 *      rasterising triangles; its value is the shader's: depth = saturate(length(fsPosition - lightPos) / farPlane) stored as
 *      unorm16, colour = the surface's base colour stored as unorm8.  A ray that hits nothing keeps the cleared texel. */
typedef struct SgiSceneBox {       /* 32 bytes */
    float   lo[3];
    float   hi[3];
    uint8_t albedo[4];             /* rgb used, a ignored (the diffuse texel is written with a = 255)         */
    int32_t insideOut;             /* 1: the box is a room seen from inside (its far faces are the surfaces)   */
} SgiSceneBox;

typedef struct SgiCubeRenderParams {
    int32_t numBoxes;                  /* <= SGI_MAX_SCENE_BOXES                                                     */
    int32_t numMaps;                   /* lights whose maps are rendered by this call (maxShadowUpdatesPerFrame = 5
                                          per frame in the reference, StratusRendererBackend.h:378; any count here)  */
    const SgiSceneBox* boxes;          /* device [numBoxes]                                                          */
    const int32_t* vplIndices;         /* device [numMaps]: indices into the bound vplData / atlasEntryByVpl tables  */
    void*   stream;                    /* CUDA stream to enqueue on; NULL = the handle's stream (see sgi_early_begin) */
} SgiCubeRenderParams;
#define SGI_MAX_SCENE_BOXES 256

typedef struct SgiHandle_* SgiHandle;

/* ---- multi-GPU row bands with peer-to-peer halos (no reference counterpart — the
 *      reference is single-GPU) ------------------------------------------------------------------
 * A band handle HOLDS frame rows [y_offset, y_offset+rows) of every per-pixel plane but COMPUTES only
 * [compute_y0, compute_y0+compute_rows).  The held rows outside the computed range are halo: the caller
 * supplies the G-buffer there (it is an input).  The denoiser's reservoir halo rows are either computed by
 * the band itself (SGI_HALO_RECOMPUTE) or written by the neighbouring band's kernels directly into this
 * handle's memory through NVLink peer stores, followed by a release flag the consumer pass waits on
 * (SGI_HALO_EXCHANGE); the GI halo rows (next frame's reprojection) always arrive from the neighbours'
 * final passes.  The final lit rows of every band end up in the root band's frame image (the "gather").
 * Handles may live in one process (peer access) or one process per GPU (CUDA IPC). */
#define SGI_HALO_ROWS 20   /* SGI_HALO_EXCHANGE: >= 16 = reach of the widest reference à-trous pass (2*(multiplier+1), BE:1920), on the 5-row tile grid */
#define SGI_HALO_ROWS_RECOMPUTE 30   /* SGI_HALO_RECOMPUTE: >= 28 = summed reach of the reference's three à-trous passes (4 + 8 + 16) */
enum { SGI_PEER_UP = 0 /* band rank-1 */, SGI_PEER_DOWN = 1 /* band rank+1 */, SGI_PEER_ROOT = 2 /* owner of the frame image */ };

typedef struct SgiBandParams {
    int32_t compute_y0;        /* first frame row this handle computes (>= y_offset, multiple of 5)        */
    int32_t compute_rows;      /* rows it computes (multiple of 5)                                          */
    int32_t rank;              /* band index, 0 = top                                                        */
    int32_t world;             /* number of bands                                                            */
    int32_t owns_frame_image;  /* 1: allocate the full_height x width half4 frame images bands store into    */
    int32_t halo_mode;         /* SGI_HALO_EXCHANGE or SGI_HALO_RECOMPUTE (below); the same on every band    */
} SgiBandParams;
/* How a band gets the reservoir rows its à-trous taps read beyond its own rows.
 *   SGI_HALO_EXCHANGE:  the neighbours' kernels store the boundary rows of EVERY pass into this handle's halo rows (one exchange
 *                       and one pair of flags per pass: three per frame with four denoise passes).
 *   SGI_HALO_RECOMPUTE: the band computes the gather and the non-final passes on as many of its halo rows as the later passes
 *                       will read — pass p on the computed rows +- sum_{j>p} 2^(j+1) rows (28, 24, 16, 0 with four passes) — from the
 *                       G-buffer halo rows the caller supplies anyway.  Same arithmetic on the same inputs as the neighbour, so the
 *                       same bytes; +10 % work on a 270-row band and no exchange or wait inside the frame.  Only the final pass's GI
 *                       rows (next frame's reprojection) and the lit rows (frame image) still travel.  Needs
 *                       2^(denoise_iters+1) - 4 held rows beyond every inner edge (28 -> 30 on the tile grid). */
enum { SGI_HALO_EXCHANGE = 0, SGI_HALO_RECOMPUTE = 1 };

typedef struct SgiPeerInfo {   /* plain data; exchanged between band processes (e.g. all_gather) */
    uint8_t  mem_handle[64];   /* cudaIpcMemHandle_t of the handle's shared slab                             */
    uint64_t local_base;       /* slab address in the exporting process (used when pid matches)              */
    int64_t  off_res[2];       /* vplGIDenoisedFbo1/2 reservoir attachments                                  */
    int64_t  off_gi[2];        /* GI ping-pong                                                               */
    int64_t  off_hist[2];      /* history ping-pong                                                          */
    int64_t  off_flags;        /* int32[SGI_BAND_FLAGS]                                                      */
    int64_t  off_image[2];     /* frame images (owner only, else -1), alternating by frame parity            */
    int32_t  y_offset, rows, compute_y0, compute_rows, rank, device, pid, _pad;
} SgiPeerInfo;

/* create / destroy / resize — UpdateWindowDimensions_ (BE:436) and InitializeVplData_ (BE:297) */
int  sgi_create(const SgiCreateParams* params, SgiHandle* out);
int  sgi_destroy(SgiHandle h);
int  sgi_resize(SgiHandle h, int32_t width, int32_t full_height, int32_t y_offset, int32_t rows);

/* hand over raw device pointers (the "BindBase/BindTexture" calls, BE:1599-1601, 1840-1861, 1888-1900) */
int  sgi_bind_device_buffers(SgiHandle h, const SgiBuffers* buffers);
/* GL-interop hand-over (GLuint names).  Returns SGI_ERR_UNSUPPORTED when the
 * library was built without GL headers (this image has none). */
int  sgi_register_gl(SgiHandle h, const uint32_t* gl_names, int32_t count);

/* renders the cube_res^2 x 6 depth and diffuse maps of the listed lights into their atlas layers (the bound shadowCubeMaps /
 * diffuseCubeMaps at atlasEntryByVpl[vplIndices[k]]); stream-ordered, needs vplData, atlasEntryByVpl and the atlases bound */
int  sgi_render_vpl_cubemaps(SgiHandle h, const SgiCubeRenderParams* rp);

/* ---- deferred direct lighting: pbr.fs, the pass whose output is the `screen` plane of this path ----
 * (csrc/k_direct.cu; parity against the reference's compiled shader text: tests/test_zz_*.py) */
typedef struct SgiPointLight {          /* GpuPointLight / `struct PointLight` (pbr.fs:15-21) */
    float position[4];
    float color[4];
    float radius;
    float farPlane;
    float _pad[2];
} SgiPointLight;

typedef struct SgiDirectParams {
    int32_t numLights;                  /* nonShadowCasters[numLights]                                  */
    int32_t numShadowLights;            /* shadowCasters[numShadowLights], shadowIndices[numShadowLights] */
    int32_t sunEnabled;                 /* INFINITE_LIGHTING_ENABLED                                    */
    int32_t _pad;
    const SgiPointLight* nonShadowCasters;   /* device */
    const SgiPointLight* shadowCasters;      /* device */
    const SgiAtlasEntry* shadowIndices;      /* device */
    float   emissionStrength;
    float   infiniteLightDepthBias, infiniteLightZnear, infiniteLightZfar;      /* pbr.glsl:53-56 */
    float   pointLightAmbientIntensity;                                         /* pbr.glsl:66, 0.003 */
    float   _pad2[3];
    void*   screenOut;                  /* device half4 [rows][W]: boundHDR(colour + emissive * emissionStrength) */
} SgiDirectParams;

/* pbr.fs over the bound G-buffer, cube shadow atlases and sun shadow map; stream-ordered */
int  sgi_direct_lighting(SgiHandle h, const SgiFrameParams* fp, const SgiDirectParams* dp);

/* ---- screen-space ambient occlusion (first half): ssao.fs + ssao_blur.fs, the two passes whose output
 *      is the `ssao` plane of this path (RenderSsaoOcclude_ / RenderSsaoBlur_, StratusRendererBackend.cpp:1160-1204) ----
 * (csrc/k_ssao.cu; parity against the reference's compiled shader text: tests/test_zz_*.py)  Whole-frame handles only. */
typedef struct SgiSsaoParams {
    float   aspectRatio;                /* viewportWidth / viewportHeight                      (BE:1165)          */
    float   projPlaneZDist;             /* 1 / tan(fovy / 2)                                   (BE:1167)          */
    float   intensity;                  /* 5.0 in the reference                                (BE:1170)          */
    float   _pad;
    const void* structureBuffer;        /* device half4 [rows][W]: (dFdx z, dFdy z, z high part, z low part), pbr_geometry_pass.fs:61-75 */
    const void* rotationLookup;         /* device half4 [4][4]: (cos, sin, 0, -) of InitSSAO_'s shuffled angles (BE:759-779)          */
    void*   occlusionScratch;           /* device half [rows][W]: ssaoOcclusionTexture                              */
    void*   ssaoOut;                    /* device half [rows][W]: ssaoOcclusionBlurredTexture = SgiBuffers.ssao     */
} SgiSsaoParams;

/* both passes over the structure buffer; stream-ordered */
int  sgi_ssao(SgiHandle h, const SgiSsaoParams* sp);

/* ---- temporal anti-aliasing (second half): taa.fs, PerformTaaPostFx_ (BE:2324-2354) ----
 * (csrc/k_taa.cu; parity against the reference's compiled shader text: tests/test_zz_*.py)  Whole-frame handles only. */
typedef struct SgiTaaParams {
    const void* screen;                 /* device rgba8 [rows][W]: finalScreenBuffer (the FXAA target, BE:694-699)                   */
    const void* prevScreen;             /* device rgba8 [rows][W]: previousFrameBuffer (RGB8, BE:450-455; alpha ignored)             */
    const void* velocity;               /* device half2 [rows][W]: currentFrame.velocity                                             */
    void*   colorOut;                   /* device half4 [rows][W]: taaFbo (rgb, 0.0625)                                              */
    void*   historyOut;                 /* device rgba8 [rows][W] or NULL: the blit into previousFrameBuffer (BE:2347-2353); may not alias prevScreen */
} SgiTaaParams;

int  sgi_taa(SgiHandle h, const SgiTaaParams* tp);

/* VPL pre-pass: lights whose sphere touches the view frustum, ordered by distance to the viewer (equal distances keep
 * their input order, as std::multiset insertion does), the nearest 10000 written as the cull's input table.  Stream-ordered;
 * bind countOut as SgiBuffers.vplCountDevice and pass totalNumLights = min(numLights, 10000) to run a frame without
 * reading N back. */
int  sgi_vpl_prepass(SgiHandle h, const SgiPrepassParams* pp);

/* stage entry points; all stream-ordered, none blocks the host */
int  sgi_cull(SgiHandle h, const SgiFrameParams* fp);          /* viscull_vpls.cs (a1)               */
int  sgi_color(SgiHandle h, const SgiFrameParams* fp);         /* vpl_light_color.cs (a2)            */
int  sgi_tile_lists(SgiHandle h, const SgiFrameParams* fp);    /* ..._stage1.cs + ..._stage2.cs (a3,a4) */
int  sgi_gather(SgiHandle h, const SgiFrameParams* fp);        /* vpl_pbr_gi.fs (a5)                  */
int  sgi_denoise_pass(SgiHandle h, const SgiFrameParams* fp, int32_t pass);  /* one vpl_pbr_gi_denoise.fs pass */
int  sgi_denoise(SgiHandle h, const SgiFrameParams* fp);       /* all passes + blit + history swap (a6) */
/* all of the above on one stream: cull, colour, tile lists, gather, denoise */
int  sgi_frame(SgiHandle h, const SgiFrameParams* fp);

/* row bands with peer-to-peer halos: set up (re-allocates the shared planes), export, attach neighbours / root.
 * After sgi_band_setup every stage entry point computes only the band's rows, pushes its boundary rows to the
 * attached neighbours and orders itself against them with device-side flags.  A band's lit rows travel to the owner's
 * frame image after its final pass (copy engine, overlapping the next frame; two images alternate by frame parity), and
 * the owner's stream waits for them one frame late.  sgi_frame_image makes the owner's stream wait for every band's rows
 * of the last frame started on the handle and returns that image (stream-ordered; every band must have been given the
 * frame, and sgi_synchronize on a band also drains its transfer). */
int  sgi_band_setup(SgiHandle h, const SgiBandParams* bp);
int  sgi_band_export(SgiHandle h, SgiPeerInfo* out);
int  sgi_band_attach(SgiHandle h, int32_t which, const SgiPeerInfo* peer);
int  sgi_frame_image(SgiHandle h, void** out_ptr, int64_t* out_bytes);
/* nonzero when a peer wait timed out since creation (a neighbour died or never signalled) */
int  sgi_band_timeouts(SgiHandle h, int32_t* out_count);
/* Diagnostics of the band protocol (blocking).  out[0], out[1]: last epochs received from the upper / lower neighbour; out[2]:
 * time-outs; out[3]: reprojection misses; out[5], out[6], out[7]: in-kernel waits since the last call that found the neighbour's
 * flag behind — how many, their summed and their longest duration, in units of 1024 SM clock cycles. */
int  sgi_band_stats(SgiHandle h, int32_t out[8]);
/* Motion limit of a band.  The final denoise pass reads the PREVIOUS frame's normal / depth / id / GI / history at
 * uv - velocity (vpl_pbr_gi_denoise.fs:280,352-360).  A band only has the rows it holds: its own, plus the halo rows, whose
 * G-buffer the caller supplies and whose GI / history the neighbour's final pass stores.  The halo height is the caller's
 * choice (y_offset / rows of the handle; at least the à-trous reach, which sgi_band_setup enforces), so it must also cover
 * the largest vertical motion of a frame: halo >= ceil(max |velocity.y| * full_height) + 1 rows.  Taps beyond the held rows
 * are clamped to them and counted; a nonzero count means this band's image differs from the whole-frame result there. */
int  sgi_band_reprojection_misses(SgiHandle h, int32_t* out_count);

/* ---- the cull where the reference runs it: EARLY.  In RenderScene the visibility cull is part of UpdatePointLights_ (BE:1982 ->
 *      BE:1395), followed by the cube-map updates, the geometry pass, SSAO and the direct lighting before the GI pass consumes its
 *      result (BE:2043-2044).  These two calls give a caller the same freedom on a second CUDA stream: the cull (and, when the
 *      next frame's atlases are already final, the colouring stage) of frame i+1 runs under the denoise passes of frame i, and
 *      sgi_frame(i+1) starts at the gather.  Outputs go to the same caller-owned buffers as sgi_cull / sgi_color.
 *        sgi_early_begin(h, stream)     `stream` waits until the frame last enqueued on this handle has finished READING the
 *                                       compacted tables and the cube atlases (its gather and tile-list stages).  After this call
 *                                       the caller may enqueue, on `stream`, its uploads of the next frame's VPL table / CSM and
 *                                       its cube-map refresh (e.g. sgi_render_vpl_cubemaps with the same stream).
 *        sgi_cull_early(h, fp, stream, with_color)   enqueues sgi_cull (with_color != 0: + sgi_color) for the next frame on
 *                                       `stream`.  The next sgi_frame on this handle waits for it instead of running its own; its
 *                                       lights / sun / cascade parameters must be the ones given here (else SGI_ERR_INVALID_ARG).
 *                                       An explicit sgi_cull in between supersedes the early result. */
int  sgi_early_begin(SgiHandle h, void* stream);
int  sgi_cull_early(SgiHandle h, const SgiFrameParams* fp, void* stream, int32_t with_color);

/* the blocking variant of the visible count (what BE:1675-1678 maps the buffer for) */
int  sgi_visible_count(SgiHandle h, int32_t* out_count);
/* library-owned planes (device pointers valid until resize/destroy) */
int  sgi_get_plane(SgiHandle h, int32_t which, void** out_ptr, int64_t* out_bytes);
/* blocking device->host copy of a library-owned plane (tests, debugging) */
int  sgi_download_plane(SgiHandle h, int32_t which, void* host_dst, int64_t bytes);
/* clear temporal state (history/prev GI), as after FrameBuffer re-creation (BE:569-585) */
int  sgi_reset_history(SgiHandle h);
int  sgi_synchronize(SgiHandle h);
void* sgi_stream(SgiHandle h);
/* number of kernel launches issued through this handle since creation */
int64_t sgi_launch_count(SgiHandle h);
const char* sgi_last_error(SgiHandle h);
const char* sgi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* STRATUSGI_H */
