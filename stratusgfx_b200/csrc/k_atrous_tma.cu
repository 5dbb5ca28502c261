// k_atrous_tma.cu — the non-final à-trous pass with its taps staged through shared memory by TMA (north_star: "bandwidth-bound
// stencils with shared-memory halos").  It serves the interior of whole-frame handles; atrousKernel (k_denoise.cu) serves the
// border tiles, the final pass and row bands.  Same taps, same order, same arithmetic: the outputs are the same bytes
// (checked on c2 and c4, profiles/r02g).
//
// A pass with tap spacing s only ever combines pixels of one residue class (x mod s, y mod s): the frame is s*s independent
// "lattice images" of (W/s) x (rows/s) points on which the filter is a dense 5x5 stencil.  A CTA owns kRes consecutive x residues,
// one y residue and a tile of kTX x kTY lattice points: 64 columns x 16 rows = 1024 outputs, four per thread (the same vertical
// register blocking as atrousKernel).  Tile + halo — (kTX+4) x kRes x (kTY+4) texels of the float4 guide plane and of the half4
// reservoir — arrive as TWO cp.async.bulk.tensor.4d boxes: the tensor maps describe a plane as
//     [ s texels (contiguous) ][ W/s lattice columns, stride s texels ][ s row residues, stride W texels ][ rows/s lattice rows, stride s*W texels ]
// so the strided lattice lands DENSE in shared memory, one elected thread issues both copies and everybody waits on one mbarrier.
// A lane's taps are then LDS.128 / LDS.64 at immediate offsets from one base address.
// Tiles whose halo would leave the plane (TMA fills zeros there, the shader wants CLAMP_TO_EDGE) and planes whose extent is not a
// multiple of the spacing keep the clamped global-memory path of k_denoise.cu (the launcher runs atrousKernel on them).
#include <cuda.h>
#include "sgi_internal.h"
#include "sgi_math.cuh"

namespace sgi {
namespace {

constexpr float PREVENT_DIV_BY_ZERO = 0.00001f;
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kLat = 4;            // outputs per thread, consecutive lattice rows of one column
constexpr int kCols = 64;          // columns of a CTA tile = kRes residues x kTX lattice columns
constexpr int kTY = 16;            // lattice rows of a CTA tile
constexpr int kThreads = kCols * (kTY / kLat);      // 256

SGI_DEV float lg2Fast(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
SGI_DEV float ex2Fast(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

SGI_DEV uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
SGI_DEV void mbarInit(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count)); }
SGI_DEV void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
SGI_DEV void mbarWait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smemAddr(bar)), "r"(parity) : "memory");
}
SGI_DEV void tmaLoad4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smemAddr(dst)), "l"(map), "r"(smemAddr(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

template <int S> struct TileGeom {
    static constexpr int kRes = S < 4 ? S : 4;           // x residues per CTA (a 16-byte run of reservoir texels at least)
    static constexpr int kTX = kCols / kRes;             // lattice columns per CTA
    static constexpr int kPX = kTX + 4, kPY = kTY + 4;   // with the two-tap halo
    static constexpr int kPoints = kPX * kRes * kPY;
};

template <int S>
__global__ void __launch_bounds__(kThreads, 4)
atrousTmaKernel(const __grid_constant__ CUtensorMap guideMap, const __grid_constant__ CUtensorMap resMap, int W, int tilesX, int tilesY,
                int ix0First, int iy0First, uint2* __restrict__ reservoirOut) {
    using G = TileGeom<S>;
    __shared__ __align__(128) float4 sGuide[G::kPoints];
    __shared__ __align__(128) uint2 sRes[G::kPoints];
    __shared__ __align__(8) uint64_t bar;
    // blockIdx.x enumerates (lattice tile column, x residue group), blockIdx.y (lattice tile row, y residue)
    const int groupsX = S / G::kRes;
    const int tx = blockIdx.x / groupsX, rx0 = (blockIdx.x - tx * groupsX) * G::kRes;
    const int ty = blockIdx.y / S, ry = blockIdx.y - ty * S;
    const int ix0 = ix0First + tx * G::kTX, iy0 = iy0First + ty * kTY;          // first lattice point of the tile (interior tiles only)
    if (threadIdx.x == 0) {
        mbarInit(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbarExpectTx(&bar, (uint32_t)(G::kPoints * (sizeof(float4) + sizeof(uint2))));
        tmaLoad4d(sGuide, &guideMap, &bar, rx0 * 4, ix0 - 2, ry, iy0 - 2);       // innermost coordinate in floats
        tmaLoad4d(sRes, &resMap, &bar, rx0 * 2, ix0 - 2, ry, iy0 - 2);          // ... in 32-bit words
    }
    __syncthreads();                     // the barrier's initialisation is visible to the waiters
    mbarWait(&bar, 0);

    const int col = threadIdx.x & (kCols - 1), grp = threadIdx.x >> 6;           // column (ixl, rx) and group of four lattice rows
    // shared-memory index of the tap (lattice row j, column offset d): ((j * kPX + ixl + d) * kRes + rx), ixl * kRes + rx = col
    const int base = (grp * kLat) * G::kPX * G::kRes + col;                      // tap row c = 0 (two above output 0), d = 0 (two to the left)
    float4 cen[kLat];
    float acc[kLat][4];
    #pragma unroll
    for (int k = 0; k < kLat; ++k) {
        cen[k] = sGuide[base + ((k + 2) * G::kPX + 2) * G::kRes];
        cen[k].w *= kLog2e;
        acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.0f;
    }
    #pragma unroll
    for (int c = 0; c < kLat + 4; ++c) {
        #pragma unroll
        for (int d = 0; d < 5; ++d) {
            const int q = base + (c * G::kPX + d) * G::kRes;
            const float4 gq = sGuide[q];
            const uint2 rv = sRes[q];
            const float rr = h2f((unsigned short)(rv.x & 0xffffu)), rg = h2f((unsigned short)(rv.x >> 16)), rb = h2f((unsigned short)(rv.y & 0xffffu));
            const float tq = gq.w * kLog2e;
            #pragma unroll
            for (int k = 0; k < kLat; ++k) {
                const int dy = c - 2 - k;
                if (dy < -2 || dy > 2) continue;
                float w = 1.0f;
                if (dy != 0 || d != 2) {
                    const float wn = fmaxf(__fmaf_rn(cen[k].z, gq.z, __fmaf_rn(cen[k].y, gq.y, __fmul_rn(cen[k].x, gq.x))), 0.0f);
                    w = ex2Fast(__fmaf_rn(128.0f, lg2Fast(wn), -fabsf(cen[k].w - tq)));
                }
                acc[k][0] = __fmaf_rn(w, rr, acc[k][0]); acc[k][1] = __fmaf_rn(w, rg, acc[k][1]);
                acc[k][2] = __fmaf_rn(w, rb, acc[k][2]); acc[k][3] += w;
            }
        }
    }
    const int ixl = col / G::kRes, rx = col - ixl * G::kRes;
    const int x = (ix0 + ixl) * S + rx0 + rx;
    #pragma unroll
    for (int k = 0; k < kLat; ++k) {
        const int y = (iy0 + grp * kLat + k) * S + ry;
        const float invDenom = __fdividef(1.0f, vmax(PREVENT_DIV_BY_ZERO, acc[k][3]));
        reservoirOut[(size_t)y * W + x] = packHalf4(acc[k][0] * invDenom, acc[k][1] * invDenom, acc[k][2] * invDenom, 1.0f);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encodeTiled() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// plane of `texelWords` 32-bit words per texel, [rows][W], as the lattice tensor of spacing s (see the file header)
bool latticeMap(CUtensorMap* map, const void* plane, int texelWords, int W, int rows, int s, int res, int px, int py, bool isFloat) {
    EncodeTiledFn enc = encodeTiled();
    if (!enc) return false;
    const cuuint64_t texel = (cuuint64_t)texelWords * 4;
    const cuuint64_t dims[4] = {(cuuint64_t)texelWords * s, (cuuint64_t)(W / s), (cuuint64_t)s, (cuuint64_t)(rows / s)};
    const cuuint64_t strides[3] = {texel * s, texel * W, texel * W * s};              // bytes, dimensions 1..3
    const cuuint32_t box[4] = {(cuuint32_t)(texelWords * res), (cuuint32_t)px, 1u, (cuuint32_t)py};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    return enc(map, isFloat ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<void*>(plane), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int S>
bool launchOne(cudaStream_t stream, int W, int rows, const float4* guide, const uint2* in, uint2* out, int* ix0, int* ix1, int* iy0, int* iy1) {
    using G = TileGeom<S>;
    if (W % S || rows % S || ((size_t)W * 8 * S) % 16) return false;
    const int LX = W / S, LY = rows / S;
    // interior tiles: the halo [ix0-2, ix0+kTX+2) x [iy0-2, iy0+kTY+2) stays inside the lattice image
    const int tilesX = (LX - 4) / G::kTX, tilesY = (LY - 4) / kTY;
    if (tilesX <= 0 || tilesY <= 0) return false;
    CUtensorMap gm, rm;
    if (!latticeMap(&gm, guide, 4, W, rows, S, G::kRes, G::kPX, G::kPY, true) || !latticeMap(&rm, in, 2, W, rows, S, G::kRes, G::kPX, G::kPY, false)) return false;
    dim3 grid(tilesX * (S / G::kRes), tilesY * S);
    atrousTmaKernel<S><<<grid, kThreads, 0, stream>>>(gm, rm, W, tilesX, tilesY, 2, 2, out);
    *ix0 = 2; *ix1 = 2 + tilesX * G::kTX; *iy0 = 2; *iy1 = 2 + tilesY * kTY;
    return true;
}

}  // namespace

// Launches the TMA kernel on the interior tiles of a whole-frame plane; returns false when the plane does not qualify (then the
// caller runs atrousKernel on everything).  On success the lattice columns [ix0, ix1) and rows [iy0, iy1) of every residue class
// were written — the pixel rectangle [ix0*s, ix1*s) x [iy0*s, iy1*s); the caller covers the rest.
bool launchAtrousTma(cudaStream_t stream, int W, int rows, int s, const float4* guide, const uint2* in, uint2* out, int* ix0, int* ix1, int* iy0, int* iy1) {
    if (s == 2) return launchOne<2>(stream, W, rows, guide, in, out, ix0, ix1, iy0, iy1);
    if (s == 4) return launchOne<4>(stream, W, rows, guide, in, out, ix0, ix1, iy0, iy1);
    if (s == 8) return launchOne<8>(stream, W, rows, guide, in, out, ix0, ix1, iy0, iy1);
    return false;
}

}  // namespace sgi
