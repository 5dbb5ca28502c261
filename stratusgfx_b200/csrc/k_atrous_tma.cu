// k_atrous_tma.cu — the non-final à-trous pass with its taps staged through shared memory by TMA (north_star: "bandwidth-bound
// stencils with shared-memory halos").  It serves every pass that neither waits for a neighbour nor stores reservoir rows into
// one: whole-frame handles and row bands that recompute their halo, final pass included; atrousKernel (k_denoise.cu) serves row
// bands that exchange halos.  The two share the weight, accumulation and epilogue code (sgi_denoise.cuh): same taps, same order,
// same arithmetic, the same bytes (checked on c2 and c4, profiles/r02g; the band tests of tests/test_parity_gpu.py).
//
// A pass with tap spacing s only ever combines pixels of one residue class (x mod s, y mod s): the frame is s*s independent
// "lattice images" of (W/s) x (rows/s) points on which the filter is a dense 5x5 stencil.  A CTA owns kRes consecutive x residues,
// one y residue and a tile of kTX x kTY lattice points: 64 columns x 16 rows = 1024 outputs, four per thread (the same vertical
// register blocking as atrousKernel).  Tile + halo — (kTX+4) x kRes x (kTY+4) texels of the float4 guide plane and of the half4
// reservoir — arrive as TWO cp.async.bulk.tensor.4d boxes: the tensor maps describe a plane as
//     [ s texels (contiguous) ][ W/s lattice columns, stride s texels ][ s row residues, stride W texels ][ rows/s lattice rows, stride s*W texels ]
// so the strided lattice lands DENSE in shared memory, one elected thread issues both copies and everybody waits on one mbarrier.
// A lane's taps are then LDS.128 / LDS.64 at immediate offsets from one base address.
// The tiles cover the lattice points whose taps stay inside the plane, [2, W/s - 2) x [2, rows/s - 2) of every residue class,
// restricted to the rows the handle computes (a row band).  The two lattice points next to each edge (TMA would fill zeros there,
// the shader wants CLAMP_TO_EDGE) and rows / columns beyond the last whole multiple of the spacing keep the clamped global-memory
// path of k_denoise.cu: the CTAs of ITS grid that reach outside the tiles' rectangle are the first CTAs of this kernel's launch.
#include <cuda.h>
#include "sgi_internal.h"
#include "sgi_math.cuh"
#include "sgi_denoise.cuh"

namespace sgi {
namespace {

constexpr int kCols = 64;          // columns of a CTA tile = kRes residues x kTX lattice columns
constexpr int kTY = 16;            // lattice rows of a CTA tile
constexpr int kThreads = kCols * (kTY / kLat);      // 256

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

// FINAL at four CTAs per SM as well (64 registers, 52 bytes of spills in the epilogue): 0.283 -> 0.263 ms for the final pass of c4
// against three CTAs at 80 registers (profiles/r02p).  PARTIAL: see the last tile of a column below.
template <int S, bool FINAL, bool PARTIAL>
__global__ void __launch_bounds__(kThreads, 4)
atrousTmaKernel(const __grid_constant__ CUtensorMap guideMap, const __grid_constant__ CUtensorMap resMap, BandDims bd, const __grid_constant__ DenoiseIO io,
                int ixFirst, int ixLast, int iyFirst, int iyEnd, int nBorder, int gridX) {
    using G = TileGeom<S>;
    __shared__ __align__(128) float4 sGuide[G::kPoints];
    __shared__ __align__(128) uint2 sRes[G::kPoints];
    __shared__ __align__(8) uint64_t bar;
    // The first nBorder CTAs of the launch are the CTAs of atrousKernel's grid that reach outside the rectangle the tiles cover
    // (CLAMP_TO_EDGE taps at the plane's edges, the rows and columns the lattice tiles leave over): clamped global-memory taps, the
    // same code as k_denoise.cu.  One launch instead of two: the ~15 us of a second, dependent kernel (launch gap + one wave of
    // 11-us CTAs) were 10 % of a whole-frame pass and half of a pass on an eighth of the frame (profiles/r02n).
    if ((int)blockIdx.x < nBorder) {
        int bx, by;
        borderCta(io, (int)blockIdx.x, &bx, &by);
        atrousCta<FINAL, S>(bd, io, bx, by, (int)threadIdx.x & (kDnX - 1), (int)threadIdx.x / kDnX);
        return;
    }
    // then (lattice tile column, x residue group) x (lattice tile row, y residue)
    constexpr int groupsX = S / G::kRes;
    const int blk = (int)blockIdx.x - nBorder;
    const int bxT = blk % gridX, byT = blk / gridX;
    const int tx = bxT / groupsX, rx0 = (bxT - tx * groupsX) * G::kRes;
    const int ty = byT / S, ry = byT - ty * S;
    // first lattice point of the tile.  The last tile of a row or column is shifted back to end where the interior ends (a few
    // outputs are computed twice, with the same result).  PARTIAL (chosen by the launcher for planes of few lattice rows: a thin row
    // band with spacing 8 has ~33 of them, and shifting a 16-row tile back would recompute up to 15): the last tile of a COLUMN
    // keeps its place and leaves its lattice rows from iyEnd on uncomputed; its box may then reach past the tensor (zero fill), where
    // only taps of uncomputed rows lie.
    const int ix0 = min(ixFirst + tx * G::kTX, ixLast), iy0 = PARTIAL ? iyFirst + ty * kTY : min(iyFirst + ty * kTY, iyEnd - kTY);
    if (threadIdx.x == 0) {
        mbarInit(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbarExpectTx(&bar, (uint32_t)(G::kPoints * (sizeof(float4) + sizeof(uint2))));
        tmaLoad4d(sGuide, &guideMap, &bar, rx0 * 4, ix0 - 2, ry, iy0 - 2);       // innermost coordinate in floats
        tmaLoad4d(sRes, &resMap, &bar, rx0 * 2, ix0 - 2, ry, iy0 - 2);          // ... in 32-bit words
    }
    const int col = threadIdx.x & (kCols - 1), grp = threadIdx.x >> 6;           // column (ixl, rx) and group of four lattice rows
    const int ixl = col / G::kRes, rx = col - ixl * G::kRes;
    const int x = (ix0 + ixl) * S + rx0 + rx;
    if (FINAL) {
        // the epilogue reads three more planes per output, first touched in this kernel: ask L2 for them while the tile is in flight
        #pragma unroll
        for (int k = 0; k < kLat; ++k) {
            const size_t p = (size_t)(min(iy0 + grp * kLat + k, iyEnd - 1) * S + ry) * bd.W + x;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(io.temporal + p));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(io.screen + p));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(io.illum + p));
        }
    }
    __syncthreads();                     // the barrier's initialisation is visible to the waiters
    if (PARTIAL && iy0 + grp * kLat >= iyEnd) return;       // none of this group's four lattice rows is computed
    mbarWait(&bar, 0);

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
            atrousTap(cen, acc, c, d, sGuide[q], sRes[q]);
        }
    }
    #pragma unroll
    for (int k = 0; k < kLat; ++k) {
        if (PARTIAL && iy0 + grp * kLat + k >= iyEnd) break;
        const int ly = (iy0 + grp * kLat + k) * S + ry;
        const f3 shadowFactor = atrousResolve(acc[k]);
        io.reservoirOut[(size_t)ly * bd.W + x] = packHalf4(shadowFactor.x, shadowFactor.y, shadowFactor.z, 1.0f);
        if (FINAL) finalEpilogue(bd, io, x, ly, shadowFactor);
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
    const cuuint64_t dims[4] = {(cuuint64_t)texelWords * s, (cuuint64_t)(W / s), (cuuint64_t)s, (cuuint64_t)(rows / s)};      // whole multiples of the spacing
    const cuuint64_t strides[3] = {texel * s, texel * W, texel * W * s};              // bytes, dimensions 1..3
    const cuuint32_t box[4] = {(cuuint32_t)(texelWords * res), (cuuint32_t)px, 1u, (cuuint32_t)py};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    return enc(map, isFloat ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<void*>(plane), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// The rectangle of pixels the tiles cover, for a handle that computes band-local rows [c0, c0 + crows): whole lattice rows inside
// that range whose taps stay inside the held plane (two lattice points from every edge; rows / columns beyond the last whole
// multiple of the spacing are left to the border CTAs as well).
template <int S>
bool planTiles(const BandDims& bd, int* ixLo, int* ixHi, int* iyLo, int* iyHi) {
    using G = TileGeom<S>;
    const int W = bd.W, rows = bd.rows;
    if (((size_t)W * 8 * S) % 16) return false;
    const int LX = W / S, LY = rows / S;
    *ixLo = 2; *ixHi = LX - 2;
    *iyLo = max(2, (bd.c0 + S - 1) / S); *iyHi = min(LY - 2, (bd.c0 + bd.crows) / S);
    return *ixHi - *ixLo >= G::kTX && *iyHi - *iyLo >= kTY;
}

template <int S, bool FINAL>
bool launchOne(cudaStream_t stream, const BandDims& bd, const DenoiseIO& io, int nBorder) {
    using G = TileGeom<S>;
    int ixLo, ixHi, iyLo, iyHi;
    if (!planTiles<S>(bd, &ixLo, &ixHi, &iyLo, &iyHi)) return false;
    const int tilesX = (ixHi - ixLo + G::kTX - 1) / G::kTX, tilesY = (iyHi - iyLo + kTY - 1) / kTY;
    CUtensorMap gm, rm;
    if (!latticeMap(&gm, io.guide, 4, bd.W, bd.rows, S, G::kRes, G::kPX, G::kPY, true) || !latticeMap(&rm, io.shadowsIn, 2, bd.W, bd.rows, S, G::kRes, G::kPX, G::kPY, false)) return false;
    const int gridX = tilesX * (S / G::kRes), gridY = tilesY * S;
    // few lattice rows and a last tile that would be shifted back by more than a quarter: leave it partial instead
    const int rem = (iyHi - iyLo) % kTY;
    const bool partial = rem != 0 && tilesY <= 8 && rem <= kTY - kLat;
    auto kf = partial ? atrousTmaKernel<S, FINAL, true> : atrousTmaKernel<S, FINAL, false>;
    kf<<<dim3((unsigned)(nBorder + gridX * gridY)), kThreads, 0, stream>>>(gm, rm, bd, io, ixLo, ixHi - G::kTX, iyLo, iyHi, nBorder, gridX);
    return true;
}

}  // namespace

// CUDA loads kernels lazily on first launch, and that load can stall behind a kernel that is spinning on a band flag
// (sgi_band.cu) — so every kernel of this file is loaded when the handle is created.
template <int S> static bool preloadSpacing() {
    cudaFuncAttributes a; bool ok = true;
    ok = ok && cudaFuncGetAttributes(&a, atrousTmaKernel<S, false, false>) == cudaSuccess;
    ok = ok && cudaFuncGetAttributes(&a, atrousTmaKernel<S, false, true>) == cudaSuccess;
    ok = ok && cudaFuncGetAttributes(&a, atrousTmaKernel<S, true, false>) == cudaSuccess;
    ok = ok && cudaFuncGetAttributes(&a, atrousTmaKernel<S, true, true>) == cudaSuccess;
    return ok;
}
bool preloadAtrousTmaKernels() { return preloadSpacing<2>() && preloadSpacing<4>() && preloadSpacing<8>(); }

static_assert(kThreads == kDnX * kDnY, "the border CTAs run atrousCta with atrousKernel's thread geometry");

// The pixel rectangle [*x0, *x1) x [*y0, *y1) (band-local rows) the TMA tiles of a pass would cover; false when the plane or the
// spacing does not qualify (then atrousKernel runs on everything).
bool planAtrousTma(const BandDims& bd, int s, int* x0, int* x1, int* y0, int* y1) {
    int ixLo, ixHi, iyLo, iyHi;
    const bool ok = s == 2 ? planTiles<2>(bd, &ixLo, &ixHi, &iyLo, &iyHi) : s == 4 ? planTiles<4>(bd, &ixLo, &ixHi, &iyLo, &iyHi) :
                    s == 8 ? planTiles<8>(bd, &ixLo, &ixHi, &iyLo, &iyHi) : false;
    if (!ok || encodeTiled() == nullptr) return false;
    *x0 = ixLo * s; *x1 = ixHi * s; *y0 = iyLo * s; *y1 = iyHi * s;
    return true;
}

// One launch for the whole pass: the nBorder border CTAs of io (CTAs of atrousKernel's grid outside the rectangle of planAtrousTma,
// io.skip* = that rectangle) followed by the lattice tiles.
bool launchAtrousTma(cudaStream_t stream, const BandDims& bd, const DenoiseIO& io, bool final_, int nBorder) {
    const int s = 1 + io.multiplier;
    if (s == 2) return final_ ? launchOne<2, true>(stream, bd, io, nBorder) : launchOne<2, false>(stream, bd, io, nBorder);
    if (s == 4) return final_ ? launchOne<4, true>(stream, bd, io, nBorder) : launchOne<4, false>(stream, bd, io, nBorder);
    if (s == 8) return final_ ? launchOne<8, true>(stream, bd, io, nBorder) : launchOne<8, false>(stream, bd, io, nBorder);
    return false;
}

}  // namespace sgi
