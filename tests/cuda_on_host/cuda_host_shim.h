// cuda_host_shim.h — compile a CUDA kernel of this repository with g++ and run it thread by thread on the host.
// TEST INFRASTRUCTURE (see host_tu.py): it checks a kernel's LOGIC — index arithmetic, exact-arithmetic chains, what is stored
// where — against the oracle / the reference's compiled shader text without a GPU: for kernels that have not met one yet
// (csrc/k_direct.cu) and for kernel changes that have not met a GPU yet.  Every device
// intrinsic is given its IEEE host meaning (-ffp-contract=off: a * b + c is never fused behind the code's back); the MUFU
// approximations become libm calls, so radiance agrees to float rounding, not to the bit.  A thread-by-thread run cannot honour a
// barrier or a warp collective: reaching one aborts the process instead of computing something else.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>          // vector types (float4, uint2, dim3 ...) and the host API declarations; no device code

#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __grid_constant__
#define __launch_bounds__(...)

static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline float __fmaf_rn(float a, float b, float c) { return std::fma(a, b, c); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __saturatef(float a) { return a < 0.0f ? 0.0f : (a > 1.0f ? 1.0f : a); }
static inline float rsqrtf(float a) { return 1.0f / std::sqrt(a); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
#define __expf(x) std::exp((float)(x))                                          // (glibc declares, but does not export, a __expf)
static inline int min(int a, int b) { return a < b ? a : b; }                 // CUDA's global integer overloads
static inline int max(int a, int b) { return a > b ? a : b; }

// the helpers of csrc/ whose body is inline PTX (host_tu.py drops the originals)
static inline void prefetchL2(const void*) {}
static inline float lg2Fast(float x) { return std::log2(x); }                 // lg2.approx.ftz.f32
static inline float ex2Fast(float x) { return std::exp2(x); }                 // ex2.approx.ftz.f32
static inline float sqrtApprox(float x) { return std::sqrt(x); }              // sqrt.approx.ftz.f32 (k_tiles.cu; estimate tier only)
static inline float rcpFast(float x) { return 1.0f / x; }                           // rcp.approx.ftz.f32 (k_gather.cu; radiance tier and margin estimates)
static inline float rsqrtFast(float x) { return 1.0f / std::sqrt(x); }              // rsqrt.approx.ftz.f32
static inline float sqrtFast(float x) { return std::sqrt(x); }                // sqrt.approx.ftz.f32 (k_gather.cu; radiance tier)
static inline unsigned long long pack2(float lo, float hi) {                  // mov.b64 %0, {lo, hi}
    uint32_t a, b; std::memcpy(&a, &lo, 4); std::memcpy(&b, &hi, 4);
    return (unsigned long long)a | ((unsigned long long)b << 32);
}
static inline void unpack2(unsigned long long v, float& lo, float& hi) {
    const uint32_t a = (uint32_t)v, b = (uint32_t)(v >> 32);
    std::memcpy(&lo, &a, 4); std::memcpy(&hi, &b, 4);
}
static inline unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {      // fma.rn.f32x2: IEEE fma per lane
    float a0, a1, b0, b1, c0, c1;
    unpack2(a, a0, a1); unpack2(b, b0, b1); unpack2(c, c0, c1);
    return pack2(std::fma(a0, b0, c0), std::fma(a1, b1, c1));
}
static inline int ldAcquireSys(const int32_t* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }      // ld.acquire.sys.global.s32
static inline void stReleaseSys(int32_t* p, int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }       // st.release.sys.global.s32

// the per-thread built-ins; the launch emulation below sets them before every call of the kernel function
// (inline variables: ONE instance per shared object even when it is built from several translation units — the inline device
// functions of sgi_internal.h are merged across them by the linker and must see the same runtime state as the kernels)
inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __nanosleep(unsigned) {}
static inline long long clock64() { return 0; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
template <typename T> static inline T __ldcg(const T* p) { return *(const volatile T*)p; }
template <typename T> static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned* p, int v) { return __atomic_fetch_add(p, (unsigned)v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicMax(T* p, T v) {
    T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}

#ifdef SGI_HOST_RUNTIME
// ---- the CUDA runtime as the ABI layer (csrc/sgi_abi.cu, csrc/sgi_band.cu) uses it, on host memory: "device" allocations are
// malloc blocks filled with a 0xCD pattern (a read of memory the library never wrote shows up), copies and memsets are synchronous,
// streams and events are opaque tokens, there is exactly one device, and nothing inter-process (CUDA IPC) is available.
#ifndef SGI_HOST_FILL
#define SGI_HOST_FILL 0xCD
#endif
namespace sgihost {
inline cudaError_t hostMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); if (!*p) return cudaErrorMemoryAllocation; std::memset(*p, SGI_HOST_FILL, n ? n : 1); return cudaSuccess; }
inline cudaError_t hostFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t hostToken(void** p) { static char token; *p = &token; return cudaSuccess; }
}
#define cudaGetDeviceCount(p) (*(p) = 1, cudaSuccess)
#define cudaSetDevice(d) cudaSuccess
#define cudaGetLastError() cudaSuccess
#define cudaGetErrorString(e) "host emulation"
#define cudaMalloc(pp, n) sgihost::hostMalloc((void**)(pp), (n))
#define cudaMallocHost(pp, n) sgihost::hostMalloc((void**)(pp), (n))
#define cudaFree(p) sgihost::hostFree((void*)(p))
#define cudaFreeHost(p) sgihost::hostFree((void*)(p))
#define cudaMemcpyAsync(dst, src, n, kind, stream) (std::memcpy((dst), (src), (n)), cudaSuccess)
#define cudaMemcpy(dst, src, n, kind) (std::memcpy((dst), (src), (n)), cudaSuccess)
#define cudaMemset(p, v, n) (std::memset((p), (v), (n)), cudaSuccess)
#define cudaStreamCreateWithFlags(ps, flags) sgihost::hostToken((void**)(ps))
#define cudaStreamCreateWithPriority(ps, flags, prio) sgihost::hostToken((void**)(ps))
#define cudaDeviceGetStreamPriorityRange(lo, hi) (*(lo) = 0, *(hi) = 0, cudaSuccess)
#define cudaEventCreateWithFlags(pe, flags) sgihost::hostToken((void**)(pe))
#define cudaEventRecord(e, s) cudaSuccess
#define cudaStreamWaitEvent(s, e, f) cudaSuccess
#define cudaStreamSynchronize(s) cudaSuccess
#define cudaStreamDestroy(s) cudaSuccess
#define cudaEventDestroy(e) cudaSuccess
#define cudaIpcGetMemHandle(ph, p) (std::memset((ph), 0, sizeof(*(ph))), cudaSuccess)      /* (only same-process peers can attach) */
#define cudaIpcOpenMemHandle(pp, h, f) cudaErrorNotSupported
#define cudaDeviceCanAccessPeer(pc, a, b) (*(pc) = 0, cudaSuccess)
#define cudaDeviceEnablePeerAccess(d, f) cudaErrorNotSupported
#endif

#ifdef SGI_HOST_TMA
// ---- TMA as csrc/k_atrous_tma.cu uses it: a tiled tensor map of rank 4 over 32-bit elements, encoded through the driver entry
// point the file looks up, and cp.async.bulk.tensor.4d into shared memory completing on an mbarrier.  Here the encode function is
// a stand-in that keeps base / dimensions / strides / box in the 128 opaque bytes of the CUtensorMap and REJECTS what the driver
// rejects (so the launcher's fall-back to the clamped kernel is taken for the same planes as on a GPU); the copy is synchronous,
// done by the issuing thread before the CTA's barrier, coordinates outside the tensor read as zero; the mbarrier calls are no-ops.
#include <cuda.h>
namespace sgihost {
struct HostTensorMap { uint64_t magic; const char* base; uint64_t elemBytes, rank, dims[4], strides[3], box[4]; };
static_assert(sizeof(HostTensorMap) <= sizeof(CUtensorMap), "the stand-in lives inside the opaque bytes");
constexpr uint64_t kTensorMapMagic = 0x5347495f544d4150ull;
inline long long tmaLoads = 0, tmaRejected = 0;        // boxes copied / encode calls refused since the library was loaded (one CTA runs at a time)
static inline CUresult encodeTiledOnHost(CUtensorMap* map, CUtensorMapDataType type, cuuint32_t rank, void* addr, const cuuint64_t* dims, const cuuint64_t* strides,
                                         const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapInterleave il, CUtensorMapSwizzle sw, CUtensorMapL2promotion,
                                         CUtensorMapFloatOOBfill fill) {
    auto refuse = [] { ++tmaRejected; return CUDA_ERROR_INVALID_VALUE; };
    if (!map || !addr || rank != 4 || il != CU_TENSOR_MAP_INTERLEAVE_NONE || sw != CU_TENSOR_MAP_SWIZZLE_NONE || fill != CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) return refuse();
    if (type != CU_TENSOR_MAP_DATA_TYPE_FLOAT32 && type != CU_TENSOR_MAP_DATA_TYPE_UINT32) return refuse();
    if (((uintptr_t)addr & 15u) != 0) return refuse();                                                   // globalAddress: 16-byte aligned
    HostTensorMap m; std::memset(&m, 0, sizeof(m));
    m.magic = kTensorMapMagic; m.base = (const char*)addr; m.elemBytes = 4; m.rank = rank;
    for (cuuint32_t i = 0; i < rank; ++i) {
        if (dims[i] == 0 || dims[i] > (1ull << 32)) return refuse();
        if (box[i] == 0 || box[i] > 256 || estr[i] != 1) return refuse();                                 // boxDim <= 256 (element strides other than 1 are not emulated)
        m.dims[i] = dims[i]; m.box[i] = box[i];
    }
    if ((box[0] * m.elemBytes) % 16 != 0) return refuse();                                               // inner box row: a multiple of 16 bytes
    for (cuuint32_t i = 0; i + 1 < rank; ++i) {
        if (strides[i] % 16 != 0 || strides[i] >= (1ull << 40)) return refuse();                          // globalStrides: multiples of 16 bytes
        m.strides[i] = strides[i];
    }
    std::memset(map, 0, sizeof(*map));
    std::memcpy(map, &m, sizeof(m));
    return CUDA_SUCCESS;
}
static inline cudaError_t driverEntryPoint(const char* name, void** fn, cudaDriverEntryPointQueryResult* q) {
    const bool known = std::strcmp(name, "cuTensorMapEncodeTiled") == 0;
    *fn = known ? (void*)&encodeTiledOnHost : nullptr;
    if (q) *q = known ? cudaDriverEntryPointSuccess : cudaDriverEntryPointSymbolNotFound;
    return cudaSuccess;
}
}  // namespace sgihost
extern "C" __attribute__((weak)) long long sgihost_tma_loads() { return sgihost::tmaLoads; }           // for the tests: was the TMA path taken?
extern "C" __attribute__((weak)) long long sgihost_tma_rejected() { return sgihost::tmaRejected; }
#define cudaGetDriverEntryPoint(name, pfn, flags, pq) sgihost::driverEntryPoint((name), (void**)(pfn), (pq))
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)(uintptr_t)p; }
static inline void mbarInit(uint64_t* bar, int) { *bar = 0; }
static inline void mbarExpectTx(uint64_t*, uint32_t) {}
static inline void mbarWait(uint64_t*, uint32_t) {}             // the copies below have completed before the CTA's barrier
static inline void tmaLoad4d(void* dst, const CUtensorMap* map, uint64_t*, int c0, int c1, int c2, int c3) {
    sgihost::HostTensorMap m; std::memcpy(&m, map, sizeof(m));
    if (m.magic != sgihost::kTensorMapMagic) { std::fprintf(stderr, "cuda_host_shim: tmaLoad4d on a tensor map that was not encoded\n"); std::abort(); }
    ++sgihost::tmaLoads;
    uint32_t* out = (uint32_t*)dst;
    const long long c[4] = {c0, c1, c2, c3};
    for (uint64_t i3 = 0; i3 < m.box[3]; ++i3) for (uint64_t i2 = 0; i2 < m.box[2]; ++i2) for (uint64_t i1 = 0; i1 < m.box[1]; ++i1) for (uint64_t i0 = 0; i0 < m.box[0]; ++i0) {
        const long long x0 = c[0] + (long long)i0, x1 = c[1] + (long long)i1, x2 = c[2] + (long long)i2, x3 = c[3] + (long long)i3;
        const bool inside = x0 >= 0 && x0 < (long long)m.dims[0] && x1 >= 0 && x1 < (long long)m.dims[1] && x2 >= 0 && x2 < (long long)m.dims[2] && x3 >= 0 && x3 < (long long)m.dims[3];
        uint32_t v = 0;
        if (inside) std::memcpy(&v, m.base + (uint64_t)x0 * m.elemBytes + (uint64_t)x1 * m.strides[0] + (uint64_t)x2 * m.strides[1] + (uint64_t)x3 * m.strides[2], 4);
        *out++ = v;
    }
}
#endif

// the host API calls a launch function of csrc/ makes around its kernels
#define cudaMemsetAsync(p, v, n, stream) (std::memset((p), (v), (n)), cudaSuccess)
#define cudaFuncSetAttribute(...) cudaSuccess
#define cudaGetDevice(p) cudaErrorNoDevice
#define cudaDeviceGetAttribute(...) cudaErrorNoDevice
#define cudaOccupancyMaxActiveBlocksPerMultiprocessor(...) cudaErrorNoDevice
#define cudaIpcCloseMemHandle(p) cudaSuccess

#ifndef SGI_HOST_LOCKSTEP
// ---- mode 1 (default): one call of the kernel function per thread, CTAs spread over the host's cores.  A thread-by-thread run
// cannot honour a barrier or a warp collective: reaching one aborts.
[[noreturn]] static inline void sgiHostUnsupported(const char* what) {
    std::fprintf(stderr, "cuda_host_shim: %s reached — this launch is not barrier-free, the host run would compute something else\n", what);
    std::abort();
}
static inline void __syncthreads() { sgiHostUnsupported("__syncthreads()"); }
static inline int __syncthreads_or(int) { sgiHostUnsupported("__syncthreads_or()"); }
static inline void __syncwarp(unsigned = 0xffffffffu) { sgiHostUnsupported("__syncwarp()"); }

#define SGI_HOST_LAUNCH(kernel, grid, block, smem, ...)                                                     \
    do {                                                                                                    \
        const dim3 g_ = (grid), b_ = (block);                                                               \
        _Pragma("omp parallel for collapse(2) schedule(dynamic, 1)")                                        \
        for (unsigned by_ = 0; by_ < g_.y; ++by_)                                                           \
            for (unsigned bx_ = 0; bx_ < g_.x; ++bx_) {                                                     \
                gridDim = g_; blockDim = b_; blockIdx.x = bx_; blockIdx.y = by_; blockIdx.z = 0;            \
                for (unsigned ty_ = 0; ty_ < b_.y; ++ty_)                                                   \
                    for (unsigned tx_ = 0; tx_ < b_.x; ++tx_) {                                             \
                        threadIdx.x = tx_; threadIdx.y = ty_; threadIdx.z = 0;                              \
                        kernel(__VA_ARGS__);                                                                \
                    }                                                                                       \
            }                                                                                               \
    } while (0)

#else
// ---- mode 2 (-DSGI_HOST_LOCKSTEP): CTAs one after the other, every CUDA thread of the CTA an OS thread.  __syncthreads is a
// barrier over the CTA's live threads, a warp collective a barrier over the live lanes of its mask with the operands exchanged
// through a per-warp slot array; __shared__ variables are function-local statics (one CTA runs at a time), dynamic shared memory
// is a per-CTA allocation.  A thread that returns from the kernel leaves the barriers it would have taken part in, as on a GPU.
// Limits (none is met by the kernels of csrc/): a lane that exits while others wait in a PARTIAL-mask collective re-evaluates that
// wait with the full mask; __shfl reads of a lane outside the mask return the caller's own value; CTAs of one launch never overlap,
// so a kernel whose CTAs wait for each other would hang here (the band kernels wait for OTHER bands only).
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#define __shared__ static

namespace sgihost {
struct FlexBarrier {                  // the expected count is re-evaluated when a participant leaves
    std::mutex m; std::condition_variable cv; int arrived = 0; unsigned gen = 0;
    template <class F> void wait(F expected) {
        std::unique_lock<std::mutex> l(m);
        const unsigned g = gen;
        if (++arrived >= expected()) { arrived = 0; ++gen; cv.notify_all(); }
        else cv.wait(l, [&] { return gen != g; });
    }
    template <class F> void left(F expected) {
        std::lock_guard<std::mutex> l(m);
        if (arrived > 0 && arrived >= expected()) { arrived = 0; ++gen; cv.notify_all(); }
    }
};
struct Warp { FlexBarrier bar; unsigned long long slot[32]; std::atomic<unsigned> exited{0}; unsigned present = 0; };
struct Cta {
    FlexBarrier bar; std::atomic<int> live{0}; std::atomic<int> orAcc[3]; std::vector<Warp> warps; void* dynSmem = nullptr;
    explicit Cta(int n) : warps((n + 31) / 32) { live = n; for (auto& a : orAcc) a = 0; for (int t = 0; t < n; ++t) warps[t / 32].present |= 1u << (t % 32); }
};
inline Cta* cta = nullptr;
inline thread_local int lane = 0, warpId = 0, orPhase = 0;
static inline Warp& warp() { return cta->warps[warpId]; }
static inline int liveLanes(Warp& w, unsigned mask) { return __builtin_popcount(mask & w.present & ~w.exited.load()); }
template <typename T> static inline unsigned long long bits(T v) { unsigned long long u = 0; std::memcpy(&u, &v, sizeof(T)); return u; }
template <typename T> static inline T from(unsigned long long u) { T v; std::memcpy(&v, &u, sizeof(T)); return v; }
// every live lane of `mask` deposits a value, then reads the slot of `src(lane)`
template <typename T, class F> static inline T exchange(unsigned mask, T v, F src) {
    Warp& w = warp();
    w.slot[lane] = bits(v);
    w.bar.wait([&] { return liveLanes(w, mask); });
    const int sl = src(lane);
    const T r = ((mask >> sl) & 1u) ? from<T>(w.slot[sl]) : v;
    w.bar.wait([&] { return liveLanes(w, mask); });
    return r;
}
static inline void* dynSmem() { return cta->dynSmem; }

// The n OS threads of a launch are created once and walk through the CTAs together (a pool barrier before and after every CTA)
// instead of being created and joined per CTA: thread creation was most of the system time of the host suite.
struct PoolBarrier {
    std::mutex m; std::condition_variable cv; int arrived = 0, n = 0; unsigned gen = 0;
    void wait() {
        std::unique_lock<std::mutex> l(m);
        const unsigned g = gen;
        if (++arrived == n) { arrived = 0; ++gen; cv.notify_all(); }
        else cv.wait(l, [&] { return gen != g; });
    }
};
template <class Body> static void runGrid(dim3 g, dim3 b, size_t smem, Body body) {
    const int n = (int)(b.x * b.y * b.z);
    const unsigned long long total = (unsigned long long)g.x * g.y * g.z;
    if (n <= 0 || total == 0) return;
    std::vector<unsigned long long> mem((smem + 15) / 8 + 2);
    void* const dyn = (void*)(((uintptr_t)mem.data() + 15) & ~(uintptr_t)15);
    Cta* cur = nullptr;
    PoolBarrier pool; pool.n = n;
    std::vector<std::thread> th;
    th.reserve(n);
    for (int t = 0; t < n; ++t) th.emplace_back([&, t] {
        for (unsigned long long i = 0; i < total; ++i) {
            if (t == 0) { delete cur; cur = new Cta(n); cur->dynSmem = dyn; cta = cur; }      // CTAs of a launch run one after the other
            pool.wait();                                                                   // the CTA's state is in place
            Cta& c = *cur;
            gridDim = g; blockDim = b;
            blockIdx.x = (unsigned)(i % g.x); blockIdx.y = (unsigned)((i / g.x) % g.y); blockIdx.z = (unsigned)(i / ((unsigned long long)g.x * g.y));
            threadIdx.x = t % b.x; threadIdx.y = (t / b.x) % b.y; threadIdx.z = t / (b.x * b.y);
            lane = t % 32; warpId = t / 32; orPhase = 0;
            body();
            Warp& w = c.warps[warpId];
            w.exited.fetch_or(1u << lane);
            w.bar.left([&] { return liveLanes(w, 0xffffffffu); });     // (a waiting partial-mask collective re-checks on its own arrival)
            c.live.fetch_sub(1);
            c.bar.left([&] { return c.live.load(); });
            pool.wait();                                                                   // every thread has left the CTA before thread 0 replaces it
        }
    });
    for (auto& x : th) x.join();
    delete cur;
    cta = nullptr;
}
}  // namespace sgihost

static inline void __syncthreads() { sgihost::Cta& c = *sgihost::cta; c.bar.wait([&] { return c.live.load(); }); }
static inline int __syncthreads_or(int p) {
    sgihost::Cta& c = *sgihost::cta;
    const int k = sgihost::orPhase++ % 3;
    if (p) c.orAcc[k].fetch_or(1);
    c.orAcc[(k + 1) % 3].store(0);
    c.bar.wait([&] { return c.live.load(); });
    return c.orAcc[k].load();
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) { sgihost::Warp& w = sgihost::warp(); w.bar.wait([&] { return sgihost::liveLanes(w, mask); }); }
template <typename T> static inline T __shfl_sync(unsigned mask, T v, int srcLane, int width = 32) {
    return sgihost::exchange(mask, v, [=](int l) { return (l & ~(width - 1)) | (srcLane & (width - 1)); });
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int laneMask, int width = 32) {
    return sgihost::exchange(mask, v, [=](int l) { const int s = l ^ laneMask; return (s & ~(width - 1)) == (l & ~(width - 1)) ? s : l; });
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    return sgihost::exchange(mask, v, [=](int l) { const int s = l - (int)delta; return s >= (l & ~(width - 1)) ? s : l; });
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    return sgihost::exchange(mask, v, [=](int l) { const int s = l + (int)delta; return s <= (l | (width - 1)) ? s : l; });
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    sgihost::Warp& w = sgihost::warp();
    w.slot[sgihost::lane] = pred ? 1ull : 0ull;
    w.bar.wait([&] { return sgihost::liveLanes(w, mask); });
    unsigned r = 0;
    const unsigned in = mask & w.present & ~w.exited.load();
    for (int l = 0; l < 32; ++l) if (((in >> l) & 1u) && w.slot[l]) r |= 1u << l;
    w.bar.wait([&] { return sgihost::liveLanes(w, mask); });
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }

#define SGI_HOST_LAUNCH(kernel, grid, block, smem, ...) sgihost::runGrid((grid), (block), (size_t)(smem), [&] { kernel(__VA_ARGS__); })
#endif
