// cuda_host_shim.h — compile a CUDA kernel of this repository with g++ and run it thread by thread on the host.
// TEST INFRASTRUCTURE (see host_tu.py): it checks a kernel's LOGIC — index arithmetic, exact-arithmetic chains, what is stored
// where — against the oracle / the reference's compiled shader text without a GPU: for kernels that have not met one yet
// (csrc/k_direct.cu) and for compile-time variants queued for an A/B (csrc/k_denoise.cu, -DSGI_ATROUS_V2).  Every device
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
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }                 // CUDA's global integer overloads
static inline int max(int a, int b) { return a > b ? a : b; }

// the helpers of csrc/ whose body is inline PTX (host_tu.py drops the originals)
static inline void prefetchL2(const void*) {}
static inline float lg2Fast(float x) { return std::log2(x); }                 // lg2.approx.ftz.f32
static inline float ex2Fast(float x) { return std::exp2(x); }                 // ex2.approx.ftz.f32
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
static inline int ldAcquireSys(const int32_t* p) { return *(const volatile int32_t*)p; }
static inline void stReleaseSys(int32_t* p, int v) { *(volatile int32_t*)p = v; }

// what a thread-by-thread run cannot do
[[noreturn]] static inline void sgiHostUnsupported(const char* what) {
    std::fprintf(stderr, "cuda_host_shim: %s reached — this launch is not barrier-free, the host run would compute something else\n", what);
    std::abort();
}
static inline void __syncthreads() { sgiHostUnsupported("__syncthreads()"); }
static inline int __syncthreads_or(int) { sgiHostUnsupported("__syncthreads_or()"); }
static inline void __syncwarp() { sgiHostUnsupported("__syncwarp()"); }
static inline void __threadfence_system() {}
static inline void __nanosleep(unsigned) {}
static inline long long clock64() { return 0; }
static inline int atomicAdd(int* p, int v) { int old; _Pragma("omp atomic capture") { old = *p; *p += v; } return old; }

// the per-thread built-ins; the launch emulation below sets them before every call of the kernel function
static thread_local uint3 threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;

// kernel<<<grid, block, smem, stream>>>(args...)  ->  one call per thread, blocks spread over the host's cores
#define SGI_HOST_LAUNCH(kernel, grid, block, ...)                                                           \
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
