// cuda_host_shim.h — compile a barrier-free CUDA kernel of this repository with g++ and run it thread by thread on the host.
// TEST INFRASTRUCTURE for kernels that have not met a GPU yet (csrc/k_direct.cu): it checks the kernel's LOGIC against the
// reference's compiled shader text before any GPU minute is spent.  Every device intrinsic used by csrc/sgi_math.cuh is given its
// IEEE host meaning (the translation unit is compiled with -ffp-contract=off, so a * b and a + b are never fused).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>          // vector types (float4, uint2, dim3 ...) and the host API declarations; no device code

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
