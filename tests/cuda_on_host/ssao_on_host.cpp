// ssao_on_host.cpp — csrc/k_ssao.cu executed on the host, thread by thread (both kernels are barrier-free).
#include "cuda_host_shim.h"
#include "k_ssao.cu"               // tests/cuda_on_host/_build/k_ssao.cu

extern "C" int host_ssao(int W, int H, const SgiSsaoParams* sp) {
    sgi::BandDims bd;
    bd.W = W; bd.fullH = H; bd.yOffset = 0; bd.rows = H; bd.cubeRes = 0; bd.csmRes = 0; bd.c0 = 0; bd.crows = H;
    return sgi::launchSsao(nullptr, bd, *sp);
}
