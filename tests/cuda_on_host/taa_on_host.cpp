// taa_on_host.cpp — csrc/k_taa.cu executed on the host, thread by thread (the kernel is barrier-free).
#include <vector>
#include "cuda_host_shim.h"
#include "k_taa.cu"                // tests/cuda_on_host/_build/k_taa.cu

extern "C" int host_taa(int W, int H, const SgiTaaParams* tp) {
    sgi::BandDims bd;
    bd.W = W; bd.fullH = H; bd.yOffset = 0; bd.rows = H; bd.cubeRes = 0; bd.csmRes = 0; bd.c0 = 0; bd.crows = H;
    std::vector<float> tables(sgi::kTaaTableFloats);
    return sgi::launchTaa(nullptr, bd, *tp, tables.data());
}
