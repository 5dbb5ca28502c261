// prepass_on_host.cpp — csrc/k_prepass.cu in lockstep mode (the rank kernel stages keys through shared memory between barriers).
#define SGI_HOST_LOCKSTEP
#include <vector>
#include "cuda_host_shim.h"
#include "k_prepass.cu"            // tests/cuda_on_host/_build/k_prepass.cu

// sgi_vpl_prepass (csrc/sgi_abi.cu): the key scratch and the in-frustum counter are library-owned
extern "C" int host_vpl_prepass(const SgiPrepassParams* pp) {
    std::vector<unsigned long long> keys((size_t)(pp->numLights > 0 ? pp->numLights : 1));
    int inside = 0;
    return sgi::launchPrepass(nullptr, *pp, keys.data(), &inside);
}
