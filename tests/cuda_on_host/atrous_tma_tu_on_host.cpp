// atrous_tma_tu_on_host.cpp — csrc/k_atrous_tma.cu as its own translation unit of the TMA build of atrous_pass_on_host.cpp
// (host_tu.py rule 7: the tensor maps and cp.async.bulk.tensor.4d are cuda_host_shim.h's stand-ins).
#define SGI_HOST_LOCKSTEP
#include "cuda_host_shim.h"
#include "k_atrous_tma.cu"         // tests/cuda_on_host/_build/k_atrous_tma.cu
