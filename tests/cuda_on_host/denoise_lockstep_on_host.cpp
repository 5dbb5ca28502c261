// denoise_lockstep_on_host.cpp — csrc/k_denoise.cu in lockstep mode, for the band runs: bandWaitCta's __syncthreads_or and the
// producer epilogue's __syncthreads are real barriers here.  Provides sgi::launchDenoisePass / launchPrepareGuide to bands_on_host.cpp.
#define SGI_HOST_LOCKSTEP
#include "cuda_host_shim.h"
#include "k_denoise.cu"            // tests/cuda_on_host/_build/k_denoise.cu
