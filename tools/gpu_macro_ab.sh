#!/bin/bash
# A/B of a compile-time experiment (-DSGI_PDL: programmatic dependent launch of the gather -> a-trous chain, csrc/sgi_internal.h;
# -DSGI_ATROUS_V2: 32-bit tap indices + packed FFMA2 accumulation, -DSGI_ATROUS_V3: templated spacing + pointer/immediate
# addressing of clamp-free threads, both in csrc/k_denoise.cu and combinable): the default build, then the build with
# the given flags (compiled on the GPU box), each through the parity suite, the bench and the band probe; the default build is
# restored at the end.   Usage: bash tools/gpu_macro_ab.sh <tag> "<nvcc flags>"     e.g.  ... r02_pdl "-DSGI_PDL"   |   ... r02_atrous "-DSGI_ATROUS_V2 -DSGI_ATROUS_V3"
tag=${1:-macro}; flags=${2:--DSGI_PDL}; out=gpurun_out/$tag; mkdir -p $out
run() {   # $1 = label
  timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_$1.log 2>&1; echo "$1 pytest exit $?"; tail -2 $out/pytest_$1.log
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $out/bench_$1.json 2> $out/bench_$1.err
  python -c "
import json; d=json.load(open('$out/bench_$1.json')); print('$1', 'frame ms', round(d['ms_per_step'],4), {k: v.get('ms') for k, v in d['stages'].items()})"
  timeout 300 python tools/band_probe.py --bands 0,5 --no-whole 2>&1 | head -4 | tail -2
}
python -m stratusgfx_b200.build --force > /dev/null 2>&1; run default
SGI_EXTRA_NVCC_FLAGS="$flags" python -m stratusgfx_b200.build --force > /dev/null 2>&1; run experiment
python -m stratusgfx_b200.build --force > /dev/null 2>&1
