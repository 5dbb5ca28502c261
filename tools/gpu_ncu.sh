#!/bin/bash
# One ncu --set full capture of the kernels matching <regex>.  Usage: bash tools/gpu_ncu.sh <tag> <regex> [skip] [count]
tag=${1:-ncu}; rx=${2:-Kernel}; skip=${3:-20}; cnt=${4:-4}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f -o $out/prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu.log 2>&1
tail -3 $out/ncu.log
ls -la $out
