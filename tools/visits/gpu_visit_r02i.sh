#!/bin/bash
# Round 2: ncu evidence for the shipped build on one B200 (c4).  1. launch list of a bench run (--metrics gpu__time_duration.sum);
# 2. one --set full capture per heavy kernel (3 launches each, after the warm-up frames).  Numbers printed under ncu are never bench values.
tag=${1:-r02i}; out=gpurun_out/$tag; mkdir -p $out
timeout 300 python bench.py --steps 30 --warmup 5 > $out/bench.json 2> $out/bench.err; tail -c 600 $out/bench.json | head -c 300; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_launches.log 2>&1
for spec in "gather:gatherStochasticKernel:4" "atrous_tma:atrousTmaKernel:6" "atrous_final:atrousKernel<.bool.1:4" "temporal:temporalKernel:3" "stage2:stage2Kernel:3"; do
  name=${spec%%:*}; rest=${spec#*:}; rx=${rest%%:*}; skip=${rest#*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c 3 -f -o $out/prof_$name \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_$name.log 2>&1
  echo "$name: $(ls -la $out/prof_$name.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
done
