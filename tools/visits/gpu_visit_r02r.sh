#!/bin/bash
# Round 2, evidence visit for the shipped build on one B200 (c4): the GPU suite, the default bench line (with the CPU baseline leg),
# the reference arm once, the ncu launch list of a bench run, one --set full capture per heavy kernel.  Numbers printed under ncu are
# never bench values.
tag=${1:-r02r}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest exit $?"; tail -3 $out/pytest.log
timeout 400 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench exit $?"
python -c "
import json; d=json.loads(open('$out/bench.json').read().strip().splitlines()[-1]); print('frame ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],3), 'roofline', d['roofline']['frac'], d['roofline']['achieved'], 'cpu', d['cpu_baseline'] and d['cpu_baseline'].get('value'), {k: v.get('ms') for k, v in d['stages'].items()})" || tail -5 $out/bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; echo "reference arm exit $?"; tail -c 400 $out/bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_launches.log 2>&1
python tools/ncu_summary.py launches $out/launches.csv 2>/dev/null | head -18
# (atrous: launches 4-6 of atrousTmaKernel = the s = 2, 4 and 8 (final) passes of one frame)
for spec in "gather:gatherStochasticKernel:4:2" "atrous:atrousTmaKernel:3:3" "temporal:temporalKernel:3:2" "stage2:stage2Kernel:3:2" "stage1:stage1Kernel:3:2"; do
  name=${spec%%:*}; rest=${spec#*:}; rx=${rest%%:*}; rest=${rest#*:}; skip=${rest%%:*}; cnt=${rest#*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f -o $out/prof_$name \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_$name.log 2>&1
  echo "$name: $(ls -la $out/prof_$name.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
done
