#!/bin/bash
# Round 2, single-GPU visit: gather instruction cuts (ftz MUFU helpers, bit-trick cube face + q-dependent texel margin, per-column /
# per-row coordinate tables, multiply-high unit decomposition, two-tier pass-0 epilogue, band code compiled out of whole-frame handles).
tag=${1:-r02k}; out=gpurun_out/$tag; mkdir -p $out
bench() {   # $1 = label, $2.. = extra flags
  l=$1; shift
  timeout 300 python bench.py --steps 30 --warmup 5 "$@" > $out/bench_$l.json 2> $out/bench_$l.err
  python -c "
import json; d=json.loads(open('$out/bench_$l.json').read().strip().splitlines()[-1]); print('$l', 'frame ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],3), {k: v.get('ms') for k, v in d['stages'].items()})" || tail -5 $out/bench_$l.err
}
timeout 900 python -m pytest tests -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest exit $?"; tail -4 $out/pytest.log
bench early --no-cpu-baseline
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gatherStochasticKernel -s 4 -c 2 -f -o $out/prof_gather \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_gather.log 2>&1
echo "gather: $(ls -la $out/prof_gather.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
