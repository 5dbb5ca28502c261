#!/bin/bash
# Round 2, single-GPU visit: TMA-staged a-trous for all three passes of whole-frame handles + compact border launch.
tag=${1:-r02j}; out=gpurun_out/$tag; mkdir -p $out
bench() {   # $1 = label, $2.. = extra flags
  l=$1; shift
  timeout 300 python bench.py --steps 30 --warmup 5 "$@" > $out/bench_$l.json 2> $out/bench_$l.err
  python -c "
import json; d=json.loads(open('$out/bench_$l.json').read().strip().splitlines()[-1]); print('$l', 'frame ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],3), d['e2e'].get('h2d_probe'), {k: v.get('ms') for k, v in d['stages'].items()})" || tail -5 $out/bench_$l.err
}
timeout 900 python -m pytest tests -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest exit $?"; tail -4 $out/pytest.log
bench early --no-cpu-baseline
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_launches.log 2>&1
python tools/ncu_summary.py launches $out/launches.csv | head -16
