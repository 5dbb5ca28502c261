#!/bin/bash
# Round 2, single-GPU visit: temporal pre-pass under the a-trous passes (own stream), bench with / without the look-ahead stream.
tag=${1:-r02e}; out=gpurun_out/$tag; mkdir -p $out
bench() {   # $1 = label, $2.. = extra flags
  l=$1; shift
  timeout 300 python bench.py --steps 30 --warmup 5 "$@" > $out/bench_$l.json 2> $out/bench_$l.err
  python -c "
import json; d=json.loads(open('$out/bench_$l.json').read().strip().splitlines()[-1]); print('$l', 'frame ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],3), {k: v.get('ms') for k, v in d['stages'].items()})" || tail -5 $out/bench_$l.err
}
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest exit $?"; tail -3 $out/pytest.log
bench early --no-cpu-baseline
bench noearly --no-cpu-baseline --no-early
