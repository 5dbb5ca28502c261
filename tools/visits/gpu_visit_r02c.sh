#!/bin/bash
# Round 2, single-GPU visit: the whole GPU suite (new: cube-map production, early cull), then the bench with and without the
# look-ahead stream and with the refreshed cube maps rendered on the device.
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_visit_r02c.sh r02c'
tag=${1:-r02c}; out=gpurun_out/$tag; mkdir -p $out
bench() {   # $1 = label, $2.. = extra flags
  l=$1; shift
  timeout 300 python bench.py --steps 30 --warmup 5 "$@" > $out/bench_$l.json 2> $out/bench_$l.err
  python -c "
import json; d=json.loads(open('$out/bench_$l.json').read().strip().splitlines()[-1]); print('$l', 'frame ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],3), {k: v.get('ms') for k, v in d['stages'].items()})" || tail -5 $out/bench_$l.err
}
timeout 900 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest exit $?"; tail -8 $out/pytest.log
bench early --no-cpu-baseline
bench noearly --no-cpu-baseline --no-early
bench early_devmaps --no-cpu-baseline --maps device
