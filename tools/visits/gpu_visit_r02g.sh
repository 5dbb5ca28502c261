#!/bin/bash
# Round 2, single-GPU visit: TMA-staged a-trous kernel against the default (bytes, then time), early chain gated on tile stage 1,
# device-rendered maps with the small-CTA cube kernel, fast-camera band test (4 processes on one GPU).
tag=${1:-r02g}; out=gpurun_out/$tag; mkdir -p $out
bench() {   # $1 = label, $2.. = extra flags
  l=$1; shift
  timeout 300 python bench.py --steps 30 --warmup 5 "$@" > $out/bench_$l.json 2> $out/bench_$l.err
  python -c "
import json; d=json.loads(open('$out/bench_$l.json').read().strip().splitlines()[-1]); print('$l', 'frame ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],3), {k: v.get('ms') for k, v in d['stages'].items()})" || tail -5 $out/bench_$l.err
}
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_zz_cubemaps.py -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest exit $?"; tail -3 $out/pytest.log
timeout 300 python -m pytest tests/test_multigpu_nccl.py -q -k "fast_camera" > $out/pytest_fast_camera.log 2>&1; echo "fast camera exit $?"; tail -3 $out/pytest_fast_camera.log
for w in c2 c4; do
  timeout 200 python tools/debug/atrous_tma_check.py $w /tmp/def_$w.npz > $out/tma_def_$w.log 2>&1
  SGI_ATROUS_TMA=1 timeout 200 python tools/debug/atrous_tma_check.py $w /tmp/tma_$w.npz > $out/tma_tma_$w.log 2>&1
  python -c "
import numpy as np
a=np.load('/tmp/def_$w.npz'); b=np.load('/tmp/tma_$w.npz')
for k in ('lit','a','b'):
    x=a[k].view(np.uint16); y=b[k].view(np.uint16); print('$w', k, 'equal' if np.array_equal(x,y) else 'DIFFER %d' % int((x!=y).sum()), 'nonzero', int((y!=0).sum()))" 2>&1 | tail -4
done
bench early --no-cpu-baseline
SGI_ATROUS_TMA=1 bench early_tma --no-cpu-baseline
bench early_devmaps --no-cpu-baseline --maps device
