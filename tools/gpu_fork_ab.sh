#!/bin/bash
# A/B of the tile-list side stream in sgi_frame: serial, forked at low priority, forked at equal priority.
# Usage: bash tools/gpu_fork_ab.sh <tag> [skip-tests]
tag=${1:-fork}; out=gpurun_out/$tag; mkdir -p $out
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log; tail -3 $out/pytest.log
fi
for v in serial low equal; do
  case $v in
    serial) export SGI_SERIAL_TILE_LISTS=1; unset SGI_TILE_STREAM_PRIO;;
    low) unset SGI_SERIAL_TILE_LISTS; unset SGI_TILE_STREAM_PRIO;;
    equal) unset SGI_SERIAL_TILE_LISTS; export SGI_TILE_STREAM_PRIO=equal;;
  esac
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $out/bench_$v.json 2> $out/bench_$v.err
  python -c "
import json; d=json.load(open('$out/bench_$v.json')); print('$v', 'frame ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],3), 'stage sum', round(sum(x['ms'] for x in d['stages'].values()),4))"
  timeout 300 python tools/band_probe.py --bands 0,5 --no-whole 2>&1 | head -4 | tail -2
done
