#!/bin/bash
# A/B visit: parity tests, then the bench with the persistent two-phase gather and with the legacy gather (SGI_GATHER_LEGACY=1),
# then one ncu --set full capture of the kernels matching <regex>.  Usage: bash tools/gpu_ab.sh <tag> [regex] [skip-tests]
tag=${1:-ab}; rx=${2:-gatherStochastic}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
if [ "$3" != "skip-tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1
  echo "pytest exit $?" >> $out/pytest.log
  tail -8 $out/pytest.log
fi
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench.json 2> $out/bench.err
echo "bench exit $?"; tail -3 $out/bench.err
SGI_GATHER_LEGACY=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_legacy.json 2> $out/bench_legacy.err
python - <<PY
import json
for f in ("bench.json", "bench_legacy.json"):
    try:
        d=json.load(open("$out/"+f))
        print(f, "ms_per_step", round(d["ms_per_step"],4), "e2e_ms", round(d["e2e"]["ms_per_step"],3))
        for k,v in d["stages"].items(): print(f"  {k:16s} {v['ms']:.4f} ms")
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 3 -c 2 -f -o $out/prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu.log 2>&1
tail -3 $out/ncu.log
ls -la $out
