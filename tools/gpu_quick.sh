#!/bin/bash
# Quick GPU-box visit: selected parity tests + the bench line.  Usage: bash tools/gpu_quick.sh <tag> [pytest -k expr]
tag=${1:-quick}
out=gpurun_out/$tag
mkdir -p $out
if [ -n "$2" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q -k "$2" > $out/pytest.log 2>&1
else
  timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1
fi
echo "pytest exit $?" >> $out/pytest.log
tail -15 $out/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench.json 2> $out/bench.err
echo "bench exit $?"; tail -3 $out/bench.err
python - <<PY
import json
d=json.load(open("$out/bench.json"))
print("ms_per_step", d["ms_per_step"], "e2e_ms", d["e2e"]["ms_per_step"])
for k,v in d["stages"].items(): print(f"  {k:16s} {v['ms']:.4f} ms  hbm_frac {v['hbm_frac']}")
PY
