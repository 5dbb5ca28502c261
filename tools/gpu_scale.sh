#!/bin/bash
# Scaling visit (gpurun --gpus N): the multi-GPU band tests, then the bench at 1, 2, 4, 8 GPUs (peer-to-peer halos only).
# Usage: bash tools/gpu_scale.sh <tag> <N> [skip-tests]
tag=${1:-scale}; n=${2:-8}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
if [ "$3" != "skip-tests" ]; then
  timeout 600 python -m pytest tests/test_multigpu_nccl.py -x -q > $out/pytest.log 2>&1
  echo "pytest exit $?" >> $out/pytest.log
  tail -4 $out/pytest.log
fi
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_1.json 2> $out/bench_1.err
for g in 2 4 8; do
  if [ $g -le $n ]; then
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $g --steps 20 --warmup 5 > $out/bench_$g.json 2> $out/bench_$g.err
    echo "bench $g exit $?"; tail -2 $out/bench_$g.err
  fi
done
python - <<PY
import json,glob
base=None
for g in (1,2,4,8):
    f="$out/bench_%d.json"%g
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if base is None: base=d["ms_per_step"]
        print("n", d["n_gpus"], "ms", round(d["ms_per_step"],4), "speedup", round(base/d["ms_per_step"],2), "eff", round(base/d["ms_per_step"]/g,3), "e2e_ms", round(d["e2e"]["ms_per_step"],3), "rows", d.get("run", d["config"])["band_rows"], "timeouts", d.get("run", d["config"])["band_wait_timeouts"])
        print("   stages", {k: v["ms"] for k,v in d["stages"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
