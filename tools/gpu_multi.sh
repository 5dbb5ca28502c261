#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): the NCCL/P2P band tests and the bench at 1..N GPUs.  Usage: bash tools/gpu_multi.sh <tag> <N> [skip-tests]
tag=${1:-multi}; n=${2:-2}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
if [ "$3" != "skip-tests" ]; then
  timeout 900 python -m pytest tests/test_multigpu_nccl.py -x -q > $out/pytest.log 2>&1
  echo "pytest exit $?" >> $out/pytest.log
  tail -6 $out/pytest.log
fi
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_1.json 2> $out/bench_1.err
for g in 2 4 8; do
  if [ $g -le $n ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $g --steps 20 --warmup 5 > $out/bench_$g.json 2> $out/bench_$g.err
    echo "bench $g exit $?"; tail -2 $out/bench_$g.err
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29512 \
        bench.py --gpus $g --steps 20 --warmup 5 --bands apron > $out/bench_apron_$g.json 2> $out/bench_apron_$g.err
    echo "bench apron $g exit $?"
  fi
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$out/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "n", d["n_gpus"], "ms", round(d["ms_per_step"],4), "value", round(d["value"],1), "e2e_ms", round(d["e2e"]["ms_per_step"],3), d["config"]["parallelism"])
    except Exception as e:
        print(f, "unreadable", e)
PY
