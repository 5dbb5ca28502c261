#!/bin/bash
# Round 2, final multi-GPU visit (gpurun --gpus 8): the band tests over NVLink (both halo modes at 2, 4, 8 GPUs), c4 at 1, 2, 4, 8 GPUs
# with the default band mode (recompute) and at 8 GPUs with the exchange, c5 at 8 GPUs (bands, and one view per GPU).
tag=${1:-r02s}; out=gpurun_out/$tag; mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu_nccl.py -q -k "nvlink" > $out/pytest_multigpu_8gpus.log 2>&1; echo "pytest exit $?" >> $out/pytest_multigpu_8gpus.log; tail -3 $out/pytest_multigpu_8gpus.log
run() {   # $1 = gpus, $2 = label, $3.. = extra flags
  g=$1; l=$2; shift; shift
  if [ $g -eq 1 ]; then
    timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline "$@" > $out/bench_$l.json 2> $out/bench_$l.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $g --steps 30 --warmup 5 --no-cpu-baseline "$@" > $out/bench_$l.json 2> $out/bench_$l.err
  fi
  echo "bench $l exit $?"
}
run 1 c4_n1
run 2 c4_n2
run 4 c4_n4
run 8 c4_n8
run 8 c4_n8_exchange --bands p2p
run 8 c5_n8_bands --workload c5
run 8 c5_n8_views --workload c5 --split views
python - <<PY
import json
base=None
for l,g in (("c4_n1",1),("c4_n2",2),("c4_n4",4),("c4_n8",8),("c4_n8_exchange",8),("c5_n8_bands",8),("c5_n8_views",8)):
    f="$out/bench_%s.json"%l
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if l=="c4_n1": base=d["ms_per_step"]
        c=d["config"]
        eff = round(base/d["ms_per_step"]/g,3) if l.startswith("c4") else None
        print(l, "ms", round(d["ms_per_step"],4), "eff", eff, "e2e", round(d["e2e"]["ms_per_step"],3), "rows", c.get("band_rows"), "timeouts", c.get("band_wait_timeouts"), "equal", c.get("image_equals_single_gpu"))
        print("   stages", {k: v["ms"] for k,v in d["stages"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
