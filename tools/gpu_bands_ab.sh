#!/bin/bash
# Multi-GPU A/B (gpurun --gpus N): the multi-GPU band tests, then c4 at N GPUs with the neighbours' kernels filling the halos
# (--bands p2p) against the bands recomputing their halo rows (--bands recompute).   Usage: bash tools/gpu_bands_ab.sh <tag> <N> [skip-tests]
tag=${1:-bands_ab}; n=${2:-2}
out=gpurun_out/$tag; mkdir -p $out
if [ "$3" != "skip-tests" ]; then
  timeout 600 python -m pytest tests/test_multigpu_nccl.py -x -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log; tail -3 $out/pytest.log
fi
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $out/bench_1.json 2> $out/bench_1.err
for mode in recompute p2p; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline --bands $mode > $out/bench_${n}_$mode.json 2> $out/bench_${n}_$mode.err
  echo "bench $n $mode exit $?"; tail -2 $out/bench_${n}_$mode.err
done
python - <<PY
import json
base=None
for f,g in (("$out/bench_1.json",1),("$out/bench_${n}_recompute.json",$n),("$out/bench_${n}_p2p.json",$n)):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if base is None: base=d["ms_per_step"]
        c=d.get("run", d["config"])      # (bench lines before the end of round 2 kept these under "config")
        print(f.split("/")[-1], "ms", round(d["ms_per_step"],4), "eff", round(base/d["ms_per_step"]/g,3), "e2e", round(d["e2e"]["ms_per_step"],3), "rows", c["band_rows"], "timeouts", c["band_wait_timeouts"], "equal", c["image_equals_single_gpu"])
        print("   stages", {k: v["ms"] for k,v in d["stages"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
