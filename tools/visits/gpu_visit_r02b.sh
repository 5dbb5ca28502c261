#!/bin/bash
# Round 2, 8-GPU visit (gpurun --gpus 8): hardware parity of the shipped band protocol at 2/4/8 GPUs, what limits 8 GPUs
# (in-kernel wait account, per-stage times, PDL build, apron mode), and workload c5 on the whole box in both splits.
# Usage: gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_visit_r02b.sh r02b'
tag=${1:-r02b}; out=gpurun_out/$tag; mkdir -p $out
LIB=stratusgfx_b200/csrc/libstratusgi.so
run() {   # $1 = label, $2 = gpus, rest = bench flags
  l=$1; g=$2; shift; shift
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $g --steps 30 --warmup 5 "$@" > $out/bench_$l.json 2> $out/bench_$l.err
  echo "bench $l exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("$out/bench_$l.json").read().strip().splitlines()[-1])
    c=d["config"]
    print("  $l: ms", round(d["ms_per_step"],4), "image_equal", c.get("image_equals_single_gpu"), "timeouts", c.get("band_wait_timeouts"), "rows", c.get("band_rows"))
    print("   stages", {k: v.get("ms") for k,v in (d.get("stages") or {}).items()})
    for s in (c.get("band_spin") or []): print("   spin", s)
    for b in (c.get("band_balancing") or []): print("   balance", b)
except Exception as e:
    print("  $l unreadable:", e); import subprocess; print(subprocess.run(["tail","-5","$out/bench_$l.err"],capture_output=True,text=True).stdout)
PY
}
nvidia-smi topo -m > $out/topo.txt 2>&1
timeout 420 python -m pytest tests/test_multigpu_nccl.py -q > $out/pytest_multigpu_8.log 2>&1; echo "multi-GPU pytest exit $?"; tail -3 $out/pytest_multigpu_8.log
run c4_n8 8
cp build/variants/PDL/libstratusgi.so $LIB
run c4_n8_pdl 8 --no-image-check
cp build/variants/default/libstratusgi.so $LIB
run c4_n8_apron 8 --bands apron --no-image-check
run c4_n4 4 --no-image-check
run c5_n8_bands 8 --workload c5
run c5_n8_views 8 --workload c5 --split views
