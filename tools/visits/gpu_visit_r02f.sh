#!/bin/bash
# Round 2, second 8-GPU visit: the band pipeline with the look-ahead stream (table broadcast, cube-map refresh, cull + colouring of
# frame i+1 off the main stream) and the temporal pre-pass; A/B against the main-stream order, device-rendered maps, the PDL build.
# Usage: gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_visit_r02f.sh r02f'
tag=${1:-r02f}; out=gpurun_out/$tag; mkdir -p $out
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
    print("  $l: ms", round(d["ms_per_step"],4), "image_equal", c.get("image_equals_single_gpu"), "timeouts", c.get("band_wait_timeouts"), "rows", c.get("band_rows"), "e2e", round(d["e2e"]["ms_per_step"],3))
    print("   stages", {k: v.get("ms") for k,v in (d.get("stages") or {}).items()})
    for s in (c.get("band_spin") or []): print("   spin", s)
    for b in (c.get("band_balancing") or [])[-1:]: print("   balance", b)
except Exception as e:
    print("  $l unreadable:", e); import subprocess; print(subprocess.run(["tail","-5","$out/bench_$l.err"],capture_output=True,text=True).stdout)
PY
}
timeout 400 python -m pytest tests/test_multigpu_nccl.py -q -k "fast_camera or (p2p_bands_over_nvlink and 8)" > $out/pytest_multigpu.log 2>&1; echo "multi-GPU pytest exit $?"; tail -3 $out/pytest_multigpu.log
run c4_n8_early 8
run c4_n8_noearly 8 --no-early --no-image-check
run c4_n8_early_devmaps 8 --maps device
cp build/variants/PDL/libstratusgi.so $LIB
run c4_n8_early_pdl 8 --no-image-check
cp build/variants/default/libstratusgi.so $LIB
run c4_n4_early 4 --no-image-check
run c4_n2_early 2 --no-image-check
