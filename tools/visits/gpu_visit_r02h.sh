#!/bin/bash
# Round 2, 2-GPU diagnosis: the same band geometry with the neighbours attached (p2p) and without (isolated: no pushes, waits, flags):
# stage by stage, what does the exchange itself cost?
tag=${1:-r02h}; out=gpurun_out/$tag; mkdir -p $out
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
    print("  $l: ms", round(d["ms_per_step"],4), "image_equal", c.get("image_equals_single_gpu"), "rows", c.get("band_rows"), "e2e", round(d["e2e"]["ms_per_step"],3))
    print("   stages", {k: v.get("ms") for k,v in (d.get("stages") or {}).items()})
    for s in (c.get("band_spin") or []): print("   spin", s)
    for b in (c.get("band_balancing") or [])[-1:]: print("   balance", b)
except Exception as e:
    print("  $l unreadable:", e); import subprocess; print(subprocess.run(["tail","-5","$out/bench_$l.err"],capture_output=True,text=True).stdout)
PY
}
G=${2:-2}
run c4_p2p $G --balance 0
run c4_isolated $G --bands isolated --balance 0
