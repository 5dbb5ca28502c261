#!/bin/bash
# Round 2, single-GPU visit after the temporal pre-pass split: GPU suite, bench (look-ahead stream on), ncu launch list of one bench run.
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_visit_r02d.sh r02d'
tag=${1:-r02d}; out=gpurun_out/$tag; mkdir -p $out
bench() {   # $1 = label, $2.. = extra flags
  l=$1; shift
  timeout 300 python bench.py --steps 30 --warmup 5 "$@" > $out/bench_$l.json 2> $out/bench_$l.err
  python -c "
import json; d=json.loads(open('$out/bench_$l.json').read().strip().splitlines()[-1]); print('$l', 'frame ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],3), {k: v.get('ms') for k, v in d['stages'].items()})" || tail -5 $out/bench_$l.err
}
timeout 900 python -m pytest tests -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest exit $?"; tail -8 $out/pytest.log
bench early --no-cpu-baseline
bench noearly --no-cpu-baseline --no-early
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_bench.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("$out/launches.csv")) if len(r)>5]
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr is None: continue
    d=dict(zip(hdr,r)); k=d.get('Kernel Name','')[:60]; 
    try: v=float(d.get('Metric Value','0').replace(',',''))
    except: continue
    u=d.get('Metric Unit','')
    v = v/1000.0 if u in ('ns','nsecond') else v
    agg.setdefault(k,[]).append(v)
for k,v in agg.items(): print(f"{k:60s} n={len(v):3d} mean_us={sum(v)/len(v):9.2f}")
PY
