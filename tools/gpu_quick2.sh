#!/bin/bash
# quick single-GPU check: parity suite (or a subset: $2), the bench line, optionally an ncu capture of one kernel ($3 = regex, $4 = skip)
tag=${1:-q}; out=gpurun_out/$tag; mkdir -p $out
tests=${2:-tests}
timeout 900 python -m pytest $tests -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest exit $?"; tail -3 $out/pytest.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $out/bench.json 2> $out/bench.err
python -c "
import json; d=json.loads(open('$out/bench.json').read().strip().splitlines()[-1]); print('frame ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],3), {k: v.get('ms') for k, v in d['stages'].items()})" || tail -5 $out/bench.err
if [ -n "$3" ]; then
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$3" -s ${4:-4} -c 2 -f -o $out/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu.log 2>&1
  echo "ncu: $(ls -la $out/prof.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
fi
