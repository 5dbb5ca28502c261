#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list and one full capture of every kernel.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag> [skip-tests]
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
if [ "$2" != "skip-tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1
  echo "pytest exit $?" >> $out/pytest.log
  tail -5 $out/pytest.log
fi
timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err
echo "bench exit $?"; cat $out/bench.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $out/bench_ref.json 2>> $out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'Kernel' -s 60 -c 12 -o $out/prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_full_bench.log 2>&1
ls -la $out
