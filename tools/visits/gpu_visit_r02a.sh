#!/bin/bash
# Round 2, first single-GPU visit (one gpurun call).  Every library variant was compiled in the build container
# (tools/prebuild_variants.py -> build/variants/<label>/libstratusgi.so), so an A/B here is a file copy.
#   1. GPU parity suite on the default build, the never-run direct-lighting / SSAO / TAA tests included
#   2. default bench line (with the CPU baseline), box facts (EGL/GL probe, host cores, topology)
#   3. A/B of the queued compile-time variants: parity suite + bench each
#   4. compute-sanitizer memcheck + racecheck on the small end-to-end tests
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_visit_r02a.sh r02a'
tag=${1:-r02a}; out=gpurun_out/$tag; mkdir -p $out
LIB=stratusgfx_b200/csrc/libstratusgi.so
cp $LIB $out/../.lib_default_backup.so
bench() {   # $1 = label, $2.. = extra flags
  l=$1; shift
  timeout 300 python bench.py --steps 30 --warmup 5 "$@" > $out/bench_$l.json 2> $out/bench_$l.err
  python -c "
import json; d=json.loads(open('$out/bench_$l.json').read().strip().splitlines()[-1]); print('$l', 'frame ms', round(d['ms_per_step'],4), {k: v.get('ms') for k, v in d['stages'].items()})"
}
{ nproc; nvidia-smi -L; ldconfig -p | grep -i -e egl -e libGL -e opengl; ls /usr/share/glvnd/egl_vendor.d 2>/dev/null; nvidia-smi topo -m 2>/dev/null | head -20; } > $out/box_facts.txt 2>&1
SGI_RUN_UNVALIDATED=1 timeout 900 python -m pytest tests -m gpu -q > $out/pytest_default.log 2>&1; echo "default pytest (incl. direct/SSAO/TAA) exit $?"; tail -15 $out/pytest_default.log
bench default
for label in KEY64 V3 V2V3 PDL ALL; do
  cp build/variants/$label/libstratusgi.so $LIB
  timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > $out/pytest_$label.log 2>&1; echo "$label pytest exit $?"; tail -1 $out/pytest_$label.log
  bench $label --no-cpu-baseline
done
cp build/variants/default/libstratusgi.so $LIB
bench default_again --no-cpu-baseline
for tool in memcheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "tiny_two_frames or fused_frame_call" > $out/sanitizer_$tool.log 2>&1
  echo "compute-sanitizer $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $out/sanitizer_$tool.log | tail -3
done
