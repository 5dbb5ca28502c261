#!/bin/bash
# First single-GPU visit of a new round, in one gpurun call (about 25 minutes of box time):
#   1. the GPU parity suite on the default build, plus the opt-in tests of the direct-lighting, SSAO and TAA kernels (never run on a GPU yet)
#   2. the default bench line, then compute-sanitizer memcheck and racecheck on the small end-to-end tests
#   3. A/B of the queued compile-time variants of DESIGN.md section 9, each through the parity suite and the bench
#      (all were shown byte-identical to the default build on the host, tests/test_denoise_kernel_on_host.py)
#   4. the default build restored
# Usage: /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_next_visit.sh r02a'
tag=${1:-next}; out=gpurun_out/$tag; mkdir -p $out
bench() {   # $1 = label
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $out/bench_$1.json 2> $out/bench_$1.err
  python -c "
import json; d=json.loads(open('$out/bench_$1.json').read().strip().splitlines()[-1]); print('$1', 'frame ms', round(d['ms_per_step'],4), {k: v.get('ms') for k, v in d['stages'].items()})"
}
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_default.log 2>&1; echo "default pytest exit $?"; tail -2 $out/pytest_default.log
SGI_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests/test_zz_direct_lighting.py tests/test_zz_ssao.py tests/test_zz_taa.py -m gpu -q > $out/pytest_direct.log 2>&1; echo "direct-lighting + SSAO + TAA pytest exit $?"; tail -3 $out/pytest_direct.log
bench default
# compute-sanitizer on the small end-to-end test (the host-side sanitizer passes of tools/sanitize_on_host.py found nothing; this is the device's word)
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "tiny_two_frames or fused_frame_call" > $out/sanitizer_$tool.log 2>&1
  echo "compute-sanitizer $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $out/sanitizer_$tool.log | tail -3
done
for v in "KEY64:-DSGI_STAGE2_KEY64" "V3:-DSGI_ATROUS_V3" "V2V3:-DSGI_ATROUS_V2 -DSGI_ATROUS_V3" "PDL:-DSGI_PDL" "ALL:-DSGI_PDL -DSGI_ATROUS_V2 -DSGI_ATROUS_V3 -DSGI_STAGE2_KEY64"; do
  label=${v%%:*}; flags=${v#*:}
  SGI_EXTRA_NVCC_FLAGS="$flags" python -m stratusgfx_b200.build --force > $out/build_$label.log 2>&1 || { echo "$label: build failed"; continue; }
  timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > $out/pytest_$label.log 2>&1; echo "$label pytest exit $?"; tail -1 $out/pytest_$label.log
  bench $label
done
python -m stratusgfx_b200.build --force > /dev/null 2>&1
bench default_again
