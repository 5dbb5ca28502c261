#!/bin/bash
# Round 2, last visit: the GPU suite on the final tree, then compute-sanitizer memcheck over the small end-to-end tests (new gather,
# merged TMA / border a-trous launches incl. the partial last tile row, coordinate tables).
tag=${1:-r02u}; out=gpurun_out/$tag; mkdir -p $out
if [ "$2" != "bands" ]; then
timeout 110 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "tiny_two_frames or fused_frame_call" > $out/sanitizer_memcheck.log 2>&1
echo "compute-sanitizer memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" $out/sanitizer_memcheck.log | tail -5
timeout 200 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest exit $?"; tail -3 $out/pytest.log
fi
# second call of the visit (r02u2): memcheck over two bands that recompute their halo (TMA tiles restricted to the computed rows,
# partial last tile row, GI push between same-process peers), racecheck over the small end-to-end test
if [ "$2" = "bands" ]; then
  timeout 80 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "recompute_their_halo_reproduce and 2" > $out/sanitizer_memcheck_bands.log 2>&1
  echo "memcheck bands exit $?"; grep -E "ERROR SUMMARY|passed|failed" $out/sanitizer_memcheck_bands.log | tail -3
  timeout 70 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "tiny_two_frames" > $out/sanitizer_racecheck.log 2>&1
  echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed" $out/sanitizer_racecheck.log | tail -3
fi
