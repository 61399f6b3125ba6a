#!/bin/bash
# compute-sanitizer over the round-2 kernels (CTA-per-block decoders, page chains, tuple walk, batched callers, work-queue encoders)
mkdir -p gpurun_out
OUT=gpurun_out/r02_sanitizer.txt
echo "# compute-sanitizer on one B200 (gpurun), round-2 code" > $OUT
run() { # tool, pytest args...
  local tool=$1; shift
  echo "compute-sanitizer --tool $tool python -m pytest $*" >> $OUT
  timeout 1500 compute-sanitizer --tool $tool python -m pytest "$@" -x -q > /tmp/san.log 2>&1
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" /tmp/san.log | tail -3 | sed 's/^/    /' >> $OUT
  grep -E "Invalid|Race reported|hazard" /tmp/san.log | head -5 | sed 's/^/    /' >> $OUT
}
run memcheck tests/test_gpu_lz4_decode.py
run memcheck tests/test_gpu_pages.py tests/test_gpu_tuples.py tests/test_gpu_batch.py
run memcheck tests/test_gpu_zstd_decode.py -k "bit_exact or mixed or malformed or large_batch"
run memcheck tests/test_gpu_lz4_encode.py tests/test_gpu_zstd_encode.py -k "roundtrip or small_block"
run racecheck tests/test_gpu_pages.py tests/test_gpu_tuples.py
run racecheck tests/test_gpu_lz4_encode.py -k "small_block"
cat $OUT
