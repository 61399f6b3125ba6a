#!/bin/bash
# compute-sanitizer over the zstd pipeline's GPU tests with the final build of round 2 (kernels run one after the other under the tool:
# the path on which stage 4 writes its long runs itself)
mkdir -p gpurun_out
OUT=gpurun_out/r02c_sanitizer.txt
echo "# compute-sanitizer on one B200 (gpurun), final build of round 2 (zstd pipeline: early pass, jobs, units, server count)" > $OUT
run() { local tool=$1; shift
  echo "compute-sanitizer --tool $tool python -m pytest $*" >> $OUT
  timeout 700 compute-sanitizer --tool $tool python -m pytest "$@" -x -q > /tmp/san.log 2>&1
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" /tmp/san.log | tail -3 | sed 's/^/    /' >> $OUT
  grep -E "Invalid|Race reported|hazard" /tmp/san.log | sort | uniq -c | head -8 | sed 's/^/    /' >> $OUT
}
run memcheck tests/test_gpu_zstd_decode.py -k "bit_exact or mixed or malformed or large_batch"
run memcheck tests/test_gpu_pages.py tests/test_gpu_zstd_encode.py -k "roundtrip or pages"
run racecheck tests/test_gpu_zstd_decode.py -k "large_batch or mixed"
cat $OUT
