#!/bin/bash
# compute-sanitizer over the GPU parity tests (memcheck: all but the full-size property tests; racecheck: the correlation tests).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --launch-timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not full_size" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "xcorr_vs_oracle or xcorr_golden" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_racecheck.log
tail -4 gpurun_out/sanitizer_memcheck.log; tail -4 gpurun_out/sanitizer_racecheck.log
