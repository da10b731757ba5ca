#!/bin/bash
# round 2: compute-sanitizer memcheck over the convolution, parity and pre-processing tests (new kernels: conv_gemm_ex geometry,
# conv_small, conv3x3_shift, preproc, the FFT kernel variants), racecheck over the convolution tests; per-launch list of a batch-1 frame
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --launch-timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_parity.py -m gpu -q \
    -k "not full_size and not 64-256 and not case8" > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_conv.py -m gpu -q -k "conv_small or strides or shifted or splitk or split_k" > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2_sanitizer_racecheck.log
timeout 300 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__cluster_size --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_frame_b1.csv \
    python scripts/tune/m2_profile.py 127/255 1 > gpurun_out/r2_launches_frame_b1.log 2>&1
tail -4 gpurun_out/r2_sanitizer_memcheck.log; tail -4 gpurun_out/r2_sanitizer_racecheck.log; wc -l gpurun_out/r2_launches_frame_b1.csv
