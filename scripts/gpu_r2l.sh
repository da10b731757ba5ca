#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q > gpurun_out/r2l_pytest_conv.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2l_pytest_conv.log
timeout 300 python scripts/tune/head_bench.py 256/512 64 8 > gpurun_out/r2l_head_bench.log 2>&1
timeout 300 python scripts/tune/head_bench.py 127/255 256 32 >> gpurun_out/r2l_head_bench.log 2>&1
HEAD_BENCH_ITERS=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3x3_shift -s 6 -c 2 -o gpurun_out/r2l_prof_shift -f \
    python scripts/tune/head_bench.py 256/512 64 8 > gpurun_out/r2l_prof_shift.log 2>&1
tail -2 gpurun_out/r2l_pytest_conv.log; cat gpurun_out/r2l_head_bench.log | cut -c1-220
