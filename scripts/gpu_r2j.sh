#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
HEAD_BENCH_ITERS=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3x3_shift -s 6 -c 2 -o gpurun_out/r2j_prof_shift -f \
    python scripts/tune/head_bench.py 256/512 64 8 > gpurun_out/r2j_prof_shift.log 2>&1
tail -3 gpurun_out/r2j_prof_shift.log | cut -c1-200
