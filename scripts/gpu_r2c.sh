#!/bin/bash
# round 2, third GPU pass: tests, bench, ncu launch list + full captures of the fused chain's kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest.log
timeout 300 python scripts/tune/head_bench.py 256/512 64 8 > gpurun_out/r2c_head_bench.log 2>&1
timeout 600 python bench.py --no-cpu > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
HEAD_BENCH_ITERS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 264 -c 88 --csv --log-file gpurun_out/r2c_launches_fused.csv \
    python scripts/tune/head_bench.py 256/512 64 8 > gpurun_out/r2c_launches_fused.log 2>&1
HEAD_BENCH_ITERS=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 24 -c 3 -o gpurun_out/r2c_prof_conv -f \
    python scripts/tune/head_bench.py 256/512 64 8 > gpurun_out/r2c_prof_conv.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:xcorr -s 2 -c 2 -o gpurun_out/r2c_prof_xcorr -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-traffic > gpurun_out/r2c_prof_xcorr.log 2>&1
grep -E "passed|failed|exit" gpurun_out/r2c_pytest.log | tail -3; cat gpurun_out/r2c_head_bench.log; cut -c1-300 gpurun_out/r2c_bench.json
