#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2i_pytest.log
HEAD_BENCH_ITERS=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 24 -c 1 -o gpurun_out/r2i_prof_conv -f \
    python scripts/tune/head_bench.py 256/512 64 8 > gpurun_out/r2i_prof_conv.log 2>&1
timeout 600 python bench.py --no-cpu --no-traffic --no-e2e-m1 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
timeout 600 python bench.py --workload stream --sequences 2 --frames 101 > gpurun_out/r2i_stream.json 2> gpurun_out/r2i_stream.err
grep -E "passed|failed|exit" gpurun_out/r2i_pytest.log | tail -3; cut -c1-200 gpurun_out/r2i_bench.json; cut -c1-200 gpurun_out/r2i_stream.json
