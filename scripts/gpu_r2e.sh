#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2e_pytest.log
timeout 300 python scripts/tune/head_bench.py 256/512 64 8 > gpurun_out/r2e_head_bench.log 2>&1
timeout 900 python bench.py --no-cpu --no-traffic --no-e2e-m1 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
HEAD_BENCH_ITERS=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 24 -c 3 -o gpurun_out/r2e_prof_conv -f \
    python scripts/tune/head_bench.py 256/512 64 8 > gpurun_out/r2e_prof_conv.log 2>&1
timeout 300 python scripts/tune/backbone_bench.py > gpurun_out/r2e_backbone_bench.txt 2>&1
timeout 600 python bench.py --workload stream --sequences 2 --frames 101 > gpurun_out/r2e_stream.json 2> gpurun_out/r2e_stream.err
grep -E "passed|failed|exit" gpurun_out/r2e_pytest.log | tail -3; cat gpurun_out/r2e_head_bench.log; cut -c1-200 gpurun_out/r2e_bench.json; tail -5 gpurun_out/r2e_backbone_bench.txt
