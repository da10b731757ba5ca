#!/bin/bash
# round 2, first GPU pass: parity tests, default bench line, fused-head microbench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
timeout 300 python scripts/tune/head_bench.py 256/512 64 2,4,8 > gpurun_out/r2a_head_bench.log 2>&1
timeout 200 python scripts/tune/head_bench.py 127/255 256 8,32 >> gpurun_out/r2a_head_bench.log 2>&1
timeout 400 python bench.py --no-cpu > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_head_bench.log; cat gpurun_out/r2a_bench.json | cut -c1-600
