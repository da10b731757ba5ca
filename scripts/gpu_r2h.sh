#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q > gpurun_out/r2h_pytest_conv.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2h_pytest_conv.log
timeout 300 python scripts/tune/head_bench.py 256/512 64 8 > gpurun_out/r2h_head_bench.log 2>&1
timeout 300 python scripts/tune/backbone_bench.py > gpurun_out/r2h_backbone_bench.txt 2>&1
tail -2 gpurun_out/r2h_pytest_conv.log; cat gpurun_out/r2h_head_bench.log; tail -2 gpurun_out/r2h_backbone_bench.txt
