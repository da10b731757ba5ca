#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q > gpurun_out/r2g_pytest_conv.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2g_pytest_conv.log
timeout 300 python scripts/tune/m2_profile.py 256/512 16 > gpurun_out/r2g_m2_profile_256.txt 2>&1
timeout 300 python scripts/tune/m2_profile.py 127/255 1 > gpurun_out/r2g_m2_profile_native_b1.txt 2>&1
timeout 600 python bench.py --no-cpu --no-traffic --no-e2e > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
tail -2 gpurun_out/r2g_pytest_conv.log; grep -E "Name|conv|gemm|xcorr|cudnn|Self CUDA time" gpurun_out/r2g_m2_profile_256.txt | cut -c1-200 | head -30
