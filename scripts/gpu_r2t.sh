#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r2t
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "tensor_memory" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -25 gpurun_out/${T}_pytest.log | cut -c1-250
for ts in 0 1; do
  HDN_B200_CONV_TS=$ts timeout 300 python scripts/tune/m2_profile.py 256/512 32 > gpurun_out/${T}_m2_b32_ts$ts.log 2>&1
  grep -E "conv_gemm|Self CUDA time total" gpurun_out/${T}_m2_b32_ts$ts.log | cut -c1-100,190-330
done
