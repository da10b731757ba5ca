#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r3a
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "split_k_clusters or tensor_memory" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -12 gpurun_out/${T}_pytest.log | cut -c1-220
for ts in 1 2; do
  echo "== CONV_TS=$ts"
  HDN_B200_CONV_TS=$ts timeout 300 python scripts/tune/backbone_bench.py 2>&1 | tail -2 | cut -c1-60
  HDN_B200_CONV_TS=$ts timeout 300 python bench.py --workload stream --sequences 2 --frames 61 --no-cpu 2>/dev/null | cut -c70-130
done
