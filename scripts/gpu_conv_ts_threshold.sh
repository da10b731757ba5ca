#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for pct in 60 40 25 15 8; do
  echo "== TS min tiles = $pct % of SMs"
  HDN_B200_TS_MIN_TILES_PCT=$pct timeout 300 python scripts/tune/backbone_bench.py 2>&1 | tail -2 | cut -c1-60
  HDN_B200_TS_MIN_TILES_PCT=$pct timeout 300 python bench.py --workload stream --sequences 2 --frames 61 --no-cpu 2>/dev/null | cut -c70-130
done
