#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for sh in 1 0; do echo "== conv_shift=$sh"; HEAD_BENCH_SHIFT=$sh timeout 300 python scripts/tune/head_bench.py 256/512 64 8,16 2>&1 | tail -2 | cut -c1-160; done
echo "== native"; for sh in 1 0; do HEAD_BENCH_SHIFT=$sh timeout 300 python scripts/tune/head_bench.py 127/255 256 32 2>&1 | tail -1 | cut -c1-160; done
