#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/exp_conv.log
for n in base halfn; do
  if [ $n = base ]; then unset HDN_B200_LIB; else export HDN_B200_LIB=$PWD/build/libhdn_$n.so; fi
  echo "== $n" >> gpurun_out/exp_conv.log
  timeout 200 python scripts/tune/head_bench.py 256/512 64 8 >> gpurun_out/exp_conv.log 2>&1
done
cat gpurun_out/exp_conv.log | cut -c1-150
