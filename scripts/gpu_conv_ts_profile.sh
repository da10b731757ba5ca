#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r2u
for ts in 0 1; do echo "== TS=$ts"; HDN_B200_CONV_TS=$ts timeout 300 python scripts/tune/conv_big.py 2>&1 | tail -7; done
HDN_B200_CONV_TS=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_ts -s 2 -c 1 -o gpurun_out/${T}_prof_ts -f python scripts/tune/conv_big.py 3 > gpurun_out/${T}_prof_ts.log 2>&1
HDN_B200_CONV_TS=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tf32x3 -s 2 -c 1 -o gpurun_out/${T}_prof_ss -f python scripts/tune/conv_big.py 3 > gpurun_out/${T}_prof_ss.log 2>&1
ls -la gpurun_out/${T}_prof_*.ncu-rep
