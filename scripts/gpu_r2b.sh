#!/bin/bash
# round 2, second GPU pass: all parity tests, fused-head microbench, default bench line (ncu traffic + CPU arm), tracker frame rate
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest.log
timeout 300 python scripts/tune/head_bench.py 256/512 64 4,8,16 > gpurun_out/r2b_head_bench.log 2>&1
timeout 600 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
timeout 300 python -m hdn_b200.runner --sequences 2 --frames 60 > gpurun_out/r2b_runner_dev.json 2> gpurun_out/r2b_runner_dev.err
HDN_B200_DEVICE_PREPROC=0 timeout 300 python -m hdn_b200.runner --sequences 2 --frames 60 > gpurun_out/r2b_runner_host.json 2> gpurun_out/r2b_runner_host.err
timeout 300 python -m hdn_b200.runner --sequences 8 --frames 40 --lockstep 8 > gpurun_out/r2b_runner_lock8.json 2> gpurun_out/r2b_runner_lock8.err
grep -E "passed|failed|exit" gpurun_out/r2b_pytest.log | tail -3; cat gpurun_out/r2b_head_bench.log; cut -c1-400 gpurun_out/r2b_bench.json; cut -c1-300 gpurun_out/r2b_runner_dev.json; cut -c1-300 gpurun_out/r2b_runner_host.json
