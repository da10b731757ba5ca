#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x > gpurun_out/r2f_pytest_conv.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_pytest_conv.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_pytest.log
timeout 300 python scripts/tune/backbone_bench.py > gpurun_out/r2f_backbone_bench.txt 2>&1
timeout 300 python scripts/tune/tracker_profile.py > gpurun_out/r2f_tracker_profile.txt 2>&1
timeout 600 python bench.py --workload stream --sequences 2 --frames 101 > gpurun_out/r2f_stream.json 2> gpurun_out/r2f_stream.err
timeout 600 python bench.py --workload stream --sequences 8 --frames 101 --lockstep 8 > gpurun_out/r2f_stream_lock8.json 2> gpurun_out/r2f_stream_lock8.err
grep -E "passed|failed|exit" gpurun_out/r2f_pytest_conv.log gpurun_out/r2f_pytest.log | tail -4; tail -4 gpurun_out/r2f_backbone_bench.txt; tail -25 gpurun_out/r2f_tracker_profile.txt; cut -c1-200 gpurun_out/r2f_stream.json; cut -c1-200 gpurun_out/r2f_stream_lock8.json
