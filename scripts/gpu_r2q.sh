#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r2q
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py tests/test_gpu_tool.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
timeout 300 python scripts/tune/backbone_bench.py > gpurun_out/${T}_backbone.log 2>&1; tail -2 gpurun_out/${T}_backbone.log
timeout 600 python bench.py --workload stream --sequences 2 --frames 101 --no-cpu > gpurun_out/${T}_stream.json 2> gpurun_out/${T}_stream.err; cut -c1-220 gpurun_out/${T}_stream.json
timeout 600 python bench.py --workload stream --sequences 8 --frames 101 --lockstep 8 --no-cpu > gpurun_out/${T}_stream_lock8.json 2> gpurun_out/${T}_stream_lock8.err; cut -c1-220 gpurun_out/${T}_stream_lock8.json
timeout 300 python scripts/tune/head_bench.py 256/512 64 8 > gpurun_out/${T}_head_bench.log 2>&1; tail -3 gpurun_out/${T}_head_bench.log | cut -c1-200
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_conv.py -m gpu -q -k "strides or shifted or fused_head_tail" > gpurun_out/${T}_racecheck.log 2>&1; tail -3 gpurun_out/${T}_racecheck.log
