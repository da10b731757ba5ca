#!/bin/bash
# round 2, fourth GPU pass: tests + smoke, the full default bench line (M2 + CPU arm + ncu traffic), the reference arm,
# config 5 sweep (1 GPU), config 4 small run
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r2d_smoke.log
timeout 900 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2d_bench_ref.json 2> gpurun_out/r2d_bench_ref.err
for b in 32 64 128 256; do timeout 300 python bench.py --workload win15 --batch $b --no-cpu >> gpurun_out/r2d_win15_sweep.jsonl 2>> gpurun_out/r2d_win15.err; done
timeout 600 python bench.py --workload stream --sequences 2 --frames 101 > gpurun_out/r2d_stream.json 2> gpurun_out/r2d_stream.err
timeout 600 python bench.py --workload stream --sequences 8 --frames 61 --lockstep 8 > gpurun_out/r2d_stream_lock8.json 2> gpurun_out/r2d_stream_lock8.err
grep -E "passed|failed|exit" gpurun_out/r2d_pytest.log | tail -3; tail -2 gpurun_out/r2d_smoke.log; cut -c1-200 gpurun_out/r2d_bench.json; cut -c1-300 gpurun_out/r2d_stream.json
