#!/bin/bash
# round 2 verification pass: all GPU tests, smoke, full default bench line, reference arm, stream, ncu launch list of the bench command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2k_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r2k_smoke.log
timeout 900 python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r2k_bench_ref.json 2> gpurun_out/r2k_bench_ref.err
timeout 600 python bench.py --workload stream --sequences 2 --frames 101 > gpurun_out/r2k_stream.json 2> gpurun_out/r2k_stream.err
timeout 600 python bench.py --workload stream --sequences 8 --frames 101 --lockstep 8 > gpurun_out/r2k_stream_lock8.json 2> gpurun_out/r2k_stream_lock8.err
timeout 300 python bench.py --impl reference --workload stream --steps 8 > gpurun_out/r2k_stream_ref.json 2> gpurun_out/r2k_stream_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2k_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-traffic --no-m2 --no-e2e-m1 > gpurun_out/r2k_launches_bench.log 2>&1
HEAD_BENCH_ITERS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 264 -c 88 --csv --log-file gpurun_out/r2k_launches_fused.csv \
    python scripts/tune/head_bench.py 256/512 64 8 > gpurun_out/r2k_launches_fused.log 2>&1
grep -E "passed|failed|exit" gpurun_out/r2k_pytest.log | tail -3; tail -2 gpurun_out/r2k_smoke.log; cut -c1-200 gpurun_out/r2k_bench.json; cut -c1-200 gpurun_out/r2k_stream.json
