#!/bin/bash
# round 2 final verification pass: all GPU tests, smoke, full default bench line, reference arm, stream, win15, ncu launch lists + summaries
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=verify
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
timeout 600 python bench.py --workload stream --sequences 2 --frames 101 > gpurun_out/${T}_stream.json 2> gpurun_out/${T}_stream.err
timeout 600 python bench.py --workload stream --sequences 8 --frames 101 --lockstep 8 > gpurun_out/${T}_stream_lock8.json 2> gpurun_out/${T}_stream_lock8.err
timeout 300 python bench.py --impl reference --workload stream --steps 8 > gpurun_out/${T}_stream_ref.json 2> gpurun_out/${T}_stream_ref.err
timeout 300 python bench.py --workload win15 --no-cpu --no-e2e > gpurun_out/${T}_win15.json 2> gpurun_out/${T}_win15.err
timeout 300 python bench.py --workload 127/255 --batch 512 --no-cpu --no-m2 > gpurun_out/${T}_native.json 2> gpurun_out/${T}_native.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-traffic --no-m2 --no-e2e-m1 > gpurun_out/${T}_launches_bench.log 2>&1
HEAD_BENCH_ITERS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 264 -c 88 --csv --log-file gpurun_out/${T}_launches_fused.csv \
    python scripts/tune/head_bench.py 256/512 64 8 > gpurun_out/${T}_launches_fused.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:xcorr_fft -s 4 -c 2 -o gpurun_out/${T}_prof_xcorr -f \
    python bench.py --no-cpu --no-e2e --no-m2 --no-traffic --no-e2e-m1 --steps 2 --warmup 1 > gpurun_out/${T}_prof_xcorr.log 2>&1
timeout 300 python scripts/tune/m2_profile.py 127/255 1 > gpurun_out/${T}_m2_b1.log 2>&1
timeout 300 python scripts/tune/backbone_bench.py > gpurun_out/${T}_backbone.log 2>&1
grep -E "passed|failed|exit" gpurun_out/${T}_pytest.log | tail -3; tail -2 gpurun_out/${T}_smoke.log; cut -c1-300 gpurun_out/${T}_bench.json; cut -c1-200 gpurun_out/${T}_stream.json
