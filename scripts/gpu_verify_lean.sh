#!/bin/bash
# last verification pass of round 2 (lean): all GPU tests, smoke, default bench line, stream (one sequence at a time / lock-step 8)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=lean
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 600 python bench.py --workload stream --sequences 2 --frames 101 --no-cpu > gpurun_out/${T}_stream.json 2> gpurun_out/${T}_stream.err
timeout 600 python bench.py --workload stream --sequences 8 --frames 101 --lockstep 8 --no-cpu > gpurun_out/${T}_stream_lock8.json 2> gpurun_out/${T}_stream_lock8.err
timeout 300 python scripts/tune/m2_profile.py 127/255 1 > gpurun_out/${T}_m2_b1.log 2>&1
timeout 300 python scripts/tune/backbone_bench.py > gpurun_out/${T}_backbone.log 2>&1
grep -E "passed|failed|exit" gpurun_out/${T}_pytest.log | tail -3; tail -2 gpurun_out/${T}_smoke.log; cut -c1-200 gpurun_out/${T}_bench.json; cut -c70-130 gpurun_out/${T}_stream.json; cut -c70-130 gpurun_out/${T}_stream_lock8.json; tail -2 gpurun_out/${T}_backbone.log
