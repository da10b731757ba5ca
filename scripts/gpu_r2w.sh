#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r2w
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log | cut -c1-250
timeout 600 python bench.py --no-cpu --no-traffic --no-e2e-m1 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2w_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "fused", d["fused"]["value"], "m2", d["m2"]["value"], d["m2"]["tflops"], d["m2"]["stage_ms"])
PY
timeout 600 python bench.py --workload stream --sequences 8 --frames 101 --lockstep 8 --no-cpu > gpurun_out/${T}_stream_lock8.json 2> gpurun_out/${T}_stream_lock8.err; cut -c1-200 gpurun_out/${T}_stream_lock8.json
timeout 600 python bench.py --workload stream --sequences 2 --frames 101 --no-cpu > gpurun_out/${T}_stream.json 2> gpurun_out/${T}_stream.err; cut -c1-200 gpurun_out/${T}_stream.json
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_conv.py -m gpu -q -k "tensor_memory and (case0 or case1 or case4)" > gpurun_out/${T}_racecheck.log 2>&1; tail -3 gpurun_out/${T}_racecheck.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_conv.py -m gpu -q -k "tensor_memory and not case7 and not case2" > gpurun_out/${T}_memcheck.log 2>&1; tail -3 gpurun_out/${T}_memcheck.log
