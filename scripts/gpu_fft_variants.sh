#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2o_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2o_pytest.log
tail -3 gpurun_out/r2o_pytest.log
for algo in fft_phased fft_pipe fft_ws; do
  timeout 300 python bench.py --no-cpu --no-e2e --no-m2 --no-traffic --no-e2e-m1 --steps 10 --xcorr-algo $algo > gpurun_out/r2o_bench_$algo.json 2> gpurun_out/r2o_bench_$algo.err
  timeout 300 python bench.py --no-cpu --no-e2e --no-m2 --no-traffic --no-e2e-m1 --steps 10 --workload win15 --xcorr-algo $algo > gpurun_out/r2o_w15_$algo.json 2>> gpurun_out/r2o_bench_$algo.err
  python - $algo <<'PY'
import json, sys
a = sys.argv[1]
for tag in ("bench", "w15"):
    try:
        d = json.load(open("gpurun_out/r2o_%s_%s.json" % (tag, a)))
        print(a, tag, "value %.0f" % d["value"], d["roofline"].get("kernel_ms"))
    except Exception as e:
        print(a, tag, "FAILED", e)
PY
done
cat > /tmp/race.py <<'PY'
import torch
from hdn_b200 import ops
for algo in ("fft_pipe", "fft_ws"):
    ops.set_xcorr_algo(algo)
    for shape, circ in (((2, 8, 61, 61), False), ((2, 8, 29, 29), True)):
        x = torch.randn(shape, device="cuda"); k = torch.randn((2, 8, 29, 29), device="cuda")
        y = (ops.xcorr_depthwise_circular if circ else ops.xcorr_depthwise)(x, k)
        torch.cuda.synchronize(); print(algo, shape, float(y.abs().max()))
PY
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python /tmp/race.py > gpurun_out/r2o_racecheck.log 2>&1; tail -8 gpurun_out/r2o_racecheck.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:xcorr_fft_ws -s 4 -c 2 -o gpurun_out/r2o_prof_ws -f \
    python bench.py --no-cpu --no-e2e --no-m2 --no-traffic --no-e2e-m1 --steps 2 --warmup 1 --xcorr-algo fft_ws > gpurun_out/r2o_prof_ws.log 2>&1
ls -la gpurun_out/r2o_prof_ws.ncu-rep
