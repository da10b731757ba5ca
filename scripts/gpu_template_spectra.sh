#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r2r
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_conv.py -m gpu -q -x -k "spectra or head_engine or fft_kernel_variants" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -15 gpurun_out/${T}_pytest.log
for algo in auto fft_phased fft_pipe; do
  timeout 300 python bench.py --no-cpu --no-e2e --no-m2 --no-traffic --no-e2e-m1 --steps 10 --shared-template 1 --xcorr-algo $algo > gpurun_out/${T}_bench_shared_$algo.json 2> gpurun_out/${T}_bench_shared_$algo.err
  python - $algo <<'PY'
import json, sys
a = sys.argv[1]
try:
    d = json.load(open("gpurun_out/r2r_bench_shared_%s.json" % a))
    print(a, "value %.0f" % d["value"], d["roofline"].get("kernel_ms"), d["config"].get("template"))
except Exception as e:
    print(a, "FAILED", e)
PY
done
HDN_B200_TEMPLATE_SPECTRA=0 timeout 300 python bench.py --no-cpu --no-e2e --no-m2 --no-traffic --no-e2e-m1 --steps 10 --shared-template 1 > gpurun_out/${T}_bench_shared_nospec.json 2> gpurun_out/${T}_bench_shared_nospec.err
python -c "
import json; d=json.load(open('gpurun_out/r2r_bench_shared_nospec.json')); print('nospec value %.0f' % d['value'], d['roofline'].get('kernel_ms'))"
timeout 300 python scripts/tune/head_bench.py 256/512 64 8 > gpurun_out/${T}_head_bench.log 2>&1; tail -2 gpurun_out/${T}_head_bench.log | cut -c1-200
cat > /tmp/race.py <<'PY'
import torch
from hdn_b200 import ops
for algo in ("fft_phased", "fft_pipe"):
    ops.set_xcorr_algo(algo)
    for shape, circ in (((3, 8, 61, 61), False), ((3, 8, 29, 29), True)):
        x = torch.randn(shape, device="cuda"); k = torch.randn((1, 8, 29, 29), device="cuda")
        sp = ops.xcorr_template_spectra([k], shape[2], shape[3], circ)
        y = ops.xcorr_depthwise_multi_spec([x], sp, 29, 29, circ)[0]
        torch.cuda.synchronize(); print(algo, shape, float(y.abs().max()))
PY
PYTHONPATH=$PWD timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python /tmp/race.py > gpurun_out/${T}_racecheck.log 2>&1; tail -6 gpurun_out/${T}_racecheck.log
