#!/bin/bash
# GPU pass for the transform-domain correlation kernel: parity tests, A/B bench (vs the direct kernels), ncu --set full.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "xcorr or fft" > gpurun_out/pytest_fft.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_fft.log
tail -5 gpurun_out/pytest_fft.log
timeout 300 python bench.py --no-cpu --no-e2e --xcorr-algo direct > gpurun_out/bench_direct.json 2> gpurun_out/bench_fft.err
timeout 300 python bench.py --no-cpu --no-e2e > gpurun_out/bench_fft.json 2>> gpurun_out/bench_fft.err
timeout 300 python bench.py --no-cpu --no-e2e --workload win15 --xcorr-algo fft > gpurun_out/bench_fft_win15.json 2>> gpurun_out/bench_fft.err
python - <<'PY'
import json
for f in ("bench_direct", "bench_fft", "bench_fft_win15"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value %.0f" % d["value"], "ms/step %.3f" % d["ms_per_step"], {k: round(v, 3) for k, v in d["roofline"]["kernel_ms"].items()})
    except Exception as e:
        print(f, "FAILED", e)
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:xcorr_fft -c 2 -o gpurun_out/prof_xcorr_fft -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/prof_xcorr_fft.log 2>&1
tail -3 gpurun_out/bench_fft.err
