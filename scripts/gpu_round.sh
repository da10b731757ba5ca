#!/bin/bash
# One GPU-box pass over the current tree: GPU parity tests, the bench line, the ncu launch list of the bench command,
# ncu --set full captures of the dominant kernels.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --workload 127/255 --batch 512 --no-cpu > gpurun_out/bench_native.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --workload win15 --no-cpu > gpurun_out/bench_win15.json 2>> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_256.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/launches_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:xcorr -c 2 -o gpurun_out/prof_xcorr_256 -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/prof_xcorr.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:xcorr -c 2 -o gpurun_out/prof_xcorr_256_direct -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --xcorr-algo direct > gpurun_out/prof_xcorr_direct.log 2>&1
timeout 300 python bench.py --no-cpu --no-e2e --xcorr-algo direct > gpurun_out/bench_direct.json 2>> gpurun_out/bench.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:xcorr -c 2 -o gpurun_out/prof_xcorr_native -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --workload 127/255 --batch 512 > gpurun_out/prof_xcorr_native.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -c 1 -s 2 -o gpurun_out/prof_conv_gemm -f \
    python scripts/tune/conv_one.py > gpurun_out/prof_conv.log 2>&1
timeout 300 python -m hdn_b200.runner --sequences 2 --frames 40 > gpurun_out/runner.json 2> gpurun_out/runner.err
timeout 300 python scripts/tune/tracker_profile.py > gpurun_out/tracker_profile.txt 2>&1
timeout 300 python scripts/tune/backbone_profile.py > gpurun_out/backbone_profile.txt 2>&1
HDN_B200_TCGEN05=0 timeout 300 python scripts/tune/backbone_profile.py > gpurun_out/backbone_profile_cudnn.txt 2>&1
timeout 300 python scripts/tune/backbone_bench.py > gpurun_out/backbone_bench.txt 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json
