#!/bin/bash
# lean 8-GPU pass with the final kernels (run under `gpurun --gpus 8`): 2-rank NCCL equivalence test, the default bench line, config 4
# (stream) and config 5 (win15) at 8 ranks
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/m2_pytest_multi.log 2>&1; echo "exit $?" >> gpurun_out/m2_pytest_multi.log
timeout 600 $TR --nproc-per-node $N --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 --no-e2e-m1 --no-cpu --no-m2 --no-traffic > gpurun_out/m2_bench_n$N.json 2> gpurun_out/m2_bench_n$N.err
timeout 900 $TR --nproc-per-node $N --master-port 29531 bench.py --gpus $N --workload stream --sequences 8 --frames 501 > gpurun_out/m2_stream_n$N.json 2> gpurun_out/m2_stream_n$N.err
timeout 300 $TR --nproc-per-node $N --master-port 29532 bench.py --gpus $N --workload win15 --batch 256 --no-cpu > gpurun_out/m2_win15_n$N.json 2> gpurun_out/m2_win15_n$N.err
tail -2 gpurun_out/m2_pytest_multi.log; cut -c1-200 gpurun_out/m2_bench_n$N.json; cut -c1-300 gpurun_out/m2_stream_n$N.json; cut -c1-200 gpurun_out/m2_win15_n$N.json
