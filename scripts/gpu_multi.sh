#!/bin/bash
# N-GPU pass (run under `gpurun --gpus N`): the bench line at N ranks, the tracker-level runner sharded over N GPUs, gloo-free NCCL path.
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 \
    > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 -m hdn_b200.runner --sequences $((4*N)) --frames 40 --lockstep 4 \
    > gpurun_out/runner_n$N.json 2> gpurun_out/runner_n$N.err
timeout 300 python -m hdn_b200.runner --sequences 8 --frames 40 --lockstep 8 > gpurun_out/runner_lockstep8.json 2> gpurun_out/runner_lockstep8.err
tail -n 2 gpurun_out/bench_n$N.json | cut -c1-400; tail -n 1 gpurun_out/runner_n$N.json; tail -n 1 gpurun_out/runner_lockstep8.json; tail -n 3 gpurun_out/bench_n$N.err; tail -n 3 gpurun_out/runner_n$N.err
