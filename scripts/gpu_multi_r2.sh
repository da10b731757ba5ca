#!/bin/bash
# N-GPU pass of round 2 (run under `gpurun --gpus N`): scaling of the default bench line, config 4 (stream) and config 5 (win15) at N
# ranks, per-rank H2D bandwidth with all ranks copying, the 2-rank NCCL equivalence test.
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/m_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/m_pytest_multi.log 2>&1; echo "exit $?" >> gpurun_out/m_pytest_multi.log
for n in 2 4 8; do
  [ $n -le $N ] || continue
  timeout 200 $TR --nproc-per-node $n --master-port 2951$n scripts/tune/h2d_ranks.py >> gpurun_out/m_h2d_ranks.jsonl 2>> gpurun_out/m_h2d_ranks.err
  timeout 600 $TR --nproc-per-node $n --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 3 --no-e2e-m1 > gpurun_out/m_bench_n$n.json 2> gpurun_out/m_bench_n$n.err
done
timeout 900 $TR --nproc-per-node $N --master-port 29531 bench.py --gpus $N --workload stream --sequences 8 --frames 501 > gpurun_out/m_stream_n$N.json 2> gpurun_out/m_stream_n$N.err
timeout 300 $TR --nproc-per-node $N --master-port 29532 bench.py --gpus $N --workload win15 --batch 256 --no-cpu > gpurun_out/m_win15_n$N.json 2> gpurun_out/m_win15_n$N.err
timeout 600 python bench.py --workload stream --sequences 8 --frames 501 --lockstep 8 > gpurun_out/m_stream_1gpu_lock8.json 2> gpurun_out/m_stream_1gpu_lock8.err
tail -2 gpurun_out/m_pytest_multi.log; cat gpurun_out/m_h2d_ranks.jsonl; for n in 2 4 8; do [ -f gpurun_out/m_bench_n$n.json ] && cut -c1-160 gpurun_out/m_bench_n$n.json; done; cut -c1-300 gpurun_out/m_stream_n$N.json
