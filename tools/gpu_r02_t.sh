#!/bin/bash
# 8 GPUs: final depth-10 bench line, then BASELINE config 5 (100 M points, depth 12)
mkdir -p gpurun_out/r02t
export PRB_ARENA_GB=28
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02t/bench_d10_8gpu.log 2>&1
tail -1 gpurun_out/r02t/bench_d10_8gpu.log | cut -c1-260
timeout 420 $TR --master-port 29612 tools/mg_phases.py dense100m_d12 > gpurun_out/r02t/phases_d12_8gpu.log 2>&1
grep "rank 0/8\]\|rank 7/8\]" gpurun_out/r02t/phases_d12_8gpu.log | tail -4 | cut -c1-1700; tail -3 gpurun_out/r02t/phases_d12_8gpu.log | cut -c1-400
