#!/bin/bash
# 8 GPUs: correctness + phases + bench after the barrier / plan / chunking changes
mkdir -p gpurun_out/r02j
export PRB_ARENA_BYTES=$((6<<30))
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 tools/mg_check.py scan5m_d10 2 > gpurun_out/r02j/mg_check_scan5m_d10.log 2>&1
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29572 tools/mg_phases.py scan5m_d10 > gpurun_out/r02j/phases_8gpu.log 2>&1
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02j/bench_8gpu.log 2>&1
grep -E "MG_CHECK" gpurun_out/r02j/mg_check_scan5m_d10.log | tail -1; grep "timeline" gpurun_out/r02j/phases_8gpu.log | tail -2 | cut -c1-1600; grep -o "CG ms[^i]*" gpurun_out/r02j/phases_8gpu.log | sort | uniq | head -8; tail -1 gpurun_out/r02j/bench_8gpu.log | cut -c1-300
