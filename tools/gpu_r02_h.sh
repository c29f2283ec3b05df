#!/bin/bash
# 8 GPUs, lean: correctness against 1 GPU, CG phase split, bench line
mkdir -p gpurun_out/r02h
export PRB_ARENA_BYTES=$((6<<30))
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/mg_check.py scan5m_d10 2 > gpurun_out/r02h/mg_check_scan5m_d10.log 2>&1
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 tools/mg_phases.py scan5m_d10 > gpurun_out/r02h/phases_8gpu.log 2>&1
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02h/bench_8gpu.log 2>&1
grep -E "MG_CHECK|rank 0\]" gpurun_out/r02h/mg_check_scan5m_d10.log | tail -3 | cut -c1-500; grep "rank 0/8\|rank 7/8\|rank 3/8" gpurun_out/r02h/phases_8gpu.log | tail -3 | cut -c1-500; tail -1 gpurun_out/r02h/bench_8gpu.log | cut -c1-300
