#!/bin/bash
mkdir -p gpurun_out/r02l
export PRB_ARENA_BYTES=$((6<<30))
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 tools/mg_check.py torus1m_d9 2 > gpurun_out/r02l/mg_check_torus.log 2>&1
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29582 tools/mg_check.py scan5m_d10 2 > gpurun_out/r02l/mg_check_scan.log 2>&1
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29583 tools/mg_phases.py scan5m_d10 > gpurun_out/r02l/phases_2gpu.log 2>&1
grep -h "MG_CHECK" gpurun_out/r02l/mg_check_*.log; grep -h "rank 0/2" gpurun_out/r02l/phases_2gpu.log | tail -2 | cut -c1-1500
