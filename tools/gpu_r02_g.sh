#!/bin/bash
mkdir -p gpurun_out/r02g
timeout 120 python tools/mg_phases.py scan5m_d10 > gpurun_out/r02g/phases_1gpu.log 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tools/mg_phases.py scan5m_d10 > gpurun_out/r02g/phases_2gpu.log 2>&1
( time timeout 400 python -m pytest tests/test_cli_gpu.py tests/test_multi_gpu.py -m gpu -q -x ) > gpurun_out/r02g/pytest_cli_mg.log 2>&1
tail -2 gpurun_out/r02g/phases_1gpu.log | cut -c1-600; tail -4 gpurun_out/r02g/phases_2gpu.log | cut -c1-600; tail -3 gpurun_out/r02g/pytest_cli_mg.log
