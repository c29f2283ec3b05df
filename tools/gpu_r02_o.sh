#!/bin/bash
mkdir -p gpurun_out/r02o
( time timeout 400 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "small_configs or depth10 or config1 or edge_cases" ) > gpurun_out/r02o/pytest_quick.log 2>&1
timeout 100 python tools/mg_phases.py scan5m_d10 > gpurun_out/r02o/phases_1gpu.log 2>&1
timeout 200 python tools/mg_phases.py multi20m_d11 > gpurun_out/r02o/phases_d11_1gpu.log 2>&1
tail -3 gpurun_out/r02o/pytest_quick.log; grep -h "timeline" gpurun_out/r02o/phases_1gpu.log | tail -1 | cut -c1-1800;  grep -h "timeline\|stages" gpurun_out/r02o/phases_d11_1gpu.log | tail -2 | cut -c1-1800
