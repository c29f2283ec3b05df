#!/bin/bash
mkdir -p gpurun_out/r02r
( time timeout 400 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "small_configs or depth10 or edge_cases or full_size_properties" ) > gpurun_out/r02r/pytest_quick.log 2>&1
timeout 100 python tools/mg_phases.py scan5m_d10 > gpurun_out/r02r/phases_1gpu.log 2>&1
( time timeout 500 python tools/quick_bench.py dense100m_d12 - 2 ) > gpurun_out/r02r/dense100m_d12_1gpu.log 2>&1
tail -3 gpurun_out/r02r/pytest_quick.log; grep -h "stages" gpurun_out/r02r/phases_1gpu.log | tail -1 | cut -c1-300; tail -6 gpurun_out/r02r/dense100m_d12_1gpu.log | cut -c1-900
