#!/bin/bash
mkdir -p gpurun_out/r02k
( time timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_blocks_gpu.py -m gpu -x -q -k "small_configs or edge_cases or blocks or radix or scan or depth10 or config1" ) > gpurun_out/r02k/pytest_quick.log 2>&1
timeout 100 python tools/mg_phases.py scan5m_d10 > gpurun_out/r02k/phases_1gpu.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02k/launches.csv python tools/step_times.py scan5m_d10 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02k/launches.csv 40 > gpurun_out/r02k/launch_summary.txt 2>&1
tail -3 gpurun_out/r02k/pytest_quick.log; grep -h "timeline\|CG ms" gpurun_out/r02k/phases_1gpu.log | tail -2 | cut -c1-1800; head -14 gpurun_out/r02k/launch_summary.txt
