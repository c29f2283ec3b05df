#!/bin/bash
mkdir -p gpurun_out/r02n
( time timeout 400 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "small_configs or depth10 or config1 or edge_cases" ) > gpurun_out/r02n/pytest_quick.log 2>&1
timeout 100 python tools/mg_phases.py scan5m_d10 > gpurun_out/r02n/phases_1gpu.log 2>&1
timeout 200 python tools/mg_phases.py multi20m_d11 > gpurun_out/r02n/phases_d11_1gpu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02n/launches_d11.csv python tools/step_times.py multi20m_d11 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02n/launches_d11.csv 30 > gpurun_out/r02n/launch_summary_d11.txt 2>&1
tail -3 gpurun_out/r02n/pytest_quick.log; grep -h "timeline\|CG ms" gpurun_out/r02n/phases_1gpu.log | tail -2 | cut -c1-1800;  grep -h "timeline\|CG ms" gpurun_out/r02n/phases_d11_1gpu.log | tail -2 | cut -c1-1800; head -24 gpurun_out/r02n/launch_summary_d11.txt
