#!/bin/bash
mkdir -p gpurun_out/r02c
python tools/div_check.py sphere100k_d8 scan5m_d10 > gpurun_out/r02c/div_check.log 2>&1
( time python -m pytest tests/test_parity_gpu.py -m gpu -x -q ) > gpurun_out/r02c/pytest_parity.log 2>&1
( time python -m pytest tests -m gpu -q --deselect tests/test_parity_gpu.py ) > gpurun_out/r02c/pytest_rest.log 2>&1
python tools/step_times.py scan5m_d10 > gpurun_out/r02c/step_times.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02c/launches.csv python tools/step_times.py scan5m_d10 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02c/launches.csv 50 > gpurun_out/r02c/launch_summary.txt 2>&1
tail -4 gpurun_out/r02c/div_check.log; tail -5 gpurun_out/r02c/pytest_parity.log; tail -3 gpurun_out/r02c/pytest_rest.log; tail -3 gpurun_out/r02c/step_times.log
