#!/bin/bash
mkdir -p gpurun_out/r02b
python tools/div_check.py sphere3k_d5 sphere2k_d2 sphere2k_d3 sphere100k_d8 torus1m_d9 scan5m_d10 > gpurun_out/r02b/div_check.log 2>&1
( time python -m pytest tests/test_parity_gpu.py -m gpu -x -q ) > gpurun_out/r02b/pytest_parity.log 2>&1
python tools/step_times.py scan5m_d10 > gpurun_out/r02b/step_times.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02b/launches.csv python tools/step_times.py scan5m_d10 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02b/launches.csv 45 > gpurun_out/r02b/launch_summary.txt 2>&1
tail -12 gpurun_out/r02b/div_check.log; tail -5 gpurun_out/r02b/pytest_parity.log; tail -3 gpurun_out/r02b/step_times.log
