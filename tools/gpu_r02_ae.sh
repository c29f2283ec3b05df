#!/bin/bash
# round 2, call AE: ncu launch list of BASELINE config 4 (20 M points, depth 11) on the final build
mkdir -p gpurun_out/r02ae
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02ae/launches_d11.csv python tools/d11_launches.py > gpurun_out/r02ae/run.log 2>&1
python tools/launch_summary.py gpurun_out/r02ae/launches_d11.csv 45 > gpurun_out/r02ae/launch_summary_d11.txt 2>&1
head -30 gpurun_out/r02ae/launch_summary_d11.txt
