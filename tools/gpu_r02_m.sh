#!/bin/bash
mkdir -p gpurun_out/r02m
timeout 300 python tools/mg_phases.py multi20m_d11 > gpurun_out/r02m/phases_d11_1gpu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_vertex_values_stream -s 2 -c 1 -o gpurun_out/r02m/prof_vv2 -f python tools/step_times.py scan5m_d10 > gpurun_out/r02m/ncu_vv.log 2>&1
grep -h "timeline\|CG ms" gpurun_out/r02m/phases_d11_1gpu.log | tail -2 | cut -c1-2000; tail -2 gpurun_out/r02m/ncu_vv.log
