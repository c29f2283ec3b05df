#!/bin/bash
# round 2, call Z: full GPU suite, smoke, final bench line, launch list, --set full of the four secondary kernels
mkdir -p gpurun_out/r02z
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r02z/pytest_gpu.log 2>&1
( time timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02z/smoke.log 2>&1
( time timeout 400 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02z/bench.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02z/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-cuda > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02z/launches.csv 60 > gpurun_out/r02z/launch_summary.txt 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k 'regex:k_splat$|k_div_fine|k_point_values|k_vertex_values_stream' -s 4 -c 4 -o gpurun_out/r02z/prof_secondary -f python tools/step_times.py scan5m_d10 > gpurun_out/r02z/ncu_secondary.log 2>&1
tail -4 gpurun_out/r02z/pytest_gpu.log; grep -E "^FAILED" gpurun_out/r02z/pytest_gpu.log | head; tail -4 gpurun_out/r02z/smoke.log; grep '^{"metric"' gpurun_out/r02z/bench.log | cut -c1-300; head -3 gpurun_out/r02z/launch_summary.txt; tail -2 gpurun_out/r02z/ncu_secondary.log
