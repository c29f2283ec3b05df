#!/bin/bash
# round 2, call AG: smoke + short bench line on the final build
mkdir -p gpurun_out/r02ag
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ag/smoke.log 2>&1
timeout 60 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02ag/bench.log 2>&1
tail -1 gpurun_out/r02ag/smoke.log; grep '^{"metric"' gpurun_out/r02ag/bench.log | cut -c1-250
