#!/bin/bash
# read back an ncu capture made by gpu_cg_cycle.sh: per-line stall samples + headline counters
tag=$1; n=${2:-24}
cd /root/repo/gpurun_out
ncu -i prof_cg_$tag.ncu-rep --page source --csv 2>/dev/null > cg_${tag}_src.csv
ncu -i prof_cg_$tag.ncu-rep --page raw --csv 2>/dev/null > cg_${tag}_raw.csv
cd ..
(cd poissonrecon_gpu_b200/csrc/build && cuobjdump -xelf all solver.o > /dev/null)
python tools/ncu_lines.py gpurun_out/cg_${tag}_src.csv poissonrecon_gpu_b200/csrc/build/solver.sm_100a.cubin _ZN3prb15k_cg_all_depthsILb0EEEvNS_8CgParamsE $n
python - <<P
import csv
r=list(csv.reader(open('/root/repo/gpurun_out/cg_${tag}_raw.csv')))
h=r[0]; u=r[1]; v=r[2]
for n in ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','lts__t_sector_hit_rate.pct','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active']:
    i=h.index(n); print(n,u[i],v[i])
P
