#!/bin/bash
# adapter micro-benchmark + ncu --set full of the token-reduction weight-gradient kernel
cd /root/repo; mkdir -p gpurun_out
timeout 300 python tools/adapter_bench.py 256 > gpurun_out/r02_adapter_bench.txt 2>&1
timeout 300 python tools/adapter_bench.py 32 >> gpurun_out/r02_adapter_bench.txt 2>&1
cat gpurun_out/r02_adapter_bench.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tokgemm_tc --launch-skip 6 -c 2 -o gpurun_out/r02_ncu_tokgemm python tools/adapter_bench.py 256 > gpurun_out/r02_ncu_tokgemm.log 2>&1
ncu -i gpurun_out/r02_ncu_tokgemm.ncu-rep --page details --csv 2>/dev/null | grep -i "DRAM Throughput\|Duration\|Memory Throughput\|Compute (SM) Throughput\|Registers Per\|Achieved Occupancy\|L2 Hit\|tensor\|Executed Ipc Active" | head -40 > gpurun_out/r02_ncu_tokgemm_summary.txt
ncu -i gpurun_out/r02_ncu_tokgemm.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]
want=['dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','sm__inst_executed_pipe_tensor','sm__pipe_tensor','lts__t_sector_hit_rate.pct','dram__throughput.avg.pct_of_peak_sustained_elapsed']
idx=[i for i,h in enumerate(hdr) if any(w in h for w in want)]
for r in rows[:4]:
    print([ (hdr[i], r[i]) for i in idx])
" >> gpurun_out/r02_ncu_tokgemm_summary.txt 2>&1
cat gpurun_out/r02_ncu_tokgemm_summary.txt | cut -c 1-1500
