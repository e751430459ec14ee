#!/bin/bash
# ncu --set full of the LoRA reduction kernels inside the 256-image step
cd /root/repo; mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lora_fused_tc|lora_colsum_tc" --launch-skip 46 -c 6 -o gpurun_out/r02_ncu_lora python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --prof-steps 1 > gpurun_out/r02_ncu_lora.log 2>&1
ncu -i gpurun_out/r02_ncu_lora.ncu-rep --page raw --csv > gpurun_out/r02_ncu_lora_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02_ncu_lora_raw.csv')))
hdr=rows[0]
want=['Kernel Name','Grid Size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__cycles_active.avg','sm__cycles_elapsed.max','lts__t_sector_hit_rate.pct','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','launch__registers_per_thread']
idx=[(h,i) for i,h in enumerate(hdr) if h in want]
for r in rows[2:]:
    print({h:r[i] for h,i in idx})
PY
