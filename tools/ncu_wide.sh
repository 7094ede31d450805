#!/bin/bash
# one `ncu --set full` capture of wide_propagate_kernel (config 5) -> gpurun_out/$1.{summary,hot}.txt
OUT=$1
ncu --set full --clock-control none --import-source on -k regex:wide_propagate --launch-skip 6 -c 1 -f -o gpurun_out/$OUT python tools/prof_wide.py 20 1 > gpurun_out/$OUT.ncu.log 2>&1
ncu -i gpurun_out/$OUT.ncu-rep --page raw --csv > gpurun_out/$OUT.raw.csv 2>/dev/null
ncu -i gpurun_out/$OUT.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/$OUT.src.csv 2>/dev/null
python tools/ncu_hot.py gpurun_out/$OUT.src.csv 45 > gpurun_out/$OUT.hot.txt 2>&1
python - <<PY > gpurun_out/$OUT.summary.txt
import csv
rows=list(csv.reader(open("gpurun_out/$OUT.raw.csv")))
hdr,units,vals=rows[0],rows[1],rows[2]
want=["gpu__time_duration.sum","launch__registers_per_thread","launch__block_size","launch__grid_size","smsp__inst_executed.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__thread_inst_executed_per_inst_executed.ratio","dram__bytes_read.sum","dram__bytes_write.sum","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active","lts__t_sector_hit_rate.pct","launch__occupancy_limit_registers","launch__waves_per_multiprocessor","sm__cycles_active.avg","smsp__cycles_active.avg"]
for w in want:
    if w in hdr:
        i=hdr.index(w); print("%s [%s] = %s"%(w,units[i],vals[i]))
st=[(float(vals[i] or 0),h) for i,h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued")]
tot=sum(v for v,_ in st)
for v,h in sorted(st,reverse=True)[:12]: print("  %5.1f%% %s"%(100*v/max(tot,1),h))
PY
rm -f gpurun_out/$OUT.src.csv gpurun_out/$OUT.raw.csv gpurun_out/$OUT.ncu-rep
