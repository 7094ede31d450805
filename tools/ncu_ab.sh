#!/bin/bash
# tools/ncu_ab.sh OUTPREFIX LIB [model N T B]: one `ncu --set full` capture of the fused filter kernel of a library build and the
# summary metrics / hot source lines of it (run under gpurun; files land in gpurun_out/)
OUT=$1; LIB=$2; shift 2
ARGS=${@:-levy_sv 4096 50 444}
export CPIMH_LIB=$(realpath $LIB)
ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-pf_batch_kernel} -c 1 -f -o gpurun_out/$OUT python ${PROF_SCRIPT:-tools/prof_pf.py} $ARGS > gpurun_out/$OUT.ncu.log 2>&1
ncu -i gpurun_out/$OUT.ncu-rep --page raw --csv > gpurun_out/$OUT.raw.csv 2>/dev/null
ncu -i gpurun_out/$OUT.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/$OUT.src.csv 2>/dev/null
python tools/ncu_hot.py gpurun_out/$OUT.src.csv 70 > gpurun_out/$OUT.hot.txt 2>&1
python - <<PY > gpurun_out/$OUT.summary.txt
import csv
rows=list(csv.reader(open("gpurun_out/$OUT.raw.csv")))
hdr,units,vals=rows[0],rows[1],rows[2]
want=["gpu__time_duration.sum","launch__registers_per_thread","launch__block_size","launch__grid_size","smsp__inst_executed.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__thread_inst_executed_per_inst_executed.ratio","dram__bytes_read.sum","dram__bytes_write.sum","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","smsp__inst_executed_op_local_ld.sum","smsp__inst_executed_op_local_st.sum","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem"]
for w in want:
    if w in hdr:
        i=hdr.index(w); print("%s [%s] = %s"%(w,units[i],vals[i]))
st=[(float(vals[i] or 0),h) for i,h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued")]
tot=sum(v for v,_ in st)
for v,h in sorted(st,reverse=True)[:12]: print("  %5.1f%% %s"%(100*v/max(tot,1),h))
PY
rm -f gpurun_out/$OUT.src.csv gpurun_out/$OUT.raw.csv
