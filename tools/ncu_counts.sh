#!/bin/bash
# tools/ncu_counts.sh -- per-workload warp-instruction counts and DRAM traffic of the filter kernels (one short run of each BASELINE
# config shape under `ncu --metrics ...`), written as profiles-ready JSON into gpurun_out/.  Run under gpurun.
MET=smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
run() {  # name script args... ; particle-steps of ONE run() call = $PS
  local name=$1; shift
  ncu --metrics $MET --clock-control none --csv --log-file gpurun_out/counts_$name.csv "$@" > gpurun_out/counts_$name.log 2>&1
}
run cfg1 python tools/prof_pf.py gss 256 100 3552
run cfg2 python tools/prof_pf.py levy_sv 4096 50 444
run cfg3 python tools/prof_pf.py sk 8192 50 148
run cfg4 python tools/prof_pf.py ar10 1024 100 740
run cfg5 python tools/prof_wide.py 20 1
python - <<'PY'
import csv, json, collections
shapes = {"cfg1": 256 * 100 * 3552, "cfg2": 4096 * 50 * 444, "cfg3": 8192 * 50 * 148, "cfg4": 1024 * 100 * 740, "cfg5": 65536 * 20 * 1}
runs = 3   # prof_pf.py / prof_wide.py run the batch three times
issue, traffic, detail = {}, {}, {}
for name, ps in shapes.items():
    rows = [r for r in csv.reader(open("gpurun_out/counts_%s.csv" % name)) if len(r) > 10]
    hdr = rows[0]; ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    tot = collections.defaultdict(float); per_kernel = collections.defaultdict(lambda: collections.defaultdict(float))
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        k = r[ki].split("(")[0].split("<")[0]
        if "pf_batch_kernel" not in r[ki] and "wide_" not in r[ki]:
            continue
        if "wide_init" in r[ki] or "wide_path" in r[ki]:
            per_kernel[k][r[mi]] += v
            continue
        tot[r[mi]] += v; per_kernel[k][r[mi]] += v
    n = ps * runs
    issue[name] = {"warp_instructions_per_particle_step": tot["smsp__inst_executed.sum"] / n,
                   "source": "profiles/r2_counts.json (ncu --metrics smsp__inst_executed.sum, tools/ncu_counts.sh: %s)" % name}
    traffic[name] = {"dram_bytes_per_particle_step": (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / n,
                     "source": "profiles/r2_counts.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, tools/ncu_counts.sh: %s)" % name}
    detail[name] = {"particle_steps": n, "kernels": {k: dict(v) for k, v in per_kernel.items()}}
json.dump(issue, open("gpurun_out/issue_counts.json", "w"), indent=1)
json.dump(traffic, open("gpurun_out/ncu_traffic.json", "w"), indent=1)
json.dump(detail, open("gpurun_out/r2_counts.json", "w"), indent=1)
print(json.dumps(issue, indent=1)); print(json.dumps(traffic, indent=1))
PY
