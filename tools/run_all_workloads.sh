#!/bin/bash
# Run bench.py on every BASELINE config shape (single GPU) and collect the JSON lines under gpurun_out/.
for w in cfg1 cfg2 cfg3 cfg4 cfg5; do
  python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-chains 2>/dev/null | tail -1 > gpurun_out/bench_$w.json
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$w.json"))
print("$w", d["config"]["workload"], "| value %.3e" % d["value"], "| e2e %.3e" % d["e2e"]["value"], "| roofline %.3f" % d["roofline"]["frac"], "| launches", d["gpu_launches"])
PY
done
