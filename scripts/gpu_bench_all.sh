#!/bin/bash
mkdir -p gpurun_out
for W in "$@"; do
  timeout 900 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err
  echo "== $W exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$W.json"))
    print({k:d[k] for k in ("value","ms_per_step","lap_ms","cost_build_ms","total_cost")}, d["e2e"], d["roofline"]["frac"], d["lap_stats"], d.get("cpu_baseline"), d.get("strong_cfg5"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/bench_$W.err").read()[-2000:])
PY
done
