#!/bin/bash
# per-pass launch times of the cluster kernels for different cluster sizes / strided-tile geometries (N=30, one GPU)
mkdir -p gpurun_out
for cfg in "3 4" "0 4" "1 4" "2 4" "3 7"; do
  set -- $cfg
  QCA_V3_CLUSTER_BITS=$1 QCA_V3_MIN_LOW=$2 python bench.py --steps 2 --warmup 3 --no-tdvp --no-cpu-baseline --no-matched --no-e2e ${N:+--num-cells $N} > gpurun_out/v3_sweep_cb$1_low$2.json 2>/dev/null
  python - <<PY
import json
d = json.load(open("gpurun_out/v3_sweep_cb$1_low$2.json"))
r = d["roofline"]
print("cb=$1 minlow=$2", "steps/s", round(d["value"], 4), "passes", d["details"]["passes_per_term"], "ms by pass", [round(x, 3) for x in r["avg_launch_ms_by_pass"]],
      "GB/s", round(r["achieved"]), "checksum", d["checksum"]["ok"])
PY
done
