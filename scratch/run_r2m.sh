#!/bin/bash
# Large registers on ONE B200: N=31 (32-bit amplitude index at its limit, 3 x 16 GiB planes) and N=32 (64-bit index
# instantiations of pass_kernel_v3, 3 x 32 GiB planes + the parked upload plane = 128 GiB of the 180 GB)
mkdir -p gpurun_out
for n in 31 32; do
  timeout 200 python bench.py --num-cells $n --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-matched --no-tdvp 2> gpurun_out/r2m_n$n.err > gpurun_out/r2m_n$n.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2m_n$n.json")); r = d["roofline"]; c = d["checksum"]; p = c["population"]; s = c["entropy"]
    n = len(p)
    print("N=$n steps/s", round(d["value"], 4), "frac", round(r["frac"], 4), "ms by pass", [round(x, 3) for x in r["avg_launch_ms_by_pass"]], "norm2", c["norm2"],
          "mirror pop", max(abs(p[i] - p[n - 1 - i]) for i in range(n)), "mirror S", max(abs(s[i] - s[n - 1 - i]) for i in range(n)), "terms", d["details"]["chebyshev_terms"], "passes", d["details"]["passes_per_term"])
except Exception as e:
    print("N=$n FAILED", e)
    print(open("gpurun_out/r2m_n$n.err").read()[-1500:])
PY
done
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
