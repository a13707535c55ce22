#!/bin/bash
# usage: scratch/run_mgpu.sh <tag> <gpus...>   (bench at each listed GPU count, N=30)
tag=$1; shift
mkdir -p gpurun_out
for g in "$@"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2953$g \
      bench.py --gpus $g --no-cpu-baseline --no-e2e --no-tdvp $EXTRA 2> gpurun_out/bench_${tag}_g${g}.err | grep "^{" > gpurun_out/bench_${tag}_g${g}.json
  grep -v "OMP_NUM\|^\*\*\*\*" gpurun_out/bench_${tag}_g${g}.err | tail -5
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${tag}_g${g}.json"))
    r = d["roofline"]
    print("$tag g=$g", "steps/s", round(d["value"], 4), "ms/step", round(d["ms_per_step"], 2), "frac", round(r["frac"], 3),
          "terms", d["config"]["chebyshev_terms"], "nvlink GB/s", round(r["nvlink_read_gbs_per_gpu"], 1),
          "ms by pass", [round(x, 3) for x in r.get("avg_launch_ms_by_pass", [])])
except Exception as e:
    print("$tag g=$g FAILED", e)
PY
done
