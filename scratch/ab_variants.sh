#!/bin/bash
# A/B timing of library builds (scratch/variants/*.so) on ONE box: bench exact leg only, two rounds interleaved.
mkdir -p gpurun_out
for v in "$@"; do
  echo "== tests/test_exact_gpu.py with $v"; QCA_B200_LIBRARY=$PWD/scratch/variants/libqca_$v.so timeout 400 python -m pytest tests/test_exact_gpu.py -m gpu -x -q 2>&1 | tail -3
done
for round in 1 2; do
for v in "$@"; do
  QCA_B200_LIBRARY=$PWD/scratch/variants/libqca_$v.so timeout 200 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-matched --no-tdvp \
      2> gpurun_out/ab_$v.err > gpurun_out/ab_${v}_$round.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_${v}_$round.json"))
    r = d["roofline"]
    print("$v round $round: steps/s", round(d["value"], 4), "frac", round(r["frac"], 4), "ms by pass", [round(x, 3) for x in r["avg_launch_ms_by_pass"]],
          "checksum", d["checksum"]["ok"], d["checksum"]["max_abs_diff_vs_committed"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("$v FAILED", e)
PY
done
done
# TDVP: parity tests and the chi = 256 step breakdown with the in-tree build
timeout 600 python -m pytest tests/test_tdvp_gpu.py -m gpu -q 2>&1 | tail -8
timeout 200 python scratch/tdvp_prof2.py > gpurun_out/r2e_tdvp_prof.txt 2>&1; head -14 gpurun_out/r2e_tdvp_prof.txt; tail -1 gpurun_out/r2e_tdvp_prof.txt
QCA_TDVP_SYNC_SPLIT=1 timeout 200 python scratch/tdvp_prof2.py 2>&1 | head -1
