#!/bin/bash
# second A/B of the tile-staging variants: which of (issue scheme, request order, ring depth) pays in which pass
mkdir -p gpurun_out
T0=$(date +%s)
echo "== tests/test_exact_gpu.py with p0first_du12"; QCA_B200_LIBRARY=$PWD/scratch/variants/libqca_p0first_du12.so timeout 300 python -m pytest tests/test_exact_gpu.py -m gpu -x -q 2>&1 | tail -2
for v in p0first_du8 allfirst_du12; do
echo "== subset with $v"; QCA_B200_LIBRARY=$PWD/scratch/variants/libqca_$v.so timeout 200 python -m pytest tests/test_exact_gpu.py -m gpu -x -q -k "fast_kernel or cluster or c_oracle_rows" 2>&1 | tail -2
done
echo "tests done after $(( $(date +%s) - T0 )) s"
one() {  # tag library env...
  tag=$1; lib=$2; shift 2
  env "$@" QCA_B200_LIBRARY=$lib timeout 150 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-matched --no-tdvp \
      2> gpurun_out/r2j_$tag.err > gpurun_out/r2j_${tag}_$round.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2j_${tag}_$round.json"))
    r = d["roofline"]
    print("$tag round $round: steps/s", round(d["value"], 4), "frac", round(r["frac"], 4), "ms by pass", [round(x, 3) for x in r["avg_launch_ms_by_pass"]],
          "checksum", d["checksum"]["ok"], d["checksum"]["max_abs_diff_vs_committed"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("$tag FAILED", e)
PY
}
for round in 1 2; do
  one base $PWD/quantum-cellular-automaton_b200/libqca_b200.so QCA_X=1
  one wi_du12 $PWD/scratch/variants/libqca_wi_du12.so QCA_X=1
  one p0first_du12 $PWD/scratch/variants/libqca_p0first_du12.so QCA_X=1
  one p0first_du8 $PWD/scratch/variants/libqca_p0first_du8.so QCA_X=1
  [ $round = 1 ] && one allfirst_du12 $PWD/scratch/variants/libqca_allfirst_du12.so QCA_X=1
done
echo "all done after $(( $(date +%s) - T0 )) s"
