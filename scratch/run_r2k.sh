#!/bin/bash
# Final 1-GPU verification of round 2: GPU parity suite, smoke, the default bench line, the ncu launch list of one timed step
# and one --set full capture per tile pass (second Chebyshev term: pass 0 with both recurrence operands).
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2k_pytest.log; tail -4 gpurun_out/r2k_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2k_smoke.log
echo "tests+smoke after $(( $(date +%s) - T0 )) s"
timeout 500 python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2k_bench.json"))
print("value", d["value"], "frac", d["roofline"]["frac"], d["roofline"]["avg_launch_ms_by_pass"], "e2e", d["e2e"]["value"], d["e2e"].get("sequential"), "checksum", d["checksum"]["ok"],
      "tdvp", d["tdvp"]["value"] if d.get("tdvp") and "value" in d["tdvp"] else d.get("tdvp"), "matched", [round(m["ratio"], 1) for m in d.get("matched", [])], d["clocks"])
PY
echo "bench after $(( $(date +%s) - T0 )) s"
QCA_NCU_RANGE=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_v3_final_n30.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-tdvp --no-matched > gpurun_out/r2k_ncu_list.log 2>&1; tail -1 gpurun_out/r2k_ncu_list.log | cut -c1-200
echo "ncu list after $(( $(date +%s) - T0 )) s"
QCA_NCU_RANGE=1 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:pass_kernel_v3 --launch-skip 3 --launch-count 3 -f -o gpurun_out/r02_pass_v3_final_n30 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-tdvp --no-matched > gpurun_out/r2k_ncu_full.log 2>&1; tail -1 gpurun_out/r2k_ncu_full.log | cut -c1-200
echo "all done after $(( $(date +%s) - T0 )) s"
