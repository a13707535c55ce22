#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2f_pytest.log; tail -4 gpurun_out/r2f_pytest.log
timeout 500 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2f_bench.json"))
print("value", d["value"], "frac", d["roofline"]["frac"], d["roofline"]["avg_launch_ms_by_pass"], "e2e", d["e2e"]["value"], d["e2e"].get("sequential"), "tdvp", d["tdvp"]["value"] if d.get("tdvp") and "value" in d["tdvp"] else d.get("tdvp"), d["tdvp"].get("split"))
PY
QCA_NCU_RANGE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:pass_kernel_v3 --launch-count 3 -f -o gpurun_out/r02_pass_v3_tma_n30 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-tdvp --no-matched > gpurun_out/r2f_ncu_full.log 2>&1; tail -1 gpurun_out/r2f_ncu_full.log | cut -c1-200
timeout 300 python scratch/loopback_prof.py 30 8 5 2>&1 | tail -3
timeout 300 python scratch/e2e_trace.py 30 3 2>&1 | tail -24
