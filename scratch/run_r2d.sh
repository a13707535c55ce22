#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2d_pytest.log; tail -6 gpurun_out/r2d_pytest.log
timeout 500 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2d_bench.json"))
print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"].get("sequential"), "tdvp", d["tdvp"]["value"] if d.get("tdvp") and "value" in d["tdvp"] else d.get("tdvp"))
PY
QCA_NCU_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_v3_n30.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-tdvp --no-matched > gpurun_out/r2d_ncu_list.log 2>&1; tail -2 gpurun_out/r2d_ncu_list.log | cut -c1-300
QCA_NCU_RANGE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:pass_kernel_v3 --launch-count 3 -f -o gpurun_out/r02_pass_v3_n30 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-tdvp --no-matched > gpurun_out/r2d_ncu_full.log 2>&1; tail -2 gpurun_out/r2d_ncu_full.log | cut -c1-300; ls -la gpurun_out/*.ncu-rep
