#!/bin/bash
# register-resident Q formation: QR tests, 1TDVP at chi = 256, then the default bench line with the final build
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_tdvp_gpu.py -m gpu -x -q -k "qr or 1tdvp" 2>&1 | tail -3 | tee gpurun_out/r2p_pytest_qr.log
timeout 150 python scratch/tdvp_prof2.py 64 256 1tdvp > gpurun_out/r2p_tdvp1_prof.txt 2>&1; grep -v "Warn\|warn" gpurun_out/r2p_tdvp1_prof.txt | head -8
timeout 400 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2p_bench.json"))
print("value", d["value"], "frac", d["roofline"]["frac"], d["roofline"]["avg_launch_ms_by_pass"], "e2e", d["e2e"]["value"], d["e2e"].get("sequential"), "checksum", d["checksum"]["ok"],
      "tdvp", d["tdvp"]["value"] if d.get("tdvp") and "value" in d["tdvp"] else d.get("tdvp"), "matched", [round(m["ratio"], 1) for m in d.get("matched", [])], d["clocks"])
PY
