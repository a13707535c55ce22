#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_exact_multigpu.py -m gpu -q -rA 2>&1 | tail -12 > gpurun_out/r2l_mgpu_pytest_2gpu.log; tail -7 gpurun_out/r2l_mgpu_pytest_2gpu.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline --no-matched --no-tdvp 2> gpurun_out/r2l_bench_2gpu.err | grep "^{" > gpurun_out/r2l_bench_2gpu.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2l_bench_2gpu.json")); r = d["roofline"]
print("2gpu steps/s", round(d["value"], 4), "ms by pass", [round(x, 3) for x in r["avg_launch_ms_by_pass"]], "nvlink", round(r["nvlink_read_gbs_per_gpu"], 1),
      "checksum", d["checksum"]["ok"], d["checksum"]["max_abs_diff_vs_committed"], "e2e", d["e2e"]["value"] if d.get("e2e") else None)
PY
grep -v "OMP_NUM\|^\*\*\*" gpurun_out/r2l_bench_2gpu.err | tail -3
