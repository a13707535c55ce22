#!/bin/bash
# One 8-GPU session (charged 8x: keep it short): parity of the sharded path at 8 ranks, strong-scaling bench lines
# with the persistent and the per-tile sharded kernels, N=33 on 8, N=30 on 4.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_exact_multigpu.py -m gpu -q -rA -k "8" 2>&1 | tail -12 > gpurun_out/r2h_mgpu_pytest_8gpu.log
tail -5 gpurun_out/r2h_mgpu_pytest_8gpu.log
run() {  # tag gpus extra-env...
  tag=$1; g=$2; shift 2
  env "$@" timeout ${TMO:-100} python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2961$g \
      bench.py --gpus $g --steps ${STEPS:-4} --warmup 3 --no-e2e --no-cpu-baseline --no-matched --no-tdvp ${NCELLS:+--num-cells $NCELLS} 2> gpurun_out/r2h_bench_${tag}.err | grep "^{" > gpurun_out/r2h_bench_${tag}.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2h_bench_${tag}.json"))
    r = d["roofline"]
    print("${tag}", "steps/s", round(d["value"], 4), "ms/step", round(d["ms_per_step"], 2), "ms by pass", [round(x, 3) for x in r["avg_launch_ms_by_pass"]],
          "nvlink GB/s", round(r["nvlink_read_gbs_per_gpu"], 1), "checksum ok", d["checksum"]["ok"], d["checksum"]["max_abs_diff_vs_committed"])
except Exception as e:
    print("${tag} FAILED", e)
PY
}
run 8gpu_persistent 8 QCA_X=1
run 8gpu_per_tile 8 QCA_PERSISTENT_CTAS=0
NCELLS=33 STEPS=2 TMO=150 run 8gpu_n33 8 QCA_X=1
run 4gpu_persistent 4 QCA_X=1
