#!/usr/bin/env python
"""Benchmark of the exact time-evolution hot path (BASELINE.json: "exact steps/s at N=30").

    python bench.py --gpus N --steps K --warmup W           # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W   # the reference algorithm on host cores

A step is what the reference's loop does per plot step for --algorithm exact
(quantum_game.py:85-119): Algorithm.measure (populations, entropies) followed by
Exact.do_time_step with step_size * plot_step_interval = 1.0, i.e. psi <- exp(-i pi/2 H) psi.

One JSON line on stdout (rank 0).  `value` is device-resident throughput (CUDA events, max over
ranks); `e2e` repeats the measurement through the host-buffer API (upload psi from pinned host
memory, step, measure, download psi); `roofline` is for the dominant kernel (the tile pass);
`cpu_baseline` times the oracle's restatement of the reference algorithm on this box's cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "exact steps/s at N=30"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--num-cells", type=int, default=30)
    ap.add_argument("--distance", type=int, default=2)
    ap.add_argument("--activation-interval", type=int, nargs=2, default=(2, 4))
    ap.add_argument("--initial-state", default="triple_blinker")
    ap.add_argument("--step-size", type=float, default=1.0, help="effective exact step (reference default 0.005*200)")
    ap.add_argument("--force-complex", action="store_true", help="keep both real planes (general complex128 state)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-num-cells", type=int, default=12, help="chain length of the bounded CPU sample")
    ap.add_argument("--no-tdvp", action="store_true", help="skip the 2TDVP chi=256 leg")
    ap.add_argument("--tdvp-cells", type=int, default=64)
    ap.add_argument("--tdvp-chi", type=int, default=256)
    ap.add_argument("--tdvp-steps", type=int, default=3)
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, cell in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if cell.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's restatement of the reference algorithm
# ---------------------------------------------------------------------------------------------
def cpu_reference_run(args, steps: int, warmup: int) -> dict:
    """Dense reference algorithm (MPO.as_matrix -> calculate_U -> U@psi + measure) at the largest
    chain that finishes in seconds; the N=30 workload itself would need a 2^30 x 2^30 complex matrix
    (1.8e19 bytes) and cannot be run by the reference at all."""
    import numpy as np
    import qca_oracle as oracle
    n = args.ref_num_cells
    d, (lo, hi) = args.distance, args.activation_interval
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    t0 = time.perf_counter()
    h = oracle.mpo_as_matrix(oracle.mpo_tensors(n, d, lo, hi))
    u = oracle.calculate_U(h, args.step_size)
    build_s = time.perf_counter() - t0
    psi = oracle.product_state_vector(oracle.initial_plist(args.initial_state, n, d))
    for _ in range(warmup):
        oracle.measure_vector(psi, n); psi = oracle.exact_step(u, psi)
    t1 = time.perf_counter()
    for _ in range(steps):
        oracle.measure_vector(psi, n); psi = oracle.exact_step(u, psi)
    loop_s = time.perf_counter() - t1
    return {"value": steps / loop_s, "unit": "steps/s", "cores": threads, "kind": "port",
            "sample": (f"oracle port of the reference's dense algorithm at N={n} (not N={args.num_cells}: U would be "
                       f"2^{2 * args.num_cells} complex128): {steps} steps of measure+U@psi after a one-off "
                       f"{build_s:.1f} s as_matrix+eigh build (build excluded from value)"),
            "build_s": build_s, "ms_per_step": 1e3 * loop_s / steps, "num_cells": n}


def reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(args.steps, 1) * 20  # a dense matvec at N=12 is milliseconds; keep the sample a few seconds
    res = cpu_reference_run(args, steps, max(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": "steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, planes=None, extra={"reference_num_cells": res["num_cells"]}),
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, planes, extra=None) -> dict:
    lo, hi = args.activation_interval
    cfg = {"workload": f"exact, {args.initial_state}, --num-cells {args.num_cells}, --distance {args.distance}, "
                       f"--activation-interval {lo} {hi}, exact step {args.step_size} (x pi/2), measure every step",
           "num_cells": args.num_cells, "l2": "state planes (>= 8 GiB at N=30) are far larger than the 126 MB L2",
           "planes": planes}
    if extra:
        cfg.update(extra)
    return cfg


# ---------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------
def b200_arm(args) -> None:
    import numpy as np
    import torch
    import qca_b200
    from qca_b200 import _lib, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and not (world == 1 and args.gpus == 1):
        raise SystemExit(f"--gpus {args.gpus} must be launched with torch.distributed.run, one rank per GPU "
                         f"(WORLD_SIZE is {world})")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the B200 arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # NCCL prints its version banner on STDOUT otherwise
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # a real (non-default) stream: handle 0 would make the library create its own stream and the
    # torch events below would not see the kernels
    torch.cuda.set_stream(torch.cuda.Stream())
    stream = torch.cuda.current_stream().cuda_stream
    assert stream != 0
    lo, hi = args.activation_interval
    rules = qca_b200.Rules(args.num_cells, range(lo, hi), args.distance)
    plist = qca_b200.states.plist(args.initial_state, rules)
    flags = _lib.QCA_FLAG_FORCE_COMPLEX if args.force_complex else 0

    def make_engine(extra_flags=0):
        if world > 1:
            return sharding.ShardedExactEngine(rules, device=local_rank, flags=flags | extra_flags, stream=stream)
        return _lib.ExactEngine(rules, device=local_rank, flags=flags | extra_flags, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        return max(sharding.gather_objects(float(x)))

    def one_step(engine):
        engine.measure()          # Algorithm.measure: D2H of 4*N sums (+ host gather when sharded)
        engine.step(args.step_size, 1)

    # ---- device-resident throughput -------------------------------------------------------
    eng = make_engine()
    eng.set_product_state(plist)
    for _ in range(args.warmup):
        one_step(eng)
    barrier()
    eng.reset_stats()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ncu_range = os.environ.get("QCA_NCU_RANGE") == "1"   # ncu --profile-from-start off: only the timed region
    if ncu_range:
        torch.cuda.profiler.start()
    ev0.record()
    for _ in range(args.steps):
        one_step(eng)
    ev1.record()
    barrier()
    if ncu_range:
        torch.cuda.profiler.stop()
    clocks = sampler.stop()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    st = eng.stats()
    norm2 = eng.norm2()
    planes = st["planes"]
    value = args.steps / (ms * 1e-3)
    whole_step_gbs = st["pass_bytes"] / (ms * 1e-3) / 1e9
    nvlink_gbs = st["remote_bytes"] / (ms * 1e-3) / 1e9
    eng.close()

    # ---- per-launch duration of the dominant kernel (event pair around every launch) ---------
    prof = make_engine(_lib.QCA_FLAG_PROFILE)
    prof.set_product_state(plist)
    prof.step(args.step_size, 1)
    barrier()
    prof.reset_stats()
    prof.step(args.step_size, 1)
    barrier()
    pst = prof.stats()
    prof.close()
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    # DRAM bytes per launch of the same kernel from the committed `ncu --set full` capture (None if absent)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "pass_kernel_traffic.json")
    if os.path.exists(tpath):
        rec = json.load(open(tpath)).get(f"N{args.num_cells}_g{world}")
        if rec and planes == rec.get("planes"):
            traffic = rec["dram_bytes_per_launch"]
    applies = max(pst["pass_launches"] // max(pst["passes_per_apply"], 1), 1)
    bytes_per_launch = pst["pass_bytes"] / max(pst["pass_launches"], 1)
    avg_ms = pst["profiled_pass_ms"] / max(pst["profiled_pass_launches"], 1)
    achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "qca::pass_kernel_v2 (tile pass of the rule operator + fused Clenshaw update)",
                "avg_launch_ms_by_pass": [m / applies for m in pst["profiled_ms_by_pass"][:pst["passes_per_apply"]]],
                "bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_ms, "launches_per_step": pst["pass_launches"],
                "peak_source": peak_src, "whole_step_gbs": whole_step_gbs, "per_gpu": True,
                "nvlink_read_gbs_per_gpu": nvlink_gbs,
                "note": "per GPU (rank 0): achieved = local operand vectors read/written once per launch / CUDA-event "
                        "launch duration; whole_step_gbs = same bytes / the timed region including measurement; "
                        "partner-rank reads over NVLink are not counted in achieved"}

    # ---- end to end through host buffers ---------------------------------------------------
    e2e = None
    if not args.no_e2e:
        namps = (1 << args.num_cells) // world
        host = torch.empty(2 * namps, dtype=torch.float64)
        try:
            host = host.pin_memory()
            pinned = True
        except Exception:
            pinned = False
        e = make_engine()
        e.set_product_state(plist)
        raw = e._eng if world > 1 else e          # this rank's slice moves through the C ABI
        raw.get_state_ptr(host.data_ptr(), namps)

        def e2e_step():
            raw.set_state_ptr(host.data_ptr(), namps)   # Exact.psi setter: H2D of this rank's complex128 slice
            if world > 1:
                e._resolve()
            pop = e.measure()[0]                        # Algorithm.measure: D2H of the sums
            e.step(args.step_size, 1)                   # Exact.do_time_step
            raw.get_state_ptr(host.data_ptr(), namps)   # Exact.psi getter: D2H of the slice
            return pop

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(args.steps):
            e2e_step()
        a1.record()
        barrier()
        wall = time.perf_counter() - t0
        e2e_ms = max_over_ranks(max(a0.elapsed_time(a1), wall * 1e3))
        e2e = {"value": args.steps / (e2e_ms * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": 16 * namps * world,
               "d2h_bytes_per_step": (16 * namps + 8 * 4 * args.num_cells) * world, "ms_per_step": e2e_ms / args.steps,
               "pinned": pinned,
               "api": "ExactEngine.set_state(host psi) -> measure -> step -> get_state(host psi) over the C ABI, "
                      "every rank moving its own slice"}
        e.close()
        del host

    cpu = None if (args.no_cpu_baseline or world > 1 or rank != 0) else cpu_reference_run(args, 200, 5)
    tdvp = None
    if not args.no_tdvp and world == 1:
        try:
            tdvp = tdvp_leg(args, local_rank)
        except Exception as exc:  # the exact leg's numbers stand on their own
            tdvp = {"error": repr(exc)}
    line = {"metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, planes, {"chebyshev_terms": st["last_terms"],
                                                     "passes_per_term": st["passes_per_apply"],
                                                     "spectral_bound": st["spectral_bound"], "norm2_after": norm2,
                                                     "sharding": f"top {world.bit_length() - 1} qubits over {world} ranks, "
                                                                 "partner reads over NVLink peer memory" if world > 1 else "none"}),
            "roofline": roofline, "cpu_baseline": None if cpu is None else {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": e2e, "gpu_launches": int(st["kernel_launches"]) * world, "clocks": clocks, "tdvp": tdvp}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# second half of the BASELINE metric: 2TDVP sweeps/s at chi = 256 (single GPU)
# ---------------------------------------------------------------------------------------------
def tdvp_leg(args, device: int) -> dict:
    """configs[4]: 2tdvp, --num-cells 64, bond cap chi = 256.  A product state only reaches chi = 256
    after thousands of steps, so (north_star: "synthetic initial states of the named sizes") the MPS
    is a seeded random one whose bonds are already at the cap, with the SVD cut-off low enough to
    keep them there.  One time step = one right plus one left sweep (tdvp.py:50-63)."""
    import numpy as np
    import torch
    import qca_b200
    n, chi = args.tdvp_cells, args.tdvp_chi
    rules = qca_b200.Rules(n, range(1, 2), 1)
    targs = qca_b200.Args(rules=rules, step_size=0.005, algorithm="2tdvp", max_bond_dim=chi, svd_epsilon=1e-14)
    rng = np.random.default_rng(0)
    dims = [min(2 ** i, 2 ** (n - i), chi) for i in range(n + 1)]
    mps = qca_b200.MPS([(rng.standard_normal((2, dims[i], dims[i + 1])) + 1j * rng.standard_normal((2, dims[i], dims[i + 1])))
                        / np.sqrt(2 * dims[i]) for i in range(n)])
    h2d = sum(a.nbytes for a in mps.A)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    algo = qca_b200.TDVP(mps, qca_b200.MPO.hamiltonian_from_rules(rules), targs, device=device)
    torch.cuda.synchronize()
    init_s = time.perf_counter() - t0
    pop, dpop, sse, bond = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n + 1)
    for _ in range(max(args.warmup, 1)):
        algo.do_time_step()
    torch.cuda.synchronize()
    algo.heff_applications, algo.heff_flops = 0, 0.0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.tdvp_steps):
        algo.do_time_step()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    flops, napply = algo.heff_flops, algo.heff_applications
    # the contraction kernel itself, in situ: one more time step with an event pair around every
    # launch of the DMMA kernel (all bond sizes of the chain, L2 state as in the real sweep)
    from qca_b200 import _lib
    _lib.zgemm_profile(True)
    algo.do_time_step()
    torch.cuda.synchronize()
    zp = _lib.zgemm_profile(False)
    # through the plug-in API with a measurement every step (D2H of the N density matrices)
    t1 = time.perf_counter()
    for _ in range(args.tdvp_steps):
        algo.measure(pop, dpop, sse, bond)
        algo.do_time_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t1
    # denominator measured here: cuBLAS DGEMM (the box has no measured FP64 figure in MEASURED_PEAKS.json)
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    torch.matmul(a, b); torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(); torch.matmul(a, b); torch.matmul(a, b); g1.record(); torch.cuda.synchronize()
    dgemm_tflops = 2 * 2.0 * 8192 ** 3 / (g0.elapsed_time(g1) * 1e-3) / 1e12
    del a, b
    whole = flops / (ms * 1e-3) / 1e12
    achieved = zp["flops"] / (zp["ms"] * 1e-3) / 1e12 if zp["ms"] > 0 else 0.0
    return {"metric": "2TDVP sweeps/s at chi=256", "value": 2 * args.tdvp_steps / (ms * 1e-3), "unit": "sweeps/s",
            "ms_per_time_step": ms / args.tdvp_steps, "steps": args.tdvp_steps, "init_s": init_s,
            "config": {"workload": f"2tdvp, --num-cells {n}, --max-bond-dim {chi}, distance 1, interval [1,2), step 0.005, "
                                   "seeded random MPS at the bond cap, svd_epsilon 1e-14", "max_bond": int(max(dims))},
            "heff_applications_per_step": napply / args.tdvp_steps,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": dgemm_tflops, "unit": "TFLOP/s",
                         "frac": achieved / dgemm_tflops, "traffic": None,
                         "kernel": "qca::zgemm_dmma_kernel (L.psi and T.R contractions of H_eff, FP64 tensor cores)",
                         "launches_per_step": zp["launches"], "avg_launch_ms": zp["ms"] / max(zp["launches"], 1),
                         "kernel_share_of_step": zp["ms"] / (ms / args.tdvp_steps),
                         "whole_step_tflops": whole, "whole_step_frac": whole / dgemm_tflops,
                         "note": "achieved = FP64 operations of the DMMA launches of one time step (8 M N K per complex GEMM) / "
                                 "their summed CUDA-event durations, measured live; whole_step_tflops = the same operations / "
                                 "the WHOLE step time (Lanczos vector work, SVD, QR, environments included); peak = cuBLAS "
                                 "DGEMM 8192^3 measured in this run (nominal B200 FP64 tensor peak: 40 TFLOP/s)"},
            "e2e": {"value": 2 * args.tdvp_steps / e2e_s, "unit": "sweeps/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 64 * n, "h2d_bytes_once": h2d,
                    "api": "TDVP.measure + TDVP.do_time_step (state resident on the device between steps, as in the reference's loop)"}}


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
