#!/usr/bin/env python
"""Benchmark of the exact time-evolution hot path (BASELINE.json: "exact steps/s at N=30").

    python bench.py --gpus N --steps K --warmup W           # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W   # the reference algorithm on host cores

A step is what the reference's loop does per plot step for --algorithm exact
(quantum_game.py:85-119): Algorithm.measure (populations, entropies) followed by
Exact.do_time_step with step_size * plot_step_interval = 1.0, i.e. psi <- exp(-i pi/2 H) psi.

One JSON line on stdout (rank 0).  `value` is device-resident throughput (CUDA events, max over
ranks); `e2e` repeats the measurement through the host-buffer API (upload psi from pinned host
memory, measure, step, download psi; on one GPU as a two-deep pipeline over a batch of states so
that the PCIe copies overlap the other state's kernels); `roofline` is for the dominant kernel (the
tile pass); `cpu_baseline` / `--impl reference` time the CPU oracle on this box's cores ON THE SAME
WORKLOAD (matrix-free C port, a bounded row sample of one operator application, extrapolated; the
reference's own dense algorithm cannot hold N > 13) and, for the sizes the reference itself can
run, its dense algorithm next to this repo's GPU path (`matched`).  `checksum` holds the
populations and entropies after two steps from the named initial state: identical protocol for
every --gpus N, so the lines of a scaling run can be compared with each other and with the
committed expectation (tests/golden/bench_checksum.json).
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



def metric_name(args) -> str:
    return f"exact steps/s at N={args.num_cells}"


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
    ap.add_argument("--ref-num-cells", type=int, default=12, help="chain length of the dense-reference matched leg")
    ap.add_argument("--no-matched", action="store_true", help="skip the small-register legs (N=9, 12, 20) and their CPU runs")
    ap.add_argument("--e2e-steps", type=int, default=10, help="upper limit of timed end-to-end steps (each moves 2 x 16 GiB at N=30)")
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
# CPU legs: the oracle on this box's host cores (bench.py is one of the three places allowed to run oracle/)
# ---------------------------------------------------------------------------------------------
def host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def pin_host_threads() -> int:
    """All cores of the affinity mask for OpenMP (C oracle) and BLAS (numpy), whatever the launcher
    exported: torch.distributed.run sets OMP_NUM_THREADS=1, which silently cut the round-1 reference
    arm to one thread.  Returns the count in effect."""
    import qca_oracle_c as oc
    n = host_threads()
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass
    return oc.set_threads(n)


def available_host_bytes() -> int:
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except OSError:
        pass
    return 8 << 30


class MatrixFreeCpuSample:
    """The benchmark workload itself (same N, rule, initial state, step) on the host: the C oracle's
    matrix-free Chebyshev stepper (oracle/qca_oracle_c.c).  One full operator application at N = 30
    takes about a minute on 8 cores and a step has ~90 of them, so a timed sample is a block of ROWS of
    one application (2^rows_log2 consecutive output amplitudes, reading the whole 2^N input vector);
    steps/s = 1 / (terms * t_rows * 2^N / rows).  The axpy / accumulate passes and the measurement
    of a step are left out, which favours the CPU."""

    def __init__(self, args):
        import numpy as np
        import qca_oracle as oracle
        import qca_oracle_c as oc
        self.oc = oc
        self.threads = pin_host_threads()
        self.d, (self.lo, self.hi) = args.distance, args.activation_interval
        self.n_target = args.num_cells
        n = args.num_cells
        while n > 20 and (16 << n) > 0.4 * available_host_bytes():   # 2^n complex128 input vector must fit
            n -= 1
        self.n = n
        self.rows_log2 = min(n, 24)
        self.psi = oc.product_state(n, oracle.initial_plist(args.initial_state, n, self.d))
        self.out = np.empty(1 << self.rows_log2, dtype=np.complex128)
        self.terms = oc.chebyshev_terms(self.n_target, args.step_size)
        self.block = 0

    def sample_seconds(self) -> float:
        """One timed row block; successive calls walk through the vector."""
        nblocks = 1 << (self.n - self.rows_log2)
        x0 = ((self.block * 2654435761) % nblocks) << self.rows_log2    # scattered blocks: low and high index ranges alike
        self.block += 1
        t0 = time.perf_counter()
        self.oc.apply_h_rows(self.psi, self.n, self.d, self.lo, self.hi, x0, 1 << self.rows_log2, self.out)
        return time.perf_counter() - t0

    def steps_per_second(self, seconds_per_block: float) -> float:
        per_apply = seconds_per_block * float(1 << (self.n_target - self.rows_log2))
        return 1.0 / (self.terms * per_apply)

    def describe(self, nsamples: int, seconds_per_block: float) -> str:
        scaled = "" if self.n == self.n_target else (f" (host memory holds only N={self.n}: the block time of the "
                                                     f"smaller register is scaled by the row count)")
        return (f"C oracle, matrix-free (oracle/qca_oracle_c.c), same workload N={self.n_target}: {nsamples} timed blocks of "
                f"2^{self.rows_log2} rows of one application of H on the full 2^{self.n} complex128 vector "
                f"({seconds_per_block * 1e3:.0f} ms per block on {self.threads} threads){scaled}; step = {self.terms} "
                f"Chebyshev terms (Gershgorin scale R=N) x 2^{self.n_target - self.rows_log2} blocks; vector updates and "
                f"measurement not counted")


def cpu_matrix_free(args, nsamples: int, warm: int) -> dict:
    smp = MatrixFreeCpuSample(args)
    for _ in range(max(warm, 1)):
        smp.sample_seconds()
    times = [smp.sample_seconds() for _ in range(max(nsamples, 1))]
    sec = sum(times) / len(times)
    value = smp.steps_per_second(sec)
    return {"value": value, "unit": "steps/s", "cores": smp.threads, "kind": "port",
            "sample": smp.describe(len(times), sec), "ms_per_step": 1e3 / value, "seconds_sampled": sum(times)}


def cpu_dense_reference(n, d, lo, hi, state, step_size, steps: int, warmup: int) -> dict:
    """The reference's own algorithm at a size it can run (MPO.as_matrix -> calculate_U, then per step
    Exact.psi -> MPS.from_vector -> MPS.measure with its QR sweeps and scipy logm, then U @ psi;
    quantum_game.py:85-119), restated by oracle/qca_oracle.py, all host threads through numpy's BLAS."""
    import qca_oracle as oracle
    threads = pin_host_threads()
    t0 = time.perf_counter()
    u = oracle.calculate_U(oracle.mpo_as_matrix(oracle.mpo_tensors(n, d, lo, hi)), step_size)
    build_s = time.perf_counter() - t0
    psi = oracle.product_state_vector(oracle.initial_plist(state, n, d))
    for _ in range(warmup):
        oracle.measure_via_mps(psi, n); psi = oracle.exact_step(u, psi)
    t1 = time.perf_counter()
    for _ in range(steps):
        oracle.measure_via_mps(psi, n); psi = oracle.exact_step(u, psi)
    loop_s = time.perf_counter() - t1
    return {"value": steps / loop_s, "unit": "steps/s", "cores": threads, "kind": "port", "build_s": build_s,
            "steps": steps, "with_build": steps / (loop_s + build_s),
            "sample": f"dense reference algorithm (oracle/qca_oracle.py) at N={n}: {steps} steps of MPS.from_vector + "
                      f"MPS.measure (QR sweeps, logm) + U@psi; one-off as_matrix+eigh build {build_s:.1f} s not in value "
                      f"(with_build: steps / (loop + build))"}


def reference_arm(args) -> None:
    """--impl reference: the CPU oracle on the GPU arm's workload, metric and unit.  Each of the
    W + K "steps" is one bounded row-block sample (see MatrixFreeCpuSample)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_matrix_free(args, max(args.steps, 1), max(args.warmup, 1))
    dense = None
    if not args.no_matched:
        lo, hi = args.activation_interval
        dense = cpu_dense_reference(min(args.ref_num_cells, 12), args.distance, lo, hi, args.initial_state, args.step_size, 20, 2)
    line = {"impl": "reference", "metric": metric_name(args), "value": res["value"], "unit": "steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "dense_reference_algorithm": dense,
            "note": "the reference's dense algorithm (4^N matrix) cannot run this workload at all; value is the oracle's "
                    "matrix-free C port on the same N, rule, state and step, extrapolated from bounded row samples; "
                    "dense_reference_algorithm is the reference's own algorithm at the largest N that builds in seconds"}
    print(json.dumps(line), flush=True)


def workload_config(args) -> dict:
    """Identical for both arms: it names the workload only (arm-specific facts live in `details`)."""
    lo, hi = args.activation_interval
    return {"workload": f"exact, {args.initial_state}, --num-cells {args.num_cells}, --distance {args.distance}, "
                        f"--activation-interval {lo} {hi}, exact step {args.step_size} (x pi/2), measure every step",
            "num_cells": args.num_cells,
            "l2": f"state vectors of 2^{args.num_cells} amplitudes (>= 8 GiB at N=30) are far larger than the 126 MB L2"}


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
        # (NCCL_DEBUG stays unset: at VERSION or above NCCL prints its version banner on STDOUT, next to the JSON line)
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

    def make_engine(extra_flags=0, on_stream=None):
        if world > 1:
            return sharding.ShardedExactEngine(rules, device=local_rank, flags=flags | extra_flags, stream=on_stream or stream)
        return _lib.ExactEngine(rules, device=local_rank, flags=flags | extra_flags, stream=on_stream or stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        return max(sharding.gather_objects(float(x)))

    def one_step(engine):
        engine.measure()          # Algorithm.measure: D2H of 4*N sums (+ the cross-rank sum when sharded)
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
    # the same engine also yields the checksum: exactly two steps from the named initial state, whatever
    # --steps/--warmup/--gpus are, so every line of a scaling run must print the same numbers
    prof = make_engine(_lib.QCA_FLAG_PROFILE)
    prof.set_product_state(plist)
    prof.step(args.step_size, 1)
    barrier()
    prof.reset_stats()
    prof.step(args.step_size, 1)
    barrier()
    pst = prof.stats()
    ck_pop, _, ck_ent, _ = prof.measure()
    checksum = make_checksum(args, ck_pop, ck_ent, prof.norm2())
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
        if rec and planes == rec.get("planes") and rec.get("passes_per_term", 3) == st["passes_per_apply"]:
            traffic = rec["dram_bytes_per_launch"]
    # one GPU, >= 14 local qubits: 14-bit tiles, 512 threads, one CTA per SM (csrc/qca_pass3.cuh); sharded engines: 13-bit
    # tiles with remote operand slots, persistent CTAs (csrc/qca_pass.cuh)
    if world > 1:
        kernel_name = "qca::pass_kernel_v2p / pass_kernel_v2" if st["local_bits"] >= 13 else "qca::pass_kernel_generic"
    else:
        kernel_name = "qca::pass_kernel_v3" if st["local_bits"] >= 14 else ("qca::pass_kernel_v2" if st["local_bits"] >= 13 else "qca::pass_kernel_generic")
    applies = max(pst["pass_launches"] // max(pst["passes_per_apply"], 1), 1)
    bytes_per_launch = pst["pass_bytes"] / max(pst["pass_launches"], 1)
    avg_ms = pst["profiled_pass_ms"] / max(pst["profiled_pass_launches"], 1)
    achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kernel_name + " (tile pass of the rule operator + fused Clenshaw update)",
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
        e2e = e2e_leg(args, world, local_rank, make_engine, barrier, max_over_ranks)

    cpu, matched, tdvp = None, None, None
    if world == 1 and rank == 0:
        if not args.no_cpu_baseline:
            cpu = cpu_matrix_free(args, 12, 2)
        if not args.no_matched:
            matched = matched_legs(args, local_rank)
        if not args.no_tdvp:
            try:
                tdvp = tdvp_leg(args, local_rank)
            except Exception as exc:  # the exact leg's numbers stand on their own
                tdvp = {"error": repr(exc)}
    line = {"metric": metric_name(args), "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "details": {"planes": planes, "chebyshev_terms": st["last_terms"], "passes_per_term": st["passes_per_apply"],
                        "spectral_bound": st["spectral_bound"], "norm2_after": norm2,
                        "sharding": (f"{world.bit_length() - 1} qubits over {world} ranks, partner reads over NVLink peer "
                                     "memory inside the tile-pass kernel") if world > 1 else "none"},
            "checksum": checksum,
            "roofline": roofline, "cpu_baseline": None if cpu is None else {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": e2e, "gpu_launches": int(st["kernel_launches"]) * world, "clocks": clocks, "matched": matched, "tdvp": tdvp}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if checksum.get("ok") is False:
        raise SystemExit(f"bench checksum differs from the committed expectation by {checksum['max_abs_diff_vs_committed']:.3e}")


def make_checksum(args, pop, ent, norm2) -> dict:
    """Populations and entropies after exactly two steps from the named initial state, compared with the
    committed expectation of the same workload (written by a 1-GPU run, tests/golden/bench_checksum.json;
    the kernels behind it are tied to the oracle at N <= 26 by the test-suite)."""
    lo, hi = args.activation_interval
    key = f"{args.initial_state}_N{args.num_cells}_d{args.distance}_{lo}_{hi}_step{args.step_size}"
    out = {"protocol": "measure after 2 steps from the initial state", "key": key, "norm2": norm2,
           "population": [float(x) for x in pop], "entropy": [float(x) for x in ent],
           "max_abs_diff_vs_committed": None, "ok": None}
    path = os.path.join(ROOT, "tests", "golden", "bench_checksum.json")
    if os.path.exists(path):
        rec = json.load(open(path)).get(key)
        if rec:
            diff = max(max(abs(a - b) for a, b in zip(out["population"], rec["population"])),
                       max(abs(a - b) for a, b in zip(out["entropy"], rec["entropy"])))
            out["max_abs_diff_vs_committed"] = diff
            out["ok"] = bool(diff < 1e-10)
    return out


def e2e_leg(args, world, local_rank, make_engine, barrier, max_over_ranks) -> dict:
    """Same metric through the host-buffer API: every step uploads that step's input state from pinned
    host memory (Exact.psi setter), measures, steps, and downloads the result (Exact.psi getter).
    One GPU: a batch of independent states goes through TWO engines on two streams, each driven by its
    own host thread (ctypes releases the GIL), so the PCIe copies of one state overlap the kernels of the
    other; the sequential single-engine variant is reported beside it.  Sharded runs use the sequential
    variant (every rank moves its own slice)."""
    import threading as th
    import torch
    namps = (1 << args.num_cells) // world
    steps = max(1, min(args.steps, args.e2e_steps))
    lo, hi = args.activation_interval
    import qca_b200
    plist = qca_b200.states.plist(args.initial_state, qca_b200.Rules(args.num_cells, range(lo, hi), args.distance))

    def pinned_buffer():
        host = torch.empty(2 * namps, dtype=torch.float64)
        try:
            return host.pin_memory(), True
        except Exception:
            return host, False

    def run_sequential(e, host, count, gate=None, sync=None):
        """gate/sync (pipelined variant): the compute phase of a step holds `gate` until the device has finished
        it, so the two engines take turns on the SMs while the other one's copies run."""
        raw = e._eng if world > 1 else e          # this rank's slice moves through the C ABI
        for _ in range(count):
            raw.set_state_ptr(host.data_ptr(), namps)   # H2D of this rank's complex128 slice
            if world > 1:
                e._resolve()
            if gate is not None:
                gate.acquire()
            try:
                e.measure()                             # D2H of the sums
                e.step(args.step_size, 1)
                if sync is not None:
                    sync()
            finally:
                if gate is not None:
                    gate.release()
            raw.get_state_ptr(host.data_ptr(), namps)   # D2H of the slice

    host0, pinned = pinned_buffer()
    e0 = make_engine()
    e0.set_product_state(plist)
    (e0._eng if world > 1 else e0).get_state_ptr(host0.data_ptr(), namps)
    run_sequential(e0, host0, 1)
    barrier()
    t0 = time.perf_counter()
    run_sequential(e0, host0, steps)
    barrier()
    seq_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    out = {"value": steps / (seq_ms * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": 16 * namps * world,
           "d2h_bytes_per_step": (16 * namps + 8 * 4 * args.num_cells) * world, "ms_per_step": seq_ms / steps,
           "steps": steps, "pinned": pinned, "variant": "sequential",
           "api": "ExactEngine.set_state(host psi) -> measure -> step -> get_state(host psi) over the C ABI, "
                  "every rank moving its own slice",
           "sequential": {"value": steps / (seq_ms * 1e-3), "ms_per_step": seq_ms / steps}}
    if world == 1:
        # two-deep pipeline over a batch of `2 * ceil(steps / 2)` states
        s1 = torch.cuda.Stream()
        host1, _ = pinned_buffer()
        host1.copy_(host0)
        e1 = make_engine(on_stream=s1.cuda_stream)
        e1.set_product_state(plist)
        # the pipeline is timed over a longer batch than the sequential variant: its fill (first upload) and drain
        # (last download) are part of the timed region and should not dominate it
        per_engine = (max(1, args.e2e_steps) + 1) // 2
        run_sequential(e1, host1, 1)
        torch.cuda.synchronize()
        gate = th.Lock()
        s0 = torch.cuda.current_stream()
        threads = [th.Thread(target=run_sequential, args=(e, h, per_engine, gate, st.synchronize))
                   for e, h, st in ((e0, host0, s0), (e1, host1, s1))]
        t0 = time.perf_counter()
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        torch.cuda.synchronize()
        pipe_ms = (time.perf_counter() - t0) * 1e3
        e1.close()
        done = 2 * per_engine
        out.update({"value": done / (pipe_ms * 1e-3), "ms_per_step": pipe_ms / done, "steps": done, "variant": "pipelined",
                    "api": out["api"] + "; two engines on two streams, one host thread each: the engines take turns on the SMs and the "
                                        "copies of one state overlap the kernels of the other (batch of independent states)",
                    "pipelined": {"value": done / (pipe_ms * 1e-3), "ms_per_step": pipe_ms / done}})
        del host1
    e0.close()
    del host0
    return out


def matched_legs(args, device: int) -> list:
    """Same-config comparisons at the sizes the reference itself runs: this repo's Exact plug-in
    (measure + do_time_step per step, host loop included) next to the reference's dense algorithm on the
    host cores (cpu_dense_reference), plus BASELINE configs[1] (N=20, beyond the dense algorithm) next to
    the matrix-free C port."""
    import numpy as np
    import torch
    import qca_b200
    import qca_oracle_c as oc
    import qca_oracle as oracle
    out = []

    def gpu_run(n, d, lo, hi, state, tau, steps):
        rules = qca_b200.Rules(n, range(lo, hi), d)
        algo = qca_b200.Exact(qca_b200.states.make(state, rules), qca_b200.MPO.hamiltonian_from_rules(rules),
                              qca_b200.Args(rules=rules, step_size=tau), device=device)
        pop, dpop, ent, bond = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n + 1)
        for _ in range(5):
            algo.measure(pop, dpop, ent, bond); algo.do_time_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            algo.measure(pop, dpop, ent, bond); algo.do_time_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = algo.engine.stats()
        algo.engine.close()
        return steps / dt, st, pop.copy()

    for (label, n, d, lo, hi, state, tau, steps, cpu_steps) in (
            ("BASELINE configs[0]: exact, single, --num-cells 9, 100 steps, distance 1", 9, 1, 1, 2, "single", 1.0, 100, 100),
            (f"exact, {args.initial_state}, --num-cells {args.ref_num_cells}, distance {args.distance}, interval "
             f"[{args.activation_interval[0]},{args.activation_interval[1]}): the largest register whose dense U builds in seconds",
             args.ref_num_cells, args.distance, args.activation_interval[0], args.activation_interval[1], args.initial_state,
             args.step_size, 200, 40)):
        g, st, pop = gpu_run(n, d, lo, hi, state, tau, steps)
        c = cpu_dense_reference(n, d, lo, hi, state, tau, cpu_steps, 2)
        out.append({"workload": label + f", effective step {tau}, measure every step", "num_cells": n,
                    "gpu_steps_per_s": g, "gpu_steps": steps, "gpu_launches_per_step": st["kernel_launches"] / (steps + 5),
                    "cpu_steps_per_s": c["value"], "cpu_steps_per_s_with_build": c["with_build"], "cpu_build_s": c["build_s"],
                    "cpu_cores": c["cores"], "cpu_kind": "dense reference algorithm, oracle port", "ratio": g / c["value"]})
    # configs[1]: N=20 blinker, 1000 steps -- the dense algorithm would need a 2^40-entry matrix
    n, d, lo, hi, state, tau = 20, 1, 1, 2, "blinker", 1.0
    g, st, pop = gpu_run(n, d, lo, hi, state, tau, 1000)
    threads = pin_host_threads()
    ref = oc.Stepper(n, d, lo, hi)
    ref.set_product_state(oracle.initial_plist(state, n, d))
    ref.step(tau)
    t0 = time.perf_counter()
    for _ in range(2):
        ref.measure(); ref.step(tau)
    c = 2 / (time.perf_counter() - t0)
    out.append({"workload": "BASELINE configs[1]: exact, blinker, --num-cells 20, 1000 steps, distance 1, effective step 1.0, "
                            "measure every step", "num_cells": n, "gpu_steps_per_s": g, "gpu_steps": 1000,
                "gpu_launches_per_step": st["kernel_launches"] / 1005, "cpu_steps_per_s": c, "cpu_cores": threads,
                "cpu_kind": "matrix-free C port (2 full steps); the dense reference algorithm cannot hold N=20", "ratio": g / c})
    return out


# ---------------------------------------------------------------------------------------------
# second half of the BASELINE metric: 2TDVP sweeps/s at chi = 256 (single GPU)
# ---------------------------------------------------------------------------------------------
def random_mps_at_cap(n: int, chi: int, seed: int = 0):
    import numpy as np
    rng = np.random.default_rng(seed)
    dims = [min(2 ** i, 2 ** (n - i), chi) for i in range(n + 1)]
    return [(rng.standard_normal((2, dims[i], dims[i + 1])) + 1j * rng.standard_normal((2, dims[i], dims[i + 1])))
            / np.sqrt(2 * dims[i]) for i in range(n)], dims


def tdvp_leg(args, device: int) -> dict:
    """configs[4]: 2tdvp, --num-cells 64, bond cap chi = 256.  A product state only reaches chi = 256
    after thousands of steps, so (north_star: "synthetic initial states of the named sizes") the headline
    MPS is a seeded random one whose bonds are already at the cap, with the SVD cut-off low enough to
    keep them there.  One time step = one right plus one left sweep (tdvp.py:50-63).  Beside it: the
    literal configs[4] start (blinker product state, the reference's default cut-off: bond-growth regime)
    and a same-config CPU comparison at the bond dimension the reference's dense H_eff can still do."""
    import numpy as np
    import torch
    import qca_b200
    from qca_b200 import _lib
    n, chi = args.tdvp_cells, args.tdvp_chi
    rules = qca_b200.Rules(n, range(1, 2), 1)
    targs = qca_b200.Args(rules=rules, step_size=0.005, algorithm="2tdvp", max_bond_dim=chi, svd_epsilon=1e-14)
    tensors, dims = random_mps_at_cap(n, chi)
    mps = qca_b200.MPS(tensors)
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
    spec_stats = {"speculative_splits": int(algo.speculative_splits), "repeated_steps": int(algo.repeated_steps),
                  "note": "splits at the bond cap run without a host read (cuSOLVER zheevd called directly, flags checked once per "
                          "step); counted over warm-up + timed steps"}
    # the contraction kernel itself, in situ: one more time step with an event pair around every
    # launch of the DMMA kernel (all bond sizes of the chain, L2 state as in the real sweep)
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
    norm_err = abs(float(np.vdot(*(2 * [algo._A[0].cpu().numpy().reshape(-1)])).real) - 1.0)  # centre tensor of a normalised MPS
    del algo
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

    # ---- literal configs[4] start: blinker product state, reference defaults (bond-growth regime) ----------
    growth = None
    try:
        gargs = qca_b200.Args(rules=rules, step_size=0.005, algorithm="2tdvp", max_bond_dim=chi, svd_epsilon=5e-5)
        galgo = qca_b200.TDVP(qca_b200.states.make("blinker", rules), qca_b200.MPO.hamiltonian_from_rules(rules), gargs, device=device)
        for _ in range(3):
            galgo.do_time_step()
        torch.cuda.synchronize()
        gsteps = 30
        t2 = time.perf_counter()
        for _ in range(gsteps):
            galgo.do_time_step()
        torch.cuda.synchronize()
        gs = time.perf_counter() - t2
        galgo.measure(pop, dpop, sse, bond)
        growth = {"workload": "2tdvp, blinker, --num-cells 64, --max-bond-dim 256, step 0.005, svd_epsilon 5e-5 (reference "
                              "defaults): steps 4..33 from the product state", "value": 2 * gsteps / gs, "unit": "sweeps/s",
                  "max_bond_after": int(bond.max()), "population_sum": float(pop.sum())}
        del galgo
    except Exception as exc:
        growth = {"error": repr(exc)}

    # ---- same-config CPU comparison at a bond dimension the reference's dense H_eff can do ----------------
    cpu = None
    if not args.no_cpu_baseline:
        try:
            cpu = tdvp_cpu_compare(device)
        except Exception as exc:
            cpu = {"error": repr(exc)}
    return {"metric": "2TDVP sweeps/s at chi=256", "value": 2 * args.tdvp_steps / (ms * 1e-3), "unit": "sweeps/s",
            "ms_per_time_step": ms / args.tdvp_steps, "steps": args.tdvp_steps, "init_s": init_s,
            "config": {"workload": f"2tdvp, --num-cells {n}, --max-bond-dim {chi}, distance 1, interval [1,2), step 0.005, "
                                   "seeded random MPS at the bond cap, svd_epsilon 1e-14", "max_bond": int(max(dims))},
            "heff_applications_per_step": napply / args.tdvp_steps, "centre_norm_error": norm_err,
            "split": spec_stats,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": dgemm_tflops, "unit": "TFLOP/s",
                         "frac": achieved / dgemm_tflops, "traffic": None,
                         "kernel": "qca::zgemm_dmma_kernel (L.psi and T.R contractions of H_eff, FP64 tensor cores)",
                         "launches_per_step": zp["launches"], "avg_launch_ms": zp["ms"] / max(zp["launches"], 1),
                         "kernel_share_of_step": zp["ms"] / (ms / args.tdvp_steps),
                         "whole_step_tflops": whole, "whole_step_frac": whole / dgemm_tflops,
                         "note": "achieved = FP64 operations of the DMMA launches of one time step (8 M N K per complex GEMM) / "
                                 "their summed CUDA-event durations, measured live; whole_step_tflops = the H_eff contraction "
                                 "operations / the WHOLE step time (Lanczos vector work, SVD, QR, environments included); peak = "
                                 "cuBLAS DGEMM 8192^3 measured in this run (nominal B200 FP64 tensor peak: 40 TFLOP/s)"},
            "e2e": {"value": 2 * args.tdvp_steps / e2e_s, "unit": "sweeps/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 64 * n, "h2d_bytes_once": h2d,
                    "api": "TDVP.measure + TDVP.do_time_step (state resident on the device between steps, as in the reference's loop)"},
            "growth_regime": growth, "cpu_baseline": cpu}


def tdvp_cpu_compare(device: int, n: int = 12, chi: int = 8) -> dict:
    """2tdvp on the same seeded random MPS (n cells, bonds at the cap chi) through the B200 TDVP and
    through the oracle's restatement of the reference's algorithm (dense (4 chi^2)^2 H_eff + eigh per
    bond, oracle/tdvp_oracle.py, gauge-consistent variant) on the host cores."""
    import numpy as np
    import torch
    import qca_b200
    import qca_oracle as oracle
    import tdvp_oracle
    threads = pin_host_threads()
    rules = qca_b200.Rules(n, range(1, 2), 1)
    tensors, dims = random_mps_at_cap(n, chi, seed=1)
    dt, eps = 0.005, 1e-12
    cpu_algo = tdvp_oracle.TDVPOracle([t.copy() for t in tensors], oracle.mpo_tensors(n, 1, 1, 2), "2tdvp", dt, chi, eps,
                                      consistent=True)
    t0 = time.perf_counter()
    cpu_algo.step()
    cpu_s = time.perf_counter() - t0
    targs = qca_b200.Args(rules=rules, step_size=dt, algorithm="2tdvp", max_bond_dim=chi, svd_epsilon=eps)
    algo = qca_b200.TDVP(qca_b200.MPS([t.copy() for t in tensors]), qca_b200.MPO.hamiltonian_from_rules(rules), targs, device=device)
    algo.do_time_step()
    torch.cuda.synchronize()
    pop_g, pop_c = np.zeros(n), tdvp_oracle.measure_mps(cpu_algo.a)[0]
    algo.measure(pop_g, np.zeros(n), np.zeros(n), np.zeros(n + 1))
    steps = 5
    t1 = time.perf_counter()
    for _ in range(steps):
        algo.do_time_step()
    torch.cuda.synchronize()
    gpu_s = (time.perf_counter() - t1) / steps
    return {"workload": f"2tdvp, {n} cells, seeded random MPS at the bond cap chi={chi}, step {dt}, svd_epsilon {eps}",
            "value": 2.0 / cpu_s, "unit": "sweeps/s", "cores": threads, "kind": "port",
            "sample": "one time step (right + left sweep) of oracle/tdvp_oracle.py: dense H_eff + eigh per bond as in the reference",
            "gpu_sweeps_per_s": 2.0 / gpu_s, "ratio": cpu_s / gpu_s,
            "population_max_abs_diff_after_1_step": float(np.abs(pop_g - pop_c).max())}


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
