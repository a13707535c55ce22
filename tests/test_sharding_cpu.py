"""CPU tests of the sharded (multi-GPU) path's host logic: the remote-term plan of the C library
reproduces the reference Hamiltonian when every rank applies its tile passes plus its remote
terms, and the torch.distributed plumbing (gloo, world_size 2) combines measurements correctly."""
import os
import socket

import numpy as np
import pytest

import pass_model
import qca_b200
import qca_oracle as oracle
from conftest import RuleNS
from qca_b200 import _lib, sharding


def k_apply_full(vec, n, d, lo, hi):
    xs = np.arange(1 << n, dtype=np.int64)
    act = pass_model.activity(xs, n, d, lo, hi)
    out = np.zeros_like(vec)
    for g in range(n):
        on = ((act >> g) & 1).astype(bool)
        sign = np.where((xs >> g) & 1, -1.0, 1.0)
        out[on] += sign[on] * vec[xs[on] ^ (1 << g)]
    return out


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("n,d,lo,hi", [(10, 1, 1, 2), (11, 2, 2, 4), (17, 2, 2, 4), (12, 3, 2, 5), (16, 1, 1, 3)])
def test_remote_plan_reproduces_operator(world, n, d, lo, hi):
    rules = RuleNS(n, d, lo, hi)
    rbits = world.bit_length() - 1
    nl = n - rbits
    rng = np.random.default_rng(world * 100 + n)
    vec = rng.standard_normal(1 << n)
    want = k_apply_full(vec, n, d, lo, hi)
    passes = _lib.plan_passes(nl)
    got = np.zeros_like(vec)
    xl = np.arange(1 << nl, dtype=np.int64)
    for rank in range(world):
        mine = sharding.local_slice(vec, world, rank)
        out = pass_model.apply_k_by_tiles(mine.copy(), passes, nl, n, d, lo, hi, prefix=rank << nl)
        ops = _lib.plan_remote(rules, world, rank)
        for op in ops:
            assert op["partner"] == rank ^ (1 << (op["qubit"] - nl)) and 0 <= op["pass_index"] < len(passes)
            assert op["sign"] == (-1 if (rank >> (op["qubit"] - nl)) & 1 else 1)
            v = (xl >> op["shift"]) & 15
            on = ((op["mask"] >> v) & 1).astype(bool) if d <= 4 else None
            # the generic kernel's criterion: activity bit of the sharded qubit
            act = pass_model.activity(xl | (rank << nl), n, d, lo, hi)
            on_bit = ((act >> op["qubit"]) & 1).astype(bool)
            if on is not None:
                assert np.array_equal(on, on_bit)
            partner = sharding.local_slice(vec, world, op["partner"])
            out[on_bit] += op["sign"] * partner[on_bit]
        if len(passes) > 1:  # spread over the passes: pass 0 takes at most one, later passes at most three
            per_pass = [sum(1 for op in ops if op["pass_index"] == p) for p in range(len(passes))]
            assert per_pass[0] <= 1 and max(per_pass) <= 3
        # terms the plan dropped must really be inactive on this rank
        planned = {op["qubit"] for op in ops}
        act = pass_model.activity(xl | (rank << nl), n, d, lo, hi)
        for q in range(nl, n):
            if q not in planned:
                assert not ((act >> q) & 1).any()
        got[rank << nl:(rank + 1) << nl] = out
    assert np.abs(got - want).max() < 1e-12


def measure_partial_model(psi_local, n, nl, rank, world, peers):
    """What qca_exact_measure_partial returns, restated in numpy (rotation drops out of |.|^2, |w|)."""
    sums = np.zeros(4 * n)
    for bit in range(nl):
        cell = n - 1 - bit
        t = psi_local.reshape(-1, 2, 1 << bit)
        a0, a1 = t[:, 0, :], t[:, 1, :]
        w = np.vdot(a1, a0)  # sum a0 conj(a1)
        sums[4 * cell:4 * cell + 4] = [np.vdot(a0, a0).real, np.vdot(a1, a1).real, w.real, w.imag]
    for j in range(world.bit_length() - 1):
        if (rank >> j) & 1:
            continue
        cell = n - 1 - (nl + j)
        other = peers[rank ^ (1 << j)]
        w = np.vdot(other, psi_local)
        sums[4 * cell:4 * cell + 4] = [np.vdot(psi_local, psi_local).real, np.vdot(other, other).real, w.real, w.imag]
    return sums


def _gloo_worker(rank, world, port, n, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert sharding.world_and_rank() == (world, rank)
        rng = np.random.default_rng(3)
        psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
        psi /= np.linalg.norm(psi)
        nl = n - (world.bit_length() - 1)
        mine = sharding.local_slice(psi, world, rank)
        slices = sharding.gather_objects(mine)
        assert np.array_equal(np.concatenate(slices), psi)
        part = measure_partial_model(mine, n, nl, rank, world, slices)
        pop, dpop, ent, bonds = sharding.combine_measurements(sharding.gather_objects(part), n)
        pop_o, dpop_o, ent_o, bonds_o = oracle.measure_vector(psi, n)
        ok = (np.abs(pop - pop_o).max() < 1e-12 and np.abs(ent - ent_o).max() < 1e-11
              and np.array_equal(bonds, bonds_o) and np.array_equal(dpop, dpop_o))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_gloo_measurement_reduction(world):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_gloo_worker, args=(world, port, 9, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_sharded_engine_refuses_without_device():
    if qca_b200.lib.qca_device_count() > 0:
        pytest.skip("a device is present")
    with pytest.raises(qca_b200.QcaError):
        _lib.ExactEngine(qca_b200.Rules(12, range(1, 2), 1), world_size=2, rank=1)
    with pytest.raises(qca_b200.QcaError) as e:
        _lib.plan_remote(qca_b200.Rules(12, range(1, 2), 1), 3, 0)
    assert e.value.code == _lib.QCA_ERR_ARG
