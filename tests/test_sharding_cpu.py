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
@pytest.mark.parametrize("n,d,lo,hi", [(10, 1, 1, 2), (11, 2, 2, 4), (17, 2, 2, 4), (12, 3, 2, 5), (16, 1, 1, 3),
                                       (21, 1, 1, 2), (22, 2, 2, 4)])
def test_remote_plan_reproduces_operator(world, n, d, lo, hi):
    """Every rank applies its tile passes plus its remote terms (as planned by the C library, for the
    sharded-qubit layout the C library chose: scattered cells for the larger registers, top cells
    otherwise) and the ranks together reproduce the operator on the full register."""
    if n >= 21 and world == 2:
        pytest.skip("same layout as the smaller cases")
    rules = RuleNS(n, d, lo, hi)
    rbits = world.bit_length() - 1
    nl = n - rbits
    positions = _lib.plan_shard(rules, world)
    assert positions == sorted(positions) and len(positions) == rbits and positions[-1] == n - 1
    scattered = rbits >= 2 and positions != list(range(nl, n))
    assert scattered == (rbits >= 2 and d <= 2 and n - 1 - (rbits - 1) * (d + 1) >= 13 + d + 1)
    rng = np.random.default_rng(world * 100 + n)
    vec = rng.standard_normal(1 << n)
    want = k_apply_full(vec, n, d, lo, hi)
    passes = _lib.plan_passes(nl)
    got = np.zeros_like(vec)
    xl = np.arange(1 << nl, dtype=np.int64)
    loads = []
    for rank in range(world):
        idx = sharding.local_indices(n, positions, rank)
        assert np.array_equal(idx, pass_model.expand(xl, positions, rank))
        mine = vec[idx]
        out = pass_model.apply_k_by_tiles(mine.copy(), passes, nl, n, d, lo, hi, positions, rank)
        ops = _lib.plan_remote(rules, world, rank)
        act = pass_model.activity(idx, n, d, lo, hi)
        load = 0.0
        for op in ops:
            j = positions.index(op["qubit"])
            assert op["partner"] == rank ^ (1 << j) and 0 <= op["pass_index"] < len(passes)
            assert op["sign"] == (-1 if (rank >> j) & 1 else 1)
            on_bit = ((act >> op["qubit"]) & 1).astype(bool)      # the generic kernel's criterion
            if op["window_bits"] <= 4:                            # the fast kernel's criterion
                v = (xl >> op["shift"]) & 15
                assert np.array_equal(((op["mask"] >> v) & 1).astype(bool), on_bit)
            partner = vec[sharding.local_indices(n, positions, op["partner"])]
            out[on_bit] += op["sign"] * partner[on_bit]
            load += on_bit.mean()
        loads.append(load)
        if len(passes) > 1:  # spread over the passes: pass 0 takes at most one, later passes at most three
            per_pass = [sum(1 for op in ops if op["pass_index"] == p) for p in range(len(passes))]
            assert per_pass[0] <= 1 and max(per_pass) <= 3
        planned = {op["qubit"] for op in ops}
        for q in positions:  # terms the plan dropped must really be inactive on this rank
            if q not in planned:
                assert not ((act >> q) & 1).any()
        got[idx] = out
    assert np.abs(got - want).max() < 1e-12
    if scattered:  # the point of the scattered layout: every rank pulls the same amount over NVLink
        assert max(loads) - min(loads) < 1e-12


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("n,d,lo,hi", [(10, 1, 1, 2), (12, 2, 2, 4), (13, 1, 1, 2), (15, 2, 2, 4), (17, 2, 2, 4), (18, 1, 1, 2),
                                       (19, 1, 1, 2), (20, 2, 2, 4), (22, 2, 2, 4), (24, 2, 1, 3), (26, 2, 2, 4), (30, 2, 2, 4),
                                       (33, 2, 2, 4)])
def test_rotation_applies_every_remote_term_exactly_once(world, n, d, lo, hi):
    """Fast-kernel placement (qca_plan_rotation): over the passes of one application every remote term
    is carried by exactly one (pass, slot) for every rotation value, no slot carries two terms, and --
    for the registers that matter -- every pass pulls the same share of every term."""
    rules = RuleNS(n, d, lo, hi)
    rbits = world.bit_length() - 1
    nl = n - rbits
    passes = _lib.plan_passes(nl)
    if nl < 13 or len(passes) < 2:   # generic kernel regime: the plan must still exist (nslots 0 = "no rotation")
        for rank in range(world):
            rot = _lib.plan_rotation(rules, world, rank)
            assert 0 <= rot["nslots"] <= 2
        return
    for rank in range(world):
        ops = _lib.plan_remote(rules, world, rank)
        rot = _lib.plan_rotation(rules, world, rank)
        assert rot["npasses"] == min(len(passes), 4) and rot["rot_shift"] == 13
        assert rot["nslots"] == -(-len(ops) // rot["npasses"]) <= 2
        rots = [(rot["rot_word"] >> (2 * v)) & 3 for v in range(16)]
        assert max(rots) < rot["npasses"]
        counts = np.bincount(rots, minlength=rot["npasses"])
        assert counts.max() - counts.min() <= 1                       # rotations equally frequent
        op_of = rot["op_of"]
        for r in range(rot["npasses"]):
            carried = [op_of[p, s, r] for p in range(rot["npasses"]) for s in range(rot["nslots"]) if op_of[p, s, r] >= 0]
            assert sorted(carried) == list(range(len(ops)))           # every term exactly once
        assert (op_of[rot["npasses"]:] == -1).all() and (op_of[:, rot["nslots"]:] == -1).all()
        # share of term j carried by pass p = fraction of rotation values that send it there
        for j in range(len(ops)):
            share = [sum(counts[r] for r in range(rot["npasses"]) if j in op_of[p, :, r]) / 16.0 for p in range(rot["npasses"])]
            assert abs(sum(share) - 1.0) < 1e-12 and max(share) - min(share) <= 1 / 16 + 1e-12
    if n <= 22:   # numerically, on the full operator: rotation-placed terms == static placement
        positions = _lib.plan_shard(rules, world)
        rng = np.random.default_rng(n)
        vec = rng.standard_normal(1 << n)
        xl = np.arange(1 << nl, dtype=np.int64)
        for rank in range(world):
            ops = _lib.plan_remote(rules, world, rank)
            rot = _lib.plan_rotation(rules, world, rank)
            r_of_x = (rot["rot_word"] >> (2 * ((xl >> rot["rot_shift"]) & 15))) & 3
            want = np.zeros(1 << nl)
            got = np.zeros(1 << nl)
            for op in ops:
                partner = vec[sharding.local_indices(n, positions, op["partner"])]
                on = ((op["mask"] >> ((xl >> op["shift"]) & 15)) & 1).astype(bool)
                want[on] += op["sign"] * partner[on]
            for p in range(rot["npasses"]):
                for s_ in range(rot["nslots"]):
                    for r in range(rot["npasses"]):
                        j = rot["op_of"][p, s_, r]
                        if j < 0:
                            continue
                        op = ops[j]
                        partner = vec[sharding.local_indices(n, positions, op["partner"])]
                        on = ((op["mask"] >> ((xl >> op["shift"]) & 15)) & 1).astype(bool) & (r_of_x == r)
                        got[on] += op["sign"] * partner[on]
            assert np.abs(got - want).max() < 1e-12   # (terms are added in another order)


def measure_partial_model(psi_local, n, nl, rank, world, peers, positions):
    """What qca_exact_measure_partial returns, restated in numpy (rotation drops out of |.|^2, |w|)."""
    sums = np.zeros(4 * n)
    for bit in range(nl):
        cell = n - 1 - pass_model.global_pos(bit, positions)
        t = psi_local.reshape(-1, 2, 1 << bit)
        a0, a1 = t[:, 0, :], t[:, 1, :]
        w = np.vdot(a1, a0)  # sum a0 conj(a1)
        sums[4 * cell:4 * cell + 4] = [np.vdot(a0, a0).real, np.vdot(a1, a1).real, w.real, w.imag]
    for j in range(world.bit_length() - 1):
        if (rank >> j) & 1:
            continue
        cell = n - 1 - positions[j]
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
        positions = [3, 8][:world.bit_length() - 1] if world == 4 else [n - 1]   # scattered and top layouts
        mine = sharding.local_slice(psi, positions, rank)
        slices = sharding.gather_objects(mine)
        assert np.array_equal(sharding.assemble(slices, positions), psi)
        part = measure_partial_model(mine, n, nl, rank, world, slices, positions)
        pop, dpop, ent, bonds = sharding.combine_measurements(sharding.gather_objects(part), n)
        # the product path gathers the sums as one tensor (gloo here, NCCL on the GPUs): same rows, same order
        rows = sharding.gather_rows(np.asarray(part, dtype=np.float64), device=0)
        assert rows.shape == (world, 4 * n) and np.array_equal(rows[rank], np.asarray(part, dtype=np.float64))
        pop2, _, ent2, _ = sharding.combine_measurements(list(rows), n)
        assert np.array_equal(pop2, pop) and np.array_equal(ent2, ent)
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


@pytest.mark.parametrize("world,n,d,lo,hi", [(2, 18, 2, 2, 4), (2, 20, 1, 1, 2), (4, 19, 2, 2, 4), (4, 22, 1, 1, 2),
                                             (8, 18, 1, 1, 2), (8, 20, 2, 2, 4), (8, 22, 1, 1, 2), (8, 26, 2, 2, 4),
                                             (2, 30, 2, 2, 4), (4, 30, 2, 2, 4), (8, 30, 2, 2, 4), (8, 33, 2, 2, 4)])
@pytest.mark.parametrize("rd", [4, 6])
def test_fast_kernel_remote_slot_bookkeeping(world, n, d, lo, hi, rd):
    """Statement-by-statement model of the remote operand slots of the fast tile-pass kernel
    (pass_model.remote_slot_trace): warp votes reproduce every lane's own decision, activity is constant
    over a bulk-copy piece, every ring slot holds the row that is read from it, every mbarrier wait uses
    the parity of the phase that row completed, and over the passes every amplitude receives every
    remote term exactly where the plan says."""
    rules = RuleNS(n, d, lo, hi)
    rbits = world.bit_length() - 1
    nl = n - rbits
    passes = _lib.plan_passes(nl)
    assert nl >= 13 and len(passes) >= 2
    rng = np.random.default_rng(n + world)
    for rank in sorted(set([0, world - 1, int(rng.integers(world))])):
        ops = _lib.plan_remote(rules, world, rank)
        rot = _lib.plan_rotation(rules, world, rank)
        if not ops or rot["nslots"] == 0:
            continue
        counts = {}
        for p, ps in enumerate(passes):
            ntiles = 1 << (nl - 13)
            tiles = sorted(set([0, ntiles - 1] + [int(t) for t in rng.integers(ntiles, size=3)]))
            for tile in tiles:
                applied = pass_model.remote_slot_trace(ps, tile, ops, rot, p, rd)
                L, H0 = ps["low_bits"], ps["high_start"]
                for k, table in enumerate(applied):
                    for (tid, e), j in table.items():
                        assert j >= 0
                        gap = H0 - L
                        base = ((tile & ((1 << gap) - 1)) << L) | ((tile >> gap) << (H0 + 13 - L))
                        y = (tid << 1) | (e << 9)
                        x = base | (y & ((1 << L) - 1)) | ((y >> L) << H0)
                        op = ops[j]
                        assert (op["mask"] >> ((x >> op["shift"]) & 15)) & 1          # the term is active at x
                        r = (rot["rot_word"] >> (2 * ((x >> rot["rot_shift"]) & 15))) & 3
                        assert rot["op_of"][p, k, r] == j                              # and this pass/slot carries it
                        counts[(x, j)] = counts.get((x, j), 0) + 1
        assert all(c == 1 for c in counts.values())
