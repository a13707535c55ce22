#!/usr/bin/env python
"""Worker of tests/test_reference_objects_cpu.py: one process per command line, because the reference keeps its
arguments in a process-wide singleton read at import time (parameters/parser.py:187-193, states.py).

    python tests/ref_objects_worker.py /root/reference --num-cells 11 --distance 2 ...

Imports the UNMODIFIED reference's ``parameters``, ``tensor_networks`` and ``states`` next to ``qca_b200`` and feeds
the reference's own ``Parser`` / ``Rules`` / ``MPO`` / ``MPS`` OBJECTS to the host side of the B200 plug-ins
(everything up to the first device call -- there is no GPU in the build container and no reference on the GPU
box, so this is the one place where the two meet).  Prints one JSON line of what it checked."""
import json
import os
import sys

ref_root = sys.argv[1]
argv = sys.argv[2:]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, ref_root)
sys.argv = ["main.py"] + argv

import numpy as np  # noqa: E402

import qca_b200  # noqa: E402
from qca_b200 import _lib  # noqa: E402
from parameters import Parser, Rules as RefRules  # noqa: E402  (the reference's)
from tensor_networks import MPO as RefMPO, MPS as RefMPS  # noqa: E402
import states as ref_states  # noqa: E402

checked = {}
ref = Parser.instance()
ours = qca_b200.Args.from_argv(argv)

# 1. same flags -> same fields (parser.py:182-212)
for name in ("num_steps", "step_size", "algorithm", "max_bond_dim", "svd_epsilon", "plot_frequency", "plot_step_interval",
             "plot_steps", "approximative_evolution_method", "taylor_steps", "initial_states", "initial_state_files"):
    assert getattr(ref, name) == getattr(ours, name), (name, getattr(ref, name), getattr(ours, name))
for name in ("ncells", "distance", "activation_interval", "periodic"):
    assert getattr(ref.rules, name) == getattr(ours.rules, name), name
checked["parser_fields"] = 16

# 2. the reference's MPO object is recognised as the rule Hamiltonian; another rule's is not
h_ref = RefMPO.hamiltonian_from_rules(ref.rules)
h_ours = qca_b200.MPO.hamiltonian_from_rules(ref.rules)            # our constructor on the REFERENCE's Rules object
assert len(h_ref.W) == len(h_ours.W) and all(np.array_equal(a, b) for a, b in zip(h_ref.W, h_ours.W))
assert h_ours.same_operator_as(qca_b200.MPO(list(h_ref.W)))
other = RefRules(ncells=ref.rules.ncells, activation_interval=range(ref.rules.activation_interval.start,
                                                                   ref.rules.activation_interval.stop + 1),
                 distance=ref.rules.distance, periodic=False)
h_other = RefMPO.hamiltonian_from_rules(other)
assert not h_ours.same_operator_as(qca_b200.MPO(list(h_other.W)))
checked["mpo_tensors"] = len(h_ref.W)

# 3. the reference's named states, as the reference's MPS objects
from qca_b200.algorithms.exact import _product_plist  # noqa: E402
n = ref.rules.ncells
nstates = 0
for name in ref.initial_states:
    psi_ref = getattr(ref_states, name)()
    assert isinstance(psi_ref, RefMPS)
    mine = qca_b200.states.make(name, ref.rules)
    if name != "rand":
        assert np.abs(psi_ref.as_vector() - mine.as_vector()).max() < 1e-15, name
        plist = _product_plist(psi_ref, n)
        assert plist is not None and np.allclose(plist, qca_b200.states.plist(name, ref.rules), atol=1e-15), name
    wrapped = qca_b200.MPS(list(psi_ref.A))                         # our container around the reference's tensors
    assert np.array_equal(wrapped.as_vector(), psi_ref.as_vector())
    nstates += 1
checked["states"] = nstates

# 4. the plug-in constructors take the reference's objects through every host-side check and then reach for the
#    device: without one they fail LOUDLY (no CPU fallback); with one they would simply run
have_gpu = _lib.lib.qca_device_count() > 0
psi0 = getattr(ref_states, (ref.initial_states or ["single"])[0])()
outcome = {}
if ref.algorithm == "exact":
    try:
        qca_b200.Exact(psi0, h_ref, ref)
        outcome["exact"] = "ran"
    except _lib.QcaError as e:
        assert e.code in (_lib.QCA_ERR_CUDA, _lib.QCA_ERR_NOMEM), e
        outcome["exact"] = "QcaError (no device)"
    try:
        qca_b200.Exact(psi0, h_other, ref)
        raise SystemExit("a different operator was accepted")
    except ValueError:
        outcome["exact_other_h"] = "ValueError"
else:
    try:
        qca_b200.TDVP(psi0, h_ref, ref)
        outcome["tdvp"] = "ran"
    except _lib.QcaError as e:
        assert e.code == _lib.QCA_ERR_CUDA, e
        outcome["tdvp"] = "QcaError (no device)"
assert have_gpu or all(v != "ran" for v in outcome.values())
checked["constructors"] = outcome
print(json.dumps(checked))
