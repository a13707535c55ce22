"""The drop-in boundary fed with the reference's OWN objects (build container only: the reference does not travel
to the GPU box, where these tests skip).  Each case runs tests/ref_objects_worker.py in a fresh process with one of
the BASELINE command lines: the reference's ``Parser`` against ``Args.from_argv``, its ``MPO`` / ``MPS`` / named
states through the host side of ``Exact`` / ``TDVP`` up to the first device call."""
import json
import os
import subprocess
import sys

import pytest

REFERENCE = os.environ.get("QCA_REFERENCE_ROOT", "/root/reference")
WORKER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_objects_worker.py")

CASES = {
    "configs0": ["--initial-states", "single", "--num-cells", "9", "--num-steps", "100"],
    "configs1": ["--initial-states", "blinker", "--num-cells", "12", "--num-steps", "1000"],
    "configs2": ["--algorithm", "2tdvp", "--initial-states", "single", "--num-cells", "15", "--num-steps", "1000",
                 "--plotting-frequency", "10", "--plot-bond-dims"],
    "configs3_rule": ["--initial-states", "triple_blinker", "gradient", "--num-cells", "11", "--distance", "2",
                      "--activation-interval", "2", "4"],
    "configs4_rule": ["--algorithm", "2tdvp", "--initial-states", "blinker", "--num-cells", "16", "--max-bond-dim", "256"],
    "defaults": [],
    "tdvp1_all_states": ["--algorithm", "1tdvp", "--num-cells", "10", "--step-size", "0.01", "--svd-epsilon", "1e-6",
                         "--initial-states", "full_blinker", "single_bottom", "all_ket_0", "all_ket_1", "only_outer",
                         "all_ket_1_but_outer", "equal_superposition", "equal_superposition_but_outer"],
}


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "tensor_networks")), reason="reference not present")
@pytest.mark.parametrize("case", sorted(CASES))
def test_reference_objects_pass_the_boundary(case):
    run = subprocess.run([sys.executable, WORKER, REFERENCE] + CASES[case], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stderr[-2000:]
    checked = json.loads(run.stdout.strip().splitlines()[-1])
    assert checked["parser_fields"] == 16 and checked["mpo_tensors"] >= 9
    assert checked["constructors"]
