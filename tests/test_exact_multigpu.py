"""Multi-GPU parity (needs >= 2 GPUs on the box; run with `gpurun --gpus N -- python -m pytest
tests -m gpu`).  One torchrun job per world size; the checks live in tests/mgpu_worker.py."""
import os
import socket
import subprocess
import sys

import pytest

import qca_b200
from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_exact_matches_single_gpu(world):
    if qca_b200.lib.qca_device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]


def test_two_slot_kernels_on_two_gpus():
    """The tile-pass kernels with two remote operand slots are normally only used with three sharded
    qubits (8 ranks) on small registers; QCA_FORCE_SLOTS2 runs them wherever one slot would do, so that
    two GPUs suffice to cover them."""
    if qca_b200.lib.qca_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, QCA_FORCE_SLOTS2="1"))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
