import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
    try:
        import qca_b200
        return qca_b200.lib.qca_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name: str):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return json.loads(str(g["spec"])), g


def golden_names(kind: str) -> list[str]:
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        spec = json.loads(str(np.load(f)["spec"]))
        if spec["kind"] == kind:
            out.append(spec["name"])
    return out


class RuleNS:
    """Minimal stand-in with the attributes of parameters.Rules."""

    def __init__(self, ncells, distance, lo, hi):
        self.ncells, self.distance, self.activation_interval, self.periodic = ncells, distance, range(lo, hi), False
