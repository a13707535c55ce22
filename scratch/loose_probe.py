"""Which flow makes the first launch fail with 'invalid argument'?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qca_b200
from qca_b200 import _lib
case = sys.argv[1]
rules = qca_b200.Rules(16, range(2, 4), 2)
plist = qca_b200.states.plist("blinker", rules)
try:
    if case == "loose1":
        e = _lib.ExactEngine(rules, device=0, flags=_lib.QCA_FLAG_LOOSE_BOUND)
        e.set_product_state(plist)
    elif case == "loose2":
        e = _lib.ExactEngine(rules, device=0, world_size=2, rank=1, flags=_lib.QCA_FLAG_LOOSE_BOUND)
        e.loopback_peers(); e.set_product_state(plist)
    elif case == "tight2":
        e = _lib.ExactEngine(rules, device=0, world_size=2, rank=1)
        e.loopback_peers(); e.set_product_state(plist)
    elif case == "loose2_sync":
        import torch
        torch.zeros(1, device="cuda")
        e = _lib.ExactEngine(rules, device=0, world_size=2, rank=1, flags=_lib.QCA_FLAG_LOOSE_BOUND)
        e.loopback_peers(); e.set_product_state(plist)
    print(case, "ok")
except Exception as exc:
    print(case, "FAILED", exc)
