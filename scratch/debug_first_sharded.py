import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["QCA_DEBUG"] = "1"
import qca_b200
from qca_b200 import _lib
rules = qca_b200.Rules(17, range(2, 4), 2)
flags = int(sys.argv[1]) if len(sys.argv) > 1 else _lib.QCA_FLAG_LOOSE_BOUND
eng = _lib.ExactEngine(rules, world_size=2, rank=1, flags=flags)
print("created", flush=True)
eng.loopback_peers()
eng.set_product_state(qca_b200.states.plist("triple_blinker", rules))
print("state set", flush=True)
eng.resolve_planes(*eng.plane_flags())
eng.step(1.0, 1)
print("stepped", eng.norm2(), flush=True)
