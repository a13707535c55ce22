"""env_grow (DMMA) vs the einsum form inside a real 2TDVP run: where do they differ?"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import qca_b200
from qca_b200.algorithms import tdvp as T

g = np.load("tests/golden/tdvp2_gradient8.npz"); spec = json.loads(str(g["spec"]))
rules = qca_b200.Rules(spec["ncells"], range(spec["lo"], spec["hi"]), spec["distance"])
args = qca_b200.Args(rules=rules, step_size=spec["step_size"], algorithm="2tdvp", max_bond_dim=spec["chi"], svd_epsilon=spec["eps"],
                     num_steps=spec["num_steps"], plot_frequency=spec["plot_freq"])
worst = []
def left_ref(prev, a, w):
    t = torch.einsum("xwy,axr->awyr", prev, a); t = torch.einsum("abwm,awyr->bmyr", w, t)
    return torch.einsum("bmyr,bys->rms", t, a.conj())
def right_ref(prev, a, w):
    t = torch.einsum("uwv,alu->awvl", prev, a); t = torch.einsum("abmw,awvl->bmvl", w, t)
    return torch.einsum("bmvl,bkv->lmk", t, a.conj())
orig_l, orig_r = T.TDVP._grow_left, T.TDVP._grow_right
mode = sys.argv[1] if len(sys.argv) > 1 else "native"
def gl(self, prev, site):
    got = orig_l(self, prev, site); want = left_ref(prev, self._A[site], self._W[site])
    worst.append(("L", site, tuple(prev.shape), tuple(self._A[site].shape), (got - want).abs().max().item(), want.abs().max().item()))
    return want if mode == "einsum" else got
def gr(self, prev, site):
    got = orig_r(self, prev, site); want = right_ref(prev, self._A[site], self._W[site])
    worst.append(("R", site, tuple(prev.shape), tuple(self._A[site].shape), (got - want).abs().max().item(), want.abs().max().item()))
    if worst[-1][4] > 1e-6 and not getattr(self, "_dumped", False):
        self._dumped = True
        torch.set_printoptions(precision=5, linewidth=220)
        a = self._A[site]
        print("MISMATCH site", site, "A shape/strides", a.shape, a.stride(), "offset", a.storage_offset(), "prev", prev.reshape(-1), prev.stride())
        print("A", a.reshape(-1)); print("got", got.reshape(-1)); print("want", want.reshape(-1))
        at = a.transpose(1, 2).contiguous(); print("at strides", at.stride(), at.data_ptr() == a.data_ptr())
        from qca_b200.linalg import env_grow
        again = env_grow(prev, a.transpose(1, 2).contiguous().clone(), self._site_operator("one_mirrored", site))
        print("again (cloned input)", (again - want).abs().max().item())
        again2 = orig_r(self, prev, site)
        print("again2 (same call)", (again2 - want).abs().max().item(), (again2 - got).abs().max().item())
    return want if mode == "einsum" else got
T.TDVP._grow_left, T.TDVP._grow_right = gl, gr
algo = qca_b200.TDVP(qca_b200.states.make(spec["state"], rules), qca_b200.MPO.hamiltonian_from_rules(rules), args)
n = spec["ncells"]
steps = g["population"].shape[0]
pop = np.zeros((steps, n)); sse = np.zeros((steps, n))
for step in range(args.num_steps):
    if step % args.plot_step_interval == 0:
        k = step // args.plot_step_interval
        d, b = np.zeros(n), np.zeros(n + 1)
        algo.measure(pop[k], d, sse[k], b)
    algo.do_time_step()
worst.sort(key=lambda r: -r[4])
print("mode", mode, "calls", len(worst))
for r in worst[:8]:
    print(r)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import tdvp_oracle
po, eo, bo, psi = tdvp_oracle.run_tdvp(spec["state"], n, spec["distance"], spec["lo"], spec["hi"], "2tdvp", spec["step_size"], spec["num_steps"],
                                       int(g["plot_step_interval"]), spec["chi"], spec["eps"], consistent=True)
print("pop diff vs consistent oracle", np.abs(pop - po).max(), "entropy", np.abs(sse - eo).max())
