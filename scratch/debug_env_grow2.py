import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qca_b200
from qca_b200.linalg import SiteOperator, env_grow
torch.set_printoptions(precision=4, linewidth=200)
rules = qca_b200.Rules(8, range(1, 2), 1)
H = qca_b200.MPO.hamiltonian_from_rules(rules)
W = [np.asarray(w) for w in H.W]
gen = torch.Generator(device="cuda").manual_seed(3)
def rnd(*s): return torch.randn(*s, dtype=torch.complex128, device="cuda", generator=gen)
for (dl, dr) in ((1, 2), (1, 1), (2, 1), (2, 2), (1, 3), (3, 2)):
    # right env at the last site
    a = rnd(2, dr, dl); prev = rnd(dl, 1, dl); w = torch.as_tensor(W[7], device="cuda")
    t = torch.einsum("uwv,alu->awvl", prev, a); t = torch.einsum("abmw,awvl->bmvl", w, t); want = torch.einsum("bmvl,bkv->lmk", t, a.conj())
    got = env_grow(prev, a.transpose(1, 2).contiguous(), SiteOperator(np.ascontiguousarray(W[7].transpose(0, 1, 3, 2)), device="cuda"))
    print("R last", dl, dr, (got - want).abs().max().item(), "vs conj", (got - want.conj()).abs().max().item(), "vs T", (got - want.permute(2, 1, 0)).abs().max().item())
    if dl == 1 and dr == 2:
        print(got.reshape(-1)); print(want.reshape(-1))
    # left env at the first site
    a = rnd(2, dl, dr); prev = rnd(dl, 1, dl); w = torch.as_tensor(W[0], device="cuda")
    t = torch.einsum("xwy,axr->awyr", prev, a); t = torch.einsum("abwm,awyr->bmyr", w, t); want = torch.einsum("bmyr,bys->rms", t, a.conj())
    got = env_grow(prev, a, SiteOperator(W[0], device="cuda"))
    print("L first", dl, dr, (got - want).abs().max().item())
