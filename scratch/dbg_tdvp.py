import sys, numpy as np, torch
ROOT='/root/repo'
for p in (ROOT, ROOT+'/oracle', ROOT+'/tests'): sys.path.insert(0,p)
import qca_b200
from qca_b200.algorithms import tdvp as T
rules = qca_b200.Rules(7, range(1,3), 1)
def make(cpu):
    args = qca_b200.Args(rules=rules, step_size=0.005, algorithm='2tdvp', max_bond_dim=8, svd_epsilon=5e-5)
    if cpu:
        real = torch.device
        torch.device = lambda *a, **k: real('cpu')
    try:
        return qca_b200.TDVP(qca_b200.states.make('equal_superposition', rules), qca_b200.MPO.hamiltonian_from_rules(rules), args)
    finally:
        if cpu: torch.device = real
g, c = make(False), make(True)
print('devices', g.dev, c.dev)
log = {}
def wrap(obj, tag):
    orig_svd = torch.linalg.svd
    def two_site(i, j, _orig=obj._two_site):
        ul, s, vr = _orig(i, j)
        log.setdefault((obj.stepno, i, j, obj.phase), {})[tag] = (s.detach().cpu().numpy().real.copy(), ul.shape[2])
        return ul, s, vr
    obj._two_site = two_site
wrap(g,'gpu'); wrap(c,'cpu')
def vec(o): return o.psi.as_vector()
for step in range(5):
    for o in (g,c):
        o.stepno = step
        o._canonicalize(0)
    print('step',step,'after canon 1-ov', abs(1-abs(np.vdot(vec(g),vec(c)))))
    for o in (g,c): o.phase='R'; o._sweep_right_two_site()
    print('step',step,'after right 1-ov', abs(1-abs(np.vdot(vec(g),vec(c)))), [a.shape[1] for a in g._A],[a.shape[1] for a in c._A])
    for o in (g,c): o.phase='L'; o._sweep_left_two_site()
    print('step',step,'after left  1-ov', abs(1-abs(np.vdot(vec(g),vec(c)))), [a.shape[1] for a in g._A],[a.shape[1] for a in c._A])
for k in sorted(log):
    if k[0] >= 2:
        e = log[k]
        print(k, 'gpu', e['gpu'][1], np.array2string(e['gpu'][0], precision=3), 'cpu', e['cpu'][1], np.array2string(e['cpu'][0], precision=3))
