import sys, time, numpy as np, torch
ROOT='/root/repo'
for p in (ROOT,): sys.path.insert(0,p)
import qca_b200
def random_mps(n, chi, seed=0):
    rng = np.random.default_rng(seed)
    dims = [min(2**i, 2**(n-i), chi) for i in range(n+1)]
    return qca_b200.MPS([ (rng.standard_normal((2,dims[i],dims[i+1])) + 1j*rng.standard_normal((2,dims[i],dims[i+1])))/np.sqrt(2*dims[i]) for i in range(n)])
for (n, chi) in [(32,64),(32,128),(64,256)]:
    rules = qca_b200.Rules(n, range(1,2), 1)
    args = qca_b200.Args(rules=rules, step_size=0.005, algorithm='2tdvp', max_bond_dim=chi, svd_epsilon=1e-14)
    t0=time.time()
    algo = qca_b200.TDVP(random_mps(n, chi), qca_b200.MPO.hamiltonian_from_rules(rules), args)
    torch.cuda.synchronize(); t_init=time.time()-t0
    algo.do_time_step(); torch.cuda.synchronize()
    # phase timers
    import qca_b200.algorithms.tdvp as T
    timers = {'svd':0.0,'expm':0.0,'qr':0.0}
    def timed(name, fn):
        def w(*a, **k):
            torch.cuda.synchronize(); t=time.time(); r=fn(*a, **k); torch.cuda.synchronize(); timers[name]+=time.time()-t; return r
        return w
    orig_svd = torch.linalg.svd; torch.linalg.svd = timed('svd', orig_svd)
    algo._expm_apply = timed('expm', algo._expm_apply)
    algo._left_qr = timed('qr', algo._left_qr)
    algo.heff_applications = 0
    torch.cuda.synchronize(); t0=time.time()
    algo.do_time_step(); torch.cuda.synchronize(); dt=time.time()-t0
    torch.linalg.svd = orig_svd
    print(f"N={n} chi={chi}: init {t_init:.2f}s, step {dt:.3f}s ({2/dt:.3f} sweeps/s), heff applications {algo.heff_applications}, timers {dict((k,round(v,3)) for k,v in timers.items())}, bonds max {max(a.shape[1] for a in algo._A)}, krylov m {T.krylov_dimension(abs(np.pi/2*0.0025)*algo._bound)}")
