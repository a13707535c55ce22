import torch, time, numpy as np
torch.manual_seed(0)
def bench(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t=time.time()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.time()-t)/reps*1e3
for n in (128, 256, 512):
    # a realistic spectrum: exponentially decaying singular values
    u,_ = torch.linalg.qr(torch.randn(n,n,dtype=torch.complex128,device='cuda'))
    v,_ = torch.linalg.qr(torch.randn(n,n,dtype=torch.complex128,device='cuda'))
    s = torch.exp(-torch.arange(n,device='cuda',dtype=torch.float64)*(18.0/n))
    m = (u*s.to(torch.complex128))@v.conj().T
    res={}
    for drv in (None,'gesvd','gesvdj'):
        try: res[str(drv)] = bench(lambda: torch.linalg.svd(m, full_matrices=False, driver=drv))
        except Exception as e: res[str(drv)] = str(e)[:40]
    res['eigh'] = bench(lambda: torch.linalg.eigh(m.conj().T@m))
    res['qr'] = bench(lambda: torch.linalg.qr(m))
    res['cpu_svd'] = None
    mc = m.cpu(); t=time.time(); torch.linalg.svd(mc, full_matrices=False); res['cpu_svd']=(time.time()-t)*1e3
    # accuracy of eigh-based SVD
    lam, V = torch.linalg.eigh(m.conj().T@m)
    lam = lam.flip(0); V = V.flip(1)
    B = m@V; sig = torch.linalg.vector_norm(B, dim=0)
    print(n, {k:(round(x,2) if isinstance(x,float) else x) for k,x in res.items()}, 'eigh-svd sigma rel err (top 8, mid, last):', ((sig-s).abs()/s)[[0,n//4,n//2,n-1]].cpu().numpy())
