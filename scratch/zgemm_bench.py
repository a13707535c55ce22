import sys, time, torch
sys.path.insert(0,'/root/repo')
from qca_b200.linalg import env_times_tensor, tensor_times_env
def bench(fn, reps=10):
    fn(); torch.cuda.synchronize(); t=time.time()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.time()-t)/reps*1e3
for chi in (64,128,256):
    w,g=6,4
    left=torch.randn(chi,w,chi,dtype=torch.complex128,device='cuda'); theta=torch.randn(g,chi,chi,dtype=torch.complex128,device='cuda')
    t=torch.randn(g,w,chi,chi,dtype=torch.complex128,device='cuda'); right=torch.randn(chi,w,chi,dtype=torch.complex128,device='cuda')
    fl = 8.0*g*w*chi**3
    a=bench(lambda: env_times_tensor(left,theta)); b=bench(lambda: torch.einsum('xwy,gxu->gwyu',left,theta))
    c=bench(lambda: tensor_times_env(t,right)); d=bench(lambda: torch.einsum('gnyu,unv->gyv',t,right))
    print(f"chi={chi}: L.theta dmma {a:.3f} ms ({fl/a/1e9:.1f} TF/s) einsum {b:.3f} ms ({fl/b/1e9:.1f} TF/s) | T.R dmma {c:.3f} ms ({fl/c/1e9:.1f} TF/s) einsum {d:.3f} ms ({fl/d/1e9:.1f} TF/s)")
