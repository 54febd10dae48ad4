import sys; sys.path.insert(0,'oracle'); sys.path.insert(0,'feastkit.jl_b200')
import numpy as np, scipy.linalg as sla, feast_oracle as fo, feastcuda as fc
eng=fc.default_engine(0)
rng=np.random.default_rng(2)
for (n,k,kb) in ((40,3,0),(40,3,2),(600,7,0),(600,7,2)):
    Af=np.zeros((n,n)); Bf=np.zeros((n,n))
    for dd in range(k+1):
        v=rng.standard_normal(n-dd)*(3.0 if dd==0 else 0.4); Af+=np.diag(v,dd)+(np.diag(v,-dd) if dd else 0)
        if dd<=kb:
            u=rng.uniform(0.05,0.1,n-dd) if dd else rng.uniform(1.0,2.0,n); Bf+=np.diag(u,dd)+(np.diag(u,-dd) if dd else 0)
    if kb==0: Bf=np.eye(n)
    eng.set_band(fc.A, fo.full_to_banded(Af,k), k, fc.SYM)
    if kb: eng.set_band(fc.B, fo.full_to_banded(Bf,kb), kb, fc.SYM)
    else: eng.clear_b()
    z=0.3+0.05j; RHS=rng.standard_normal((n,9))+1j*rng.standard_normal((n,9))
    X,_,_=eng.block_solve(z,RHS,solver="direct")
    W=np.linalg.solve(z*Bf-Af,RHS)
    Y=eng.apply(fc.A,RHS)
    print(n,k,kb,"solve err",np.abs(X-W).max()/np.abs(W).max(),"apply err",np.abs(Y-Af@RHS).max())
    if kb:
        Yb=eng.apply(fc.B,RHS); print("   applyB err",np.abs(Yb-Bf@RHS).max())
