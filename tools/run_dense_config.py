"""configs[1] at full size: dense real symmetric n=8192 (Householder-similar to diag(linspace(0,100,n))), M0=128, 8 nodes."""
import sys, time; sys.path.insert(0,'oracle'); sys.path.insert(0,'feastkit.jl_b200')
import numpy as np, feastcuda as fc
n=int(sys.argv[1]) if len(sys.argv)>1 else 8192; M0=128
rng=np.random.default_rng(42); v=rng.standard_normal(n); v/=np.linalg.norm(v); d=np.linspace(0.0,100.0,n); Dv=d*v
t=time.time(); A=np.diag(d)-2*np.outer(v,Dv)-2*np.outer(Dv,v)+4*(v@Dv)*np.outer(v,v); A=0.5*(A+A.T); print("build",time.time()-t)
Emin,Emax=50.0,50.0+80.5*(100.0/(n-1)); inside=d[(d>=Emin)&(d<=Emax)]
Q0=np.random.default_rng(12345).standard_normal((n,M0)); Q0/=np.linalg.norm(Q0,axis=0)
for rep in range(2):
    t=time.time(); r=fc.dfeast_syev(A,Emin,Emax,M0,fc.feastinit(),Q0=Q0); dt=time.time()-t
    st=r.stats
    print("rep",rep,"time",dt,"info",r.info,"M",r.M,len(inside),"loops",r.loop,"epsout",r.epsout,"eig err",np.abs(np.sort(r.lambda_)-inside).max() if r.M==len(inside) else None)
    print({k:st[k] for k in ("ms_total","ms_solve","ms_ortho","ms_project","ms_eig","ms_resid","ms_h2d","ms_d2h","kernel_launches","node_solves")})
