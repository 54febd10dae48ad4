import sys, time, os; sys.path.insert(0,'oracle'); sys.path.insert(0,'feastkit.jl_b200'); sys.path.insert(0,'.')
import numpy as np, feast_oracle as fo
import feastcuda as fc
N=int(sys.argv[1]); M0=int(sys.argv[2]); rel=float(sys.argv[3]); kmax=int(sys.argv[4]); maxloop=int(sys.argv[5])
rel0=float(sys.argv[6]); kmax0=int(sys.argv[7]); keep=int(sys.argv[8]); reps=int(sys.argv[9]) if len(sys.argv)>9 else 1
t=time.time(); A=fo.laplacian_3d(N).astype(float).tocsr(); print("build A", time.time()-t, A.nnz)
ev=fo.laplacian_3d_eigs(N)[:80]
Emin,Emax=0.0,0.5*(ev[34]+ev[35])
print("interval",Emin,Emax)
Q0=fo.seeded_subspace(N**3,M0,complex_storage=False)
eng=fc.default_engine(0)
eng.set_sparse(fc.A, A, fc.SYM); eng.clear_b()
fpm=fc.feastinit(); fpm[3]=maxloop
fc.feastdefault_(fpm)
Z,W=fc.feast_contour(Emin,Emax,fpm)
for rep in range(reps):
    eng.reset_stats()
    t=time.time()
    r=eng.solve_interval(Emin,Emax,M0,list(fpm),Z,W,Q0=Q0,x_real=True,filter="true",inner_rel=rel,ritz_guess=True,solver_tol=1e-13,solver_maxiter=kmax,solver_restart=0,inner_rel0=rel0,maxiter0=kmax0,keep_going=bool(keep))
    dt=time.time()-t
    st=r.stats
    print("rep",rep,"time",dt,"info",r.info,"M",r.M,"loops",r.loop,"epsout",r.epsout)
    print({k:v for k,v in st.items() if k!="node_iters"}); print("node_iters",st["node_iters"][:8])
    if r.M: print("eig err", np.abs(np.sort(r.lambda_)-ev[:r.M]).max(), "res max", r.res.max())
    if st["spmm_sampled"]: 
        ms=st["ms_spmm_sampled"]/st["spmm_sampled"]; print("spmm ms", ms, "alg GB/s", st["bytes_spmm_alg"]/ms/1e6)
