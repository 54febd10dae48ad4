import sys, time, os; sys.path.insert(0,'oracle'); sys.path.insert(0,'feastkit.jl_b200'); sys.path.insert(0,'.')
import numpy as np, feast_oracle as fo
import feastcuda as fc
N=int(sys.argv[1]); M0=int(sys.argv[2]); rel=float(sys.argv[3]); kmax=int(sys.argv[4]); reps=int(sys.argv[5]) if len(sys.argv)>5 else 1
rel0=float(sys.argv[6]) if len(sys.argv)>6 else 0.0
maxloop=int(sys.argv[7]) if len(sys.argv)>7 else 20
t=time.time(); A=fo.laplacian_3d(N).astype(float).tocsr(); print("build A", time.time()-t, A.nnz)
ev=fo.laplacian_3d_eigs(N)[:80]
Emin,Emax=0.0,0.5*(ev[34]+ev[35])
Q0=fo.seeded_subspace(N**3,M0,complex_storage=False)
eng=fc.default_engine(0)
eng.set_sparse(fc.A, A, fc.SYM); eng.clear_b()
fpm=fc.feastinit(); fpm[3]=maxloop; fc.feastdefault_(fpm)
Z,W=fc.feast_contour(Emin,Emax,fpm)
for rep in range(reps):
    eng.reset_stats()
    t=time.time()
    r=eng.solve_interval(Emin,Emax,M0,list(fpm),Z,W,Q0=Q0,x_real=True,filter="true",solver="mslanczos",inner_rel=rel,inner_rel0=rel0,ritz_guess=True,solver_maxiter=kmax,check_every=16,adaptive=bool(int(os.environ.get("ADAPTIVE","1"))),mixed=bool(int(os.environ.get("MIXED","0"))))
    dt=time.time()-t
    st=r.stats
    print("rep",rep,"time",dt,"info",r.info,"M",r.M,"loops",r.loop,"epsout",r.epsout, "env", {k:v for k,v in os.environ.items() if k.startswith("FEASTCUDA")})
    print({k:v for k,v in st.items() if k not in ("node_iters",)})
    if r.M: print("eig err", np.abs(np.sort(r.lambda_)-ev[:r.M]).max(), "res max", r.res.max())
    for i,name in enumerate(fc._lib.KERN_NAMES):
        if st["n_kern"][i]:
            ms=st["ms_kern"][i]/st["n_kern"][i]; print(f"  kern {name}: {ms:.4f} ms avg over {st['n_kern'][i]}, alg GB/s {st['bytes_kern'][i]/ms/1e6:.1f}")
