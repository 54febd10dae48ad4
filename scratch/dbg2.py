import sys; sys.path.insert(0,'oracle'); sys.path.insert(0,'feastkit.jl_b200'); sys.path.insert(0,'.')
import numpy as np, feast_oracle as fo, feast_port as fp
import feastcuda as fc
N=14
A=fo.laplacian_3d(N).astype(float); ev=fo.laplacian_3d_eigs(N)
Emin,Emax=0.0,0.5*(ev[9]+ev[10]); M0=24
Q0=fo.seeded_subspace(N**3,M0,complex_storage=False).astype(complex)
fpm=fo.feastinit(); fo.feastdefault(fpm)
Z,W=fo.feast_contour(Emin,Emax,fpm)
eng=fc.default_engine(0)
eng.set_sparse(fc.A, A.tocsc(), fc.SYM); eng.clear_b()
for e in [0,6,7]:
    for kw in [dict(solver_tol=1e-12, solver_maxiter=4000, inner_rel=1e-9), dict(solver_tol=1e-9, solver_maxiter=4000), dict(solver_tol=1e-12, solver_maxiter=4000, inner_rel=1e-9, solver_restart=0)]:
        X,its,res=eng.block_solve(Z[e],Q0,**kw)
        true=np.linalg.norm(Q0-fp.shifted_apply(A,None,Z[e],X),axis=0)
        print(e,kw,"its",its.min(),its.max(),"res",res.max(),"true",true.max(), "krylov_iters stat", eng.stats()["krylov_iters"])
        eng.reset_stats()
X,its,res=eng.block_solve(Z[7],Q0[:,:1],solver_tol=1e-9, solver_maxiter=4000)
print("m=1", its, res)
