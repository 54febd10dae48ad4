import sys; sys.path.insert(0,'oracle')
import numpy as np, feast_oracle as fo, feast_port as fp
N=14
A=fo.laplacian_3d(N).astype(float); ev=fo.laplacian_3d_eigs(N)
Emin,Emax=0.0,0.5*(ev[9]+ev[10]); M0=24
Q0=fo.seeded_subspace(N**3,M0,complex_storage=False)
fpm=fo.feastinit(); fo.feastdefault(fpm)
Z,W=fo.feast_contour(Emin,Emax,fpm)
for e in [0,6,7]:
    X,its,true,ok=fp.block_bicgstab(A,None,Z[e],Q0.astype(complex),None,rtol=1e-12,maxiter=4000,rel_to_initial=1e-9)
    print(e,Z[e],its, true.max(), ok.all())
