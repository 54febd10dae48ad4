import sys; sys.path.insert(0,'oracle')
import numpy as np, feast_oracle as fo
N=14
A=fo.laplacian_3d(N).astype(float).tocsc(); ev=fo.laplacian_3d_eigs(N)
Emin,Emax=0.0,0.5*(ev[9]+ev[10]); M0=24
Q0=fo.seeded_subspace(N**3,M0,complex_storage=False)
fpm=fo.feastinit(); fpm[3]=60
ro=fo.feast_scsrev(A,Emin,Emax,M0,list(fpm),Q0=Q0.astype(complex),filter="reference")
print(ro.info, ro.M, ro.loop, ro.epsout)
