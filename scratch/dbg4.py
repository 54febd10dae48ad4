import sys; sys.path.insert(0,'oracle')
import numpy as np, feast_oracle as fo
for N,M0,k in [(8,40,10),(10,48,10),(8,32,4),(9,64,10)]:
    A=fo.laplacian_3d(N).astype(float).tocsc(); ev=fo.laplacian_3d_eigs(N)
    Emin,Emax=0.0,0.5*(ev[k-1]+ev[k])
    Q0=fo.seeded_subspace(N**3,M0,complex_storage=False)
    fpm=fo.feastinit(); fpm[3]=60
    ro=fo.feast_scsrev(A,Emin,Emax,M0,list(fpm),Q0=Q0.astype(complex),filter="reference")
    print(N,M0,k,"->",ro.info, ro.M, ro.loop, ro.epsout)
