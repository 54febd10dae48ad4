import sys, time; sys.path.insert(0,'oracle'); sys.path.insert(0,'scratch')
import numpy as np
import feast_oracle as fo
from proto_bicg import bicgstab_block
N=int(sys.argv[1]); m=int(sys.argv[2]); tol=float(sys.argv[3])
A = fo.laplacian_3d(N).astype(float); ev = fo.laplacian_3d_eigs(N)
Emax = 0.5*(ev[34]+ev[35]); Emin=0.0
fpm = fo.feastinit(); fo.feastdefault(fpm)
Z, W = fo.feast_contour(Emin, Emax, fpm)
Q0 = fo.seeded_subspace(N**3, m)
for e in range(8):
    X, its, true = bicgstab_block(A, Z[e], Q0, np.zeros_like(Q0), tol, 6000)
    print(N, "node", e, "iters max/mean", its.max(), its.mean(), "true", true.max(), flush=True)
