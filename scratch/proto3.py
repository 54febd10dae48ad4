import sys, time; sys.path.insert(0,'oracle')
import numpy as np
import feast_oracle as fo, feast_port as fp
N=int(sys.argv[1]); m=int(sys.argv[2]); rel=float(sys.argv[3]); kmax=int(sys.argv[4]); maxloop=int(sys.argv[5])
A = fo.laplacian_3d(N).astype(float); ev = fo.laplacian_3d_eigs(N)
Emax = 0.5*(ev[34]+ev[35]); Emin=0.0
fpm = fo.feastinit(); fpm[3]=maxloop
Q0 = fo.seeded_subspace(N**3, m, complex_storage=False)
t=time.time()
r = fp.feast_hrr_bicgstab(A, None, Emin, Emax, m, fpm, Q0, inner_rtol=1e-13, inner_rel=rel, inner_maxiter=kmax, verbose=True)
print("RESULT rel",rel,"kmax",kmax,"info",r.info,"M",r.M,"loops",r.loop,"epsout",r.epsout,"col_iters",r.stats["col_iters"],"sum node iters",sum(sum(x) for x in r.stats["node_iters"]),"time",time.time()-t)
print("eig err", np.abs(np.sort(r.lambda_)-ev[:r.M]).max())
