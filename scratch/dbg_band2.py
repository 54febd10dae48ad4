import sys; sys.path.insert(0,'oracle'); sys.path.insert(0,'feastkit.jl_b200')
import numpy as np, scipy.linalg as sla, feast_oracle as fo, feastcuda as fc
rng=np.random.default_rng(2)
n,k,M0=600,7,24
Af=np.zeros((n,n)); Bf=np.zeros((n,n))
for dd in range(k+1):
    v=rng.standard_normal(n-dd)*(3.0 if dd==0 else 0.4); Af+=np.diag(v,dd)+(np.diag(v,-dd) if dd else 0)
    if dd<=2:
        u=rng.uniform(0.05,0.1,n-dd) if dd else rng.uniform(1.0,2.0,n); Bf+=np.diag(u,dd)+(np.diag(u,-dd) if dd else 0)
ws=sla.eigh(Af,Bf,eigvals_only=True); lo=250
Emin,Emax=0.5*(ws[lo-1]+ws[lo]),0.5*(ws[lo+9]+ws[lo+10])
print("interval",Emin,Emax, "gap below", ws[lo]-ws[lo-1], "above", ws[lo+10]-ws[lo+9], "24th dist", ws[lo+17]-Emax, Emin-ws[lo-8])
Q0=fo.seeded_subspace(n,M0,complex_storage=False)
print("== band"); r=fc.feast_sbgv(fo.full_to_banded(Af,k),fo.full_to_banded(Bf,2),k,2,Emin,Emax,M0,fc.feastinit(),Q0=Q0)
print(r.info,r.M,r.loop,r.epsout)
print("== dense"); r=fc.feast_sygv(Af,Bf,Emin,Emax,M0,fc.feastinit(),Q0=Q0)
print(r.info,r.M,r.loop,r.epsout)
ro=fo.feast_sygv(Af,Bf,Emin,Emax,M0,fo.feastinit(),Q0=Q0.astype(complex),filter="true")
print("oracle",ro.info,ro.M,ro.loop,ro.epsout)
