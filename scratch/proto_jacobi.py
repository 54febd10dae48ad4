import numpy as np
rng=np.random.default_rng(0)
def jacobi_herm(C, sweeps=30):
    r=C.shape[0]; C=C.copy(); V=np.eye(r,dtype=complex)
    rp = r + (r&1)
    for sw in range(sweeps):
        off = np.sqrt(np.sum(np.abs(C)**2)-np.sum(np.abs(np.diag(C))**2)); nrm=np.linalg.norm(C)
        if off <= 1e-15*nrm: break
        # round robin
        idx=list(range(rp))
        for step in range(rp-1):
            pairs=[(idx[i], idx[rp-1-i]) for i in range(rp//2)]
            rots=[]
            for (p,q) in pairs:
                if p>q: p,q=q,p
                if q>=r: rots.append(None); continue
                g=C[p,q]; ag=abs(g)
                if ag <= 1e-300 or ag <= 1e-18*np.sqrt(abs(C[p,p].real*C[q,q].real)) and False:
                    rots.append(None); continue
                a=C[p,p].real; b=C[q,q].real
                ph=g/ag
                th=(b-a)/(2*ag)
                t=(1.0 if th>=0 else -1.0)/(abs(th)+np.sqrt(th*th+1))
                c=1/np.sqrt(t*t+1); s=t*c
                rots.append((p,q,c,s,ph))
            # rows: C <- J^H C
            Cn=C.copy()
            for rt in rots:
                if rt is None: continue
                p,q,c,s,ph=rt
                Cn[p,:]=c*C[p,:]-s*ph*C[q,:]
                Cn[q,:]=s*C[p,:]+c*ph*C[q,:]     # J^H row q: conj(J[:,q]) = [s, c*conj(e^{-i phi})] = [s, c e^{i phi}]
            C=Cn; Cn=C.copy(); Vn=V.copy()
            for rt in rots:
                if rt is None: continue
                p,q,c,s,ph=rt
                Cn[:,p]=c*C[:,p]-s*np.conj(ph)*C[:,q]
                Cn[:,q]=s*C[:,p]+c*np.conj(ph)*C[:,q]
                Vn[:,p]=c*V[:,p]-s*np.conj(ph)*V[:,q]
                Vn[:,q]=s*V[:,p]+c*np.conj(ph)*V[:,q]
            C=Cn; V=Vn
            idx=[idx[0]]+[idx[-1]]+idx[1:-1]
    return np.diag(C).real, V, sw
for r in [1,2,5,8,33,64]:
    M=rng.standard_normal((r,r))+1j*rng.standard_normal((r,r)); M=(M+M.conj().T)/2
    w,V,sw=jacobi_herm(M)
    print(r, sw, np.abs(np.sort(w)-np.linalg.eigvalsh(M)).max(), np.abs(M@V-V*w).max(), np.abs(V.conj().T@V-np.eye(r)).max())
