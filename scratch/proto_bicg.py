import sys, time; sys.path.insert(0,'oracle')
import numpy as np, scipy.sparse as sp
import feast_oracle as fo

def bicgstab_block(A, z, Bm, X0, tol, maxiter, log=None):
    n, m = Bm.shape
    S = lambda X: z*X - A@X
    X = X0.copy()
    R = Bm - S(X)
    Rh = R.copy()
    bn = np.linalg.norm(Bm, axis=0)
    rho = np.sum(Rh.conj()*R, axis=0)
    P = R.copy()
    active = np.ones(m, bool)
    its = np.zeros(m, int)
    for it in range(maxiter):
        V = S(P)
        den = np.sum(Rh.conj()*V, axis=0)
        alpha = np.where(active, rho/den, 0)
        Sv = R - alpha*V
        T = S(Sv)
        tt = np.sum(np.abs(T)**2, axis=0)
        ts = np.sum(T.conj()*Sv, axis=0)
        omega = np.where(active, ts/tt, 0)
        rht = np.sum(Rh.conj()*T, axis=0)
        rho_new = -omega*rht
        beta = np.where(active, (rho_new/rho)*(alpha/np.where(omega==0,1,omega)), 0)
        X += alpha*P + omega*Sv
        R = Sv - omega*T
        P = R + beta*(P - omega*V)
        rho = np.where(active, rho_new, rho)
        rn = np.linalg.norm(R, axis=0)
        its[active] += 1
        active &= ~(rn <= tol*(1+bn))
        if log is not None and it % 50 == 0: log.append((it, rn.max(), active.sum()))
        if not active.any(): break
    true = np.linalg.norm(Bm - S(X), axis=0)
    return X, its, true

if __name__ == "__main__":
    N = int(sys.argv[1]); m = int(sys.argv[2])
    A = fo.laplacian_3d(N).astype(float)
    ev = fo.laplacian_3d_eigs(N)
    # interval containing first 35 eigenvalues (scaled analog of C3)
    Emax = 0.5*(ev[34]+ev[35]); Emin = 0.0
    print("n", N**3, "interval", Emin, Emax, "ev35/36", ev[34], ev[35], "ev64,65", ev[63], ev[64])
    fpm = fo.feastinit(); fo.feastdefault(fpm)
    Z, W = fo.feast_contour(Emin, Emax, fpm)
    Q0 = fo.seeded_subspace(N**3, m)
    for e in [0, 1, 2, 3, 7]:
        log = []
        t = time.time()
        X, its, true = bicgstab_block(A, Z[e], Q0, np.zeros_like(Q0), float(sys.argv[3]), 4000, log)
        print("node", e, Z[e], "iters max/mean", its.max(), its.mean(), "true res max", true.max(), "time", time.time()-t)
