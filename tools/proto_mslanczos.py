"""Prototype: FEAST with multi-shift two-pass Lanczos inner solves (real symmetric, B=I)."""
import sys, time
sys.path.insert(0, 'oracle')
import numpy as np, scipy.sparse as sp, scipy.linalg as sla
import feast_oracle as fo
import os
USE_F = int(os.environ.get("USE_F", "1"))

def lanczos_T(A, b, kmax, Z, F, tolrel, check_every=8):
    """per-column Lanczos in lock step; returns alpha,beta (k x m), norms, k used.
    stop when for every column, every node: beta_{k+1}|y_k| <= tolrel (relative to ||b||)."""
    n, m = b.shape
    nb = np.linalg.norm(b, axis=0)
    u_prev = np.zeros_like(b); u = b / nb
    alpha = np.zeros((kmax, m)); beta = np.zeros((kmax + 1, m))
    ne = len(Z)
    d = np.zeros((ne, m), complex); g = np.zeros((ne, m), complex)
    k = 0
    for j in range(kmax):
        w = A @ u - beta[j] * u_prev
        a = np.einsum('ij,ij->j', u, w)
        w -= a * u
        bn = np.linalg.norm(w, axis=0)
        alpha[j] = a; beta[j + 1] = bn
        # incremental last component of (zI - T_j)^-1 e1
        for e in range(ne):
            if j == 0:
                d[e] = Z[e] - a; g[e] = 1.0 / d[e]
            else:
                dn = (Z[e] - a) - beta[j] ** 2 / d[e]
                g[e] = beta[j] * g[e] / dn; d[e] = dn
        u_prev = u; u = w / bn
        k = j + 1
        resid = bn[None, :] * np.abs(g) * (np.abs(F) if USE_F else 1.0)
        if (k % check_every == 0) and resid.max() <= tolrel:
            break
    return alpha[:k], beta[:k + 1], nb, k, resid.max()

def coeffs(alpha, beta, nb, Z, W, F):
    k, m = alpha.shape
    C = np.zeros((k, m))
    for c in range(m):
        for e in range(len(Z)):
            # solve (zI - T) y = e1 with banded solver
            ab = np.zeros((3, k), complex)
            ab[1] = Z[e] - alpha[:, c]
            ab[0, 1:] = -beta[1:k, c]; ab[2, :-1] = -beta[1:k, c]
            rhs = np.zeros(k, complex); rhs[0] = 1
            y = sla.solve_banded((1, 1), ab, rhs)
            C[:, c] += np.real(2 * W[e] * F[e, c] * y) * nb[c]
    return C

def apply_V(A, b, alpha, beta, C):
    n, m = b.shape
    nb = np.linalg.norm(b, axis=0)
    u_prev = np.zeros_like(b); u = b / nb
    Q = np.zeros_like(b)
    for j in range(alpha.shape[0]):
        Q += C[j] * u
        w = A @ u - beta[j] * u_prev - alpha[j] * u
        u_prev = u; u = w / beta[j + 1]
    return Q

def feast_ms(A, Emin, Emax, M0, ne=8, tolrel=1e-2, kmax=4000, maxloop=30, tol=1e-12, seed=12345):
    n = A.shape[0]
    fpm = fo.feastinit(); fpm[1] = ne; fo.feastdefault(fpm)
    Z, W = fo.feast_contour(Emin, Emax, fpm)
    Q = fo.seeded_subspace(n, M0, complex_storage=False)
    lam = None; total_k = 0
    for loop in range(maxloop + 1):
        if lam is None:
            b = Q; F = np.ones((ne, Q.shape[1]), complex); base = 0
        else:
            R = A @ Q - Q * lam      # eigen residuals (A q - theta q)
            b = R; F = 1.0 / (Z[:, None] - lam[None, :])
        al, be, nb, k, rmax = lanczos_T(A, b, kmax, Z, F, tolrel)
        total_k += k
        C = coeffs(al, be, nb, Z, W, F)
        Y = apply_V(A, b, al, be, C)
        if lam is not None:
            rho = np.real((2 * W[:, None] / (Z[:, None] - lam[None, :])).sum(0))
            Y = Y + Q * rho
        # RR
        Qo, Rr, P = sla.qr(Y, mode='economic', pivoting=True)
        dR = np.abs(np.diag(Rr)); rank = int((dR > max(1.5e-8, 2.2e-16 * n) * dR[0]).sum())
        Qo = Qo[:, :rank]
        S = Qo.T @ (A @ Qo); S = 0.5 * (S + S.T)
        th, V = np.linalg.eigh(S)
        X = Qo @ V
        inside = (th >= Emin) & (th <= Emax)
        M = inside.sum()
        Rs = A @ X - X * th
        res = np.linalg.norm(Rs, axis=0) / np.maximum(np.abs(th), 1)
        eps = res[inside].max() if M else np.inf
        print(f"loop {loop}: k={k} rmax={rmax:.2e} rank={rank} M={M} epsout={eps:.3e}", flush=True)
        if M and eps <= tol:
            break
        order = np.concatenate([np.where(inside)[0], np.where(~inside)[0]])
        Q = X[:, order]; lam = th[order]
    return th[inside], total_k, loop

if __name__ == "__main__":
    N = int(sys.argv[1]); M0 = int(sys.argv[2]); tolrel = float(sys.argv[3]); kmax = int(sys.argv[4])
    A = fo.laplacian_3d(N).astype(float).tocsr()
    ev = fo.laplacian_3d_eigs(N)
    Emin, Emax = 0.0, 0.5 * (ev[34] + ev[35])
    t = time.time()
    lam, tk, loops = feast_ms(A, Emin, Emax, M0, tolrel=tolrel, kmax=kmax)
    print("N", N, "total Lanczos steps", tk, "loops", loops, "time", time.time() - t, "eig err", np.abs(np.sort(lam) - ev[:len(lam)]).max())
