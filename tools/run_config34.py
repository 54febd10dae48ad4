"""BASELINE configs[3] (complex Hermitian FEM stiffness/mass pair, zfeast_hcsrgv!) and configs[4] (general complex
Toeplitz-Kronecker pencil, pzifeast_gcsrgv!) at a chosen grid; checks the analytic spectrum and residuals.
usage: python tools/run_config34.py {3|4} nx ny nz M0 [maxiter] [inner_rel]"""
import sys, time
sys.path.insert(0, 'oracle'); sys.path.insert(0, 'feastkit.jl_b200'); sys.path.insert(0, 'tests')
import numpy as np, scipy.sparse as sp
import feastcuda as fc

which = int(sys.argv[1]); nx, ny, nz, M0 = map(int, sys.argv[2:6])
maxiter = int(sys.argv[6]) if len(sys.argv) > 6 else 2000
inner_rel = float(sys.argv[7]) if len(sys.argv) > 7 else 1e-3
rng = np.random.default_rng(12345)
if which == 3:
    from test_gpu_configs import _fem_pair
    t = time.time(); A, B, w = _fem_pair(nx, ny, nz); n = A.shape[0]; print("build", time.time() - t, "n", n, "nnz", A.nnz, flush=True)
    want = int(sys.argv[8]) if len(sys.argv) > 8 else 60
    while w[want] - w[want - 1] < 1e-8 * w[want]: want += 1
    Emin, Emax = 0.0, 0.5 * (w[want - 1] + w[want])
    Q0 = rng.standard_normal((n, M0)) + 0j; Q0 /= np.linalg.norm(Q0, axis=0)
    fpm = fc.feastinit(); fpm[1] = 16; fpm[3] = 40
    t = time.time()
    import os
    r = fc.zfeast_hcsrgv(A, B, Emin, Emax, M0, fpm, Q0=Q0, solver_maxiter=maxiter, ritz_guess=True, inner_rel=inner_rel,
                         b_delta=float(os.environ.get("B_DELTA", "0")))
    dt = time.time() - t
    print("config3 time", dt, "info", r.info, "M", r.M, "want", want, "loops", r.loop, "epsout", r.epsout)
    if r.M == want: print("eig err rel", np.abs(np.sort(r.lambda_) - w[:want]).max() / w[want])
else:
    dims = (nx, ny, nz)
    coef = [(0.4 + 0.1j, 1.0 + 0.05j, 0.9 - 0.05j), (0.3 - 0.1j, 0.8 + 0.1j, 0.75 + 0.05j), (0.5 + 0.2j, 0.6 - 0.05j, 0.65 + 0.02j)]
    eps = 0.05
    toe = lambda n, a, b, c: sp.diags([b * np.ones(n - 1), a * np.ones(n), c * np.ones(n - 1)], [-1, 0, 1])
    I = [sp.identity(n) for n in dims]; T = [toe(n, *abc) for n, abc in zip(dims, coef)]; S = [toe(n, 0.0, abc[1], abc[2]) for n, abc in zip(dims, coef)]
    ksum = lambda X: sp.kron(sp.kron(X[0], I[1]), I[2]) + sp.kron(sp.kron(I[0], X[1]), I[2]) + sp.kron(sp.kron(I[0], I[1]), X[2])
    t = time.time(); A = ksum(T).tocsc(); B = (sp.identity(A.shape[0]) + eps * ksum(S)).tocsc(); n = A.shape[0]; print("build", time.time() - t, "n", n, flush=True)
    la, ls = [], []
    for nn, (a, b, c) in zip(dims, coef):
        th = np.arange(1, nn + 1) * np.pi / (nn + 1); la.append(a + 2 * np.sqrt(b * c) * np.cos(th)); ls.append(2 * np.sqrt(b * c) * np.cos(th))
    lamA = (la[0][:, None, None] + la[1][None, :, None] + la[2][None, None, :]).ravel(); lamS = (ls[0][:, None, None] + ls[1][None, :, None] + ls[2][None, None, :]).ravel()
    lam = lamA / (1 + eps * lamS)
    # a disc at the left edge of the spectrum holding ~35 eigenvalues
    order = np.argsort(lam.real); edge = lam[order[0]]
    Emid = complex(edge.real, lam[order[:40]].imag.mean())
    dist = np.abs(lam - Emid); rad = 0.5 * (np.sort(dist)[34] + np.sort(dist)[35])
    inside = lam[dist <= rad]; print("center", Emid, "radius", rad, "inside", len(inside), "margin", np.abs(dist - rad).min(), flush=True)
    Q0 = rng.standard_normal((n, M0)) + 0j; Q0 /= np.linalg.norm(Q0, axis=0)
    fpm = fc.feastinit(); fpm[7] = 24; fpm[2] = 10; fpm[3] = 30
    t = time.time()
    r = fc.pzifeast_gcsrgv(A, B, Emid, rad, M0, fpm, Q0=Q0, solver_maxiter=maxiter, inner_rel=inner_rel)
    dt = time.time() - t
    print("config4 time", dt, "info", r.info, "M", r.M, "inside", len(inside), "loops", r.loop, "epsout", r.epsout)
    if r.M == len(inside): print("eig err", max(min(abs(g - x) for x in inside) for g in r.lambda_))
print({k: v for k, v in r.stats.items() if k in ("ms_total", "ms_solve", "krylov_iters", "spmm_launches", "node_solves", "cheb_degree", "lz_steps_p1")})
for i, nm in enumerate(fc._lib.KERN_NAMES):
    if r.stats["n_kern"][i]:
        avg = r.stats["ms_kern"][i] / r.stats["n_kern"][i]
        print(f"  {nm}: avg {avg:.4f} ms, {r.stats['bytes_kern'][i] / avg / 1e6:.0f} GB/s algorithmic, {r.stats['n_kern'][i]} samples")
