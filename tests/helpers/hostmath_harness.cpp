// Test harness: exposes the engine's host-side dense routines (feastkit.jl_b200/csrc/host_math.hpp) through a C ABI so that the
// CPU suite can check them against LAPACK.  Built by tests/test_cpu_hostmath.py with g++ (no CUDA needed); test code only.
#include "../../feastkit.jl_b200/csrc/host_math.hpp"

using namespace feastcuda;

static std::vector<zc> from_colmajor(long n, const double* M) {
  const zc* m = reinterpret_cast<const zc*>(M);
  std::vector<zc> R((size_t)n * n);
  for (long i = 0; i < n; ++i)
    for (long j = 0; j < n; ++j) R[(size_t)i * n + j] = m[(size_t)j * n + i];
  return R;
}
static void to_colmajor(long n, const std::vector<zc>& R, double* M) {
  zc* m = reinterpret_cast<zc*>(M);
  for (long i = 0; i < n; ++i)
    for (long j = 0; j < n; ++j) m[(size_t)j * n + i] = R[(size_t)i * n + j];
}

extern "C" {
// eigenpairs of S v = lambda B v; returns the numerical rank of B, or -1 on failure
int hm_pencil_eig(long n, const double* S, const double* B, double* lambda, double* V) {
  std::vector<zc> lam, Vr;
  int rank = 0;
  if (!host_pencil_eig((int)n, from_colmajor(n, S), from_colmajor(n, B), lam, Vr, &rank)) return -1;
  for (long k = 0; k < n; ++k) { lambda[2 * k] = lam[k].real(); lambda[2 * k + 1] = lam[k].imag(); }
  to_colmajor(n, Vr, V);
  return rank;
}
int hm_complex_eig(long n, const double* A, double* lambda, double* V) {
  std::vector<zc> lam, Vr;
  if (!host_complex_eig((int)n, from_colmajor(n, A), lam, Vr)) return -1;
  for (long k = 0; k < n; ++k) { lambda[2 * k] = lam[k].real(); lambda[2 * k + 1] = lam[k].imag(); }
  to_colmajor(n, Vr, V);
  return 0;
}
// one pass of the Gram-driven orthonormalisation: G (cur x cur, column-major) -> T (cur x nout, column-major, caller allocates
// cur x cur); out3 = {accepted, kept, dropped}
void hm_ortho_pass(long cur, long done, const double* G, double thr_abs, double safe, double* T, long* out3) {
  OrthoPass op = ortho_pass(from_colmajor(cur, G), (int)cur, (int)done, thr_abs, safe);
  const long nout = op.accepted + op.kept;
  zc* t = reinterpret_cast<zc*>(T);
  for (long i = 0; i < cur; ++i)
    for (long j = 0; j < nout; ++j) t[(size_t)j * cur + i] = op.T[(size_t)i * nout + j];
  out3[0] = op.accepted; out3[1] = op.kept; out3[2] = op.dropped;
}
// X = M^-1 Bm (n x n right-hand sides), 0 on success
int hm_lu_solve(long n, const double* M, const double* Bm, double* X) {
  std::vector<zc> b = from_colmajor(n, Bm);
  if (!host_lu_solve((int)n, from_colmajor(n, M), b)) return -1;
  to_colmajor(n, b, X);
  return 0;
}
}
