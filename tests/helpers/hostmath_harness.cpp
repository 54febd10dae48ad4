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
// X = M^-1 Bm (n x n right-hand sides), 0 on success
int hm_lu_solve(long n, const double* M, const double* Bm, double* X) {
  std::vector<zc> b = from_colmajor(n, Bm);
  if (!host_lu_solve((int)n, from_colmajor(n, M), b)) return -1;
  to_colmajor(n, b, X);
  return 0;
}
}
