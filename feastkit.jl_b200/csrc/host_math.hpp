// Tiny host-side pieces of the path: feastdefault!/contours (core/feast_parameters.jl:41-386,
// core/feast_tools.jl:212-371) and the m x m algebra that steers the device orthonormalisation
// (the rank decision of _feast_qr_compress!, core/feast_aux.jl:101-131).
#pragma once
#include "zolotarev_tables.hpp"
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <limits>
#include <vector>

namespace feastcuda {

typedef std::complex<double> zc;
constexpr int64_t FEAST_UNINIT = -111;

// returns 0 on success, 1 where the reference throws ArgumentError
inline int host_feastdefault(int64_t* fpm) {
  auto g = [&](int k) -> int64_t& { return fpm[k - 1]; };
  const int64_t U = FEAST_UNINIT;
  int dig[7] = {0, 0, 0, 0, 0, 0, 0};
  if (g(30) != U && g(30) > 0) {
    int64_t rem = g(30);
    for (int i = 1; i <= 6; ++i) { dig[7 - i] = (int)(rem % 10); rem /= 10; }
  }
  if (g(1) == U) g(1) = 0; else if (g(1) > 1) return 1;
  if (g(14) == U) g(14) = 0; else if (g(14) < 0 || g(14) > 2) return 1;
  if (g(16) == U) {
    g(16) = 0;
    if (dig[3] == 2) g(16) = 1;
    if (dig[4] == 3) g(16) = 1;
    if (dig[4] == 1 && dig[2] == 4) g(16) = 1;
  } else if (g(16) < 0 || g(16) > 2) return 1;
  if (g(16) == 2 && (dig[4] == 3 || (dig[4] == 1 && dig[2] == 4))) return 1;
  if (g(2) == U || g(2) <= 0) {
    g(2) = 8;
    if (dig[3] == 2) g(2) = 4;
    if (g(14) == 2) g(2) = 3;
  } else if ((g(16) == 0 || g(16) == 2) && g(2) > 20) {
    const int64_t v = g(2);
    if (!(v == 24 || v == 32 || v == 40 || v == 48 || v == 56)) return 1;
  }
  if (g(3) == U) g(3) = 12; else if (g(3) < 0 || g(3) > 16) return 1;
  if (g(4) == U || g(4) <= 0) { g(4) = 20; if (dig[3] == 2) g(4) = 50; }
  if (g(5) == U) g(5) = 0; else if (g(5) != 0 && g(5) != 1) return 1;
  if (g(6) == U) g(6) = 1; else if (g(6) != 0 && g(6) != 1) return 1;
  if (g(7) == U) g(7) = 5; else if (g(7) < 0 || g(7) > 7) return 1;
  if (g(8) == U || g(8) <= 0) {
    g(8) = 16;
    if (dig[3] == 2) g(8) = 8;
    if (g(14) == 2) g(8) = 6;
  } else if (g(8) < 2) return 1;
  else if (g(16) == 0 && g(8) > 40) {
    const int64_t v = g(8);
    if (!(v == 48 || v == 64 || v == 80 || v == 96 || v == 112)) return 1;
  }
  if (g(9) == U) g(9) = 0;
  if (g(10) == U) { g(10) = 1; if (dig[5] == 1) g(10) = 0; } else if (g(10) != 0 && g(10) != 1) return 1;
  if (g(11) == U) g(11) = 0;
  if (g(12) == U) g(12) = 0;
  if (g(13) == U) g(13) = 0; else if (g(13) < 0 || g(13) > 3) return 1;
  if (g(15) == U) { g(15) = 0; if (dig[4] == 1) g(15) = 2; } else if (g(15) < 0 || g(15) > 2) return 1;
  if (g(14) == 2) g(15) = 1;
  if (g(17) == U) g(17) = 0;
  if (g(18) == U) {
    g(18) = 100;
    if (dig[3] == 1 && dig[6] <= 5) {
      if (dig[4] == 2) g(18) = 30;
      if (dig[4] == 1 && dig[2] != 3 && dig[2] != 4) g(18) = 30;
    }
  } else if (g(18) < 0) return 1;
  if (g(19) == U) g(19) = 0; else if (g(19) < -180 || g(19) > 180) return 1;
  for (int k = 20; k <= 28; ++k) if (g(k) == U) g(k) = 0;
  if (g(29) == U) g(29) = 0;
  if (g(31) == U) g(31) = 40;
  if (g(32) == U) g(32) = 10;
  for (int k = 33; k <= 35; ++k) if (g(k) == U) g(k) = 0;
  if (g(36) == U) g(36) = 1;
  if (g(37) == U) g(37) = 0;
  if (g(38) == U) g(38) = 1;
  if (g(39) == U) g(39) = 0;
  if (g(40) == U) g(40) = 0;
  if (g(41) == U) g(41) = 1;
  if (g(42) == U) g(42) = 1;
  if (g(43) == U) g(43) = 0;
  if (g(44) == U) g(44) = 0;
  if (g(45) == U) g(45) = 1;
  if (g(46) == U) g(46) = 40;
  for (int k = 47; k <= 49; ++k) if (g(k) == U) g(k) = 0;
  for (int k = 50; k <= 58; ++k) if (g(k) == U) g(k) = 0;
  for (int k = 59; k <= 64; ++k) if (g(k) == U) g(k) = 0;
  return 0;
}

inline double host_feast_tolerance(const int64_t* fpm) {
  const int64_t e = fpm[2];
  if (e < 0 || e > 16) return 1e-12;
  return std::pow(10.0, -(double)e);
}

// Gauss-Legendre nodes (ascending) and weights on [-1,1] by Newton iteration on P_n
inline void gauss_legendre(int n, std::vector<double>& x, std::vector<double>& w) {
  x.assign(n, 0.0);
  w.assign(n, 0.0);
  const double pi = 3.14159265358979323846;
  for (int i = 0; i < (n + 1) / 2; ++i) {
    double t = std::cos(pi * (i + 0.75) / (n + 0.5));
    double pp = 1.0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 0; j < n; ++j) {
        const double p3 = p2;
        p2 = p1;
        p1 = ((2.0 * j + 1.0) * t * p2 - j * p3) / (j + 1.0);
      }
      pp = n * (t * p1 - p2) / (t * t - 1.0);
      const double dt = p1 / pp;
      t -= dt;
      if (std::fabs(dt) < 1e-16) break;
    }
    // final evaluation of the derivative at the converged node
    double p1 = 1.0, p2 = 0.0;
    for (int j = 0; j < n; ++j) {
      const double p3 = p2;
      p2 = p1;
      p1 = ((2.0 * j + 1.0) * t * p2 - j * p3) / (j + 1.0);
    }
    pp = n * (t * p1 - p2) / (t * t - 1.0);
    x[i] = -t;
    x[n - 1 - i] = t;
    w[i] = 2.0 / ((1.0 - t * t) * pp * pp);
    w[n - 1 - i] = w[i];
  }
  if (n % 2 == 1) x[n / 2] = 0.0;
}

// feast_contour, core/feast_tools.jl:212-284 (Gauss fpm[16]=0, trapezoid fpm[16]=1, Zolotarev fpm[16]=2)
inline int host_contour(double Emin, double Emax, int64_t* fpm, zc* Z, zc* W) {
  if (fpm[1] == FEAST_UNINIT || fpm[1] <= 0)
    if (host_feastdefault(fpm)) return 1;
  const int ne = (int)fpm[1];
  const int64_t t16 = fpm[15];
  const double aspect = fpm[17] * 0.01;
  const double pi = 3.14159265358979323846;
  const double r = (Emax - Emin) / 2.0, Emid = Emin + r;
  const double ba = -pi / 2, ab = pi / 2;
  std::vector<double> xg, wg;
  if (t16 == 0) gauss_legendre(ne, xg, wg);
  const int zoff = (t16 == 2) ? zolo_offset(ne) : -1;
  for (int e = 0; e < ne; ++e) {
    if (t16 == 2) {
      // zolotarev_point (core/feast_tools.jl:182-210): table value, or the reference's fallback for an ne without a table
      zc xe, we;
      if (zoff >= 0) { const ZoloNode& zn = ZOLO_NODES[zoff + e]; xe = zc(zn.xr, zn.xi); we = zc(zn.wr, zn.wi); }
      else { const double th = pi * (2.0 * (e + 1) - 1.0) / (2.0 * ne); xe = zc(std::cos(th), std::sin(th)); we = zc(0.0, pi / ne); }
      Z[e] = xe * r + Emid;       // core/feast_tools.jl:263-266 (the ellipse ratio fpm[18] does not enter)
      W[e] = we * r;
    } else if (t16 == 0) {
      const double th = ba * xg[e] + ab;
      Z[e] = zc(Emid + r * std::cos(th), r * aspect * std::sin(th));
      const zc jac(r * aspect * std::cos(th), r * std::sin(th));
      W[e] = 0.25 * wg[e] * jac;
    } else {
      const double th = pi - (pi / ne) / 2 - (pi / ne) * e;
      Z[e] = zc(Emid + r * std::cos(th), r * aspect * std::sin(th));
      const zc jac(r * aspect * std::cos(th), r * std::sin(th));
      W[e] = (1.0 / (2.0 * ne)) * jac;
    }
  }
  return 0;
}

// feast_gcontour, core/feast_tools.jl:286-371
inline int host_gcontour(zc Emid, double r, int64_t* fpm, zc* Z, zc* W) {
  if (fpm[7] == FEAST_UNINIT || fpm[7] <= 0)
    if (host_feastdefault(fpm)) return 1;
  const int ne = (int)fpm[7];
  const int64_t t16 = fpm[15];
  const double aspect = fpm[17] * 0.01;
  const double pi = 3.14159265358979323846;
  const double rot = (fpm[18] / 180.0) * pi;
  const zc nr = r * zc(std::cos(rot), std::sin(rot));
  const zc I(0.0, 1.0);
  const double ba = -pi / 2, ab = pi / 2;
  if (t16 == 0) {
    const int nu = ne / 2, nl = ne - nu;
    std::vector<double> xu, wu, xl, wl;
    if (nu > 0) gauss_legendre(nu, xu, wu);
    gauss_legendre(nl, xl, wl);
    for (int e = 0; e < nu; ++e) {
      const double th = ba * xu[e] + ab;
      Z[e] = Emid + nr * std::cos(th) + nr * I * aspect * std::sin(th);
      W[e] = 0.25 * wu[e] * (nr * I * std::sin(th) + nr * aspect * std::cos(th));
    }
    for (int e = nu; e < ne; ++e) {
      const int i = e - nu;
      const double th = -ba * xl[i] - ab;
      Z[e] = Emid + nr * std::cos(th) + nr * I * aspect * std::sin(th);
      W[e] = 0.25 * wl[i] * (nr * I * std::sin(th) + nr * aspect * std::cos(th));
    }
  } else {
    for (int e = 0; e < ne; ++e) {
      const double th = pi - (2 * pi / ne) / 2 - (2 * pi / ne) * e;
      Z[e] = Emid + nr * std::cos(th) + nr * I * aspect * std::sin(th);
      W[e] = (1.0 / ne) * (nr * I * std::sin(th) + nr * aspect * std::cos(th));
    }
  }
  return 0;
}

// feast_inside_gcontour, core/feast_tools.jl:623-650
inline bool host_inside_gcontour(zc lam, zc Emid, double r, const int64_t* fpm) {
  zc w = lam - Emid;
  double aspect = 1.0, rot = 0.0;
  const double pi = 3.14159265358979323846;
  if (fpm[17] > 0) aspect = fpm[17] * 0.01;
  if (fpm[18] != 0) rot = (fpm[18] / 180.0) * pi;
  if (rot != 0.0) w *= zc(std::cos(-rot), std::sin(-rot));
  const double x = w.real() / r, y = w.imag() / (r * aspect);
  return x * x + y * y <= 1.0;
}

// node block of a rank, parallel/feast_mpi.jl:36-43 (0-based start)
inline void host_node_partition(int64_t ne, int nranks, int rank, int64_t* start, int64_t* count) {
  const int64_t base = ne / nranks, rem = ne % nranks;
  *start = rank * base + std::min<int64_t>(rank, rem);
  *count = base + (rank < rem ? 1 : 0);
}

// ---- one pass of the Gram-driven rank-revealing orthonormalisation ---------------------------------
// Input: G = Z^H Z (cur x cur, row-major, Hermitian) of the current block whose first `done` columns
// are (nearly) orthonormal.  Runs a Cholesky with the first `done` pivots forced in order and the rest
// chosen by largest remaining squared norm (= pivoted QR's |R_kk|^2, core/feast_aux.jl:113-124).
// A pivot is accepted while sqrt(d) > thr_abs (the reference's rank threshold) and d >= safe*d_init
// (enough digits left in the Schur complement; otherwise the column is re-examined next pass after an
// explicit projection).  Output T (cur x ncols_out, row-major): Znext = Z*T = [Q_accepted | W_rest].
struct OrthoPass {
  int accepted = 0;   // orthonormal columns at the front of Znext
  int kept = 0;       // residual columns carried to the next pass
  int dropped = 0;    // columns below the rank threshold
  std::vector<zc> T;  // cur x (accepted + kept)
};

inline OrthoPass ortho_pass(const std::vector<zc>& G, int cur, int done, double thr_abs, double safe) {
  OrthoPass out;
  std::vector<zc> S(G);
  std::vector<double> dinit(cur);
  for (int j = 0; j < cur; ++j) dinit[j] = S[(size_t)j * cur + j].real();
  std::vector<int> order;
  std::vector<char> used(cur, 0);
  std::vector<zc> Lf((size_t)cur * cur, zc(0));  // Lf[row j][step t]: Cholesky factor entries for every index
  bool accepting = true;
  for (int t = 0; t < cur && accepting; ++t) {
    int piv = -1;
    if (t < done) piv = t;
    else {
      double best = -1.0;
      for (int j = 0; j < cur; ++j)
        if (!used[j] && S[(size_t)j * cur + j].real() > best) { best = S[(size_t)j * cur + j].real(); piv = j; }
      if (piv < 0) break;
      const double d = best;
      if (!(d > 0.0) || std::sqrt(d) <= thr_abs) {  // all remaining columns are below the rank threshold
        for (int j = 0; j < cur; ++j) if (!used[j]) { used[j] = 2; out.dropped++; }
        break;
      }
      if (d < safe * dinit[piv]) { accepting = false; break; }
    }
    const double d = S[(size_t)piv * cur + piv].real();
    if (!(d > 0.0)) { accepting = false; break; }
    const double sd = std::sqrt(d);
    used[piv] = 1;
    order.push_back(piv);
    const int tt = (int)order.size() - 1;
    for (int j = 0; j < cur; ++j) {
      if (used[j] == 1 && j != piv) continue;
      Lf[(size_t)j * cur + tt] = (j == piv) ? zc(sd, 0.0) : S[(size_t)j * cur + piv] / sd;
    }
    for (int i = 0; i < cur; ++i) {
      if (used[i]) continue;
      const zc li = Lf[(size_t)i * cur + tt];
      for (int j = 0; j < cur; ++j) {
        if (used[j]) continue;
        S[(size_t)i * cur + j] -= li * std::conj(Lf[(size_t)j * cur + tt]);
      }
    }
  }
  const int k = (int)order.size();
  std::vector<int> rest;
  for (int j = 0; j < cur; ++j) if (used[j] == 0) rest.push_back(j);
  out.accepted = k;
  out.kept = (int)rest.size();
  const int nout = k + out.kept;
  out.T.assign((size_t)cur * std::max(nout, 1), zc(0));
  if (nout == 0) return out;
  // Lc (k x k lower): Lc[i][t] = Lf[order[i]][t];  Minv = Lc^-1 by forward substitution
  std::vector<zc> Minv((size_t)k * k, zc(0));
  for (int c = 0; c < k; ++c) {
    for (int i = c; i < k; ++i) {
      zc s = (i == c) ? zc(1.0) : zc(0.0);
      for (int t = c; t < i; ++t) s -= Lf[(size_t)order[i] * cur + t] * Minv[(size_t)t * k + c];
      Minv[(size_t)i * k + c] = s / Lf[(size_t)order[i] * cur + i];
    }
  }
  // accepted output column a: Z_P * (Lc^-H)[:, a]  ->  T[order[i]][a] = conj(Minv[a][i])
  for (int a = 0; a < k; ++a)
    for (int i = 0; i <= a; ++i) out.T[(size_t)order[i] * nout + a] = std::conj(Minv[(size_t)a * k + i]);
  // rest column b (index j): Z_j - Z_P * Lc^-H * conj(Lf[j][0:k])
  for (int b = 0; b < out.kept; ++b) {
    const int j = rest[b];
    out.T[(size_t)j * nout + k + b] = zc(1.0);
    for (int i = 0; i < k; ++i) {
      zc s(0.0);
      for (int a = i; a < k; ++a) s += std::conj(Minv[(size_t)a * k + i]) * std::conj(Lf[(size_t)j * cur + a]);
      out.T[(size_t)order[i] * nout + k + b] = -s;
    }
  }
  return out;
}

// ---- small dense complex algebra for the general (non-Hermitian) reduced problem ---------------------------------
// eigen(Aq, Sq) of kernel/feast_kernel.jl:812 (LAPACK zggev in the reference) for r <= 128: C = Sq^-1 Aq by LU with
// partial pivoting, then Hessenberg + shifted QR (complex Schur form) with back-substituted eigenvectors.
// All matrices row-major r x r.  Returns false if Sq is singular or the QR iteration does not converge.
inline bool host_lu_solve(int n, std::vector<zc> M, std::vector<zc>& Bm /* n x n, overwritten by M^-1 Bm */) {
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = std::abs(M[(size_t)k * n + k]);
    for (int i = k + 1; i < n; ++i)
      if (std::abs(M[(size_t)i * n + k]) > best) { best = std::abs(M[(size_t)i * n + k]); p = i; }
    if (!(best > 0.0)) return false;
    if (p != k)
      for (int j = 0; j < n; ++j) { std::swap(M[(size_t)k * n + j], M[(size_t)p * n + j]); std::swap(Bm[(size_t)k * n + j], Bm[(size_t)p * n + j]); }
    const zc inv = zc(1.0) / M[(size_t)k * n + k];
    for (int i = k + 1; i < n; ++i) {
      const zc l = M[(size_t)i * n + k] * inv;
      if (l == zc(0.0)) continue;
      for (int j = k + 1; j < n; ++j) M[(size_t)i * n + j] -= l * M[(size_t)k * n + j];
      for (int j = 0; j < n; ++j) Bm[(size_t)i * n + j] -= l * Bm[(size_t)k * n + j];
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    const zc inv = zc(1.0) / M[(size_t)i * n + i];
    for (int j = 0; j < n; ++j) {
      zc s = Bm[(size_t)i * n + j];
      for (int k = i + 1; k < n; ++k) s -= M[(size_t)i * n + k] * Bm[(size_t)k * n + j];
      Bm[(size_t)i * n + j] = s * inv;
    }
  }
  return true;
}

inline bool host_complex_eig(int n, std::vector<zc> A, std::vector<zc>& lam, std::vector<zc>& V) {
  auto a = [&](int i, int j) -> zc& { return A[(size_t)i * n + j]; };
  std::vector<zc> Z((size_t)n * n, zc(0.0));
  for (int i = 0; i < n; ++i) Z[(size_t)i * n + i] = 1.0;
  double anorm = 0.0;
  for (auto& v : A) anorm = std::max(anorm, std::abs(v));
  if (!(anorm > 0.0)) anorm = 1.0;
  const double eps = 2.220446049250313e-16;
  // Householder reduction to upper Hessenberg form, H = Q^H A Q, Z = Q
  std::vector<zc> v(n);
  for (int k = 0; k + 2 < n; ++k) {
    double xn2 = 0.0;
    for (int i = k + 1; i < n; ++i) xn2 += std::norm(a(i, k));
    double tail = 0.0;
    for (int i = k + 2; i < n; ++i) tail += std::norm(a(i, k));
    if (!(tail > 0.0)) continue;
    const double xn = std::sqrt(xn2);
    const zc x0 = a(k + 1, k);
    const zc ph = std::abs(x0) > 0.0 ? x0 / std::abs(x0) : zc(1.0);
    const int len = n - k - 1;
    for (int i = 0; i < len; ++i) v[i] = a(k + 1 + i, k);
    v[0] += ph * xn;
    double vn2 = 0.0;
    for (int i = 0; i < len; ++i) vn2 += std::norm(v[i]);
    if (!(vn2 > 0.0)) continue;
    const double f = 2.0 / vn2;
    for (int j = 0; j < n; ++j) {          // A <- H A
      zc sdot(0.0);
      for (int i = 0; i < len; ++i) sdot += std::conj(v[i]) * a(k + 1 + i, j);
      sdot *= f;
      for (int i = 0; i < len; ++i) a(k + 1 + i, j) -= v[i] * sdot;
    }
    for (int i = 0; i < n; ++i) {          // A <- A H, Z <- Z H
      zc sdot(0.0), zdot(0.0);
      for (int j = 0; j < len; ++j) { sdot += a(i, k + 1 + j) * v[j]; zdot += Z[(size_t)i * n + k + 1 + j] * v[j]; }
      sdot *= f;
      zdot *= f;
      for (int j = 0; j < len; ++j) { a(i, k + 1 + j) -= sdot * std::conj(v[j]); Z[(size_t)i * n + k + 1 + j] -= zdot * std::conj(v[j]); }
    }
    for (int i = k + 2; i < n; ++i) a(i, k) = 0.0;
  }
  // single-shift QR iteration on the active window [l, hi] (Wilkinson shift, Givens rotations), full Schur form
  int hi = n - 1, iter = 0, total = 0;
  std::vector<zc> cs(n), sn(n);
  while (hi >= 0) {
    int l = hi;
    while (l > 0) {
      double sc = std::abs(a(l - 1, l - 1)) + std::abs(a(l, l));
      if (sc == 0.0) sc = anorm;
      if (std::abs(a(l, l - 1)) <= eps * sc) { a(l, l - 1) = 0.0; break; }
      --l;
    }
    if (l == hi) { --hi; iter = 0; continue; }
    if (++total > 60 * n + 200) return false;
    zc mu;
    if (iter == 10 || iter == 20) mu = a(hi, hi) + zc(std::abs(a(hi, hi - 1).real()) + std::abs(a(hi, hi - 1).imag()), 0.0);
    else {
      const zc p = a(hi - 1, hi - 1), q = a(hi - 1, hi), r = a(hi, hi - 1), t = a(hi, hi);
      const zc half = 0.5 * (p - t), disc = std::sqrt(half * half + q * r);
      const zc m1 = 0.5 * (p + t) + disc, m2 = 0.5 * (p + t) - disc;
      mu = (std::abs(m1 - t) < std::abs(m2 - t)) ? m1 : m2;
    }
    ++iter;
    for (int i = l; i <= hi; ++i) a(i, i) -= mu;
    for (int k = l; k < hi; ++k) {          // left rotations: zero a(k+1, k)
      const zc x = a(k, k), y = a(k + 1, k);
      const double nr = std::sqrt(std::norm(x) + std::norm(y));
      zc c(1.0), s2(0.0);
      if (nr > 0.0) { c = x / nr; s2 = y / nr; }
      cs[k] = c;
      sn[k] = s2;
      for (int j = k; j < n; ++j) {
        const zc u = a(k, j), w = a(k + 1, j);
        a(k, j) = std::conj(c) * u + std::conj(s2) * w;
        a(k + 1, j) = -s2 * u + c * w;
      }
    }
    for (int k = l; k < hi; ++k) {          // right rotations (G^H), also accumulated into Z
      const zc c = cs[k], s2 = sn[k];
      const int top = std::min(k + 1, hi);
      for (int i = 0; i <= top; ++i) {
        const zc u = a(i, k), w = a(i, k + 1);
        a(i, k) = u * c + w * s2;
        a(i, k + 1) = -u * std::conj(s2) + w * std::conj(c);
      }
      for (int i = 0; i < n; ++i) {
        const zc u = Z[(size_t)i * n + k], w = Z[(size_t)i * n + k + 1];
        Z[(size_t)i * n + k] = u * c + w * s2;
        Z[(size_t)i * n + k + 1] = -u * std::conj(s2) + w * std::conj(c);
      }
    }
    for (int i = l; i <= hi; ++i) a(i, i) += mu;
  }
  // eigenvectors of the triangular factor by back substitution, then x = Z y
  lam.resize(n);
  V.assign((size_t)n * n, zc(0.0));
  std::vector<zc> y(n);
  const double small = eps * anorm;
  for (int k = 0; k < n; ++k) {
    lam[k] = a(k, k);
    y[k] = 1.0;
    for (int i = k - 1; i >= 0; --i) {
      zc sdot(0.0);
      for (int j = i + 1; j <= k; ++j) sdot += a(i, j) * y[j];
      zc d = a(i, i) - lam[k];
      if (std::abs(d) < small) d = small;
      y[i] = -sdot / d;
    }
    double nrm = 0.0;
    for (int i = 0; i < n; ++i) {
      zc sdot(0.0);
      for (int j = 0; j <= k; ++j) sdot += Z[(size_t)i * n + j] * y[j];
      V[(size_t)i * n + k] = sdot;
      nrm += std::norm(sdot);
    }
    nrm = nrm > 0.0 ? std::sqrt(nrm) : 1.0;
    for (int i = 0; i < n; ++i) V[(size_t)i * n + k] /= nrm;
  }
  return true;
}

// Eigenpairs of the small pencil S v = lambda B v (row-major r x r) that tolerates a rank-deficient B, the way the
// reference's eigen(Sq, Aq) (LAPACK QZ, kernel/feast_kernel.jl:186, :532) does for FEAST moment matrices when the search
// space is larger than the number of eigenvalues inside: Householder QR with column pivoting B P = Q R reveals the
// numerical rank k; rows k.. of Q^H S P and of R carry only round-off when the null spaces of S and B coincide (the moment
// case: both are Q0^H f(A) Q0), so the k finite eigenpairs come from the leading k x k blocks and the remaining r-k
// directions are reported as lambda = +inf with the coordinate vector P e_j of a pivoted-out column (V stays regular).  A regular pencil with singular B
// (significant trailing rows of Q^H S P) eliminates them by a Schur complement instead.
inline bool host_pencil_eig(int n, const std::vector<zc>& S, const std::vector<zc>& B, std::vector<zc>& lam, std::vector<zc>& V,
                            int* rank_out = nullptr) {
  for (const zc& v : S) if (!std::isfinite(v.real()) || !std::isfinite(v.imag())) return false;
  for (const zc& v : B) if (!std::isfinite(v.real()) || !std::isfinite(v.imag())) return false;
  std::vector<zc> R = B, T = S;
  std::vector<int> perm(n);
  for (int j = 0; j < n; ++j) perm[j] = j;
  auto r_ = [&](int i, int j) -> zc& { return R[(size_t)i * n + j]; };
  auto t_ = [&](int i, int j) -> zc& { return T[(size_t)i * n + j]; };
  std::vector<double> cn(n);
  std::vector<zc> w(n);
  double r11 = 0.0;
  int k = 0;
  for (; k < n; ++k) {
    int p = k;
    double best = -1.0;
    for (int j = k; j < n; ++j) {
      double s = 0.0;
      for (int i = k; i < n; ++i) s += std::norm(r_(i, j));
      cn[j] = s;
      if (s > best) { best = s; p = j; }
    }
    best = std::sqrt(std::max(best, 0.0));
    if (k == 0) r11 = best;
    if (!(best > 1e-13 * r11) || !(best > 0.0)) break;
    if (p != k) {
      for (int i = 0; i < n; ++i) { std::swap(r_(i, k), r_(i, p)); std::swap(t_(i, k), t_(i, p)); }
      std::swap(perm[k], perm[p]);
    }
    // Householder H = I - 2 w w^H / (w^H w) zeroing R[k+1.., k]; applied from the left to R and to T (= Q^H S P)
    const zc x0 = r_(k, k);
    const zc phase = std::abs(x0) > 0.0 ? x0 / std::abs(x0) : zc(1.0);
    for (int i = k; i < n; ++i) w[i] = r_(i, k);
    w[k] += phase * best;
    double wn = 0.0;
    for (int i = k; i < n; ++i) wn += std::norm(w[i]);
    if (wn > 0.0) {
      for (int pass = 0; pass < 2; ++pass) {
        std::vector<zc>& M = pass == 0 ? R : T;
        for (int j = (pass == 0 ? k : 0); j < n; ++j) {
          zc d(0.0);
          for (int i = k; i < n; ++i) d += std::conj(w[i]) * M[(size_t)i * n + j];
          d *= 2.0 / wn;
          for (int i = k; i < n; ++i) M[(size_t)i * n + j] -= d * w[i];
        }
      }
    }
  }
  if (rank_out) *rank_out = k;
  lam.assign(n, zc(std::numeric_limits<double>::infinity(), 0.0));
  V.assign((size_t)n * n, zc(0.0));
  for (int j = k; j < n; ++j) V[(size_t)perm[j] * n + j] = 1.0;
  if (k == 0) return true;
  std::vector<zc> T11((size_t)k * k), R11((size_t)k * k);
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < k; ++j) { T11[(size_t)i * k + j] = t_(i, j); R11[(size_t)i * k + j] = r_(i, j); }
  const int m = n - k;
  std::vector<zc> Xel;  // m x k: w2 = -Xel w1 when the trailing rows of T are significant
  if (m > 0) {
    double tn = 0.0, bn = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) (i < k ? tn : bn) = std::max(i < k ? tn : bn, std::abs(t_(i, j)));
    if (bn > 1e-8 * std::max(tn, 1e-300)) {
      std::vector<zc> T22((size_t)m * m), T21((size_t)m * m, zc(0.0));
      // solve T22 X = T21 column block by block of width m (host_lu_solve works on square right-hand sides)
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) T22[(size_t)i * m + j] = t_(k + i, k + j);
      Xel.assign((size_t)m * k, zc(0.0));
      for (int c0 = 0; c0 < k; c0 += m) {
        std::fill(T21.begin(), T21.end(), zc(0.0));
        for (int i = 0; i < m; ++i)
          for (int j = 0; j < std::min(m, k - c0); ++j) T21[(size_t)i * m + j] = t_(k + i, c0 + j);
        if (!host_lu_solve(m, T22, T21)) return false;
        for (int i = 0; i < m; ++i)
          for (int j = 0; j < std::min(m, k - c0); ++j) Xel[(size_t)i * k + c0 + j] = T21[(size_t)i * m + j];
      }
      for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
          zc st(0.0), sr(0.0);
          for (int l = 0; l < m; ++l) { st += t_(i, k + l) * Xel[(size_t)l * k + j]; sr += r_(i, k + l) * Xel[(size_t)l * k + j]; }
          T11[(size_t)i * k + j] -= st;
          R11[(size_t)i * k + j] -= sr;
        }
    }
  }
  if (!host_lu_solve(k, R11, T11)) return false;
  std::vector<zc> lk, Vk;
  if (!host_complex_eig(k, T11, lk, Vk)) return false;
  for (int c = 0; c < k; ++c) {
    lam[c] = lk[c];
    std::vector<zc> full(n, zc(0.0));
    for (int i = 0; i < k; ++i) full[i] = Vk[(size_t)i * k + c];
    if (!Xel.empty())
      for (int l = 0; l < m; ++l) {
        zc s(0.0);
        for (int i = 0; i < k; ++i) s += Xel[(size_t)l * k + i] * Vk[(size_t)i * k + c];
        full[k + l] = -s;
      }
    double nrm = 0.0;
    for (int i = 0; i < n; ++i) nrm += std::norm(full[i]);
    nrm = nrm > 0.0 ? std::sqrt(nrm) : 1.0;
    for (int i = 0; i < n; ++i) V[(size_t)perm[i] * n + c] = full[i] / nrm;
  }
  return true;
}

}  // namespace feastcuda
