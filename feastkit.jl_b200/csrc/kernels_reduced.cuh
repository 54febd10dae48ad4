// On-device reduced eigensolver for the r x r (r <= 128) Rayleigh-Ritz pencil  Sq v = lambda Aq v.
// Replaces eigen(Hermitian(Sq), Hermitian(Aq)) -> LAPACK zhegv (dense/feast_dense.jl:272,
// sparse/feast_sparse.jl:412, banded/feast_banded.jl:750).
//
// One CTA: (1) Cholesky Aq = L L^H, (2) C = L^-1 Sq L^-H, (3) cyclic parallel-order two-sided Jacobi on
// the Hermitian C with accumulated rotations, (4) X = L^-H V, (5) ascending sort.  Output vectors
// satisfy X^H Aq X = I like zhegv's.  Matrices live in global memory (L1/L2 resident, <= 256 KB each).
#pragma once
#include "cxmath.cuh"
#include "kernels_block.cuh"

namespace feastcuda {

template <typename R>
struct ReducedEigArgs {
  int r;
  int generalized;       // 0: Aq = I
  cx<R>* S;              // in: Sq (row-major r x r, Hermitian); overwritten (becomes C)
  cx<R>* Bq;             // in: Aq (row-major); overwritten by L (lower)
  cx<R>* V;              // work r x r
  cx<R>* W;              // work r x r
  R* lam;                // out: ascending eigenvalues
  cx<R>* Xout;           // out: eigenvectors, row-major r x r, column k <-> lam[k]
  int* status;           // out: 0 ok, 1 Aq not positive definite, 2 Jacobi did not converge
  int* sweeps;           // out
};

template <typename R>
__global__ void __launch_bounds__(512) k_reduced_eig(ReducedEigArgs<R> a) {
  const int r = a.r, tid = threadIdx.x, NT = blockDim.x;
  cx<R>* C = a.S;
  cx<R>* L = a.Bq;
  cx<R>* V = a.V;
  cx<R>* W = a.W;
  __shared__ R s_c[FC_MAXCOLS / 2 + 1], s_s[FC_MAXCOLS / 2 + 1];
  __shared__ cx<R> s_ph[FC_MAXCOLS / 2 + 1];
  __shared__ int s_p[FC_MAXCOLS / 2 + 1], s_q[FC_MAXCOLS / 2 + 1];
  __shared__ R s_red[512];
  __shared__ int s_flag;
  __shared__ int s_perm[FC_MAXCOLS];
  __shared__ R s_d[FC_MAXCOLS];
  if (tid == 0) { s_flag = 0; *a.status = 0; *a.sweeps = 0; }
  __syncthreads();

  if (a.generalized) {
    // (1) left-looking Cholesky, thread i owns row i of L
    for (int j = 0; j < r; ++j) {
      cx<R> s = czero<R>();
      if (tid >= j && tid < r) {
        s = L[tid * r + j];
        for (int k = 0; k < j; ++k) s = s - L[tid * r + k] * conj(L[j * r + k]);
        if (tid == j) { s_d[j] = s.x; if (!(s.x > R(0))) s_flag = 1; }
      }
      __syncthreads();
      if (s_flag) break;
      if (tid >= j && tid < r) {
        const R d = sqrt(s_d[j]);
        L[tid * r + j] = (tid == j) ? mk<R>(d, R(0)) : mk<R>(s.x / d, s.y / d);
      }
      __syncthreads();
    }
    if (s_flag) { if (tid == 0) *a.status = 1; return; }
    // (2a) W = L^-1 S : thread c owns column c
    if (tid < r) {
      const int c = tid;
      for (int i = 0; i < r; ++i) {
        cx<R> s = C[i * r + c];
        for (int k = 0; k < i; ++k) s = s - L[i * r + k] * W[k * r + c];
        const R d = L[i * r + i].x;
        W[i * r + c] = mk<R>(s.x / d, s.y / d);
      }
    }
    __syncthreads();
    // (2b) Y = L^-1 W^H (stored in V), then C = Y^H
    if (tid < r) {
      const int c = tid;
      for (int i = 0; i < r; ++i) {
        cx<R> s = conj(W[c * r + i]);
        for (int k = 0; k < i; ++k) s = s - L[i * r + k] * V[k * r + c];
        const R d = L[i * r + i].x;
        V[i * r + c] = mk<R>(s.x / d, s.y / d);
      }
    }
    __syncthreads();
    for (int e = tid; e < r * r; e += NT) { const int i = e / r, j = e % r; C[e] = conj(V[j * r + i]); }
    __syncthreads();
  }
  // Hermitian part of C (also removes round-off asymmetry), V = I
  for (int e = tid; e < r * r; e += NT) {
    const int i = e / r, j = e % r;
    if (i < j) {
      const cx<R> u = C[i * r + j], l = C[j * r + i];
      const cx<R> h = mk<R>(R(0.5) * (u.x + l.x), R(0.5) * (u.y - l.y));
      C[i * r + j] = h;
      C[j * r + i] = conj(h);
    } else if (i == j) C[e].y = R(0);
    W[e] = czero<R>();
  }
  __syncthreads();
  for (int e = tid; e < r * r; e += NT) V[e] = ((e / r) == (e % r)) ? mk<R>(R(1), R(0)) : czero<R>();
  __syncthreads();

  // (3) Jacobi sweeps, round-robin ordering
  const int rp = r + (r & 1);
  const int np = rp / 2;
  int sweep = 0;
  const int max_sweeps = 60;
  R prev_off = R(0);
  for (; sweep < max_sweeps && r > 1; ++sweep) {
    R off = R(0), tot = R(0);
    for (int e = tid; e < r * r; e += NT) {
      const R v = abs2(C[e]);
      tot += v;
      if ((e / r) != (e % r)) off += v;
    }
    s_red[tid] = off;
    __syncthreads();
    for (int w = NT / 2; w > 0; w >>= 1) { if (tid < w) s_red[tid] += s_red[tid + w]; __syncthreads(); }
    const R off_all = s_red[0];
    __syncthreads();
    s_red[tid] = tot;
    __syncthreads();
    for (int w = NT / 2; w > 0; w >>= 1) { if (tid < w) s_red[tid] += s_red[tid + w]; __syncthreads(); }
    const R tot_all = s_red[0];
    __syncthreads();
    const R epsm = (sizeof(R) == 8) ? R(2.220446049250313e-16) : R(1.1920929e-7);
    if (off_all <= R(r) * epsm * epsm * tot_all) break;
    // at the round-off floor the off-diagonal mass stops shrinking: accept
    if (sweep > 3 && off_all >= R(0.25) * prev_off && off_all <= R(1e4) * R(r) * R(r) * epsm * epsm * tot_all) break;
    prev_off = off_all;
    for (int step = 0; step < rp - 1; ++step) {
      if (tid < np) {
        int p, q;
        if (tid == 0) { p = step % (rp - 1); q = rp - 1; }
        else { p = (step + tid) % (rp - 1); q = (step - tid + (rp - 1)) % (rp - 1); }
        if (p > q) { const int t = p; p = q; q = t; }
        R c = R(1), s = R(0);
        cx<R> ph = mk<R>(R(1), R(0));
        if (q < r) {
          const cx<R> g = C[p * r + q];
          const R ag = sqrt(abs2(g));
          if (ag > R(0) && abs2(g) > tiny_of<R>()) {
            const R ap = C[p * r + p].x, aq = C[q * r + q].x;
            ph = mk<R>(g.x / ag, g.y / ag);
            const R th = (aq - ap) / (R(2) * ag);
            const R t = ((th >= R(0)) ? R(1) : R(-1)) / (fabs(th) + sqrt(th * th + R(1)));
            c = R(1) / sqrt(t * t + R(1));
            s = t * c;
          }
        } else { q = -1; }
        s_p[tid] = p; s_q[tid] = q; s_c[tid] = c; s_s[tid] = s; s_ph[tid] = ph;
      }
      __syncthreads();
      // rows: C <- J^H C
      for (int e = tid; e < np * r; e += NT) {
        const int i = e / r, j = e % r;
        const int p = s_p[i], q = s_q[i];
        if (q >= 0 && s_s[i] != R(0)) {
          const R c = s_c[i], s = s_s[i];
          const cx<R> ph = s_ph[i];
          const cx<R> cp = C[p * r + j], cq = ph * C[q * r + j];
          C[p * r + j] = c * cp - s * cq;
          C[q * r + j] = s * cp + c * cq;
        }
      }
      __syncthreads();
      // columns: C <- C J, V <- V J
      for (int e = tid; e < np * r; e += NT) {
        const int i = e % np, j = e / np;
        const int p = s_p[i], q = s_q[i];
        if (q >= 0 && s_s[i] != R(0)) {
          const R c = s_c[i], s = s_s[i];
          const cx<R> phc = conj(s_ph[i]);
          cx<R> cp = C[j * r + p], cq = phc * C[j * r + q];
          C[j * r + p] = c * cp - s * cq;
          C[j * r + q] = s * cp + c * cq;
          cp = V[j * r + p]; cq = phc * V[j * r + q];
          V[j * r + p] = c * cp - s * cq;
          V[j * r + q] = s * cp + c * cq;
        }
      }
      __syncthreads();
      if (tid < np && s_q[tid] >= 0 && s_s[tid] != R(0)) {
        const int p = s_p[tid], q = s_q[tid];
        C[p * r + q] = czero<R>();
        C[q * r + p] = czero<R>();
        C[p * r + p].y = R(0);
        C[q * r + q].y = R(0);
      }
      __syncthreads();
    }
  }
  if (tid == 0) { *a.sweeps = sweep; if (sweep >= max_sweeps) *a.status = 2; }

  // (4) X = L^-H V (generalized), result in W; standard: W = V
  if (a.generalized) {
    if (tid < r) {
      const int c = tid;
      for (int i = r - 1; i >= 0; --i) {
        cx<R> s = V[i * r + c];
        for (int k = i + 1; k < r; ++k) s = s - conj(L[k * r + i]) * W[k * r + c];
        const R d = L[i * r + i].x;
        W[i * r + c] = mk<R>(s.x / d, s.y / d);
      }
    }
  } else {
    for (int e = tid; e < r * r; e += NT) W[e] = V[e];
  }
  __syncthreads();
  // (5) ascending sort (stable insertion on an index array)
  if (tid == 0) {
    for (int i = 0; i < r; ++i) { s_perm[i] = i; s_d[i] = C[i * r + i].x; }
    for (int i = 1; i < r; ++i) {
      const int pi = s_perm[i];
      const R v = s_d[pi];
      int k = i;
      while (k > 0 && s_d[s_perm[k - 1]] > v) { s_perm[k] = s_perm[k - 1]; --k; }
      s_perm[k] = pi;
    }
  }
  __syncthreads();
  for (int e = tid; e < r * r; e += NT) { const int i = e / r, k = e % r; a.Xout[e] = W[i * r + s_perm[k]]; }
  if (tid < r) a.lam[tid] = s_d[s_perm[tid]];
}

}  // namespace feastcuda
