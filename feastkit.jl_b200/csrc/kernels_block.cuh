// Block-vector kernels on row-major n x ld complex blocks: the BiCGStab vector updates with fused
// reductions, the w_e-weighted accumulation (dense/feast_dense.jl:231, sparse/feast_sparse.jl:369,
// kernel/feast_kernel.jl:143,762-766), layout conversion at the ABI (Julia column-major <-> device
// row-major), tall-skinny Gram products Q^H Y (dense/feast_dense.jl:253,263; kernel/feast_kernel.jl:
// 790-807) and the back-projection X = Q*V (dense/feast_dense.jl:287-290, kernel/feast_kernel.jl:838-845).
#pragma once
#include "cxmath.cuh"

namespace feastcuda {

constexpr int FC_MAXCOLS = 128;  // columns handled per launch by every kernel

// Per-solve BiCGStab scalars, one entry per RHS column (device resident; no host sync per iteration).
template <typename R>
struct KrylovState {
  cx<R> rho[FC_MAXCOLS], alpha[FC_MAXCOLS], omega[FC_MAXCOLS], beta[FC_MAXCOLS];
  R target[FC_MAXCOLS], rnorm[FC_MAXCOLS], rn0[FC_MAXCOLS];
  int active[FC_MAXCOLS], iters[FC_MAXCOLS], flags[FC_MAXCOLS], diverged[FC_MAXCOLS];
  int n_active;
};

// ---- elementwise thread mapping: thread owns column c = tid % mp, rows strided -----------------
struct EwMap {
  int c, rsub, rpb;
  __device__ __forceinline__ EwMap(int mp) { c = threadIdx.x % mp; rsub = threadIdx.x / mp; rpb = blockDim.x / mp; }
};

template <typename R>
__device__ __forceinline__ void block_reduce_cols(R v, int mp, int m, R* out_partial_row /*[pstride]*/) {
  // sums the per-thread value over the threads that share a column; thread c<m writes the block result
  __shared__ R red[256];
  __syncthreads();
  red[threadIdx.x] = v;
  __syncthreads();
  if ((int)threadIdx.x < mp && (int)threadIdx.x < m) {
    R s = R(0);
    for (int q = threadIdx.x; q < (int)blockDim.x; q += mp) s += red[q];
    out_partial_row[threadIdx.x] = s;
  }
}

// s = r - alpha_c v
template <typename R>
__global__ void __launch_bounds__(256) k_upd_s(int64_t n, int m, int mp, int64_t ld, const KrylovState<R>* st,
                                               const cx<R>* __restrict__ r, const cx<R>* __restrict__ v,
                                               cx<R>* __restrict__ s) {
  EwMap e(mp);
  if (e.c >= m) return;
  const cx<R> al = st->alpha[e.c];
  for (int64_t row = (int64_t)blockIdx.x * e.rpb + e.rsub; row < n; row += (int64_t)gridDim.x * e.rpb) {
    const int64_t i = row * ld + e.c;
    s[i] = r[i] - al * v[i];
  }
}

// x += alpha p + omega s ; r = s - omega t ; p = r + beta (p - omega v) ; partial ||r||^2
template <typename R>
__global__ void __launch_bounds__(256) k_upd_xrp(int64_t n, int m, int mp, int64_t ld, const KrylovState<R>* st,
                                                 cx<R>* __restrict__ x, cx<R>* __restrict__ r, cx<R>* __restrict__ p,
                                                 const cx<R>* __restrict__ v, const cx<R>* __restrict__ s,
                                                 const cx<R>* __restrict__ t, R* __restrict__ partial, int pstride) {
  EwMap e(mp);
  R acc = R(0);
  if (e.c < m) {
    const cx<R> al = st->alpha[e.c], om = st->omega[e.c], be = st->beta[e.c];
    for (int64_t row = (int64_t)blockIdx.x * e.rpb + e.rsub; row < n; row += (int64_t)gridDim.x * e.rpb) {
      const int64_t i = row * ld + e.c;
      const cx<R> pi = p[i], si = s[i], ti = t[i], vi = v[i];
      x[i] = x[i] + al * pi + om * si;
      const cx<R> rn = si - om * ti;
      r[i] = rn;
      p[i] = rn + be * (pi - om * vi);
      acc += abs2(rn);
    }
  }
  block_reduce_cols<R>(acc, mp, m, partial + (int64_t)blockIdx.x * pstride);
}

// Y[:,c] = beta*Y[:,c] + w * X[:,c]   (weighted accumulation of the filtered subspace)
template <typename R>
__global__ void __launch_bounds__(256) k_axpby_cols(int64_t n, int m, int mp, int64_t ldx, int64_t ldy, cx<R> w, R beta,
                                                    const cx<R>* __restrict__ X, cx<R>* __restrict__ Y) {
  EwMap e(mp);
  if (e.c >= m) return;
  for (int64_t row = (int64_t)blockIdx.x * e.rpb + e.rsub; row < n; row += (int64_t)gridDim.x * e.rpb) {
    const cx<R> xv = X[row * ldx + e.c];
    cx<R>* y = Y + row * ldy + e.c;
    *y = (beta == R(0)) ? (w * xv) : (beta * (*y) + w * xv);
  }
}

// Y[:,c] = f_c * X[:,c]  (Ritz-pair initial guess X0 = Q diag(1/(z-theta)); column normalisation)
template <typename R>
__global__ void __launch_bounds__(256) k_scale_cols(int64_t n, int m, int mp, int64_t ldx, int64_t ldy,
                                                    const cx<R>* __restrict__ f, const cx<R>* __restrict__ X,
                                                    cx<R>* __restrict__ Y) {
  EwMap e(mp);
  if (e.c >= m) return;
  const cx<R> fc = f[e.c];
  for (int64_t row = (int64_t)blockIdx.x * e.rpb + e.rsub; row < n; row += (int64_t)gridDim.x * e.rpb)
    Y[row * ldy + e.c] = fc * X[row * ldx + e.c];
}

// X[:,c] = Bk[:,c] for the columns with mask[c] != 0 (restore parked BiCGStab columns)
template <typename R>
__global__ void __launch_bounds__(256) k_restore_cols(int64_t n, int m, int mp, int64_t ld, const int* __restrict__ mask,
                                                      const cx<R>* __restrict__ Bk, cx<R>* __restrict__ X) {
  EwMap e(mp);
  if (e.c >= m || !mask[e.c]) return;
  for (int64_t row = (int64_t)blockIdx.x * e.rpb + e.rsub; row < n; row += (int64_t)gridDim.x * e.rpb) X[row * ld + e.c] = Bk[row * ld + e.c];
}

template <typename R>
__global__ void __launch_bounds__(256) k_zero_imag(int64_t n, int m, int mp, int64_t ld, cx<R>* __restrict__ X) {
  EwMap e(mp);
  if (e.c >= m) return;
  for (int64_t row = (int64_t)blockIdx.x * e.rpb + e.rsub; row < n; row += (int64_t)gridDim.x * e.rpb) X[row * ld + e.c].y = R(0);
}

template <typename R>
__global__ void __launch_bounds__(256) k_colnorm2(int64_t n, int m, int mp, int64_t ld, const cx<R>* __restrict__ X,
                                                  R* __restrict__ partial, int pstride) {
  EwMap e(mp);
  R acc = R(0);
  if (e.c < m)
    for (int64_t row = (int64_t)blockIdx.x * e.rpb + e.rsub; row < n; row += (int64_t)gridDim.x * e.rpb) acc += abs2(X[row * ld + e.c]);
  block_reduce_cols<R>(acc, mp, m, partial + (int64_t)blockIdx.x * pstride);
}

// ---- reductions of per-block partials and the BiCGStab scalar recurrences ----------------------
// out[s*m + c] = sum_b partial[(s*nblocks + b)*pstride + c]; one block of 1024 threads.
template <typename T>
__device__ __forceinline__ void reduce_partials(const T* __restrict__ partial, int nslots, int nblocks, int pstride, int m,
                                                T* sm_out /*[nslots*FC_MAXCOLS]*/, T* sm_tmp /*[1024]*/) {
  // the block's threads are split into `nparts` groups of `mp` (smallest power of two >= m, at least 32) columns: fewer
  // columns -> more groups -> shorter dependent chains; the summation order is fixed for a given (m, nblocks)
  int mp = 32;
  while (mp < m) mp <<= 1;
  const int c = threadIdx.x % mp, part = threadIdx.x / mp, nparts = blockDim.x / mp;
  for (int s = 0; s < nslots; ++s) {
    T acc0 = zero_of<T>::v(), acc1 = zero_of<T>::v();
    if (c < m) {
      const T* base = partial + (int64_t)s * nblocks * pstride + c;
      int b = part;
      for (; b + nparts < nblocks; b += 2 * nparts) {
        acc0 = acc0 + base[(int64_t)b * pstride];
        acc1 = acc1 + base[(int64_t)(b + nparts) * pstride];
      }
      if (b < nblocks) acc0 = acc0 + base[(int64_t)b * pstride];
    }
    __syncthreads();
    sm_tmp[threadIdx.x] = acc0 + acc1;
    __syncthreads();
    if (part == 0 && c < m) {
      T t = sm_tmp[c];
      for (int q = 1; q < nparts; ++q) t = t + sm_tmp[q * mp + c];
      sm_out[s * FC_MAXCOLS + c] = t;
    }
  }
  __syncthreads();
}

template <typename R> __device__ __forceinline__ R tiny_of();
template <> __device__ __forceinline__ double tiny_of<double>() { return 2.2250738585072014e-308; }
template <> __device__ __forceinline__ float tiny_of<float>() { return 1.17549435e-38f; }

// after v = S p : alpha = rho / (rhat^H v)
template <typename R>
__global__ void __launch_bounds__(1024) k_bicg_scal1(KrylovState<R>* st, const cx<R>* partial, int nblocks, int pstride, int m) {
  __shared__ cx<R> so[FC_MAXCOLS];
  __shared__ cx<R> tmp[1024];
  reduce_partials<cx<R>>(partial, 1, nblocks, pstride, m, so, tmp);
  const int c = threadIdx.x;
  if (c < m) {
    const cx<R> den = so[c];
    const bool ok = st->active[c] && (abs2(den) > tiny_of<R>());
    st->alpha[c] = ok ? (st->rho[c] / den) : czero<R>();
    st->flags[c] = ok ? 1 : 0;  // bit0: this iteration is live for the column
  }
}

// after t = S s : omega = (t^H s)/(t^H t); rho' = -omega (rhat^H t); beta = (rho'/rho)(alpha/omega)
template <typename R>
__global__ void __launch_bounds__(1024) k_bicg_scal2(KrylovState<R>* st, const cx<R>* partial, int nblocks, int pstride, int m) {
  __shared__ cx<R> so[3 * FC_MAXCOLS];
  __shared__ cx<R> tmp[1024];
  reduce_partials<cx<R>>(partial, 3, nblocks, pstride, m, so, tmp);
  const int c = threadIdx.x;
  if (c < m) {
    const cx<R> ts = so[c], rht = so[2 * FC_MAXCOLS + c];
    const R tt = so[FC_MAXCOLS + c].x;
    const bool ok = st->flags[c] != 0;
    const bool ok2 = ok && (tt > tiny_of<R>());
    cx<R> om = ok2 ? mk<R>(ts.x / tt, ts.y / tt) : czero<R>();
    // Sleijpen & van der Vorst's safeguard: on indefinite shifted systems the minimal-residual omega can be
    // (nearly) zero, which wrecks the BiCG coefficients; keep |cos(t,s)| >= 0.7 by enlarging omega
    const R ss = so[FC_MAXCOLS + c].y;
    if (ok2 && ss > tiny_of<R>()) {
      const R cosang = sqrt(abs2(ts) / (tt * ss));
      if (cosang > R(0) && cosang < R(0.7)) om = om * (R(0.7) / cosang);
    }
    const cx<R> rho = st->rho[c];
    const cx<R> rho_new = -(om * rht);
    const bool ok3 = ok2 && (abs2(om) > tiny_of<R>()) && (abs2(rho) > tiny_of<R>());
    st->omega[c] = om;
    st->beta[c] = ok3 ? ((rho_new / rho) * (st->alpha[c] / om)) : czero<R>();
    if (ok3) st->rho[c] = rho_new;
    st->flags[c] = ok3 ? 1 : 0;
  }
}

// after the x/r/p update: ||r||, iteration count, convergence mask, number of live columns
template <typename R>
__global__ void __launch_bounds__(1024) k_bicg_scal3(KrylovState<R>* st, const R* partial, int nblocks, int pstride, int m) {
  __shared__ R so[FC_MAXCOLS];
  __shared__ R tmp[1024];
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  reduce_partials<R>(partial, 1, nblocks, pstride, m, so, tmp);
  const int c = threadIdx.x;
  if (c < m) {
    const R rn = sqrt(so[c]);
    const int was = st->active[c];
    if (was) { st->iters[c] += 1; st->rnorm[c] = rn; }
    int now = was && st->flags[c] && (rn > st->target[c]);
    // divergence guard: a column whose residual blew up (or went NaN) is parked; the host restores its iterate
    if (was && !(rn <= R(1e4) * st->rn0[c])) { now = 0; st->diverged[c] = 1; }
    st->active[c] = now;
    if (now) atomicAdd(&cnt, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) st->n_active = cnt;
}

// generic: out[s*m + c] = sum of partials (host-visible results: norms, Gram blocks)
template <typename T>
__global__ void __launch_bounds__(1024) k_reduce_partials(const T* partial, int nslots, int nblocks, int pstride, int m, T* out) {
  extern __shared__ unsigned char smraw[];
  T* so = reinterpret_cast<T*>(smraw);
  T* tmp = so + nslots * FC_MAXCOLS;
  reduce_partials<T>(partial, nslots, nblocks, pstride, m, so, tmp);
  for (int i = threadIdx.x; i < nslots * m; i += blockDim.x) out[i] = so[(i / m) * FC_MAXCOLS + (i % m)];
}

// ---- layout conversion at the ABI ---------------------------------------------------------------
// host/Julia: column-major n x m (ldh = n).  device: row-major n x ld.
template <typename TIN, typename R>
__device__ __forceinline__ cx<R> to_cx(TIN v);
template <> __device__ __forceinline__ cx<double> to_cx<double, double>(double v) { return mk<double>(v, 0.0); }
template <> __device__ __forceinline__ cx<double> to_cx<cx<double>, double>(cx<double> v) { return v; }
template <> __device__ __forceinline__ cx<float> to_cx<float, float>(float v) { return mk<float>(v, 0.f); }
template <> __device__ __forceinline__ cx<float> to_cx<cx<float>, float>(cx<float> v) { return v; }

template <typename TIN, typename R>
__global__ void __launch_bounds__(256) k_col2row(int64_t n, int m, int64_t ldh, int64_t ld, const TIN* __restrict__ src,
                                                 cx<R>* __restrict__ dst) {
  __shared__ cx<R> tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j;
    const int64_t r = r0 + tx;
    if (c < m && r < n) tile[j][tx] = to_cx<TIN, R>(src[(int64_t)c * ldh + r]);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + i;
    const int c = c0 + tx;
    if (c < m && r < n) dst[r * ld + c] = tile[tx][i];
  }
}

template <typename TOUT, typename R> __device__ __forceinline__ TOUT from_cx(cx<R> v);
template <> __device__ __forceinline__ double from_cx<double, double>(cx<double> v) { return v.x; }
template <> __device__ __forceinline__ cx<double> from_cx<cx<double>, double>(cx<double> v) { return v; }
template <> __device__ __forceinline__ float from_cx<float, float>(cx<float> v) { return v.x; }
template <> __device__ __forceinline__ cx<float> from_cx<cx<float>, float>(cx<float> v) { return v; }

template <typename TOUT, typename R>
__global__ void __launch_bounds__(256) k_row2col(int64_t n, int m, int64_t ld, int64_t ldh, const cx<R>* __restrict__ src,
                                                 TOUT* __restrict__ dst) {
  __shared__ cx<R> tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + i;
    const int c = c0 + tx;
    if (c < m && r < n) tile[i][tx] = src[r * ld + c];
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j;
    const int64_t r = r0 + tx;
    if (c < m && r < n) dst[(int64_t)c * ldh + r] = from_cx<TOUT, R>(tile[tx][j]);
  }
}

// ---- tall-skinny Gram: C[a x b] = X[:, :a]^H Y[:, :b], reduction over the n rows ------------------
// grid = (row chunks, ceil(a/64), ceil(b/64)); 256 threads = 16 x 16, each a 4 x 4 register tile.
// Per-chunk partial tiles go to `partial[chunk][a*b]` (row-major a x b) and are summed by k_sum_chunks.
template <typename R>
__global__ void __launch_bounds__(256) k_gram(int64_t n, int a, int b, int64_t ldx, int64_t ldy, const cx<R>* __restrict__ X,
                                              const cx<R>* __restrict__ Y, cx<R>* __restrict__ partial) {
  constexpr int TR = 16;
  __shared__ cx<R> xs[TR][64];
  __shared__ cx<R> ys[TR][64];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * 64, j0 = blockIdx.z * 64;
  int64_t rpb = (n + gridDim.x - 1) / gridDim.x;
  rpb = ((rpb + TR - 1) / TR) * TR;
  const int64_t rbeg = (int64_t)blockIdx.x * rpb, rend = min(n, rbeg + rpb);
  cx<R> acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = czero<R>();
  for (int64_t r0 = rbeg; r0 < rend; r0 += TR) {
    __syncthreads();
    for (int q = threadIdx.x; q < TR * 64; q += 256) {
      const int rr = q >> 6, cc = q & 63;
      const int64_t row = r0 + rr;
      xs[rr][cc] = (row < rend && i0 + cc < a) ? X[row * ldx + i0 + cc] : czero<R>();
      ys[rr][cc] = (row < rend && j0 + cc < b) ? Y[row * ldy + j0 + cc] : czero<R>();
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < TR; ++k) {
      cx<R> xv[4], yv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i] = xs[k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) yv[j] = ys[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) fma_conj_acc(acc[i][j], xv[i], yv[j]);
    }
  }
  cx<R>* out = partial + (int64_t)blockIdx.x * a * b;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ii = i0 + ty + 16 * i, jj = j0 + tx + 16 * j;
      if (ii < a && jj < b) out[(int64_t)ii * b + jj] = acc[i][j];
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_sum_chunks(int64_t len, int nchunks, const T* __restrict__ partial, T* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
    T s = zero_of<T>::v();
    for (int q = 0; q < nchunks; ++q) s = s + partial[(int64_t)q * len + i];
    out[i] = s;
  }
}

// ---- row transform: Y[:, :b] = X[:, :a] * T (T row-major a x b, leading dim ldt) ------------------
// grid = (ceil(n/64), ceil(b/64)); block 256 = 16 x 16; thread tile 4 rows x 4 cols; K tiled by 16.
template <typename R>
__global__ void __launch_bounds__(256) k_rowtransform(int64_t n, int a, int b, int64_t ldx, int64_t ldy, int ldt,
                                                      const cx<R>* __restrict__ X, const cx<R>* __restrict__ T,
                                                      cx<R>* __restrict__ Y) {
  constexpr int KT = 16;
  __shared__ cx<R> xs[64][KT + 1];
  __shared__ cx<R> ts[KT][64];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t r0 = (int64_t)blockIdx.x * 64;
  const int j0 = blockIdx.y * 64;
  cx<R> acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = czero<R>();
  for (int k0 = 0; k0 < a; k0 += KT) {
    __syncthreads();
    for (int q = threadIdx.x; q < 64 * KT; q += 256) {
      const int rr = q / KT, kk = q % KT;
      const int64_t row = r0 + rr;
      xs[rr][kk] = (row < n && k0 + kk < a) ? X[row * ldx + k0 + kk] : czero<R>();
    }
    for (int q = threadIdx.x; q < KT * 64; q += 256) {
      const int kk = q >> 6, cc = q & 63;
      ts[kk][cc] = (k0 + kk < a && j0 + cc < b) ? T[(int64_t)(k0 + kk) * ldt + j0 + cc] : czero<R>();
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < KT; ++k) {
      cx<R> xv[4], tv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i] = xs[ty + 16 * i][k];
#pragma unroll
      for (int j = 0; j < 4; ++j) tv[j] = ts[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) fma_acc(acc[i][j], xv[i], tv[j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = r0 + ty + 16 * i;
    if (row < n) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cc = j0 + tx + 16 * j;
        if (cc < b) Y[row * ldy + cc] = acc[i][j];
      }
    }
  }
}

}  // namespace feastcuda
