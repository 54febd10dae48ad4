// Banded operator kernels: shifted band formation in LAPACK gbtrf layout, band LU with partial pivoting, band solve on
// row-major right-hand-side blocks, band mat-vec.
//
// Replaces `fill_shifted_banded!` (banded/feast_banded.jl:216-237,273-296,511-559), `LAPACK.gbtrf!` / `gbtrs!`
// (banded/feast_banded.jl:108,141,678,683) and the band mat-vecs (banded/feast_banded.jl:239-259,298-314).
// Operators arrive expanded to general band storage (2k+1) x n (diagonal in row k); factors use ldf = 3k+1 rows with
// kl = ku = k: A[i,j] at F[(2k + i - j) + j*ldf], the top k rows are the fill-in workspace of the pivoted factorisation.
// Bound: latency (a band factorisation is n dependent steps) for the LU, HBM for the solve over M0 columns.
#pragma once
#include "cxmath.cuh"

namespace feastcuda {

typedef cx<double> zdb;

__global__ void __launch_bounds__(256) k_band_shift(int n, int k, int ka, int kb, const zdb* __restrict__ A, const zdb* __restrict__ B,
                                                    zdb z, zdb* __restrict__ F) {
  const int ldf = 3 * k + 1;
  const int64_t total = (int64_t)ldf * n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(idx % ldf), j = (int)(idx / ldf);
    zdb v = czero<double>();
    const int i = j + r - 2 * k;
    if (r >= k && i >= 0 && i < n) {
      const int d = i - j;
      if (d >= -ka && d <= ka) v = v - A[(ka + d) + (int64_t)j * (2 * ka + 1)];
      if (B != nullptr) {
        if (d >= -kb && d <= kb) v = v + z * B[(kb + d) + (int64_t)j * (2 * kb + 1)];
      } else if (d == 0) v = v + z;
    }
    F[idx] = v;
  }
}

// unblocked band LU (zgbtf2) with kl = ku = k; one CTA; info = first zero pivot (1-based) or 0
__global__ void __launch_bounds__(256) k_band_lu(int n, int k, zdb* __restrict__ F, int* __restrict__ ipiv, int* __restrict__ info) {
  const int ldf = 3 * k + 1, kv = 2 * k;
  const int tid = threadIdx.x, NT = blockDim.x;
  __shared__ double s_val[256];
  __shared__ int s_idx[256];
  __shared__ int s_jp, s_ju;
  if (tid == 0) s_ju = 0;
  __syncthreads();
  for (int j = 0; j < n; ++j) {
    const int km = min(k, n - 1 - j);
    zdb* colj = F + (int64_t)j * ldf + kv;   // colj[i] = A[j+i, j]
    double best = -1.0;
    int bi = 0;
    for (int i = tid; i <= km; i += NT) {
      const double a = fabs(colj[i].x) + fabs(colj[i].y);
      if (a > best) { best = a; bi = i; }
    }
    s_val[tid] = best;
    s_idx[tid] = bi;
    __syncthreads();
    for (int w = NT / 2; w > 0; w >>= 1) {
      if (tid < w) {
        const double o = s_val[tid + w];
        const int oi = s_idx[tid + w];
        if (o > s_val[tid] || (o == s_val[tid] && oi < s_idx[tid])) { s_val[tid] = o; s_idx[tid] = oi; }
      }
      __syncthreads();
    }
    if (tid == 0) {
      s_jp = s_idx[0];
      ipiv[j] = j + s_idx[0];
      s_ju = max(s_ju, min(j + k + s_idx[0], n - 1));
      if (!(s_val[0] > 0.0) && *info == 0) *info = j + 1;
    }
    __syncthreads();
    const int jp = s_jp, ju = s_ju;
    if (jp != 0) {
      for (int c = j + tid; c <= ju; c += NT) {   // swap rows j and j+jp over the columns j..ju
        zdb* e0 = F + (int64_t)c * ldf + (kv + j - c);
        const zdb t = e0[0];
        e0[0] = e0[jp];
        e0[jp] = t;
      }
    }
    __syncthreads();
    const zdb pv = colj[0];
    const bool ok = (fabs(pv.x) + fabs(pv.y)) > 0.0;
    const zdb inv = ok ? (mk<double>(1.0, 0.0) / pv) : czero<double>();
    for (int i = 1 + tid; i <= km; i += NT) colj[i] = colj[i] * inv;
    __syncthreads();
    const int nc = ju - j;
    for (int e = tid; e < nc * km; e += NT) {
      const int cc = e / km + 1, i = e % km + 1;       // column j+cc, row j+i
      zdb* colc = F + (int64_t)(j + cc) * ldf + (kv - cc);   // colc[i] = A[j+i, j+cc]
      colc[i] = colc[i] - colj[i] * colc[0];
    }
    __syncthreads();
  }
}

// band solve (zgbtrs, no transpose) on a row-major block: one thread per right-hand-side column
__global__ void __launch_bounds__(64) k_band_solve(int n, int k, const zdb* __restrict__ F, const int* __restrict__ ipiv, int m,
                                                   int64_t ld, zdb* __restrict__ X) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  const int ldf = 3 * k + 1, kv = 2 * k;
  zdb* x = X + c;
  for (int j = 0; j < n; ++j) {
    const int lm = min(k, n - 1 - j), p = ipiv[j];
    if (p != j) { const zdb t = x[(int64_t)p * ld]; x[(int64_t)p * ld] = x[(int64_t)j * ld]; x[(int64_t)j * ld] = t; }
    const zdb xj = x[(int64_t)j * ld];
    const zdb* colj = F + (int64_t)j * ldf + kv;
    for (int i = 1; i <= lm; ++i) x[(int64_t)(j + i) * ld] = x[(int64_t)(j + i) * ld] - colj[i] * xj;
  }
  for (int j = n - 1; j >= 0; --j) {
    const zdb* colj = F + (int64_t)j * ldf + kv;
    const zdb xj = x[(int64_t)j * ld] / colj[0];
    x[(int64_t)j * ld] = xj;
    const int i0 = max(0, j - kv);
    for (int i = i0; i < j; ++i) x[(int64_t)i * ld] = x[(int64_t)i * ld] - colj[i - j] * xj;
  }
}

// Y = A X for a general band matrix (2k+1) x n, row-major blocks
__global__ void __launch_bounds__(256) k_band_apply(int n, int k, const zdb* __restrict__ AB, int m, int64_t ld,
                                                    const zdb* __restrict__ X, zdb* __restrict__ Y) {
  const int64_t total = (int64_t)n * m;
  const int rows = 2 * k + 1;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / m), c = (int)(idx % m);
    zdb s = czero<double>();
    const int j0 = max(0, i - k), j1 = min(n - 1, i + k);
    for (int j = j0; j <= j1; ++j) fma_acc(s, AB[(k + i - j) + (int64_t)j * rows], X[(int64_t)j * ld + c]);
    Y[(int64_t)i * ld + c] = s;
  }
}

}  // namespace feastcuda
