// Elementwise kernels of the two-sided (non-Hermitian) multi-shift Lanczos filter for general pencils (feastcuda.cu: msl2_filter).
// Blocks are compact complex: n x ld doubles with ld = 2 * nc, element e of a row = complex column e (re, im); a thread owns one column,
// rows strided.  Per-column complex scalars travel BY VALUE in the launch parameters (<= 64 columns per launch): the recurrence scalars
// are advanced on the host (two tiny reductions per step against ~30 SpMM launches), so no scalar ever needs a host->device copy.
#pragma once
#include "kernels_lanczos.cuh"

namespace feastcuda {

constexpr int TS_MAXC = 64;
struct TsScal { double2 a[TS_MAXC]; double2 b[TS_MAXC]; };

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// out = x0 - a_c x1 - b_c x2   (out may alias any operand)
__global__ void __launch_bounds__(256) k_ts_lin3(int64_t n, int nc, int pp, int64_t ld, const double* x0, const double* x1, const double* x2,
                                                 double* out, TsScal s) {
  EwMap2 e(pp);
  if (e.pc >= nc) return;
  const double2 a = s.a[e.pc], b = s.b[e.pc];
  for (int64_t row = (int64_t)blockIdx.x * e.rpb + e.rsub; row < n; row += (int64_t)gridDim.x * e.rpb) {
    const int64_t off = row * ld + 2 * e.pc;
    const double2 v0 = ldg2(x0 + off), v1 = ldg2(x1 + off), v2 = ldg2(x2 + off);
    const double2 t1 = cmul(a, v1), t2 = cmul(b, v2);
    stg2(out + off, make_double2(v0.x - t1.x - t2.x, v0.y - t1.y - t2.y));
  }
}

// v *= a_c.x, w *= b_c.x (real factors)
__global__ void __launch_bounds__(256) k_ts_scale2(int64_t n, int nc, int pp, int64_t ld, double* v, double* w, TsScal s) {
  EwMap2 e(pp);
  if (e.pc >= nc) return;
  const double fa = s.a[e.pc].x, fb = s.b[e.pc].x;
  for (int64_t row = (int64_t)blockIdx.x * e.rpb + e.rsub; row < n; row += (int64_t)gridDim.x * e.rpb) {
    const int64_t off = row * ld + 2 * e.pc;
    double2 a = ldg2(v + off);
    stg2(v + off, make_double2(fa * a.x, fa * a.y));
    if (w != nullptr) {
      double2 b = ldg2(w + off);
      stg2(w + off, make_double2(fb * b.x, fb * b.y));
    }
  }
}

// Q += a_c v
__global__ void __launch_bounds__(256) k_ts_caxpy(int64_t n, int nc, int pp, int64_t ld, const double* v, double* Q, TsScal s) {
  EwMap2 e(pp);
  if (e.pc >= nc) return;
  const double2 a = s.a[e.pc];
  for (int64_t row = (int64_t)blockIdx.x * e.rpb + e.rsub; row < n; row += (int64_t)gridDim.x * e.rpb) {
    const int64_t off = row * ld + 2 * e.pc;
    const double2 t = cmul(a, ldg2(v + off));
    double2 q = ldg2(Q + off);
    stg2(Q + off, make_double2(q.x + t.x, q.y + t.y));
  }
}

// NQ real sums per column, written as partial[(q * gridDim.x + blockIdx.x) * pstride + c] (the layout k_reduce_partials sums)
template <int NQ>
__device__ __forceinline__ void ts_block_reduce(const double (&v)[NQ], int pp, int nc, double* partial, int pstride) {
  __shared__ double red[256 * NQ];
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NQ; ++q) red[q * 256 + threadIdx.x] = v[q];
  __syncthreads();
  for (int i = threadIdx.x; i < NQ * pp; i += blockDim.x) {
    const int q = i / pp, c = i % pp;
    if (c < nc) {
      double s = 0.0;
      for (int t = c; t < (int)blockDim.x; t += pp) s += red[q * 256 + t];
      partial[((int64_t)q * gridDim.x + blockIdx.x) * pstride + c] = s;
    }
  }
}

// conj(x) . y per column: slot 0 = real part, slot 1 = imaginary part
__global__ void __launch_bounds__(256) k_ts_dot(int64_t n, int nc, int pp, int64_t ld, const double* x, const double* y, double* partial,
                                                int pstride) {
  EwMap2 e(pp);
  double acc[2] = {0.0, 0.0};
  if (e.pc < nc)
    for (int64_t row = (int64_t)blockIdx.x * e.rpb + e.rsub; row < n; row += (int64_t)gridDim.x * e.rpb) {
      const int64_t off = row * ld + 2 * e.pc;
      const double2 a = ldg2(x + off), b = ldg2(y + off);
      acc[0] += a.x * b.x + a.y * b.y;
      acc[1] += a.x * b.y - a.y * b.x;
    }
  ts_block_reduce<2>(acc, pp, nc, partial, pstride);
}

// slots: 0 |v|^2, 1 |w|^2, 2 Re conj(w).v, 3 Im conj(w).v
__global__ void __launch_bounds__(256) k_ts_dot3(int64_t n, int nc, int pp, int64_t ld, const double* v, const double* w, double* partial,
                                                 int pstride) {
  EwMap2 e(pp);
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  if (e.pc < nc)
    for (int64_t row = (int64_t)blockIdx.x * e.rpb + e.rsub; row < n; row += (int64_t)gridDim.x * e.rpb) {
      const int64_t off = row * ld + 2 * e.pc;
      const double2 a = ldg2(v + off), b = ldg2(w + off);
      acc[0] += a.x * a.x + a.y * a.y;
      acc[1] += b.x * b.x + b.y * b.y;
      acc[2] += b.x * a.x + b.y * a.y;
      acc[3] += b.x * a.y - b.y * a.x;
    }
  ts_block_reduce<4>(acc, pp, nc, partial, pstride);
}

}  // namespace feastcuda
