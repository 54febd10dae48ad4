// Multi-RHS CSR kernels: Y = zb*(B X) + za*(A X) on ROW-MAJOR block vectors (n x ld, the m active
// columns of a row are contiguous), with the Krylov dot products fused into the same pass.
//
// Replaces the reference's per-column operator (sparse/feast_sparse.jl:20-27 mul! of
// SparseShiftedOperator, :142-148 identity form) and the mul!(rhs,B,basis) / mul!(aq,A,q) SpMMs
// (sparse/feast_sparse.jl:331,392,402) and the residual loop (:449-462).
//
// Mapping: a group of G lanes owns one matrix row; lane g of the group owns columns g, g+G, ...
// (NC of them), so a row of X is read as one contiguous 16*m-byte segment per stored entry.  The
// group's lanes fetch G (col,val) pairs of the row with one coalesced load and broadcast them with
// shuffles.  HBM-bound: per launch it must move nnz*(sizeof(val)+4) + 2*n*m*16 bytes (plus 16*n*m per
// extra fused operand); the gather hits L1/L2 for stencil-like patterns.
#pragma once
#include "cxmath.cuh"

namespace feastcuda {

enum SpmmMode {
  SPMM_PLAIN = 0,     // Y = op(X)
  SPMM_DOT_RHAT = 1,  // Y = op(X);  d0 = rhat^H y                       (BiCGStab: v = S p)
  SPMM_DOT_TS = 2,    // Y = op(X);  d0 = y^H x, d1 = (y^H y, x^H x), d2 = rhat^H y (BiCGStab: t = S s)
  SPMM_EIGRES = 3,    // no store;   y = A x - lam_c (B x);  d0 = |y|^2    (FEAST residuals)
  SPMM_RESID = 4      // Y = aux - op(X);  d0 = |y|^2                      (true linear residual)
};

template <typename R, typename TA>
struct SpmmArgs {
  int64_t n;
  int m;
  int64_t ld;
  const int* a_ptr; const int* a_col; const TA* a_val;  // a_ptr == nullptr: skip A
  const int* b_ptr; const int* b_col; const TA* b_val;  // b_ptr == nullptr: B = I
  int skip_b;                                           // 1: zb term absent altogether
  cx<R> za, zb;
  const cx<R>* X;
  cx<R>* Y;
  const cx<R>* aux;   // rhat (modes 1,2) or RHS (mode 4)
  const cx<R>* lam;   // per-column lambda (mode 3)
  cx<R>* partial;     // [slot][block][pstride]
  int pstride;
};

template <int MODE> struct spmm_ndots { static constexpr int v = (MODE == SPMM_PLAIN) ? 0 : (MODE == SPMM_DOT_TS ? 3 : 1); };

template <typename R, typename TA, int G, int NC>
__device__ __forceinline__ void csr_row_gather(const int* __restrict__ ptr, const int* __restrict__ col,
                                               const TA* __restrict__ val, int64_t row, bool valid,
                                               const cx<R>* __restrict__ X, int64_t ld, int m, int g,
                                               unsigned gmask, cx<R> (&acc)[NC]) {
  constexpr int U = (G >= 4) ? 4 : G;
  int p0 = 0, p1 = 0;
  if (valid) { p0 = ptr[row]; p1 = ptr[row + 1]; }
  for (int pb = p0; pb < p1; pb += G) {
    const int cnt = min(G, p1 - pb);
    int myj = (int)row;
    TA mya = zero_of<TA>::v();
    if (g < cnt) { myj = col[pb + g]; mya = val[pb + g]; }
    for (int t = 0; t < cnt; t += U) {
      int jj[U];
      TA aa[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        jj[u] = __shfl_sync(gmask, myj, t + u, G);
        if constexpr (sizeof(TA) == sizeof(R)) aa[u] = __shfl_sync(gmask, mya, t + u, G);
        else aa[u] = shfl_cx(gmask, mya, t + u, G);
      }
      cx<R> xv[U][NC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const cx<R>* xr = X + (int64_t)jj[u] * ld;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          const int c = g + G * k;
          xv[u][k] = (c < m) ? xr[c] : czero<R>();
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int k = 0; k < NC; ++k) fma_acc(acc[k], aa[u], xv[u][k]);
    }
  }
}

template <typename R, typename TA, int G, int NC, int MODE>
__global__ void __launch_bounds__(256) k_spmm(SpmmArgs<R, TA> a) {
  constexpr int RPW = 32 / G;
  constexpr int ND = spmm_ndots<MODE>::v;
  const int lane = threadIdx.x & 31, g = lane % G, sub = lane / G;
  constexpr unsigned gm0 = (G >= 32) ? 0xffffffffu : ((1u << (G & 31)) - 1u);
  const unsigned gmask = gm0 << (sub * G);
  const int wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int64_t step = (int64_t)wpb * RPW;
  int64_t rpb = (a.n + gridDim.x - 1) / gridDim.x;
  rpb = ((rpb + step - 1) / step) * step;
  const int64_t row_begin = (int64_t)blockIdx.x * rpb;
  const int64_t row_end = min(a.n, row_begin + rpb);

  cx<R> dots[ND > 0 ? ND : 1][NC];
#pragma unroll
  for (int d = 0; d < (ND > 0 ? ND : 1); ++d)
#pragma unroll
    for (int k = 0; k < NC; ++k) dots[d][k] = czero<R>();
  cx<R> lamc[NC];
  if constexpr (MODE == SPMM_EIGRES) {
#pragma unroll
    for (int k = 0; k < NC; ++k) { const int c = g + G * k; lamc[k] = (c < a.m) ? a.lam[c] : czero<R>(); }
  }

  for (int64_t rb = row_begin + (int64_t)wib * RPW; rb < row_end; rb += step) {
    const int64_t row = rb + sub;
    const bool valid = row < row_end;
    cx<R> accA[NC], accB[NC], xo[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) { accA[k] = czero<R>(); accB[k] = czero<R>(); xo[k] = czero<R>(); }
    const bool need_own = (MODE == SPMM_DOT_TS) || (a.b_ptr == nullptr && !a.skip_b);
    if (need_own && valid) {
      const cx<R>* xr = a.X + row * a.ld;
#pragma unroll
      for (int k = 0; k < NC; ++k) { const int c = g + G * k; if (c < a.m) xo[k] = xr[c]; }
    }
    if (a.a_ptr != nullptr) csr_row_gather<R, TA, G, NC>(a.a_ptr, a.a_col, a.a_val, row, valid, a.X, a.ld, a.m, g, gmask, accA);
    if (!a.skip_b) {
      if (a.b_ptr != nullptr) csr_row_gather<R, TA, G, NC>(a.b_ptr, a.b_col, a.b_val, row, valid, a.X, a.ld, a.m, g, gmask, accB);
      else {
#pragma unroll
        for (int k = 0; k < NC; ++k) accB[k] = xo[k];
      }
    }
    if (valid) {
      cx<R> y[NC];
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        if constexpr (MODE == SPMM_EIGRES) y[k] = accA[k] - lamc[k] * accB[k];
        else y[k] = a.za * accA[k] + a.zb * accB[k];
      }
      if constexpr (MODE == SPMM_RESID) {
        const cx<R>* br = a.aux + row * a.ld;
#pragma unroll
        for (int k = 0; k < NC; ++k) { const int c = g + G * k; y[k] = ((c < a.m) ? br[c] : czero<R>()) - y[k]; }
      }
      if constexpr (MODE != SPMM_EIGRES) {
        cx<R>* yr = a.Y + row * a.ld;
#pragma unroll
        for (int k = 0; k < NC; ++k) { const int c = g + G * k; if (c < a.m) yr[c] = y[k]; }
      }
      if constexpr (MODE == SPMM_DOT_RHAT || MODE == SPMM_DOT_TS) {
        const cx<R>* hr = a.aux + row * a.ld;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          const int c = g + G * k;
          if (c < a.m) {
            const cx<R> h = hr[c];
            if constexpr (MODE == SPMM_DOT_RHAT) fma_conj_acc(dots[0][k], h, y[k]);
            else {
              fma_conj_acc(dots[0][k], y[k], xo[k]);
              dots[1][k].x += abs2(y[k]);
              dots[1][k].y += abs2(xo[k]);
              fma_conj_acc(dots[2][k], h, y[k]);
            }
          }
        }
      }
      if constexpr (MODE == SPMM_EIGRES || MODE == SPMM_RESID) {
#pragma unroll
        for (int k = 0; k < NC; ++k) { const int c = g + G * k; if (c < a.m) dots[0][k].x += abs2(y[k]); }
      }
    }
  }

  if constexpr (ND > 0) {
    __shared__ cx<R> red[8 * 32 * NC];
    const int ngroups = wpb * RPW;
    const int width = G * NC;
    for (int d = 0; d < ND; ++d) {
      __syncthreads();
#pragma unroll
      for (int k = 0; k < NC; ++k) red[(wib * RPW + sub) * width + g + G * k] = dots[d][k];
      __syncthreads();
      for (int c = threadIdx.x; c < width; c += blockDim.x) {
        if (c < a.m) {
          cx<R> s = czero<R>();
          for (int q = 0; q < ngroups; ++q) s = s + red[q * width + c];
          a.partial[((int64_t)d * gridDim.x + blockIdx.x) * a.pstride + c] = s;
        }
      }
    }
  }
}

}  // namespace feastcuda
