// Dense operator kernels: shifted-matrix formation, batched complex LU with partial pivoting (one batch entry per
// quadrature node), triangular solves on row-major right-hand-side blocks and the complex GEMM they all share.
//
// Replaces, for dense inputs, `_feast_dense_shifted_identity_minus!` / `z*B - A` (core/feast_aux.jl:59-74,
// dense/feast_dense.jl:190-194), `lu(shifted)` -> LAPACK zgetrf (dense/feast_dense.jl:196), `ldiv!` -> zgetrs (:207) and the
// projections `A*Qr`, `B*Qr` (:252,262).
//
// The GEMM runs on the FP64 tensor pipe: mma.sync.m8n8k4.f64 (DMMA); a complex product is four real MMAs on split
// re/im planes staged in shared memory.  tcgen05/TMEM has no FP64 kind, so this register-fragment path IS the FP64
// tensor-core path of sm_100a.  LU layout is LAPACK's: column-major n x n, unit-lower L below the diagonal, U on and above,
// pivots as row indices.  Bound: FP64 tensor pipe for the trailing updates (8/3 n^3 flop per node), HBM/latency for panels.
#pragma once
#include <cooperative_groups.h>

#include "cxmath.cuh"

namespace feastcuda {

typedef cx<double> zdd;

// real column-major input -> complex storage
__global__ void __launch_bounds__(256) k_dense_widen(int64_t total, const double* __restrict__ src, zdd* __restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) dst[i] = mk<double>(src[i], 0.0);
}

// ---- LU_b = z_b * B - A   (B == nullptr: identity) ------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dense_shift(int64_t n, const zdd* __restrict__ A, const zdd* __restrict__ B,
                                                     const zdd* __restrict__ z, zdd* __restrict__ LU, int64_t batch_stride) {
  const zdd zb = z[blockIdx.y];
  zdd* out = LU + (int64_t)blockIdx.y * batch_stride;
  const int64_t total = n * n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx % n, j = idx / n;
    zdd v = -A[idx];
    if (B != nullptr) v = v + zb * B[idx];
    else if (i == j) v = v + zb;
    out[idx] = v;
  }
}

// ---- panel factorisation: columns [k0, k0+nbw) of every batch entry, one CTA per entry ----------------------------------
// classic right-looking, partial pivoting by max(|re|+|im|) like izamax; row swaps are applied inside the panel only
// (k_dense_laswp does the rest of the row).  info[b] = first zero pivot (1-based) or 0.
constexpr int FC_LU_NB = 32;

__global__ void __launch_bounds__(512) k_dense_panel_lu(int n, int k0, int nbw, zdd* __restrict__ LU, int64_t batch_stride,
                                                        int* __restrict__ ipiv, int64_t piv_stride, int* __restrict__ info) {
  zdd* M = LU + (int64_t)blockIdx.x * batch_stride;
  int* piv = ipiv + (int64_t)blockIdx.x * piv_stride;
  __shared__ double s_val[512];
  __shared__ int s_idx[512];
  __shared__ zdd s_u[FC_LU_NB];
  __shared__ int s_p;
  const int tid = threadIdx.x, NT = blockDim.x;
  for (int j = 0; j < nbw; ++j) {
    const int col = k0 + j;
    double best = -1.0;
    int bi = col;
    for (int i = col + tid; i < n; i += NT) {
      const zdd v = M[i + (int64_t)col * n];
      const double a = fabs(v.x) + fabs(v.y);
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
      s_p = s_idx[0];
      piv[col] = s_idx[0];
      if (!(s_val[0] > 0.0) && info[blockIdx.x] == 0) info[blockIdx.x] = col + 1;
    }
    __syncthreads();
    const int p = s_p;
    if (p != col && tid < nbw) {
      const int64_t c = (int64_t)(k0 + tid) * n;
      const zdd t = M[col + c];
      M[col + c] = M[p + c];
      M[p + c] = t;
    }
    __syncthreads();
    const zdd pv = M[col + (int64_t)col * n];
    const bool ok = (fabs(pv.x) + fabs(pv.y)) > 0.0;
    const zdd inv = ok ? (mk<double>(1.0, 0.0) / pv) : czero<double>();
    if (tid < nbw) s_u[tid] = M[col + (int64_t)(k0 + tid) * n];
    __syncthreads();
    for (int i = col + 1 + tid; i < n; i += NT) {
      const zdd l = M[i + (int64_t)col * n] * inv;
      M[i + (int64_t)col * n] = l;
      for (int c = j + 1; c < nbw; ++c) {
        zdd* e = M + i + (int64_t)(k0 + c) * n;
        *e = *e - l * s_u[c];
      }
    }
    __syncthreads();
  }
}

// The same panel factorisation spread over `cpn` CTAs per batch entry (cooperative launch, two grid-wide barriers per
// column): every thread owns the rows i = (part*NT + tid) mod (cpn*NT) for the whole panel, so an update and the next pivot
// search never cross CTAs; CTA 0 of an entry performs the row swap.  For n = 8192 the single-CTA panel was the largest
// item of the LU (one SM streaming a 4 MB panel 32 times).
__global__ void __launch_bounds__(512) k_dense_panel_lu_coop(int n, int k0, int nbw, zdd* __restrict__ LU, int64_t batch_stride,
                                                             int* __restrict__ ipiv, int64_t piv_stride, int* __restrict__ info, int cpn,
                                                             double* __restrict__ pval, int* __restrict__ pidx) {
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const int node = blockIdx.x / cpn, part = blockIdx.x % cpn;
  zdd* M = LU + (int64_t)node * batch_stride;
  int* piv = ipiv + (int64_t)node * piv_stride;
  __shared__ double s_val[512];
  __shared__ int s_idx[512];
  __shared__ zdd s_u[FC_LU_NB];
  __shared__ int s_p;
  const int tid = threadIdx.x, NT = blockDim.x;
  const int own = part * NT + tid, stride = cpn * NT;
  for (int j = 0; j < nbw; ++j) {
    const int col = k0 + j;
    double best = -1.0;
    int bi = col;
    // first owned row >= col
    int i0 = own + ((col - own + stride - 1) / stride) * stride;
    if (own >= col) i0 = own;
    for (int i = i0; i < n; i += stride) {
      const zdd v = M[i + (int64_t)col * n];
      const double a = fabs(v.x) + fabs(v.y);
      if (a > best || (a == best && i < bi)) { best = a; bi = i; }
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
    if (tid == 0) { pval[blockIdx.x] = s_val[0]; pidx[blockIdx.x] = s_idx[0]; }
    grid.sync();
    if (tid == 0) {
      double bv = -1.0;
      int bp = col;
      for (int q = 0; q < cpn; ++q) {
        const double o = pval[node * cpn + q];
        const int oi = pidx[node * cpn + q];
        if (o > bv || (o == bv && oi < bp)) { bv = o; bp = oi; }
      }
      s_p = bp;
      if (part == 0) {
        piv[col] = bp;
        if (!(bv > 0.0) && info[node] == 0) info[node] = col + 1;
      }
    }
    __syncthreads();
    const int p = s_p;
    if (part == 0 && p != col && tid < nbw) {
      const int64_t c = (int64_t)(k0 + tid) * n;
      const zdd t = M[col + c];
      M[col + c] = M[p + c];
      M[p + c] = t;
    }
    grid.sync();
    const zdd pv = M[col + (int64_t)col * n];
    const bool ok = (fabs(pv.x) + fabs(pv.y)) > 0.0;
    const zdd inv = ok ? (mk<double>(1.0, 0.0) / pv) : czero<double>();
    if (tid < nbw) s_u[tid] = M[col + (int64_t)(k0 + tid) * n];
    __syncthreads();
    int i1 = own + ((col + 1 - own + stride - 1) / stride) * stride;
    if (own >= col + 1) i1 = own;
    for (int i = i1; i < n; i += stride) {
      const zdd l = M[i + (int64_t)col * n] * inv;
      M[i + (int64_t)col * n] = l;
      for (int c = j + 1; c < nbw; ++c) {
        zdd* e = M + i + (int64_t)(k0 + c) * n;
        *e = *e - l * s_u[c];
      }
    }
    __syncthreads();
  }
}

// row interchanges of the panel [k0, k0+nbw) applied to the columns [c0, c1) outside it; one thread per column
__global__ void __launch_bounds__(256) k_dense_laswp(int n, int k0, int nbw, int c0, int c1, zdd* __restrict__ LU,
                                                     int64_t batch_stride, const int* __restrict__ ipiv, int64_t piv_stride) {
  zdd* M = LU + (int64_t)blockIdx.y * batch_stride;
  const int* piv = ipiv + (int64_t)blockIdx.y * piv_stride;
  const int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= c1) return;
  zdd* colp = M + (int64_t)c * n;
  for (int j = 0; j < nbw; ++j) {
    const int r = k0 + j, p = piv[r];
    if (p != r) { const zdd t = colp[r]; colp[r] = colp[p]; colp[p] = t; }
  }
}

// U12 = L11^-1 A12 (unit lower L11 = panel rows/cols [k0, k0+nbw)) for the columns [c0, c1) of A12; one thread per column
__global__ void __launch_bounds__(128) k_dense_trsm_u12(int n, int k0, int nbw, int c0, int c1, zdd* __restrict__ LU, int64_t batch_stride) {
  zdd* M = LU + (int64_t)blockIdx.y * batch_stride;
  __shared__ zdd sL[FC_LU_NB][FC_LU_NB + 1];
  for (int e = threadIdx.x; e < nbw * nbw; e += blockDim.x) {
    const int i = e % nbw, j = e / nbw;
    sL[i][j] = M[(k0 + i) + (int64_t)(k0 + j) * n];
  }
  __syncthreads();
  const int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= c1) return;
  zdd* colp = M + (int64_t)c * n + k0;
  for (int i = 1; i < nbw; ++i) {
    zdd s = colp[i];
    for (int j = 0; j < i; ++j) s = s - sL[i][j] * colp[j];
    colp[i] = s;
  }
}

// ipiv (sequence of transpositions) -> perm with  (P b)[i] = b[perm[i]]
__global__ void k_dense_piv_to_perm(int n, const int* __restrict__ ipiv, int* __restrict__ perm, int64_t piv_stride) {
  const int* piv = ipiv + (int64_t)blockIdx.x * piv_stride;
  int* pm = perm + (int64_t)blockIdx.x * piv_stride;
  if (threadIdx.x != 0) return;
  for (int i = 0; i < n; ++i) pm[i] = i;
  for (int i = 0; i < n; ++i) {
    const int p = piv[i];
    if (p != i) { const int t = pm[i]; pm[i] = pm[p]; pm[p] = t; }
  }
}

// X[i, :] = RHS[perm[i], :]   (row-major n x ld blocks, m active columns)
// blockIdx.y = batch entry (one per quadrature node): perm and X advance by their batch strides, RHS is shared
__global__ void __launch_bounds__(256) k_dense_gather_rows(int64_t n, int m, int64_t ld, const int* __restrict__ perm, int64_t piv_stride,
                                                           const zdd* __restrict__ RHS, zdd* __restrict__ X, int64_t xbatch) {
  perm += (int64_t)blockIdx.y * piv_stride;
  X += (int64_t)blockIdx.y * xbatch;
  const int64_t total = n * m;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / m;
    const int c = (int)(idx % m);
    X[i * ld + c] = RHS[(int64_t)perm[i] * ld + c];
  }
}

// In-place triangular solve of one diagonal block on a row-major block of right-hand sides:
//   LOWER: rows [r0, r0+bs) of X <- L11^-1 X (unit diagonal);  UPPER: X <- U11^-1 X.  One thread per RHS column.
template <bool LOWER>
__global__ void __launch_bounds__(128) k_dense_trsm_rows(int n, int r0, int bs, const zdd* __restrict__ LU, int64_t lubatch, int m,
                                                         int64_t ld, zdd* __restrict__ X, int64_t xbatch) {
  LU += (int64_t)blockIdx.y * lubatch;
  X += (int64_t)blockIdx.y * xbatch;
  __shared__ zdd sT[FC_LU_NB][FC_LU_NB + 1];
  for (int e = threadIdx.x; e < bs * bs; e += blockDim.x) {
    const int i = e % bs, j = e / bs;
    sT[i][j] = LU[(r0 + i) + (int64_t)(r0 + j) * n];
  }
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  zdd* xp = X + (int64_t)r0 * ld + c;
  if (LOWER) {
    for (int i = 1; i < bs; ++i) {
      zdd s = xp[(int64_t)i * ld];
      for (int j = 0; j < i; ++j) s = s - sT[i][j] * xp[(int64_t)j * ld];
      xp[(int64_t)i * ld] = s;
    }
  } else {
    for (int i = bs - 1; i >= 0; --i) {
      zdd s = xp[(int64_t)i * ld];
      for (int j = i + 1; j < bs; ++j) s = s - sT[i][j] * xp[(int64_t)j * ld];
      xp[(int64_t)i * ld] = s / sT[i][i];
    }
  }
}

// ---- complex GEMM on the FP64 tensor pipe ---------------------------------------------------------------------------------
// C(M x N) = beta * C + alpha * A(M x K) * B(K x N), alpha = +-1, beta in {0, 1}; every operand is addressed with (row stride,
// column stride) so that column-major LU blocks and row-major vector blocks mix freely.  CTA tile 64 x 64, K tile 16,
// 8 warps as 4 (M) x 2 (N), warp tile 16 x 32 = 2 x 4 mma tiles of m8n8k4.
struct ZgemmArgs {
  int M, N, K;
  const zdd* A; int64_t ars, acs, abatch;
  const zdd* B; int64_t brs, bcs, bbatch;
  zdd* C; int64_t crs, ccs, cbatch;
  double alpha;
  int beta;
};

__device__ __forceinline__ void dmma_8x8x4(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) k_zgemm_dmma(ZgemmArgs g) {
  constexpr int TM = 64, TN = 64, TK = 16, LDS = 72;   // LDS = 8 mod 16: conflict-free fragment reads
  __shared__ double sAr[TK][LDS], sAi[TK][LDS], sBr[TK][LDS], sBi[TK][LDS];
  const zdd* A = g.A + (int64_t)blockIdx.z * g.abatch;
  const zdd* B = g.B + (int64_t)blockIdx.z * g.bbatch;
  zdd* C = g.C + (int64_t)blockIdx.z * g.cbatch;
  const int i0 = blockIdx.x * TM, j0 = blockIdx.y * TN;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;
  const int fr = lane >> 2, fk = lane & 3;
  double cr[2][4][2], ci[2][4][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) { cr[a][b][0] = cr[a][b][1] = 0.0; ci[a][b][0] = ci[a][b][1] = 0.0; }
  const bool a_mfast = (g.ars == 1), b_nfast = (g.bcs == 1);
  for (int k0 = 0; k0 < g.K; k0 += TK) {
    __syncthreads();
    for (int e = tid; e < TM * TK; e += 256) {
      const int mm = a_mfast ? (e % TM) : (e / TK), kk = a_mfast ? (e / TM) : (e % TK);
      zdd v = czero<double>();
      if (i0 + mm < g.M && k0 + kk < g.K) v = A[(int64_t)(i0 + mm) * g.ars + (int64_t)(k0 + kk) * g.acs];
      sAr[kk][mm] = v.x;
      sAi[kk][mm] = v.y;
    }
    for (int e = tid; e < TN * TK; e += 256) {
      const int nn = b_nfast ? (e % TN) : (e / TK), kk = b_nfast ? (e / TN) : (e % TK);
      zdd v = czero<double>();
      if (j0 + nn < g.N && k0 + kk < g.K) v = B[(int64_t)(k0 + kk) * g.brs + (int64_t)(j0 + nn) * g.bcs];
      sBr[kk][nn] = v.x;
      sBi[kk][nn] = v.y;
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < TK; ks += 4) {
      double ar[2], ai[2], nai[2], br[4], bi[4];
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        ar[a] = sAr[ks + fk][wm + 8 * a + fr];
        ai[a] = sAi[ks + fk][wm + 8 * a + fr];
        nai[a] = -ai[a];
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        br[b] = sBr[ks + fk][wn + 8 * b + fr];
        bi[b] = sBi[ks + fk][wn + 8 * b + fr];
      }
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          dmma_8x8x4(cr[a][b], ar[a], br[b]);
          dmma_8x8x4(cr[a][b], nai[a], bi[b]);
          dmma_8x8x4(ci[a][b], ar[a], bi[b]);
          dmma_8x8x4(ci[a][b], ai[a], br[b]);
        }
    }
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int i = i0 + wm + 8 * a + fr, j = j0 + wn + 8 * b + 2 * fk + q;
        if (i < g.M && j < g.N) {
          zdd* cp = C + (int64_t)i * g.crs + (int64_t)j * g.ccs;
          zdd v = mk<double>(g.alpha * cr[a][b][q], g.alpha * ci[a][b][q]);
          if (g.beta) v = v + *cp;
          *cp = v;
        }
      }
}

// ---- the same GEMM with an asynchronous, double-buffered operand pipeline -----------------------------------------------
// Operands stay interleaved (re, im) in shared memory and are staged by cp.async (LDGSTS, 16 bytes = one complex entry per
// request, zero-filled outside the matrix), two K tiles in flight; a fragment read is one LDS.128 that delivers both planes.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}

__global__ void __launch_bounds__(256, 2) k_zgemm_dmma_async(ZgemmArgs g) {
  constexpr int TM = 64, TN = 64, TK = 16;
  extern __shared__ __align__(16) unsigned char zg_smem[];
  zdd (*sA)[TK][TM] = reinterpret_cast<zdd (*)[TK][TM]>(zg_smem);                                   // [2][TK][TM]
  zdd (*sB)[TK][TN] = reinterpret_cast<zdd (*)[TK][TN]>(zg_smem + 2 * TK * TM * sizeof(zdd));       // [2][TK][TN]
  const zdd* A = g.A + (int64_t)blockIdx.z * g.abatch;
  const zdd* B = g.B + (int64_t)blockIdx.z * g.bbatch;
  zdd* C = g.C + (int64_t)blockIdx.z * g.cbatch;
  const int i0 = blockIdx.x * TM, j0 = blockIdx.y * TN;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;
  const int fr = lane >> 2, fk = lane & 3;
  double cr[2][4][2], ci[2][4][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) { cr[a][b][0] = cr[a][b][1] = 0.0; ci[a][b][0] = ci[a][b][1] = 0.0; }
  const bool a_mfast = (g.ars == 1), b_nfast = (g.bcs == 1);
  auto issue = [&](int st, int k0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = tid + q * 256;
      const int mm = a_mfast ? (e % TM) : (e / TK), ka = a_mfast ? (e / TM) : (e % TK);
      const bool va = (i0 + mm < g.M) && (k0 + ka < g.K);
      cp_async16(&sA[st][ka][mm], va ? (A + (int64_t)(i0 + mm) * g.ars + (int64_t)(k0 + ka) * g.acs) : A, va);
      const int nn = b_nfast ? (e % TN) : (e / TK), kb = b_nfast ? (e / TN) : (e % TK);
      const bool vb = (j0 + nn < g.N) && (k0 + kb < g.K);
      cp_async16(&sB[st][kb][nn], vb ? (B + (int64_t)(k0 + kb) * g.brs + (int64_t)(j0 + nn) * g.bcs) : B, vb);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue(0, 0);
  int st = 0;
  for (int k0 = 0; k0 < g.K; k0 += TK) {
    const bool more = k0 + TK < g.K;
    if (more) {
      issue(st ^ 1, k0 + TK);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < TK; ks += 4) {
      zdd av[2], bv[4];
#pragma unroll
      for (int a = 0; a < 2; ++a) av[a] = sA[st][ks + fk][wm + 8 * a + fr];
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = sB[st][ks + fk][wn + 8 * b + fr];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          dmma_8x8x4(cr[a][b], av[a].x, bv[b].x);
          dmma_8x8x4(cr[a][b], -av[a].y, bv[b].y);
          dmma_8x8x4(ci[a][b], av[a].x, bv[b].y);
          dmma_8x8x4(ci[a][b], av[a].y, bv[b].x);
        }
    }
    __syncthreads();
    st ^= 1;
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int i = i0 + wm + 8 * a + fr, j = j0 + wn + 8 * b + 2 * fk + q;
        if (i < g.M && j < g.N) {
          zdd* cp = C + (int64_t)i * g.crs + (int64_t)j * g.ccs;
          zdd v = mk<double>(g.alpha * cr[a][b][q], g.alpha * ci[a][b][q]);
          if (g.beta) v = v + *cp;
          *cp = v;
        }
      }
}

}  // namespace feastcuda
