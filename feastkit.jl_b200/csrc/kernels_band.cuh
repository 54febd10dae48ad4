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

// =====================================================================================================
// Batched-over-nodes variants (round 2; the kernels above remain for wide bands and as the A/B reference, FEASTCUDA_BAND_IMPL=1).
// A band factorisation is n dependent steps, so the parallelism is ACROSS quadrature nodes and right-hand-side columns: every node of a
// sweep is factored by its own warp in ONE launch, and all (node, 32-column group) pairs are solved by their own warp in ONE launch.
// CPU restatement: oracle/feast_port.py band_lu / band_solve_window.
// =====================================================================================================
constexpr int FC_BAND_BATCH = 32;      // nodes per launch (pointers travel by value in the launch parameters)
constexpr int FC_BAND_LU_MAXK = 32;    // widest half-bandwidth served by the warp LU
constexpr int FC_BAND_WIN_MAXK = 16;   // widest half-bandwidth served by the register-window solve

struct BandBatch {
  zdb* F[FC_BAND_BATCH];      // factor (3k+1) x n of node q
  int* ipiv[FC_BAND_BATCH];   // n pivots + the info word
};

__device__ __forceinline__ void band_cp16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void band_cp4(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void band_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void band_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void band_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// The warp LU with the active window in SHARED memory: columns j .. j + 2k (all a step can touch) live in a ring of W column slots
// (slot = column mod W); a column enters by cp.async W - 2k - 1 steps before its first possible use and returns to global memory
// right after its own pivot step.  In k_band_lu_warp every phase of a step (pivot search, swap, scaling, update) reads what the
// previous phase stored through L2 (stores do not allocate in L1): 4 400 cycles per column at k = 7; here the phases meet in
// shared memory.  KMAX = 7 (W = 32, 11 KB) or 15 (W = 64, 46 KB).
template <int KMAX, int W>
__global__ void __launch_bounds__(32) k_band_lu_smem(int n, int k, BandBatch bb) {
  constexpr int LDFMAX = 3 * KMAX + 1, PEND = W - 2 * KMAX - 1;
  static_assert(PEND >= 1 && (W & (W - 1)) == 0, "ring size");
  __shared__ __align__(16) zdb sW[W * LDFMAX];
  zdb* F = bb.F[blockIdx.x];
  int* ipiv = bb.ipiv[blockIdx.x];
  const int ldf = 3 * k + 1, kv = 2 * k, lane = threadIdx.x;
  const unsigned FULL = 0xffffffffu;
  auto slot = [&](int c) -> zdb* { return sW + (c & (W - 1)) * ldf; };
  // one commit group per column, in column order (empty groups past the matrix keep the count uniform)
  for (int c = 0; c < W; ++c) {
    if (c < n)
      for (int e = lane; e < ldf; e += 32) band_cp16(slot(c) + e, F + (int64_t)c * ldf + e);
    band_commit();
  }
  int ju = 0, info = 0;
  for (int j = 0; j < n; ++j) {
    asm volatile("cp.async.wait_group %0;" ::"n"(PEND) : "memory");   // columns <= j + 2k have landed
    __syncwarp();
    const int km = min(k, n - 1 - j);
    zdb* colj = slot(j) + kv;                // colj[i] = A[j+i, j]
    double best = -1.0;
    int bi = 0x7fffffff;
    for (int i = lane; i <= km; i += 32) {
      const zdb v = colj[i];
      const double a = fabs(v.x) + fabs(v.y);
      if (a > best) { best = a; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(FULL, best, o);
      const int oi = __shfl_xor_sync(FULL, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    const int jp = (best < 0.0) ? 0 : bi;
    if (lane == 0) ipiv[j] = j + jp;
    ju = max(ju, min(j + k + jp, n - 1));
    if (best > 0.0) {
      if (jp != 0) {
        for (int c = j + lane; c <= ju; c += 32) {   // swap rows j and j+jp over the columns j..ju
          zdb* e0 = slot(c) + (kv + j - c);
          const zdb t = e0[0];
          e0[0] = e0[jp];
          e0[jp] = t;
        }
        __syncwarp();
      }
      const zdb inv = mk<double>(1.0, 0.0) / colj[0];
      for (int i = 1 + lane; i <= km; i += 32) colj[i] = colj[i] * inv;
      __syncwarp();
      const int nc = ju - j;
      for (int e = lane; e < nc * km; e += 32) {
        const int cc = e / km + 1, i = e % km + 1;       // column j+cc, row j+i
        zdb* colc = slot(j + cc) + (kv - cc);             // colc[i] = A[j+i, j+cc]
        colc[i] = colc[i] - colj[i] * colc[0];
      }
      __syncwarp();
    } else if (info == 0) {
      info = j + 1;                                     // exactly singular (zgbtf2: info = j, the column is skipped)
    }
    // column j is final: back to global memory, its slot takes column j + W
    for (int e = lane; e < ldf; e += 32) F[(int64_t)j * ldf + e] = slot(j)[e];
    __syncwarp();
    if (j + W < n)
      for (int e = lane; e < ldf; e += 32) band_cp16(slot(j) + e, F + (int64_t)(j + W) * ldf + e);
    band_commit();
  }
  band_wait0();
  if (lane == 0) ipiv[n] = info;
}

// zgbtf2 (kl = ku = k) with ONE WARP per node: no block barriers, the pivot search is a shuffle butterfly, the (km x (ju - j)) rank-1
// update is spread over the lanes; the window a step touches (2k + 1 columns of 3k + 1 rows: 5 KB at k = 7) lives in L1.
__global__ void __launch_bounds__(32) k_band_lu_warp(int n, int k, BandBatch bb) {
  zdb* F = bb.F[blockIdx.x];
  int* ipiv = bb.ipiv[blockIdx.x];
  const int ldf = 3 * k + 1, kv = 2 * k, lane = threadIdx.x;
  const unsigned FULL = 0xffffffffu;
  int ju = 0, info = 0;
  for (int j = 0; j < n; ++j) {
    const int km = min(k, n - 1 - j);
    zdb* colj = F + (int64_t)j * ldf + kv;   // colj[i] = A[j+i, j]
    double best = -1.0;
    int bi = 0x7fffffff;
    for (int i = lane; i <= km; i += 32) {
      const zdb v = colj[i];
      const double a = fabs(v.x) + fabs(v.y);
      if (a > best) { best = a; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(FULL, best, o);
      const int oi = __shfl_xor_sync(FULL, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    const int jp = (best < 0.0) ? 0 : bi;     // a column of NaNs: no candidate, flagged below
    if (lane == 0) ipiv[j] = j + jp;
    ju = max(ju, min(j + k + jp, n - 1));
    if (!(best > 0.0)) {                        // exactly singular (zgbtf2: info = j, the column is skipped)
      if (info == 0) info = j + 1;
      continue;
    }
    if (jp != 0) {
      for (int c = j + lane; c <= ju; c += 32) {   // swap rows j and j+jp over the columns j..ju
        zdb* e0 = F + (int64_t)c * ldf + (kv + j - c);
        const zdb t = e0[0];
        e0[0] = e0[jp];
        e0[jp] = t;
      }
      __syncwarp();
    }
    const zdb inv = mk<double>(1.0, 0.0) / colj[0];
    for (int i = 1 + lane; i <= km; i += 32) colj[i] = colj[i] * inv;
    __syncwarp();
    const int nc = ju - j;
    for (int e = lane; e < nc * km; e += 32) {
      const int cc = e / km + 1, i = e % km + 1;       // column j+cc, row j+i
      zdb* colc = F + (int64_t)(j + cc) * ldf + (kv - cc);   // colc[i] = A[j+i, j+cc]
      colc[i] = colc[i] - colj[i] * colc[0];
    }
    __syncwarp();
  }
  if (lane == 0) ipiv[n] = info;
}

// w -= a * b
__device__ __forceinline__ void band_fnma(zdb& w, const zdb a, const zdb b) {
  w.x = fma(-a.x, b.x, w.x); w.x = fma(a.y, b.y, w.x);
  w.y = fma(-a.x, b.y, w.y); w.y = fma(-a.y, b.x, w.y);
}

// zgbtrs (no transpose) for all nodes of a sweep: one warp per (node, group of 32 right-hand-side columns), one thread per column.
// The rows a step touches slide through a REGISTER window (K+1 values forward, 2K+1 backward; K >= k is the compile-time size), so x
// moves once per row and sweep: read RHS -> store y, read y -> store x.  Factor columns, pivots and the rows entering the window are
// staged CH steps ahead by cp.async into a double-buffered, warp-private shared-memory ring (no block barriers; a chunk of CH steps
// is longer than the HBM latency).  X must not alias RHS.  X_q = X + q * xbatch.
template <int K>
__global__ void __launch_bounds__(32) k_band_solve_win(int n, int k, BandBatch bb, int m, int64_t ld, const zdb* __restrict__ RHS,
                                                       zdb* __restrict__ X, int64_t xbatch) {
  constexpr int CH = 16, LDFMAX = 3 * K + 1;
  __shared__ __align__(16) zdb sF[2][CH * LDFMAX];
  __shared__ __align__(16) zdb sX[2][CH * 32];
  __shared__ int sP[2][CH];
  const int lane = threadIdx.x, node = blockIdx.y;
  const int c = blockIdx.x * 32 + lane;
  const bool act = c < m;
  const int cr = act ? c : (m - 1);             // idle lanes shadow the last column (loads only)
  const zdb* F = bb.F[node];
  const int* ipiv = bb.ipiv[node];
  const zdb* b = RHS + cr;
  zdb* x = X + (int64_t)node * xbatch + cr;
  const int ldf = 3 * k + 1, kv = 2 * k;
  const int nch = (n + CH - 1) / CH;

  // ---------------- forward: L y = P b ----------------
  auto stage_fwd = [&](int ch, int buf) {
    const int j0 = ch * CH, cnt = min(CH, n - j0);
    const zdb* src = F + (int64_t)j0 * ldf;
    for (int e = lane; e < cnt * ldf; e += 32) band_cp16(&sF[buf][e], src + e);
    if (lane < cnt) band_cp4(&sP[buf][lane], ipiv + j0 + lane);
    for (int s = 0; s < cnt; ++s) {
      const int row = j0 + s + 1 + K;            // enters the window at the end of step j0 + s
      if (row < n) band_cp16(&sX[buf][s * 32 + lane], b + (int64_t)row * ld);
    }
  };
  zdb w[K + 1];
#pragma unroll
  for (int i = 0; i <= K; ++i) w[i] = (i < n) ? b[(int64_t)i * ld] : czero<double>();
  stage_fwd(0, 0);
  band_commit();
  for (int ch = 0; ch < nch; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < nch) stage_fwd(ch + 1, buf ^ 1);
    band_commit();
    band_wait1();
    __syncwarp();
    const int j0 = ch * CH, cnt = min(CH, n - j0);
    for (int s = 0; s < cnt; ++s) {
      const int j = j0 + s;
      const zdb* Lc = &sF[buf][s * ldf + kv];    // Lc[i] = multiplier of row j+i
      const int p = sP[buf][s] - j;
      if (p != 0) {
        const zdb t = w[0];
#pragma unroll
        for (int i = 1; i <= K; ++i)
          if (i == p) { w[0] = w[i]; w[i] = t; }
      }
      const zdb xj = w[0];
      if (act) x[(int64_t)j * ld] = xj;
#pragma unroll
      for (int i = 1; i <= K; ++i)
        if (i <= k) band_fnma(w[i], Lc[i], xj);
#pragma unroll
      for (int i = 1; i <= K; ++i) w[i - 1] = w[i];
      w[K] = (j + 1 + K < n) ? sX[buf][s * 32 + lane] : czero<double>();
    }
    __syncwarp();
  }
  band_wait0();
  __threadfence_block();
  __syncwarp();

  // ---------------- backward: U x = y ----------------
  auto stage_bwd = [&](int ch, int buf) {
    const int jhi = n - 1 - ch * CH, jlo = max(0, jhi - CH + 1), cnt = jhi - jlo + 1;
    const zdb* src = F + (int64_t)jlo * ldf;
    for (int e = lane; e < cnt * ldf; e += 32) band_cp16(&sF[buf][e], src + e);
    for (int s = 0; s < cnt; ++s) {
      const int row = jhi - s - 1 - 2 * K;       // enters the window at the end of step jhi - s
      if (row >= 0) band_cp16(&sX[buf][s * 32 + lane], x + (int64_t)row * ld);
    }
  };
  zdb v[2 * K + 1];
#pragma unroll
  for (int d = 0; d <= 2 * K; ++d) v[d] = (n - 1 - d >= 0) ? x[(int64_t)(n - 1 - d) * ld] : czero<double>();
  stage_bwd(0, 0);
  band_commit();
  for (int ch = 0; ch < nch; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < nch) stage_bwd(ch + 1, buf ^ 1);
    band_commit();
    band_wait1();
    __syncwarp();
    const int jhi = n - 1 - ch * CH, jlo = max(0, jhi - CH + 1);
    for (int s = 0; s <= jhi - jlo; ++s) {
      const int j = jhi - s;
      const zdb* Uc = &sF[buf][(j - jlo) * ldf];  // Uc[kv - d] = U[j-d, j]
      const zdb xj = v[0] / Uc[kv];
      if (act) x[(int64_t)j * ld] = xj;
#pragma unroll
      for (int d = 1; d <= 2 * K; ++d)
        if (d <= kv) band_fnma(v[d], Uc[kv - d], xj);
#pragma unroll
      for (int d = 1; d <= 2 * K; ++d) v[d - 1] = v[d];
      v[2 * K] = (j - 1 - 2 * K >= 0) ? sX[buf][s * 32 + lane] : czero<double>();
    }
    __syncwarp();
  }
  band_wait0();
}

// The same substitutions with the window of a column spread over G LANES (G = 4, 8, 16, 32 >= 2k + 1): lane i of a column's group
// holds row j + i (forward) resp. j - i (backward) of the running right-hand side.  A step is: broadcast x_j from lane 0 (after the
// pivot swap, two shuffles), ONE complex FMA per lane with its own multiplier, shift the window by a shuffle, the lane at the far end
// takes the entering row from the staging ring.  The dependent chain of a step is shuffle -> 2 FMA -> shuffle (~80 cycles) where the
// one-thread-per-column window kernel above walks k (resp. 2k) FMAs, 2k register moves and a division per step with a single warp
// on its SM (measured: 1 800 cycles per row for both sweeps at k = 7).  32 / G columns per warp.
template <typename T>
__device__ __forceinline__ cx<T> band_shfl(cx<T> v, int src, int width) {
  cx<T> r;
  r.x = __shfl_sync(0xffffffffu, v.x, src, width);
  r.y = __shfl_sync(0xffffffffu, v.y, src, width);
  return r;
}
template <typename T>
__device__ __forceinline__ cx<T> band_shfl_down(cx<T> v, int width) {
  cx<T> r;
  r.x = __shfl_down_sync(0xffffffffu, v.x, 1, width);
  r.y = __shfl_down_sync(0xffffffffu, v.y, 1, width);
  return r;
}

template <int G>
__global__ void __launch_bounds__(32) k_band_solve_lanes(int n, int k, BandBatch bb, int m, int64_t ld, const zdb* __restrict__ RHS,
                                                         zdb* __restrict__ X, int64_t xbatch) {
  constexpr int CH = 16, CPW = 32 / G, KMAX = (G - 1) / 2, LDFMAX = 3 * KMAX + 1;
  __shared__ __align__(16) zdb sF[2][CH * LDFMAX];
  __shared__ __align__(16) zdb sX[2][CH * CPW];
  __shared__ int sP[2][CH];
  const int lane = threadIdx.x, node = blockIdx.y;
  const int g = lane / G, i = lane % G;            // column of the warp, window slot
  const int c = blockIdx.x * CPW + g;
  const bool act = c < m;
  const int cr = act ? c : (m - 1);                // idle groups shadow the last column (loads only)
  const zdb* F = bb.F[node];
  const int* ipiv = bb.ipiv[node];
  zdb* Xn = X + (int64_t)node * xbatch;
  const int ldf = 3 * k + 1, kv = 2 * k;
  const int nch = (n + CH - 1) / CH;
  const int col0 = blockIdx.x * CPW;

  // ---------------- forward: L y = P b ----------------
  auto stage_fwd = [&](int ch, int buf) {
    const int j0 = ch * CH, cnt = min(CH, n - j0);
    const zdb* src = F + (int64_t)j0 * ldf;
    for (int e = lane; e < cnt * ldf; e += 32) band_cp16(&sF[buf][e], src + e);
    if (lane < cnt) band_cp4(&sP[buf][lane], ipiv + j0 + lane);
    for (int e = lane; e < cnt * CPW; e += 32) {
      const int s = e / CPW, gg = e % CPW;
      const int row = j0 + s + 1 + k;              // enters the window at the end of step j0 + s
      if (row < n) band_cp16(&sX[buf][e], RHS + (int64_t)row * ld + min(col0 + gg, m - 1));
    }
  };
  zdb w = (i <= k && i < n) ? RHS[(int64_t)i * ld + cr] : czero<double>();
  stage_fwd(0, 0);
  band_commit();
  for (int ch = 0; ch < nch; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < nch) stage_fwd(ch + 1, buf ^ 1);
    band_commit();
    band_wait1();
    __syncwarp();
    const int j0 = ch * CH, cnt = min(CH, n - j0);
    for (int s = 0; s < cnt; ++s) {
      const int j = j0 + s;
      const int p = sP[buf][s] - j;
      const zdb lmul = (i >= 1 && i <= k) ? sF[buf][s * ldf + kv + i] : czero<double>();   // multiplier of row j + i
      const zdb enter = (j + 1 + k < n) ? sX[buf][s * CPW + g] : czero<double>();
      const zdb v0 = band_shfl(w, 0, G);
      zdb xj = v0;
      if (p != 0) {
        const zdb vp = band_shfl(w, p, G);
        if (i == 0) w = vp;
        else if (i == p) w = v0;
        xj = vp;
      }
      if (i == 0 && act) Xn[(int64_t)j * ld + cr] = xj;
      band_fnma(w, lmul, xj);                       // lanes outside 1..k carry a zero multiplier
      w = band_shfl_down(w, G);
      if (i == k) w = enter;
    }
    __syncwarp();
  }
  band_wait0();
  __threadfence_block();
  __syncwarp();

  // ---------------- backward: U x = y ----------------
  auto stage_bwd = [&](int ch, int buf) {
    const int jhi = n - 1 - ch * CH, jlo = max(0, jhi - CH + 1), cnt = jhi - jlo + 1;
    const zdb* src = F + (int64_t)jlo * ldf;
    for (int e = lane; e < cnt * ldf; e += 32) band_cp16(&sF[buf][e], src + e);
    for (int e = lane; e < cnt * CPW; e += 32) {
      const int s = e / CPW, gg = e % CPW;
      const int row = jhi - s - 1 - kv;            // enters the window at the end of step jhi - s
      if (row >= 0) band_cp16(&sX[buf][e], Xn + (int64_t)row * ld + min(col0 + gg, m - 1));
    }
  };
  zdb v = (i <= kv && n - 1 - i >= 0) ? Xn[(int64_t)(n - 1 - i) * ld + cr] : czero<double>();
  stage_bwd(0, 0);
  band_commit();
  for (int ch = 0; ch < nch; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < nch) stage_bwd(ch + 1, buf ^ 1);
    band_commit();
    band_wait1();
    __syncwarp();
    const int jhi = n - 1 - ch * CH, jlo = max(0, jhi - CH + 1);
    for (int s = 0; s <= jhi - jlo; ++s) {
      const int j = jhi - s;
      const zdb* Uc = &sF[buf][(j - jlo) * ldf];   // Uc[kv - d] = U[j-d, j]
      const zdb rinv = mk<double>(1.0, 0.0) / Uc[kv];
      const zdb umul = (i >= 1 && i <= kv) ? Uc[kv - i] : czero<double>();
      const zdb enter = (j - 1 - kv >= 0) ? sX[buf][s * CPW + g] : czero<double>();
      const zdb xj = band_shfl(v, 0, G) * rinv;
      if (i == 0 && act) Xn[(int64_t)j * ld + cr] = xj;
      band_fnma(v, umul, xj);
      v = band_shfl_down(v, G);
      if (i == kv) v = enter;
    }
    __syncwarp();
  }
  band_wait0();
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
