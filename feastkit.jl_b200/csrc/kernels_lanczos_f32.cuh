// Mixed-precision variant of the multi-shift Lanczos filter (kernels_lanczos.cuh): FP32 Krylov vectors and matrix
// entries, FP64 everything else.  FEAST's own parameter for it is fpm[42] ("Mixed precision (0=double, 1=single solver)",
// default 1, core/feast_parameters.jl:316-319 -- declared by the reference, consumed by none of its drivers).
//
// Why it is safe: from the second refinement loop on the recurrence starts from the Ritz RESIDUAL block A q - theta q
// (computed in FP64 by k_lz_spmm<LZ_RES>), so the FP32 solve only produces a correction of relative accuracy
// ~max(inner_rel, eps32 * cond) to a quantity that is already O(epsout); the accumulator Q, the Lanczos scalars
// (alpha, beta, the shifted-residual recurrences, the quadrature coefficients), the Rayleigh-Ritz stage and the residuals
// stay FP64.  It is iterative refinement with a low-precision inner solver; run_interval falls back to FP64 vectors when a
// sweep stops contracting.
//
// Layout: row-major n x ld FLOAT blocks, ld a multiple of 4; a 16-byte element is FOUR columns and a lane owns one element
// (m <= 128 columns -> at most 32 lanes per row, so no per-lane chunk loop).  The accumulator is an n x ld DOUBLE block.
// Both passes run the same instruction sequence (explicit __fmul_rn/__fmaf_rn, CSR-order gather), so pass 2 reproduces
// pass 1's vectors bit for bit, exactly like the FP64 kernels.
//
// HBM-bound.  Algorithmic bytes per launch:   LZ_P1  nnz*8 + 4(n+1) + 3*n*m*4      update  3*n*m*4
//                                             LZ_P2  nnz*8 + 4(n+1) + 3*n*m*4 + 2*n*m*8
#pragma once
#include "kernels_lanczos.cuh"

namespace feastcuda {

struct LzArgs32 {
  int64_t n;
  int m;                 // active columns
  int64_t ld;            // row stride in elements of the vector type (floats for U/prev/out, doubles for Q); multiple of 4
  const int* ptr; const int* col; const float* val;
  const float* U;        // gathered operand
  const float* prev;     // own-row operand u_{j-1} (the host passes U at j = 0, ratio_b is 0 there)
  float* out;            // own-row result (may alias prev)
  double* Q;             // FP64 accumulator (LZ_P2)
  const double* s_inv_beta; const double* s_ratio_b; const double* s_ratio_a;   // per-column scalars of this step (rounded to FP32 on load)
  const double* s_coef;  // LZ_P2: c_j / beta_j (kept FP64)
  double* partial;       // [gridDim.x][pstride]
  int pstride;
  int tile_rows;
  const int* done;
  const double* s_coef_prev;   // LZ_P2_PAIR: c_{j-1} / beta_{j-1}
  LzTail tail;           // pass 1: the step's scalar recurrences, run by the last CTA (kernels_lanczos.cuh)
  // row-sharded runs (see LzArgs): pre-resolved gather distances in 16-byte units (the FLOAT blocks' row stride), tile order, lazy wait
  const int* goff;
  const int* tile_order;
  int halo_start;
  unsigned long long wait_seq;
  const unsigned long long* kdone;
  int nranks, rank;
};

// gather metadata of an entry: element offset col * ld (single GPU) or the signed distance to the row in its owner's block (row-sharded)
template <bool SHARD> struct Lz32Off;
template <> struct Lz32Off<false> {
  typedef unsigned T;
  static __device__ __forceinline__ T meta(const LzArgs32& a, int p, unsigned ldu) { return (unsigned)a.col[p] * ldu; }
  static __device__ __forceinline__ T own(unsigned eo_own) { return eo_own; }
  static __device__ __forceinline__ const float* at(const float* base, T o) { return base + o; }
};
template <> struct Lz32Off<true> {
  typedef int T;
  static __device__ __forceinline__ T meta(const LzArgs32& a, int p, unsigned) { return a.goff[p]; }
  static __device__ __forceinline__ T own(unsigned eo_own) { return (int)(eo_own >> 2); }
  static __device__ __forceinline__ const float* at(const float* base, T o) { return base + 4 * (long long)o; }
};

__device__ __forceinline__ float4 ldg4f(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void stg4f(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// four per-column scalars of element e, zero beyond the active columns
__device__ __forceinline__ float4 lz32_scal(const double* s, int e, int m) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s != nullptr) {
    const int c = 4 * e;
    if (c < m) v.x = (float)s[c];
    if (c + 1 < m) v.y = (float)s[c + 1];
    if (c + 2 < m) v.z = (float)s[c + 2];
    if (c + 3 < m) v.w = (float)s[c + 3];
  }
  return v;
}

__device__ __forceinline__ float4 lz32_t(float4 acc, float4 ib, float4 rb, float4 pv) {
  float4 t;
  t.x = __fmaf_rn(-rb.x, pv.x, __fmul_rn(acc.x, ib.x));
  t.y = __fmaf_rn(-rb.y, pv.y, __fmul_rn(acc.y, ib.y));
  t.z = __fmaf_rn(-rb.z, pv.z, __fmul_rn(acc.z, ib.z));
  t.w = __fmaf_rn(-rb.w, pv.w, __fmul_rn(acc.w, ib.w));
  return t;
}
__device__ __forceinline__ float4 lz32_next(float4 t, float4 ra, float4 uo) {
  float4 r;
  r.x = __fmaf_rn(-ra.x, uo.x, t.x);
  r.y = __fmaf_rn(-ra.y, uo.y, t.y);
  r.z = __fmaf_rn(-ra.z, uo.z, t.z);
  r.w = __fmaf_rn(-ra.w, uo.w, t.w);
  return r;
}
__device__ __forceinline__ void lz32_dot(double (&d)[4], float4 x, float4 y) {
  d[0] = fma((double)x.x, (double)y.x, d[0]);
  d[1] = fma((double)x.y, (double)y.y, d[1]);
  d[2] = fma((double)x.z, (double)y.z, d[2]);
  d[3] = fma((double)x.w, (double)y.w, d[3]);
}

// acc += sum_p val[p] * U[col[p], element]  over the stored entries [p0, p1) of the row, CSR order (see lz_gather)
template <int G, int UNMAX, bool SHARD>
__device__ __forceinline__ void lz32_gather(const LzArgs32& a, unsigned row_eo_, int p0, int p1, typename Lz32Off<SHARD>::T myo, float mya,
                                            typename Lz32Off<SHARD>::T myo2, float mya2, int g, unsigned gmask, const float* Ul, float4& acc) {
  typedef Lz32Off<SHARD> OF;
  constexpr int UN = (G >= UNMAX) ? UNMAX : G;   // gathers in flight per lane
  constexpr bool PF2 = (G <= 4);
  const unsigned ldu = (unsigned)a.ld;
  const typename OF::T row_eo = OF::own(row_eo_);
  for (int pb = p0; pb < p1; pb += G) {
    const int cnt = min(G, p1 - pb);
    if (PF2 && pb == p0 + G) {
      myo = myo2;
      mya = mya2;
    } else if (pb != p0) {
      myo = row_eo;
      mya = 0.f;
      if (g < cnt) { myo = OF::meta(a, pb + g, ldu); mya = a.val[pb + g]; }
    }
    for (int t = 0; t < cnt; t += UN) {
      typename OF::T eo[UN];
      float aa[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        eo[u] = __shfl_sync(gmask, myo, t + u, G);     // slots beyond cnt carry weight 0 and a valid offset
        aa[u] = __shfl_sync(gmask, mya, t + u, G);
      }
      float4 xv[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) xv[u] = ldg4f(OF::at(Ul, eo[u]));
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        acc.x = __fmaf_rn(aa[u], xv[u].x, acc.x);
        acc.y = __fmaf_rn(aa[u], xv[u].y, acc.y);
        acc.z = __fmaf_rn(aa[u], xv[u].z, acc.z);
        acc.w = __fmaf_rn(aa[u], xv[u].w, acc.w);
      }
    }
  }
}

// MODE: LZ_P1 (out = (A u) inv_beta - ratio_b prev; partial = u . out) or LZ_P2 (out = t - ratio_a u; Q += coef u)
// (THREADS, MINB, UNMAX): 512 x 2 CTAs/SM x 4 gathers in flight (64 registers) or fewer resident warps with more loads in flight each
template <int G, int MODE, int THREADS, int MINB = 1024 / THREADS, int UNMAX = 4, bool SHARD = false>
__global__ void __launch_bounds__(THREADS, MINB) k_lz32_spmm(LzArgs32 a) {
  static_assert(MODE == LZ_P1 || lz_is_p2(MODE), "FP32 vectors exist only inside the two Lanczos passes");
  typedef Lz32Off<SHARD> OF;
  typedef typename OF::T off_t;
  if (a.done != nullptr && *a.done != 0) return;
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x & 31, g = lane % G, sub = lane / G;
  constexpr unsigned gm0 = (G >= 32) ? 0xffffffffu : ((1u << (G & 31)) - 1u);
  const unsigned gmask = gm0 << (sub * G);
  const int wib = threadIdx.x >> 5, wpb = THREADS >> 5;
  constexpr int STEP = (THREADS / 32) * RPW;
  const int P = (a.m + 3) >> 2;                       // 16-byte elements (4 columns) per row
  const int n = (int)a.n;
  // round-robin tiles of consecutive rows, as in k_lz_spmm: one moving front over the matrix
  const int spt = max(1, a.tile_rows / STEP);
  const int TR = spt * STEP;
  const int ntiles = (n + TR - 1) / TR;
  const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int niter = my_tiles * spt;
  const int lane_row = wib * RPW + sub;
  int it_f = 0, s_f = 0, base_f = (int)blockIdx.x * TR, t_f = (int)blockIdx.x;
  if (SHARD && a.tile_order != nullptr && my_tiles > 0) base_f = a.tile_order[t_f] * TR;
  auto next_row = [&]() -> int {
    const int r = (it_f < niter) ? base_f + lane_row : n;
    ++it_f;
    if (++s_f == spt) {
      s_f = 0;
      if (SHARD && a.tile_order != nullptr) {
        t_f += (int)gridDim.x;
        base_f = (it_f < niter) ? a.tile_order[t_f] * TR : 0;
      } else base_f += ((int)gridDim.x - 1) * TR + STEP;
    } else base_f += STEP;
    return r < n ? r : n;
  };

  __shared__ float4 s_sc[3][FC_MAXCOLS / 4];     // [0] inv_beta, [1] ratio_b, [2] ratio_a
  __shared__ double s_cf[FC_MAXCOLS];            // coef (pass 2)
  __shared__ double s_cp[FC_MAXCOLS];            // coef of the previous step (q_mode 2)
  for (int i = threadIdx.x; i < 3 * (FC_MAXCOLS / 4); i += THREADS) {
    const int w = i / (FC_MAXCOLS / 4), pc = i % (FC_MAXCOLS / 4);
    s_sc[w][pc] = lz32_scal(w == 0 ? a.s_inv_beta : (w == 1 ? a.s_ratio_b : a.s_ratio_a), pc, a.m);
  }
  for (int i = threadIdx.x; i < FC_MAXCOLS; i += THREADS) {
    s_cf[i] = (lz_is_p2(MODE) && a.s_coef != nullptr && i < a.m) ? a.s_coef[i] : 0.0;
    s_cp[i] = (MODE == LZ_P2_PAIR && a.s_coef_prev != nullptr && i < a.m) ? a.s_coef_prev[i] : 0.0;
  }
  __syncthreads();
  double dot[4] = {0.0, 0.0, 0.0, 0.0};

  const unsigned ldu = (unsigned)a.ld;
  const int pcl = 4 * min(g, P - 1);             // lanes beyond the active elements read a clamped (valid) element
  const float* Ul = a.U + pcl;
  const float* Pl = a.prev + pcl;
  float* Ol = a.out + pcl;
  double* Ql = a.Q + pcl;
  int r_cur = next_row(), r_nxt = next_row();
  int p0_cur = 0, p1_cur = 0, p0_nxt = 0, p1_nxt = 0;
  if (r_cur < n) { p0_cur = a.ptr[r_cur]; p1_cur = a.ptr[r_cur + 1]; }
  if (r_nxt < n) { p0_nxt = a.ptr[r_nxt]; p1_nxt = a.ptr[r_nxt + 1]; }
  constexpr bool PF2 = (G <= 4);
  off_t o_cur = OF::own((r_cur < n ? (unsigned)r_cur : 0u) * ldu), o_cur2 = o_cur;
  float a_cur = 0.f, a_cur2 = 0.f;
  if (g < p1_cur - p0_cur) { o_cur = OF::meta(a, p0_cur + g, ldu); a_cur = a.val[p0_cur + g]; }
  if (PF2 && g + G < p1_cur - p0_cur) { o_cur2 = OF::meta(a, p0_cur + G + g, ldu); a_cur2 = a.val[p0_cur + G + g]; }

  int it_h = 0x7fffffff;     // first iteration of this CTA that may touch a halo tile (see k_lz_spmm)
  if (SHARD && a.wait_seq != 0 && a.tile_order != nullptr) {
    const int first = a.halo_start > (int)blockIdx.x ? (a.halo_start - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    it_h = first * spt;
  }
  for (int it = 0; it < niter; ++it) {
    if (SHARD && it == it_h) {
      if (lane < a.nranks && lane != a.rank) { while (ld_sys_u64(a.kdone + lane) < a.wait_seq) { } }
      __syncwarp();
    }
    const int row = r_cur;
    const bool valid = row < n;
    const int r_fut = next_row();
    int p0_fut = 0, p1_fut = 0;
    if (r_fut < n) { p0_fut = a.ptr[r_fut]; p1_fut = a.ptr[r_fut + 1]; }
    off_t o_nxt = OF::own((r_nxt < n ? (unsigned)r_nxt : 0u) * ldu), o_nxt2 = o_nxt;
    float a_nxt = 0.f, a_nxt2 = 0.f;
    if (g < p1_nxt - p0_nxt) { o_nxt = OF::meta(a, p0_nxt + g, ldu); a_nxt = a.val[p0_nxt + g]; }
    if (PF2 && g + G < p1_nxt - p0_nxt) { o_nxt2 = OF::meta(a, p0_nxt + G + g, ldu); a_nxt2 = a.val[p0_nxt + G + g]; }

    const unsigned eo_own = (valid ? (unsigned)row : 0u) * ldu;
    const float4 uo = ldg4f(Ul + eo_own);
    const float4 pv = ldg4f(Pl + eo_own);
    double2 q0 = make_double2(0.0, 0.0), q1 = q0;
    if constexpr (MODE == LZ_P2 || MODE == LZ_P2_PAIR) {
      q0 = ldg2(Ql + eo_own);
      q1 = ldg2(Ql + eo_own + 2);
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    lz32_gather<G, UNMAX, SHARD>(a, eo_own, p0_cur, p1_cur, o_cur, a_cur, o_cur2, a_cur2, g, gmask, Ul, acc);
    if (valid && g < P) {
      const float4 t = lz32_t(acc, s_sc[0][g], s_sc[1][g], pv);
      if constexpr (MODE == LZ_P1) {
        stg4f(Ol + eo_own, t);
        lz32_dot(dot, uo, t);
      } else {
        stg4f(Ol + eo_own, lz32_next(t, s_sc[2][g], uo));
        if constexpr (MODE != LZ_P2_SKIP) {
          if constexpr (MODE == LZ_P2_PAIR) {
            const double* cp = s_cp + 4 * g;
            q0.x = fma(cp[0], (double)pv.x, q0.x);
            q0.y = fma(cp[1], (double)pv.y, q0.y);
            q1.x = fma(cp[2], (double)pv.z, q1.x);
            q1.y = fma(cp[3], (double)pv.w, q1.y);
          }
          const double* cf = s_cf + 4 * g;
          q0.x = fma(cf[0], (double)uo.x, q0.x);
          q0.y = fma(cf[1], (double)uo.y, q0.y);
          q1.x = fma(cf[2], (double)uo.z, q1.x);
          q1.y = fma(cf[3], (double)uo.w, q1.y);
          stg2(Ql + eo_own, q0);
          stg2(Ql + eo_own + 2, q1);
        }
      }
    }
    r_cur = r_nxt; p0_cur = p0_nxt; p1_cur = p1_nxt; o_cur = o_nxt; a_cur = a_nxt; o_cur2 = o_nxt2; a_cur2 = a_nxt2;
    r_nxt = r_fut; p0_nxt = p0_fut; p1_nxt = p1_fut;
  }

  __shared__ double red[(THREADS / 32) * 32 * 4];
  if constexpr (MODE == LZ_P1) {
    // fixed-order CTA reduction (pass 2 relies on pass 1's exact scalars)
    const int width = 4 * G;   // columns per row group
#pragma unroll
    for (int q = 0; q < 4; ++q) red[(wib * RPW + sub) * width + 4 * g + q] = dot[q];
    __syncthreads();
    const int ngroups = wpb * RPW;
    for (int c = threadIdx.x; c < width; c += THREADS) {
      if (c < a.m) {
        double s = 0.0;
        for (int q = 0; q < ngroups; ++q) s += red[q * width + c];
        a.partial[(int64_t)blockIdx.x * a.pstride + c] = s;
      }
    }
  }
  __syncthreads();
  lz_tail(a.tail, a.partial, a.pstride, a.m, red);
}

// ---- elementwise kernels: a thread owns one 4-column element, rows strided (pp = power of two >= elements per row) -------
__device__ __forceinline__ void block_reduce_quads(const double (&v)[4], int pp, int m, double* out_row) {
  __shared__ double red4[256 * 4];
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) red4[threadIdx.x * 4 + q] = v[q];
  __syncthreads();
  for (int c = threadIdx.x; c < 4 * pp; c += blockDim.x) {
    if (c < m) {
      const int e = c >> 2, q = c & 3;
      double s = 0.0;
      for (int t = e; t < (int)blockDim.x; t += pp) s += red4[t * 4 + q];
      out_row[c] = s;
    }
  }
}

// pass 1, second half of a step: T (in place) <- T - ratio_a * U ; partial = |T|^2
__global__ void __launch_bounds__(256) k_lz32_update(int64_t n, int m, int pp, int64_t ld, const double* __restrict__ s_ratio_a,
                                                     const float* __restrict__ U, float* __restrict__ T, double* __restrict__ partial,
                                                     int pstride, const int* __restrict__ done, LzTail tail) {
  if (done != nullptr && *done != 0) return;
  EwMap2 e(pp);
  const int P = (m + 3) >> 2;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  if (e.pc < P) {
    const float4 ra = lz32_scal(s_ratio_a, e.pc, m);
    const int64_t stride = (int64_t)gridDim.x * e.rpb;
    int64_t row = (int64_t)blockIdx.x * e.rpb + e.rsub;
    for (; row + 3 * stride < n; row += 4 * stride) {
      float4 tv[4], uv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t off = (row + q * stride) * ld + 4 * e.pc;
        tv[q] = ldg4f(T + off);
        uv[q] = ldg4f(U + off);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t off = (row + q * stride) * ld + 4 * e.pc;
        const float4 r = lz32_next(tv[q], ra, uv[q]);
        stg4f(T + off, r);
        lz32_dot(acc, r, r);
      }
    }
    for (; row < n; row += stride) {
      const int64_t off = row * ld + 4 * e.pc;
      const float4 r = lz32_next(ldg4f(T + off), ra, ldg4f(U + off));
      stg4f(T + off, r);
      lz32_dot(acc, r, r);
    }
  }
  block_reduce_quads(acc, pp, m, partial + (int64_t)blockIdx.x * pstride);
  __shared__ double tail_scratch[FC_MAXCOLS + 256];
  lz_tail(tail, partial, pstride, m, tail_scratch);
}

// Q (double) += coef * U (float)   (last pass-2 step)
__global__ void __launch_bounds__(256) k_lz32_axpy(int64_t n, int m, int pp, int64_t ld, const double* __restrict__ s_coef,
                                                   const float* __restrict__ U, double* __restrict__ Q) {
  EwMap2 e(pp);
  const int P = (m + 3) >> 2;
  if (e.pc >= P) return;
  double cf[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) cf[q] = (4 * e.pc + q < m) ? s_coef[4 * e.pc + q] : 0.0;
  for (int64_t row = (int64_t)blockIdx.x * e.rpb + e.rsub; row < n; row += (int64_t)gridDim.x * e.rpb) {
    const int64_t off = row * ld + 4 * e.pc;
    const float4 u = ldg4f(U + off);
    double2 q0 = ldg2(Q + off), q1 = ldg2(Q + off + 2);
    q0.x = fma(cf[0], (double)u.x, q0.x);
    q0.y = fma(cf[1], (double)u.y, q0.y);
    q1.x = fma(cf[2], (double)u.z, q1.x);
    q1.y = fma(cf[3], (double)u.w, q1.y);
    stg2(Q + off, q0);
    stg2(Q + off + 2, q1);
  }
}

// FP64 compact block (row stride ld doubles) -> FP32 block (row stride ld floats); columns >= m are written as zeros
__global__ void __launch_bounds__(256) k_lz32_narrow(int64_t n, int m, int pp, int64_t ld, const double* __restrict__ X, float* __restrict__ Y) {
  EwMap2 e(pp);
  const int P = (m + 3) >> 2;
  if (e.pc >= P) return;
  const int c = 4 * e.pc;
  for (int64_t row = (int64_t)blockIdx.x * e.rpb + e.rsub; row < n; row += (int64_t)gridDim.x * e.rpb) {
    const int64_t off = row * ld + c;
    float4 v;
    v.x = (float)X[off];
    v.y = (c + 1 < m) ? (float)X[off + 1] : 0.f;
    v.z = (c + 2 < m) ? (float)X[off + 2] : 0.f;
    v.w = (c + 3 < m) ? (float)X[off + 3] : 0.f;
    stg4f(Y + off, v);
  }
}

// matrix entries to FP32, once per operator
__global__ void __launch_bounds__(256) k_lz32_vals(int64_t nnz, const double* __restrict__ v, float* __restrict__ o) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) o[i] = (float)v[i];
}

}  // namespace feastcuda
