// Matrix-free real symmetric operators (feast_matvec / feast_sparse_matvec!, interfaces/feast_interfaces.jl:465-481,
// sparse/feast_sparse.jl:1284-1471): the caller's DEVICE callback produces W = A U for a row-major block, these kernels are the
// rest of a multi-shift Lanczos step -- the arithmetic that k_lz_spmm fuses behind its gather when the operator is a stored CSR
// matrix (same helper functions lz_t / lz_next, hence the same rounding in pass 1 and pass 2).
// HBM-bound elementwise kernels.  Bytes per launch (n rows, m columns): P1 4*8nm, P2 6*8nm, RES 4*8nm.
#pragma once
#include "kernels_lanczos.cuh"

namespace feastcuda {

// MODE LZ_P1 : out = W*inv_beta - ratio_b*prev ; partial = u . out
//      LZ_P2 : t as P1 ; out = t - ratio_a*u ; Q += coef*u
//      LZ_RES: out = W - theta*u ; partial = |out|^2 ; Q = coef*u
template <int MODE>
__global__ void __launch_bounds__(256) k_mf_step(int64_t n, int m, int pp, int64_t ld, const double* __restrict__ W,
                                                 const double* __restrict__ U, const double* prev, double* out, double* __restrict__ Q,
                                                 const double* __restrict__ s0, const double* __restrict__ s_rb,
                                                 const double* __restrict__ s_ra, const double* __restrict__ s_cf,
                                                 double* __restrict__ partial, int pstride, const int* __restrict__ done) {
  if (done != nullptr && *done != 0) return;
  EwMap2 e(pp);
  const int P = (m + 1) >> 1;
  double2 acc = make_double2(0.0, 0.0);
  if (e.pc < P) {
    const double2 a0 = lz_scal<false>(s0, e.pc, m), rb = lz_scal<false>(s_rb, e.pc, m), ra = lz_scal<false>(s_ra, e.pc, m),
                  cf = lz_scal<false>(s_cf, e.pc, m);
    for (int64_t row = (int64_t)blockIdx.x * e.rpb + e.rsub; row < n; row += (int64_t)gridDim.x * e.rpb) {
      const int64_t off = row * ld + 2 * e.pc;
      const double2 w = ldg2(W + off), u = ldg2(U + off);
      if constexpr (MODE == LZ_RES) {
        double2 t;
        t.x = __fma_rn(-a0.x, u.x, w.x);
        t.y = __fma_rn(-a0.y, u.y, w.y);
        stg2(out + off, t);
        acc.x = fma(t.x, t.x, acc.x);
        acc.y = fma(t.y, t.y, acc.y);
        if (Q != nullptr) stg2(Q + off, make_double2(cf.x * u.x, cf.y * u.y));
      } else {
        const double2 t = lz_t(w, a0, rb, ldg2(prev + off));
        if constexpr (MODE == LZ_P1) {
          stg2(out + off, t);
          acc.x = fma(u.x, t.x, acc.x);
          acc.y = fma(u.y, t.y, acc.y);
        } else {
          stg2(out + off, lz_next(t, ra, u));
          double2 q = ldg2(Q + off);
          q.x = fma(cf.x, u.x, q.x);
          q.y = fma(cf.y, u.y, q.y);
          stg2(Q + off, q);
        }
      }
    }
  }
  if constexpr (MODE != LZ_P2) block_reduce_pairs<false>(acc, pp, P, m, partial + (int64_t)blockIdx.x * pstride);
}

}  // namespace feastcuda
