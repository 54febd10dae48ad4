// Multi-shift Lanczos inner solver for the standard real-symmetric FEAST problem (B = I, real basis).
//
// All ne shifted systems (z_e I - A) x_e = b of one refinement loop share the Krylov space K(A, b), so the
// engine runs ONE real Lanczos recurrence per right-hand-side column (all columns in lock step) instead of
// ne complex Krylov solves (the reference does ne*M0 GMRES calls, sparse/feast_sparse.jl:164-236,318-370),
// and gets the w_e-weighted sum  Q = sum_e Re(2 w_e x_e)  (sparse/feast_sparse.jl:369) as  V_k c  with
// c = ||b|| sum_e Re(2 w_e (z_e I - T_k)^-1 e_1)  from the k x k Lanczos tridiagonal T_k of each column.
// Two passes over the recurrence (no basis is stored):
//   pass 1  k_lz_spmm<LZ_P1> + k_lz_update: builds T_k; the host tracks every shifted residual
//           beta_{k+1} |e_k^T (z_e I - T_k)^-1 e_1| and stops the sweep;
//   pass 2  k_lz_spmm<LZ_P2>: re-runs the recurrence with the stored scalars (bit-identical vectors) and
//           accumulates Q += c_j v_j in the same kernel -- the accumulation is the solve's epilogue.
// Vectors are REAL row-major n x ld blocks (ld even; the m active columns of a row are contiguous), each lane
// owns column PAIRS (one 16-byte load per stored entry).  Lanczos vectors are kept unnormalised:
// u_j = beta_j v_j, the 1/beta_j factors ride along as per-column scalars.
//
// HBM-bound.  Algorithmic bytes per launch (n rows, m columns, nnz entries, 8-byte values, 4-byte indices):
//   LZ_P1   nnz*12 + 4(n+1) + 3*n*m*8   (gather u_j once, own-row u_{j-1}, write A v_j - beta_j v_{j-1})
//   update  3*n*m*8
//   LZ_P2   nnz*12 + 4(n+1) + 5*n*m*8   (+ read/write of the accumulator)
#pragma once
#include "cxmath.cuh"
#include "kernels_block.cuh"

namespace feastcuda {

enum LzMode {
  LZ_P1 = 0,     // out = (A u) * inv_beta - ratio_b * prev ;   partial[0] = u . out
  LZ_P2 = 1,     // t as LZ_P1; out = t - ratio_a * u ;  Q += coef * u
  LZ_RES = 2,    // out = A u - theta * u ;  partial[0] = |out|^2 ;  Q = coef * u      (Ritz-residual start block)
  LZ_PLAIN = 3,  // out = A u
  // pass 2 with the accumulator touched every SECOND step: step j holds u_j (U) and u_{j-1} (prev) as own-row operands, so one
  // read-modify-write of Q serves two steps (compile-time variants: a run-time switch cost the narrow kernels their schedule)
  LZ_P2_SKIP = 4,   // as LZ_P2 without any Q traffic
  LZ_P2_PAIR = 5,   // as LZ_P2 with Q += coef_prev * u_{j-1} + coef * u_j
  // one step of the three-term Chebyshev iteration for M x = rhs (M = the CSR operand; generalized problems: the inner solves with B):
  //   out = u + c1 (u - prev) + c2 dinv[row] (rhs - M u)          (u gathered, prev / rhs own-row, out may alias prev)
  // c1 = c2-independent of the column, so a solve is a FIXED polynomial in M: no dot products, no data-dependent control flow
  LZ_CHEB = 6,
  LZ_CHEB_DOT = 7   // as LZ_CHEB;  partial[0] = Re(conj(out) . rhs)   (last step: beta^2 = u_{j+1}^H s_{j+1})
};
__host__ __device__ constexpr bool lz_is_p2(int mode) { return mode == LZ_P2 || mode == LZ_P2_SKIP || mode == LZ_P2_PAIR; }

// ---- per-step scalar recurrences -------------------------------------------------------------------------------
// T arrays: [step][FC_MAXCOLS]; `scale` is the running max(|alpha|, beta) used by the breakdown test.
// Shift recurrences: for every (node e, column c) the last entry g of (z_e I - T_j)^-1 e_1 is advanced by the LU
// pivots d of the shifted tridiagonal, so that the residual of every shifted system after j+1 steps,
// beta_{j+1} |g| (relative to ||b||), is known on the device; when all are below `target` the flag done_k = j+1 is
// raised and every later launch of the sweep returns immediately.
struct LzScalars {
  double* alpha; double* beta; double* inv_beta; double* ratio_b; double* ratio_a;   // [kmax + 1][FC_MAXCOLS]
  double* scale;                                                                     // [FC_MAXCOLS]
  cx<double>* d; cx<double>* g;                                                      // [ne][FC_MAXCOLS]
  const cx<double>* z;                                                               // [ne]
  int ne;
  double target;
  int* done_k;
  double* maxres;                                                                    // [kmax + 1]
};

// Row-sharded multi-GPU runs (one process per GPU, every rank owns a block of rows of A and of every vector): the per-step dot
// products are summed over the ranks by a ONE-SHOT exchange through peer memory -- every rank stores its m partial sums into a
// mailbox in each peer's HBM (NVLink stores), raises a sequence flag there, waits for the flags of all peers in its own mailbox and
// adds the nranks rows in rank order (so every rank gets the same bits and takes the same decisions).  The exchange runs in the
// tail of the kernel that produced the partial sums: no NCCL call, no extra launch.  Two consecutive exchanges use different
// slots; a rank can be at most one exchange ahead of its peers (it needs their contribution to finish the current one).
constexpr int LZ_MAXRANKS = 16;
constexpr int LZ_MBOX_SLOTS = 4;
struct LzMailbox {   // lives in every rank's shared arena (zeroed at allocation)
  double data[LZ_MBOX_SLOTS][LZ_MAXRANKS][FC_MAXCOLS];
  unsigned long long flag[LZ_MBOX_SLOTS][LZ_MAXRANKS];
  unsigned long long kdone[LZ_MAXRANKS];   // kdone[p]: sequence number of the last kernel rank p has completed (LZ_TAIL_SIGNAL)
};
struct LzXchg {
  int nranks, rank;                    // nranks <= 1: single GPU, nothing below is used
  unsigned long long seq;              // sequence number of this exchange (host counter, identical on every rank)
  LzMailbox* mbox[LZ_MAXRANKS];        // every rank's mailbox as mapped in THIS process (mbox[rank] is local memory)
};

__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// so[0..m) (shared memory, local sums) -> sums over all ranks.  Called by every thread of ONE CTA per rank.
// m = 0: pure barrier (pass 2: "every rank has finished this step").
__device__ __forceinline__ void lz_exchange(const LzXchg& x, double* so, int m) {
  if (x.nranks <= 1) return;
  const int slot = (int)(x.seq & (LZ_MBOX_SLOTS - 1));
  __syncthreads();
  for (int i = threadIdx.x; i < x.nranks * m; i += blockDim.x) {
    const int p = i / m, c = i % m;
    x.mbox[p]->data[slot][x.rank][c] = so[c];
  }
  __syncthreads();
  if ((int)threadIdx.x < x.nranks) {
    __threadfence_system();     // cumulative: everything this CTA has observed (the other CTAs' rows, the stores above) precedes the flag
    st_sys_u64(&x.mbox[threadIdx.x]->flag[slot][x.rank], x.seq);
    const unsigned long long* mine = &x.mbox[x.rank]->flag[slot][threadIdx.x];
    while (ld_sys_u64(mine) != x.seq) { }
  }
  __syncthreads();
  if ((int)threadIdx.x < m) {
    double s = 0.0;
    for (int p = 0; p < x.nranks; ++p) s += ld_sys_f64(&x.mbox[x.rank]->data[slot][p][threadIdx.x]);
    so[threadIdx.x] = s;
  }
  __syncthreads();
}

// "this rank has completed kernel number seq" to every peer, without waiting (pass 2 of a row-sharded run: the consumers wait lazily, just
// before their halo tiles)
__device__ __forceinline__ void lz_signal(const LzXchg& x) {
  if (x.nranks <= 1) return;
  __syncthreads();
  if ((int)threadIdx.x < x.nranks && (int)threadIdx.x != x.rank) {
    __threadfence_system();
    st_sys_u64(&x.mbox[threadIdx.x]->kdone[x.rank], x.seq);
  }
}

// fixed-order sum of the per-CTA partial rows (L2 reads: the rows were written by other CTAs of the same launch).  The loads of a
// thread are issued in batches of eight before any of them is consumed: the sum costs a few L2 latencies, not one per row.
__device__ __forceinline__ void lz_reduce_rows(const double* partial, int nblocks, int pstride, int m, double* so /*[FC_MAXCOLS]*/,
                                               double* tmp /*[blockDim.x]*/) {
  int mp = 32;
  while (mp < m) mp <<= 1;
  const int c = threadIdx.x % mp, part = threadIdx.x / mp, nparts = max(1, (int)blockDim.x / mp);
  double acc = 0.0;
  if (c < m && part < nparts) {
    const double* base = partial + c;
    for (int b0 = part; b0 < nblocks; b0 += 8 * nparts) {
      double v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int b = b0 + q * nparts;
        v[q] = (b < nblocks) ? __ldcg(base + (int64_t)b * pstride) : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) acc += v[q];
    }
  }
  __syncthreads();
  tmp[threadIdx.x] = acc;
  __syncthreads();
  if (part == 0 && c < m) {
    double t = tmp[c];
    for (int q = 1; q < nparts; ++q) t += tmp[q * mp + c];
    so[c] = t;
  }
  __syncthreads();
}

// beta_0 = ||b||  from the (global) sums of |b|^2
__device__ __forceinline__ void lz_scalars_init(const LzScalars& s, const double* so, int m) {
  const int c = threadIdx.x;
  if (c < m) {
    const double b = sqrt(so[c]);
    const bool ok = b > 1e-290;
    s.beta[c] = b;
    s.inv_beta[c] = ok ? 1.0 / b : 0.0;
    s.ratio_b[c] = 0.0;
    s.scale[c] = 0.0;
  }
  if (threadIdx.x == 0) *s.done_k = 0;
}

// after LZ_P1 of step j: alpha_j = (u_j . t) / beta_j ; ratio_a = alpha_j / beta_j
__device__ __forceinline__ void lz_scalars_alpha(const LzScalars& s, int j, const double* so, int m) {
  const int c = threadIdx.x;
  if (c < m) {
    const int64_t o = (int64_t)j * FC_MAXCOLS + c;
    const double ibv = s.inv_beta[o];
    const double al = so[c] * ibv;
    s.alpha[o] = al;
    s.ratio_a[o] = al * ibv;
    s.scale[c] = fmax(s.scale[c], fabs(al));
  }
}

// after the update of step j: beta_{j+1} = ||u_{j+1}||; a column whose Krylov space is exhausted is frozen (inv = 0);
// then the shifted-residual recurrences and the convergence flag.  tmp: [blockDim.x] doubles of shared memory
__device__ __forceinline__ void lz_scalars_beta(const LzScalars& s, int j, const double* so, int m, double* tmp) {
  const int c = threadIdx.x;
  const int64_t row = (int64_t)j * FC_MAXCOLS;
  if (c < m) {
    const int64_t o = row + c, o1 = o + FC_MAXCOLS;
    const double b = sqrt(so[c]);
    const double sc = s.scale[c];
    const bool ok = (b > 1e-290) && (b > 1e-13 * sc) && (s.inv_beta[o] != 0.0);
    s.beta[o1] = ok ? b : 0.0;
    s.inv_beta[o1] = ok ? 1.0 / b : 0.0;
    s.ratio_b[o1] = ok ? b * s.inv_beta[o] : 0.0;
    s.scale[c] = fmax(sc, b);
  }
  __syncthreads();
  double mx = 0.0;
  for (int idx = threadIdx.x; idx < s.ne * m; idx += blockDim.x) {
    const int e = idx / m, cc = idx % m;
    const cx<double> z = s.z[e];
    const double al = s.alpha[row + cc], bj = s.beta[row + cc], bn = s.beta[row + FC_MAXCOLS + cc];
    const int64_t so_ = (int64_t)e * FC_MAXCOLS + cc;
    cx<double> d, g;
    if (j == 0) {
      d = mk<double>(z.x - al, z.y);
      g = mk<double>(1.0, 0.0) / d;
    } else {
      const cx<double> dp = s.d[so_];
      d = mk<double>(z.x - al, z.y) - mk<double>(bj * bj, 0.0) / dp;
      g = (bj * s.g[so_]) / d;
    }
    s.d[so_] = d;
    s.g[so_] = g;
    const double r = bn * sqrt(abs2(g));
    mx = fmax(mx, r);
  }
  __syncthreads();
  tmp[threadIdx.x] = mx;
  __syncthreads();
  for (int w = blockDim.x / 2; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) tmp[threadIdx.x] = fmax(tmp[threadIdx.x], tmp[threadIdx.x + w]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    s.maxres[j] = tmp[0];
    if (tmp[0] <= s.target) *s.done_k = j + 1;
  }
}

// Tail of a producer kernel: the LAST CTA to finish (ticket counter) sums the per-CTA partial rows in a fixed order, exchanges the
// sums with the other ranks and runs the scalar recurrences that used to be separate one-CTA launches.
enum LzTailKind { LZ_TAIL_NONE = 0, LZ_TAIL_INIT = 1, LZ_TAIL_ALPHA = 2, LZ_TAIL_BETA = 3, LZ_TAIL_BARRIER = 4, LZ_TAIL_SIGNAL = 5 };
constexpr int LZ_TAIL_GROUP = 16;   // CTAs per first-level group of the two-level tail reduction
struct LzTail {
  int kind;          // LzTailKind
  int j;             // Lanczos step
  int* ticket;       // device counters, zero between launches: [0] groups finished, [1 + g] CTAs of group g finished
  double* grows;     // [ngroups][FC_MAXCOLS] group sums (first level of the reduction)
  LzScalars s;
  LzXchg x;
};

// every thread of the CTA calls this after the CTA's partial row is written (or with nothing written for LZ_TAIL_BARRIER);
// blockDim.x must be a power of two >= 128
// scratch: FC_MAXCOLS + blockDim.x doubles of shared memory the caller no longer needs
__device__ __forceinline__ void lz_tail(const LzTail& t, const double* partial, int pstride, int m, double* scratch) {
  if (t.kind == LZ_TAIL_NONE) return;
  __shared__ int s_last;
  double* t_so = scratch;
  double* t_tmp = scratch + FC_MAXCOLS;
  __threadfence();      // this CTA's rows are visible device-wide before its ticket; the CTA that raises the cross-rank flag adds the
  __syncthreads();      // system-scope fence (lz_exchange), which is cumulative over everything it has observed
  // two levels, both in a fixed order: the last CTA of every group of LZ_TAIL_GROUP consecutive CTAs sums the group's rows, the last
  // group to finish sums the group sums (a single CTA summing several hundred rows would cost more than the kernels it replaced)
  const int grp = (int)blockIdx.x / LZ_TAIL_GROUP, ngroups = ((int)gridDim.x + LZ_TAIL_GROUP - 1) / LZ_TAIL_GROUP;
  const int gsize = min(LZ_TAIL_GROUP, (int)gridDim.x - grp * LZ_TAIL_GROUP);
  if (threadIdx.x == 0) {
    const int tk = atomicAdd(t.ticket + 1 + grp, 1);
    s_last = (tk == gsize - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (t.kind != LZ_TAIL_BARRIER && t.kind != LZ_TAIL_SIGNAL) {
    lz_reduce_rows(partial + (int64_t)grp * LZ_TAIL_GROUP * pstride, gsize, pstride, m, t_so, t_tmp);
    if ((int)threadIdx.x < m) t.grows[(int64_t)grp * FC_MAXCOLS + threadIdx.x] = t_so[threadIdx.x];
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    t.ticket[1 + grp] = 0;
    const int tk = atomicAdd(t.ticket, 1);
    s_last = (tk == ngroups - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (t.kind == LZ_TAIL_BARRIER) {
    lz_exchange(t.x, t_so, 0);
  } else if (t.kind == LZ_TAIL_SIGNAL) {
    lz_signal(t.x);
  } else {
    lz_reduce_rows(t.grows, ngroups, FC_MAXCOLS, m, t_so, t_tmp);
    lz_exchange(t.x, t_so, m);
    if (t.kind == LZ_TAIL_INIT) lz_scalars_init(t.s, t_so, m);
    else if (t.kind == LZ_TAIL_ALPHA) lz_scalars_alpha(t.s, t.j, t_so, m);
    else lz_scalars_beta(t.s, t.j, t_so, m, t_tmp);
  }
  if (threadIdx.x == 0) *t.ticket = 0;
}

// one-CTA forms (matrix-free and generalized paths, whose producers are not fused)
__global__ void __launch_bounds__(1024) k_lz_scal_init(LzScalars s, const double* partial, int nblocks, int pstride, int m) {
  __shared__ double so[FC_MAXCOLS];
  __shared__ double tmp[1024];
  reduce_partials<double>(partial, 1, nblocks, pstride, m, so, tmp);
  lz_scalars_init(s, so, m);
}
__global__ void __launch_bounds__(1024) k_lz_scal1(LzScalars s, int j, const double* partial, int nblocks, int pstride, int m) {
  if (*s.done_k != 0) return;
  __shared__ double so[FC_MAXCOLS];
  __shared__ double tmp[1024];
  reduce_partials<double>(partial, 1, nblocks, pstride, m, so, tmp);
  lz_scalars_alpha(s, j, so, m);
}
__global__ void __launch_bounds__(1024) k_lz_scal2(LzScalars s, int j, const double* partial, int nblocks, int pstride, int m) {
  if (*s.done_k != 0) return;
  __shared__ double so[FC_MAXCOLS];
  __shared__ double tmp[1024];
  reduce_partials<double>(partial, 1, nblocks, pstride, m, so, tmp);
  lz_scalars_beta(s, j, so, m, tmp);
}

struct LzArgs {
  int64_t n;
  int m;                 // active columns
  int64_t ld;            // row stride in doubles (even)
  const int* ptr; const int* col; const void* val;   // val: double (real symmetric) or cx<double> (complex Hermitian)
  const double* U;       // gathered operand
  const double* prev;    // own-row operand u_{j-1}; at j = 0 any finite block (the host passes U; ratio_b is 0 there)
  double* out;           // own-row result (may alias prev)
  double* Q;             // accumulator (LZ_P2, LZ_RES)
  const double* s_inv_beta; const double* s_ratio_b; const double* s_ratio_a;   // per-column scalars of this step
  const double* s_coef;  // LZ_P2: c_j / beta_j ; LZ_RES: rho(theta_c)
  const double* s_theta; // LZ_RES
  double* partial;       // [gridDim.x][pstride]
  int pstride;
  int tile_rows;         // rows per round-robin tile (rounded to the CTA's rows-per-iteration)
  const int* done;       // device flag: nonzero once pass 1 has converged -> the launch is a no-op (nullptr: always run)
  const double* s_coef_prev;   // LZ_P2_PAIR: c_{j-1} / beta_{j-1}
  const double* own;     // LZ_RES: own-row operand of the theta term (generalized problems: B q); nullptr -> U
  const double* rhs;     // LZ_CHEB*: right-hand side block (own row)
  const double* dinv;    // LZ_CHEB*: 1 / diag(M), one double per row
  double c1, c2;         // LZ_CHEB*: recurrence coefficients of this step
  LzTail tail;           // what the last CTA does after the row loop (pass 1: the step's scalars; sharded pass 2: the step barrier)
  // row-sharded runs: `n` rows are this rank's block; the uploaded column index of an entry is (owner rank << LZ_OWNER_SHIFT) | row
  // local to the owner, resolved once per row stride into `goff`
  const int* goff;       // row-sharded runs: per stored entry, the distance (16-byte units) from a local block to the gathered row (k_lz_resolve)
  const int* tile_order; // row-sharded runs: the order in which the row tiles are dealt to the CTAs (nullptr: natural order); halo tiles last
  int halo_start;        // position in tile_order of the first tile that reads (and is read by) a peer
  unsigned long long wait_seq;          // > 0: before its first halo tile a warp waits until every peer has completed kernel wait_seq
  const unsigned long long* kdone;      // the local mailbox's kdone[] (written by the peers)
  int nranks, rank;
};
constexpr int LZ_OWNER_SHIFT = 26;
struct LzArenas { const char* base[LZ_MAXRANKS]; };   // arena base of every rank as mapped in this process
__host__ __device__ constexpr bool lz_is_cheb(int mode) { return mode == LZ_CHEB || mode == LZ_CHEB_DOT; }

__device__ __forceinline__ double2 ldg2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void stg2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }

// Element semantics.  Real problems: a 16-byte element is a PAIR of real columns, matrix entries are doubles.  Complex
// Hermitian problems (CPLX): an element is ONE complex column (re, im), matrix entries are cx<double>; T_k stays real, so
// the per-column scalars apply to both components and a dot product is the sum of the two component products.
template <bool CPLX> struct LzVal;
template <> struct LzVal<false> {
  typedef double T;
  static __device__ __forceinline__ T zero() { return 0.0; }
  static __device__ __forceinline__ T load(const void* v, int i) { return reinterpret_cast<const double*>(v)[i]; }
  static __device__ __forceinline__ T shfl(unsigned mask, T a, int src, int width) { return __shfl_sync(mask, a, src, width); }
  static __device__ __forceinline__ void fma_acc(double2& acc, T a, double2 x) { acc.x = fma(a, x.x, acc.x); acc.y = fma(a, x.y, acc.y); }
};
template <> struct LzVal<true> {
  typedef double2 T;
  static __device__ __forceinline__ T zero() { return make_double2(0.0, 0.0); }
  static __device__ __forceinline__ T load(const void* v, int i) { return reinterpret_cast<const double2*>(v)[i]; }
  static __device__ __forceinline__ T shfl(unsigned mask, T a, int src, int width) {
    return make_double2(__shfl_sync(mask, a.x, src, width), __shfl_sync(mask, a.y, src, width));
  }
  static __device__ __forceinline__ void fma_acc(double2& acc, T a, double2 x) {
    acc.x = fma(a.x, x.x, acc.x); acc.x = fma(-a.y, x.y, acc.x);
    acc.y = fma(a.x, x.y, acc.y); acc.y = fma(a.y, x.x, acc.y);
  }
};
// per-element scalars from a per-column array: real -> (s[2e], s[2e+1]); complex -> (s[e], s[e])
template <bool CPLX>
__device__ __forceinline__ double2 lz_scal(const double* s, int e, int m) {
  double2 v = make_double2(0.0, 0.0);
  if (s != nullptr) {
    if (CPLX) { if (e < m) v.x = v.y = s[e]; }
    else { if (2 * e < m) v.x = s[2 * e]; if (2 * e + 1 < m) v.y = s[2 * e + 1]; }
  }
  return v;
}
template <bool CPLX> __device__ __forceinline__ int lz_elems(int m) { return CPLX ? m : ((m + 1) >> 1); }

// acc[k] += sum_p val[p] * U[col[p], pair(g + G k)]  for the stored entries [p0, p1) of `row`, in CSR order.
// (myo, mya): the group's first G entries, already fetched by the caller one iteration ahead; myo is the ELEMENT offset
// col*ld of the gathered row (32-bit: the host guarantees n*ld < 2^32), so an address is one IMAD.WIDE from the lane's
// base pointer Ul[k] = U + 2*(g + G k).  Padding slots of the last batch gather the lane's own row with weight 0 and lanes
// beyond the active columns read a clamped (valid) column: the loop body carries no predicates at all.
// Narrow groups (G <= 4 lanes per row) also receive the row's SECOND chunk of G entries prefetched (myo2, mya2): a
// 7-point row then needs no dependent metadata load inside the loop even with 4 lanes per row.
// Row-sharded runs (SHARD): the metadata of an entry is a pre-resolved SIGNED 32-bit distance, in 16-byte units, from the local block to
// the gathered row (k_lz_resolve: the ranks' arenas are mapped into one contiguous virtual range, peer_arena.hpp, and the owner's copy
// of a block sits at the same arena offset as the local one), so a halo row in a peer's HBM is read by the same load as a local row --
// over NVLink -- at the same instruction count as the single-GPU kernel.
template <bool SHARD> struct LzOff;
template <> struct LzOff<false> {
  typedef unsigned T;
  static __device__ __forceinline__ T meta(const LzArgs& a, int p, unsigned ldu) { return (unsigned)a.col[p] * ldu; }
  static __device__ __forceinline__ T own(unsigned eo_own) { return eo_own; }
  static __device__ __forceinline__ const double* at(const double* base, T o) { return base + o; }
};
template <> struct LzOff<true> {
  typedef int T;
  static __device__ __forceinline__ T meta(const LzArgs& a, int p, unsigned) { return a.goff[p]; }
  static __device__ __forceinline__ T own(unsigned eo_own) { return (int)(eo_own >> 1); }
  static __device__ __forceinline__ const double* at(const double* base, T o) { return base + 2 * (long long)o; }
};

template <int G, int NC, bool CPLX, bool SHARD>
__device__ __forceinline__ void lz_gather(const LzArgs& a, unsigned row_eo_, int p0, int p1, typename LzOff<SHARD>::T myo, typename LzVal<CPLX>::T mya,
                                          typename LzOff<SHARD>::T myo2, typename LzVal<CPLX>::T mya2, int g, unsigned gmask,
                                          const double* const (&Ul)[NC], double2 (&acc)[NC]) {
  typedef LzVal<CPLX> V;
  typedef LzOff<SHARD> OF;
  constexpr int UN = (G >= 4) ? 4 : G;
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
      mya = V::zero();
      if (g < cnt) { myo = OF::meta(a, pb + g, ldu); mya = V::load(a.val, pb + g); }
    }
    for (int t = 0; t < cnt; t += UN) {
      typename OF::T eo[UN];
      typename V::T aa[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        eo[u] = __shfl_sync(gmask, myo, t + u, G);
        aa[u] = V::shfl(gmask, mya, t + u, G);
      }
      double2 xv[UN][NC];
#pragma unroll
      for (int u = 0; u < UN; ++u)
#pragma unroll
        for (int k = 0; k < NC; ++k) xv[u][k] = ldg2(OF::at(Ul[k], eo[u]));
#pragma unroll
      for (int u = 0; u < UN; ++u)
#pragma unroll
        for (int k = 0; k < NC; ++k) V::fma_acc(acc[k], aa[u], xv[u][k]);
    }
  }
}

__device__ __forceinline__ double2 lz_scal2(const double* s, int c0, int m) {
  double2 v = make_double2(0.0, 0.0);
  if (s != nullptr) {
    if (c0 < m) v.x = s[c0];
    if (c0 + 1 < m) v.y = s[c0 + 1];
  }
  return v;
}

// the two Lanczos vector updates, shared by pass 1 (split over two kernels) and pass 2 (fused): identical rounding
__device__ __forceinline__ double2 lz_t(double2 acc, double2 ib, double2 rb, double2 pv) {
  double2 t;
  t.x = __fma_rn(-rb.x, pv.x, __dmul_rn(acc.x, ib.x));
  t.y = __fma_rn(-rb.y, pv.y, __dmul_rn(acc.y, ib.y));
  return t;
}
__device__ __forceinline__ double2 lz_next(double2 t, double2 ra, double2 uo) {
  double2 r;
  r.x = __fma_rn(-ra.x, uo.x, t.x);
  r.y = __fma_rn(-ra.y, uo.y, t.y);
  return r;
}

template <int G, int NC, int MODE, int THREADS, bool CPLX = false, bool SHARD = false>
__global__ void __launch_bounds__(THREADS, (NC >= 3) ? 1 : 1024 / THREADS) k_lz_spmm(LzArgs a) {
  typedef LzVal<CPLX> V;
  typedef LzOff<SHARD> OF;
  typedef typename OF::T off_t;
  if (a.done != nullptr && *a.done != 0) return;
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x & 31, g = lane % G, sub = lane / G;
  constexpr unsigned gm0 = (G >= 32) ? 0xffffffffu : ((1u << (G & 31)) - 1u);
  const unsigned gmask = gm0 << (sub * G);
  const int wib = threadIdx.x >> 5, wpb = THREADS >> 5;
  constexpr int STEP = (THREADS / 32) * RPW;          // rows the CTA covers per iteration
  const int P = lz_elems<CPLX>(a.m);                  // 16-byte elements per row: column pairs (real) or complex columns
  const int n = (int)a.n;
  // Rows are dealt to the CTAs in tiles of `tile_rows` consecutive rows, round robin: the whole grid sweeps the matrix as
  // one moving front, so a vector row fetched as somebody's far neighbour is still in L2 when the front reaches it
  // (each row of U comes from HBM once), while the near neighbours of a tile stay in the CTA's L1.
  const int spt = max(1, a.tile_rows / STEP);         // iterations per tile
  const int TR = spt * STEP;
  const int ntiles = (n + TR - 1) / TR;
  const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int niter = my_tiles * spt;
  // iteration -> row, advanced incrementally (no integer divisions in the loop): `base_f` is the first row of the CTA's
  // iteration `it_f`, the furthest one the metadata pipeline has looked at
  // Row-sharded runs deal the tiles in the order `tile_order` (host: the tiles that touch halo rows are spread evenly among the
  // interior ones, so the NVLink latency of a halo gather is hidden behind local rows instead of stalling every CTA at once at
  // both ends of the sweep).
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

  // per-column scalars of this step live in shared memory (one 16-byte read per use instead of 20 live registers)
  constexpr int SCW = CPLX ? FC_MAXCOLS : FC_MAXCOLS / 2;   // elements per row: complex columns or real column pairs
  __shared__ double2 s_sc[5][SCW];   // [0] inv_beta | theta, [1] ratio_b, [2] ratio_a, [3] coef, [4] coef of the previous step
  for (int i = threadIdx.x; i < 5 * SCW; i += THREADS) {
    const int w = i / SCW, pc = i % SCW;
    const double* src = (w == 0) ? (MODE == LZ_RES ? a.s_theta : a.s_inv_beta)
                                 : (w == 1 ? a.s_ratio_b : (w == 2 ? a.s_ratio_a : (w == 3 ? a.s_coef : a.s_coef_prev)));
    s_sc[w][pc] = lz_scal<CPLX>(src, pc, a.m);
  }
  __syncthreads();
  double2 dot[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) dot[k] = make_double2(0.0, 0.0);

  // software pipeline over the CSR metadata: row pointers two iterations ahead, the first G (offset, val) pairs one
  // iteration ahead, so the vector gathers of an iteration never wait behind a pointer chase.  (Measured on B200: a
  // deeper pipeline with L2 prefetch of the next iteration's rows costs more in registers/spills than it hides.)
  const unsigned ldu = (unsigned)a.ld;
  const double* Ul[NC];
  const double* Pl[NC];   // prev: never null (the host passes U with ratio_b = 0 at j = 0)
  double* Ol[NC];
  double* Ql[NC];
  const double* Wl[NC];   // LZ_RES: own-row operand of the theta term; LZ_CHEB*: the right-hand side
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    const int pcl = 2 * min(g + G * k, P - 1);
    Ul[k] = a.U + pcl;
    Pl[k] = a.prev + pcl;
    Ol[k] = a.out + pcl;
    Ql[k] = a.Q + pcl;
    Wl[k] = (lz_is_cheb(MODE) ? a.rhs : (a.own != nullptr ? a.own : a.U)) + pcl;
  }
  int r_cur = next_row(), r_nxt = next_row();
  int p0_cur = 0, p1_cur = 0, p0_nxt = 0, p1_nxt = 0;
  if (r_cur < n) { p0_cur = a.ptr[r_cur]; p1_cur = a.ptr[r_cur + 1]; }
  if (r_nxt < n) { p0_nxt = a.ptr[r_nxt]; p1_nxt = a.ptr[r_nxt + 1]; }
  constexpr bool PF2 = (G <= 4);
  off_t o_cur = OF::own((r_cur < n ? (unsigned)r_cur : 0u) * ldu), o_cur2 = o_cur;
  typename V::T a_cur = V::zero(), a_cur2 = V::zero();
  if (g < p1_cur - p0_cur) { o_cur = OF::meta(a, p0_cur + g, ldu); a_cur = V::load(a.val, p0_cur + g); }
  if (PF2 && g + G < p1_cur - p0_cur) { o_cur2 = OF::meta(a, p0_cur + G + g, ldu); a_cur2 = V::load(a.val, p0_cur + G + g); }

  // first iteration of this CTA that touches a halo tile (tiles are dealt in the order of `tile_order`, halo tiles last)
  int it_h = 0x7fffffff;
  if (SHARD && a.wait_seq != 0 && a.tile_order != nullptr) {
    const int first = a.halo_start > (int)blockIdx.x ? (a.halo_start - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    it_h = first * spt;
  }
  for (int it = 0; it < niter; ++it) {
    if (SHARD && it == it_h) {
      // the peers' previous kernel is complete: their rows of the gathered block are final, and they no longer read the rows this
      // kernel is about to overwrite
      if (lane < a.nranks && lane != a.rank) { while (ld_sys_u64(a.kdone + lane) < a.wait_seq) { } }
      __syncwarp();
    }
    const int row = r_cur;
    const bool valid = row < n;
    const int r_fut = next_row();
    int p0_fut = 0, p1_fut = 0;
    if (r_fut < n) { p0_fut = a.ptr[r_fut]; p1_fut = a.ptr[r_fut + 1]; }
    off_t o_nxt = OF::own((r_nxt < n ? (unsigned)r_nxt : 0u) * ldu), o_nxt2 = o_nxt;
    typename V::T a_nxt = V::zero(), a_nxt2 = V::zero();
    if (g < p1_nxt - p0_nxt) { o_nxt = OF::meta(a, p0_nxt + g, ldu); a_nxt = V::load(a.val, p0_nxt + g); }
    if (PF2 && g + G < p1_nxt - p0_nxt) { o_nxt2 = OF::meta(a, p0_nxt + G + g, ldu); a_nxt2 = V::load(a.val, p0_nxt + G + g); }

    // own-row operands first (independent of the gather); invalid rows and inactive lanes read valid dummies
    const unsigned eo_own = (valid ? (unsigned)row : 0u) * ldu;
    double2 acc[NC], uo[NC], pv[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      acc[k] = make_double2(0.0, 0.0);
      if constexpr (MODE == LZ_RES) uo[k] = ldg2(Wl[k] + eo_own);
      else if constexpr (MODE != LZ_PLAIN) uo[k] = ldg2(Ul[k] + eo_own);
      if constexpr (MODE == LZ_P1 || lz_is_p2(MODE) || lz_is_cheb(MODE)) pv[k] = ldg2(Pl[k] + eo_own);
    }
    double2 rh[NC];
    double di = 0.0;
    if constexpr (lz_is_cheb(MODE)) {
#pragma unroll
      for (int k = 0; k < NC; ++k) rh[k] = ldg2(Wl[k] + eo_own);
      di = a.dinv[valid ? row : 0] * a.c2;
    }
    lz_gather<G, NC, CPLX, SHARD>(a, eo_own, p0_cur, p1_cur, o_cur, a_cur, o_cur2, a_cur2, g, gmask, Ul, acc);
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const int pc = g + G * k;
      if (valid && pc < P) {
        if constexpr (MODE == LZ_PLAIN) {
          stg2(Ol[k] + eo_own, acc[k]);
        } else if constexpr (lz_is_cheb(MODE)) {
          double2 o;
          o.x = fma(di, rh[k].x - acc[k].x, fma(a.c1, uo[k].x - pv[k].x, uo[k].x));
          o.y = fma(di, rh[k].y - acc[k].y, fma(a.c1, uo[k].y - pv[k].y, uo[k].y));
          stg2(Ol[k] + eo_own, o);
          if constexpr (MODE == LZ_CHEB_DOT) {
            dot[k].x = fma(o.x, rh[k].x, dot[k].x);
            dot[k].y = fma(o.y, rh[k].y, dot[k].y);
          }
        } else if constexpr (MODE == LZ_RES) {
          const double2 th = s_sc[0][pc], cf = s_sc[3][pc];
          double2 t;
          t.x = __fma_rn(-th.x, uo[k].x, acc[k].x);
          t.y = __fma_rn(-th.y, uo[k].y, acc[k].y);
          stg2(Ol[k] + eo_own, t);
          dot[k].x = fma(t.x, t.x, dot[k].x);
          dot[k].y = fma(t.y, t.y, dot[k].y);
          if (a.Q != nullptr) stg2(Ql[k] + eo_own, make_double2(cf.x * uo[k].x, cf.y * uo[k].y));
        } else {
          const double2 t = lz_t(acc[k], s_sc[0][pc], s_sc[1][pc], pv[k]);
          if constexpr (MODE == LZ_P1) {
            stg2(Ol[k] + eo_own, t);
            dot[k].x = fma(uo[k].x, t.x, dot[k].x);
            dot[k].y = fma(uo[k].y, t.y, dot[k].y);
          } else {
            stg2(Ol[k] + eo_own, lz_next(t, s_sc[2][pc], uo[k]));
            if constexpr (MODE != LZ_P2_SKIP) {
              const double2 cf = s_sc[3][pc];
              double2 q = ldg2(Ql[k] + eo_own);
              if constexpr (MODE == LZ_P2_PAIR) {
                const double2 cp = s_sc[4][pc];
                q.x = fma(cp.x, pv[k].x, q.x);
                q.y = fma(cp.y, pv[k].y, q.y);
              }
              q.x = fma(cf.x, uo[k].x, q.x);
              q.y = fma(cf.y, uo[k].y, q.y);
              stg2(Ql[k] + eo_own, q);
            }
          }
        }
      }
    }
    r_cur = r_nxt; p0_cur = p0_nxt; p1_cur = p1_nxt; o_cur = o_nxt; a_cur = a_nxt; o_cur2 = o_nxt2; a_cur2 = a_nxt2;
    r_nxt = r_fut; p0_nxt = p0_fut; p1_nxt = p1_fut;
  }

  __shared__ double2 red[(THREADS / 32) * 32 * NC];   // CTA reduction of the dot products, then the tail's scratch
  if constexpr (MODE == LZ_P1 || MODE == LZ_RES || MODE == LZ_CHEB_DOT) {
    // fixed-order CTA reduction: every launch sums in the same order (pass 2 relies on pass 1's exact scalars)
    const int width = G * NC;   // pairs per row group
#pragma unroll
    for (int k = 0; k < NC; ++k) red[(wib * RPW + sub) * width + g + G * k] = dot[k];
    __syncthreads();
    const int ngroups = wpb * RPW;
    for (int pc = threadIdx.x; pc < width; pc += THREADS) {
      if (pc < P) {
        double sx = 0.0, sy = 0.0;
        for (int q = 0; q < ngroups; ++q) { const double2 v = red[q * width + pc]; sx += v.x; sy += v.y; }
        if (CPLX) a.partial[(int64_t)blockIdx.x * a.pstride + pc] = sx + sy;   // Re(conj(u) t) = sum of both components
        else {
          double* o = a.partial + (int64_t)blockIdx.x * a.pstride + 2 * pc;
          o[0] = sx;
          if (2 * pc + 1 < a.m) o[1] = sy;
        }
      }
    }
  }
  __syncthreads();
  lz_tail(a.tail, a.partial, a.pstride, a.m, reinterpret_cast<double*>(red));
}

// ---- elementwise kernels on real blocks: a thread owns one column PAIR, rows strided ---------------------
struct EwMap2 {
  int pc, rsub, rpb;
  __device__ __forceinline__ EwMap2(int pp) { pc = threadIdx.x % pp; rsub = threadIdx.x / pp; rpb = blockDim.x / pp; }
};

template <bool CPLX>
__device__ __forceinline__ void block_reduce_pairs(double2 v, int pp, int P, int m, double* out_row) {
  __shared__ double2 red2[256];
  __syncthreads();
  red2[threadIdx.x] = v;
  __syncthreads();
  if ((int)threadIdx.x < pp && (int)threadIdx.x < P) {
    double sx = 0.0, sy = 0.0;
    for (int q = threadIdx.x; q < (int)blockDim.x; q += pp) { sx += red2[q].x; sy += red2[q].y; }
    if (CPLX) out_row[threadIdx.x] = sx + sy;
    else {
      out_row[2 * threadIdx.x] = sx;
      if (2 * (int)threadIdx.x + 1 < m) out_row[2 * threadIdx.x + 1] = sy;
    }
  }
}

// pass 1, second half of a step: T (in place) <- T - ratio_a * U ; partial = |T|^2
template <bool CPLX>
__global__ void __launch_bounds__(256) k_lz_update(int64_t n, int m, int pp, int64_t ld, const double* __restrict__ s_ratio_a,
                                                   const double* __restrict__ U, double* __restrict__ T,
                                                   double* __restrict__ partial, int pstride, const int* __restrict__ done,
                                                   LzTail tail) {
  if (done != nullptr && *done != 0) return;
  EwMap2 e(pp);
  const int P = lz_elems<CPLX>(m);
  double2 acc = make_double2(0.0, 0.0);
  if (e.pc < P) {
    const double2 ra = lz_scal<CPLX>(s_ratio_a, e.pc, m);
    const int64_t stride = (int64_t)gridDim.x * e.rpb;
    int64_t row = (int64_t)blockIdx.x * e.rpb + e.rsub;
    for (; row + 3 * stride < n; row += 4 * stride) {   // four independent rows in flight per thread (narrow blocks are latency bound)
      double2 tv[4], uv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t off = (row + q * stride) * ld + 2 * e.pc;
        tv[q] = ldg2(T + off);
        uv[q] = ldg2(U + off);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t off = (row + q * stride) * ld + 2 * e.pc;
        const double2 r = lz_next(tv[q], ra, uv[q]);
        stg2(T + off, r);
        acc.x = fma(r.x, r.x, acc.x);
        acc.y = fma(r.y, r.y, acc.y);
      }
    }
    for (; row < n; row += stride) {
      const int64_t off = row * ld + 2 * e.pc;
      const double2 r = lz_next(ldg2(T + off), ra, ldg2(U + off));
      stg2(T + off, r);
      acc.x = fma(r.x, r.x, acc.x);
      acc.y = fma(r.y, r.y, acc.y);
    }
  }
  block_reduce_pairs<CPLX>(acc, pp, P, m, partial + (int64_t)blockIdx.x * pstride);
  __shared__ double tail_scratch[FC_MAXCOLS + 256];
  lz_tail(tail, partial, pstride, m, tail_scratch);
}

// Q += coef * U  (last pass-2 step: no further Lanczos vector is needed)
template <bool CPLX>
__global__ void __launch_bounds__(256) k_lz_axpy(int64_t n, int m, int pp, int64_t ld, const double* __restrict__ s_coef,
                                                 const double* __restrict__ U, double* __restrict__ Q) {
  EwMap2 e(pp);
  const int P = lz_elems<CPLX>(m);
  if (e.pc >= P) return;
  const double2 cf = lz_scal<CPLX>(s_coef, e.pc, m);
  for (int64_t row = (int64_t)blockIdx.x * e.rpb + e.rsub; row < n; row += (int64_t)gridDim.x * e.rpb) {
    const int64_t off = row * ld + 2 * e.pc;
    const double2 u = ldg2(U + off);
    double2 q = ldg2(Q + off);
    q.x = fma(cf.x, u.x, q.x);
    q.y = fma(cf.y, u.y, q.y);
    stg2(Q + off, q);
  }
}

// first Chebyshev step from X_0 = 0:  X = c * dinv[row] * R
template <bool CPLX>
__global__ void __launch_bounds__(256) k_lz_cheb_first(int64_t n, int m, int pp, int64_t ld, double c, const double* __restrict__ dinv,
                                                       const double* __restrict__ R, double* __restrict__ X, const int* __restrict__ done) {
  if (done != nullptr && *done != 0) return;
  EwMap2 e(pp);
  const int P = lz_elems<CPLX>(m);
  if (e.pc >= P) return;
  for (int64_t row = (int64_t)blockIdx.x * e.rpb + e.rsub; row < n; row += (int64_t)gridDim.x * e.rpb) {
    const int64_t off = row * ld + 2 * e.pc;
    const double f = c * dinv[row];
    const double2 r = ldg2(R + off);
    stg2(X + off, make_double2(f * r.x, f * r.y));
  }
}

// engine block (complex storage, row stride ldz) -> compact Lanczos block.  Real problems keep the real part of the m columns
// as column pairs (the pad column of an odd m is zeroed); complex problems copy the m complex columns.  partial = |x|^2
template <bool CPLX>
__global__ void __launch_bounds__(256) k_lz_real_part(int64_t n, int m, int pp, int64_t ldz, int64_t ld,
                                                      const cx<double>* __restrict__ Z, double* __restrict__ X,
                                                      double* __restrict__ partial, int pstride, LzTail tail) {
  EwMap2 e(pp);
  const int P = lz_elems<CPLX>(m);
  double2 acc = make_double2(0.0, 0.0);
  if (e.pc < P) {
    for (int64_t row = (int64_t)blockIdx.x * e.rpb + e.rsub; row < n; row += (int64_t)gridDim.x * e.rpb) {
      double2 v;
      if (CPLX) {
        const cx<double> z = Z[row * ldz + e.pc];
        v = make_double2(z.x, z.y);
      } else {
        const int c0 = 2 * e.pc;
        v.x = Z[row * ldz + c0].x;
        v.y = (c0 + 1 < m) ? Z[row * ldz + c0 + 1].x : 0.0;
      }
      stg2(X + row * ld + 2 * e.pc, v);
      acc.x = fma(v.x, v.x, acc.x);
      acc.y = fma(v.y, v.y, acc.y);
    }
  }
  if (partial != nullptr) block_reduce_pairs<CPLX>(acc, pp, P, m, partial + (int64_t)blockIdx.x * pstride);
  __shared__ double tail_scratch[FC_MAXCOLS + 256];
  lz_tail(tail, partial, pstride, m, tail_scratch);
}

// cross-rank barrier of row-sharded runs (before a kernel gathers rows that the peers' previous kernel wrote)
__global__ void __launch_bounds__(128) k_lz_barrier(LzXchg x) {
  __shared__ double dummy[1];
  lz_exchange(x, dummy, 0);
}

// pre-resolved gather offsets of a row-sharded operator: enc = (owner << LZ_OWNER_SHIFT) | row local to the owner
__global__ void __launch_bounds__(256) k_lz_resolve(int64_t nnz, const int* __restrict__ enc, long long row_bytes, int self,
                                                    LzArenas ar, int* __restrict__ goff) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned e = (unsigned)enc[i];
    const int owner = (int)(e >> LZ_OWNER_SHIFT);
    const long long local = (long long)(e & ((1u << LZ_OWNER_SHIFT) - 1u));
    goff[i] = (int)(((long long)(ar.base[owner] - ar.base[self]) + local * row_bytes) / 16);
  }
}

// compact Lanczos block -> engine block (real problems: zero imaginary part)
template <bool CPLX>
__global__ void __launch_bounds__(256) k_lz_to_complex(int64_t n, int m, int pp, int64_t ld, int64_t ldz,
                                                       const double* __restrict__ X, cx<double>* __restrict__ Z) {
  EwMap2 e(pp);
  const int P = lz_elems<CPLX>(m);
  if (e.pc >= P) return;
  for (int64_t row = (int64_t)blockIdx.x * e.rpb + e.rsub; row < n; row += (int64_t)gridDim.x * e.rpb) {
    const double2 v = ldg2(X + row * ld + 2 * e.pc);
    if (CPLX) Z[row * ldz + e.pc] = mk<double>(v.x, v.y);
    else {
      const int c0 = 2 * e.pc;
      Z[row * ldz + c0] = mk<double>(v.x, 0.0);
      if (c0 + 1 < m) Z[row * ldz + c0 + 1] = mk<double>(v.y, 0.0);
    }
  }
}

}  // namespace feastcuda
