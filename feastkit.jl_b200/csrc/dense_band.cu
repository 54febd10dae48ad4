// Dense and banded operator paths of libfeastcuda: per-node LU factors cached on the device (the reference caches
// `lu(z*B - A)` per node too, dense/feast_dense.jl:186-204, banded/feast_banded.jl:100-112), direct solves on the
// engine's row-major right-hand-side blocks, operator application for the Rayleigh-Ritz stage.
#include "dense_band.cuh"

#include <algorithm>
#include <cstring>
#include <cstdlib>

#include "kernels_band.cuh"
#include "kernels_dense.cuh"

namespace feastcuda {

typedef feastcuda_handle_s H;

static void launched(H* h) {
  h->stats.kernel_launches++;
  FC_CUDA(cudaGetLastError());
}

// =====================================================================================================
// dense
// =====================================================================================================
// the caller's column-major matrix goes straight to the device (no host copy is kept); real input is widened there
void dense_set(H* h, int which, int64_t n, const double* a, int64_t lda, bool cplx, int structure) {
  (void)structure;
  FC_REQUIRE(n > 0 && a != nullptr && lda >= n, "set_dense: bad arguments");
  FC_REQUIRE(n < 2147483647 / 2, "set_dense: n too large");
  FC_REQUIRE(which == FEASTCUDA_A || which == FEASTCUDA_B, "which must be A or B");
  HostDense& d = (which == FEASTCUDA_A) ? h->denseA : h->denseB;
  if (which == FEASTCUDA_B) FC_REQUIRE(h->kind == OP_DENSE && h->denseA.set && h->denseA.n == n, "set A (same size, dense) before B");
  DBuf& dst = (which == FEASTCUDA_A) ? h->dDenseA : h->dDenseB;
  const size_t es = cplx ? sizeof(zd) : sizeof(double);
  dst.ensure((size_t)n * n * sizeof(zd));
  if (cplx) {
    FC_CUDA(cudaMemcpy2DAsync(dst.p, (size_t)n * es, a, (size_t)lda * es, (size_t)n * es, (size_t)n, cudaMemcpyHostToDevice, h->stream));
  } else {
    h->stage.ensure((size_t)n * n * es);
    FC_CUDA(cudaMemcpy2DAsync(h->stage.p, (size_t)n * es, a, (size_t)lda * es, (size_t)n * es, (size_t)n, cudaMemcpyHostToDevice, h->stream));
    const int64_t total = n * n;
    k_dense_widen<<<(int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)h->sms * 8)), 256, 0, h->stream>>>(
        total, h->stage.as<double>(), dst.as<zd>());
    launched(h);
  }
  FC_CUDA(cudaStreamSynchronize(h->stream));
  d.n = n;
  d.cplx = cplx;
  d.a.clear();
  d.set = true;
  if (which == FEASTCUDA_A) {
    if (h->kind != OP_DENSE || (h->has_b && h->denseB.n != n)) { h->has_b = false; h->denseB.set = false; }   // a B of another size never survives a new A
    h->kind = OP_DENSE;
    h->n = n;
  } else {
    FC_REQUIRE(h->kind == OP_DENSE && h->denseA.set && h->denseA.n == n, "set A (same size, dense) before B");
    h->has_b = true;
  }
  h->dense_uploaded = true;
  release_factor_cache(h);
}

void dense_prepare(H* h) {
  FC_REQUIRE(h->denseA.set, "operator A has not been set");
  h->n = h->denseA.n;
}

static void zgemm(H* h, int M, int N, int K, const zd* A, int64_t ars, int64_t acs, const zd* B, int64_t brs, int64_t bcs, zd* C,
                  int64_t crs, int64_t ccs, double alpha, int beta, int batch = 1, int64_t ab = 0, int64_t bb = 0, int64_t cb = 0) {
  if (M <= 0 || N <= 0) return;
  ZgemmArgs g;
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.ars = ars; g.acs = acs; g.abatch = ab;
  g.B = B; g.brs = brs; g.bcs = bcs; g.bbatch = bb;
  g.C = C; g.crs = crs; g.ccs = ccs; g.cbatch = cb;
  g.alpha = alpha; g.beta = beta;
  dim3 grid((M + 63) / 64, (N + 63) / 64, batch);
  static const int variant = getenv("FEASTCUDA_GEMM") ? atoi(getenv("FEASTCUDA_GEMM")) : 2;   // 2: cp.async double-buffered operands (default), 1: synchronous staging
  if (variant == 2) {
    static bool attr = false;
    if (!attr) { FC_CUDA(cudaFuncSetAttribute(k_zgemm_dmma_async, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); attr = true; }
    k_zgemm_dmma_async<<<grid, 256, 64 * 1024, h->stream>>>(g);
  } else {
    k_zgemm_dmma<<<grid, 256, 0, h->stream>>>(g);
  }
  launched(h);
}

// ---- batched factorisation / solves: all quadrature nodes of a sweep go through every launch together ---------------
// dense_pool holds one column-major n x n factor per contour node (slot = node index), dense_piv [ipiv(n) | perm(n)] per
// slot, dense_xpool one n x ld solution block per node.  The single-CTA panel kernels and the small triangular solves are
// latency bound: batching over nodes is what fills the GPU there; the trailing GEMMs are batched through blockIdx.z.
static void dense_ensure_pools(H* h, int nslots) {
  const int64_t n = h->n;
  if (h->dense_slots < nslots) {
    h->dense_pool.release();
    h->dense_piv.release();
    h->dense_pool.ensure((size_t)nslots * n * n * sizeof(zd));
    h->dense_piv.ensure((size_t)nslots * 2 * n * sizeof(int));
    h->dense_slots = nslots;
    h->lu_shift.assign(nslots, zc(NAN, NAN));
  }
  if ((int)h->lu_shift.size() < nslots) h->lu_shift.resize(nslots, zc(NAN, NAN));
}

// factorise z_q B - A for the contiguous node range [first, first+count)
static bool dense_factor_range(H* h, int first, int count, const zc* shifts) {
  const int n = (int)h->n;
  const int64_t nn = (int64_t)n * n, pst = 2 * (int64_t)n;
  zd* LU = h->dense_pool.as<zd>() + (int64_t)first * nn;
  int* ipiv = h->dense_piv.as<int>() + (int64_t)first * pst;
  DBuf dz, dinfo, dpart;
  dz.ensure((size_t)count * sizeof(zd));
  dinfo.ensure((size_t)count * sizeof(int));
  // CTAs per node of the cooperative panel kernel: enough to spread a tall panel, few enough to be co-resident
  static const int coop_env = getenv("FEASTCUDA_PANEL_CPN") ? atoi(getenv("FEASTCUDA_PANEL_CPN")) : -1;
  int cpn = 1;
  if (n >= 2048) cpn = std::max(1, std::min(32, h->sms / std::max(1, count)));
  if (coop_env >= 1) cpn = coop_env;
  dpart.ensure((size_t)count * cpn * (sizeof(double) + sizeof(int)) + 64);
  std::vector<zd> zz(count);
  for (int q = 0; q < count; ++q) zz[q] = mk<double>(shifts[q].real(), shifts[q].imag());
  FC_CUDA(cudaMemcpyAsync(dz.p, zz.data(), (size_t)count * sizeof(zd), cudaMemcpyHostToDevice, h->stream));
  FC_CUDA(cudaMemsetAsync(dinfo.p, 0, (size_t)count * sizeof(int), h->stream));
  {
    dim3 grid((unsigned)std::min<int64_t>((nn + 255) / 256, (int64_t)h->sms * 4), count);
    k_dense_shift<<<grid, 256, 0, h->stream>>>(n, h->dDenseA.as<zd>(), h->has_b ? h->dDenseB.as<zd>() : nullptr, dz.as<zd>(), LU, nn);
    launched(h);
  }
  // Two-level blocking.  Panels are FC_LU_NB = 32 columns wide (latency-bound pivot searches), but a rank-32 update of the whole trailing
  // matrix reads and writes C for 8 flop per byte (ncu: 1.7 TB/s, tensor pipe 37 % busy at n = 8192).  So the panels of an OUTER block
  // of `nbo` columns update only the rest of that block; the columns to its right get the block's row interchanges, the triangular
  // solve with the block's unit-lower L11 (panel by panel) and ONE rank-nbo update when the block is complete (ncu, rank 128 at n = 8192:
  // 24.2 TFLOP/s = 0.66 of the ZGEMM rate measured on the box, DMMA pipe 65 % busy).  nbo = 32 is the one-level right-looking algorithm of
  // round 1 (same operations in the same order).
  static const int nbo_env = getenv("FEASTCUDA_DENSE_NBO") ? atoi(getenv("FEASTCUDA_DENSE_NBO")) : 256;   // measured at n = 8192: 32 -> 1 298 ms, 128 -> 1 049, 256 -> 937, 512 -> 954
  const int nbo = std::max(FC_LU_NB, (nbo_env / FC_LU_NB) * FC_LU_NB);
  auto u12 = [&](int k0, int nbw, int c0, int c1) {       // U12 = L11^-1 A12 for the columns [c0, c1) (row interchanges already applied)
    if (c1 <= c0) return;
    k_dense_trsm_u12<<<dim3((c1 - c0 + 127) / 128, count), 128, 0, h->stream>>>(n, k0, nbw, c0, c1, LU, nn);
    launched(h);
  };
  for (int o0 = 0; o0 < n; o0 += nbo) {
    const int oend = std::min(n, o0 + nbo);
    for (int k0 = o0; k0 < oend; k0 += FC_LU_NB) {
      const int nbw = std::min(FC_LU_NB, oend - k0);
      if (cpn > 1) {
        int n_ = n, k0_ = k0, nbw_ = nbw, cpn_ = cpn;
        zd* lu_ = LU;
        int64_t nn_ = nn, pst_ = pst;
        int* ipiv_ = ipiv;
        int* info_ = dinfo.as<int>();
        double* pv_ = dpart.as<double>();
        int* pi_ = reinterpret_cast<int*>(dpart.as<double>() + (size_t)count * cpn);
        void* args[] = {&n_, &k0_, &nbw_, &lu_, &nn_, &ipiv_, &pst_, &info_, &cpn_, &pv_, &pi_};
        FC_CUDA(cudaLaunchCooperativeKernel((void*)k_dense_panel_lu_coop, dim3(count * cpn), dim3(512), args, 0, h->stream));
      } else {
        k_dense_panel_lu<<<count, 512, 0, h->stream>>>(n, k0, nbw, LU, nn, ipiv, pst, dinfo.as<int>());
      }
      launched(h);
      if (k0 > 0) {
        k_dense_laswp<<<dim3((k0 + 255) / 256, count), 256, 0, h->stream>>>(n, k0, nbw, 0, k0, LU, nn, ipiv, pst);
        launched(h);
      }
      const int c0 = k0 + nbw, inblk = oend - c0;          // the rest of the outer block
      if (inblk > 0) {
        k_dense_laswp<<<dim3((inblk + 255) / 256, count), 256, 0, h->stream>>>(n, k0, nbw, c0, oend, LU, nn, ipiv, pst);
        launched(h);
        u12(k0, nbw, c0, oend);
        // A22[c0:n, c0:oend] -= L21[c0:n, k0:c0] * U12[k0:c0, c0:oend]   (all column-major, ld = n), every node in the same launch
        zgemm(h, n - c0, inblk, nbw, LU + c0 + (int64_t)k0 * n, 1, n, LU + k0 + (int64_t)c0 * n, 1, n, LU + c0 + (int64_t)c0 * n, 1, n,
              -1.0, 1, count, nn, nn, nn);
      }
    }
    const int rest = n - oend;
    if (rest > 0) {
      // the columns right of the block: all of the block's row interchanges in order, U12 = L11^-1 A12 panel by panel, one rank-(oend - o0) update
      k_dense_laswp<<<dim3((rest + 255) / 256, count), 256, 0, h->stream>>>(n, o0, oend - o0, oend, n, LU, nn, ipiv, pst);
      launched(h);
      for (int k0 = o0; k0 < oend; k0 += FC_LU_NB) {
        const int nbw = std::min(FC_LU_NB, oend - k0), below = oend - k0 - nbw;
        u12(k0, nbw, oend, n);
        if (below > 0)   // A12[k0+nbw:oend, oend:n] -= L11[k0+nbw:oend, k0:k0+nbw] * U12[k0:k0+nbw, oend:n]
          zgemm(h, below, rest, nbw, LU + (k0 + nbw) + (int64_t)k0 * n, 1, n, LU + k0 + (int64_t)oend * n, 1, n,
                LU + (k0 + nbw) + (int64_t)oend * n, 1, n, -1.0, 1, count, nn, nn, nn);
      }
      zgemm(h, rest, rest, oend - o0, LU + oend + (int64_t)o0 * n, 1, n, LU + o0 + (int64_t)oend * n, 1, n, LU + oend + (int64_t)oend * n, 1, n,
            -1.0, 1, count, nn, nn, nn);
    }
  }
  k_dense_piv_to_perm<<<count, 32, 0, h->stream>>>(n, ipiv, ipiv + n, pst);
  launched(h);
  std::vector<int> info(count, 0);
  FC_CUDA(cudaMemcpyAsync(info.data(), dinfo.p, (size_t)count * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  FC_CUDA(cudaStreamSynchronize(h->stream));
  dz.release();
  dinfo.release();
  dpart.release();
  bool all_ok = true;
  for (int q = 0; q < count; ++q) {
    if (info[q] != 0) { all_ok = false; h->lu_shift[first + q] = zc(NAN, NAN); }
    else h->lu_shift[first + q] = shifts[q];
  }
  return all_ok;
}

// X_q = (z_q B - A)^-1 RHS for the node range [first, first+count); X_q = Xbase + q * xbatch (row-major n x ld blocks)
static void dense_solve_range(H* h, int first, int count, int m, const zd* RHS, zd* Xbase, int64_t xbatch) {
  const int n = (int)h->n;
  const int64_t ld = h->ws_ld, nn = (int64_t)n * n, pst = 2 * (int64_t)n;
  const zd* LU = h->dense_pool.as<zd>() + (int64_t)first * nn;
  const int* perm = h->dense_piv.as<int>() + (int64_t)first * pst + n;
  {
    const int64_t total = (int64_t)n * m;
    const int gx = (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)h->sms * 4));
    k_dense_gather_rows<<<dim3(gx, count), 256, 0, h->stream>>>(n, m, ld, perm, pst, RHS, Xbase, xbatch);
    launched(h);
  }
  const int cgrid = (m + 127) / 128;
  for (int r0 = 0; r0 < n; r0 += FC_LU_NB) {     // forward: L y = P b
    const int bs = std::min(FC_LU_NB, n - r0);
    k_dense_trsm_rows<true><<<dim3(cgrid, count), 128, 0, h->stream>>>(n, r0, bs, LU, nn, m, ld, Xbase, xbatch);
    launched(h);
    const int rest = n - r0 - bs;
    if (rest > 0)
      zgemm(h, rest, m, bs, LU + (r0 + bs) + (int64_t)r0 * n, 1, n, Xbase + (int64_t)r0 * ld, ld, 1, Xbase + (int64_t)(r0 + bs) * ld, ld, 1,
            -1.0, 1, count, nn, xbatch, xbatch);
  }
  const int nblk = (n + FC_LU_NB - 1) / FC_LU_NB;
  for (int b = nblk - 1; b >= 0; --b) {          // backward: U x = y
    const int r0 = b * FC_LU_NB, bs = std::min(FC_LU_NB, n - r0);
    k_dense_trsm_rows<false><<<dim3(cgrid, count), 128, 0, h->stream>>>(n, r0, bs, LU, nn, m, ld, Xbase, xbatch);
    launched(h);
    if (r0 > 0)
      zgemm(h, r0, m, bs, LU + (int64_t)r0 * n, 1, n, Xbase + (int64_t)r0 * ld, ld, 1, Xbase, ld, 1, -1.0, 1, count, nn, xbatch, xbatch);
  }
}

bool dense_node_solve(H* h, int node, zc z, int m, const zd* RHS, zd* X) {
  dense_ensure_pools(h, std::max(node + 1, h->dense_slots));
  if (!(h->lu_shift[node] == z)) {
    if (!dense_factor_range(h, node, 1, &z)) return false;   // exactly singular shifted matrix -> LAPACK info > 0 (dense/feast_dense.jl:198-203)
  }
  dense_solve_range(h, node, 1, m, RHS, X, 0);
  return true;
}

// all nodes [first, first+count) of a sweep at once; solutions land in dense_xpool (slot q -> node first+q)
bool dense_batch_solve(H* h, int ne_total, int first, int count, const zc* shifts, int m, const zd* RHS, zd** Xpool, int64_t* xbatch) {
  dense_ensure_pools(h, ne_total);
  bool need = false;
  for (int q = 0; q < count; ++q) need = need || !(h->lu_shift[first + q] == shifts[q]);
  if (need && !dense_factor_range(h, first, count, shifts)) return false;
  const int64_t xb = (int64_t)h->n * h->ws_ld;
  h->dense_xpool.ensure((size_t)count * xb * sizeof(zd));
  dense_solve_range(h, first, count, m, RHS, h->dense_xpool.as<zd>(), xb);
  *Xpool = h->dense_xpool.as<zd>();
  *xbatch = xb;
  return true;
}

void dense_apply(H* h, int which, int m, const zd* X, zd* Y) {
  const int n = (int)h->n;
  const int64_t ld = h->ws_ld;
  const zd* Mx = (which == FEASTCUDA_A) ? h->dDenseA.as<zd>() : h->dDenseB.as<zd>();
  zgemm(h, n, m, n, Mx, 1, n, X, ld, 1, Y, ld, 1, 1.0, 0);
}

// =====================================================================================================
// banded: operators are expanded to general band storage (2k+1) x n, diagonal in row k (banded/feast_banded.jl:263-271);
// factors in LAPACK gbtrf layout (3k+1) x n with kl = ku = k.
// =====================================================================================================
void band_set(H* h, int which, int64_t n, int64_t k, const double* ab, int64_t ldab, bool cplx, int structure) {
  FC_REQUIRE(n > 0 && k >= 0 && ab != nullptr, "set_band: bad arguments");
  FC_REQUIRE(which == FEASTCUDA_A || which == FEASTCUDA_B, "which must be A or B");
  const bool general = (structure == FEASTCUDA_GEN);
  FC_REQUIRE(ldab >= (general ? 2 * k + 1 : k + 1), "set_band: storage insufficient for the bandwidth");
  HostBand& d = (which == FEASTCUDA_A) ? h->bandA : h->bandB;
  if (which == FEASTCUDA_B) FC_REQUIRE(h->kind == OP_BAND && h->bandA.set && h->bandA.n == n, "set A (same size, banded) before B");
  d.n = n;
  d.k = k;
  d.cplx = cplx;
  const int64_t rows = 2 * k + 1;
  d.ab.assign((size_t)2 * rows * n, 0.0);
  auto in = [&](int64_t r, int64_t j) -> zc {
    const size_t o = (size_t)j * ldab + r;
    return cplx ? zc(ab[2 * o], ab[2 * o + 1]) : zc(ab[o], 0.0);
  };
  auto out = [&](int64_t i, int64_t j, zc v) {   // entry (i, j) of the full matrix, |i - j| <= k
    const size_t o = 2 * ((size_t)j * rows + (k + i - j));
    d.ab[o] = v.real();
    d.ab[o + 1] = v.imag();
  };
  for (int64_t j = 0; j < n; ++j) {
    if (general) {
      for (int64_t i = std::max<int64_t>(0, j - k); i <= std::min<int64_t>(n - 1, j + k); ++i) out(i, j, in(k + i - j, j));
    } else {
      // upper storage (k+1) x n, A[i,j] (i <= j) in row k + i - j (banded/feast_banded.jl:205-214); mirror with conj
      for (int64_t i = std::max<int64_t>(0, j - k); i <= j; ++i) {
        const zc v = in(k + i - j, j);
        out(i, j, v);
        if (i != j) out(j, i, std::conj(v));
      }
    }
  }
  d.set = true;
  if (which == FEASTCUDA_A) {
    if (h->kind != OP_BAND || (h->has_b && h->bandB.n != n)) { h->has_b = false; h->bandB.set = false; }   // a B of another size never survives a new A
    h->kind = OP_BAND;
    h->n = n;
  } else {
    FC_REQUIRE(h->kind == OP_BAND && h->bandA.set && h->bandA.n == n, "set A (same size, banded) before B");
    h->has_b = true;
  }
  h->band_uploaded = false;
  release_factor_cache(h);
}

void band_prepare(H* h) {
  FC_REQUIRE(h->bandA.set, "operator A has not been set");
  h->n = h->bandA.n;
  if (h->band_uploaded) return;
  auto up = [&](const HostBand& s, DBuf& d) {
    const size_t bytes = s.ab.size() * sizeof(double);
    d.ensure(bytes);
    FC_CUDA(cudaMemcpyAsync(d.p, s.ab.data(), bytes, cudaMemcpyHostToDevice, h->stream));
  };
  up(h->bandA, h->dBandA);
  if (h->has_b) up(h->bandB, h->dBandB);
  FC_CUDA(cudaStreamSynchronize(h->stream));
  h->band_uploaded = true;
  release_factor_cache(h);
}

// 3 (default): batched over nodes, LU with the active window in a shared-memory ring + lane-parallel substitutions (k <= 15);
// 2: warp LU on global memory + register-window substitutions; 1: the one-CTA LU / thread-per-column solve of round 1
static int band_impl() {
  static const int v = getenv("FEASTCUDA_BAND_IMPL") ? atoi(getenv("FEASTCUDA_BAND_IMPL")) : 3;
  return v;
}

static int band_halfwidth(H* h) { return std::max((int)h->bandA.k, h->has_b ? (int)h->bandB.k : 0); }

static void band_ensure_node(H* h, int node) {
  if ((int)h->lu_cache.size() <= node) {
    h->lu_cache.resize(node + 1);
    h->piv_cache.resize(node + 1);
    h->lu_shift.resize(node + 1, zc(NAN, NAN));
  }
}

// factor (z_q B - A) for every node of [first, first + count) whose cached shift differs; all of them in ONE LU launch
static bool band_factor_range(H* h, int first, int count, const zc* shifts) {
  const int n = (int)h->n;
  const int ka = (int)h->bandA.k, kb = h->has_b ? (int)h->bandB.k : 0;
  const int k = std::max(ka, kb);
  const int ldf = 3 * k + 1;
  band_ensure_node(h, first + count - 1);
  std::vector<int> todo;
  for (int q = 0; q < count; ++q)
    if (!(h->lu_shift[first + q] == shifts[q])) todo.push_back(q);
  if (todo.empty()) return true;
  const bool warp_lu = band_impl() >= 2 && k <= FC_BAND_LU_MAXK;
  for (int q : todo) {
    const int node = first + q;
    h->lu_shift[node] = zc(NAN, NAN);
    h->lu_cache[node].ensure((size_t)ldf * n * sizeof(zd));
    h->piv_cache[node].ensure((size_t)(n + 1) * sizeof(int));
    FC_CUDA(cudaMemsetAsync(h->piv_cache[node].as<int>() + n, 0, sizeof(int), h->stream));
    const int64_t total = (int64_t)ldf * n;
    k_band_shift<<<(int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)h->sms * 8)), 256, 0, h->stream>>>(
        n, k, ka, kb, h->dBandA.as<zd>(), h->has_b ? h->dBandB.as<zd>() : nullptr, mk<double>(shifts[q].real(), shifts[q].imag()),
        h->lu_cache[node].as<zd>());
    launched(h);
    if (!warp_lu) {
      k_band_lu<<<1, 256, 0, h->stream>>>(n, k, h->lu_cache[node].as<zd>(), h->piv_cache[node].as<int>(), h->piv_cache[node].as<int>() + n);
      launched(h);
    }
  }
  if (warp_lu) {
    for (size_t t0 = 0; t0 < todo.size(); t0 += FC_BAND_BATCH) {
      const int cnt = (int)std::min<size_t>(FC_BAND_BATCH, todo.size() - t0);
      BandBatch bb;
      memset(&bb, 0, sizeof(bb));
      for (int t = 0; t < cnt; ++t) {
        bb.F[t] = h->lu_cache[first + todo[t0 + t]].as<zd>();
        bb.ipiv[t] = h->piv_cache[first + todo[t0 + t]].as<int>();
      }
      const int ev = sample_begin(h, FEASTCUDA_KERN_BAND_LU);
      if (band_impl() >= 3 && k <= 7) k_band_lu_smem<7, 32><<<cnt, 32, 0, h->stream>>>(n, k, bb);
      else if (band_impl() >= 3 && k <= 15) k_band_lu_smem<15, 64><<<cnt, 32, 0, h->stream>>>(n, k, bb);
      else k_band_lu_warp<<<cnt, 32, 0, h->stream>>>(n, k, bb);
      launched(h);
      sample_end(h, ev);
      // algorithmic bytes of the launch: every factor is read once and written once (the 5 KB window of a step lives in L1), pivots written
      h->stats.bytes_kern[FEASTCUDA_KERN_BAND_LU] = (double)cnt * (2.0 * ldf * (double)n * sizeof(zd) + 4.0 * n);
    }
  }
  std::vector<int> info(todo.size(), 0);
  for (size_t t = 0; t < todo.size(); ++t)
    FC_CUDA(cudaMemcpyAsync(&info[t], h->piv_cache[first + todo[t]].as<int>() + n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  FC_CUDA(cudaStreamSynchronize(h->stream));
  bool all_ok = true;
  for (size_t t = 0; t < todo.size(); ++t) {
    if (info[t] != 0) all_ok = false;                       // exactly singular shifted matrix -> LAPACK info > 0 (banded:108-112)
    else h->lu_shift[first + todo[t]] = shifts[todo[t]];
  }
  return all_ok;
}

template <int G>
static void band_launch_lanes(H* h, int n, int k, const BandBatch& bb, int cnt, int m, int64_t ld, const zd* RHS, zd* X, int64_t xbatch) {
  constexpr int CPW = 32 / G;
  k_band_solve_lanes<G><<<dim3((m + CPW - 1) / CPW, cnt), 32, 0, h->stream>>>(n, k, bb, m, ld, RHS, X, xbatch);
  launched(h);
}

template <int K>
static void band_launch_win(H* h, int n, int k, const BandBatch& bb, int cnt, int m, int64_t ld, const zd* RHS, zd* X, int64_t xbatch) {
  k_band_solve_win<K><<<dim3((m + 31) / 32, cnt), 32, 0, h->stream>>>(n, k, bb, m, ld, RHS, X, xbatch);
  launched(h);
}

// X_q = (z_q B - A)^-1 RHS for the node range [first, first + count) with cached factors; X_q = Xbase + q * xbatch (row-major n x ld)
static void band_solve_range(H* h, int first, int count, int m, const zd* RHS, zd* Xbase, int64_t xbatch) {
  const int n = (int)h->n;
  const int k = band_halfwidth(h);
  const int ldf = 3 * k + 1;
  const int64_t ld = h->ws_ld;
  if (band_impl() >= 2 && k <= FC_BAND_WIN_MAXK) {
    FC_REQUIRE(RHS != Xbase, "band solve: the solution block must not alias the right-hand side");
    for (int q0 = 0; q0 < count; q0 += FC_BAND_BATCH) {
      const int cnt = std::min(FC_BAND_BATCH, count - q0);
      BandBatch bb;
      memset(&bb, 0, sizeof(bb));
      for (int t = 0; t < cnt; ++t) {
        bb.F[t] = h->lu_cache[first + q0 + t].as<zd>();
        bb.ipiv[t] = h->piv_cache[first + q0 + t].as<int>();
      }
      zd* X = Xbase + (int64_t)q0 * xbatch;
      const int ev = sample_begin(h, FEASTCUDA_KERN_BAND_SOLVE);
      if (band_impl() >= 3 && k <= 1) band_launch_lanes<4>(h, n, k, bb, cnt, m, ld, RHS, X, xbatch);
      else if (band_impl() >= 3 && k <= 3) band_launch_lanes<8>(h, n, k, bb, cnt, m, ld, RHS, X, xbatch);
      else if (band_impl() >= 3 && k <= 7) band_launch_lanes<16>(h, n, k, bb, cnt, m, ld, RHS, X, xbatch);
      else if (band_impl() >= 3 && k <= 15) band_launch_lanes<32>(h, n, k, bb, cnt, m, ld, RHS, X, xbatch);
      else if (k <= 2) band_launch_win<2>(h, n, k, bb, cnt, m, ld, RHS, X, xbatch);
      else if (k <= 4) band_launch_win<4>(h, n, k, bb, cnt, m, ld, RHS, X, xbatch);
      else if (k <= 8) band_launch_win<8>(h, n, k, bb, cnt, m, ld, RHS, X, xbatch);
      else band_launch_win<16>(h, n, k, bb, cnt, m, ld, RHS, X, xbatch);
      sample_end(h, ev);
      // per node (minimum traffic): the factor is read once per sweep, the pivots once, x moves four times (RHS -> y -> x)
      h->stats.bytes_kern[FEASTCUDA_KERN_BAND_SOLVE] =
          (double)cnt * (2.0 * ldf * (double)n * sizeof(zd) + 4.0 * n + 4.0 * (double)n * m * sizeof(zd));
    }
    return;
  }
  for (int q = 0; q < count; ++q) {
    zd* X = Xbase + (int64_t)q * xbatch;
    FC_CUDA(cudaMemcpy2DAsync(X, (size_t)ld * sizeof(zd), RHS, (size_t)ld * sizeof(zd), (size_t)m * sizeof(zd), (size_t)n,
                              cudaMemcpyDeviceToDevice, h->stream));
    k_band_solve<<<(m + 63) / 64, 64, 0, h->stream>>>(n, k, h->lu_cache[first + q].as<zd>(), h->piv_cache[first + q].as<int>(), m, ld, X);
    launched(h);
  }
}

bool band_node_solve(H* h, int node, zc z, int m, const zd* RHS, zd* X) {
  if (!band_factor_range(h, node, 1, &z)) return false;
  band_solve_range(h, node, 1, m, RHS, X, 0);
  return true;
}

// all nodes [first, first+count) of a sweep at once: one LU launch for the nodes that need a factor, one solve launch for all
// (node, column group) pairs; solutions land in the shared solution pool (slot q -> node first+q).  Returns false without touching
// anything when the pool would not fit (the caller then solves node by node).
bool band_batch_fits(H* h, int count) {
  return band_impl() >= 2 && (double)count * (double)h->n * (double)h->ws_ld * sizeof(zd) <= 32.0 * 1073741824.0;
}

bool band_batch_solve(H* h, int first, int count, const zc* shifts, int m, const zd* RHS, zd** Xpool, int64_t* xbatch) {
  if (!band_factor_range(h, first, count, shifts)) return false;
  const int64_t xb = (int64_t)h->n * h->ws_ld;
  h->dense_xpool.ensure((size_t)count * xb * sizeof(zd));
  band_solve_range(h, first, count, m, RHS, h->dense_xpool.as<zd>(), xb);
  *Xpool = h->dense_xpool.as<zd>();
  *xbatch = xb;
  return true;
}

void band_apply(H* h, int which, int m, const zd* X, zd* Y) {
  const int n = (int)h->n;
  const HostBand& s = (which == FEASTCUDA_A) ? h->bandA : h->bandB;
  const zd* AB = (which == FEASTCUDA_A) ? h->dBandA.as<zd>() : h->dBandB.as<zd>();
  const int64_t total = (int64_t)n * m;
  k_band_apply<<<(int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)h->sms * 8)), 256, 0, h->stream>>>(
      n, (int)s.k, AB, m, h->ws_ld, X, Y);
  launched(h);
}

}  // namespace feastcuda
