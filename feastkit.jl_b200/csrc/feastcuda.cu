// libfeastcuda: engine + C ABI.  See include/feastcuda.h for the contract and DESIGN.md for the layout.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "engine.hpp"
#include "host_math.hpp"
#include "kernels_block.cuh"
#include "kernels_reduced.cuh"
#include "kernels_sparse.cuh"
#include "kernels_lanczos.cuh"
#include "kernels_lanczos_f32.cuh"
#include "kernels_matfree.cuh"
#include "kernels_twosided.cuh"
#include "dense_band.cuh"

using namespace feastcuda;
typedef feastcuda_handle_s H;

static thread_local std::string g_last_error;

// =====================================================================================================
// small utilities
// =====================================================================================================
static inline int pow2_ge(int m) { int p = 1; while (p < m) p <<= 1; return p; }
static inline zd tozd(zc v) { return mk<double>(v.real(), v.imag()); }

static void* pinned_buf(H* h, size_t bytes) {
  if (bytes > h->pinned_cap) {
    if (h->pinned) cudaFreeHost(h->pinned);
    h->pinned = nullptr;
    h->pinned_cap = 0;
    FC_CUDA(cudaMallocHost(&h->pinned, bytes));
    h->pinned_cap = bytes;
  }
  return h->pinned;
}

static void sync(H* h) { FC_CUDA(cudaStreamSynchronize(h->stream)); }
static void check_launch(H* h) { h->stats.kernel_launches++; FC_CUDA(cudaGetLastError()); }

static zd* blk(H* h, int slot) { return h->blk[slot].as<zd>(); }

static void arena_allocate(H* h, int ld);
static void fill_xchg(H* h, LzXchg& x);
static const int* resolve_goff(H* h, int64_t rowbytes);
static void xbarrier(H* h);
static void allreduce_small(H* h, double* dev, size_t count);
static void sharded_apply(H* h, int m, const cx<double>* X, cx<double>* Y, const double* theta, std::vector<double>* norms2);

static void ensure_workspace(H* h, int64_t n, int m0) {
  FC_REQUIRE(m0 >= 1 && m0 <= FC_MAXCOLS,
             "M0 must be in [1,128]: the block kernels hold at most 128 columns (slice a wider interval into sub-intervals)");
  const int ld = std::max(m0, 1);
  if (h->ws_n != n || h->ws_ld < ld) {
    h->ws_n = n;
    h->ws_ld = ld;
    h->have_subspace = false;
    if (h->row_sharded) arena_allocate(h, ld);
    else for (int s = 0; s < BS_COUNT; ++s) h->blk[s].ensure((size_t)n * ld * sizeof(zd));
  }
  const int maxblocks = h->sms * 8;
  h->partial.ensure((size_t)3 * maxblocks * FC_MAXCOLS * sizeof(zd));
  h->partial_r.ensure((size_t)maxblocks * FC_MAXCOLS * sizeof(double));
  h->kstate.ensure(sizeof(KrylovState<double>));
  h->small.ensure((size_t)8 * FC_MAXCOLS * FC_MAXCOLS * sizeof(zd));
  h->small2.ensure((size_t)4 * FC_MAXCOLS * sizeof(zd) + 64);
  h->red_ws.ensure((size_t)4 * FC_MAXCOLS * sizeof(zd));
}

static int ew_grid(H* h, int64_t n, int mp) {
  const int rpb = 256 / mp;
  int64_t g = (n + rpb - 1) / rpb;
  return (int)std::max<int64_t>(1, std::min<int64_t>(g, (int64_t)h->sms * 8));
}

// =====================================================================================================
// operators: CSR ingest (SparseMatrixCSC -> device CSR)
// =====================================================================================================
static void ingest_csr(HostCsr& out, int64_t n, int64_t nnz, const int64_t* ptr, const int64_t* idx, const double* val,
                       bool cplx, int base, int fmt, int structure) {
  FC_REQUIRE(n > 0 && nnz >= 0 && ptr && (nnz == 0 || (idx && val)), "set_csr: bad arguments");
  FC_REQUIRE(nnz < (int64_t)2147483647 && n < (int64_t)2147483647, "set_csr: int32 index range exceeded");
  FC_REQUIRE(ptr[n] - base == nnz && ptr[0] - base == 0, "set_csr: pointer array inconsistent with nnz");
  for (int64_t i = 0; i < n; ++i) FC_REQUIRE(ptr[i] <= ptr[i + 1], "set_csr: pointer array must be non-decreasing");
  out.n = n;
  out.nnz = nnz;
  out.cplx = cplx;
  out.structure = structure;
  const int vs = cplx ? 2 : 1;
  out.ptr.assign(n + 1, 0);
  out.col.assign(nnz, 0);
  out.val.assign((size_t)nnz * vs, 0.0);
  const bool transpose = (fmt == FEASTCUDA_CSC) && (structure == FEASTCUDA_GEN);
  const bool conjugate = (fmt == FEASTCUDA_CSC) && (structure == FEASTCUDA_HERM) && cplx;
  if (!transpose) {
    for (int64_t i = 0; i <= n; ++i) out.ptr[i] = (int)(ptr[i] - base);
    for (int64_t p = 0; p < nnz; ++p) {
      const int64_t j = idx[p] - base;
      FC_REQUIRE(j >= 0 && j < n, "set_csr: index out of range");
      out.col[p] = (int)j;
      if (cplx) { out.val[2 * p] = val[2 * p]; out.val[2 * p + 1] = conjugate ? -val[2 * p + 1] : val[2 * p + 1]; }
      else out.val[p] = val[p];
    }
  } else {
    std::vector<int> cnt(n + 1, 0);
    for (int64_t p = 0; p < nnz; ++p) {
      const int64_t j = idx[p] - base;
      FC_REQUIRE(j >= 0 && j < n, "set_csr: index out of range");
      cnt[j + 1]++;
    }
    for (int64_t i = 0; i < n; ++i) cnt[i + 1] += cnt[i];
    for (int64_t i = 0; i <= n; ++i) out.ptr[i] = cnt[i];
    std::vector<int> pos(cnt.begin(), cnt.end() - 1);
    for (int64_t c = 0; c < n; ++c)
      for (int64_t p = ptr[c] - base; p < ptr[c + 1] - base; ++p) {
        const int64_t rrow = idx[p] - base;
        const int q = pos[rrow]++;
        out.col[q] = (int)c;
        if (cplx) { out.val[2 * (size_t)q] = val[2 * p]; out.val[2 * (size_t)q + 1] = val[2 * p + 1]; }
        else out.val[q] = val[p];
      }
  }
  out.set = true;
}

static void upload_csr(H* h, const HostCsr& src, DevCsr& dst, bool as_complex) {
  dst.ptr.ensure((src.n + 1) * sizeof(int));
  dst.col.ensure(std::max<int64_t>(src.nnz, 1) * sizeof(int));
  FC_CUDA(cudaMemcpyAsync(dst.ptr.p, src.ptr.data(), (src.n + 1) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  if (src.nnz) FC_CUDA(cudaMemcpyAsync(dst.col.p, src.col.data(), src.nnz * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  if (as_complex) {
    std::vector<double> tmp;
    const double* v = src.val.data();
    if (!src.cplx) {
      tmp.assign((size_t)2 * src.nnz, 0.0);
      for (int64_t p = 0; p < src.nnz; ++p) tmp[2 * p] = src.val[p];
      v = tmp.data();
    }
    dst.val.ensure(std::max<int64_t>(src.nnz, 1) * sizeof(zd));
    if (src.nnz) FC_CUDA(cudaMemcpyAsync(dst.val.p, v, src.nnz * sizeof(zd), cudaMemcpyHostToDevice, h->stream));
    sync(h);
  } else {
    dst.val.ensure(std::max<int64_t>(src.nnz, 1) * sizeof(double));
    if (src.nnz) FC_CUDA(cudaMemcpyAsync(dst.val.p, src.val.data(), src.nnz * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    sync(h);
  }
  dst.uploaded = true;
}

// contiguous block of rows owned by `rank` (the rule of feastcuda_node_partition applied to rows) and the owner of a global row
static void row_block(int64_t n, int nranks, int rank, int64_t* r0, int64_t* cnt) {
  const int64_t base = n / nranks, rem = n % nranks;
  *r0 = rank * base + std::min<int64_t>(rank, rem);
  *cnt = base + (rank < rem ? 1 : 0);
}

// row-sharded upload: this rank's rows only; a column index becomes (owner << LZ_OWNER_SHIFT) | row local to the owner
static void upload_csr_rows(H* h, const HostCsr& src, DevCsr& dst) {
  FC_REQUIRE(!src.cplx, "row sharding: real symmetric operators only");
  const int64_t n = src.n;
  const int P = h->nranks;
  int64_t r0 = 0, cnt = 0;
  row_block(n, P, h->rank, &r0, &cnt);
  const int64_t base = n / P, rem = n % P, bound = rem * (base + 1);
  FC_REQUIRE(base + 1 < ((int64_t)1 << LZ_OWNER_SHIFT), "row sharding: too many rows per rank for the column encoding");
  const int p0 = src.ptr[r0], p1 = src.ptr[r0 + cnt];
  std::vector<int> ptr(cnt + 1), col((size_t)std::max(1, p1 - p0));
  for (int64_t i = 0; i <= cnt; ++i) ptr[i] = src.ptr[r0 + i] - p0;
  for (int p = p0; p < p1; ++p) {
    const int64_t c = src.col[p];
    int64_t owner, first;
    if (c < bound) { owner = c / (base + 1); first = owner * (base + 1); }
    else { owner = rem + (base > 0 ? (c - bound) / base : 0); first = bound + (owner - rem) * base; }
    col[p - p0] = (int)(((unsigned)owner << LZ_OWNER_SHIFT) | (unsigned)(c - first));
  }
  dst.ptr.ensure((cnt + 1) * sizeof(int));
  dst.col.ensure(col.size() * sizeof(int));
  dst.val.ensure((size_t)std::max(1, p1 - p0) * sizeof(double));
  FC_CUDA(cudaMemcpyAsync(dst.ptr.p, ptr.data(), (cnt + 1) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  FC_CUDA(cudaMemcpyAsync(dst.col.p, col.data(), col.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  if (p1 > p0) FC_CUDA(cudaMemcpyAsync(dst.val.p, src.val.data() + p0, (size_t)(p1 - p0) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  sync(h);
  dst.uploaded = true;
  h->nnz_loc = p1 - p0;
  for (int sl = 0; sl < 4; ++sl) h->goff_rowbytes4[sl] = 0;
  h->tile_order_tr = 0;
}

static void finalize_sparse(H* h) {
  FC_REQUIRE(h->hA.set, "operator A has not been set");
  const bool need_cplx = h->hA.cplx || (h->has_b && h->hB.cplx);
  if (need_cplx != h->dev_complex) { h->dA.uploaded = false; h->dB.uploaded = false; h->dev_complex = need_cplx; }
  if (h->row_sharded) {
    FC_REQUIRE(!h->has_b && !need_cplx, "row sharding serves standard real symmetric sparse problems (multi-shift Lanczos filter)");
    int64_t r0 = 0, cnt = 0;
    row_block(h->hA.n, h->nranks, h->rank, &r0, &cnt);
    FC_REQUIRE(cnt >= 1, "row sharding: fewer rows than ranks");
    h->n_glob = h->hA.n;
    h->row0 = r0;
    h->n = cnt;
    h->nloc_max = h->hA.n / h->nranks + ((h->hA.n % h->nranks) ? 1 : 0);
    if (!h->dA.uploaded) upload_csr_rows(h, h->hA, h->dA);
    return;
  }
  if (!h->dA.uploaded) upload_csr(h, h->hA, h->dA, need_cplx);
  if (h->has_b && !h->dB.uploaded) upload_csr(h, h->hB, h->dB, need_cplx);
  h->n = h->hA.n;
  h->n_glob = h->hA.n;
  h->row0 = 0;
  h->nnz_loc = h->hA.nnz;
}

// =====================================================================================================
// launches
// =====================================================================================================
struct OpDesc {   // Y = zb*(B X) + za*(A X); use_b = 0 drops the B term, B unset means identity
  zd za, zb;
  bool use_a, use_b;
};

static int spmm_grid(H* h, int64_t n, int m) {
  const int G = m <= 4 ? 4 : (m <= 8 ? 8 : (m <= 16 ? 16 : 32));
  const int rpi = 8 * (32 / G);
  int64_t g = (n + rpi - 1) / rpi;
  return (int)std::max<int64_t>(1, std::min<int64_t>(g, (int64_t)h->sms * 4));
}

template <typename TA, int MODE>
static void spmm_dispatch(H* h, SpmmArgs<double, TA>& a, int grid) {
  const int m = a.m;
#define FC_L(G, NC) k_spmm<double, TA, G, NC, MODE><<<grid, 256, 0, h->stream>>>(a)
  if (m <= 4) FC_L(4, 1);
  else if (m <= 8) FC_L(8, 1);
  else if (m <= 16) FC_L(16, 1);
  else if (m <= 32) FC_L(32, 1);
  else if (m <= 64) FC_L(32, 2);
  else if (m <= 96) FC_L(32, 3);
  else FC_L(32, 4);
#undef FC_L
}

// ---- sampled kernel timings: CUDA events on the launching stream around selected launches ---------------------
// sample_begin returns a slot (or -1 when the pool is exhausted); `tag` lets the caller discard samples later
int feastcuda::sample_begin(H* h, int kind, int tag) {
  if (h->ev_used + 2 > 512) return -1;
  while (h->ev_pool.size() < h->ev_used + 2) {
    cudaEvent_t e;
    FC_CUDA(cudaEventCreate(&e));
    h->ev_pool.push_back(e);
  }
  const int a = (int)h->ev_used, b = a + 1;
  h->ev_used += 2;
  FC_CUDA(cudaEventRecord(h->ev_pool[a], h->stream));
  h->ev_pending.push_back({kind, tag, a, b});
  return (int)h->ev_pending.size() - 1;
}
void feastcuda::sample_end(H* h, int slot) {
  if (slot >= 0) FC_CUDA(cudaEventRecord(h->ev_pool[h->ev_pending[slot].b], h->stream));
}
// the stream must be idle; samples whose tag is >= tag_limit are dropped (launches that ran as no-ops)
static void drain_events(H* h, int tag_limit = 2147483647) {
  for (auto& sm : h->ev_pending) {
    float ms = 0.f;
    if (sm.tag < tag_limit && cudaEventElapsedTime(&ms, h->ev_pool[sm.a], h->ev_pool[sm.b]) == cudaSuccess) {
      h->stats.ms_kern[sm.kind] += ms;
      h->stats.n_kern[sm.kind]++;
      if (sm.kind == FEASTCUDA_KERN_SPMM_Z) { h->stats.ms_spmm_sampled += ms; h->stats.spmm_sampled++; }
    }
  }
  h->ev_pending.clear();
  h->ev_used = 0;
}

// returns the grid size used (= number of partial rows written per slot)
template <int MODE>
static int launch_spmm(H* h, const OpDesc& op, int m, const zd* X, zd* Y, const zd* aux, const zd* lam, bool sample = false) {
  FC_REQUIRE(h->kind == OP_SPARSE, "sparse operator required");
  const int grid = spmm_grid(h, h->n, m);
  const int ev = sample ? sample_begin(h, FEASTCUDA_KERN_SPMM_Z) : -1;
  if (h->dev_complex) {
    SpmmArgs<double, zd> a;
    a.n = h->n; a.m = m; a.ld = h->ws_ld;
    a.a_ptr = op.use_a ? h->dA.ptr.as<int>() : nullptr; a.a_col = h->dA.col.as<int>(); a.a_val = h->dA.val.as<zd>();
    const bool bmat = op.use_b && h->has_b;
    a.b_ptr = bmat ? h->dB.ptr.as<int>() : nullptr; a.b_col = h->dB.col.as<int>(); a.b_val = h->dB.val.as<zd>();
    a.skip_b = op.use_b ? 0 : 1;
    a.za = op.za; a.zb = op.zb; a.X = X; a.Y = Y; a.aux = aux; a.lam = lam;
    a.partial = h->partial.as<zd>(); a.pstride = FC_MAXCOLS;
    spmm_dispatch<zd, MODE>(h, a, grid);
  } else {
    SpmmArgs<double, double> a;
    a.n = h->n; a.m = m; a.ld = h->ws_ld;
    a.a_ptr = op.use_a ? h->dA.ptr.as<int>() : nullptr; a.a_col = h->dA.col.as<int>(); a.a_val = h->dA.val.as<double>();
    const bool bmat = op.use_b && h->has_b;
    a.b_ptr = bmat ? h->dB.ptr.as<int>() : nullptr; a.b_col = h->dB.col.as<int>(); a.b_val = h->dB.val.as<double>();
    a.skip_b = op.use_b ? 0 : 1;
    a.za = op.za; a.zb = op.zb; a.X = X; a.Y = Y; a.aux = aux; a.lam = lam;
    a.partial = h->partial.as<zd>(); a.pstride = FC_MAXCOLS;
    spmm_dispatch<double, MODE>(h, a, grid);
  }
  check_launch(h);
  h->stats.spmm_launches++;
  sample_end(h, ev);
  {
    const double vs = h->dev_complex ? 16.0 : 8.0;
    double b = (double)h->hA.nnz * (vs + 4.0) + 4.0 * (h->hA.n + 1) + 2.0 * (double)h->hA.n * m * 16.0;
    if (h->has_b && op.use_b) b += (double)h->hB.nnz * (vs + 4.0) + 4.0 * (h->hB.n + 1);
    if (sample) h->stats.bytes_kern[FEASTCUDA_KERN_SPMM_Z] = b;
  }
  return grid;
}

static OpDesc op_shifted(zc z) { return OpDesc{mk<double>(-1.0, 0.0), tozd(z), true, true}; }
static OpDesc op_A() { return OpDesc{mk<double>(1.0, 0.0), mk<double>(0.0, 0.0), true, false}; }
static OpDesc op_B() { return OpDesc{mk<double>(0.0, 0.0), mk<double>(1.0, 0.0), false, true}; }

template <typename T>
static void reduce_to_host(H* h, const T* partial, int nslots, int nblocks, int m, T* host_out) {
  T* dev = h->red_ws.as<T>();
  const size_t sm = ((size_t)nslots * FC_MAXCOLS + 1024) * sizeof(T);
  k_reduce_partials<T><<<1, 1024, sm, h->stream>>>(partial, nslots, nblocks, FC_MAXCOLS, m, dev);
  check_launch(h);
  FC_CUDA(cudaMemcpyAsync(host_out, dev, (size_t)nslots * m * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
  sync(h);
}

static void col_norms(H* h, int m, const zd* X, std::vector<double>& out) {
  const int mp = pow2_ge(m), grid = ew_grid(h, h->ws_n, mp);
  k_colnorm2<double><<<grid, 256, 0, h->stream>>>(h->ws_n, m, mp, h->ws_ld, X, h->partial_r.as<double>(), FC_MAXCOLS);
  check_launch(h);
  out.assign(m, 0.0);
  reduce_to_host<double>(h, h->partial_r.as<double>(), 1, grid, m, out.data());
  for (auto& v : out) v = std::sqrt(v);
}

static void copy_cols(H* h, int m, const zd* src, zd* dst) {
  FC_CUDA(cudaMemcpy2DAsync(dst, (size_t)h->ws_ld * sizeof(zd), src, (size_t)h->ws_ld * sizeof(zd), (size_t)m * sizeof(zd),
                            (size_t)h->ws_n, cudaMemcpyDeviceToDevice, h->stream));
}
static void zero_cols(H* h, int m, zd* dst) {
  FC_CUDA(cudaMemset2DAsync(dst, (size_t)h->ws_ld * sizeof(zd), 0, (size_t)m * sizeof(zd), (size_t)h->ws_n, h->stream));
}

static void axpby_cols(H* h, int m, zc w, double beta, const zd* X, zd* Y) {
  const int mp = pow2_ge(m), grid = ew_grid(h, h->ws_n, mp);
  k_axpby_cols<double><<<grid, 256, 0, h->stream>>>(h->ws_n, m, mp, h->ws_ld, h->ws_ld, tozd(w), beta, X, Y);
  check_launch(h);
}

static void scale_cols(H* h, int m, const std::vector<zc>& f, const zd* X, zd* Y) {
  zd* df = h->small2.as<zd>();
  FC_CUDA(cudaMemcpyAsync(df, f.data(), (size_t)m * sizeof(zd), cudaMemcpyHostToDevice, h->stream));
  const int mp = pow2_ge(m), grid = ew_grid(h, h->ws_n, mp);
  k_scale_cols<double><<<grid, 256, 0, h->stream>>>(h->ws_n, m, mp, h->ws_ld, h->ws_ld, df, X, Y);
  check_launch(h);
  sync(h);  // f is host memory owned by the caller
}

// host (column-major) <-> device (row-major) staging; host buffers go through a pinned bounce buffer
// Row-sharded runs: the host array is the GLOBAL n_glob x m block, only this rank's rows [row0, row0 + n) cross the bus.
template <typename TIN>
static void upload_block(H* h, int64_t n, int m, const TIN* host, zd* dst) {
  const size_t bytes = (size_t)n * m * sizeof(TIN);
  h->stage.ensure(bytes);
  Timer t;
  if (h->row_sharded && n == h->n)
    FC_CUDA(cudaMemcpy2DAsync(h->stage.p, (size_t)n * sizeof(TIN), host + h->row0, (size_t)h->n_glob * sizeof(TIN), (size_t)n * sizeof(TIN),
                              (size_t)m, cudaMemcpyHostToDevice, h->stream));
  else FC_CUDA(cudaMemcpyAsync(h->stage.p, host, bytes, cudaMemcpyHostToDevice, h->stream));
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((m + 31) / 32));
  k_col2row<TIN, double><<<grid, 256, 0, h->stream>>>(n, m, n, h->ws_ld, h->stage.as<TIN>(), dst);
  check_launch(h);
  sync(h);
  h->stats.ms_h2d += t.ms();
}
template <typename TOUT>
static void download_block(H* h, int64_t n, int m, const zd* src, TOUT* host) {
  if (m <= 0) return;
  const size_t bytes = (size_t)n * m * sizeof(TOUT);
  h->stage.ensure(bytes);
  Timer t;
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((m + 31) / 32));
  k_row2col<TOUT, double><<<grid, 256, 0, h->stream>>>(n, m, h->ws_ld, n, src, h->stage.as<TOUT>());
  check_launch(h);
  if (h->row_sharded && n == h->n)   // this rank's rows of the global host block; the other rows are left untouched
    FC_CUDA(cudaMemcpy2DAsync(host + h->row0, (size_t)h->n_glob * sizeof(TOUT), h->stage.p, (size_t)n * sizeof(TOUT), (size_t)n * sizeof(TOUT),
                              (size_t)m, cudaMemcpyDeviceToHost, h->stream));
  else FC_CUDA(cudaMemcpyAsync(host, h->stage.p, bytes, cudaMemcpyDeviceToHost, h->stream));
  sync(h);
  h->stats.ms_d2h += t.ms();
}

// C (a x b, row-major host) = X[:, :a]^H Y[:, :b]
static void gram_host(H* h, int a, int b, const zd* X, const zd* Y, std::vector<zc>& C) {
  const int64_t n = h->ws_n;
  int chunks = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)h->sms * 2));
  h->gram_partial.ensure((size_t)chunks * a * b * sizeof(zd));
  dim3 grid(chunks, (a + 63) / 64, (b + 63) / 64);
  k_gram<double><<<grid, 256, 0, h->stream>>>(n, a, b, h->ws_ld, h->ws_ld, X, Y, h->gram_partial.as<zd>());
  check_launch(h);
  zd* out = h->small.as<zd>();
  const int64_t len = (int64_t)a * b;
  k_sum_chunks<zd><<<(int)std::min<int64_t>((len + 255) / 256, 1024), 256, 0, h->stream>>>(len, chunks, h->gram_partial.as<zd>(), out);
  check_launch(h);
  if (h->row_sharded) allreduce_small(h, reinterpret_cast<double*>(out), (size_t)2 * len);   // local rows -> all rows
  C.resize((size_t)len);
  FC_CUDA(cudaMemcpyAsync(C.data(), out, (size_t)len * sizeof(zd), cudaMemcpyDeviceToHost, h->stream));
  sync(h);
}

// Y[:, :b] = X[:, :a] * T  (T row-major a x b on host)
static void rowtransform(H* h, int a, int b, const zd* X, const std::vector<zc>& T, zd* Y) {
  zd* dT = h->small.as<zd>() + (size_t)4 * FC_MAXCOLS * FC_MAXCOLS;
  FC_CUDA(cudaMemcpyAsync(dT, T.data(), (size_t)a * b * sizeof(zd), cudaMemcpyHostToDevice, h->stream));
  dim3 grid((unsigned)((h->ws_n + 63) / 64), (unsigned)((b + 63) / 64));
  k_rowtransform<double><<<grid, 256, 0, h->stream>>>(h->ws_n, a, b, h->ws_ld, h->ws_ld, b, X, dT, Y);
  check_launch(h);
  sync(h);
}

// =====================================================================================================
// block BiCGStab (all m columns in lock step), see oracle/feast_port.py:block_bicgstab for the CPU port
// =====================================================================================================
struct SolveOut {
  std::vector<int> iters;
  std::vector<double> truenorm, target;
  std::vector<char> ok;
  int it_total = 0;
};

static void block_bicgstab(H* h, zc z, int m, const zd* RHS, zd* X, bool use_x0, const feastcuda_solver_opts& o, double tol,
                           SolveOut& out) {
  const int64_t n = h->ws_n;
  const int64_t ld = h->ws_ld;
  zd *r = blk(h, BS_KR), *rh = blk(h, BS_KRH), *p = blk(h, BS_KP), *v = blk(h, BS_KV), *s = blk(h, BS_KS), *t = blk(h, BS_KT);
  zd* xb = blk(h, BS_KB);  // iterate at the start of the current round (restored for diverged columns)
  const OpDesc S = op_shifted(z);
  const int mp = pow2_ge(m);
  const int egrid = ew_grid(h, n, mp);
  KrylovState<double>* dst = h->kstate.as<KrylovState<double>>();
  static thread_local KrylovState<double> hst;
  const int maxiter = std::max(1, o.maxiter);
  const int max_restarts = std::max(0, o.restart);
  const int check_every = o.check_every > 0 ? o.check_every : 8;

  std::vector<double> bn, rn(m);
  col_norms(h, m, RHS, bn);
  out.target.assign(m, 0.0);
  for (int c = 0; c < m; ++c) out.target[c] = tol + tol * bn[c];
  if (!use_x0) zero_cols(h, m, X);
  out.iters.assign(m, 0);
  out.it_total = 0;
  memset(&hst, 0, sizeof(hst));
  bool first = true;
  for (int restart = 0; restart <= max_restarts; ++restart) {
    const int g0 = launch_spmm<SPMM_RESID>(h, S, m, X, r, RHS, nullptr);
    std::vector<zd> tmp(m);
    reduce_to_host<zd>(h, h->partial.as<zd>(), 1, g0, m, tmp.data());
    for (int c = 0; c < m; ++c) rn[c] = std::sqrt(tmp[c].x);
    if (first) {
      if (o.inner_rel > 0)
        for (int c = 0; c < m; ++c) out.target[c] = std::max(out.target[c], o.inner_rel * rn[c]);
      first = false;
    }
    int nact = 0;
    for (int c = 0; c < m; ++c) {
      hst.active[c] = rn[c] > out.target[c] ? 1 : 0;
      nact += hst.active[c];
      hst.target[c] = out.target[c];
      hst.rnorm[c] = rn[c];
      hst.rn0[c] = rn[c];
      hst.diverged[c] = 0;
      hst.rho[c] = mk<double>(rn[c] * rn[c], 0.0);
      hst.alpha[c] = hst.omega[c] = hst.beta[c] = mk<double>(0.0, 0.0);
      hst.flags[c] = 0;
      hst.iters[c] = out.iters[c];
    }
    hst.n_active = nact;
    if (nact == 0 || out.it_total >= maxiter) break;
    FC_CUDA(cudaMemcpyAsync(dst, &hst, sizeof(hst), cudaMemcpyHostToDevice, h->stream));
    copy_cols(h, m, r, rh);
    copy_cols(h, m, r, p);
    copy_cols(h, m, X, xb);
    int* flag = reinterpret_cast<int*>(pinned_buf(h, 64));
    while (out.it_total < maxiter) {
      const bool sample = (out.it_total % 16) == 0;
      const int g1 = launch_spmm<SPMM_DOT_RHAT>(h, S, m, p, v, rh, nullptr, sample);
      k_bicg_scal1<double><<<1, 1024, 0, h->stream>>>(dst, h->partial.as<zd>(), g1, FC_MAXCOLS, m);
      check_launch(h);
      k_upd_s<double><<<egrid, 256, 0, h->stream>>>(n, m, mp, ld, dst, r, v, s);
      check_launch(h);
      const int g2 = launch_spmm<SPMM_DOT_TS>(h, S, m, s, t, rh, nullptr);
      k_bicg_scal2<double><<<1, 1024, 0, h->stream>>>(dst, h->partial.as<zd>(), g2, FC_MAXCOLS, m);
      check_launch(h);
      k_upd_xrp<double><<<egrid, 256, 0, h->stream>>>(n, m, mp, ld, dst, X, r, p, v, s, t, h->partial_r.as<double>(), FC_MAXCOLS);
      check_launch(h);
      k_bicg_scal3<double><<<1, 1024, 0, h->stream>>>(dst, h->partial_r.as<double>(), egrid, FC_MAXCOLS, m);
      check_launch(h);
      out.it_total++;
      h->stats.krylov_iters++;
      if (out.it_total % check_every == 0 || out.it_total >= maxiter) {
        FC_CUDA(cudaMemcpyAsync(flag, &dst->n_active, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        sync(h);
        if (*flag == 0) break;
      }
    }
    FC_CUDA(cudaMemcpyAsync(&hst, dst, sizeof(hst), cudaMemcpyDeviceToHost, h->stream));
    sync(h);
    bool any_div = false;
    for (int c = 0; c < m; ++c) { out.iters[c] = hst.iters[c]; any_div = any_div || hst.diverged[c]; }
    if (any_div) {
      k_restore_cols<double><<<egrid, 256, 0, h->stream>>>(n, m, mp, ld, dst->diverged, xb, X);
      check_launch(h);
    }
  }
  // true residual of the returned block (the reference's explicit gate, sparse/feast_sparse.jl:189-198)
  {
    const int g0 = launch_spmm<SPMM_RESID>(h, S, m, X, r, RHS, nullptr);
    std::vector<zd> tmp(m);
    reduce_to_host<zd>(h, h->partial.as<zd>(), 1, g0, m, tmp.data());
    out.truenorm.assign(m, 0.0);
    out.ok.assign(m, 0);
    for (int c = 0; c < m; ++c) {
      out.truenorm[c] = std::sqrt(tmp[c].x);
      out.ok[c] = out.truenorm[c] <= 10.0 * out.target[c];
    }
  }
  int64_t ci = 0;
  for (int c = 0; c < m; ++c) ci += out.iters[c];
  h->stats.col_iters += ci;
  drain_events(h);
}

// =====================================================================================================
// multi-shift two-pass Lanczos filter (kernels_lanczos.cuh; CPU port: oracle/feast_port.py:feast_hrr_mslanczos)
// Computes, for the real basis columns [c0, c0+nc) of slot `basis_slot`, the filtered block
//     Qacc = sum_e Re(2 w_e (z_e I - A)^-1 q)           (sparse/feast_sparse.jl:318-370 with B = I)
// into the same columns of BS_ACC (complex storage, zero imaginary part).  With Ritz values theta (previous
// loop) the solves start from x0 = q/(z_e - theta):  x_e = q/(z_e-theta) + (z_e I - A)^-1 (A q - theta q)/(z_e-theta),
// every residual system shares the start vector A q - theta q, so one recurrence still serves all nodes.
// =====================================================================================================
struct MslOut { int k = 0; double maxres = 0.0; bool converged = false; };

static inline double* rblk(H* h, int slot) { return reinterpret_cast<double*>(blk(h, slot)); }

// rows per round-robin tile: the configured size (64), halved while the static deal of the tiles leaves the busiest CTA with more than
// 3 % extra work (few rows per rank in row-sharded runs: 125 000 rows over 296 CTAs are 6.6 tiles of 64 rows each -> 7 for some)
static int lz_pick_tile(H* h, int64_t n, int rows_per_step, int ctas_per_sm = 0) {
  const int64_t ctas = (int64_t)h->sms * (ctas_per_sm > 0 ? ctas_per_sm : h->lz_ctas_per_sm);
  int t = h->lz_tile_rows;
  for (;;) {
    const int64_t tr = (int64_t)std::max(1, t / rows_per_step) * rows_per_step;
    const int64_t nt = (n + tr - 1) / tr;
    const double per = (double)nt / (double)ctas;
    if (per <= 1.0 || std::ceil(per) / per <= 1.03 || tr <= rows_per_step) return (int)tr;
    t = (int)tr / 2;
  }
}

static int lz_grid_spmm(H* h, int64_t n, int rows_per_step, int ctas_per_sm = 0, int tile_rows = 0) {
  const int64_t tr = (int64_t)std::max(1, (tile_rows > 0 ? tile_rows : h->lz_tile_rows) / rows_per_step) * rows_per_step;   // as in k_lz_spmm
  const int64_t ntiles = (n + tr - 1) / tr;
  return (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)h->sms * (ctas_per_sm > 0 ? ctas_per_sm : h->lz_ctas_per_sm)));
}

// Row-sharded runs: the order in which k_lz_spmm deals its row tiles.  Tiles holding a row with a stored entry in a peer's block (halo
// gathers over NVLink; for a symmetric pattern these are also the rows the peers read) come in the second half of the sweep.
static const int* shard_tile_order(H* h, int rows_per_step, int tile_rows) {
  const int tr = std::max(1, tile_rows / rows_per_step) * rows_per_step;
  if (h->tile_order_tr == tr && h->tile_order.p) return h->tile_order.as<int>();
  const int64_t n = h->n;
  const int ntiles = (int)((n + tr - 1) / tr);
  std::vector<int> inner, halo, order;
  const HostCsr& A = h->hA;
  const unsigned self = (unsigned)h->rank;
  // owner of a global column: the block rule of row_block
  const int64_t ng = h->n_glob, base = ng / h->nranks, rem = ng % h->nranks, bound = rem * (base + 1);
  auto owner_of = [&](int64_t c) -> unsigned { return (unsigned)(c < bound ? c / (base + 1) : rem + (base > 0 ? (c - bound) / base : 0)); };
  for (int t = 0; t < ntiles; ++t) {
    bool remote = false;
    const int64_t r0 = h->row0 + (int64_t)t * tr, r1 = std::min<int64_t>(h->row0 + n, r0 + tr);
    for (int64_t r = r0; r < r1 && !remote; ++r)
      for (int p = A.ptr[r]; p < A.ptr[r + 1]; ++p)
        if (owner_of(A.col[p]) != self) { remote = true; break; }
    (remote ? halo : inner).push_back(t);
  }
  // interior tiles in their natural order; the halo tiles are spread evenly over the second half of the sweep: late enough that the
  // peers' completion signals have long arrived when the first of them is reached (lazy wait of pass 2), and interleaved with interior
  // tiles so that the NVLink latency of their gathers hides behind local rows instead of stalling every CTA at the end
  order.reserve(ntiles);
  const size_t first_half = std::min(inner.size(), (size_t)std::max(0, ntiles / 2));
  order.insert(order.end(), inner.begin(), inner.begin() + first_half);
  h->halo_start = (int)order.size();
  {
    const size_t rest = (size_t)ntiles - order.size();
    size_t ih = 0, ii = first_half;
    for (size_t v = 0; v < rest; ++v) {
      const bool want_halo = ih < halo.size() && ((double)(ih + 1) * rest <= (double)(v + 1) * halo.size() || ii >= inner.size());
      order.push_back(want_halo ? halo[ih++] : inner[ii++]);
    }
  }
  h->tile_order.ensure((size_t)std::max(1, ntiles) * sizeof(int));
  FC_CUDA(cudaMemcpyAsync(h->tile_order.p, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  sync(h);
  h->tile_order_tr = tr;
  return h->tile_order.as<int>();
}

template <int MODE, bool CPLX>
static void lz_launch(H* h, LzArgs& a, int* grid_out) {
  const int P = CPLX ? a.m : (a.m + 1) / 2;
  // G lanes per row, NC column-pair chunks per lane
#define FC_LZ(G, NC)                                                                       \
  do {                                                                                     \
    if (NC == 1 && h->lz_threads >= 1024 && a.goff == nullptr) {                                                \
      if constexpr (NC == 1) {                                                             \
        const int grid = lz_grid_spmm(h, a.n, 32 * (32 / G));                              \
        *grid_out = grid;                                                                  \
        k_lz_spmm<G, NC, MODE, 1024, CPLX><<<grid, 1024, 0, h->stream>>>(a);               \
      }                                                                                    \
    } else {                                                                               \
      a.tile_rows = lz_pick_tile(h, a.n, 16 * (32 / G), NC >= 3 ? 1 : 0);                  \
      const int grid = lz_grid_spmm(h, a.n, 16 * (32 / G), NC >= 3 ? 1 : 0, a.tile_rows);  \
      *grid_out = grid;                                                                    \
      if (a.goff != nullptr) {                                                             \
        if constexpr (!CPLX && NC <= 2 && MODE <= LZ_P2_PAIR) {                            \
          a.tile_order = shard_tile_order(h, 16 * (32 / G), a.tile_rows);                               \
          a.halo_start = h->halo_start;                                                    \
          a.nranks = h->nranks; a.rank = h->rank;                                          \
          a.kdone = reinterpret_cast<const LzMailbox*>((const char*)h->arena + h->arena_mbox_off)->kdone; \
          k_lz_spmm<G, NC, MODE, 512, CPLX, true><<<grid, 512, 0, h->stream>>>(a);         \
        } else throw FcError(FEASTCUDA_ERR_UNSUPPORTED, "row-sharded gather: real blocks only"); \
      } else k_lz_spmm<G, NC, MODE, 512, CPLX><<<grid, 512, 0, h->stream>>>(a);            \
    }                                                                                      \
  } while (0)
  const int Pd = P;   // lanes per row = elements per row (wider groups for narrow blocks measured slower: idle lanes still issue)
  if (Pd <= 1) FC_LZ(1, 1);
  else if (Pd <= 2) FC_LZ(2, 1);
  else if (Pd <= 4) FC_LZ(4, 1);
  else if (Pd <= 8) FC_LZ(8, 1);
  else if (Pd <= 16) FC_LZ(16, 1);
  else if (Pd <= 32) FC_LZ(32, 1);
  else if (Pd <= 64) FC_LZ(32, 2);
  else if constexpr (CPLX) {   // complex blocks wider than 64 columns (real blocks: 128 columns = 64 pairs)
    if (Pd <= 96) FC_LZ(32, 3);
    else FC_LZ(32, 4);
  } else throw FcError(FEASTCUDA_ERR_ARG, "Lanczos SpMM: more than 128 real columns");
#undef FC_LZ
  check_launch(h);
  h->stats.spmm_launches++;
}

// Wide complex blocks (more than `slice` columns) in column slices: a 96-column complex row is 1.5 KB, too wide for the L1-resident
// neighbour reuse and the register budget of the gather kernel; every slice re-reads the matrix (one launch per slice).
template <int MODE, bool CPLX>
static void lz_launch_sliced(H* h, LzArgs& a, int* grid_out, int slice) {
  if (!CPLX || slice <= 0 || a.m <= slice) { lz_launch<MODE, CPLX>(h, a, grid_out); return; }
  const int m = a.m;
  const int nsl = (m + slice - 1) / slice, w = (m + nsl - 1) / nsl;   // equal slices
  for (int c0 = 0; c0 < m; c0 += w) {
    LzArgs b = a;
    b.m = std::min(w, m - c0);
    const int64_t o = 2 * (int64_t)c0;     // doubles per complex column
    b.U = a.U + o;
    if (a.prev) b.prev = a.prev + o;
    if (a.out) b.out = a.out + o;
    if (a.Q) b.Q = a.Q + o;
    if (a.own) b.own = a.own + o;
    if (a.rhs) b.rhs = a.rhs + o;
    if (a.s_inv_beta) b.s_inv_beta = a.s_inv_beta + c0;
    if (a.s_ratio_b) b.s_ratio_b = a.s_ratio_b + c0;
    if (a.s_ratio_a) b.s_ratio_a = a.s_ratio_a + c0;
    if (a.s_coef) b.s_coef = a.s_coef + c0;
    if (a.s_coef_prev) b.s_coef_prev = a.s_coef_prev + c0;
    if (a.s_theta) b.s_theta = a.s_theta + c0;
    if (a.partial) b.partial = a.partial + c0;
    lz_launch<MODE, CPLX>(h, b, grid_out);
  }
}

template <int MODE>
static void lz32_launch(H* h, LzArgs32& a, int* grid_out) {
  const int P = (a.m + 3) / 4;   // 4-column elements per row = lanes per row
  // shape: 2 CTAs x 512 threads, 4 gathers in flight per lane (64 registers).  Measured on B200 and rejected (n = 1e6, 64 columns,
  // pass 1 / pass 2): 3 x 256 threads with 8 gathers in flight (80 registers) 0.264 / 0.354 ms and 2 x 384 / 8: 0.270 / 0.356 ms
  // against 0.242 / 0.344 ms -- resident warps matter more than loads in flight per warp.
#define FC_LZ32(G)                                                                 \
  do {                                                                             \
    a.tile_rows = lz_pick_tile(h, a.n, 16 * (32 / G));                             \
    const int grid = lz_grid_spmm(h, a.n, 16 * (32 / G), 0, a.tile_rows);          \
    *grid_out = grid;                                                              \
    if (a.goff != nullptr) {                                                       \
      a.tile_order = shard_tile_order(h, 16 * (32 / G), a.tile_rows);              \
      a.halo_start = h->halo_start;                                                \
      a.nranks = h->nranks; a.rank = h->rank;                                      \
      a.kdone = reinterpret_cast<const LzMailbox*>((const char*)h->arena + h->arena_mbox_off)->kdone; \
      k_lz32_spmm<G, MODE, 512, 2, 4, true><<<grid, 512, 0, h->stream>>>(a);       \
    } else k_lz32_spmm<G, MODE, 512><<<grid, 512, 0, h->stream>>>(a);              \
  } while (0)
  if (P <= 1) FC_LZ32(1);
  else if (P <= 2) FC_LZ32(2);
  else if (P <= 4) FC_LZ32(4);
  else if (P <= 8) FC_LZ32(8);
  else if (P <= 16) FC_LZ32(16);
  else FC_LZ32(32);
#undef FC_LZ32
  check_launch(h);
  h->stats.spmm_launches++;
}

// FP32 vectors: nvec32 float blocks + nvec64 double blocks per launch
static double lz32_bytes_spmm(H* h, int m, int nvec32, int nvec64) {
  return (double)h->nnz_loc * 8.0 + 4.0 * (double)(h->n + 1) + (double)h->n * m * (4.0 * nvec32 + 8.0 * nvec64);
}

static double lz_bytes_spmm(H* h, int m, int nvec, bool cplx) {
  const double es = cplx ? 16.0 : 8.0;
  return (double)h->nnz_loc * (es + 4.0) + 4.0 * (double)(h->n + 1) + (double)nvec * (double)h->n * m * es;
}

// c_j = ||b|| sum_e Re(2 w_e F_e [(z_e I - T_k)^-1 e_1]_j) / beta_j for the unnormalised Lanczos vectors of every column: complex Thomas
// solves of the shifted tridiagonals (the pivots have Im d_j >= Im z_e > 0: no pivoting needed); F_e = 1/(z_e - theta) with a Ritz start
static void lz_host_coefficients(int k, int nc, const std::vector<double>& al, const std::vector<double>& be, bool have_ritz,
                                 const double* theta, const zc* Zne, const zc* Wne, int ne, std::vector<double>& coef) {
  const size_t rowsz = (size_t)FC_MAXCOLS;
  {
    std::vector<zc> dd(k), ff(k);
    for (int c = 0; c < nc; ++c) {
      const double b0 = be[c];
      if (!(b0 > 0.0)) continue;
      int kc = k;  // a frozen column's T ends where beta vanished
      for (int j = 1; j < k; ++j)
        if (be[(size_t)j * rowsz + c] == 0.0) { kc = j; break; }
      for (int e = 0; e < ne; ++e) {
        const zc z = Zne[e];
        const zc F = have_ritz ? zc(1.0) / (z - theta[c]) : zc(1.0);
        const zc wf = 2.0 * Wne[e] * F * b0;
        dd[0] = z - al[c];
        ff[0] = 1.0;
        for (int j = 1; j < kc; ++j) {
          const double bj = be[(size_t)j * rowsz + c];
          const zc w = -bj / dd[j - 1];
          dd[j] = (z - al[(size_t)j * rowsz + c]) + w * bj;
          ff[j] = -w * ff[j - 1];
        }
        zc y = ff[kc - 1] / dd[kc - 1];
        coef[(size_t)(kc - 1) * rowsz + c] += (wf * y).real();
        for (int j = kc - 2; j >= 0; --j) {
          y = (ff[j] + be[(size_t)(j + 1) * rowsz + c] * y) / dd[j];
          coef[(size_t)j * rowsz + c] += (wf * y).real();
        }
      }
      for (int j = 0; j < kc; ++j) coef[(size_t)j * rowsz + c] /= be[(size_t)j * rowsz + c];
    }
  }
}

template <bool CPLX>
static void msl_filter(H* h, int basis_slot, int c0, int nc, bool have_ritz, const double* theta, const zc* Zne, const zc* Wne,
                       int ne, double target, int kmax, int check_every, MslOut& out, bool mixed = false) {
  const bool matfree = h->kind == OP_MATFREE;
  FC_REQUIRE(((h->kind == OP_SPARSE && h->dev_complex == CPLX) || (matfree && !CPLX)) && !h->has_b,
             "multi-shift Lanczos needs a standard sparse Hermitian (or matrix-free real symmetric) problem");
  FC_REQUIRE(CPLX || (c0 & 1) == 0, "column slices must start at an even column");
  const int64_t n = h->ws_n;
  // the work blocks are COMPACT: row stride = the slice's own column count (in doubles: even(nc) real, 2 nc complex), so a
  // rank that owns 8 of 64 columns streams dense 64-byte rows instead of touching 64 bytes out of every 512
  // mixed precision (FP32 Lanczos vectors, real problems): 4-column elements, so every compact block is padded to a multiple of 4
  mixed = mixed && !CPLX && !matfree && (int64_t)((nc + 3) & ~3) <= 2 * h->ws_ld;   // the padded FP64 blocks must fit their slots
  const int64_t ldz = h->ws_ld, ld = CPLX ? 2 * (int64_t)nc : (mixed ? ((nc + 3) & ~3) : ((nc + 1) & ~1));
  FC_REQUIRE((double)n * (double)ld < 4294967296.0, "multi-shift Lanczos: n*ld must be below 2^32 (32-bit gather offsets)");
  kmax = std::max(1, std::min(kmax, 16384));
  check_every = std::max(1, check_every);
  const int P = CPLX ? nc : (nc + 1) / 2;
  const int pp = pow2_ge(P);
  const int egrid = (int)std::max<int64_t>(1, std::min<int64_t>((n + (256 / pp) - 1) / (256 / pp), (int64_t)h->sms * h->lz_egrid_mult));
  const int pp4 = pow2_ge((nc + 3) / 4);
  const int egrid4 = (int)std::max<int64_t>(1, std::min<int64_t>((n + (256 / pp4) - 1) / (256 / pp4), (int64_t)h->sms * h->lz_egrid_mult));

  // ---- device scalars ------------------------------------------------------------------------------------------
  const size_t rowsz = (size_t)FC_MAXCOLS;
  const size_t per = (size_t)(kmax + 2) * rowsz;
  const size_t scal_doubles = 5 * per + rowsz /*scale*/ + 2 * rowsz /*theta, rho*/ + (size_t)(kmax + 2) /*maxres*/;
  h->lz_scal.ensure(scal_doubles * sizeof(double));
  h->lz_coef.ensure(per * sizeof(double));
  h->lz_state.ensure(((size_t)2 * ne * rowsz + ne) * sizeof(zd) + 64);
  double* base = h->lz_scal.as<double>();
  FC_CUDA(cudaMemsetAsync(base, 0, scal_doubles * sizeof(double), h->stream));
  LzScalars S;
  S.alpha = base; S.beta = base + per; S.inv_beta = base + 2 * per; S.ratio_b = base + 3 * per; S.ratio_a = base + 4 * per;
  S.scale = base + 5 * per;
  double* d_theta = S.scale + rowsz;
  double* d_rho = d_theta + rowsz;
  S.maxres = d_rho + rowsz;
  zd* stz = h->lz_state.as<zd>();
  S.d = stz; S.g = stz + (size_t)ne * rowsz;
  zd* d_z = stz + (size_t)2 * ne * rowsz;
  S.z = d_z;
  S.ne = ne;
  S.target = target;
  S.done_k = reinterpret_cast<int*>(d_z + ne);
  FC_CUDA(cudaMemsetAsync(S.done_k, 0, 64, h->stream));
  if (h->lz_ticket.cap == 0) {      // the tails leave their counters at zero
    h->lz_ticket.ensure(256 * sizeof(int));
    FC_CUDA(cudaMemsetAsync(h->lz_ticket.p, 0, 256 * sizeof(int), h->stream));
  }
  h->lz_grows.ensure((size_t)128 * FC_MAXCOLS * sizeof(double));
  int* ticket = h->lz_ticket.as<int>();
  double* grows = h->lz_grows.as<double>();
  FC_CUDA(cudaMemcpyAsync(d_z, Zne, (size_t)ne * sizeof(zd), cudaMemcpyHostToDevice, h->stream));
  // what the last CTA of a producer kernel does with the partial sums (pass 1) / the step barrier of row-sharded runs (pass 2)
  const bool sharded = h->row_sharded;
  auto mk_tail = [&](int kind, int j) {
    LzTail t;
    memset(&t, 0, sizeof(t));
    if (matfree || kind == LZ_TAIL_NONE) return t;
    t.kind = kind; t.j = j; t.ticket = ticket; t.grows = grows; t.s = S;
    fill_xchg(h, t.x);
    return t;
  };
  const int bar_kind = sharded ? LZ_TAIL_BARRIER : LZ_TAIL_NONE;
  xbarrier(h);   // row-sharded: no peer is still reading the blocks this sweep is about to overwrite

  // ---- real work blocks (each aliases a complex slot; n x ld doubles) -------------------------------------------
  double* RQ = rblk(h, BS_KS);
  double* RB = rblk(h, BS_KR);
  double* UA = rblk(h, BS_KRH);
  double* UB = rblk(h, BS_KP);
  double* QA = rblk(h, BS_KV);
  double* MW = rblk(h, BS_KT);   // matrix-free: W = A U from the caller's callback
  auto mf_apply = [&](const double* X, double* Y) {
    h->mf_apply(h->mf_ctx, n, (int64_t)nc, X, ld, Y, ld, (void*)h->stream);
    h->stats.kernel_launches++;
    h->stats.spmm_launches++;
  };
  if (matfree) FC_CUDA(cudaMemsetAsync(MW, 0, (size_t)n * (size_t)ld * sizeof(double), h->stream));   // pad column of an odd slice
  const zd* basis = blk(h, basis_slot) + c0;
  double* part = h->partial_r.as<double>();

  std::vector<double> rho(nc, 0.0);
  if (mixed) FC_CUDA(cudaMemsetAsync(QA, 0, (size_t)n * (size_t)ld * sizeof(double), h->stream));   // pad columns included
  if (!have_ritz) {
    k_lz_real_part<CPLX><<<egrid, 256, 0, h->stream>>>(n, nc, pp, ldz, ld, basis, RB, part, FC_MAXCOLS, mk_tail(LZ_TAIL_INIT, 0));
    check_launch(h);
    if (matfree) {
      k_lz_scal_init<<<1, 1024, 0, h->stream>>>(S, part, egrid, FC_MAXCOLS, nc);
      check_launch(h);
    }
    if (!mixed) FC_CUDA(cudaMemset2DAsync(QA, (size_t)ld * sizeof(double), 0, (size_t)(2 * P) * sizeof(double), (size_t)n, h->stream));
  } else {
    // rho(theta) = Re sum_e 2 w_e / (z_e - theta): what the rational filter does to an exact eigenvector
    for (int c = 0; c < nc; ++c) {
      zc acc(0.0);
      for (int e = 0; e < ne; ++e) acc += 2.0 * Wne[e] / (Zne[e] - theta[c]);
      rho[c] = acc.real();
    }
    FC_CUDA(cudaMemcpyAsync(d_theta, theta, (size_t)nc * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    FC_CUDA(cudaMemcpyAsync(d_rho, rho.data(), (size_t)nc * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    k_lz_real_part<CPLX><<<egrid, 256, 0, h->stream>>>(n, nc, pp, ldz, ld, basis, RQ, nullptr, FC_MAXCOLS, mk_tail(bar_kind, 0));
    check_launch(h);
    LzArgs a;
    memset(&a, 0, sizeof(a));
    a.n = n; a.m = nc; a.ld = ld;
    a.ptr = h->dA.ptr.as<int>(); a.col = h->dA.col.as<int>(); a.val = h->dA.val.p;
    a.U = RQ; a.prev = RQ; a.out = RB; a.Q = QA; a.s_coef = d_rho; a.s_theta = d_theta;
    a.partial = part; a.pstride = FC_MAXCOLS; a.tile_rows = h->lz_tile_rows;
    a.tail = mk_tail(LZ_TAIL_INIT, 0);
    if (sharded) a.goff = resolve_goff(h, ld * (int64_t)sizeof(double));
    int g = 0;
    const int ev = sample_begin(h, FEASTCUDA_KERN_LZ_RES);
    if (matfree) {
      mf_apply(RQ, MW);
      k_mf_step<LZ_RES><<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, MW, RQ, nullptr, RB, QA, d_theta, nullptr, nullptr, d_rho, part, FC_MAXCOLS,
                                                      nullptr);
      check_launch(h);
      g = egrid;
    } else lz_launch<LZ_RES, CPLX>(h, a, &g);
    sample_end(h, ev);
    h->stats.bytes_kern[FEASTCUDA_KERN_LZ_RES] = matfree ? 4.0 * 8.0 * (double)n * nc : lz_bytes_spmm(h, nc, 3, CPLX);
    if (matfree) {
      k_lz_scal_init<<<1, 1024, 0, h->stream>>>(S, part, g, FC_MAXCOLS, nc);
      check_launch(h);
    }
    sync(h);  // theta / rho are host buffers
  }

  auto cur = [&](int j) -> double* { return j == 0 ? RB : ((j & 1) ? UA : UB); };
  // mixed precision: the FP64 start block (basis or Ritz residual) is rounded once; the three FP32 vectors alias FP64 slots
  float* F0 = reinterpret_cast<float*>(RQ);
  auto cur32 = [&](int j) -> float* { return j == 0 ? F0 : reinterpret_cast<float*>((j & 1) ? UA : UB); };
  if (mixed) {
    if (!h->dA.val32_ready) {
      h->dA.val32.ensure((size_t)std::max<int64_t>(h->nnz_loc, 1) * sizeof(float));
      k_lz32_vals<<<std::max(1, h->sms * 4), 256, 0, h->stream>>>(h->nnz_loc, h->dA.val.as<double>(), h->dA.val32.as<float>());
      check_launch(h);
      h->dA.val32_ready = true;
    }
    k_lz32_narrow<<<egrid4, 256, 0, h->stream>>>(n, nc, pp4, ld, RB, F0);
    check_launch(h);
    xbarrier(h);   // row-sharded: step 0 gathers the peers' rows of the rounded start block
  }
  auto args32 = [&](int j) {
    LzArgs32 a;
    memset(&a, 0, sizeof(a));
    a.n = n; a.m = nc; a.ld = ld;
    a.ptr = h->dA.ptr.as<int>(); a.col = h->dA.col.as<int>(); a.val = h->dA.val32.as<float>();
    a.U = cur32(j); a.prev = j > 0 ? cur32(j - 1) : cur32(j); a.out = cur32(j + 1);
    a.s_inv_beta = S.inv_beta + (size_t)j * rowsz; a.s_ratio_b = S.ratio_b + (size_t)j * rowsz;
    a.tile_rows = h->lz_tile_rows;
    if (sharded) a.goff = resolve_goff(h, ld * (int64_t)sizeof(float));     // the FLOAT blocks' row stride
    return a;
  };

  // ---- pass 1: build T_k, device-side convergence flag ----------------------------------------------------------
  Timer t1;
  int* flag = reinterpret_cast<int*>(pinned_buf(h, 64));
  flag[0] = 0;
  int done = 0, kfinal = 0;
  while (done < kmax) {
    const int batch = std::min(check_every, kmax - done);
    for (int j = done; j < done + batch; ++j) {
      const bool smp = (j % 16) == 3;
      int g = 0;
      int ev = smp ? sample_begin(h, FEASTCUDA_KERN_LZ_P1, j) : -1;
      if (matfree) {
        mf_apply(cur(j), MW);
        k_mf_step<LZ_P1><<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, MW, cur(j), j > 0 ? cur(j - 1) : cur(j), cur(j + 1), nullptr,
                                                       S.inv_beta + (size_t)j * rowsz, S.ratio_b + (size_t)j * rowsz, nullptr, nullptr, part,
                                                       FC_MAXCOLS, S.done_k);
        check_launch(h);
        g = egrid;
      } else if (mixed) {
        LzArgs32 a = args32(j);
        a.partial = part; a.pstride = FC_MAXCOLS; a.done = S.done_k;
        a.tail = mk_tail(LZ_TAIL_ALPHA, j);
        lz32_launch<LZ_P1>(h, a, &g);
      } else {
        LzArgs a;
        memset(&a, 0, sizeof(a));
        a.n = n; a.m = nc; a.ld = ld;
        a.ptr = h->dA.ptr.as<int>(); a.col = h->dA.col.as<int>(); a.val = h->dA.val.p;
        a.U = cur(j); a.prev = j > 0 ? cur(j - 1) : cur(j); a.out = cur(j + 1);
        a.s_inv_beta = S.inv_beta + (size_t)j * rowsz; a.s_ratio_b = S.ratio_b + (size_t)j * rowsz;
        a.partial = part; a.pstride = FC_MAXCOLS; a.tile_rows = h->lz_tile_rows; a.done = S.done_k;
        a.tail = mk_tail(LZ_TAIL_ALPHA, j);
        if (sharded) a.goff = resolve_goff(h, ld * (int64_t)sizeof(double));
        lz_launch<LZ_P1, CPLX>(h, a, &g);
      }
      sample_end(h, ev);
      if (matfree) {
        k_lz_scal1<<<1, 1024, 0, h->stream>>>(S, j, part, g, FC_MAXCOLS, nc);
        check_launch(h);
      }
      ev = smp ? sample_begin(h, FEASTCUDA_KERN_LZ_UPD, j) : -1;
      if (mixed)
        k_lz32_update<<<egrid4, 256, 0, h->stream>>>(n, nc, pp4, ld, S.ratio_a + (size_t)j * rowsz, cur32(j), cur32(j + 1), part, FC_MAXCOLS,
                                                     S.done_k, mk_tail(LZ_TAIL_BETA, j));
      else
        k_lz_update<CPLX><<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, S.ratio_a + (size_t)j * rowsz, cur(j), cur(j + 1), part, FC_MAXCOLS,
                                                   S.done_k, mk_tail(LZ_TAIL_BETA, j));
      check_launch(h);
      sample_end(h, ev);
      if (matfree) {
        k_lz_scal2<<<1, 1024, 0, h->stream>>>(S, j, part, egrid, FC_MAXCOLS, nc);
        check_launch(h);
      }
    }
    done += batch;
    FC_CUDA(cudaMemcpyAsync(flag, S.done_k, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    sync(h);
    if (flag[0] != 0) { kfinal = flag[0]; out.converged = true; break; }
  }
  if (kfinal == 0) kfinal = done;
  h->stats.bytes_kern[FEASTCUDA_KERN_LZ_P1] = matfree ? 4.0 * 8.0 * (double)n * nc : (mixed ? lz32_bytes_spmm(h, nc, 3, 0) : lz_bytes_spmm(h, nc, 3, CPLX));
  h->stats.bytes_kern[FEASTCUDA_KERN_LZ_UPD] = 3.0 * (double)n * nc * (mixed ? 4.0 : (CPLX ? 16.0 : 8.0));
  drain_events(h, kfinal);
  h->stats.lz_steps_p1 += kfinal;
  if (mixed) h->stats.lz_steps_fp32 += kfinal;
  h->stats.krylov_iters += kfinal;
  h->stats.col_iters += (int64_t)kfinal * nc;
  h->stats.ms_lz_p1 += t1.ms();

  // ---- coefficients c_j = ||b|| sum_e Re(2 w_e F_e [(z_e I - T_k)^-1 e_1]_j), divided by beta_j (unnormalised vectors) ---
  const int k = kfinal;
  std::vector<double> al((size_t)k * rowsz), be((size_t)(k + 1) * rowsz), coef((size_t)k * rowsz, 0.0);
  FC_CUDA(cudaMemcpyAsync(al.data(), S.alpha, al.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  FC_CUDA(cudaMemcpyAsync(be.data(), S.beta, be.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  double mr = 0.0;
  FC_CUDA(cudaMemcpyAsync(&mr, S.maxres + (k - 1), sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  sync(h);
  out.k = k;
  out.maxres = mr;
  lz_host_coefficients(k, nc, al, be, have_ritz, theta, Zne, Wne, ne, coef);
  double* d_coef = h->lz_coef.as<double>();
  FC_CUDA(cudaMemcpyAsync(d_coef, coef.data(), coef.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));

  // ---- pass 2: same recurrence from the stored scalars, accumulation fused ---------------------------------------
  // Q is read-modify-written every SECOND step: step j holds u_j and u_{j-1} as own-row operands, so odd steps add both
  // contributions and even steps carry no Q traffic at all (the staged and matrix-free variants accumulate every step)
  Timer t2;
  const bool paired = h->lz_paired && !matfree;
  for (int j = 0; j < k; ++j) {
    if (j == k - 1) {
      for (int jj = (paired && j >= 1 && ((j - 1) & 1) == 0) ? j - 1 : j; jj <= j; ++jj) {   // a skipped even step k-2 is added here
        if (mixed) k_lz32_axpy<<<egrid4, 256, 0, h->stream>>>(n, nc, pp4, ld, d_coef + (size_t)jj * rowsz, cur32(jj), QA);
        else k_lz_axpy<CPLX><<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, d_coef + (size_t)jj * rowsz, cur(jj), QA);
        check_launch(h);
      }
      break;
    }
    const int q_mode = paired ? ((j & 1) ? 2 : 1) : 0;
    const bool smp = (j % 16) == 3 || (j % 16) == 10;   // one heavy (odd) and one light (even) step per 16
    int g = 0;
    const int ev = smp ? sample_begin(h, FEASTCUDA_KERN_LZ_P2, j) : -1;
    if (matfree) {
      mf_apply(cur(j), MW);
      k_mf_step<LZ_P2><<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, MW, cur(j), j > 0 ? cur(j - 1) : cur(j), cur(j + 1), QA,
                                                     S.inv_beta + (size_t)j * rowsz, S.ratio_b + (size_t)j * rowsz,
                                                     S.ratio_a + (size_t)j * rowsz, d_coef + (size_t)j * rowsz, nullptr, 0, nullptr);
      check_launch(h);
    } else if (mixed) {
      LzArgs32 a = args32(j);
      a.Q = QA; a.s_ratio_a = S.ratio_a + (size_t)j * rowsz; a.s_coef = d_coef + (size_t)j * rowsz;
      a.s_coef_prev = j > 0 ? d_coef + (size_t)(j - 1) * rowsz : nullptr;
      if (sharded) {
        a.wait_seq = (h->nranks > 1 && j > 0) ? h->xseq : 0;     // the signal of the previous pass-2 kernel
        a.tail = mk_tail(LZ_TAIL_SIGNAL, j);
      }
      if (q_mode == 0) lz32_launch<LZ_P2>(h, a, &g);
      else if (q_mode == 1) lz32_launch<LZ_P2_SKIP>(h, a, &g);
      else lz32_launch<LZ_P2_PAIR>(h, a, &g);
    } else {
      LzArgs a;
      memset(&a, 0, sizeof(a));
      a.n = n; a.m = nc; a.ld = ld;
      a.ptr = h->dA.ptr.as<int>(); a.col = h->dA.col.as<int>(); a.val = h->dA.val.p;
      a.U = cur(j); a.prev = j > 0 ? cur(j - 1) : cur(j); a.out = cur(j + 1); a.Q = QA;
      a.s_inv_beta = S.inv_beta + (size_t)j * rowsz; a.s_ratio_b = S.ratio_b + (size_t)j * rowsz;
      a.s_ratio_a = S.ratio_a + (size_t)j * rowsz; a.s_coef = d_coef + (size_t)j * rowsz;
      a.s_coef_prev = j > 0 ? d_coef + (size_t)(j - 1) * rowsz : nullptr;
      a.tile_rows = h->lz_tile_rows;
      if (sharded) {
        // no rank waits at the end of a pass-2 kernel: it signals its completion to the peers, and the NEXT kernel waits for the peers'
        // signals only before its halo tiles (which come last)
        a.goff = resolve_goff(h, ld * (int64_t)sizeof(double));
        a.wait_seq = (h->nranks > 1 && j > 0) ? h->xseq : 0;     // the signal of the previous pass-2 kernel
        a.tail = mk_tail(LZ_TAIL_SIGNAL, j);
      }
      if (q_mode == 0) lz_launch<LZ_P2, CPLX>(h, a, &g);
      else if (q_mode == 1) lz_launch<LZ_P2_SKIP, CPLX>(h, a, &g);
      else lz_launch<LZ_P2_PAIR, CPLX>(h, a, &g);
    }
    sample_end(h, ev);
  }
  k_lz_to_complex<CPLX><<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, ldz, QA, blk(h, BS_ACC) + c0);
  check_launch(h);
  xbarrier(h);   // row-sharded: every rank has finished its last gather of the Lanczos blocks
  sync(h);  // coef is a host buffer
  // algorithmic bytes of an AVERAGE pass-2 launch: paired accumulation moves the accumulator in every second launch only
  h->stats.bytes_kern[FEASTCUDA_KERN_LZ_P2] =
      matfree ? 6.0 * 8.0 * (double)n * nc
              : (mixed ? 0.5 * (lz32_bytes_spmm(h, nc, 3, paired ? 0 : 2) + lz32_bytes_spmm(h, nc, 3, 2))
                       : 0.5 * (lz_bytes_spmm(h, nc, paired ? 3 : 5, CPLX) + lz_bytes_spmm(h, nc, 5, CPLX)));
  drain_events(h);
  h->stats.lz_steps_p2 += k;
  h->stats.ms_lz_p2 += t2.ms();
}

// =====================================================================================================
// Generalized Hermitian problems A x = lambda B x (B Hermitian positive definite) on the multi-shift Lanczos filter
// (CPU restatement: oracle/feast_port.py:mslanczos_filter_gen_cheb).  (z B - A)^-1 B q = (z I - B^-1 A)^-1 q and B^-1 A is
// self-adjoint in the B-inner product, so ONE recurrence in that inner product serves every node with a REAL tridiagonal, like
// the standard case.  The inner solves with B are a FIXED Chebyshev polynomial P = p_K(D^-1 B) D^-1 (k_lz_spmm<LZ_CHEB>: no dot
// products, no data-dependent control flow), so the recurrence is an exact Lanczos process for the neighbouring pencil
// (A, P^-1); the images s_j = P^-1 u_j ride along (s_{j+1} is the right-hand side of the solve that produced u_{j+1}):
//     t = A u_j / beta_j - (beta_j / beta_{j-1}) s_{j-1},   alpha_j beta_j = Re(u_j^H t)                 k_lz_spmm<LZ_P1>
//     s_{j+1} = t - (alpha_j / beta_j) s_j                                                             k_lz_update
//     u_{j+1} = P s_{j+1},   beta_{j+1}^2 = Re(u_{j+1}^H s_{j+1})                                      k_lz_spmm<LZ_CHEB / LZ_CHEB_DOT>
// The per-step scalar kernels, the shifted-residual recurrences and the host-side tridiagonal solves are those of msl_filter.
// =====================================================================================================
static bool cheb_prepare(H* h) {
  if (h->cheb_ready) return h->cheb_usable;
  h->cheb_ready = true;
  h->cheb_usable = false;
  const HostCsr& B = h->hB;
  if (!B.set || B.structure == FEASTCUDA_GEN) return false;
  const int64_t n = B.n;
  const int vs = B.cplx ? 2 : 1;
  std::vector<double> d(n, 0.0), dinv(n), sq(n);
  double hi = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    double rs = 0.0;
    for (int p = B.ptr[i]; p < B.ptr[i + 1]; ++p) {
      const double re = B.val[(size_t)vs * p], im = B.cplx ? B.val[(size_t)vs * p + 1] : 0.0;
      rs += std::hypot(re, im);
      if (B.col[p] == i) d[i] = re;
    }
    if (!(d[i] > 0.0)) return false;          // not positive definite: the caller keeps the per-node Krylov solves
    dinv[i] = 1.0 / d[i];
    sq[i] = std::sqrt(dinv[i]);
    hi = std::max(hi, rs * dinv[i]);          // Gershgorin bound of D^-1 B
  }
  // smallest Ritz value of a short Lanczos run on D^-1/2 B D^-1/2 (an upper bound of the smallest eigenvalue), times a safety factor
  const int steps = (int)std::min<int64_t>(40, n);
  std::vector<zc> v(n), vp(n, zc(0.0)), w(n);
  double nrm = 0.0;
  for (int64_t i = 0; i < n; ++i) { v[i] = std::cos(1.0 + (double)i * 0.7548776662466927); nrm += std::norm(v[i]); }
  nrm = std::sqrt(nrm);
  for (auto& x : v) x /= nrm;
  std::vector<double> al, be;
  double beta = 0.0;
  for (int it = 0; it < steps; ++it) {
    for (int64_t i = 0; i < n; ++i) {
      zc acc(0.0);
      for (int p = B.ptr[i]; p < B.ptr[i + 1]; ++p) {
        const int j = B.col[p];
        const zc bij = B.cplx ? zc(B.val[2 * (size_t)p], B.val[2 * (size_t)p + 1]) : zc(B.val[p], 0.0);
        acc += bij * (sq[j] * v[j]);
      }
      w[i] = sq[i] * acc;
    }
    double a = 0.0;
    for (int64_t i = 0; i < n; ++i) a += (std::conj(v[i]) * w[i]).real();
    double b2 = 0.0;
    for (int64_t i = 0; i < n; ++i) { w[i] -= a * v[i] + beta * vp[i]; b2 += std::norm(w[i]); }
    al.push_back(a);
    beta = std::sqrt(b2);
    if (!(beta > 1e-12 * hi)) break;
    be.push_back(beta);
    for (int64_t i = 0; i < n; ++i) { vp[i] = v[i]; v[i] = w[i] / beta; }
  }
  // smallest eigenvalue of the Lanczos tridiagonal by bisection on its Sturm sequence
  const int k = (int)al.size();
  auto count_below = [&](double x) {
    int cnt = 0;
    double q = 1.0;
    for (int i = 0; i < k; ++i) {
      const double b2 = i > 0 ? be[i - 1] * be[i - 1] : 0.0;
      q = al[i] - x - (i > 0 ? b2 / q : 0.0);
      if (q == 0.0) q = 1e-300;
      if (q < 0.0) ++cnt;
    }
    return cnt;
  };
  double lo_b = -hi, hi_b = 2.0 * hi;
  for (int it = 0; it < 200; ++it) {
    const double mid = 0.5 * (lo_b + hi_b);
    if (count_below(mid) >= 1) hi_b = mid; else lo_b = mid;
  }
  const double ritz_min = 0.5 * (lo_b + hi_b);
  if (!(ritz_min > 1e-3 * hi)) return false;     // too ill-conditioned for a fixed low-degree polynomial
  h->cheb_lo = 0.85 * ritz_min;
  h->cheb_hi = hi;
  h->cheb_dinv.ensure((size_t)n * sizeof(double));
  FC_CUDA(cudaMemcpyAsync(h->cheb_dinv.p, dinv.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  sync(h);
  h->cheb_usable = true;
  if (getenv("FEASTCUDA_VERBOSE")) fprintf(stderr, "[feastcuda r%d] Chebyshev interval of D^-1 B: [%.4e, %.4e]\n", h->rank, h->cheb_lo, h->cheb_hi);
  return true;
}

static int cheb_degree(double lo, double hi, double delta) {
  const double kap = hi / lo, sg = (std::sqrt(kap) - 1.0) / (std::sqrt(kap) + 1.0);
  int K = 2;
  while (2.0 * std::pow(sg, K) / (1.0 + std::pow(sg, 2 * K)) > delta && K < 400) ++K;
  return K;
}

template <bool CPLX>
static void msl_filter_gen(H* h, int basis_slot, int c0, int nc, bool have_ritz, const double* theta, const zc* Zne, const zc* Wne,
                           int ne, double target, int kmax, int check_every, double cheb_delta, MslOut& out) {
  FC_REQUIRE(h->kind == OP_SPARSE && h->dev_complex == CPLX && h->has_b && h->cheb_usable,
             "generalized multi-shift Lanczos needs a sparse Hermitian pencil with a positive definite B");
  FC_REQUIRE(CPLX || (c0 & 1) == 0, "column slices must start at an even column");
  const int64_t n = h->ws_n;
  const int64_t ldz = h->ws_ld, ld = CPLX ? 2 * (int64_t)nc : ((nc + 1) & ~1);
  FC_REQUIRE((double)n * (double)ld < 4294967296.0, "multi-shift Lanczos: n*ld must be below 2^32 (32-bit gather offsets)");
  kmax = std::max(1, std::min(kmax, 16384));
  check_every = std::max(1, check_every);
  const int P = CPLX ? nc : (nc + 1) / 2;
  const int pp = pow2_ge(P);
  const int egrid = (int)std::max<int64_t>(1, std::min<int64_t>((n + (256 / pp) - 1) / (256 / pp), (int64_t)h->sms * h->lz_egrid_mult));
  const int K = cheb_degree(h->cheb_lo, h->cheb_hi, cheb_delta);
  const int slice = getenv("FEASTCUDA_GEN_SLICE") ? atoi(getenv("FEASTCUDA_GEN_SLICE")) : 128;   // complex columns per gather launch (slices measured SLOWER: 2 x 48 columns 0.52 ms against 0.30 ms unsliced)
  const double* dinv = h->cheb_dinv.as<double>();

  // ---- device scalars (layout as in msl_filter) -------------------------------------------------------------------
  const size_t rowsz = (size_t)FC_MAXCOLS;
  const size_t per = (size_t)(kmax + 2) * rowsz;
  const size_t scal_doubles = 5 * per + rowsz + 2 * rowsz + (size_t)(kmax + 2);
  h->lz_scal.ensure(scal_doubles * sizeof(double));
  h->lz_coef.ensure(per * sizeof(double));
  h->lz_state.ensure(((size_t)2 * ne * rowsz + ne) * sizeof(zd) + 64);
  double* base = h->lz_scal.as<double>();
  FC_CUDA(cudaMemsetAsync(base, 0, scal_doubles * sizeof(double), h->stream));
  LzScalars S;
  S.alpha = base; S.beta = base + per; S.inv_beta = base + 2 * per; S.ratio_b = base + 3 * per; S.ratio_a = base + 4 * per;
  S.scale = base + 5 * per;
  double* d_theta = S.scale + rowsz;
  double* d_rho = d_theta + rowsz;
  S.maxres = d_rho + rowsz;
  zd* stz = h->lz_state.as<zd>();
  S.d = stz; S.g = stz + (size_t)ne * rowsz;
  zd* d_z = stz + (size_t)2 * ne * rowsz;
  S.z = d_z;
  S.ne = ne;
  S.target = target;
  S.done_k = reinterpret_cast<int*>(d_z + ne);
  FC_CUDA(cudaMemcpyAsync(d_z, Zne, (size_t)ne * sizeof(zd), cudaMemcpyHostToDevice, h->stream));

  // ---- compact work blocks ------------------------------------------------------------------------------------------
  double* RQ = rblk(h, BS_KS);                                             // the basis slice (start block only)
  double* SV[2] = {rblk(h, BS_KR), rblk(h, BS_KRH)};                       // s_j / s_{j-1}
  double* VV[3] = {rblk(h, BS_KP), rblk(h, BS_KT), rblk(h, BS_KX)};        // u_j and the two Chebyshev iterates
  double* QA = rblk(h, BS_KV);
  const zd* basis = blk(h, basis_slot) + c0;
  double* part = h->partial_r.as<double>();
  const double th_c = 0.5 * (h->cheb_hi + h->cheb_lo), de_c = 0.5 * (h->cheb_hi - h->cheb_lo), sg_c = th_c / de_c;

  auto args0 = [&]() {
    LzArgs a;
    memset(&a, 0, sizeof(a));
    a.n = n; a.m = nc; a.ld = ld; a.partial = part; a.pstride = FC_MAXCOLS; a.tile_rows = h->lz_tile_rows;
    return a;
  };
  auto use_A = [&](LzArgs& a) { a.ptr = h->dA.ptr.as<int>(); a.col = h->dA.col.as<int>(); a.val = h->dA.val.p; };
  auto use_B = [&](LzArgs& a) { a.ptr = h->dB.ptr.as<int>(); a.col = h->dB.col.as<int>(); a.val = h->dB.val.p; };
  const double cheb_bytes = (double)h->hB.nnz * ((CPLX ? 16.0 : 8.0) + 4.0) + 4.0 * (double)(n + 1) + 8.0 * (double)n +
                            4.0 * (double)n * nc * (CPLX ? 16.0 : 8.0);
  // X = P R: K Chebyshev steps; the result lands in xa or xb (returned); the last step leaves Re(X^H R) in `part`
  int g_last = 0, cur_j = 0;
  int64_t cheb_count = 0;
  auto cheb_solve = [&](const double* R, double* xa, double* xb, const int* done) -> double* {
    k_lz_cheb_first<CPLX><<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, 1.0 / th_c, dinv, R, xa, done);
    check_launch(h);
    double rho_prev = 1.0 / sg_c;
    double* x = xa;      // x_i
    double* xo = xb;     // x_{i-1} (i = 1: never read with weight, prev = x)
    for (int i = 1; i < K; ++i) {
      const double rho = 1.0 / (2.0 * sg_c - rho_prev);
      LzArgs a = args0();
      use_B(a);
      a.U = x; a.prev = (i == 1) ? x : xo; a.out = xo; a.rhs = R; a.dinv = dinv; a.done = done;
      a.c1 = rho * rho_prev; a.c2 = 2.0 * rho / de_c;
      const bool smp = (cheb_count++ % 64) == 5;
      const int ev = smp ? sample_begin(h, FEASTCUDA_KERN_LZ_CHEB, cur_j) : -1;
      if (i == K - 1) lz_launch_sliced<LZ_CHEB_DOT, CPLX>(h, a, &g_last, slice);
      else lz_launch_sliced<LZ_CHEB, CPLX>(h, a, &g_last, slice);
      sample_end(h, ev);
      std::swap(x, xo);
      rho_prev = rho;
    }
    return x;
  };

  // ---- start block: s_0 = B q (or A q - theta B q), u_0 = P s_0, beta_0^2 = Re(u_0^H s_0) ----------------------------
  std::vector<double> rho(nc, 0.0);
  k_lz_real_part<CPLX><<<egrid, 256, 0, h->stream>>>(n, nc, pp, ldz, ld, basis, RQ, nullptr, FC_MAXCOLS, LzTail{});
  check_launch(h);
  FC_CUDA(cudaMemsetAsync(QA, 0, (size_t)n * (size_t)ld * sizeof(double), h->stream));
  int g = 0;
  if (!have_ritz) {
    LzArgs a = args0();
    use_B(a);
    a.U = RQ; a.prev = RQ; a.out = SV[0];
    lz_launch_sliced<LZ_PLAIN, CPLX>(h, a, &g, slice);
  } else {
    for (int c = 0; c < nc; ++c) {
      zc acc(0.0);
      for (int e = 0; e < ne; ++e) acc += 2.0 * Wne[e] / (Zne[e] - theta[c]);
      rho[c] = acc.real();
    }
    FC_CUDA(cudaMemcpyAsync(d_theta, theta, (size_t)nc * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    FC_CUDA(cudaMemcpyAsync(d_rho, rho.data(), (size_t)nc * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    LzArgs a = args0();
    use_B(a);
    a.U = RQ; a.prev = RQ; a.out = VV[1];                       // B q
    lz_launch_sliced<LZ_PLAIN, CPLX>(h, a, &g, slice);
    LzArgs r = args0();
    use_A(r);
    r.U = RQ; r.prev = RQ; r.out = SV[0]; r.own = VV[1]; r.s_theta = d_theta; r.s_coef = d_rho; r.Q = nullptr;   // A q - theta (B q)
    lz_launch_sliced<LZ_RES, CPLX>(h, r, &g, slice);
    k_lz_axpy<CPLX><<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, d_rho, RQ, QA);      // Q = rho(theta) q
    check_launch(h);
    sync(h);   // theta / rho are host buffers
  }
  double* Ucur = cheb_solve(SV[0], VV[0], VV[1], nullptr);
  double* free_a = (Ucur == VV[0]) ? VV[1] : VV[0];
  double* free_b = VV[2];
  k_lz_scal_init<<<1, 1024, 0, h->stream>>>(S, part, g_last, FC_MAXCOLS, nc);
  check_launch(h);
  // the start vectors are needed again by pass 2: keep copies (u_0 in RQ's slot once q is no longer needed, s_0 in BS_KB)
  double* U0 = RQ;
  double* S0 = rblk(h, BS_KB);
  FC_CUDA(cudaMemcpyAsync(U0, Ucur, (size_t)n * (size_t)ld * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  FC_CUDA(cudaMemcpyAsync(S0, SV[0], (size_t)n * (size_t)ld * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));

  // one Lanczos step with the given buffers; pass 1 raises the scalars, pass 2 replays them
  auto step = [&](int j, double*& U, double*& fa, double*& fb, int& si, bool pass1, const int* done) {
    double* s_cur = SV[si];
    double* s_oth = SV[si ^ 1];
    cur_j = pass1 ? j : 0;
    LzArgs a = args0();
    use_A(a);
    a.U = U; a.prev = j > 0 ? s_oth : U; a.out = s_oth; a.done = done;
    a.s_inv_beta = S.inv_beta + (size_t)j * rowsz; a.s_ratio_b = S.ratio_b + (size_t)j * rowsz;
    int gg = 0;
    const bool smp = pass1 && (j % 16) == 3;
    int ev = smp ? sample_begin(h, FEASTCUDA_KERN_LZ_P1, j) : -1;
    lz_launch_sliced<LZ_P1, CPLX>(h, a, &gg, slice);
    sample_end(h, ev);
    if (pass1) {
      k_lz_scal1<<<1, 1024, 0, h->stream>>>(S, j, part, gg, FC_MAXCOLS, nc);
      check_launch(h);
    }
    k_lz_update<CPLX><<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, S.ratio_a + (size_t)j * rowsz, s_cur, s_oth, part, FC_MAXCOLS, done, LzTail{});
    check_launch(h);
    double* unew = cheb_solve(s_oth, fa, fb, done);
    if (pass1) {
      k_lz_scal2<<<1, 1024, 0, h->stream>>>(S, j, part, g_last, FC_MAXCOLS, nc);
      check_launch(h);
    }
    double* other = (unew == fa) ? fb : fa;
    fa = U;          // the old u_j is free now
    fb = other;
    U = unew;
    si ^= 1;
  };

  // ---- pass 1 ---------------------------------------------------------------------------------------------------------
  Timer t1;
  int* flag = reinterpret_cast<int*>(pinned_buf(h, 64));
  flag[0] = 0;
  int done = 0, kfinal = 0, si = 0;
  {
    double *U = Ucur, *fa = free_a, *fb = free_b;
    while (done < kmax) {
      const int batch = std::min(check_every, kmax - done);
      for (int j = done; j < done + batch; ++j) step(j, U, fa, fb, si, true, S.done_k);
      done += batch;
      FC_CUDA(cudaMemcpyAsync(flag, S.done_k, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      sync(h);
      if (flag[0] != 0) { kfinal = flag[0]; out.converged = true; break; }
    }
  }
  if (kfinal == 0) kfinal = done;
  h->stats.bytes_kern[FEASTCUDA_KERN_LZ_P1] = lz_bytes_spmm(h, nc, 3, CPLX);
  h->stats.bytes_kern[FEASTCUDA_KERN_LZ_CHEB] = cheb_bytes;
  drain_events(h, kfinal);
  h->stats.lz_steps_p1 += kfinal;
  h->stats.krylov_iters += kfinal;
  h->stats.col_iters += (int64_t)kfinal * nc;
  h->stats.ms_lz_p1 += t1.ms();

  // ---- coefficients (host, as in msl_filter) ----------------------------------------------------------------------------
  const int k = kfinal;
  std::vector<double> al((size_t)k * rowsz), be((size_t)(k + 1) * rowsz), coef((size_t)k * rowsz, 0.0);
  FC_CUDA(cudaMemcpyAsync(al.data(), S.alpha, al.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  FC_CUDA(cudaMemcpyAsync(be.data(), S.beta, be.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  double mr = 0.0;
  FC_CUDA(cudaMemcpyAsync(&mr, S.maxres + (k - 1), sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  sync(h);
  out.k = k;
  out.maxres = mr;
  lz_host_coefficients(k, nc, al, be, have_ritz, theta, Zne, Wne, ne, coef);
  double* d_coef = h->lz_coef.as<double>();
  FC_CUDA(cudaMemcpyAsync(d_coef, coef.data(), coef.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));

  // ---- pass 2: replay from (u_0, s_0) with the stored scalars, Q += (c_j / beta_j) u_j ----------------------------------
  Timer t2;
  {
    // buffers: u_0 and s_0 were saved; every other block is free again
    double* U = VV[0];
    FC_CUDA(cudaMemcpyAsync(U, U0, (size_t)n * (size_t)ld * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    FC_CUDA(cudaMemcpyAsync(SV[0], S0, (size_t)n * (size_t)ld * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    double *fa = VV[1], *fb = VV[2];
    int si2 = 0;
    for (int j = 0; j < k; ++j) {
      k_lz_axpy<CPLX><<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, d_coef + (size_t)j * rowsz, U, QA);
      check_launch(h);
      if (j == k - 1) break;
      step(j, U, fa, fb, si2, false, nullptr);
    }
  }
  k_lz_to_complex<CPLX><<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, ldz, QA, blk(h, BS_ACC) + c0);
  check_launch(h);
  sync(h);
  drain_events(h);
  h->stats.lz_steps_p2 += k;
  h->stats.ms_lz_p2 += t2.ms();
  h->stats.cheb_degree = K;
}

// =====================================================================================================
// General (non-Hermitian) pencils on a SHARED Krylov space: two-sided multi-shift Lanczos, two passes.
// (z B - A)^-1 B q = (z I - C)^-1 q with C = B^-1 A; the non-Hermitian Lanczos relation  C V_k = V_k T_k + delta_{k+1} v_{k+1} e_k^T  (T_k complex
// tridiagonal, W_k^H V_k = diag(d)) does not depend on z, so ONE recurrence per column serves every node of the full contour:
//     x_e = ||b|| V_k (z_e I - T_k)^-1 e_1,      residual_e = delta_{k+1} |e_k^T (z_e I - T_k)^-1 e_1| ||b||      (||v_j||_2 = 1)
// and  q = sum_e w_e x_e = V_k c  (kernel/feast_kernel.jl:762-766 accumulates q += w_e Y_e node by node).  Pass 1 builds T_k (the recurrence
// scalars live on the HOST: two small reductions per step against ~30 SpMM launches), pass 2 replays the v-sequence alone and accumulates.
// The rows are scaled by diag(B)^-1 once (C is unchanged); the solves with B' = D^-1 B are a fixed number of Jacobi sweeps
// x <- x + (r - B' x)  (k_lz_spmm<LZ_CHEB> with c1 = 0, c2 = 1), available when B' is strictly diagonally dominant by rows and by columns
// (mass-like perturbations of the identity, configs[4]); the same polynomial in B'^H gives the exact adjoint for the w-sequence.
// The reference runs ne x M0 GMRES solves per loop here (sparse/feast_sparse.jl:873-1006 -> solve_shifted_iterative!, :164-236).
// =====================================================================================================
static void host_transpose_conj(int64_t n, const std::vector<int>& ptr, const std::vector<int>& col, const std::vector<zc>& val,
                                std::vector<int>& tptr, std::vector<int>& tcol, std::vector<zc>& tval) {
  const size_t nnz = col.size();
  tptr.assign(n + 1, 0);
  tcol.assign(nnz, 0);
  tval.assign(nnz, zc(0.0));
  for (size_t p = 0; p < nnz; ++p) tptr[col[p] + 1]++;
  for (int64_t i = 0; i < n; ++i) tptr[i + 1] += tptr[i];
  std::vector<int> pos(tptr.begin(), tptr.end() - 1);
  for (int64_t i = 0; i < n; ++i)
    for (int p = ptr[i]; p < ptr[i + 1]; ++p) {
      const int q = pos[col[p]]++;
      tcol[q] = (int)i;
      tval[q] = std::conj(val[p]);
    }
}

static void upload_csr_z(H* h, int64_t n, const std::vector<int>& ptr, const std::vector<int>& col, const std::vector<zc>& val, DevCsr& dst) {
  dst.ptr.ensure((n + 1) * sizeof(int));
  dst.col.ensure(std::max<size_t>(col.size(), 1) * sizeof(int));
  dst.val.ensure(std::max<size_t>(val.size(), 1) * sizeof(zd));
  FC_CUDA(cudaMemcpyAsync(dst.ptr.p, ptr.data(), (n + 1) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  if (!col.empty()) {
    FC_CUDA(cudaMemcpyAsync(dst.col.p, col.data(), col.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    FC_CUDA(cudaMemcpyAsync(dst.val.p, val.data(), val.size() * sizeof(zd), cudaMemcpyHostToDevice, h->stream));
  }
  sync(h);
  dst.uploaded = true;
}

static bool gen2_prepare(H* h) {
  if (h->g2_ready) return h->g2_usable;
  h->g2_ready = true;
  h->g2_usable = false;
  const HostCsr& A = h->hA;
  if (!A.set) return false;
  const int64_t n = A.n;
  auto entry = [](const HostCsr& M, int p) { return M.cplx ? zc(M.val[2 * (size_t)p], M.val[2 * (size_t)p + 1]) : zc(M.val[p], 0.0); };
  std::vector<zc> dinv(n, zc(1.0));
  double q = 0.0;
  std::vector<zc> bval;
  if (h->has_b) {
    const HostCsr& B = h->hB;
    std::vector<double> colsum(n, 0.0);
    for (int64_t i = 0; i < n; ++i) {
      zc d(0.0);
      for (int p = B.ptr[i]; p < B.ptr[i + 1]; ++p) if (B.col[p] == i) d = entry(B, p);
      if (!(std::abs(d) > 0.0)) return false;
      dinv[i] = zc(1.0) / d;
    }
    bval.resize(B.nnz);
    for (int64_t i = 0; i < n; ++i) {
      double rs = 0.0;
      for (int p = B.ptr[i]; p < B.ptr[i + 1]; ++p) {
        bval[p] = entry(B, p) * dinv[i];
        if (B.col[p] != i) { rs += std::abs(bval[p]); colsum[B.col[p]] += std::abs(bval[p]); }
        else bval[p] = zc(1.0);
      }
      q = std::max(q, rs);
    }
    for (int64_t i = 0; i < n; ++i) q = std::max(q, colsum[i]);
    if (!(q < 0.9)) return false;      // the Jacobi sweeps need a strictly diagonally dominant B (rows and columns)
  }
  std::vector<zc> aval(A.nnz);
  for (int64_t i = 0; i < n; ++i)
    for (int p = A.ptr[i]; p < A.ptr[i + 1]; ++p) aval[p] = entry(A, p) * dinv[i];
  std::vector<int> tptr, tcol;
  std::vector<zc> tval;
  upload_csr_z(h, n, A.ptr, A.col, aval, h->g2A);
  host_transpose_conj(n, A.ptr, A.col, aval, tptr, tcol, tval);
  upload_csr_z(h, n, tptr, tcol, tval, h->g2Ah);
  if (h->has_b) {
    upload_csr_z(h, n, h->hB.ptr, h->hB.col, bval, h->g2B);
    host_transpose_conj(n, h->hB.ptr, h->hB.col, bval, tptr, tcol, tval);
    upload_csr_z(h, n, tptr, tcol, tval, h->g2Bh);
  }
  std::vector<double> ones(n, 1.0);
  h->g2_ones.ensure((size_t)n * sizeof(double));
  FC_CUDA(cudaMemcpyAsync(h->g2_ones.p, ones.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  sync(h);
  h->g2_q = q;
  h->g2_usable = true;
  if (getenv("FEASTCUDA_VERBOSE")) fprintf(stderr, "[feastcuda r%d] two-sided Lanczos: Jacobi contraction bound of D^-1 B: %.3f\n", h->rank, q);
  return true;
}

static void msl2_filter(H* h, int basis_slot, int c0, int nc, bool have_ritz, const zc* theta, const zc* Zne, const zc* Wne, int ne,
                        double target, int kmax, double jac_delta, MslOut& out) {
  FC_REQUIRE(h->kind == OP_SPARSE && h->g2_usable && nc >= 1 && nc <= TS_MAXC, "two-sided Lanczos filter: sparse pencil, at most 64 columns per rank");
  const int64_t n = h->ws_n;
  const int64_t ldz = h->ws_ld, ld = 2 * (int64_t)nc;
  FC_REQUIRE((double)n * (double)ld < 4294967296.0, "multi-shift Lanczos: n*ld must be below 2^32 (32-bit gather offsets)");
  kmax = std::max(2, std::min(kmax, 16384));
  const int pp = pow2_ge(nc);
  const int egrid = (int)std::max<int64_t>(1, std::min<int64_t>((n + (256 / pp) - 1) / (256 / pp), (int64_t)h->sms * h->lz_egrid_mult));
  const bool hasb = h->has_b;
  int KJ = 1;
  if (hasb && h->g2_q > 0) KJ = std::max(1, (int)std::ceil(std::log(jac_delta) / std::log(h->g2_q)));
  double* Vb[2] = {rblk(h, BS_KR), rblk(h, BS_KRH)};
  double* Wb[2] = {rblk(h, BS_KP), rblk(h, BS_KV)};
  double* X[2] = {rblk(h, BS_KS), rblk(h, BS_KT)};
  double* T = rblk(h, BS_KX);
  double* QA = rblk(h, BS_KB);
  double* RQ = rblk(h, BS_RHS);
  h->partial_r.ensure((size_t)4 * std::max(egrid, h->sms * 8) * FC_MAXCOLS * sizeof(double));   // four sums per column and CTA (k_ts_dot3)
  double* part = h->partial_r.as<double>();
  const zd* basis = blk(h, basis_slot) + c0;
  const size_t blk_bytes = (size_t)n * (size_t)ld * sizeof(double);

  auto args0 = [&](const DevCsr& M) {
    LzArgs a;
    memset(&a, 0, sizeof(a));
    a.n = n; a.m = nc; a.ld = ld; a.partial = part; a.pstride = FC_MAXCOLS; a.tile_rows = h->lz_tile_rows;
    a.ptr = M.ptr.as<int>(); a.col = M.col.as<int>(); a.val = M.val.p;
    return a;
  };
  int g = 0;
  auto spmm = [&](const DevCsr& M, const double* src, double* dst) {
    LzArgs a = args0(M);
    a.U = src; a.prev = src; a.out = dst;
    lz_launch<LZ_PLAIN, true>(h, a, &g);
  };
  // x ~ M^-1 r by KJ Jacobi sweeps from x_1 = r (unit diagonal); returns the buffer holding the result (r itself when KJ == 1)
  auto jacobi = [&](const DevCsr& M, const double* r) -> const double* {
    const double* x = r;
    for (int i = 1; i < KJ; ++i) {
      double* o = (x == X[0]) ? X[1] : X[0];
      LzArgs a = args0(M);
      a.U = x; a.prev = x; a.out = o; a.rhs = r; a.dinv = h->g2_ones.as<double>(); a.c1 = 0.0; a.c2 = 1.0;
      lz_launch<LZ_CHEB, true>(h, a, &g);
      x = o;
    }
    return x;
  };
  auto applyC = [&](const double* src) -> const double* {          // B'^-1 (A' src)
    spmm(h->g2A, src, T);
    return hasb ? jacobi(h->g2B, T) : T;
  };
  auto applyCh = [&](const double* src) -> const double* {         // A'^H (B'^-H src)
    const double* s = hasb ? jacobi(h->g2Bh, src) : src;
    spmm(h->g2Ah, s, T);
    return T;
  };
  auto reduce = [&](int nslots, std::vector<double>& hostv) {
    double* dev = h->red_ws.as<double>();
    k_reduce_partials<double><<<1, 1024, ((size_t)nslots * FC_MAXCOLS + 1024) * sizeof(double), h->stream>>>(part, nslots, egrid, FC_MAXCOLS, nc, dev);
    check_launch(h);
    hostv.resize((size_t)nslots * nc);
    FC_CUDA(cudaMemcpyAsync(hostv.data(), dev, hostv.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    sync(h);
  };
  TsScal sc;
  auto clear_sc = [&]() { memset(&sc, 0, sizeof(sc)); };

  // ---- start block: v_0 = q (or C q - theta q), normalised; w_0 = v_0 ----------------------------------------------------
  k_lz_real_part<true><<<egrid, 256, 0, h->stream>>>(n, nc, pp, ldz, ld, basis, RQ, nullptr, FC_MAXCOLS, LzTail{});
  check_launch(h);
  FC_CUDA(cudaMemsetAsync(QA, 0, blk_bytes, h->stream));
  if (!have_ritz) {
    FC_CUDA(cudaMemcpyAsync(Vb[0], RQ, blk_bytes, cudaMemcpyDeviceToDevice, h->stream));
  } else {
    const double* cq = applyC(RQ);
    clear_sc();
    for (int c = 0; c < nc; ++c) sc.a[c] = make_double2(theta[c].real(), theta[c].imag());
    k_ts_lin3<<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, cq, RQ, RQ, Vb[0], sc);      // C q - theta q  (b = 0)
    check_launch(h);
    clear_sc();
    for (int c = 0; c < nc; ++c) {
      zc acc(0.0);
      for (int e = 0; e < ne; ++e) acc += Wne[e] / (Zne[e] - theta[c]);
      sc.a[c] = make_double2(acc.real(), acc.imag());
    }
    k_ts_caxpy<<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, RQ, QA, sc);                // q * sum_e w_e / (z_e - theta)
    check_launch(h);
  }
  std::vector<double> hv;
  k_ts_dot3<<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, Vb[0], Vb[0], part, FC_MAXCOLS);
  check_launch(h);
  reduce(4, hv);
  std::vector<double> beta0(nc);
  std::vector<char> dead(nc, 0);
  clear_sc();
  for (int c = 0; c < nc; ++c) {
    beta0[c] = std::sqrt(std::max(hv[c], 0.0));
    if (!(beta0[c] > 1e-290)) { dead[c] = 1; beta0[c] = 0.0; }
    sc.a[c].x = dead[c] ? 0.0 : 1.0 / beta0[c];
  }
  k_ts_scale2<<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, Vb[0], nullptr, sc);
  check_launch(h);
  FC_CUDA(cudaMemcpyAsync(Wb[0], Vb[0], blk_bytes, cudaMemcpyDeviceToDevice, h->stream));
  FC_CUDA(cudaMemcpyAsync(RQ, Vb[0], blk_bytes, cudaMemcpyDeviceToDevice, h->stream));      // pass 2 starts from the same v_0
  FC_CUDA(cudaMemsetAsync(Vb[1], 0, blk_bytes, h->stream));
  FC_CUDA(cudaMemsetAsync(Wb[1], 0, blk_bytes, h->stream));

  // ---- pass 1 (host scalars) --------------------------------------------------------------------------------------------------
  Timer t1;
  const size_t K1 = (size_t)kmax + 2;
  std::vector<zc> alpha(K1 * nc, zc(0.0)), betap(K1 * nc, zc(0.0)), dd(K1 * nc, zc(1.0));
  std::vector<double> delta(K1 * nc, 0.0), gam(K1 * nc, 0.0), scale(nc, 0.0);
  std::vector<int> kc(nc, 0);                    // steps of column c (its T is kc x kc); 0 while the column is alive
  std::vector<zc> lud((size_t)ne * nc), lug((size_t)ne * nc);
  int cur = 0, k = 0;
  double maxres = 0.0;
  auto at = [&](int j, int c) { return (size_t)j * nc + c; };
  for (int j = 0; j < kmax; ++j) {
    const double* cv = applyC(Vb[cur]);
    k_ts_dot<<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, Wb[cur], cv, part, FC_MAXCOLS);
    check_launch(h);
    reduce(2, hv);
    clear_sc();
    for (int c = 0; c < nc; ++c) {
      zc al(0.0);
      if (!dead[c] && !kc[c]) al = zc(hv[c], hv[nc + c]) / dd[at(j, c)];
      alpha[at(j, c)] = al;
      scale[c] = std::max(scale[c], std::abs(al));
      sc.a[c] = make_double2(al.real(), al.imag());
      const zc bp = betap[at(j, c)];
      sc.b[c] = make_double2(bp.real(), bp.imag());
    }
    k_ts_lin3<<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, cv, Vb[cur], Vb[cur ^ 1], Vb[cur ^ 1], sc);
    check_launch(h);
    const double* cw = applyCh(Wb[cur]);
    for (int c = 0; c < nc; ++c) {
      const zc al = std::conj(alpha[at(j, c)]);
      sc.a[c] = make_double2(al.real(), al.imag());
      zc gm(0.0);
      if (j > 0 && !dead[c] && !kc[c]) gm = std::conj(delta[at(j, c)] * dd[at(j, c)] / dd[at(j - 1, c)]);
      sc.b[c] = make_double2(gm.real(), gm.imag());
    }
    k_ts_lin3<<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, cw, Wb[cur], Wb[cur ^ 1], Wb[cur ^ 1], sc);
    check_launch(h);
    k_ts_dot3<<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, Vb[cur ^ 1], Wb[cur ^ 1], part, FC_MAXCOLS);
    check_launch(h);
    reduce(4, hv);
    clear_sc();
    maxres = 0.0;
    for (int c = 0; c < nc; ++c) {
      double dl = 0.0, gm = 0.0;
      zc dn(1.0);
      bool alive = !dead[c] && !kc[c];
      if (alive) {
        dl = std::sqrt(std::max(hv[c], 0.0));
        gm = std::sqrt(std::max(hv[nc + c], 0.0));
        const bool ok = dl > 1e-290 && gm > 1e-290 && dl > 1e-13 * std::max(scale[c], 1e-300);
        if (ok) dn = zc(hv[2 * nc + c], hv[3 * nc + c]) / (dl * gm);
        if (!ok || !(std::abs(dn) > 1e-10)) {      // invariant subspace reached, or a (near) breakdown of the two-sided process: the column stops here
          alive = false;
          kc[c] = j + 1;
          if (!ok) dl = 0.0;
        }
        scale[c] = std::max(scale[c], dl);
      }
      delta[at(j + 1, c)] = dl;
      gam[at(j + 1, c)] = gm;
      dd[at(j + 1, c)] = alive ? dn : zc(1.0);
      betap[at(j + 1, c)] = alive ? gm * dn / dd[at(j, c)] : zc(0.0);
      sc.a[c].x = alive ? 1.0 / dl : 0.0;
      sc.b[c].x = alive ? 1.0 / gm : 0.0;
      // shifted residuals of every node: LU pivots of z I - T advanced by one step
      if (!dead[c] && (kc[c] == 0 || kc[c] == j + 1)) {
        for (int e = 0; e < ne; ++e) {
          const size_t o = (size_t)e * nc + c;
          zc d = Zne[e] - alpha[at(j, c)];
          zc gg;
          if (j == 0) gg = zc(1.0) / d;
          else {
            d -= betap[at(j, c)] * delta[at(j, c)] / lud[o];
            gg = delta[at(j, c)] * lug[o] / d;
          }
          lud[o] = d;
          lug[o] = gg;
          maxres = std::max(maxres, dl * std::abs(gg));
        }
      }
    }
    k_ts_scale2<<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, Vb[cur ^ 1], Wb[cur ^ 1], sc);
    check_launch(h);
    cur ^= 1;
    k = j + 1;
    h->stats.krylov_iters++;
    if (!(maxres > target)) { out.converged = true; break; }
  }
  out.k = k;
  out.maxres = maxres;
  h->stats.lz_steps_p1 += k;
  h->stats.col_iters += (int64_t)k * nc;
  h->stats.ms_lz_p1 += t1.ms();

  // ---- coefficients: c_j = beta_0 sum_e w_e F_e [(z_e I - T)^-1 e_1]_j (complex Thomas solves on the host) ---------------------
  std::vector<zc> coef((size_t)k * nc, zc(0.0)), ludv(k), luf(k);
  for (int c = 0; c < nc; ++c) {
    if (dead[c]) continue;
    const int kk = kc[c] ? std::min(kc[c], k) : k;
    for (int e = 0; e < ne; ++e) {
      const zc z = Zne[e];
      const zc wf = Wne[e] * (have_ritz ? zc(1.0) / (z - theta[c]) : zc(1.0)) * beta0[c];
      ludv[0] = z - alpha[at(0, c)];
      luf[0] = 1.0;
      for (int j = 1; j < kk; ++j) {
        const zc l = delta[at(j, c)] / ludv[j - 1];
        ludv[j] = (z - alpha[at(j, c)]) - l * betap[at(j, c)];
        luf[j] = l * luf[j - 1];
      }
      zc y = luf[kk - 1] / ludv[kk - 1];
      coef[at(kk - 1, c)] += wf * y;
      for (int j = kk - 2; j >= 0; --j) {
        y = (luf[j] + betap[at(j + 1, c)] * y) / ludv[j];
        coef[at(j, c)] += wf * y;
      }
    }
  }

  // ---- pass 2: replay the v-sequence, Q += c_j v_j ----------------------------------------------------------------------------------
  Timer t2;
  FC_CUDA(cudaMemcpyAsync(Vb[0], RQ, blk_bytes, cudaMemcpyDeviceToDevice, h->stream));
  FC_CUDA(cudaMemsetAsync(Vb[1], 0, blk_bytes, h->stream));
  cur = 0;
  for (int j = 0; j < k; ++j) {
    clear_sc();
    for (int c = 0; c < nc; ++c) sc.a[c] = make_double2(coef[at(j, c)].real(), coef[at(j, c)].imag());
    k_ts_caxpy<<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, Vb[cur], QA, sc);
    check_launch(h);
    if (j == k - 1) break;
    const double* cv = applyC(Vb[cur]);
    clear_sc();
    for (int c = 0; c < nc; ++c) {
      sc.a[c] = make_double2(alpha[at(j, c)].real(), alpha[at(j, c)].imag());
      sc.b[c] = make_double2(betap[at(j, c)].real(), betap[at(j, c)].imag());
    }
    k_ts_lin3<<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, cv, Vb[cur], Vb[cur ^ 1], Vb[cur ^ 1], sc);
    check_launch(h);
    clear_sc();
    for (int c = 0; c < nc; ++c) {
      const bool alive = !dead[c] && (kc[c] == 0 || j + 1 < kc[c]);
      sc.a[c].x = (alive && delta[at(j + 1, c)] > 0) ? 1.0 / delta[at(j + 1, c)] : 0.0;
    }
    k_ts_scale2<<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, Vb[cur ^ 1], nullptr, sc);
    check_launch(h);
    cur ^= 1;
  }
  k_lz_to_complex<true><<<egrid, 256, 0, h->stream>>>(n, nc, pp, ld, ldz, QA, blk(h, BS_ACC) + c0);
  check_launch(h);
  sync(h);
  h->stats.lz_steps_p2 += k;
  h->stats.ms_lz_p2 += t2.ms();
  h->stats.cheb_degree = KJ;
}

// =====================================================================================================
// rank-revealing orthonormalisation (K7: _feast_qr_compress!, core/feast_aux.jl:101-131)
// in: Z0 (n x ncols in slot `src`), out: orthonormal basis in the returned slot, rank
// =====================================================================================================
static int orthonormalize(H* h, int ncols, int src_slot, int tmp_slot, double rank_tol, int* out_slot) {
  const int64_t n = h->ws_n;
  int cur = ncols, done = 0, a_slot = src_slot, b_slot = tmp_slot;
  double thr_abs = -1.0;
  const double eps = 2.220446049250313e-16;
  std::vector<zc> G;
  for (int pass = 0; pass < 16; ++pass) {
    gram_host(h, cur, cur, blk(h, a_slot), blk(h, a_slot), G);
    h->stats.ortho_passes++;
    if (pass == 0) {
      double dmax = 0.0;
      for (int j = 0; j < cur; ++j) dmax = std::max(dmax, G[(size_t)j * cur + j].real());
      if (!(dmax > 0.0)) { *out_slot = a_slot; return 0; }
      thr_abs = std::max(rank_tol, eps * (double)std::max<int64_t>(h->row_sharded ? h->n_glob : n, ncols)) * std::sqrt(dmax);
    }
    if (done == cur) {
      double err = 0.0;
      for (int i = 0; i < cur; ++i)
        for (int j = 0; j < cur; ++j) err = std::max(err, std::abs(G[(size_t)i * cur + j] - (i == j ? zc(1.0) : zc(0.0))));
      if (err <= 1e-14) break;
    }
    OrthoPass op = ortho_pass(G, cur, done, thr_abs, 1e-8);
    const int nout = op.accepted + op.kept;
    if (nout == 0) { *out_slot = a_slot; return 0; }
    rowtransform(h, cur, nout, blk(h, a_slot), op.T, blk(h, b_slot));
    std::swap(a_slot, b_slot);
    done = op.accepted;
    cur = nout;
  }
  *out_slot = a_slot;
  return done;
}

// =====================================================================================================
// reduced eigenproblem on device (K9)
// =====================================================================================================
static void reduced_eig(H* h, int r, const std::vector<zc>& Sq, const std::vector<zc>* Aq, std::vector<double>& lam,
                        std::vector<zc>& V, int* status) {
  const size_t rr = (size_t)r * r;
  zd* base = h->small.as<zd>();
  zd *dS = base, *dB = base + rr, *dV = base + 2 * rr, *dW = base + 3 * rr, *dX = base + 4 * rr;
  double* dlam = reinterpret_cast<double*>(base + 5 * rr);
  int* dstat = reinterpret_cast<int*>(h->small2.as<char>());
  FC_CUDA(cudaMemcpyAsync(dS, Sq.data(), rr * sizeof(zd), cudaMemcpyHostToDevice, h->stream));
  if (Aq) FC_CUDA(cudaMemcpyAsync(dB, Aq->data(), rr * sizeof(zd), cudaMemcpyHostToDevice, h->stream));
  ReducedEigArgs<double> a;
  a.r = r; a.generalized = Aq ? 1 : 0; a.S = dS; a.Bq = dB; a.V = dV; a.W = dW; a.lam = dlam; a.Xout = dX;
  a.status = dstat; a.sweeps = dstat + 1;
  k_reduced_eig<double><<<1, 512, 0, h->stream>>>(a);
  check_launch(h);
  lam.resize(r);
  V.resize(rr);
  int st[2] = {0, 0};
  FC_CUDA(cudaMemcpyAsync(lam.data(), dlam, r * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  FC_CUDA(cudaMemcpyAsync(V.data(), dX, rr * sizeof(zd), cudaMemcpyDeviceToHost, h->stream));
  FC_CUDA(cudaMemcpyAsync(st, dstat, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  sync(h);
  *status = st[0];
  h->stats.jacobi_sweeps += st[1];
}

static void hermitian_part(std::vector<zc>& M, int r) {  // _feast_hermitian_part!, core/feast_aux.jl:84-92
  for (int i = 0; i < r; ++i)
    for (int j = i; j < r; ++j) {
      const zc v = 0.5 * (M[(size_t)i * r + j] + std::conj(M[(size_t)j * r + i]));
      M[(size_t)i * r + j] = v;
      M[(size_t)j * r + i] = std::conj(v);
    }
}

// =====================================================================================================
// NCCL (resolved at run time; the single-GPU path has no NCCL dependency)
// =====================================================================================================
struct NcclId { char internal[128]; };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static void nccl_load() {
  if (g_nccl.lib) return;
  const char* env = getenv("FEASTCUDA_NCCL_LIB");
  const char* cands[] = {env, "libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; i < 3 && !g_nccl.lib; ++i)
    if (cands[i]) g_nccl.lib = dlopen(cands[i], RTLD_NOW | RTLD_GLOBAL);
  if (!g_nccl.lib) throw FcError(FEASTCUDA_ERR_NCCL, "libnccl.so.2 not found (set FEASTCUDA_NCCL_LIB)");
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.Broadcast = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclBroadcast");
  g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclAllGather");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
    throw FcError(FEASTCUDA_ERR_NCCL, "libnccl: missing symbols");
}
#define FC_NCCL(call)                                                                                          \
  do {                                                                                                         \
    int r_ = (call);                                                                                           \
    if (r_ != 0)                                                                                               \
      throw FcError(FEASTCUDA_ERR_NCCL, std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error")); \
  } while (0)

static void allreduce_block(H* h, zd* buf, int64_t count_zd) {
  if (h->nranks <= 1) return;
  Timer t;
  FC_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count_zd * 2, /*ncclDouble*/ 8, /*ncclSum*/ 0, h->nccl_comm, h->stream));
  sync(h);
  h->stats.ms_allreduce += t.ms();
  h->stats.allreduce_bytes += count_zd * (int64_t)sizeof(zd);
}

// =====================================================================================================
// Row-sharded runs: the arena (all block slots + the mailbox of the one-shot reductions) and its peer mappings
// =====================================================================================================
static void allreduce_small(H* h, double* dev, size_t count) {   // sum over ranks, in place (device buffer)
  if (h->nranks <= 1) return;
  FC_NCCL(g_nccl.AllReduce(dev, dev, count, /*ncclDouble*/ 8, /*ncclSum*/ 0, h->nccl_comm, h->stream));
  h->stats.allreduce_bytes += (int64_t)(count * sizeof(double));
}

static void nccl_barrier(H* h) {
  if (h->nranks <= 1 || !h->nccl_comm) return;
  double* d = h->small2.as<double>();
  if (!d) return;
  g_nccl.AllReduce(d, d, 1, 8, 0, h->nccl_comm, h->stream);
  cudaStreamSynchronize(h->stream);
}

static void arena_release(H* h) {
  if (!h->arena) return;
  cudaStreamSynchronize(h->stream);
  nccl_barrier(h);   // nobody unmaps while a peer may still be reading its rows
  for (int s = 0; s < BS_COUNT; ++s) { h->blk[s].p = nullptr; h->blk[s].cap = 0; h->blk[s].owned = true; }
  peer_arena_destroy(h->parena);
  for (int p = 0; p < 16; ++p) h->peer_arena[p] = nullptr;
  h->arena = nullptr;
  h->arena_bytes = 0;
}

static void arena_allocate(H* h, int ld) {
  FC_REQUIRE(h->nranks >= 1 && h->nranks <= LZ_MAXRANKS, "row sharding: at most 16 ranks");
  h->small2.ensure((size_t)4 * FC_MAXCOLS * sizeof(zd) + 64);
  arena_release(h);
  for (int s = 0; s < BS_COUNT; ++s) h->blk[s].release();
  const size_t slot = (((size_t)h->nloc_max * (size_t)ld * sizeof(zd)) + 255) & ~(size_t)255;
  h->arena_slot_bytes = slot;
  h->arena_mbox_off = (size_t)BS_COUNT * slot;
  h->arena_bytes = h->arena_mbox_off + ((sizeof(LzMailbox) + 255) & ~(size_t)255);
  std::string err = peer_arena_create(h->parena, h->device, h->nranks, h->rank, h->ipc_token, h->arena_bytes, [&]() { nccl_barrier(h); });
  if (h->nranks > 1) {   // all or nothing
    double fl = err.empty() ? 0.0 : 1.0;
    double* d = h->small2.as<double>();
    FC_CUDA(cudaMemcpyAsync(d, &fl, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    FC_NCCL(g_nccl.AllReduce(d, d, 1, 8, 0, h->nccl_comm, h->stream));
    FC_CUDA(cudaMemcpyAsync(&fl, d, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    sync(h);
    if (fl > 0 && err.empty()) { peer_arena_destroy(h->parena); err = "a peer rank could not map the shared arena"; }
  }
  if (!err.empty()) throw FcError(FEASTCUDA_ERR_UNSUPPORTED, "row sharding: " + err);
  // distances between rows must fit the kernels' signed 32-bit offsets (16-byte units)
  FC_REQUIRE((double)h->parena.stride * (double)h->nranks / 16.0 < 2147483647.0, "row sharding: the ranks' arenas span more than 32 GB");
  for (int p = 0; p < h->nranks; ++p) h->peer_arena[p] = (void*)(uintptr_t)(h->parena.base + (size_t)p * h->parena.stride);
  h->arena = h->peer_arena[h->rank];
  FC_CUDA(cudaMemsetAsync(h->arena, 0, h->arena_bytes, h->stream));
  for (int s = 0; s < BS_COUNT; ++s) {
    h->blk[s].p = (char*)h->arena + (size_t)s * slot;
    h->blk[s].cap = slot;
    h->blk[s].owned = false;
  }
  for (int sl = 0; sl < 4; ++sl) h->goff_rowbytes4[sl] = 0;
  h->xseq = 0;
  sync(h);
  nccl_barrier(h);   // every rank's mailbox is zeroed before anyone writes into it
}

static void fill_xchg(H* h, LzXchg& x) {
  memset(&x, 0, sizeof(x));
  x.nranks = 1;
  if (!h->row_sharded || h->nranks <= 1) return;
  x.nranks = h->nranks;
  x.rank = h->rank;
  x.seq = ++h->xseq;
  for (int p = 0; p < h->nranks; ++p) x.mbox[p] = reinterpret_cast<LzMailbox*>((char*)h->peer_arena[p] + h->arena_mbox_off);
}

// every rank has finished its previous kernels on the block slots (before a kernel gathers rows the peers just wrote)
static void xbarrier(H* h) {
  if (!h->row_sharded || h->nranks <= 1) return;
  LzXchg x;
  fill_xchg(h, x);
  k_lz_barrier<<<1, 128, 0, h->stream>>>(x);
  check_launch(h);
}

// gather offsets of A's local rows for blocks whose rows are `rowbytes` apart (two strides are in use: the compact Lanczos blocks
// and the engine's complex blocks viewed as interleaved real columns)
static const int* resolve_goff(H* h, int64_t rowbytes) {
  const size_t nnz1 = (size_t)std::max<int64_t>(h->nnz_loc, 1);
  for (int sl = 0; sl < 4; ++sl)
    if (h->goff_rowbytes4[sl] == rowbytes) return h->goff.as<int>() + (size_t)sl * nnz1;
  FC_REQUIRE(rowbytes % 16 == 0, "row sharding: rows must be multiples of 16 bytes");
  const int use = h->goff_next;
  h->goff_next = (h->goff_next + 1) & 3;
  h->goff.ensure((size_t)4 * nnz1 * sizeof(int));
  int* out = h->goff.as<int>() + (size_t)use * nnz1;
  LzArenas ar;
  memset(&ar, 0, sizeof(ar));
  for (int p = 0; p < h->nranks; ++p) ar.base[p] = (const char*)h->peer_arena[p];
  k_lz_resolve<<<std::max(1, h->sms * 4), 256, 0, h->stream>>>(h->nnz_loc, h->dA.col.as<int>(), (long long)rowbytes, h->rank, ar, out);
  check_launch(h);
  h->goff_rowbytes4[use] = rowbytes;
  return out;
}

// =====================================================================================================
// work sharding across ranks
// =====================================================================================================
struct WorkItem { int node, c0, nc; };

static std::vector<WorkItem> build_items(int ne, int active, int nranks, int rank, int shard, const std::vector<double>& cost) {
  std::vector<WorkItem> items;
  if (nranks <= 1) {
    for (int e = 0; e < ne; ++e) items.push_back({e, 0, active});
    return items;
  }
  if (shard == FEASTCUDA_SHARD_NODES) {
    int64_t s, c;
    host_node_partition(ne, nranks, rank, &s, &c);
    for (int64_t e = s; e < s + c; ++e) items.push_back({(int)e, 0, active});
    return items;
  }
  if (shard == FEASTCUDA_SHARD_COLUMNS) {
    const int base = active / nranks, rem = active % nranks;
    const int c0 = rank * base + std::min(rank, rem), nc = base + (rank < rem ? 1 : 0);
    if (nc > 0)
      for (int e = 0; e < ne; ++e) items.push_back({e, c0, nc});
    return items;
  }
  // balanced: the (node, column) cells, node-major, are cut into nranks contiguous runs of equal cost
  double total = 0.0;
  for (int e = 0; e < ne; ++e) total += cost[e] * active;
  const double lo = total * rank / nranks, hi = total * (rank + 1) / nranks;
  double acc = 0.0;
  for (int e = 0; e < ne; ++e) {
    const double ce = cost[e];
    const double n0 = acc, n1 = acc + ce * active;
    if (n1 > lo && n0 < hi && ce > 0) {
      int c0 = (int)std::llround(std::max(0.0, (lo - n0) / ce));
      int c1 = (int)std::llround(std::min((double)active, (hi - n0) / ce));
      c0 = std::max(0, std::min(active, c0));
      c1 = std::max(c0, std::min(active, c1));
      if (c1 > c0) items.push_back({e, c0, c1 - c0});
    }
    acc = n1;
  }
  return items;
}

// =====================================================================================================
// per-node block solve, any operator kind
// =====================================================================================================
static bool node_solve(H* h, int node, zc z, int m, const zd* RHS, zd* X, bool use_x0, const feastcuda_solver_opts& o, double tol,
                       SolveOut& so) {
  if (h->kind == OP_SPARSE) {
    block_bicgstab(h, z, m, RHS, X, use_x0, o, tol, so);
    bool all_ok = true;
    for (int c = 0; c < m; ++c) all_ok = all_ok && so.ok[c];
    return all_ok;
  }
  if (h->kind == OP_DENSE) return dense_node_solve(h, node, z, m, RHS, X);
  if (h->kind == OP_BAND) return band_node_solve(h, node, z, m, RHS, X);
  if (h->kind == OP_MATFREE)
    throw FcError(FEASTCUDA_ERR_UNSUPPORTED, "matrix-free operators are served by the multi-shift Lanczos filter only (real basis, true filter)");
  throw FcError(FEASTCUDA_ERR_STATE, "no operator set");
}

static void apply_op(H* h, int which, int m, const zd* X, zd* Y) {
  if (h->row_sharded) {
    FC_REQUIRE(which == FEASTCUDA_A, "row sharding: B = I");
    sharded_apply(h, m, X, Y, nullptr, nullptr);
    return;
  }
  if (h->kind == OP_SPARSE) {
    if (which == FEASTCUDA_A) launch_spmm<SPMM_PLAIN>(h, op_A(), m, X, Y, nullptr, nullptr);
    else launch_spmm<SPMM_PLAIN>(h, op_B(), m, X, Y, nullptr, nullptr);
  } else if (h->kind == OP_DENSE) dense_apply(h, which, m, X, Y);
  else if (h->kind == OP_BAND) band_apply(h, which, m, X, Y);
  else if (h->kind == OP_MATFREE) {
    // real operator on a complex block: the interleaved (re, im) columns are 2m real columns of a row-major block
    if (which != FEASTCUDA_A) throw FcError(FEASTCUDA_ERR_UNSUPPORTED, "matrix-free problems are standard (B = I)");
    h->mf_apply(h->mf_ctx, h->n, 2 * (int64_t)m, reinterpret_cast<const double*>(X), 2 * h->ws_ld, reinterpret_cast<double*>(Y), 2 * h->ws_ld,
                (void*)h->stream);
    h->stats.kernel_launches++;
  } else throw FcError(FEASTCUDA_ERR_STATE, "no operator set");
}

// Row-sharded runs: Y = A X (theta == nullptr) or R = A X - X diag(theta) with ||R_j||^2 in `norms2`, on the engine's complex blocks
// viewed as 2m interleaved REAL columns (A is real), by the Lanczos gather kernel -- halo rows come from the peers' HBM
static void sharded_apply(H* h, int m, const zd* X, zd* Y, const double* theta, std::vector<double>* norms2) {
  FC_REQUIRE(h->kind == OP_SPARSE && !h->dev_complex && !h->has_b, "row sharding: standard real symmetric sparse problems only");
  xbarrier(h);   // the peers have written the rows this launch gathers
  const int64_t ldr = 2 * (int64_t)h->ws_ld;
  double* part = h->partial_r.as<double>();
  double* d_theta = reinterpret_cast<double*>(h->small2.as<zd>() + 8);
  if (norms2) norms2->assign(m, 0.0);
  for (int c0 = 0; c0 < m; c0 += 64) {     // at most 128 real columns per launch
    const int mc = std::min(64, m - c0);
    LzArgs a;
    memset(&a, 0, sizeof(a));
    a.n = h->n; a.m = 2 * mc; a.ld = ldr;
    a.ptr = h->dA.ptr.as<int>(); a.col = h->dA.col.as<int>(); a.val = h->dA.val.p;
    a.goff = resolve_goff(h, ldr * (int64_t)sizeof(double));
    a.U = reinterpret_cast<const double*>(X + c0); a.prev = a.U;
    a.out = reinterpret_cast<double*>(Y + c0);
    a.partial = part; a.pstride = FC_MAXCOLS; a.tile_rows = h->lz_tile_rows;
    int g = 0;
    if (theta == nullptr) lz_launch<LZ_PLAIN, false>(h, a, &g);
    else {
      std::vector<double> th2(2 * mc);
      for (int c = 0; c < mc; ++c) th2[2 * c] = th2[2 * c + 1] = theta[c0 + c];
      FC_CUDA(cudaMemcpyAsync(d_theta, th2.data(), th2.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      a.s_theta = d_theta;
      lz_launch<LZ_RES, false>(h, a, &g);
      double* dev = h->red_ws.as<double>();
      k_reduce_partials<double><<<1, 1024, ((size_t)FC_MAXCOLS + 1024) * sizeof(double), h->stream>>>(part, 1, g, FC_MAXCOLS, 2 * mc, dev);
      check_launch(h);
      allreduce_small(h, dev, (size_t)2 * mc);
      std::vector<double> tmp(2 * mc);
      FC_CUDA(cudaMemcpyAsync(tmp.data(), dev, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      sync(h);   // th2 / tmp are host buffers
      for (int c = 0; c < mc; ++c) (*norms2)[c0 + c] = tmp[2 * c] + tmp[2 * c + 1];
    }
  }
}

// res_j = ||A x_j - lam_j B x_j||_2 (not yet divided by max(|lam|,1))
static void eig_residual_norms(H* h, int m, const zd* X, const std::vector<zc>& lam, std::vector<double>& out) {
  out.assign(m, 0.0);
  if (m == 0) return;
  if (h->row_sharded) {
    std::vector<double> th(m), n2;
    for (int c = 0; c < m; ++c) th[c] = lam[c].real();
    sharded_apply(h, m, X, blk(h, BS_KT), th.data(), &n2);
    for (int c = 0; c < m; ++c) out[c] = std::sqrt(n2[c]);
    return;
  }
  if (h->kind == OP_SPARSE) {
    zd* dl = h->small2.as<zd>() + 8;
    FC_CUDA(cudaMemcpyAsync(dl, lam.data(), (size_t)m * sizeof(zd), cudaMemcpyHostToDevice, h->stream));
    OpDesc d = op_shifted(zc(0.0));
    const int g = launch_spmm<SPMM_EIGRES>(h, d, m, X, nullptr, nullptr, dl);
    std::vector<zd> tmp(m);
    reduce_to_host<zd>(h, h->partial.as<zd>(), 1, g, m, tmp.data());
    for (int c = 0; c < m; ++c) out[c] = std::sqrt(tmp[c].x);
  } else {
    // dense / band: R = A X - (B X) diag(lam) with the generic kernels
    zd *ax = blk(h, BS_KS), *bx = blk(h, BS_KT);
    apply_op(h, FEASTCUDA_A, m, X, ax);
    const zd* bxp = X;
    if (h->has_b) { apply_op(h, FEASTCUDA_B, m, X, bx); bxp = bx; }
    std::vector<zc> f(m);
    for (int c = 0; c < m; ++c) f[c] = -lam[c];
    zd* tmpb = blk(h, BS_KV);
    scale_cols(h, m, f, bxp, tmpb);
    axpby_cols(h, m, zc(1.0), 1.0, ax, tmpb);
    col_norms(h, m, tmpb, out);
  }
}

// dense and banded operators: every node of the sweep goes through the batched LU / substitution kernels together
static bool dense_items_batched(H* h, std::vector<WorkItem>& items, int ne, int active, const zc* Zne, const zc* Wne, double wfac,
                                const zd* rhs, bool* failed) {
  if ((h->kind != OP_DENSE && h->kind != OP_BAND) || items.empty()) return false;
  for (size_t q = 0; q < items.size(); ++q)
    if (items[q].c0 != 0 || items[q].nc != active || items[q].node != items[0].node + (int)q) return false;
  if (h->kind == OP_BAND && !band_batch_fits(h, (int)items.size())) return false;
  std::vector<zc> zs;
  for (const WorkItem& it : items) zs.push_back(Zne[it.node]);
  zd* Xp = nullptr;
  int64_t xb = 0;
  const bool ok = (h->kind == OP_DENSE) ? dense_batch_solve(h, ne, items[0].node, (int)items.size(), zs.data(), active, rhs, &Xp, &xb)
                                        : band_batch_solve(h, items[0].node, (int)items.size(), zs.data(), active, rhs, &Xp, &xb);
  h->stats.node_solves += (int64_t)items.size();
  if (!ok) *failed = true;
  else
    for (size_t q = 0; q < items.size(); ++q) axpby_cols(h, active, wfac * Wne[items[q].node], 1.0, Xp + (int64_t)q * xb, blk(h, BS_ACC));
  items.clear();
  return true;
}

// =====================================================================================================
// H-RR refinement loop (dense/feast_dense.jl:78-351, sparse/feast_sparse.jl:246-499, banded:561-823)
// =====================================================================================================
static feastcuda_solver_opts default_opts() {
  feastcuda_solver_opts o;
  memset(&o, 0, sizeof(o));
  o.solver = FEASTCUDA_SOLVER_BICGSTAB;
  o.maxiter = 500;
  o.restart = 3;
  o.check_every = 8;
  o.filter = FEASTCUDA_FILTER_REFERENCE;
  o.shard = FEASTCUDA_SHARD_NODES;
  return o;
}

static bool pencil_is_real(H* h) {
  if (h->kind == OP_SPARSE) return !h->dev_complex;
  if (h->kind == OP_DENSE) return !h->denseA.cplx && !(h->has_b && h->denseB.cplx);
  if (h->kind == OP_BAND) return !h->bandA.cplx && !(h->has_b && h->bandB.cplx);
  if (h->kind == OP_MATFREE) return true;
  return false;
}

static void prepare_operator(H* h) {
  if (h->kind == OP_SPARSE) finalize_sparse(h);
  else if (h->kind == OP_DENSE) dense_prepare(h);
  else if (h->kind == OP_BAND) band_prepare(h);
  else if (h->kind == OP_MATFREE) FC_REQUIRE(h->mf_apply != nullptr && !h->has_b, "matrix-free operator: callback missing or B set");
  else throw FcError(FEASTCUDA_ERR_STATE, "no operator set");
}

static void run_interval(H* h, double Emin, double Emax, int m0, int64_t* fpm, const zc* Zne, const zc* Wne, int ne,
                         const feastcuda_solver_opts* optsp, int64_t* Mout, int64_t* info, double* epsout, int64_t* loopout,
                         bool subspace_is_real) {
  Timer ttotal;
  if (host_feastdefault(fpm)) throw FcError(FEASTCUDA_ERR_ARG, "invalid fpm (feastdefault!)");
  prepare_operator(h);
  const int64_t n = h->n;
  // check_feast_srci_input, core/feast_aux.jl:369-399 (the shim throws ArgumentError before the ccall)
  FC_REQUIRE(n > 0, "Matrix size N must be positive");
  FC_REQUIRE(m0 > 0 && m0 <= (h->row_sharded ? h->n_glob : n), "Number of eigenvalues M0 must be between 1 and N");
  FC_REQUIRE(Emin < Emax, "Search interval [Emin, Emax] must be valid");
  FC_REQUIRE(ne >= 1 && ne <= 128 && Zne && Wne, "contour required");
  FC_REQUIRE(h->have_subspace && h->sub_m0 == m0 && h->ws_n == n, "initial subspace not uploaded for this (n, M0)");
  feastcuda_solver_opts o = optsp ? *optsp : default_opts();
  const double tol = (o.tol == 0.0) ? std::pow(10.0, -(double)fpm[2]) : o.tol;
  const double eps_tol = std::max(host_feast_tolerance(fpm), o.eps_floor > 0 ? o.eps_floor : 0.0);
  const int maxloop = (int)fpm[3];
  const bool real_mode = (o.filter == FEASTCUDA_FILTER_TRUE) && pencil_is_real(h) && subspace_is_real;
  const bool matfree = (h->kind == OP_MATFREE);
  const bool iterative = (h->kind == OP_SPARSE) || matfree;
  int shard = o.shard;
  if (!iterative) shard = FEASTCUDA_SHARD_NODES;

  int active = m0, rank = 0, M = 0, M_found = 0, loop_count = 0;
  int info_code = 0;
  double eps_val = INFINITY;
  std::vector<double> lam(m0, 0.0), res(m0, 0.0);
  std::vector<double> cost(ne, 1.0);
  bool have_ritz = false;
  // fpm[42] ("single-precision solver"): FP32 Lanczos vectors until a sweep stops contracting the residual, then FP64
  bool use_fp32 = o.mixed == 1 || (o.mixed == 2 && fpm[41] == 1);
  double eps_before_sweep = INFINITY;
  int qb = BS_QB, xr = BS_XR;  // current basis / Ritz-vector slots (swapped every loop)
  int res_slot = -1;           // slot holding the Ritz vectors that M_found refers to
  std::vector<zc> Sq, Aq, V;
  std::vector<double> lam_red;

  for (int loop_idx = 0; loop_idx <= maxloop; ++loop_idx) {
    loop_count = loop_idx;
    h->stats.loops++;
    Timer tsolve;
    zd* basis = blk(h, qb);
    const zd* rhs = basis;
    // generalized Hermitian pencils with a positive definite B: the same recurrence in the B-inner product (msl_filter_gen)
    const bool use_gen = iterative && !matfree && h->has_b && o.solver == FEASTCUDA_SOLVER_MSLANCZOS && o.filter == FEASTCUDA_FILTER_TRUE &&
                         h->hA.structure != FEASTCUDA_GEN && (real_mode || h->dev_complex) && cheb_prepare(h);
    if (h->has_b && !use_gen) { apply_op(h, FEASTCUDA_B, active, basis, blk(h, BS_RHS)); rhs = blk(h, BS_RHS); }
    zero_cols(h, active, blk(h, BS_ACC));
    // multi-shift Lanczos: standard Hermitian problems with the true filter rho = Re g -- real symmetric with a real basis (real
    // arithmetic) or complex Hermitian (complex vectors, real tridiagonal: the coefficients of rho are real either way)
    const bool cplx_msl = iterative && !matfree && h->dev_complex && !h->has_b && o.filter == FEASTCUDA_FILTER_TRUE && h->hA.structure != FEASTCUDA_GEN;
    const bool use_msl = use_gen || (iterative && (o.solver == FEASTCUDA_SOLVER_MSLANCZOS || matfree) && !h->has_b && (real_mode || cplx_msl));
    if (matfree && !use_msl)
      throw FcError(FEASTCUDA_ERR_UNSUPPORTED, "matrix-free operators need a real initial subspace and filter = TRUE (multi-shift Lanczos)");
    if (h->row_sharded && !(use_msl && real_mode && !use_gen))
      throw FcError(FEASTCUDA_ERR_UNSUPPORTED, "row sharding serves the real symmetric multi-shift Lanczos filter (real basis, filter = TRUE)");
    std::vector<WorkItem> items;
    if (!use_msl) items = build_items(ne, active, h->nranks, h->rank, shard, cost);
    std::vector<double> node_cost(ne, 0.0), node_cols(ne, 0.0);
    bool failed = false;
    if (use_msl) {
      // one real Lanczos recurrence per column serves every node; ranks own contiguous column-pair slices
      const bool cplx_vec = use_gen ? h->dev_complex : cplx_msl;
      const int epc = cplx_vec ? 1 : 2;                     // columns per 16-byte element
      const int npairs = (active + epc - 1) / epc;
      const int pb = npairs / h->nranks, pr = npairs % h->nranks;
      const int p0 = h->rank * pb + std::min(h->rank, pr), pn = pb + (h->rank < pr ? 1 : 0);
      int c0 = epc * p0, nc = std::max(0, std::min(active, epc * (p0 + pn)) - c0);
      if (h->row_sharded) { c0 = 0; nc = active; }     // every rank filters all columns of its own rows
      const bool first = !(o.ritz_guess && have_ritz);
      double target = (first && o.inner_rel0 > 0) ? o.inner_rel0 : o.inner_rel;
      if (!(target > 0)) target = tol;
      if (o.adaptive && !first && std::isfinite(eps_val) && eps_val > 0) {
        // the sweep contracts the eigen-residual by roughly its inner target: when the tolerance is within reach, aim at it
        const double t = 2.0 * eps_tol / eps_val;   // measured: a sweep contracts the residual by 0.04..0.25 x its target
        if (t >= 1e-6) target = std::min(0.1, t);
      }
      const int kmax = (first && o.maxiter0 > 0) ? o.maxiter0 : o.maxiter;
      if (nc > 0) {
        MslOut mo;
        if (use_gen) {
          // accuracy of the Chebyshev inner solves with B: the filter of the neighbouring pencil (A, P^-1) only has to stay below the
          // sweep's own contraction
          const double d0 = o.b_delta > 0 ? o.b_delta : 1e-4;
          const double dfac = getenv("FEASTCUDA_BDELTA_FACTOR") ? atof(getenv("FEASTCUDA_BDELTA_FACTOR")) : 1e-3;
          const double delta = first ? d0 : std::min(d0, std::max(1e-10, dfac * target));
          if (h->dev_complex) msl_filter_gen<true>(h, qb, c0, nc, !first, lam.data() + c0, Zne, Wne, ne, target, kmax, o.check_every > 0 ? o.check_every : 16, delta, mo);
          else msl_filter_gen<false>(h, qb, c0, nc, !first, lam.data() + c0, Zne, Wne, ne, target, kmax, o.check_every > 0 ? o.check_every : 16, delta, mo);
        } else if (cplx_msl) msl_filter<true>(h, qb, c0, nc, !first, lam.data() + c0, Zne, Wne, ne, target, kmax, o.check_every > 0 ? o.check_every : 16, mo);
        else msl_filter<false>(h, qb, c0, nc, !first, lam.data() + c0, Zne, Wne, ne, target, kmax, o.check_every > 0 ? o.check_every : 16, mo,
                               use_fp32);
        h->stats.node_solves += ne;
        for (int e = 0; e < ne && e < 128; ++e) h->stats.node_iters[e] = mo.k;
        if (getenv("FEASTCUDA_VERBOSE"))
          fprintf(stderr, "[feastcuda r%d] loop %d: lanczos k=%d maxres=%.3e converged=%d cols=[%d,%d) target=%.1e fp32=%d\n", h->rank,
                  loop_idx, mo.k, mo.maxres, (int)mo.converged, c0, c0 + nc, target, (int)use_fp32);
      }
    }
    if (dense_items_batched(h, items, ne, active, Zne, Wne, 2.0, rhs, &failed) && failed) info_code = 8;
    for (const WorkItem& it : items) {
      const zc z = Zne[it.node];
      zd* X = blk(h, BS_KX) + it.c0;
      bool use_x0 = false;
      if (iterative && o.ritz_guess && have_ritz) {
        std::vector<zc> f(it.nc);
        for (int c = 0; c < it.nc; ++c) f[c] = zc(1.0) / (z - lam[it.c0 + c]);
        scale_cols(h, it.nc, f, basis + it.c0, X);
        use_x0 = true;
      }
      SolveOut so;
      feastcuda_solver_opts on = o;
      if (!use_x0) {  // no Ritz information yet: the first sweep may get its own accuracy / budget
        if (o.inner_rel0 > 0) on.inner_rel = o.inner_rel0;
        if (o.maxiter0 > 0) on.maxiter = o.maxiter0;
      }
      const bool ok = node_solve(h, it.node, z, it.nc, rhs + it.c0, X, use_x0, on, tol, so);
      h->stats.node_solves++;
      if (iterative) {
        h->stats.node_iters[it.node] = so.it_total;
        node_cost[it.node] += (double)so.it_total * it.nc;
        node_cols[it.node] += it.nc;
      }
      if (!ok && (!iterative || o.inner_rel <= 0.0)) { failed = true; info_code = iterative ? 5 : 8; break; }
      axpby_cols(h, it.nc, 2.0 * Wne[it.node], 1.0, X, blk(h, BS_ACC) + it.c0);
    }
    sync(h);
    drain_events(h);    // samples of the direct (band) kernels
    h->stats.ms_solve += tsolve.ms();
    if (h->nranks > 1) {
      // one exchange per refinement loop: MPI.Allreduce(Q_proj), parallel/feast_mpi.jl:119,341,858
      int64_t fl[1] = {failed ? 1 : 0};
      if (!h->row_sharded) allreduce_block(h, blk(h, BS_ACC), (int64_t)n * h->ws_ld);   // row-sharded: the local rows are complete
      // failure flag + measured node costs ride along in a tiny second reduction
      std::vector<double> pack(2 * ne + 1, 0.0);
      for (int e = 0; e < ne; ++e) { pack[e] = node_cost[e]; pack[ne + e] = node_cols[e]; }
      pack[2 * ne] = (double)fl[0];
      double* dp = reinterpret_cast<double*>(h->small2.as<zd>() + 2 * FC_MAXCOLS);
      FC_CUDA(cudaMemcpyAsync(dp, pack.data(), pack.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      FC_NCCL(g_nccl.AllReduce(dp, dp, pack.size(), 8, 0, h->nccl_comm, h->stream));
      FC_CUDA(cudaMemcpyAsync(pack.data(), dp, pack.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      sync(h);
      for (int e = 0; e < ne; ++e) { node_cost[e] = pack[e]; node_cols[e] = pack[ne + e]; }
      if (pack[2 * ne] > 0) { failed = true; if (!info_code) info_code = iterative ? 5 : 8; }
    }
    if (failed) break;
    if (iterative)
      for (int e = 0; e < ne; ++e) cost[e] = node_cols[e] > 0 ? std::max(1.0, node_cost[e] / node_cols[e]) : 1.0;

    if (real_mode) {
      const int mp = pow2_ge(active);
      k_zero_imag<double><<<ew_grid(h, n, mp), 256, 0, h->stream>>>(n, active, mp, h->ws_ld, blk(h, BS_ACC));
      check_launch(h);
    }
    Timer tortho;
    int qslot = BS_ACC;
    rank = orthonormalize(h, active, BS_ACC, BS_KP, std::sqrt(2.220446049250313e-16), &qslot);
    h->stats.ms_ortho += tortho.ms();
    if (rank == 0) { info_code = 5; break; }
    Timer tproj;
    zd* Q = blk(h, qslot);
    apply_op(h, FEASTCUDA_A, rank, Q, blk(h, BS_KS));
    gram_host(h, rank, rank, Q, blk(h, BS_KS), Sq);
    hermitian_part(Sq, rank);
    if (h->has_b) {
      apply_op(h, FEASTCUDA_B, rank, Q, blk(h, BS_KT));
      gram_host(h, rank, rank, Q, blk(h, BS_KT), Aq);
      hermitian_part(Aq, rank);
    }
    h->stats.ms_project += tproj.ms();
    Timer teig;
    int est = 0;
    reduced_eig(h, rank, Sq, h->has_b ? &Aq : nullptr, lam_red, V, &est);
    h->stats.ms_eig += teig.ms();
    if (getenv("FEASTCUDA_VERBOSE"))
      fprintf(stderr, "[feastcuda r%d] loop %d: rank=%d eig status=%d lam_red[0..2]=%.6e %.6e %.6e\n", h->rank, loop_idx, rank, est,
              lam_red[0], lam_red[rank > 1 ? 1 : 0], lam_red[rank > 2 ? 2 : 0]);
    if (est != 0) {
      // the reduced overlap matrix is not numerically positive definite (status 1) or the Jacobi sweeps did not converge (status 2):
      // the reference falls back to the general eigen(Sq, Aq) and keeps the real parts (dense/feast_dense.jl:274-283)
      std::vector<zc> lc, Vc;
      bool ok = h->has_b ? host_pencil_eig(rank, Sq, Aq, lc, Vc) : host_complex_eig(rank, Sq, lc, Vc);
      for (int i = 0; ok && i < rank; ++i) ok = std::isfinite(lc[i].real());
      if (!ok) { info_code = 8; break; }
      std::vector<int> ord(rank);
      for (int i = 0; i < rank; ++i) ord[i] = i;
      std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return lc[a].real() < lc[b].real(); });
      lam_red.resize(rank);
      V.assign((size_t)rank * rank, zc(0.0));
      for (int k = 0; k < rank; ++k) {
        lam_red[k] = lc[ord[k]].real();
        for (int i = 0; i < rank; ++i) V[(size_t)i * rank + k] = Vc[(size_t)i * rank + ord[k]];
      }
    }
    // _feast_reorder_by_interval! (stable partition, core/feast_aux.jl:144-197) folded into V's column order;
    // unit 2-norm of the M inside Ritz vectors (dense/feast_dense.jl:301-305): ||Q v|| = ||v|| for orthonormal Q
    std::vector<int> perm;
    for (int i = 0; i < rank; ++i) if (Emin <= lam_red[i] && lam_red[i] <= Emax) perm.push_back(i);
    M = (int)perm.size();
    for (int i = 0; i < rank; ++i) if (!(Emin <= lam_red[i] && lam_red[i] <= Emax)) perm.push_back(i);
    const bool keep_going = o.keep_going && iterative && o.inner_rel > 0 && loop_idx < maxloop;
    if (M == 0 && !keep_going) { info_code = 5; break; }
    std::vector<zc> T((size_t)rank * rank);
    for (int k = 0; k < rank; ++k) {
      const int srcc = perm[k];
      double nrm = 1.0;
      if (k < M) {
        double s2 = 0.0;
        for (int i = 0; i < rank; ++i) s2 += std::norm(V[(size_t)i * rank + srcc]);
        nrm = s2 > 0 ? std::sqrt(s2) : 1.0;
      }
      for (int i = 0; i < rank; ++i) T[(size_t)i * rank + k] = V[(size_t)i * rank + srcc] / nrm;
      lam[k] = lam_red[srcc];
    }
    Timer tres;
    rowtransform(h, rank, rank, Q, T, blk(h, xr));
    res_slot = xr;
    std::vector<zc> lamc(M);
    for (int j = 0; j < M; ++j) lamc[j] = zc(lam[j], 0.0);
    std::vector<double> rnorm;
    eig_residual_norms(h, M, blk(h, xr), lamc, rnorm);
    double max_res = 0.0;
    for (int j = 0; j < M; ++j) {
      res[j] = rnorm[j] / std::max(std::fabs(lam[j]), 1.0);
      max_res = std::max(max_res, res[j]);
    }
    h->stats.ms_resid += tres.ms();
    eps_val = M > 0 ? max_res : INFINITY;
    M_found = M;
    // a refined FP32 sweep that gained less than a factor 4 has hit single precision's attainable accuracy for this spectrum
    if (use_fp32 && have_ritz && std::isfinite(eps_before_sweep) && !(eps_val <= 0.25 * eps_before_sweep)) use_fp32 = false;
    eps_before_sweep = eps_val;
    if (getenv("FEASTCUDA_VERBOSE"))
      fprintf(stderr, "[feastcuda r%d] loop %d: M=%d rank=%d epsout=%.3e items=%zu iters(last)=%d\n", h->rank, loop_idx, M, rank,
              eps_val, items.size(), (int)h->stats.node_iters[items.empty() ? 0 : items.back().node]);
    if (M > 0 && eps_val <= eps_tol) break;
    if (loop_idx == maxloop) { info_code = 5; break; }
    active = rank;
    std::swap(qb, xr);   // Q_basis <- Ritz vectors (dense/feast_dense.jl:336-337)
    have_ritz = true;
  }
  if (M_found == 0 && info_code == 0) info_code = 5;
  // keep the public result slot fixed: the final Ritz vectors go to BS_XR
  if (res_slot >= 0 && res_slot != BS_XR) {
    copy_cols(h, std::max(rank, 1), blk(h, res_slot), blk(h, BS_XR));
    sync(h);
  }
  h->res_m0 = m0;
  h->res_M = M_found;
  h->res_rank = rank;
  h->res_lambda.assign(lam.begin(), lam.end());
  h->res_res.assign(res.begin(), res.end());
  h->res_general = false;
  h->have_subspace = false;  // the basis slot was consumed
  *Mout = M_found;
  *info = info_code;
  *epsout = eps_val;
  *loopout = loop_count;
  h->stats.ms_total += ttotal.ms();
}

// =====================================================================================================
// General (non-Hermitian) FEAST: full contour, one-sided Rayleigh-Ritz -- feast_grci! (kernel/feast_kernel.jl:646-962)
// driven as feast_gcsrgv! / feast_gegv! / _feast_banded_general do (sparse/feast_sparse.jl:873-1006,
// dense/feast_dense.jl:402-593, banded/feast_banded.jl:1088-1284).
// Differences from the reference, both in HOW, not in WHAT converges:
//   * the filtered block is orthonormalised (rank-revealing, as in the Hermitian drivers) before the projection, so the
//     reduced pencil (Q^H A Q, Q^H B Q) is well conditioned; the reference projects on the raw block (kernel:790-807);
//   * residuals are ||A x - lambda B x|| / max(|lambda|, 1); the reference drops B there (kernel:900-906), so its
//     generalized runs can only stop at fpm[4].
// =====================================================================================================
static void run_contour(H* h, zc Emid, double r, int m0, int64_t* fpm, const zc* Zne, const zc* Wne, int ne,
                        const feastcuda_solver_opts* optsp, int64_t* Mout, int64_t* info, double* epsout, int64_t* loopout) {
  Timer ttotal;
  if (host_feastdefault(fpm)) throw FcError(FEASTCUDA_ERR_ARG, "invalid fpm (feastdefault!)");
  prepare_operator(h);
  const int64_t n = h->n;
  // check_feast_grci_input, core/feast_aux.jl:401-425
  FC_REQUIRE(n > 0, "Matrix size N must be positive");
  FC_REQUIRE(m0 > 0 && m0 <= n, "Number of eigenvalues M0 must be between 1 and N");
  FC_REQUIRE(r > 0, "Search radius r must be positive");
  FC_REQUIRE(ne >= 1 && ne <= 128 && Zne && Wne, "contour required");
  FC_REQUIRE(h->have_subspace && h->sub_m0 == m0 && h->ws_n == n, "initial subspace not uploaded for this (n, M0)");
  feastcuda_solver_opts o = optsp ? *optsp : default_opts();
  const double tol = (o.tol == 0.0) ? std::pow(10.0, -(double)fpm[2]) : o.tol;
  const double eps_tol = std::max(host_feast_tolerance(fpm), o.eps_floor > 0 ? o.eps_floor : 0.0);
  const int maxloop = (int)fpm[3];
  const bool iterative = (h->kind == OP_SPARSE);
  int shard = iterative ? o.shard : FEASTCUDA_SHARD_NODES;

  int active = m0, rank = 0, M = 0, loop = 0, info_code = 0;
  double eps_val = 0.0;
  std::vector<zc> lam(m0, zc(0.0));
  std::vector<double> res(m0, 0.0), cost(ne, 1.0);
  int qb = BS_QB, xr = BS_XR, res_slot = -1;
  std::vector<zc> Sq, Aq, V, lam_red;
  while (true) {
    h->stats.loops++;
    Timer tsolve;
    zd* basis = blk(h, qb);
    const zd* rhs = basis;
    // sparse pencils: ONE two-sided Lanczos recurrence per column serves all nodes (msl2_filter); ranks own column slices
    const int cb = active / h->nranks, cr = active % h->nranks;
    const int sl_c0 = h->rank * cb + std::min(h->rank, cr), sl_nc = cb + (h->rank < cr ? 1 : 0);
    const bool use_msl2 = iterative && o.solver == FEASTCUDA_SOLVER_MSLANCZOS && sl_nc <= TS_MAXC && gen2_prepare(h);
    if (h->has_b && !use_msl2) { apply_op(h, FEASTCUDA_B, active, basis, blk(h, BS_RHS)); rhs = blk(h, BS_RHS); }
    zero_cols(h, active, blk(h, BS_ACC));
    std::vector<WorkItem> items;
    if (!use_msl2) items = build_items(ne, active, h->nranks, h->rank, shard, cost);
    bool failed = false;
    if (use_msl2 && sl_nc > 0) {
      const bool first = !(o.ritz_guess && loop > 0);
      double target = (first && o.inner_rel0 > 0) ? o.inner_rel0 : o.inner_rel;
      if (!(target > 0)) target = tol;
      if (o.adaptive && !first && std::isfinite(eps_val) && eps_val > 0) {
        const double t = 2.0 * eps_tol / eps_val;
        if (t >= 1e-6) target = std::min(0.1, t);
      }
      MslOut mo;
      msl2_filter(h, qb, sl_c0, sl_nc, !first, lam.data() + sl_c0, Zne, Wne, ne, target, (first && o.maxiter0 > 0) ? o.maxiter0 : o.maxiter,
                  o.b_delta > 0 ? o.b_delta : 1e-10, mo);
      h->stats.node_solves += ne;
      for (int e = 0; e < ne && e < 128; ++e) h->stats.node_iters[e] = mo.k;
      if (getenv("FEASTCUDA_VERBOSE"))
        fprintf(stderr, "[feastcuda r%d] general loop %d: two-sided lanczos k=%d maxres=%.3e converged=%d cols=[%d,%d) target=%.1e\n", h->rank, loop,
                mo.k, mo.maxres, (int)mo.converged, sl_c0, sl_c0 + sl_nc, target);
    }
    if (dense_items_batched(h, items, ne, active, Zne, Wne, 1.0, rhs, &failed) && failed) info_code = 8;
    for (const WorkItem& it : items) {
      zd* X = blk(h, BS_KX) + it.c0;
      SolveOut so;
      bool use_x0 = false;
      if (iterative && o.ritz_guess && loop > 0) {
        // Ritz-pair start x0 = q/(z - theta): the solve only has to correct the eigen-residual, so inexact inner solves
        // (inner_rel) keep contracting it from sweep to sweep
        std::vector<zc> f(it.nc);
        for (int c = 0; c < it.nc; ++c) f[c] = zc(1.0) / (Zne[it.node] - lam[it.c0 + c]);
        scale_cols(h, it.nc, f, basis + it.c0, X);
        use_x0 = true;
      }
      const bool ok = node_solve(h, it.node, Zne[it.node], it.nc, rhs + it.c0, X, use_x0, o, tol, so);
      h->stats.node_solves++;
      if (iterative) h->stats.node_iters[it.node] = so.it_total;
      if (!ok && (!iterative || o.inner_rel <= 0.0)) { failed = true; info_code = iterative ? 5 : 8; break; }
      axpby_cols(h, it.nc, Wne[it.node], 1.0, X, blk(h, BS_ACC) + it.c0);   // q += w_e Y (kernel:762-766)
    }
    sync(h);
    drain_events(h);    // samples of the direct (band) kernels
    h->stats.ms_solve += tsolve.ms();
    if (h->nranks > 1) {
      allreduce_block(h, blk(h, BS_ACC), (int64_t)n * h->ws_ld);
      double fl = failed ? 1.0 : 0.0;
      double* dp = reinterpret_cast<double*>(h->small2.as<zd>() + 2 * FC_MAXCOLS);
      FC_CUDA(cudaMemcpyAsync(dp, &fl, sizeof(double), cudaMemcpyHostToDevice, h->stream));
      FC_NCCL(g_nccl.AllReduce(dp, dp, 1, 8, 0, h->nccl_comm, h->stream));
      FC_CUDA(cudaMemcpyAsync(&fl, dp, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      sync(h);
      if (fl > 0) { failed = true; if (!info_code) info_code = iterative ? 5 : 8; }
    }
    if (failed) break;
    Timer tortho;
    int qslot = BS_ACC;
    rank = orthonormalize(h, active, BS_ACC, BS_KP, std::sqrt(2.220446049250313e-16), &qslot);
    h->stats.ms_ortho += tortho.ms();
    if (rank == 0) { info_code = 5; break; }
    Timer tproj;
    zd* Q = blk(h, qslot);
    apply_op(h, FEASTCUDA_A, rank, Q, blk(h, BS_KS));
    gram_host(h, rank, rank, Q, blk(h, BS_KS), Aq);
    std::vector<zc> C = Aq;
    if (h->has_b) {
      apply_op(h, FEASTCUDA_B, rank, Q, blk(h, BS_KT));
      gram_host(h, rank, rank, Q, blk(h, BS_KT), Sq);
      if (!host_lu_solve(rank, Sq, C)) { info_code = 8; break; }
    }
    h->stats.ms_project += tproj.ms();
    Timer teig;
    if (!host_complex_eig(rank, C, lam_red, V)) { info_code = 8; break; }
    h->stats.ms_eig += teig.ms();
    // Julia's eigen orders lexicographically by (real, imag); then the stable inside-first partition (core/feast_aux.jl:208-257)
    std::vector<int> order(rank);
    for (int i = 0; i < rank; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      if (lam_red[a].real() != lam_red[b].real()) return lam_red[a].real() < lam_red[b].real();
      return lam_red[a].imag() < lam_red[b].imag();
    });
    std::vector<int> perm;
    for (int i : order) if (host_inside_gcontour(lam_red[i], Emid, r, fpm)) perm.push_back(i);
    M = (int)perm.size();
    for (int i : order) if (!host_inside_gcontour(lam_red[i], Emid, r, fpm)) perm.push_back(i);
    if (M == 0) { info_code = 5; break; }
    std::vector<zc> T((size_t)rank * rank);
    for (int k = 0; k < rank; ++k) {
      const int src = perm[k];
      double s2 = 0.0;   // ||Q v|| = ||v|| for orthonormal Q: every Ritz vector leaves with unit 2-norm (kernel:864-876)
      for (int i = 0; i < rank; ++i) s2 += std::norm(V[(size_t)i * rank + src]);
      const double nrm = s2 > 0 ? std::sqrt(s2) : 1.0;
      for (int i = 0; i < rank; ++i) T[(size_t)i * rank + k] = V[(size_t)i * rank + src] / nrm;
      lam[k] = lam_red[src];
    }
    Timer tres;
    rowtransform(h, rank, rank, Q, T, blk(h, xr));
    res_slot = xr;
    std::vector<zc> lamc(lam.begin(), lam.begin() + M);
    std::vector<double> rnorm;
    eig_residual_norms(h, M, blk(h, xr), lamc, rnorm);
    double max_res = 0.0;
    for (int j = 0; j < M; ++j) {
      res[j] = rnorm[j] / std::max(std::abs(lam[j]), 1.0);
      max_res = std::max(max_res, res[j]);
    }
    h->stats.ms_resid += tres.ms();
    eps_val = max_res;
    if (getenv("FEASTCUDA_VERBOSE"))
      fprintf(stderr, "[feastcuda r%d] general loop %d: M=%d rank=%d epsout=%.3e\n", h->rank, loop, M, rank, eps_val);
    if (eps_val <= eps_tol || loop >= maxloop) break;   // kernel:917-925 (no error code at fpm[4], like the reference)
    ++loop;
    active = rank;
    std::swap(qb, xr);
  }
  // feast_sort_general!: the M inside pairs ascending in |lambda|, stable (core/feast_tools.jl:684-713)
  if (res_slot >= 0 && M > 0 && info_code == 0) {
    std::vector<int> ord(M);
    for (int i = 0; i < M; ++i) ord[i] = i;
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return std::norm(lam[a]) < std::norm(lam[b]); });
    std::vector<zc> P((size_t)rank * rank, zc(0.0)), lam2(lam);
    std::vector<double> res2(res);
    for (int k = 0; k < rank; ++k) {
      const int src = k < M ? ord[k] : k;
      P[(size_t)src * rank + k] = 1.0;
      lam[k] = lam2[src];
      if (k < M) res[k] = res2[src];
    }
    const int other = (res_slot == BS_XR) ? BS_QB : BS_XR;
    rowtransform(h, rank, rank, blk(h, res_slot), P, blk(h, other));
    res_slot = other;
  }
  if (M == 0 && info_code == 0) info_code = 5;
  if (res_slot >= 0 && res_slot != BS_XR) {
    copy_cols(h, std::max(rank, 1), blk(h, res_slot), blk(h, BS_XR));
    sync(h);
  }
  h->res_m0 = m0;
  h->res_M = info_code == 0 ? M : 0;
  h->res_rank = rank;
  h->res_lambda.assign((size_t)2 * m0, 0.0);
  for (int j = 0; j < m0; ++j) { h->res_lambda[2 * j] = lam[j].real(); h->res_lambda[2 * j + 1] = lam[j].imag(); }
  h->res_res.assign(res.begin(), res.end());
  h->res_general = true;
  h->have_subspace = false;
  *Mout = h->res_M;
  *info = info_code;
  *epsout = eps_val;
  *loopout = loop;
  h->stats.ms_total += ttotal.ms();
}

// =====================================================================================================
// C ABI
// =====================================================================================================
#define FC_TRY(hh)  \
  H* h_ = (hh);     \
  try {
#define FC_CATCH                                                  \
  }                                                               \
  catch (const FcError& e) {                                      \
    if (h_) h_->err = e.what(); else g_last_error = e.what();     \
    return e.code;                                                \
  }                                                               \
  catch (const std::exception& e) {                               \
    if (h_) h_->err = e.what(); else g_last_error = e.what();     \
    return FEASTCUDA_ERR_STATE;                                   \
  }                                                               \
  return FEASTCUDA_OK;

static void bind_device(H* h) { FC_CUDA(cudaSetDevice(h->device)); }

extern "C" {

int feastcuda_version(void) { return 100; }

int feastcuda_create(feastcuda_handle* out, int device) {
  FC_TRY(nullptr)
  FC_REQUIRE(out != nullptr, "null handle pointer");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    throw FcError(FEASTCUDA_ERR_CUDA, std::string("no CUDA device available (libfeastcuda has no CPU fallback): ") + cudaGetErrorString(e));
  FC_REQUIRE(device >= 0 && device < count, "device index out of range");
  H* h = new H();
  memset(&h->stats, 0, sizeof(h->stats));
  h->device = device;
  FC_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  FC_CUDA(cudaGetDeviceProperties(&prop, device));
  h->sms = prop.multiProcessorCount;
  FC_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  if (const char* e = getenv("FEASTCUDA_LZ_THREADS")) h->lz_threads = atoi(e);
  if (const char* e = getenv("FEASTCUDA_LZ_CTAS")) h->lz_ctas_per_sm = std::max(1, std::min(8, atoi(e)));
  if (const char* e = getenv("FEASTCUDA_LZ_TILE")) h->lz_tile_rows = std::max(1, atoi(e));
  if (const char* e = getenv("FEASTCUDA_LZ_PAIRED")) h->lz_paired = atoi(e);
  if (const char* e = getenv("FEASTCUDA_LZ_EGRID")) h->lz_egrid_mult = std::max(1, std::min(8, atoi(e)));
  *out = h;
  FC_CATCH
}

int feastcuda_destroy(feastcuda_handle h) {
  if (!h) return FEASTCUDA_OK;
  cudaSetDevice(h->device);
  if (h->arena) arena_release(h);
  if (h->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->nccl_comm);
  for (int s = 0; s < BS_COUNT; ++s) h->blk[s].release();
  DBuf* bufs[] = {&h->partial, &h->partial_r, &h->kstate, &h->small, &h->small2, &h->gram_partial, &h->stage, &h->red_ws,
                  &h->lz_scal, &h->lz_coef, &h->lz_state, &h->lz_ticket, &h->lz_grows, &h->tile_order, &h->goff, &h->cheb_dinv, &h->dA.val32, &h->dB.val32, &h->dA.ptr, &h->dA.col, &h->dA.val, &h->dB.ptr, &h->dB.col, &h->dB.val, &h->dDenseA, &h->dDenseB, &h->dBandA, &h->dBandB, &h->dense_pool, &h->dense_piv, &h->dense_xpool};
  for (DBuf* b : bufs) b->release();
  for (auto& b : h->lu_cache) b.release();
  for (auto& b : h->piv_cache) b.release();
  for (auto e : h->ev_pool) cudaEventDestroy(e);
  for (auto e : h->ev_run) if (e) cudaEventDestroy(e);
  if (h->pinned) cudaFreeHost(h->pinned);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return FEASTCUDA_OK;
}

const char* feastcuda_last_error(feastcuda_handle h) { return h ? h->err.c_str() : g_last_error.c_str(); }

int feastcuda_feastinit(int64_t* fpm) {
  if (!fpm) return FEASTCUDA_ERR_ARG;
  for (int i = 0; i < 64; ++i) fpm[i] = FEAST_UNINIT;
  return FEASTCUDA_OK;
}
int feastcuda_feastdefault(int64_t* fpm) {
  if (!fpm) return FEASTCUDA_ERR_ARG;
  return host_feastdefault(fpm) ? FEASTCUDA_ERR_ARG : FEASTCUDA_OK;
}
int feastcuda_contour(double Emin, double Emax, int64_t* fpm, double* Zne, double* Wne) {
  if (!fpm || !Zne || !Wne) return FEASTCUDA_ERR_ARG;
  const int rc = host_contour(Emin, Emax, fpm, reinterpret_cast<zc*>(Zne), reinterpret_cast<zc*>(Wne));
  return rc == 0 ? FEASTCUDA_OK : (rc == 2 ? FEASTCUDA_ERR_UNSUPPORTED : FEASTCUDA_ERR_ARG);
}
int feastcuda_gcontour(double Emid_re, double Emid_im, double r, int64_t* fpm, double* Zne, double* Wne) {
  if (!fpm || !Zne || !Wne) return FEASTCUDA_ERR_ARG;
  return host_gcontour(zc(Emid_re, Emid_im), r, fpm, reinterpret_cast<zc*>(Zne), reinterpret_cast<zc*>(Wne)) ? FEASTCUDA_ERR_ARG : FEASTCUDA_OK;
}
int feastcuda_node_partition(int64_t ne, int nranks, int rank, int64_t* start, int64_t* count) {
  if (!start || !count || nranks < 1 || rank < 0 || rank >= nranks || ne < 0) return FEASTCUDA_ERR_ARG;
  host_node_partition(ne, nranks, rank, start, count);
  return FEASTCUDA_OK;
}

static int set_csr_common(H* h, int which, int64_t n, int64_t nnz, const int64_t* ptr, const int64_t* idx, const double* val,
                          bool cplx, int base, int fmt, int structure) {
  FC_TRY(h)
  FC_REQUIRE(h != nullptr, "null handle");
  FC_REQUIRE(which == FEASTCUDA_A || which == FEASTCUDA_B, "which must be A or B");
  FC_REQUIRE(base == 0 || base == 1, "index_base must be 0 or 1");
  bind_device(h);
  if (which == FEASTCUDA_A) {
    ingest_csr(h->hA, n, nnz, ptr, idx, val, cplx, base, fmt, structure);
    h->dA.uploaded = false;
    h->dA.val32_ready = false;
    if (h->kind != OP_SPARSE || (h->has_b && h->hB.n != n)) { h->has_b = false; h->hB.set = false; h->dB.uploaded = false; }   // a B of another size never survives a new A
    h->cheb_ready = false;
    h->g2_ready = false;
    h->kind = OP_SPARSE;
    h->n = n;
    release_factor_cache(h);
  } else {
    FC_REQUIRE(h->kind == OP_SPARSE && h->hA.set && h->hA.n == n, "set A (same size, sparse) before B");
    ingest_csr(h->hB, n, nnz, ptr, idx, val, cplx, base, fmt, structure);
    h->dB.uploaded = false;
    h->has_b = true;
    h->cheb_ready = false;
    h->g2_ready = false;
  }
  FC_CATCH
}
int feastcuda_set_csr_d(feastcuda_handle h, int which, int64_t n, int64_t nnz, const int64_t* ptr, const int64_t* idx,
                        const double* val, int index_base, int fmt, int structure) {
  return set_csr_common(h, which, n, nnz, ptr, idx, val, false, index_base, fmt, structure);
}
int feastcuda_set_csr_z(feastcuda_handle h, int which, int64_t n, int64_t nnz, const int64_t* ptr, const int64_t* idx,
                        const double* val, int index_base, int fmt, int structure) {
  return set_csr_common(h, which, n, nnz, ptr, idx, val, true, index_base, fmt, structure);
}
int feastcuda_clear_b(feastcuda_handle h) {
  if (!h) return FEASTCUDA_ERR_ARG;
  h->has_b = false;
  h->hB.set = false;
  h->cheb_ready = false;
  h->g2_ready = false;
  h->denseB.set = false;
  h->bandB.set = false;
  h->dB.uploaded = false;
  h->dense_uploaded = false;
  h->band_uploaded = false;
  release_factor_cache(h);
  return FEASTCUDA_OK;
}

int feastcuda_set_matfree_d(feastcuda_handle h, int64_t n, feastcuda_apply_fn apply_a, void* ctx) {
  FC_TRY(h)
  FC_REQUIRE(h != nullptr, "null handle");
  FC_REQUIRE(n >= 1 && apply_a != nullptr, "matrix-free operator: n >= 1 and a callback are required");
  bind_device(h);
  release_factor_cache(h);
  h->kind = OP_MATFREE;
  h->mf_apply = apply_a;
  h->mf_ctx = ctx;
  h->n = n;
  h->has_b = false;
  h->hA.set = false;
  h->hB.set = false;
  h->dev_complex = false;
  FC_CATCH
}

int feastcuda_set_dense_d(feastcuda_handle h, int which, int64_t n, const double* a, int64_t lda, int structure) {
  FC_TRY(h)
  FC_REQUIRE(h != nullptr, "null handle");
  bind_device(h);
  dense_set(h, which, n, a, lda, false, structure);
  FC_CATCH
}
int feastcuda_set_dense_z(feastcuda_handle h, int which, int64_t n, const double* a, int64_t lda, int structure) {
  FC_TRY(h)
  FC_REQUIRE(h != nullptr, "null handle");
  bind_device(h);
  dense_set(h, which, n, a, lda, true, structure);
  FC_CATCH
}
int feastcuda_set_band_d(feastcuda_handle h, int which, int64_t n, int64_t k, const double* ab, int64_t ldab, int structure) {
  FC_TRY(h)
  FC_REQUIRE(h != nullptr, "null handle");
  bind_device(h);
  band_set(h, which, n, k, ab, ldab, false, structure);
  FC_CATCH
}
int feastcuda_set_band_z(feastcuda_handle h, int which, int64_t n, int64_t k, const double* ab, int64_t ldab, int structure) {
  FC_TRY(h)
  FC_REQUIRE(h != nullptr, "null handle");
  bind_device(h);
  band_set(h, which, n, k, ab, ldab, true, structure);
  FC_CATCH
}


int feastcuda_upload_subspace(feastcuda_handle h, int64_t m0, const double* Q0, int q0_real) {
  FC_TRY(h)
  FC_REQUIRE(h != nullptr, "null handle");
  FC_REQUIRE(h->kind != OP_NONE, "set the operator first");
  bind_device(h);
  prepare_operator(h);
  const int64_t ng = h->row_sharded ? h->n_glob : h->n;
  FC_REQUIRE(m0 >= 1 && m0 <= ng, "Number of eigenvalues M0 must be between 1 and N");
  ensure_workspace(h, h->n, (int)m0);
  FC_REQUIRE(!(h->row_sharded && Q0 == nullptr), "row sharding: pass the initial subspace (the library seed is generated per handle)");
  if (Q0 == nullptr) {
    // library seed: deterministic real Gaussian columns (xorshift + Box-Muller), unit 2-norm -- the stand-in for
    // _feast_seeded_subspace_complex! (core/feast_tools.jl:22-43), whose Julia RNG stream cannot be reproduced
    std::vector<double> q((size_t)h->n * m0);
    uint64_t s = 0x9E3779B97F4A7C15ull ^ ((uint64_t)h->n * 1315423911ull + (uint64_t)m0);
    auto next = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)((s >> 11) + 1) / 9007199254740993.0; };
    for (int64_t j = 0; j < m0; ++j) {
      double nrm = 0.0;
      for (int64_t i = 0; i < h->n; ++i) {
        const double u1 = next(), u2 = next();
        const double g = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
        q[(size_t)j * h->n + i] = g;
        nrm += g * g;
      }
      nrm = nrm > 0 ? std::sqrt(nrm) : 1.0;
      for (int64_t i = 0; i < h->n; ++i) q[(size_t)j * h->n + i] /= nrm;
    }
    upload_block<double>(h, h->n, (int)m0, q.data(), blk(h, BS_QB));
    h->sub_real = true;
  } else if (q0_real) {
    upload_block<double>(h, h->n, (int)m0, Q0, blk(h, BS_QB));
    h->sub_real = true;
  } else {
    upload_block<zd>(h, h->n, (int)m0, reinterpret_cast<const zd*>(Q0), blk(h, BS_QB));
    h->sub_real = false;
  }
  h->have_subspace = true;
  h->sub_m0 = m0;
  FC_CATCH
}

int feastcuda_run_interval(feastcuda_handle h, double Emin, double Emax, int64_t m0, int64_t* fpm, const double* Zne,
                           const double* Wne, int64_t ne, const feastcuda_solver_opts* opts, int64_t* M, int64_t* info,
                           double* epsout, int64_t* loop) {
  FC_TRY(h)
  FC_REQUIRE(h != nullptr && fpm && M && info && epsout && loop, "null argument");
  bind_device(h);
  if (!h->ev_run[0]) { FC_CUDA(cudaEventCreate(&h->ev_run[0])); FC_CUDA(cudaEventCreate(&h->ev_run[1])); }
  FC_CUDA(cudaEventRecord(h->ev_run[0], h->stream));
  run_interval(h, Emin, Emax, (int)m0, fpm, reinterpret_cast<const zc*>(Zne), reinterpret_cast<const zc*>(Wne), (int)ne, opts, M, info,
               epsout, loop, h->sub_real);
  FC_CUDA(cudaEventRecord(h->ev_run[1], h->stream));
  FC_CUDA(cudaEventSynchronize(h->ev_run[1]));
  float ms_ev = 0.f;
  FC_CUDA(cudaEventElapsedTime(&ms_ev, h->ev_run[0], h->ev_run[1]));
  h->stats.ms_dev_run += ms_ev;
  FC_CATCH
}

int feastcuda_fetch_results(feastcuda_handle h, int64_t m0, int x_real, double* lambda, double* X, double* res) {
  FC_TRY(h)
  FC_REQUIRE(h != nullptr, "null handle");
  FC_REQUIRE(h->res_m0 == m0, "no results for this M0");
  bind_device(h);
  const int M = (int)h->res_M;
  const int nl = h->res_general ? 2 : 1;
  if (lambda) for (int j = 0; j < M * nl; ++j) lambda[j] = h->res_lambda[j];
  if (res) for (int j = 0; j < M; ++j) res[j] = h->res_res[j];
  if (X && M > 0) {
    if (x_real) download_block<double>(h, h->n, M, blk(h, BS_XR), X);
    else download_block<zd>(h, h->n, M, blk(h, BS_XR), reinterpret_cast<zd*>(X));
  }
  FC_CATCH
}

int feastcuda_solve_interval(feastcuda_handle h, double Emin, double Emax, int64_t m0, int64_t* fpm, const double* Zne,
                             const double* Wne, int64_t ne, const double* Q0, const feastcuda_solver_opts* opts, double* lambda,
                             double* X, double* res, int64_t* M, int64_t* info, double* epsout, int64_t* loop) {
  int rc = feastcuda_upload_subspace(h, m0, Q0, opts ? opts->q0_real : 0);
  if (rc) return rc;
  rc = feastcuda_run_interval(h, Emin, Emax, m0, fpm, Zne, Wne, ne, opts, M, info, epsout, loop);
  if (rc) return rc;
  return feastcuda_fetch_results(h, m0, opts ? opts->x_real : 0, lambda, X, res);
}

static int run_contour_guarded(feastcuda_handle h, double Emid_re, double Emid_im, double r, int64_t m0, int64_t* fpm, const double* Zne,
                               const double* Wne, int64_t ne, const feastcuda_solver_opts* opts, int64_t* M, int64_t* info, double* epsout,
                               int64_t* loop) {
  FC_TRY(h)
  FC_REQUIRE(h != nullptr && fpm && M && info && epsout && loop, "null argument");
  bind_device(h);
  run_contour(h, zc(Emid_re, Emid_im), r, (int)m0, fpm, reinterpret_cast<const zc*>(Zne), reinterpret_cast<const zc*>(Wne), (int)ne, opts, M,
              info, epsout, loop);
  FC_CATCH
}

int feastcuda_solve_contour(feastcuda_handle h, double Emid_re, double Emid_im, double r, int64_t m0, int64_t* fpm,
                            const double* Zne, const double* Wne, int64_t ne, const double* Q0, const feastcuda_solver_opts* opts,
                            double* lambda, double* X, double* res, int64_t* M, int64_t* info, double* epsout, int64_t* loop) {
  int rc = feastcuda_upload_subspace(h, m0, Q0, opts ? opts->q0_real : 0);
  if (rc) return rc;
  rc = run_contour_guarded(h, Emid_re, Emid_im, r, m0, fpm, Zne, Wne, ne, opts, M, info, epsout, loop);
  if (rc) return rc;
  return feastcuda_fetch_results(h, m0, 0, lambda, X, res);
}

// ---- stage-level entry points ---------------------------------------------------------------------
static void stage_begin(H* h, int64_t m) {
  FC_REQUIRE(h != nullptr, "null handle");
  bind_device(h);
  prepare_operator(h);
  FC_REQUIRE(m >= 1 && m <= FC_MAXCOLS, "m must be in [1,128]");
  ensure_workspace(h, h->n, std::max<int>((int)m, h->ws_n == h->n ? h->ws_ld : 0));
}

int feastcuda_spmm_shifted(feastcuda_handle h, double z_re, double z_im, int64_t m, const double* X, double* Y) {
  FC_TRY(h)
  stage_begin(h, m);
  FC_REQUIRE(h->kind == OP_SPARSE, "sparse operator required");
  upload_block<zd>(h, h->n, (int)m, reinterpret_cast<const zd*>(X), blk(h, BS_KP));
  launch_spmm<SPMM_PLAIN>(h, op_shifted(zc(z_re, z_im)), (int)m, blk(h, BS_KP), blk(h, BS_KV), nullptr, nullptr);
  download_block<zd>(h, h->n, (int)m, blk(h, BS_KV), reinterpret_cast<zd*>(Y));
  FC_CATCH
}

int feastcuda_apply(feastcuda_handle h, int which, int64_t m, const double* X, double* Y) {
  FC_TRY(h)
  stage_begin(h, m);
  upload_block<zd>(h, h->n, (int)m, reinterpret_cast<const zd*>(X), blk(h, BS_KP));
  if (which == FEASTCUDA_B && !h->has_b) copy_cols(h, (int)m, blk(h, BS_KP), blk(h, BS_KV));
  else apply_op(h, which, (int)m, blk(h, BS_KP), blk(h, BS_KV));
  download_block<zd>(h, h->n, (int)m, blk(h, BS_KV), reinterpret_cast<zd*>(Y));
  FC_CATCH
}

int feastcuda_block_solve(feastcuda_handle h, double z_re, double z_im, int64_t m, const double* RHS, const double* X0,
                          const feastcuda_solver_opts* opts, double* Xout, int64_t* iters, double* resid) {
  FC_TRY(h)
  stage_begin(h, m);
  feastcuda_solver_opts o = opts ? *opts : default_opts();
  upload_block<zd>(h, h->n, (int)m, reinterpret_cast<const zd*>(RHS), blk(h, BS_RHS));
  if (X0) upload_block<zd>(h, h->n, (int)m, reinterpret_cast<const zd*>(X0), blk(h, BS_KX));
  SolveOut so;
  const double tol = o.tol == 0.0 ? 1e-12 : o.tol;
  if (h->kind != OP_SPARSE) release_factor_cache(h);
  node_solve(h, 0, zc(z_re, z_im), (int)m, blk(h, BS_RHS), blk(h, BS_KX), X0 != nullptr, o, tol, so);
  if (h->kind != OP_SPARSE) {
    // direct solves report the true residual through the generic kernels
    so.iters.assign(m, 0);
    so.truenorm.assign(m, 0.0);
    release_factor_cache(h);
  }
  download_block<zd>(h, h->n, (int)m, blk(h, BS_KX), reinterpret_cast<zd*>(Xout));
  for (int c = 0; c < m; ++c) {
    if (iters) iters[c] = so.iters.empty() ? 0 : so.iters[c];
    if (resid) resid[c] = so.truenorm.empty() ? 0.0 : so.truenorm[c];
  }
  FC_CATCH
}

int feastcuda_accumulate(feastcuda_handle h, double w_re, double w_im, int64_t m, const double* Y, double* Qacc) {
  FC_TRY(h)
  stage_begin(h, m);
  upload_block<zd>(h, h->n, (int)m, reinterpret_cast<const zd*>(Y), blk(h, BS_KP));
  upload_block<zd>(h, h->n, (int)m, reinterpret_cast<const zd*>(Qacc), blk(h, BS_ACC));
  axpby_cols(h, (int)m, zc(w_re, w_im), 1.0, blk(h, BS_KP), blk(h, BS_ACC));
  download_block<zd>(h, h->n, (int)m, blk(h, BS_ACC), reinterpret_cast<zd*>(Qacc));
  FC_CATCH
}

// stage calls that do not involve the operator still need (n, m) workspaces
static void stage_begin_free(H* h, int64_t n, int64_t m) {
  FC_REQUIRE(h != nullptr, "null handle");
  bind_device(h);
  FC_REQUIRE(n >= 1 && m >= 1 && m <= FC_MAXCOLS, "bad block shape");
  ensure_workspace(h, n, std::max<int>((int)m, h->ws_n == n ? h->ws_ld : 0));
}

int feastcuda_orthonormalize(feastcuda_handle h, int64_t n, int64_t m, const double* W, double rank_tol, double* Q, int64_t* rank) {
  FC_TRY(h)
  stage_begin_free(h, n, m);
  FC_REQUIRE(rank != nullptr, "null rank");
  upload_block<zd>(h, n, (int)m, reinterpret_cast<const zd*>(W), blk(h, BS_ACC));
  int slot = BS_ACC;
  const int rk = orthonormalize(h, (int)m, BS_ACC, BS_KP, rank_tol > 0 ? rank_tol : std::sqrt(2.220446049250313e-16), &slot);
  *rank = rk;
  if (rk > 0) download_block<zd>(h, n, rk, blk(h, slot), reinterpret_cast<zd*>(Q));
  FC_CATCH
}

int feastcuda_gram(feastcuda_handle h, int64_t n, int64_t m, const double* X, const double* Y, double* C) {
  FC_TRY(h)
  stage_begin_free(h, n, m);
  upload_block<zd>(h, n, (int)m, reinterpret_cast<const zd*>(X), blk(h, BS_KP));
  upload_block<zd>(h, n, (int)m, reinterpret_cast<const zd*>(Y), blk(h, BS_KV));
  std::vector<zc> G;
  gram_host(h, (int)m, (int)m, blk(h, BS_KP), blk(h, BS_KV), G);
  zc* out = reinterpret_cast<zc*>(C);
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) out[(size_t)j * m + i] = G[(size_t)i * m + j];
  FC_CATCH
}

int feastcuda_reduced_eig(feastcuda_handle h, int64_t r, const double* Sq, const double* Aq, double* lambda, double* V, int64_t* sweeps) {
  FC_TRY(h)
  FC_REQUIRE(h != nullptr, "null handle");
  bind_device(h);
  FC_REQUIRE(r >= 1 && r <= FC_MAXCOLS && Sq && lambda && V, "bad arguments");
  h->small.ensure((size_t)8 * FC_MAXCOLS * FC_MAXCOLS * sizeof(zd));
  h->small2.ensure((size_t)4 * FC_MAXCOLS * sizeof(zd) + 64);
  const zc* s = reinterpret_cast<const zc*>(Sq);
  const zc* a = reinterpret_cast<const zc*>(Aq);
  std::vector<zc> S((size_t)r * r), A;
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < r; ++j) S[(size_t)i * r + j] = s[(size_t)j * r + i];
  if (a) {
    A.resize((size_t)r * r);
    for (int i = 0; i < r; ++i)
      for (int j = 0; j < r; ++j) A[(size_t)i * r + j] = a[(size_t)j * r + i];
  }
  std::vector<double> lam;
  std::vector<zc> Vr;
  int st = 0;
  const int64_t sw0 = h->stats.jacobi_sweeps;
  reduced_eig(h, (int)r, S, a ? &A : nullptr, lam, Vr, &st);
  if (st == 1) throw FcError(FEASTCUDA_ERR_ARG, "reduced_eig: Aq is not positive definite");
  for (int i = 0; i < r; ++i) lambda[i] = lam[i];
  zc* vo = reinterpret_cast<zc*>(V);
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < r; ++j) vo[(size_t)j * r + i] = Vr[(size_t)i * r + j];
  if (sweeps) *sweeps = h->stats.jacobi_sweeps - sw0;
  FC_CATCH
}

int feastcuda_rowtransform(feastcuda_handle h, int64_t n, int64_t a, int64_t b, const double* X, const double* T, double* Y) {
  FC_TRY(h)
  stage_begin_free(h, n, std::max(a, b));
  FC_REQUIRE(a >= 1 && b >= 1 && X && T && Y, "bad arguments");
  upload_block<zd>(h, n, (int)a, reinterpret_cast<const zd*>(X), blk(h, BS_KP));
  const zc* t = reinterpret_cast<const zc*>(T);
  std::vector<zc> Tr((size_t)a * b);
  for (int64_t i = 0; i < a; ++i)
    for (int64_t j = 0; j < b; ++j) Tr[(size_t)i * b + j] = t[(size_t)j * a + i];
  rowtransform(h, (int)a, (int)b, blk(h, BS_KP), Tr, blk(h, BS_KV));
  download_block<zd>(h, n, (int)b, blk(h, BS_KV), reinterpret_cast<zd*>(Y));
  FC_CATCH
}

int feastcuda_eig_general(feastcuda_handle h, int64_t r, const double* A, const double* B, double* lambda, double* V) {
  FC_TRY(h)
  FC_REQUIRE(h != nullptr, "null handle");
  FC_REQUIRE(r >= 1 && r <= FC_MAXCOLS && A && lambda && V, "bad arguments");
  const zc* a = reinterpret_cast<const zc*>(A);
  const zc* b = reinterpret_cast<const zc*>(B);
  std::vector<zc> C((size_t)r * r), S, lam, Vr;
  for (int64_t i = 0; i < r; ++i)
    for (int64_t j = 0; j < r; ++j) C[(size_t)i * r + j] = a[(size_t)j * r + i];
  if (b) {
    S.resize((size_t)r * r);
    for (int64_t i = 0; i < r; ++i)
      for (int64_t j = 0; j < r; ++j) S[(size_t)i * r + j] = b[(size_t)j * r + i];
    if (!host_pencil_eig((int)r, C, S, lam, Vr))
      throw FcError(FEASTCUDA_ERR_STATE, "eig_general: the reduced pencil holds non-finite entries or its QR iteration did not converge");
  } else if (!host_complex_eig((int)r, C, lam, Vr)) {
    throw FcError(FEASTCUDA_ERR_STATE, "eig_general: QR iteration did not converge");
  }
  zc* vo = reinterpret_cast<zc*>(V);
  for (int64_t k = 0; k < r; ++k) {
    lambda[2 * k] = lam[k].real();
    lambda[2 * k + 1] = lam[k].imag();
    for (int64_t i = 0; i < r; ++i) vo[(size_t)k * r + i] = Vr[(size_t)i * r + k];
  }
  FC_CATCH
}

int feastcuda_residuals(feastcuda_handle h, int64_t m, const double* X, const double* lambda, double* res) {
  FC_TRY(h)
  stage_begin(h, m);
  upload_block<zd>(h, h->n, (int)m, reinterpret_cast<const zd*>(X), blk(h, BS_KP));
  std::vector<zc> lam(m);
  for (int j = 0; j < m; ++j) lam[j] = zc(lambda[2 * j], lambda[2 * j + 1]);
  std::vector<double> rn;
  eig_residual_norms(h, (int)m, blk(h, BS_KP), lam, rn);
  for (int j = 0; j < m; ++j) res[j] = rn[j] / std::max(std::abs(lam[j]), 1.0);
  FC_CATCH
}

int feastcuda_nccl_unique_id(char* id128) {
  FC_TRY(nullptr)
  FC_REQUIRE(id128 != nullptr, "null id");
  nccl_load();
  FC_NCCL(g_nccl.GetUniqueId(id128));
  FC_CATCH
}

int feastcuda_nccl_init(feastcuda_handle h, int nranks, int rank, const char* id128) {
  FC_TRY(h)
  FC_REQUIRE(h != nullptr && id128 != nullptr && nranks >= 1 && rank >= 0 && rank < nranks, "bad arguments");
  bind_device(h);
  nccl_load();
  NcclId id;
  memcpy(id.internal, id128, 128);
  FC_NCCL(g_nccl.CommInitRank(&h->nccl_comm, nranks, id, rank));
  unsigned long long tk = 1469598103934665603ull;      // FNV-1a of the unique id: the same on every rank of this communicator
  for (int i = 0; i < 128; ++i) { tk ^= (unsigned char)id128[i]; tk *= 1099511628211ull; }
  h->ipc_token = tk;
  h->nranks = nranks;
  h->rank = rank;
  FC_CATCH
}

int feastcuda_set_row_sharding(feastcuda_handle h, int on) {
  FC_TRY(h)
  FC_REQUIRE(h != nullptr, "null handle");
  bind_device(h);
  const bool want = on != 0 && (h->nranks > 1 || getenv("FEASTCUDA_FORCE_ROWS") != nullptr);   // the env knob runs the sharded kernels on one rank (A/B timing)
  if (want != h->row_sharded) {
    if (h->arena) arena_release(h);
    for (int s = 0; s < BS_COUNT; ++s) h->blk[s].release();
    h->row_sharded = want;
    h->ws_n = 0;
    h->ws_ld = 0;
    h->have_subspace = false;
    h->dA.uploaded = false;
    h->dB.uploaded = false;
    h->dA.val32_ready = false;
  }
  FC_CATCH
}

int feastcuda_row_range(feastcuda_handle h, int64_t* row0, int64_t* nrows, int64_t* nglobal) {
  FC_TRY(h)
  FC_REQUIRE(h != nullptr && row0 && nrows && nglobal, "null argument");
  bind_device(h);
  prepare_operator(h);
  *row0 = h->row_sharded ? h->row0 : 0;
  *nrows = h->n;
  *nglobal = h->row_sharded ? h->n_glob : h->n;
  FC_CATCH
}

int feastcuda_get_stats(feastcuda_handle h, feastcuda_stats* out) {
  if (!h || !out) return FEASTCUDA_ERR_ARG;
  if (h->kind == OP_SPARSE && h->hA.set) {
    const double vs = h->dev_complex ? 16.0 : 8.0;
    const int m = h->ws_ld;
    double b = (double)h->hA.nnz * (vs + 4.0) + 4.0 * (h->hA.n + 1) + 2.0 * (double)h->hA.n * m * 16.0;
    if (h->has_b) b += (double)h->hB.nnz * (vs + 4.0) + 4.0 * (h->hB.n + 1);
    h->stats.bytes_spmm_alg = b;
  }
  *out = h->stats;
  return FEASTCUDA_OK;
}
int feastcuda_reset_stats(feastcuda_handle h) {
  if (!h) return FEASTCUDA_ERR_ARG;
  memset(&h->stats, 0, sizeof(h->stats));
  return FEASTCUDA_OK;
}

}  // extern "C"
