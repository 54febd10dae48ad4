// Host-side engine state of libfeastcuda (one handle = one GPU = one solve at a time).
#pragma once
#include <cuda_runtime.h>

#include <chrono>
#include <complex>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/feastcuda.h"
#include "cxmath.cuh"
#include "peer_arena.hpp"

namespace feastcuda {

typedef std::complex<double> zc;

struct FcError : std::runtime_error {
  int code;
  FcError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define FC_CUDA(call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess)                                                                              \
      throw FcError(FEASTCUDA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " at " +  \
                                            __FILE__ + ":" + std::to_string(__LINE__));                \
  } while (0)

#define FC_REQUIRE(cond, msg)                                      \
  do {                                                             \
    if (!(cond)) throw FcError(FEASTCUDA_ERR_ARG, std::string(msg)); \
  } while (0)

struct DBuf {
  void* p = nullptr;
  size_t cap = 0;
  bool owned = true;   // false: a view into the shared arena of a row-sharded run
  void ensure(size_t bytes) {
    if (bytes <= cap) return;
    if (!owned) throw FcError(FEASTCUDA_ERR_STATE, "arena slot too small");
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    FC_CUDA(cudaMalloc(&p, bytes));
    cap = bytes;
  }
  void release() {
    if (p && owned) cudaFree(p);
    p = nullptr;
    cap = 0;
    owned = true;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct HostCsr {
  int64_t n = 0, nnz = 0;
  std::vector<int> ptr, col;
  std::vector<double> val;  // nnz (real) or 2*nnz (complex interleaved)
  bool cplx = false;
  bool set = false;
  int structure = 0;   // FEASTCUDA_SYM / HERM / GEN as declared at set time
};

struct DevCsr {
  DBuf ptr, col, val;
  DBuf val32;               // FP32 copy of the entries (mixed-precision Lanczos vectors, real operators only)
  bool uploaded = false, val32_ready = false;
};

struct HostDense {  // column-major n x n, complex interleaved or real
  int64_t n = 0;
  std::vector<double> a;
  bool cplx = false, set = false;
};

struct HostBand {   // full general band (2k+1) x n, diagonal row k, values expanded from the input storage
  int64_t n = 0, k = 0;
  std::vector<double> ab;  // interleaved complex always (2*(2k+1)*n)
  bool cplx = false, set = false;
};

enum OperatorKind { OP_NONE = 0, OP_SPARSE = 1, OP_DENSE = 2, OP_BAND = 3, OP_MATFREE = 4 };

// block-vector slots (each n x ld complex, row-major)
enum BlockSlot { BS_QB = 0, BS_RHS, BS_ACC, BS_XR, BS_KX, BS_KR, BS_KRH, BS_KP, BS_KV, BS_KS, BS_KT, BS_KB, BS_COUNT };

struct Timer {
  std::chrono::steady_clock::time_point t0;
  Timer() : t0(std::chrono::steady_clock::now()) {}
  double ms() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

}  // namespace feastcuda

struct feastcuda_handle_s {
  int device = 0;
  int sms = 148;
  cudaStream_t stream = nullptr;
  int kind = feastcuda::OP_NONE;
  feastcuda_apply_fn mf_apply = nullptr;   // OP_MATFREE: the caller's device mat-vec (real symmetric A, B = I)
  void* mf_ctx = nullptr;
  bool dev_complex = false;  // value type of the uploaded operators
  feastcuda::HostCsr hA, hB;
  feastcuda::DevCsr dA, dB;
  feastcuda::HostDense denseA, denseB;
  feastcuda::HostBand bandA, bandB;
  bool has_b = false;
  int64_t n = 0;

  // workspaces
  int64_t ws_n = 0;
  int ws_ld = 0;
  feastcuda::DBuf blk[feastcuda::BS_COUNT];
  feastcuda::DBuf partial, partial_r, kstate, small, small2, gram_partial, stage, red_ws;
  feastcuda::DBuf lz_scal, lz_coef, lz_state;   // multi-shift Lanczos: per-step scalars, pass-2 coefficients, shift recurrences
  // generalized Lanczos filter: Chebyshev inner solver for B (spectral interval of D^-1 B, 1/diag(B) on the device)
  bool cheb_ready = false, cheb_usable = false;
  double cheb_lo = 0.0, cheb_hi = 0.0;
  feastcuda::DBuf cheb_dinv;
  // general pencils on the two-sided Lanczos filter: row-scaled operators D^-1 A, D^-1 B and their conjugate transposes
  bool g2_ready = false, g2_usable = false;
  double g2_q = 0.0;                            // Jacobi contraction bound of D^-1 B (0: B = I)
  feastcuda::DevCsr g2A, g2Ah, g2B, g2Bh;
  feastcuda::DBuf g2_ones;
  int lz_egrid_mult = 4;    // CTAs per SM of the elementwise Lanczos kernels (partial rows the scalar kernels reduce)
  int lz_paired = 1;        // pass 2 accumulates Q every second step (0: every step)
  int lz_threads = 512;     // CTA size of the Lanczos SpMM (512 or 1024)
  int lz_ctas_per_sm = 2;   // persistent CTAs per SM of the Lanczos SpMM (contiguous row chunks keep the band in L1)
  int lz_tile_rows = 64;    // rows per round-robin tile of the Lanczos SpMM
  void* pinned = nullptr;
  size_t pinned_cap = 0;
  std::vector<cudaEvent_t> ev_pool;   // reusable event pairs for sampled kernel timings
  size_t ev_used = 0;
  cudaEvent_t ev_run[2] = {nullptr, nullptr};
  struct EvSample { int kind, tag, a, b; };
  std::vector<EvSample> ev_pending;

  // dense / band factor caches (device), one per quadrature node
  std::vector<feastcuda::DBuf> lu_cache;
  std::vector<feastcuda::DBuf> piv_cache;
  std::vector<feastcuda::zc> lu_shift;
  feastcuda::DBuf dDenseA, dDenseB, dBandA, dBandB;
  feastcuda::DBuf dense_pool, dense_piv, dense_xpool;   // batched dense factors / pivots / per-node solutions
  int dense_slots = 0;
  bool dense_uploaded = false, band_uploaded = false;

  // last solve's results (device: blk[BS_XR]); host copies of the small arrays
  int64_t res_m0 = 0, res_M = 0, res_rank = 0;
  std::vector<double> res_lambda, res_res;   // lambda interleaved complex for the general solve
  bool res_general = false;
  bool have_subspace = false;
  bool sub_real = false;
  int64_t sub_m0 = 0;

  // multi-GPU
  void* nccl_comm = nullptr;
  int nranks = 1, rank = 0;
  // row-sharded runs (feastcuda_set_row_sharding): every rank owns a contiguous block of rows of A and of every block vector; `n` is
  // the LOCAL row count then.  All block slots and the mailbox of the one-shot reductions live in ONE allocation per rank (the arena);
  // the arenas of all ranks are mapped into one contiguous virtual range in every process (peer_arena.hpp), so the gather kernels read
  // halo rows straight from the owner's HBM over NVLink.
  bool row_sharded = false;
  int64_t n_glob = 0, row0 = 0, nloc_max = 0;
  feastcuda::PeerArena parena;                 // the ranks' arenas in one contiguous virtual range (peer_arena.hpp)
  unsigned long long ipc_token = 0;            // names the descriptor-passing sockets of this communicator
  void* arena = nullptr;                       // local arena
  size_t arena_bytes = 0, arena_slot_bytes = 0, arena_mbox_off = 0;
  void* peer_arena[16] = {nullptr};            // every rank's arena as mapped in this process (peer_arena[rank] == arena)
  feastcuda::DBuf goff;                        // pre-resolved gather offsets of A's local rows (kernels_lanczos.cuh: k_lz_resolve), two row strides cached
  int64_t goff_rowbytes4[4] = {0, 0, 0, 0};
  int goff_next = 0;
  feastcuda::DBuf tile_order;                  // order in which the gather kernels deal their row tiles (halo tiles spread out)
  int tile_order_tr = 0, halo_start = 0;
  feastcuda::DBuf lz_ticket, lz_grows;         // counters and group sums of the two-level tail reductions (kernels_lanczos.cuh: lz_tail)
  int64_t nnz_loc = 0;                         // stored entries of the local rows
  unsigned long long xseq = 0;                 // sequence number of the last cross-rank exchange

  feastcuda_stats stats;
  std::string err;
};

namespace feastcuda {
// drop every cached per-node factor (device memory included)
inline void release_factor_cache(feastcuda_handle_s* h) {
  for (auto& b : h->lu_cache) b.release();
  for (auto& b : h->piv_cache) b.release();
  h->lu_cache.clear();
  h->piv_cache.clear();
  h->lu_shift.clear();
  h->dense_pool.release();
  h->dense_piv.release();
  h->dense_xpool.release();
  h->dense_slots = 0;
}
}  // namespace feastcuda
