// Dense and banded operator paths (direct per-node LU).  Filled in by dense_band_impl below.
#pragma once
#include "engine.hpp"
namespace feastcuda {
void dense_set(feastcuda_handle_s* h, int which, int64_t n, const double* a, int64_t lda, bool cplx, int structure);
void dense_prepare(feastcuda_handle_s* h);
bool dense_node_solve(feastcuda_handle_s* h, int node, zc z, int m, const zd* RHS, zd* X);
bool dense_batch_solve(feastcuda_handle_s* h, int ne_total, int first, int count, const zc* shifts, int m, const zd* RHS, zd** Xpool,
                       int64_t* xbatch);
void dense_apply(feastcuda_handle_s* h, int which, int m, const zd* X, zd* Y);
void band_set(feastcuda_handle_s* h, int which, int64_t n, int64_t k, const double* ab, int64_t ldab, bool cplx, int structure);
void band_prepare(feastcuda_handle_s* h);
bool band_node_solve(feastcuda_handle_s* h, int node, zc z, int m, const zd* RHS, zd* X);
bool band_batch_fits(feastcuda_handle_s* h, int count);
bool band_batch_solve(feastcuda_handle_s* h, int first, int count, const zc* shifts, int m, const zd* RHS, zd** Xpool, int64_t* xbatch);
void band_apply(feastcuda_handle_s* h, int which, int m, const zd* X, zd* Y);
// sampled kernel timings (CUDA events on the library's stream; feastcuda.cu)
int sample_begin(feastcuda_handle_s* h, int kind, int tag = 0);
void sample_end(feastcuda_handle_s* h, int slot);
}  // namespace feastcuda
