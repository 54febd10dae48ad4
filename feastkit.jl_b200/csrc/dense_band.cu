// Dense and banded operator paths -- placeholders until the batched LU kernels land.
#include "dense_band.cuh"
namespace feastcuda {
static void nyi() { throw FcError(FEASTCUDA_ERR_UNSUPPORTED, "dense/banded operators: not built yet"); }
void dense_set(feastcuda_handle_s*, int, int64_t, const double*, int64_t, bool, int) { nyi(); }
void dense_prepare(feastcuda_handle_s*) { nyi(); }
bool dense_node_solve(feastcuda_handle_s*, int, zc, int, const zd*, zd*) { nyi(); return false; }
void dense_apply(feastcuda_handle_s*, int, int, const zd*, zd*) { nyi(); }
void band_set(feastcuda_handle_s*, int, int64_t, int64_t, const double*, int64_t, bool, int) { nyi(); }
void band_prepare(feastcuda_handle_s*) { nyi(); }
bool band_node_solve(feastcuda_handle_s*, int, zc, int, const zd*, zd*) { nyi(); return false; }
void band_apply(feastcuda_handle_s*, int, int, const zd*, zd*) { nyi(); }
}  // namespace feastcuda
