// Example of a user-supplied matrix-free operator for feastcuda_set_matfree_d (include/feastcuda.h): the 7-point Dirichlet
// Laplacian on an nx x ny x nz grid applied to a row-major block, no matrix stored.  Built into lib/libfeastcuda_examples.so by
// __graft_entry__.build(); tests/test_gpu_matfree.py drives feast_matvec with it and checks the eigenpairs against the oracle's
// assembled matrix (oracle/feast_oracle.py:laplacian_3d -- lexicographic ordering, x slowest).
#include <cuda_runtime.h>
#include <cstdint>

struct LaplacianGrid { int nx, ny, nz; };

// one thread per (row, column pair); consecutive threads walk the columns of a row: coalesced 16-byte accesses
__global__ void __launch_bounds__(256) k_laplacian3d(LaplacianGrid g, int64_t n, int ncols, const double* __restrict__ X, int64_t ldx,
                                                    double* __restrict__ Y, int64_t ldy) {
  const int pairs = (ncols + 1) >> 1;
  const int64_t total = n * pairs;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = t / pairs;
    const int c = 2 * (int)(t % pairs);
    const bool two = (c + 1 < ncols) && ((ldx & 1) == 0) && ((ldy & 1) == 0);
    const int iz = (int)(row % g.nz), iy = (int)((row / g.nz) % g.ny), ix = (int)(row / ((int64_t)g.nz * g.ny));
    auto ld = [&](int64_t r) -> double2 {
      const double* p = X + r * ldx + c;
      return two ? *reinterpret_cast<const double2*>(p) : make_double2(p[0], (c + 1 < ncols) ? p[1] : 0.0);
    };
    const double2 x0 = ld(row);
    double2 acc = make_double2(6.0 * x0.x, 6.0 * x0.y);
    auto sub = [&](int64_t r) { const double2 v = ld(r); acc.x -= v.x; acc.y -= v.y; };
    if (iz > 0) sub(row - 1);
    if (iz + 1 < g.nz) sub(row + 1);
    if (iy > 0) sub(row - g.nz);
    if (iy + 1 < g.ny) sub(row + g.nz);
    if (ix > 0) sub(row - (int64_t)g.nz * g.ny);
    if (ix + 1 < g.nx) sub(row + (int64_t)g.nz * g.ny);
    double* q = Y + row * ldy + c;
    if (two) *reinterpret_cast<double2*>(q) = acc;
    else { q[0] = acc.x; if (c + 1 < ncols) q[1] = acc.y; }
  }
}

// signature of feastcuda_apply_fn; ctx points to a LaplacianGrid that outlives the solve
extern "C" void feastcuda_example_laplacian3d(void* ctx, int64_t n, int64_t ncols, const double* X, int64_t ldx, double* Y, int64_t ldy,
                                              void* stream) {
  const LaplacianGrid g = *static_cast<const LaplacianGrid*>(ctx);
  const int64_t total = n * ((ncols + 1) / 2);
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  k_laplacian3d<<<grid > 0 ? grid : 1, 256, 0, static_cast<cudaStream_t>(stream)>>>(g, n, (int)ncols, X, ldx, Y, ldy);
}
