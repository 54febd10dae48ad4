// Complex arithmetic helpers shared by every kernel of libfeastcuda.
// Block vectors are stored interleaved (re,im), exactly as Julia's Complex{T} and LAPACK do.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace feastcuda {

template <typename R>
struct alignas(2 * sizeof(R)) cx {
  R x, y;
};

using zd = cx<double>;
using zf = cx<float>;

#define FC_HD __host__ __device__ __forceinline__

template <typename R> FC_HD cx<R> mk(R a, R b) { cx<R> r; r.x = a; r.y = b; return r; }
template <typename R> FC_HD cx<R> czero() { return mk<R>(R(0), R(0)); }
template <typename R> FC_HD cx<R> operator+(cx<R> a, cx<R> b) { return mk<R>(a.x + b.x, a.y + b.y); }
template <typename R> FC_HD cx<R> operator-(cx<R> a, cx<R> b) { return mk<R>(a.x - b.x, a.y - b.y); }
template <typename R> FC_HD cx<R> operator-(cx<R> a) { return mk<R>(-a.x, -a.y); }
template <typename R> FC_HD cx<R> operator*(cx<R> a, cx<R> b) { return mk<R>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <typename R> FC_HD cx<R> operator*(R a, cx<R> b) { return mk<R>(a * b.x, a * b.y); }
template <typename R> FC_HD cx<R> operator*(cx<R> a, R b) { return mk<R>(a.x * b, a.y * b); }
template <typename R> FC_HD cx<R> conj(cx<R> a) { return mk<R>(a.x, -a.y); }
template <typename R> FC_HD R abs2(cx<R> a) { return a.x * a.x + a.y * a.y; }
template <typename R> FC_HD cx<R> operator/(cx<R> a, cx<R> b) {
  // Smith's algorithm: no overflow for the tiny/huge Krylov scalars
  R ar = b.x < 0 ? -b.x : b.x, ai = b.y < 0 ? -b.y : b.y;
  if (ar >= ai) {
    R t = b.y / b.x, d = b.x + b.y * t;
    return mk<R>((a.x + a.y * t) / d, (a.y - a.x * t) / d);
  } else {
    R t = b.x / b.y, d = b.x * t + b.y;
    return mk<R>((a.x * t + a.y) / d, (a.y * t - a.x) / d);
  }
}
// acc += a*b with a real or complex matrix entry a
template <typename R> FC_HD void fma_acc(cx<R>& acc, R a, cx<R> b) { acc.x = fma(a, b.x, acc.x); acc.y = fma(a, b.y, acc.y); }
template <typename R> FC_HD void fma_acc(cx<R>& acc, cx<R> a, cx<R> b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
// acc += conj(a)*b
template <typename R> FC_HD void fma_conj_acc(cx<R>& acc, cx<R> a, cx<R> b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}

template <typename T> struct zero_of;
template <> struct zero_of<double> { static FC_HD double v() { return 0.0; } };
template <> struct zero_of<float> { static FC_HD float v() { return 0.f; } };
template <typename R> struct zero_of<cx<R>> { static FC_HD cx<R> v() { return czero<R>(); } };

template <typename R> __device__ __forceinline__ cx<R> shfl_cx(unsigned mask, cx<R> v, int src, int width) {
  cx<R> r;
  r.x = __shfl_sync(mask, v.x, src, width);
  r.y = __shfl_sync(mask, v.y, src, width);
  return r;
}

}  // namespace feastcuda
