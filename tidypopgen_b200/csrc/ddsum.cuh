// Double-double accumulation for the consumers of the pairwise matrices (wg17.cu, relate.cu).
//
// The reference averages with R's mean(): a long-double sum and one refinement pass
// (R src/main/summary.c, real_mean), i.e. the correctly rounded mean up to an ulp of the
// 80-bit format.  The GPU has no 80-bit type; sums are kept as unevaluated pairs hi + lo
// (~106 bits), combined in a fixed tree, and rounded to double once at the end — at least as
// close to the exact mean as R's value, and run-to-run identical.
#pragma once
#include <cuda_runtime.h>

namespace tpg {

struct dd {
  double hi, lo;
};

__host__ __device__ __forceinline__ dd dd_zero() { return dd{0.0, 0.0}; }

// a + b exactly as s + e (Knuth two-sum; no FMA contraction: the library is built with --fmad=false)
__host__ __device__ __forceinline__ void two_sum(double a, double b, double& s, double& e) {
  s = a + b;
  const double bb = s - a;
  e = (a - (s - bb)) + (b - bb);
}

__host__ __device__ __forceinline__ dd dd_add_d(dd a, double b) {
  double s, e;
  two_sum(a.hi, b, s, e);
  e += a.lo;
  dd r;
  r.hi = s + e;
  r.lo = e - (r.hi - s);
  return r;
}

__host__ __device__ __forceinline__ dd dd_add(dd a, dd b) {
  double s, e, t, f;
  two_sum(a.hi, b.hi, s, e);
  two_sum(a.lo, b.lo, t, f);
  e += t;
  double hi = s + e;
  double lo = e - (hi - s);
  lo += f;
  dd r;
  r.hi = hi + lo;
  r.lo = lo - (r.hi - hi);
  return r;
}

__host__ __device__ __forceinline__ dd dd_neg(dd a) { return dd{-a.hi, -a.lo}; }

// (hi + lo) / n rounded to double (n > 0 an integer count; 0 gives NaN like mean(numeric(0))).
__host__ __device__ __forceinline__ double dd_mean(dd a, double n) {
  if (n == 0.0) return nan("");
  const double q1 = a.hi / n;
  if (!(q1 - q1 == 0.0)) return q1;  // Inf / NaN pass through
  const double r = fma(-q1, n, a.hi);
  const double q2 = (r + a.lo) / n;
  return q1 + q2;
}

#ifdef __CUDACC__
// Fixed-order butterfly over the 32 lanes; every lane ends with the same total.
__device__ __forceinline__ dd dd_warp_sum(dd v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dd w;
    w.hi = __shfl_xor_sync(0xFFFFFFFFu, v.hi, o);
    w.lo = __shfl_xor_sync(0xFFFFFFFFu, v.lo, o);
    // add in a lane-independent order so that both partners compute identical bits
    const bool low = (threadIdx.x & o) == 0;
    v = low ? dd_add(v, w) : dd_add(w, v);
  }
  return v;
}
__device__ __forceinline__ long long i64_warp_sum(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}
#endif

}  // namespace tpg
