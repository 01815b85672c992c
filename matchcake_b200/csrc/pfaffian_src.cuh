// Matrix sources shared by the Pfaffian kernels and their backward passes.
#pragma once
#include "common.cuh"

namespace mcb200 {

// A source returns A[i][j] of matrix `item` for arbitrary 0 <= i, j < 8*MT (0 outside the real matrix).
struct PlainSrc {
  const double* A;
  int m;
  __device__ __forceinline__ double at(long long item, int i, int j) const {
    if (i >= m || j >= m || i == j) return 0.0;
    // the skew part (A - A^T)/2, like sqrt|det| / the reference's signed Pfaffian see it to first order (their
    // gradient is the antisymmetric Pf/2 A^-T, tests/test_utils/test_pfaffian.py:86-115); exact for skew input
    const double* a = A + item * (long long)m * m;
    return 0.5 * (a[i * m + j] - a[j * m + i]);
  }
};

struct ProbsSrc {
  const double* cov;
  const int8_t* outcomes;
  int Q, k, m;
  __device__ __forceinline__ double at(long long item, int i, int j) const {
    if (i >= m || j >= m || i == j) return 0.0;
    const long long b = item / Q;
    const int q = (int)(item - b * Q);
    const double* a = cov + b * (long long)m * m;
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    double v = a[lo * m + hi];
    if ((lo & 1) == 0 && hi == lo + 1) v += outcomes[(long long)q * k + (lo >> 1)] ? 1.0 : -1.0;
    return i < j ? v : -v;
  }
};

struct SectorSrc {
  const double* cov;
  const int32_t* index_sets;
  int D, T, m;
  __device__ __forceinline__ double at(long long item, int i, int j) const {
    if (i >= m || j >= m || i == j) return 0.0;
    const long long b = item / T;
    const int t = (int)(item - b * T);
    const int32_t* idx = index_sets + (long long)t * m;
    const double* a = cov + b * (long long)D * D;
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    const double v = a[(long long)idx[lo] * D + idx[hi]];
    return i < j ? v : -v;
  }
};


}  // namespace mcb200
