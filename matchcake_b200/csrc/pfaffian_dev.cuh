// Device-side scalar Parlett-Reid Pfaffian of a real skew-symmetric matrix held in shared or global memory (one
// thread block per matrix): the path for matrices larger than the tile kernels take (m > 128) and for the
// block-per-instance fused projector kernel.
// Only the strict upper triangle A[i][j], i < j, is stored, read and updated (skew symmetry supplies the
// rest), so the arithmetic is the ~m^3/3 flop of the structured algorithm, not the 2m^3/3 of a full update.
//
// Algorithm (M. Wimmer, ACM TOMS 38 (2012), "pfaffian_LTL"; what torch_pfaffian.pfaffian(sign=True) is
// documented to run -- /root/reference/src/matchcake/utils/_pfaffian.py:22-24):
//   for k = 0, 2, 4, ...:  p = argmax_{j>k} |A[k][j]|;  symmetric swap (k+1 <-> p), sign flip;
//                          Pf *= A[k][k+1];  tau_j = A[k][j] / A[k][k+1];
//                          A[i][j] += A[k+1][i] * tau_j - tau_i * A[k+1][j]   (k+2 <= i < j)
#pragma once
#include "common.cuh"

namespace mcb200 {

// Symmetric row/column transposition (a <-> b), a < b, on upper-triangular storage; rows < k0 are dead.
// Work item t in [k0, m) is handled by the calling thread.
__device__ __forceinline__ void skew_swap_item(double* A, int ld, int a, int b, int t) {
  if (t < a) {
    double x = A[t * ld + a];
    A[t * ld + a] = A[t * ld + b];
    A[t * ld + b] = x;
  } else if (t > a && t < b) {
    double x = A[a * ld + t];
    A[a * ld + t] = -A[t * ld + b];
    A[t * ld + b] = -x;
  } else if (t > b) {
    double x = A[a * ld + t];
    A[a * ld + t] = A[b * ld + t];
    A[b * ld + t] = x;
  } else if (t == b) {
    A[a * ld + b] = -A[a * ld + b];
  }
}

// Whole thread block, matrix shared by the block.  `red` is shared scratch of >= 3 doubles + 2 ints (40 B).
// All threads must call; returns the signed Pfaffian in all threads.
__device__ inline double block_pfaffian_upper(double* A, int m, int ld, double* red) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  if (m == 0) return 1.0;
  if (m & 1) return 0.0;
  int* red_i = reinterpret_cast<int*>(red + 2);
  double pf = 1.0;
  for (int k = 0; k + 1 < m; k += 2) {
    if (tid < 32) {
      double best = -1.0;
      int bj = m;
      for (int j = k + 1 + tid; j < m; j += 32) {
        double a = fabs(A[k * ld + j]);
        if (a > best) {
          best = a;
          bj = j;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        double oa = __shfl_xor_sync(0xffffffffu, best, o);
        int oj = __shfl_xor_sync(0xffffffffu, bj, o);
        if (oa > best || (oa == best && oj < bj)) {
          best = oa;
          bj = oj;
        }
      }
      if (tid == 0) {
        red[0] = best;
        red_i[0] = bj;
      }
    }
    __syncthreads();
    const double best = red[0];
    const int p = red_i[0];
    if (!(best > 0.0)) return 0.0;
    if (p != k + 1) {
      for (int t = k + tid; t < m; t += nthr) skew_swap_item(A, ld, k + 1, p, t);
      pf = -pf;
      __syncthreads();
    }
    const double piv = A[k * ld + k + 1];
    pf *= piv;
    if (k + 2 >= m) break;
    __syncthreads();  // everyone has read piv / row k before it is scaled
    for (int j = k + 2 + tid; j < m; j += nthr) A[k * ld + j] = A[k * ld + j] / piv;
    __syncthreads();
    // trailing update over the strict upper triangle of rows/cols >= k+2
    const int n = m - (k + 2);
    const int items = n * n;
    for (int it = tid; it < items; it += nthr) {
      int i = k + 2 + it / n, j = k + 2 + it % n;
      if (j > i)
        A[i * ld + j] += A[(k + 1) * ld + i] * A[k * ld + j] - A[k * ld + i] * A[(k + 1) * ld + j];
    }
    __syncthreads();
  }
  return pf;
}

}  // namespace mcb200
