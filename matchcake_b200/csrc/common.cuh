// Shared helpers for the mcb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/mcb200.h"

namespace mcb200 {

extern thread_local std::string g_last_error;
extern std::atomic<long long> g_launch_count;
// Sticky per-thread option (mcb200_set_abs_floor): every magnitude-type output v (|Pf| x scale, probabilities,
// projector expectation values, Gram entries) is returned as sqrt(v^2 + floor2).  0 = off.
extern thread_local double g_abs_floor2;
// Sticky per-thread option (mcb200_set_gram_pivot_growth): growth bound of the static block pivots of mcb200_gram for
// d <= 64 (pfaffian_tiles.cuh: kStaticBlockGrowthPhysical = 1e5 by default, kStaticBlockGrowth = 1e3 = strict).
extern thread_local double g_gram_growth;

int set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int check_launch(const char* what);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kWarp = 32;
constexpr int kNumSMs = 148;  // B200
constexpr int kMaxDynSmem = 227 * 1024;

// torch_pfaffian.PfaffianDet returns sqrt(|det A| + EPSILON) = sqrt(Pf^2 + EPSILON) (utils/_pfaffian.py:110-118)
__device__ __forceinline__ double abs_floor(double v, double floor2) {
  return floor2 > 0.0 ? sqrt(fma(v, v, floor2)) : fabs(v);
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// index (with +1 padding per row) into a warp/CTA private square matrix held in shared memory
__device__ __forceinline__ int sidx(int i, int j, int ld) { return i * ld + j; }

// warp argmax of |v| over lanes; returns lane of the (first) maximum and broadcasts nothing else.
__device__ __forceinline__ int warp_argmax_abs(double v, int idx, double* vmax_out) {
  double a = fabs(v);
  int bi = idx;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double oa = __shfl_xor_sync(0xffffffffu, a, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oa > a || (oa == a && oi < bi)) {
      a = oa;
      bi = oi;
    }
  }
  *vmax_out = a;
  return bi;
}

}  // namespace mcb200
