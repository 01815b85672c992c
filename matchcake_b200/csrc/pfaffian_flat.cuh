// Branch-free static elimination of FULL matrices (m = 8 MT, a multiple of 32) on top of pfaffian_tiles.cuh: same tile
// layout, same shared row vectors, same 4 x 4 block pivots, same rank-4 DMMA updates, same guard as
// warp_pfaffian_static_steps.
//
// What limits the one-cell-at-a-time Gram kernel (ncu, profiles/r02/ncu_c3_gram_tiles.txt): 3 warps per scheduler, each
// a chain of dependent steps -- rows -> shared vectors -> block pivot + guard (a warp reduction, then a branch) ->
// reciprocal -> coefficients -> fragments -> DMMAs -> rows of the next quadruple; "wait" (fixed-latency dependencies)
// is the largest stall reason and the FP64 units are about half busy.  FlatStep has no branch inside the elimination:
// the guard of a quadruple is only RECORDED, the update is applied regardless, and a matrix whose guard failed anywhere
// is redone from its inputs by the pivoted routine (1.7 % of config C3's cells).  The 2 MT steps of a matrix become one
// basic block: the scheduler sinks the 2 MT warp reductions of the guards off the chain, the reciprocal no longer waits
// for the guard, the B fragment is built from UNSCALED coefficients while the reciprocal is in flight and multiplied by
// 1/q at the end (one more DMUL per fragment entry, four fewer dependent operations per step), |B| is summed as a tree,
// the Newton iteration of 1/q is one cubic step.  2842 -> ~2000 instructions per 64 x 64 cell.
//
// Measured on a 4096-sample Gram of config C3 (profiles/r02/gram_flat_experiments.md): 61.9 ms (gram_tiles_kernel, 12
// warps per SM at 168 registers) -> 59.6 ms (these steps, same shape) -> 57.3 ms (8 warps at 255 registers, every tile
// in registers).  Rejected there: a second, independent elimination interleaved in the same warp (the tail of the
// previous cell next to the head of the current one) -- 57.5 ms with the guard branches, 62.4 ms without -- and a
// column-interleaved layout of the row vectors (-35 % shared-memory wavefronts but 6 selects per fragment entry, 64.0 ms).
#pragma once
#include "pfaffian_tiles.cuh"

namespace mcb200 {

// Largest |entry| (as the high word of the double) of rows R .. R+3 over the columns > R; M a multiple of 32.
template <int M, int R>
__device__ __forceinline__ unsigned static_rows_max_key_full(const double* scratch) {
  constexpr int VS = M + 4;
  const int lane = lane_id();
  unsigned key = 0u;
#pragma unroll
  for (int q = 0; q < M / 32; ++q) {
    if (32 * q + 31 > R) {  // compile time: this group of 32 columns still holds columns > R
      const int cc = lane + 32 * q;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned k2 = (unsigned)__double2hiint(scratch[k * VS + cc]) & 0x7fffffffu;
        const unsigned k3 = cc > R ? k2 : 0u;
        key = k3 > key ? k3 : key;
      }
    }
  }
  return __reduce_max_sync(0xffffffffu, key);
}

template <int MT, int SR, int D>
struct FlatStep {
  using T = Tiles<MT>;
  static_assert(T::M % 32 == 0, "full matrices of 32 k indices");
  static constexpr int R = 4 * D, TR = R >> 3, VS = T::M + 4;
  static constexpr int TI0 = (D & 1) ? TR + 1 : TR;
  static constexpr bool kUpdate = R + 4 < T::M;  // something is left to update after this quadruple

  // (1) rows R .. R+3 -> shared vectors (the caller puts a __syncwarp() between this and run())
  __device__ static __forceinline__ void store(const double (&c)[T::NT][2], const double2* sm, const StaticLane<MT>& L) {
    if (((lane_id() >> 4) & 1) == (D & 1)) static_store_rows<MT, SR, TR, TR>(c, sm, L.st);
  }
  // (2) - (4) block pivot, update; returns whether the guard (growth of the update <= `growth`) held
  __device__ static __forceinline__ bool run(double (&c)[T::NT][2], double2* sm, const double* scratch,
                                             const StaticLane<MT>& L, double& pf, double growth) {
    const double* u = scratch;           // row R      (slot 0)
    const double* v = scratch + 2 * VS;  // row R + 1  (slot 2)
    const double* u2 = scratch + VS;     // row R + 2  (slot 1)
    const double a = u[R + 1];
    const double2 bc = *reinterpret_cast<const double2*>(u + R + 2);
    const double2 de = *reinterpret_cast<const double2*>(v + R + 2);
    const double f = u2[R + 3];
    const double q = fma(a, f, fma(bc.y, de.x, -bc.x * de.y));
    const double sB = ((fabs(a) + fabs(bc.x)) + (fabs(bc.y) + fabs(de.x))) + (fabs(de.y) + fabs(f));
    const double s = __hiloint2double((int)static_rows_max_key_full<T::M, R>(scratch) + 0x100, 0);
    const bool ok = fabs(q) * growth >= s * sB && q != 0.0;
    pf *= q;
    if constexpr (kUpdate) {
      const double invq = fast_rcp3(q);
      const double p0 = L.y_ent[0][R] * L.y_sgn[0], p1 = L.y_ent[1][R] * L.y_sgn[1], p2 = L.y_ent[2][R] * L.y_sgn[2];
      double af[MT], bf[MT];
#pragma unroll
      for (int t = TI0; t < MT; ++t) {
        af[t] = L.a_vec[8 * t];
        bf[t] = fma(p0, L.y_vec[0][8 * t], fma(p1, L.y_vec[1][8 * t], p2 * L.y_vec[2][8 * t])) * invq;
      }
      if constexpr (!(D & 1)) {
        if ((lane_id() >> 2) <= ((R + 3) & 7)) af[TR] = bf[TR] = 0.0;  // indices <= R+3 are dead
      }
      tiles_update_from<MT, SR, TI0>(c, sm, af, bf);
    }
    return ok;
  }
};

// All 2 MT quadruples of one full matrix, no branch.  Returns false when a guard failed (pf and c are then garbage).
template <int MT, int SR, int D = 0>
__device__ __forceinline__ bool warp_pfaffian_flat_steps(double (&c)[Tiles<MT>::NT][2], double* scratch, double2* sm,
                                                         double& pf, const StaticLane<MT>& L, double growth) {
  if constexpr (D >= 2 * MT) {
    return true;
  } else {
    using S = FlatStep<MT, SR, D>;
    S::store(c, sm, L);
    __syncwarp();
    const bool ok = S::run(c, sm, scratch, L, pf, growth);
    __syncwarp();
    return warp_pfaffian_flat_steps<MT, SR, D + 1>(c, scratch, sm, pf, L, growth) && ok;
  }
}

}  // namespace mcb200
