// Register-resident Pfaffian of a skew-symmetric matrix of size m <= 8*MT, one warp per matrix, on the FP64
// tensor pipe.
//
// Layout: the matrix is cut into 8x8 tiles; the warp keeps the NT = MT(MT+1)/2 upper tiles (diagonal tiles in
// full) as mma.sync.m8n8k4.f64 accumulator fragments: lane (gr = lane>>2, tg = lane&3) holds, for tile (ti,tj),
// c[t][0] = A[8ti+gr][8tj+2tg] and c[t][1] = A[8ti+gr][8tj+2tg+1].  72 doubles per lane for m = 64.
//
// Algorithm: Parlett-Reid with partial pivoting done LOGICALLY -- nothing is ever swapped.  `active` is the set
// of rows/columns not yet eliminated; each step takes r = lowest active index, p = argmax_j |A[r][j]| over the
// other active j, multiplies Pf by A[r][p] and applies the skew rank-2 update
//       A[i][j] += tau_i v_j - v_i tau_j,   tau_j = A[r][j] / A[r][p],  v_j = A[j][p]   (0 on inactive j)
// to the tiles that still hold active rows, as ONE DMMA per tile: A-fragment columns (tau, -v, 0, 0), B-fragment
// rows (v; tau; 0; 0).  The sign of the permutation (r_1,p_1,r_2,p_2,...) is tracked with a population count.
// Row r and row/column p are pulled out of the fragments through a small per-warp shared-memory scratch; the
// tile-row / tile-column they live in is warp-uniform, so the extraction is a switch over MT cases and only the
// taken case issues instructions.  The pivot search is one REDUX over the high words of |A[r][j]|.
// Reference semantics: torch_pfaffian.pfaffian(sign=True) (Parlett-Reid) -- utils/_pfaffian.py:17-36,109.
#pragma once
#include "common.cuh"

namespace mcb200 {

template <int MT>
struct Tiles {
  static constexpr int NT = MT * (MT + 1) / 2;
  static constexpr int M = 8 * MT;
  // scratch doubles per warp: rowbuf[M] | pbuf[M] | 4 fragment vectors of M + 8 (the +8 shifts consecutive vectors
  // by half a 128-byte bank line, so the two vectors read by one fragment load never collide)
  static constexpr int kVec = M + 8;
  static constexpr int kScratch = 2 * M + 4 * kVec;
  __host__ __device__ static constexpr int idx(int ti, int tj) { return ti * MT - (ti * (ti - 1)) / 2 + (tj - ti); }
};

__device__ __forceinline__ void dmma_acc(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// Tile access.  The first SR tile-rows (the ones that die first as the elimination proceeds) may live in shared
// memory instead of registers -- fragment layout, sm[tile * 32] is this lane's (c0, c1) -- which cuts the register
// footprint of a 64 x 64 matrix from 144 to 84 and doubles the number of resident warps.
template <int MT, int SR, int TI, int TJ>
__device__ __forceinline__ double2 tile_get(const double (&c)[Tiles<MT>::NT][2], const double2* sm) {
  if constexpr (TI < SR) return sm[Tiles<MT>::idx(TI, TJ) * 32];
  else return make_double2(c[Tiles<MT>::idx(TI, TJ)][0], c[Tiles<MT>::idx(TI, TJ)][1]);
}

template <int MT, int SR, int TI, int TJ0 = TI>
__device__ __forceinline__ void tiles_store_row(const double (&c)[Tiles<MT>::NT][2], const double2* sm, double* row_st) {
  if constexpr (TJ0 < MT) {
    *reinterpret_cast<double2*>(row_st + 8 * TJ0) = tile_get<MT, SR, TI, TJ0>(c, sm);
    tiles_store_row<MT, SR, TI, TJ0 + 1>(c, sm, row_st);
  }
}

template <int MT, int SR, int TI>
__device__ __forceinline__ void tiles_store_row_right(const double (&c)[Tiles<MT>::NT][2], const double2* sm,
                                                      double* row_st) {
  tiles_store_row<MT, SR, TI, TI + 1>(c, sm, row_st);
}

template <int MT, int SR, int TJ, int TI0 = 0>
__device__ __forceinline__ void tiles_store_col(const double (&c)[Tiles<MT>::NT][2], const double2* sm, double* col_st,
                                                int e) {
  if constexpr (TI0 <= TJ) {
    const double2 v = tile_get<MT, SR, TI0, TJ>(c, sm);
    col_st[8 * TI0] = e ? v.y : v.x;
    tiles_store_col<MT, SR, TJ, TI0 + 1>(c, sm, col_st, e);
  }
}

template <int MT, int SR, int TI, int TJ>
__device__ __forceinline__ void tile_update(double (&c)[Tiles<MT>::NT][2], double2* sm, double a, double b) {
  if constexpr (TI < SR) {
    const double2 v = sm[Tiles<MT>::idx(TI, TJ) * 32];
    double t[2] = {v.x, v.y};
    dmma_acc(t, a, b);
    sm[Tiles<MT>::idx(TI, TJ) * 32] = make_double2(t[0], t[1]);
  } else {
    dmma_acc(c[Tiles<MT>::idx(TI, TJ)], a, b);
  }
}

template <int MT, int SR, int TI, int TJ>
__device__ __forceinline__ void tiles_update_row(double (&c)[Tiles<MT>::NT][2], double2* sm, const double (&af)[MT],
                                                 const double (&bf)[MT]) {
  if constexpr (TJ < MT) {
    tile_update<MT, SR, TI, TJ>(c, sm, af[TI], bf[TJ]);
    tiles_update_row<MT, SR, TI, TJ + 1>(c, sm, af, bf);
  }
}

template <int MT, int SR, int TI>
__device__ __forceinline__ void tiles_update_from(double (&c)[Tiles<MT>::NT][2], double2* sm, const double (&af)[MT],
                                                  const double (&bf)[MT]) {
  if constexpr (TI < MT) {
    tiles_update_row<MT, SR, TI, TI>(c, sm, af, bf);
    tiles_update_from<MT, SR, TI + 1>(c, sm, af, bf);
  }
}

#define MCB200_TILE_SWITCH(var, MTv, CALL)            \
  switch (var) {                                      \
    case 0: { constexpr int K_ = 0; if constexpr (K_ < MTv) { CALL; } } break; \
    case 1: { constexpr int K_ = 1; if constexpr (K_ < MTv) { CALL; } } break; \
    case 2: { constexpr int K_ = 2; if constexpr (K_ < MTv) { CALL; } } break; \
    case 3: { constexpr int K_ = 3; if constexpr (K_ < MTv) { CALL; } } break; \
    case 4: { constexpr int K_ = 4; if constexpr (K_ < MTv) { CALL; } } break; \
    case 5: { constexpr int K_ = 5; if constexpr (K_ < MTv) { CALL; } } break; \
    case 6: { constexpr int K_ = 6; if constexpr (K_ < MTv) { CALL; } } break; \
    default: { constexpr int K_ = 7; if constexpr (K_ < MTv) { CALL; } } break; \
  }

// c: fragments of the upper tiles (entries at rows/cols >= m must be 0); with SR > 0 the tiles of the first SR
// tile-rows are read from / written to sm (this lane's slot, see tile_get) instead.  scratch: Tiles<MT>::kScratch
// doubles private to the warp.  Returns the signed Pfaffian in every lane.
// `start` (even): rows/columns below it have already been eliminated (by warp_pfaffian_static_steps); the result is
// then the Pfaffian of the remaining Schur complement.
template <int MT, int SR = 0>
__device__ double warp_pfaffian_tiles_pivoted(double (&c)[Tiles<MT>::NT][2], int m, double* scratch,
                                              double2* sm = nullptr, int start = 0) {
  using T = Tiles<MT>;
  constexpr int H = (T::M + 31) / 32;  // vector entries / pivot candidates per lane
  const int lane = lane_id(), gr = lane >> 2, tg = lane & 3;
  if (m == 0) return 1.0;
  if (m & 1) return 0.0;
  double* rowbuf = scratch;
  double* pbuf = scratch + T::M;
  // fragment vectors.  The rank-2 update A += tau v^T - v tau^T is issued as a K = 4 DMMA with
  //   A-fragment columns (tau/2, -v/2, tau/2, -v/2)  and  B-fragment rows (v; tau; v; tau),
  // so that every lane reads one of two vectors per fragment (even / odd tg) and the two 64-byte groups of a
  // fragment load fall into different halves of the 128-byte bank line (conflict free, broadcast within a group).
  double* tauh = scratch + 2 * T::M;          // tau / 2      (A, even tg)
  double* negvh = tauh + T::kVec;             // -v / 2       (A, odd tg)
  double* vv = negvh + T::kVec;               // v            (B, even tg)
  double* tauv = vv + T::kVec;                // tau          (B, odd tg)
  // per-lane store / load addresses (loop invariant)
  double* row_st = rowbuf + 2 * tg;   // row r -> rowbuf[8 tj + 2 tg .. +1]
  double* prow_st = pbuf + 2 * tg;    // row p -> pbuf[8 tj + 2 tg .. +1]
  double* pcol_st = pbuf + gr;        // column p -> pbuf[8 ti + gr]
  const double* srcA = ((tg & 1) ? negvh : tauh) + gr;
  const double* srcB = ((tg & 1) ? tauv : vv) + gr;
  // active set as two 32-bit words (uniform) + this lane's own entries j = lane, lane + 32
  unsigned act_lo = (m >= 32) ? 0xffffffffu : ((1u << m) - 1u);
  unsigned act_hi = (m >= 64) ? 0xffffffffu : ((m > 32) ? ((1u << (m - 32)) - 1u) : 0u);
  if (start > 0) {
    const unsigned long long keep = ~((1ull << start) - 1ull);
    act_lo &= (unsigned)keep;
    act_hi &= (unsigned)(keep >> 32);
  }
  bool on0 = lane < m && lane >= start, on1 = lane + 32 < m && lane + 32 >= start;
  const int j1 = (lane + 32 < T::M) ? lane + 32 : lane;  // second entry of this lane (only used when on1)
  double pf = 1.0;
  unsigned parity = 0;
  const int steps = (m - start) >> 1;
  for (int s = 0; s < steps; ++s) {
    const int r = act_lo ? (__ffs((int)act_lo) - 1) : (31 + __ffs((int)act_hi));
    const int tr = r >> 3;
    // (1) row r, columns in tile-columns >= tr
    if (gr == (r & 7)) {
      MCB200_TILE_SWITCH(tr, MT, (tiles_store_row<MT, SR, K_>(c, sm, row_st)))
    }
    if (lane == (r & 31)) {  // r leaves the active set before the search
      if (r < 32) on0 = false; else on1 = false;
    }
    __syncwarp();
    // (2) pivot: largest |A[r][j]| over the other active j; compare the high 32 bits of the doubles
    unsigned key = on0 ? ((unsigned)__double2hiint(rowbuf[lane]) & 0x7fffffffu) : 0u;
    int bj = lane;
    if (H > 1) {
      const unsigned k2 = on1 ? ((unsigned)__double2hiint(rowbuf[j1]) & 0x7fffffffu) : 0u;
      if (k2 > key) {
        key = k2;
        bj = j1;
      }
    }
    const unsigned kmax = __reduce_max_sync(0xffffffffu, key);
    if (kmax == 0u) return 0.0;  // singular (every candidate is zero / denormal): Pf = 0
    const unsigned ball = __ballot_sync(0xffffffffu, key == kmax);
    const int p = __shfl_sync(0xffffffffu, bj, __ffs((int)ball) - 1);
    const int tp = p >> 3;
    const double piv = rowbuf[p];
    // (3) v_j = A[j][p]: column p of the tiles above/at the diagonal; row p (NOT negated here) to the right
    if (tg == ((p & 7) >> 1)) {
      MCB200_TILE_SWITCH(tp, MT, (tiles_store_col<MT, SR, K_>(c, sm, pcol_st, p & 1)))
    }
    if (gr == (p & 7)) {
      MCB200_TILE_SWITCH(tp, MT, (tiles_store_row_right<MT, SR, K_>(c, sm, prow_st)))
    }
    if (lane == (p & 31)) {
      if (p < 32) on0 = false; else on1 = false;
    }
    // sign: number of active indices strictly between r and p
    {
      const unsigned long long act = ((unsigned long long)act_hi << 32) | act_lo;
      parity ^= (unsigned)__popcll(act & ((1ull << p) - 1ull) & ~((2ull << r) - 1ull));
      const unsigned long long nact = act & ~((1ull << r) | (1ull << p));
      act_lo = (unsigned)nact;
      act_hi = (unsigned)(nact >> 32);
    }
    pf *= piv;
    __syncwarp();
    if (s == steps - 1) break;
    // (4) masked vectors tau_j = A[r][j] / piv, v_j = A[j][p] (minus sign for the part read from row p)
    const double inv = 1.0 / piv;
    {
      const double t0 = on0 ? rowbuf[lane] * inv : 0.0;
      double v0 = on0 ? pbuf[lane] : 0.0;
      if ((lane >> 3) > tp) v0 = -v0;
      if (lane < T::M) {
        tauh[lane] = 0.5 * t0;
        negvh[lane] = -0.5 * v0;
        vv[lane] = v0;
        tauv[lane] = t0;
      }
      if (H > 1) {
        const int j = j1;
        const double t1 = on1 ? rowbuf[j] * inv : 0.0;
        double v1 = on1 ? pbuf[j] : 0.0;
        if ((j >> 3) > tp) v1 = -v1;
        if (lane + 32 < T::M) {
          tauh[j] = 0.5 * t1;
          negvh[j] = -0.5 * v1;
          vv[j] = v1;
          tauv[j] = t1;
        }
      }
    }
    __syncwarp();
    double af[MT], bf[MT];
#pragma unroll
    for (int t = 0; t < MT; ++t) {
      af[t] = srcA[8 * t];
      bf[t] = srcB[8 * t];
    }
    // (5) rank-2 update on the tensor pipe: one DMMA per live upper tile
    MCB200_TILE_SWITCH(tr, MT, (tiles_update_from<MT, SR, K_>(c, sm, af, bf)))
    __syncwarp();
  }
  return (parity & 1u) ? -pf : pf;
}

// ---------------------------------------------------------------------------------------------------------------------
// Static-order fast path.  Pair S = (2S, 2S+1) is eliminated with pivot A[2S][2S+1] -- no search, no permutation,
// every tile / lane index is a compile-time constant, so a step is: predicated row stores, a reciprocal, 2 loads + 2
// multiplies per live tile index and one DMMA per live tile (several times fewer instructions than the pivoted
// step), and two pairs share one pass over the tiles (rank-4 update, see warp_pfaffian_static_steps).  For the matrices of this path (Lambda|_W + Lambda_y, Lambda_i + Lambda_j) the natural pivot is the
// conditional probability of the target outcome and is rarely small.  A guard keeps the accuracy of partial pivoting:
// the step is taken only if |pivot| >= kStaticPivotRatio * max_j |A[2S][j]| (multipliers <= 1/ratio); otherwise the
// routine stops and reports how many indices it eliminated, and the pivoted routine finishes the remaining Schur
// complement (warp_pfaffian_tiles_pivoted(..., start)).
constexpr double kStaticPivotRatio = 1e-3;

// 1 / x for a normal, non-zero x (the guard has just checked it): hardware seed (~20 bits) + two Newton steps instead
// of the IEEE division sequence with its special-case handling; full double accuracy, half the instructions.
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

template <int MT, int SR, int TR, int TJ>
__device__ __forceinline__ void static_store_rows(const double (&c)[Tiles<MT>::NT][2], const double2* sm, double* dst) {
  if constexpr (TJ < MT) {
    *reinterpret_cast<double2*>(dst + 8 * TJ) = tile_get<MT, SR, TR, TJ>(c, sm);
    static_store_rows<MT, SR, TR, TJ + 1>(c, sm, dst);
  }
}

// Guard of one static pivot: |piv| against the largest |row[j]|, j > last (or against the bound 2 when BOUNDED: the
// caller guarantees |A[i][j]| <= 2 for the matrix and all its Schur complements -- Lambda|_W + Lambda_y of a physical
// state, whose conditional covariance after each projected wire is again a covariance).
template <int M, bool BOUNDED>
__device__ __forceinline__ bool static_pivot_ok(double piv, const double* row, int last) {
  if constexpr (BOUNDED) {
    return fabs(piv) >= 2.0 * kStaticPivotRatio;
  } else {
    const int lane = lane_id();
    unsigned key = 0u;
    if (lane > last && lane < M) key = (unsigned)__double2hiint(row[lane]) & 0x7fffffffu;
    if constexpr (M > 32) {
      if (lane + 32 < M && lane + 32 > last) {
        const unsigned k2 = (unsigned)__double2hiint(row[lane + 32]) & 0x7fffffffu;
        key = k2 > key ? k2 : key;
      }
    }
    const unsigned kmax = __reduce_max_sync(0xffffffffu, key);
    return fabs(piv) >= kStaticPivotRatio * __hiloint2double((int)kmax, 0) && piv != 0.0;
  }
}

// Double step D eliminates the pairs (4D, 4D+1) and (4D+2, 4D+3) and applies BOTH rank-2 updates to the trailing tiles
// as one K = 4 DMMA per tile:  A += w tau^T - tau w^T + w' tau'^T - tau' w'^T  with
//   A-fragment columns (w, -tau, w', -tau'),  B-fragment rows (tau; w; tau'; w')
// (tau = row 4D / piv1, w = row 4D+1; tau', w' the same for rows 4D+2, 4D+3 AFTER the first pair's update, which is
// applied to those two rows only, as vectors in shared memory).  Half the tensor instructions of two rank-2 steps and
// no wasted K lanes.  Returns the number of eliminated indices (m when the whole matrix went through); pf is
// multiplied by the pivots taken.
template <int MT, int SR, bool BOUNDED = false, int D = 0>
__device__ __forceinline__ int warp_pfaffian_static_steps(double (&c)[Tiles<MT>::NT][2], int m, double* scratch,
                                                          double2* sm, double& pf) {
  using T = Tiles<MT>;
  if constexpr (4 * D >= T::M) {
    return m;
  } else {
    constexpr int R = 4 * D, TR = R >> 3;
    constexpr int VS = T::M + 4;  // vector stride: the 4 vectors start 4 doubles apart modulo a 128-byte bank line
    if (R >= m) return m;
    const int lane = lane_id(), gr = lane >> 2, tg = lane & 3;
    // row R + k lives in slot brev2(k): consecutive rows are then 8 doubles apart modulo the bank line, so the 16-byte
    // row stores of a quarter warp (two rows) as well as the 8-byte fragment loads of a half warp (four vectors, four
    // consecutive entries each) are conflict free
    double* u = scratch;             // row R
    double* v = scratch + 2 * VS;    // row R + 1
    double* u2 = scratch + VS;       // row R + 2
    double* v2 = scratch + 3 * VS;   // row R + 3
    // (1) rows R .. R+3 of tile row TR -> shared vectors (held by the lanes with gr >> 2 == D & 1)
    if ((gr >> 2) == (D & 1))
      static_store_rows<MT, SR, TR, TR>(c, sm, scratch + (((gr & 1) << 1) | ((gr >> 1) & 1)) * VS + 2 * tg);
    __syncwarp();
    const double piv1 = u[R + 1];
    if (!static_pivot_ok<T::M, BOUNDED>(piv1, u, R + 1)) return R;
    pf *= piv1;
    if (R + 2 >= m) return m;  // last pair: nothing left to update
    const double inv1 = fast_rcp(piv1);
    // (2) first pair's update on rows R+2, R+3 only:  x[c] += w_a tau_c - tau_a w_c
    {
      const double t2 = u[R + 2] * inv1, t3 = u[R + 3] * inv1, w2 = v[R + 2], w3 = v[R + 3];
#pragma unroll
      for (int q = (R + 2) / 32; q < (T::M + 31) / 32; ++q) {  // entries below R + 2 are dead
        const int cc = lane + 32 * q;
        if (cc >= R + 2 && cc < T::M) {
          const double tc = u[cc] * inv1, wc = v[cc];
          u2[cc] += w2 * tc - t2 * wc;
          v2[cc] += w3 * tc - t3 * wc;
        }
      }
    }
    __syncwarp();
    const double piv2 = u2[R + 3];
    const bool odd = tg & 1;
    double af[MT], bf[MT];
    if (!static_pivot_ok<T::M, BOUNDED>(piv2, u2, R + 3)) {
      // second pivot rejected: commit the first pair alone (rank-2 update with both K halves carrying it) and stop
      const double sa = odd ? -0.5 * inv1 : 0.5, sb = odd ? 1.0 : inv1;
      const double* srcA = (odd ? u : v) + gr;
      const double* srcB = (odd ? v : u) + gr;
#pragma unroll
      for (int t = TR; t < MT; ++t) {
        af[t] = sa * srcA[8 * t];
        bf[t] = sb * srcB[8 * t];
      }
      if (gr <= ((R + 1) & 7)) af[TR] = bf[TR] = 0.0;
      tiles_update_from<MT, SR, TR>(c, sm, af, bf);
      __syncwarp();
      return R + 2;
    }
    pf *= piv2;
    if (R + 4 >= m) return m;
    const double inv2 = fast_rcp(piv2);
    // (3) fragments: lane tg reads vector {w, u, w', u'}[tg] for A and {u, w, u', w'}[tg] for B
    const double sa = odd ? ((tg & 2) ? -inv2 : -inv1) : 1.0;
    const double sb = odd ? 1.0 : ((tg & 2) ? inv2 : inv1);
    const int ka = tg ^ 1;
    const double* srcA = scratch + (((ka & 1) << 1) | (ka >> 1)) * VS + gr;
    const double* srcB = scratch + (((tg & 1) << 1) | (tg >> 1)) * VS + gr;
#pragma unroll
    for (int t = TR; t < MT; ++t) {
      af[t] = sa * srcA[8 * t];
      bf[t] = sb * srcB[8 * t];
    }
    if (gr <= ((R + 3) & 7)) af[TR] = bf[TR] = 0.0;  // indices <= R+3 are dead: they neither change nor contribute
    // (4) one DMMA per live upper tile; tile row TR is dead once its second quadruple has been eliminated
    constexpr int TI0 = (D & 1) ? TR + 1 : TR;
    tiles_update_from<MT, SR, TI0>(c, sm, af, bf);
    __syncwarp();
    return warp_pfaffian_static_steps<MT, SR, BOUNDED, D + 1>(c, m, scratch, sm, pf);
  }
}

// c: fragments of the upper tiles (entries at rows/cols >= m must be 0); with SR > 0 the tiles of the first SR
// tile-rows are read from / written to sm (this lane's slot, see tile_get) instead.  scratch: Tiles<MT>::kScratch
// doubles private to the warp.  Returns the signed Pfaffian in every lane.
template <int MT, int SR = 0, bool BOUNDED = false>
__device__ double warp_pfaffian_tiles(double (&c)[Tiles<MT>::NT][2], int m, double* scratch, double2* sm = nullptr) {
  if (m == 0) return 1.0;
  if (m & 1) return 0.0;
  double pf = 1.0;
  const int done = warp_pfaffian_static_steps<MT, SR, BOUNDED>(c, m, scratch, sm, pf);
  if (done < m) pf *= warp_pfaffian_tiles_pivoted<MT, SR>(c, m, scratch, sm, done);
  return pf;
}

}  // namespace mcb200
