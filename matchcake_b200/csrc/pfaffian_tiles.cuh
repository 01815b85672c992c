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
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
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
// Static-order fast path: 4 x 4 BLOCK pivots.  Quadruple D eliminates the indices R .. R+3 (R = 4D) with the leading
// 4 x 4 block B = A[R:R+4, R:R+4] as the pivot -- no search, no permutation, every tile / lane index a compile-time
// constant.  With  a = B01, b = B02, c = B03, d = B12, e = B13, f = B23:
//      q = Pf(B) = a f - b e + c d,        B^-1 = 1/q [[0,-f,e,-d],[f,0,-c,b],[-e,c,0,-a],[d,-b,a,0]],
//      Pf(A) = q Pf(A'),                   A' = A_22 + r^T (B^-1 r)      (r = the four rows R .. R+3, trailing columns),
// so the whole rank-4 update is ONE K = 4 DMMA per live upper tile with A-fragment columns r_k and B-fragment rows
// y_k = (B^-1 r)_k.  Lane tg = k builds its own row y_k on the fly while it loads the B fragment: three scalars
// (block entries times 1/q) times three of the four row vectors -- one reciprocal, one warp barrier and no second pass
// over the vectors per quadruple (the sequential variant it replaces spent two reciprocals, two barriers, a rank-2
// update of rows R+2, R+3 in shared memory and a multiply per fragment entry).  For the matrices of this path
// (Lambda|_W + Lambda_y, Lambda_i + Lambda_j) q is a product of two conditional probabilities of the target outcome
// and is rarely small.  A guard keeps the accuracy of partial pivoting: the quadruple is taken only if the entries
// of the update r^T B^-1 r cannot exceed kStaticBlockGrowth times the current scale,  s sB / |q| <= growth  with
// s = largest |entry| of the four rows (2 when BOUNDED) and sB = sum of the |entries| of B (random Gaussian matrices:
// same accuracy as scalar pivots guarded at 1e-3 of their row, measured); otherwise the routine stops, reports how many indices it eliminated, and the pivoted routine finishes the remaining
// Schur complement (warp_pfaffian_tiles_pivoted(..., start)).  A block pivot is also the more robust choice: it does
// not care whether A[R][R+1] alone is small.
constexpr double kStaticBlockGrowth = 1e3;           // arbitrary matrices (utils.pfaffian, sector features)
// Lambda|_W + Lambda_y and Lambda_i + Lambda_j: the Schur complements of these sums of covariances stay O(1), the bound
// s sB / |q| only measures the rounding amplification of the update (<= 1e5 eps); on config C3's cells it reaches 6e4
// at the 99.9 % quantile with a worst error of 1e-14 (measured), and 1e3 would hand 12 % of the cells to the pivoted
// routine for nothing.
constexpr double kStaticBlockGrowthPhysical = 1e5;
constexpr double kStaticPivotRatio = 1e-3;  // scalar-pivot guard of the block-per-matrix variant (pfaffian_block_tiles.cu)

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

// The same with one cubic step y (1 + e + e^2), e = 1 - x y ~ 2^-20 from the seed: relative error e^3, three dependent
// operations after the seed instead of four.
__device__ __forceinline__ double fast_rcp3(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  const double t = fma(e, e, e);
  return fma(y, t, y);
}

template <int MT, int SR, int TR, int TJ>
__device__ __forceinline__ void static_store_rows(const double (&c)[Tiles<MT>::NT][2], const double2* sm, double* dst) {
  if constexpr (TJ < MT) {
    *reinterpret_cast<double2*>(dst + 8 * TJ) = tile_get<MT, SR, TR, TJ>(c, sm);
    static_store_rows<MT, SR, TR, TJ + 1>(c, sm, dst);
  }
}

// Per-lane constants of the static path (the same for every quadruple): row R + k lives in vector slot brev2(k), so
// that consecutive rows are 8 doubles apart modulo the 128-byte bank line (vector stride M + 4): the 16-byte row stores
// of a quarter warp and the 8-byte fragment loads of a half warp are conflict free.
template <int MT>
struct StaticLane {
  const double* a_vec;     // A fragment: row R + tg at column gr
  const double* y_vec[3];  // the three rows != tg at column gr (B fragment: y_tg = sum_j coef_j * row_j)
  const double* y_ent[3];  // block entry behind coef_j: slot row + column offset (add R)
  double y_sgn[3];         // its sign in B^-1
  double* st;              // where this lane stores its part of row R + (gr & 3)
};

__host__ __device__ constexpr int static_slot(int k) { return ((k & 1) << 1) | (k >> 1); }

template <int MT>
__device__ __forceinline__ StaticLane<MT> static_lane_setup(double* scratch) {
  constexpr int VS = Tiles<MT>::M + 4;
  const int lane = lane_id(), gr = lane >> 2, tg = lane & 3;
  StaticLane<MT> L;
  L.a_vec = scratch + static_slot(tg) * VS + gr;
  L.st = scratch + static_slot(gr & 3) * VS + 2 * tg;
  // y_0 = (-f r1 + e r2 - d r3)/q, y_1 = (f r0 - c r2 + b r3)/q, y_2 = (-e r0 + c r1 - a r3)/q, y_3 = (d r0 - b r1 + a r2)/q
  // entries: a = row0[R+1], b = row0[R+2], c = row0[R+3], d = row1[R+2], e = row1[R+3], f = row2[R+3]
  int row[3], erow[3], ecol[3];
  if (tg == 0) {
    row[0] = 1; erow[0] = 2; ecol[0] = 3; L.y_sgn[0] = -1.0;  // f
    row[1] = 2; erow[1] = 1; ecol[1] = 3; L.y_sgn[1] = 1.0;   // e
    row[2] = 3; erow[2] = 1; ecol[2] = 2; L.y_sgn[2] = -1.0;  // d
  } else if (tg == 1) {
    row[0] = 0; erow[0] = 2; ecol[0] = 3; L.y_sgn[0] = 1.0;   // f
    row[1] = 2; erow[1] = 0; ecol[1] = 3; L.y_sgn[1] = -1.0;  // c
    row[2] = 3; erow[2] = 0; ecol[2] = 2; L.y_sgn[2] = 1.0;   // b
  } else if (tg == 2) {
    row[0] = 0; erow[0] = 1; ecol[0] = 3; L.y_sgn[0] = -1.0;  // e
    row[1] = 1; erow[1] = 0; ecol[1] = 3; L.y_sgn[1] = 1.0;   // c
    row[2] = 3; erow[2] = 0; ecol[2] = 1; L.y_sgn[2] = -1.0;  // a
  } else {
    row[0] = 0; erow[0] = 1; ecol[0] = 2; L.y_sgn[0] = 1.0;   // d
    row[1] = 1; erow[1] = 0; ecol[1] = 2; L.y_sgn[1] = -1.0;  // b
    row[2] = 2; erow[2] = 0; ecol[2] = 1; L.y_sgn[2] = 1.0;   // a
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    L.y_vec[j] = scratch + static_slot(row[j]) * VS + gr;
    L.y_ent[j] = scratch + static_slot(erow[j]) * VS + ecol[j];
  }
  return L;
}

// Largest |entry| (as the high word of the double) of rows R .. R+3 over the columns > last.
template <int M>
__device__ __forceinline__ unsigned static_rows_max_key(const double* scratch, int last) {
  constexpr int VS = M + 4;
  const int lane = lane_id();
  unsigned key = 0u;
#pragma unroll
  for (int q = 0; q < (M + 31) / 32; ++q) {
    const int cc = lane + 32 * q;
    if (cc > last && cc < M) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned k2 = (unsigned)__double2hiint(scratch[k * VS + cc]) & 0x7fffffffu;
        key = k2 > key ? k2 : key;
      }
    }
  }
  return __reduce_max_sync(0xffffffffu, key);
}

// Returns the number of eliminated indices (m when the whole matrix went through); pf is multiplied by the pivots taken.
template <int MT, int SR, bool BOUNDED = false, bool PHYSICAL = BOUNDED, int D = 0>
__device__ __forceinline__ int warp_pfaffian_static_steps(double (&c)[Tiles<MT>::NT][2], int m, double* scratch,
                                                          double2* sm, double& pf, const StaticLane<MT>& L) {
  using T = Tiles<MT>;
  if constexpr (4 * D >= T::M) {
    return m;
  } else {
    constexpr int R = 4 * D, TR = R >> 3;
    constexpr int VS = T::M + 4;
    if (R >= m) return m;
    const int lane = lane_id(), gr = lane >> 2;
    const double* u = scratch;             // row R      (slot 0)
    const double* v = scratch + 2 * VS;    // row R + 1  (slot 2)
    const double* u2 = scratch + VS;       // row R + 2  (slot 1)
    // (1) rows R .. R+3 of tile row TR -> shared vectors (held by the lanes with gr >> 2 == D & 1)
    if ((gr >> 2) == (D & 1)) static_store_rows<MT, SR, TR, TR>(c, sm, L.st);
    __syncwarp();
    const double a = u[R + 1];
    if (R + 2 >= m) {  // last pair on its own (m = 2 mod 4): nothing left to update, any pivot will do
      pf *= a;
      return m;
    }
    // (2) block pivot
    const double2 bc = *reinterpret_cast<const double2*>(u + R + 2);
    const double2 de = *reinterpret_cast<const double2*>(v + R + 2);
    const double f = u2[R + 3];
    const double q = fma(a, f, fma(bc.y, de.x, -bc.x * de.y));
    constexpr int TI0 = (D & 1) ? TR + 1 : TR;  // tile row TR is dead once its second quadruple has been eliminated
    double af[MT], bf[MT];
#ifndef MCB200_STATIC_CHAIN_OLD
    // The chain per quadruple is what bounds the callers (2 - 5 warps per scheduler): the reciprocal (one cubic Newton
    // step, pfaffian_flat.cuh) does not wait for the guard's branch, |B| is summed as a tree, and the B fragment is
    // built from the unscaled block entries while the reciprocal is in flight, then multiplied by 1/q -- one more DMUL
    // per fragment entry, four fewer dependent operations (measured on config C3's Gram kernel: 61.9 -> 59.6 ms).
    const double invq = fast_rcp3(q);  // q = 0 gives inf / NaN here, which the guard below keeps from being used
    // guard: growth of the update  ||r||^2 ||B^-1|| / s  <=  s sB / |q|  <=  kStaticBlockGrowth
    const double sB = ((fabs(a) + fabs(bc.x)) + (fabs(bc.y) + fabs(de.x))) + (fabs(de.y) + fabs(f));  // >= max |B entry|
    double s = 2.0;
    if constexpr (!BOUNDED)
      s = __hiloint2double((int)static_rows_max_key<T::M>(scratch, R) + 0x100, 0);  // >= max |entry| of the four rows
    const bool ok = fabs(q) * (PHYSICAL ? kStaticBlockGrowthPhysical : kStaticBlockGrowth) >= s * sB && q != 0.0;
    if (!ok) return R;
    pf *= q;
    if (R + 4 >= m) return m;
    // (3) fragments: A = row R + tg, B = y_tg = (sum_j (+-B entry_j) * row_j) / q
    const double p0 = L.y_ent[0][R] * L.y_sgn[0], p1 = L.y_ent[1][R] * L.y_sgn[1], p2 = L.y_ent[2][R] * L.y_sgn[2];
#pragma unroll
    for (int t = TI0; t < MT; ++t) {
      af[t] = L.a_vec[8 * t];
      bf[t] = fma(p0, L.y_vec[0][8 * t], fma(p1, L.y_vec[1][8 * t], p2 * L.y_vec[2][8 * t])) * invq;
    }
#else
    // guard: growth of the update  ||r||^2 ||B^-1|| / s  <=  s sB / |q|  <=  kStaticBlockGrowth
    const double sB = fabs(a) + fabs(bc.x) + fabs(bc.y) + fabs(de.x) + fabs(de.y) + fabs(f);  // >= max |B entry|
    double s = 2.0;
    if constexpr (!BOUNDED)
      s = __hiloint2double((int)static_rows_max_key<T::M>(scratch, R) + 0x100, 0);  // >= max |entry| of the four rows
    const bool ok = fabs(q) * (PHYSICAL ? kStaticBlockGrowthPhysical : kStaticBlockGrowth) >= s * sB && q != 0.0;
    if (!ok) return R;
    pf *= q;
    if (R + 4 >= m) return m;
    const double invq = fast_rcp(q);
    // (3) fragments: A = row R + tg, B = y_tg = sum_j coef_j * row_j
    const double k0 = L.y_ent[0][R] * (L.y_sgn[0] * invq), k1 = L.y_ent[1][R] * (L.y_sgn[1] * invq),
                 k2 = L.y_ent[2][R] * (L.y_sgn[2] * invq);
#pragma unroll
    for (int t = TI0; t < MT; ++t) {
      af[t] = L.a_vec[8 * t];
      bf[t] = fma(k0, L.y_vec[0][8 * t], fma(k1, L.y_vec[1][8 * t], k2 * L.y_vec[2][8 * t]));
    }
#endif
    if constexpr (!(D & 1)) {
      if (gr <= ((R + 3) & 7)) af[TR] = bf[TR] = 0.0;  // indices <= R+3 are dead: they neither change nor contribute
    }
    // (4) one DMMA per live upper tile
    tiles_update_from<MT, SR, TI0>(c, sm, af, bf);
    __syncwarp();
    return warp_pfaffian_static_steps<MT, SR, BOUNDED, PHYSICAL, D + 1>(c, m, scratch, sm, pf, L);
  }
}

// c: fragments of the upper tiles (entries at rows/cols >= m must be 0); with SR > 0 the tiles of the first SR
// tile-rows are read from / written to sm (this lane's slot, see tile_get) instead.  scratch: Tiles<MT>::kScratch
// doubles private to the warp.  Returns the signed Pfaffian in every lane.
template <int MT, int SR = 0, bool BOUNDED = false, bool PHYSICAL = BOUNDED>
__device__ double warp_pfaffian_tiles(double (&c)[Tiles<MT>::NT][2], int m, double* scratch, double2* sm = nullptr) {
  if (m == 0) return 1.0;
  if (m & 1) return 0.0;
  double pf = 1.0;
  const StaticLane<MT> L = static_lane_setup<MT>(scratch);
  const int done = warp_pfaffian_static_steps<MT, SR, BOUNDED, PHYSICAL>(c, m, scratch, sm, pf, L);
  if (done < m) pf *= warp_pfaffian_tiles_pivoted<MT, SR>(c, m, scratch, sm, done);
  return pf;
}

}  // namespace mcb200
