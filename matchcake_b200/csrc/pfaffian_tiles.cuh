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
  // scratch doubles per warp: rowbuf[M] | pbuf[M] | tau[M] | v[M] | negv[M] | zero[M]
  static constexpr int kScratch = 6 * M;
  __host__ __device__ static constexpr int idx(int ti, int tj) { return ti * MT - (ti * (ti - 1)) / 2 + (tj - ti); }
};

__device__ __forceinline__ void dmma_acc(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

template <int MT, int TI>
__device__ __forceinline__ void tiles_store_row(const double (&c)[Tiles<MT>::NT][2], double* row_st) {
#pragma unroll
  for (int tj = TI; tj < MT; ++tj)
    *reinterpret_cast<double2*>(row_st + 8 * tj) =
        make_double2(c[Tiles<MT>::idx(TI, tj)][0], c[Tiles<MT>::idx(TI, tj)][1]);
}

template <int MT, int TI>
__device__ __forceinline__ void tiles_store_row_right(const double (&c)[Tiles<MT>::NT][2], double* row_st) {
#pragma unroll
  for (int tj = TI + 1; tj < MT; ++tj)
    *reinterpret_cast<double2*>(row_st + 8 * tj) =
        make_double2(c[Tiles<MT>::idx(TI, tj)][0], c[Tiles<MT>::idx(TI, tj)][1]);
}

template <int MT, int TJ>
__device__ __forceinline__ void tiles_store_col(const double (&c)[Tiles<MT>::NT][2], double* col_st, int e) {
#pragma unroll
  for (int ti = 0; ti <= TJ; ++ti) col_st[8 * ti] = e ? c[Tiles<MT>::idx(ti, TJ)][1] : c[Tiles<MT>::idx(ti, TJ)][0];
}

template <int MT, int TI>
__device__ __forceinline__ void tiles_update_from(double (&c)[Tiles<MT>::NT][2], const double (&af)[MT],
                                                  const double (&bf)[MT]) {
#pragma unroll
  for (int ti = TI; ti < MT; ++ti)
#pragma unroll
    for (int tj = ti; tj < MT; ++tj) dmma_acc(c[Tiles<MT>::idx(ti, tj)], af[ti], bf[tj]);
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

// c: fragments of the upper tiles (entries at rows/cols >= m must be 0).  scratch: Tiles<MT>::kScratch doubles
// private to the warp.  Returns the signed Pfaffian in every lane.
template <int MT>
__device__ double warp_pfaffian_tiles(double (&c)[Tiles<MT>::NT][2], int m, double* scratch) {
  using T = Tiles<MT>;
  constexpr int H = (T::M + 31) / 32;  // vector entries / pivot candidates per lane
  const int lane = lane_id(), gr = lane >> 2, tg = lane & 3;
  if (m == 0) return 1.0;
  if (m & 1) return 0.0;
  double* rowbuf = scratch;
  double* pbuf = scratch + T::M;
  double* tauv = scratch + 2 * T::M;
  double* vv = scratch + 3 * T::M;
  double* negv = scratch + 4 * T::M;
  double* zero = scratch + 5 * T::M;
#pragma unroll
  for (int h = 0; h < H; ++h)
    if (lane + 32 * h < T::M) zero[lane + 32 * h] = 0.0;
  // per-lane store / load addresses (loop invariant)
  double* row_st = rowbuf + 2 * tg;   // row r -> rowbuf[8 tj + 2 tg .. +1]
  double* prow_st = pbuf + 2 * tg;    // row p -> pbuf[8 tj + 2 tg .. +1]
  double* pcol_st = pbuf + gr;        // column p -> pbuf[8 ti + gr]
  // fragment sources, fixed per lane: A-fragment column tg = (tau, -v, 0, 0), B-fragment row tg = (v; tau; 0; 0)
  const double* srcA = ((tg == 0) ? tauv : (tg == 1) ? negv : zero) + gr;
  const double* srcB = ((tg == 0) ? vv : (tg == 1) ? tauv : zero) + gr;
  // active set as two 32-bit words (uniform) + this lane's own entries j = lane, lane + 32
  unsigned act_lo = (m >= 32) ? 0xffffffffu : ((1u << m) - 1u);
  unsigned act_hi = (m >= 64) ? 0xffffffffu : ((m > 32) ? ((1u << (m - 32)) - 1u) : 0u);
  bool on0 = lane < m, on1 = lane + 32 < m;
  const int j1 = (lane + 32 < T::M) ? lane + 32 : lane;  // second entry of this lane (only used when on1)
  double pf = 1.0;
  unsigned parity = 0;
  const int steps = m >> 1;
  for (int s = 0; s < steps; ++s) {
    const int r = act_lo ? (__ffs((int)act_lo) - 1) : (31 + __ffs((int)act_hi));
    const int tr = r >> 3;
    // (1) row r, columns in tile-columns >= tr
    if (gr == (r & 7)) {
      MCB200_TILE_SWITCH(tr, MT, (tiles_store_row<MT, K_>(c, row_st)))
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
      MCB200_TILE_SWITCH(tp, MT, (tiles_store_col<MT, K_>(c, pcol_st, p & 1)))
    }
    if (gr == (p & 7)) {
      MCB200_TILE_SWITCH(tp, MT, (tiles_store_row_right<MT, K_>(c, prow_st)))
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
        tauv[lane] = t0;
        vv[lane] = v0;
        negv[lane] = -v0;
      }
      if (H > 1) {
        const int j = j1;
        const double t1 = on1 ? rowbuf[j] * inv : 0.0;
        double v1 = on1 ? pbuf[j] : 0.0;
        if ((j >> 3) > tp) v1 = -v1;
        if (lane + 32 < T::M) {
          tauv[j] = t1;
          vv[j] = v1;
          negv[j] = -v1;
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
    MCB200_TILE_SWITCH(tr, MT, (tiles_update_from<MT, K_>(c, af, bf)))
    __syncwarp();
  }
  return (parity & 1u) ? -pf : pf;
}

}  // namespace mcb200
