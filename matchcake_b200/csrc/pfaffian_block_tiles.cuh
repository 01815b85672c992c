// Device routines of the block-per-matrix tile Pfaffian (64 < m <= 256), shared by pfaffian_block_tiles.cu (tiles in
// shared memory / in an L2 slab) and pfaffian_panel.cu (panel-blocked variant for 128 < m <= 256).
#pragma once
#include "common.cuh"
#include "pfaffian_src.cuh"
#include "pfaffian_tiles.cuh"

namespace mcb200 {

// MAXM (template parameter below) = 128 or 256 = number of threads of the block = capacity of the shared vectors;
// vector stride MAXM + 4 (4 mod 16 doubles, see pfaffian_tiles.cuh)

__device__ __forceinline__ int bt_idx(int MT, int ti, int tj) { return ti * MT - (ti * (ti - 1)) / 2 + (tj - ti); }

// A[i][j] for i < j, straight out of the fragment layout
__device__ __forceinline__ double bt_upper(const double2* tiles, int MT, int i, int j) {
  const double2 v = tiles[bt_idx(MT, i >> 3, j >> 3) * 32 + (i & 7) * 4 + ((j & 7) >> 1)];
  return (j & 1) ? v.y : v.x;
}
__device__ __forceinline__ double bt_elem(const double2* tiles, int MT, int i, int j) {
  return i < j ? bt_upper(tiles, MT, i, j) : (i > j ? -bt_upper(tiles, MT, j, i) : 0.0);
}

// largest |vec[c]| (high words) over c in (last, m), identical in every warp; every lane of every warp takes part
template <int MAXM>
__device__ __forceinline__ unsigned bt_max_key(const double* vec, int last, int m) {
  const int lane = lane_id();
  unsigned key = 0u;
#pragma unroll
  for (int q = 0; q < MAXM / 32; ++q) {
    const int c = lane + 32 * q;
    if (c > last && c < m) {
      const unsigned k2 = (unsigned)__double2hiint(vec[c]) & 0x7fffffffu;
      key = k2 > key ? k2 : key;
    }
  }
  return __reduce_max_sync(0xffffffffu, key);
}

// rank-(2 or 4) update of every tile (ti, tj), ti >= ti0: the fragments are sa * srcA[8 ti + gr], sb * srcB[8 tj + gr];
// entries with index <= dead inside tile row / column tr_dead are masked (they neither change nor contribute).
__device__ __forceinline__ void bt_update(double2* tiles, int MT, int ti0, const double* srcA, double sa,
                                          const double* srcB, double sb, int dead) {
  const int lane = lane_id(), w = warp_id(), gr = lane >> 2;
  const int td = dead >> 3;
  for (int ti = ti0 + w; ti < MT; ti += (int)(blockDim.x >> 5)) {
    double af = sa * srcA[8 * ti];
    if (ti == td && gr <= (dead & 7)) af = 0.0;
    double2* row = tiles + bt_idx(MT, ti, ti) * 32 + lane;
    for (int tj = ti; tj < MT; ++tj) {
      double bf = sb * srcB[8 * tj];
      if (tj == td && gr <= (dead & 7)) bf = 0.0;
      const double2 v = row[(tj - ti) * 32];
      double t[2] = {v.x, v.y};
      dmma_acc(t, af, bf);
      row[(tj - ti) * 32] = make_double2(t[0], t[1]);
    }
  }
}

// Pivoted continuation (Parlett-Reid with logical pivoting) on the indices >= done of the matrix held in `tiles`
// (packed upper tiles of an MT-tile-row matrix); pf = product of the pivots taken so far.  Returns the signed Pfaffian,
// valid in every thread; ends with a block barrier.  Needs vec = 4 * (MAXM + 4) doubles, m <= MAXM <= blockDim.x.
template <int MAXM>
__device__ double bt_pivoted_tail(double2* tiles, int MT, int m, int done, double* vec, double pf) {
  constexpr int kBtMaxM = MAXM, kBtVS = MAXM + 4, kWords = MAXM / 32;
  const int tid = threadIdx.x, lane = tid & 31, gr = lane >> 2, tg = lane & 3;
  double* u = vec;
  double* v = vec + 2 * kBtVS;
  double* u2 = vec + kBtVS;
  double* v2 = vec + 3 * kBtVS;
  // ---- pivoted continuation on the indices >= done
  if (done < m) {
    __syncthreads();
    unsigned act[kWords];
#pragma unroll
    for (int q = 0; q < kWords; ++q) {
      const int lo = 32 * q, hi = lo + 32;
      unsigned bits = 0u;
      if (m > lo) bits = (m >= hi) ? 0xffffffffu : ((1u << (m - lo)) - 1u);
      if (done > lo) bits &= (done >= hi) ? 0u : ~((1u << (done - lo)) - 1u);
      act[q] = bits;
    }
    auto is_act = [&](int c) { return (act[c >> 5] >> (c & 31)) & 1u; };
    unsigned parity = 0;
    double* rowbuf = u;   // A[r][c]
    double* colbuf = v;   // A[c][p]
    double* tauv = u2;    // tau
    double* vv = v2;      // v
    const int steps = (m - done) >> 1;
    for (int s = 0; s < steps; ++s) {
      int r = 0;
#pragma unroll
      for (int q = kWords - 1; q >= 0; --q)
        if (act[q]) r = 32 * q + __ffs((int)act[q]) - 1;
      act[r >> 5] &= ~(1u << (r & 31));
      if (tid < kBtMaxM) rowbuf[tid] = (tid < m && is_act(tid)) ? bt_elem(tiles, MT, r, tid) : 0.0;
      __syncthreads();
      // pivot: largest |A[r][c]| over the active c (first index among equals)
      unsigned key = 0u;
      int bj = 0;
#pragma unroll
      for (int q = 0; q < kBtMaxM / 32; ++q) {
        const int c = lane + 32 * q;
        const unsigned k2 = (unsigned)__double2hiint(rowbuf[c]) & 0x7fffffffu;
        if (k2 > key) {
          key = k2;
          bj = c;
        }
      }
      const unsigned kmax = __reduce_max_sync(0xffffffffu, key);
      if (kmax == 0u) {  // singular
        pf = 0.0;
        break;
      }
      const unsigned cand = (key == kmax) ? (unsigned)bj : 0xffffffffu;
      const int p = (int)__reduce_min_sync(0xffffffffu, cand);
      const double piv = rowbuf[p];
      // sign: active indices strictly between r and p (r is already cleared)
      {
        int cnt = 0;
#pragma unroll
        for (int q = 0; q < kWords; ++q) {
          unsigned mask = act[q];
          const int lo = 32 * q;
          if (p <= lo) mask = 0u;
          else if (p < lo + 32) mask &= (1u << (p - lo)) - 1u;
          if (r >= lo + 31) mask = 0u;
          else if (r >= lo) mask &= ~((2u << (r - lo)) - 1u);
          cnt += __popc(mask);
        }
        parity ^= (unsigned)cnt;
      }
      act[p >> 5] &= ~(1u << (p & 31));
      pf *= piv;
      if (s == steps - 1) break;
      const double inv = fast_rcp(piv);
      if (tid < kBtMaxM) {
        const bool on = tid < m && is_act(tid);
        colbuf[tid] = on ? bt_elem(tiles, MT, tid, p) : 0.0;
      }
      __syncthreads();
      if (tid < kBtMaxM) {
        const double t = rowbuf[tid] * inv;  // 0 on inactive entries (p itself was cleared above -> mask it here)
        const bool on = tid < m && is_act(tid);
        tauv[tid] = on ? t : 0.0;
        vv[tid] = on ? colbuf[tid] : 0.0;
      }
      __syncthreads();
      // A += tau v^T - v tau^T: A-fragment columns (tau/2, -v/2, tau/2, -v/2), B-fragment rows (v; tau; v; tau)
      {
        const bool odd = tg & 1;
        bt_update(tiles, MT, 0, (odd ? vv : tauv) + gr, odd ? -0.5 : 0.5, (odd ? tauv : vv) + gr, 1.0, -1);
      }
      __syncthreads();
    }
    if (parity & 1u) pf = -pf;
  }
  __syncthreads();
  return pf;
}

// Signed Pfaffian of matrix `item` of `src` (m <= MAXM), computed by the whole block of MAXM threads; the result is
// valid in every thread.  tiles: MT(MT+1)/2 * 32 double2 (shared or global), vec: 4 * (MAXM + 4) doubles of shared
// memory.  Ends with a block barrier.
// LOAD = false: the tiles already hold the matrix (the panel kernel leaves its Schur complement there).
template <int MAXM, class Src, bool LOAD = true>
__device__ double block_tiles_pfaffian(const Src& src, long long item, int m, double2* tiles, double* vec) {
  constexpr int kBtMaxM = MAXM, kBtVS = MAXM + 4, kWords = MAXM / 32;
  const int MT = (m + 7) >> 3;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, gr = lane >> 2, tg = lane & 3;
  // slots: row R+k -> slot brev2(k) (static path); the pivoted path uses the same four slots for its fragment vectors
  double* u = vec;
  double* v = vec + 2 * kBtVS;
  double* u2 = vec + kBtVS;
  double* v2 = vec + 3 * kBtVS;
  {
    // ---- load: tile t, lane (gr, tg) <- (A[8ti+gr][8tj+2tg], A[8ti+gr][8tj+2tg+1]); rows/cols >= m are zero
    if constexpr (LOAD) {
      for (int ti = w; ti < MT; ti += MAXM / 32)
        for (int tj = ti; tj < MT; ++tj)
          tiles[bt_idx(MT, ti, tj) * 32 + lane] = make_double2(src.at(item, 8 * ti + gr, 8 * tj + 2 * tg),
                                                               src.at(item, 8 * ti + gr, 8 * tj + 2 * tg + 1));
    }
    __syncthreads();
    double pf = 1.0;
    int done = m;  // number of indices eliminated by the static path
    // ---- static order, two pairs per pass
    for (int R = 0; R < m; R += 4) {
      if (tid < kBtMaxM) {
        const int c = tid;
        u[c] = (c > R && c < m) ? bt_upper(tiles, MT, R, c) : 0.0;
        v[c] = (c > R + 1 && c < m) ? bt_upper(tiles, MT, R + 1, c) : 0.0;
        u2[c] = (c > R + 2 && c < m) ? bt_upper(tiles, MT, R + 2, c) : 0.0;
        v2[c] = (c > R + 3 && c < m) ? bt_upper(tiles, MT, R + 3, c) : 0.0;
      }
      __syncthreads();
      const double piv1 = u[R + 1];
      const unsigned k1 = bt_max_key<MAXM>(u, R + 1, m);
      if (!(fabs(piv1) >= kStaticPivotRatio * __hiloint2double((int)k1, 0)) || piv1 == 0.0) {
        done = R;
        break;
      }
      pf *= piv1;
      if (R + 2 >= m) break;
      const double inv1 = fast_rcp(piv1);
      {
        const double t2 = u[R + 2] * inv1, t3 = u[R + 3] * inv1, w2 = v[R + 2], w3 = v[R + 3];
        if (tid < kBtMaxM && tid >= R + 2) {
          const double tc = u[tid] * inv1, wc = v[tid];
          u2[tid] += w2 * tc - t2 * wc;
          v2[tid] += w3 * tc - t3 * wc;
        }
      }
      __syncthreads();
      const double piv2 = u2[R + 3];
      const unsigned k2 = bt_max_key<MAXM>(u2, R + 3, m);
      const bool odd = tg & 1;
      if (!(fabs(piv2) >= kStaticPivotRatio * __hiloint2double((int)k2, 0)) || piv2 == 0.0) {
        // commit the first pair alone and hand over
        bt_update(tiles, MT, R >> 3, (odd ? u : v) + gr, odd ? -0.5 * inv1 : 0.5, (odd ? v : u) + gr, odd ? 1.0 : inv1,
                  R + 1);
        done = R + 2;
        break;
      }
      pf *= piv2;
      if (R + 4 >= m) break;
      const double inv2 = fast_rcp(piv2);
      const int ka = tg ^ 1;
      bt_update(tiles, MT, (R + 4) >> 3 > (R >> 3) ? (R >> 3) + 1 : (R >> 3),
                vec + (((ka & 1) << 1) | (ka >> 1)) * kBtVS + gr, odd ? ((tg & 2) ? -inv2 : -inv1) : 1.0,
                vec + (((tg & 1) << 1) | (tg >> 1)) * kBtVS + gr, odd ? 1.0 : ((tg & 2) ? inv2 : inv1), R + 3);
      __syncthreads();
    }
    if (done < m) return bt_pivoted_tail<MAXM>(tiles, MT, m, done, vec, pf);
    __syncthreads();
    return pf;
  }
}

// Gram cells for 64 < d <= 256: one block per evaluated cell (i, j), matrix cov0[i] + cov1[j] (gram_matrix.py rules as in
// the other Gram kernels: only i > j, or j > i when n0 < n1, unless FULL; REFERENCE mirrors the value).
struct GramPairSrc {
  const double* a;
  const double* b;
  int m;
  __device__ __forceinline__ double at(long long, int i, int j) const {
    if (i >= m || j >= m || i == j) return 0.0;
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    const double v = a[lo * m + hi] + b[lo * m + hi];
    return i < j ? v : -v;
  }
};

struct GramBtArgs {
  const double* cov0;
  const double* cov1;
  long long n0, n1;
  int d;
  long long row_begin, rows;
  int mode;
  double scale;
  double* K;
  long long ldk;
  double floor2;
};

}  // namespace mcb200
