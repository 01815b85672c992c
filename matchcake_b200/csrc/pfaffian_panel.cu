// K3 for 128 < m <= 256, panel-blocked: one block per matrix, everything in shared memory.
//
// The slab variant (pfaffian_block_tiles.cu, GLOBAL = true) keeps the 528 upper tiles of a 256 x 256 matrix in an
// L2-resident slab and sweeps ALL live tiles once per eliminated quadruple: 64 read-modify-write passes over up to
// 270 KB, 14 MB of DRAM traffic per matrix with 4 resident blocks per SM (ncu: 3.7 TB/s of HBM, FP64 tensor pipe 8 %
// busy, profiles/r02/ncu_pf256_block.txt).  Here the elimination is split at index H = 128:
//
//   A = [ A11  A12 ]     Pf(A) = Pf(A11) Pf(S),    S = A22 + A12^T A11^-1 A12   (Schur complement, skew)
//       [-A12^T A22 ]
//
//   1. PANEL: the tile rows of the first 128 indices -- the 136 upper tiles of A11 and the 16 x CT tiles of A12 (<= 200 KB)
//      -- are loaded into shared memory and eliminated in the static order, two pairs per pass, with the rank-4 DMMA
//      update of pfaffian_block_tiles.cuh restricted to those tile rows.  A22 is not touched.
//   2. SCHUR: the rows of A12 left behind by the elimination are exactly the vectors of the deferred rank-2 updates:
//      with a_s, b_s = rows 2s, 2s+1 at the time pair s was eliminated and p_s its pivot,
//          S = A22 + sum_s (b_s^T a_s - a_s^T b_s) / p_s = A22 + U^T (J U),
//      ONE K = 128 product per tile of S with the accumulators in registers (A22 is read from global memory exactly once,
//      straight into the accumulators): 32 DMMAs per tile instead of 32 read-modify-write passes.
//   3. S (<= 128 x 128) overwrites the dead A11 tiles and the shared-memory routine for m <= 128 finishes it (static
//      order, pivoted continuation on S if one of its guards fails).
// If a pivot guard fails inside the panel (never on this path's covariance sums; random matrices ~never) the block
// redoes the matrix with the pivoted slab routine -- the caller's workspace is kept for that.
// Reference semantics: utils/_pfaffian.py:17-119 (torch_pfaffian), as in pfaffian_tiles.cuh.
#include "common.cuh"
#include "pfaffian_block_tiles.cuh"
#include "tiles_api.cuh"

namespace mcb200 {

constexpr int kPanelH = 128;               // indices eliminated by the panel phase
constexpr int kPanelHT = kPanelH / 8;      // its tile rows
constexpr int kPanelA11 = kPanelHT * (kPanelHT + 1) / 2;  // 136 upper tiles of A11 (= room for S afterwards)
constexpr int kPanelThreads = 512;
constexpr int kPanelVS = 256 + 4;          // stride of the four row vectors

// tile (ti, tj), ti < 16, tj >= ti:  A11 packed like a 128 x 128 matrix, A12 row-major behind it
__device__ __forceinline__ int panel_tile(int CT, int ti, int tj) {
  return tj < kPanelHT ? bt_idx(kPanelHT, ti, tj) : kPanelA11 + ti * CT + (tj - kPanelHT);
}
// A[i][j], i < j, i < 128
__device__ __forceinline__ double panel_upper(const double2* tiles, int CT, int i, int j) {
  const double2 v = tiles[panel_tile(CT, i >> 3, j >> 3) * 32 + (i & 7) * 4 + ((j & 7) >> 1)];
  return (j & 1) ? v.y : v.x;
}
// address of U[k][c] = A[k][128 + c] (k < 128) as a double
__device__ __forceinline__ const double* panel_u(const double2* tiles, int CT, int k, int c) {
  return reinterpret_cast<const double*>(tiles + (kPanelA11 + (k >> 3) * CT + (c >> 3)) * 32 + (k & 7) * 4 + ((c & 7) >> 1)) +
         (c & 1);
}

// Tile layouts of the two static-order passes.  TWO = true: the panel (rows < 16; A11 packed, A12 row-major behind it,
// `cols` = MT tile columns).  TWO = false: S, packed like a `cols`-tile-row matrix.  In both, the tiles tj0, tj0+1, ...
// of one row are 32 double2 apart as long as they stay on one side of tile column 16 (TWO) / always (packed).
template <bool TWO>
struct PanelLayout {
  int CT, cols;
  __device__ __forceinline__ int rows() const { return TWO ? kPanelHT : cols; }
  __device__ __forceinline__ int idx(int ti, int tj) const { return TWO ? panel_tile(CT, ti, tj) : bt_idx(cols, ti, tj); }
  __device__ __forceinline__ double upper(const double2* tiles, int i, int j) const {
    const double2 v = tiles[idx(i >> 3, j >> 3) * 32 + (i & 7) * 4 + ((j & 7) >> 1)];
    return (j & 1) ? v.y : v.x;
  }
};

// Rank-4 update of the tile rows ti0 .. rows-1.  Work unit = 4 consecutive tiles of one row: the four 16-byte tile loads
// and the four B-fragment loads are issued together, then four DMMAs, then four stores -- no per-tile multiplies, masks
// or index arithmetic (the vectors are pre-masked, the A-side scale is applied once per tile row).
template <bool TWO>
__device__ __forceinline__ void panel_update(double2* tiles, const PanelLayout<TWO>& L, int ti0, const double* avec,
                                             double sa, const double* bvec) {
  const int lane = lane_id(), w = warp_id(), nw = (int)(blockDim.x >> 5);
  int t = w;  // rolling chunk counter: the warps share the whole trapezoid round-robin
  for (int ti = ti0; ti < L.rows(); ++ti) {
    const double af = sa * avec[8 * ti];
#pragma unroll
    for (int run = 0; run < (TWO ? 2 : 1); ++run) {
      const int lo = run == 0 ? ti : kPanelHT;
      const int hi = (TWO && run == 0) ? kPanelHT : L.cols;
      const int nch = (hi - lo + 3) >> 2;
      for (; t < nch; t += nw) {
        const int tj0 = lo + 4 * t;
        double2* p = tiles + L.idx(ti, tj0) * 32 + lane;
        const double* bp = bvec + 8 * tj0;
        if (tj0 + 4 <= hi) {
          double2 v0 = p[0], v1 = p[32], v2 = p[64], v3 = p[96];
          const double b0 = bp[0], b1 = bp[8], b2 = bp[16], b3 = bp[24];
          double a0[2] = {v0.x, v0.y}, a1[2] = {v1.x, v1.y}, a2[2] = {v2.x, v2.y}, a3[2] = {v3.x, v3.y};
          dmma_acc(a0, af, b0);
          dmma_acc(a1, af, b1);
          dmma_acc(a2, af, b2);
          dmma_acc(a3, af, b3);
          p[0] = make_double2(a0[0], a0[1]);
          p[32] = make_double2(a1[0], a1[1]);
          p[64] = make_double2(a2[0], a2[1]);
          p[96] = make_double2(a3[0], a3[1]);
        } else {
          for (int q = 0; tj0 + q < hi; ++q) {
            const double2 v0 = p[32 * q];
            double a0[2] = {v0.x, v0.y};
            dmma_acc(a0, af, bp[8 * q]);
            p[32 * q] = make_double2(a0[0], a0[1]);
          }
        }
      }
      t -= nch;
    }
  }
}

// A[x][y] (x != y) of the matrix in `tiles`; min(x, y) must be a row the layout holds.
template <bool TWO>
__device__ __forceinline__ double panel_elem(const double2* tiles, const PanelLayout<TWO>& L, int x, int y) {
  return x < y ? L.upper(tiles, x, y) : -L.upper(tiles, y, x);
}
template <bool TWO>
__device__ __forceinline__ void panel_set(double2* tiles, const PanelLayout<TWO>& L, int x, int y, double val) {
  const int i = x < y ? x : y, j = x < y ? y : x;
  double* p = reinterpret_cast<double*>(tiles + L.idx(i >> 3, j >> 3) * 32 + (i & 7) * 4 + ((j & 7) >> 1)) + (j & 1);
  *p = x < y ? val : -val;
}

// Symmetric exchange of the indices a < b (rows and columns) among the live indices > R: A <- P A P^T, Pf <- -Pf.
// One thread per index; every storage cell is read and written by exactly one thread.  No barrier inside.
template <bool TWO>
__device__ __forceinline__ void panel_swap(double2* tiles, const PanelLayout<TWO>& L, int m, int R, int a, int b) {
  const int i = threadIdx.x;
  if (i <= R || i >= m) return;
  if (i == a) {
    panel_set<TWO>(tiles, L, a, b, -panel_elem<TWO>(tiles, L, a, b));
  } else if (i != b) {
    const double ea = panel_elem<TWO>(tiles, L, i, a), eb = panel_elem<TWO>(tiles, L, i, b);
    panel_set<TWO>(tiles, L, i, a, eb);
    panel_set<TWO>(tiles, L, i, b, ea);
  }
}

// (largest |vec[c]| over R1 < c < hi as the high word, its lowest index): identical in every warp
__device__ __forceinline__ void panel_argmax(const double* vec, int R1, int hi, unsigned& kmax, int& p) {
  const int lane = lane_id();
  unsigned key = 0u;
  int bj = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int c = lane + 32 * q;
    if (c > R1 && c < hi) {
      const unsigned k2 = (unsigned)__double2hiint(vec[c]) & 0x7fffffffu;
      if (k2 > key) {
        key = k2;
        bj = c;
      }
    }
  }
  kmax = __reduce_max_sync(0xffffffffu, key);
  p = (int)__reduce_min_sync(0xffffffffu, (key == kmax && kmax != 0u) ? (unsigned)bj : 0xffffffffu);
}

// Elimination of the indices [0, stop) of the matrix in `tiles` (stop = 128 for the panel, m for S) in the natural order
// with THRESHOLD PIVOTING inside [0, stop): pair (R, R+1) takes the natural pivot A[R][R+1] while it is at least
// kStaticPivotRatio times the largest entry of row R; otherwise index R+1 is exchanged (rows and columns, O(m) work) with
// the best candidate below `stop` that passes the same test.  Two pairs are committed per pass as ONE rank-4 DMMA update
// when the second natural pivot passes as well, else the first pair alone (the second one then leads the next pass and
// may pivot).  vec: 8 vectors of kPanelVS doubles -- rows R .. R+3 in slots brev2(k) (raw), then the same rows masked to
// the live columns (the update's fragments).  Returns the number of indices eliminated (< stop only if a row has no
// acceptable pivot below `stop`); pf is multiplied by the signed pivots.
// TWO: records 1 / pivot per pair in pinv and keeps the A12 part of every eliminated row as it was when its pair was
// eliminated (rows R+2, R+3 exist only as vectors once the first pair has been applied to them: they are written back).
template <bool TWO>
__device__ int panel_static(double2* tiles, const PanelLayout<TWO>& L, int m, int stop, double* vec, double* pinv,
                            double& pf) {
  const int tid = threadIdx.x, lane = tid & 31, gr = lane >> 2, tg = lane & 3;
  double* u = vec;                     // row R      (slot 0)
  double* v = vec + 2 * kPanelVS;      // row R + 1  (slot 2)
  double* u2 = vec + kPanelVS;         // row R + 2  (slot 1)
  double* v2 = vec + 3 * kPanelVS;     // row R + 3  (slot 3)
  double* zv = vec + 4 * kPanelVS;     // masked copies, same slot order
  // fragments of this lane: A = row R + (tg ^ 1) scaled by (+-) 1 / pivot, B = row R + tg
  const int ka = tg ^ 1;
  const double* avec = zv + (((ka & 1) << 1) | (ka >> 1)) * kPanelVS + gr;
  const double* bvec = zv + (((tg & 1) << 1) | (tg >> 1)) * kPanelVS + gr;
  int R = 0;
  while (R < stop) {
    const bool has2 = R + 3 < stop;  // a second pair exists inside [0, stop)
    {
      const int c = tid & 255;
      if (tid < 256) {
        u[c] = (c > R && c < m) ? L.upper(tiles, R, c) : 0.0;
        v[c] = (c > R + 1 && c < m) ? L.upper(tiles, R + 1, c) : 0.0;
      } else if (tid < 512 && has2) {
        u2[c] = (c > R + 2 && c < m) ? L.upper(tiles, R + 2, c) : 0.0;
        v2[c] = (c > R + 3 && c < m) ? L.upper(tiles, R + 3, c) : 0.0;
      }
    }
    __syncthreads();
    double piv1 = u[R + 1];
    {
      const unsigned k1 = bt_max_key<256>(u, R + 1, m);
      const double thr = kStaticPivotRatio * __hiloint2double((int)k1, 0);
      if (!(fabs(piv1) >= thr) || piv1 == 0.0) {
        // ---- rare: exchange index R+1 with the best candidate below `stop`
        unsigned kin;
        int p;
        panel_argmax(u, R + 1, stop, kin, p);
        const double cand = kin != 0u ? u[p] : 0.0;
        if (!(fabs(cand) >= thr) || cand == 0.0) return R;  // no acceptable pivot below `stop`
        __syncthreads();  // everyone has read u / the tiles of this pass
        panel_swap<TWO>(tiles, L, m, R, R + 1, p);
        if (tid == 0) {
          const double t = u[R + 1];
          u[R + 1] = u[p];
          u[p] = t;
        }
        __syncthreads();
        {
          const int c = tid & 255;
          if (tid < 256) {
            v[c] = (c > R + 1 && c < m) ? L.upper(tiles, R + 1, c) : 0.0;
          } else if (tid < 512 && has2) {
            u2[c] = (c > R + 2 && c < m) ? L.upper(tiles, R + 2, c) : 0.0;
            v2[c] = (c > R + 3 && c < m) ? L.upper(tiles, R + 3, c) : 0.0;
          }
        }
        __syncthreads();
        piv1 = u[R + 1];
        pf = -pf;
      }
    }
    if (R + 2 >= m) {  // last pair of the matrix: nothing left to update
      pf *= piv1;
      return m;
    }
    const double inv1 = fast_rcp(piv1);
    if (has2) {
      const double t2 = u[R + 2] * inv1, t3 = u[R + 3] * inv1, w2 = v[R + 2], w3 = v[R + 3];
      if (tid < 256 && tid >= R + 2) {
        const int c = tid;
        const double tc = u[c] * inv1, wc = v[c];
        u2[c] += w2 * tc - t2 * wc;
        v2[c] += w3 * tc - t3 * wc;
      }
    }
    __syncthreads();
    bool quad = false;
    double piv2 = 0.0, inv2 = 0.0;
    if (has2) {
      piv2 = u2[R + 3];
      const unsigned k2 = bt_max_key<256>(u2, R + 3, m);
      quad = fabs(piv2) >= kStaticPivotRatio * __hiloint2double((int)k2, 0) && piv2 != 0.0;
      if (quad) inv2 = fast_rcp(piv2);
    }
    pf *= quad ? piv1 * piv2 : piv1;
    const int dead = quad ? R + 3 : R + 1;  // indices <= dead leave the matrix with this pass
    if (tid < 256) {
      const int c = tid;
      const bool live = c > dead;
      const double nu = u2[c], nv = v2[c];
      zv[c] = live ? u[c] : 0.0;
      zv[2 * kPanelVS + c] = live ? v[c] : 0.0;
      zv[kPanelVS + c] = (live && quad) ? nu : 0.0;
      zv[3 * kPanelVS + c] = (live && quad) ? nv : 0.0;
      if (TWO && quad && c >= kPanelH && c < 8 * L.cols) {
        *const_cast<double*>(panel_u(tiles, L.CT, R + 2, c - kPanelH)) = nu;
        *const_cast<double*>(panel_u(tiles, L.CT, R + 3, c - kPanelH)) = nv;
      }
      if (TWO && c == 0) {
        pinv[R >> 1] = inv1;
        if (quad) pinv[(R >> 1) + 1] = inv2;
      }
    }
    __syncthreads();
    if (dead + 1 < m)
      panel_update<TWO>(tiles, L, (dead + 1) >> 3, avec,
                        (tg & 2) ? ((tg & 1) ? -inv2 : inv2) : ((tg & 1) ? -inv1 : inv1), bvec);
    __syncthreads();
    R = dead + 1;
  }
  return stop;
}

// Phase 2: S = A22 + U^T (J U) into the packed tiles [0, CT (CT + 1) / 2) (layout of a CT-tile-row matrix).
// Work unit = up to 4 consecutive tiles of one tile row (they share the A fragment); the unit list is walked round-robin.
template <class Src>
__device__ void panel_schur(const Src& src, long long item, int m, double2* tiles, const double* pinv) {
  const int MT = (m + 7) >> 3, CT = MT - kPanelHT;
  const int lane = lane_id(), w = warp_id(), gr = lane >> 2, tg = lane & 3, nw = (int)(blockDim.x >> 5);
  const double sgn = (tg & 1) ? 1.0 : -1.0;   // V[2s] = -U[2s+1] / p_s,  V[2s+1] = +U[2s] / p_s, folded into the A fragment
  int unit = 0;
  for (int ti = 0; ti < CT; ++ti) {
    for (int tj0 = ti; tj0 < CT; tj0 += 4, ++unit) {
      if (unit % nw != w) continue;
      const int cnt = min(4, CT - tj0);
      double acc[4][2];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = kPanelH + 8 * ti + gr, c = kPanelH + 8 * (tj0 + q) + 2 * tg;
        acc[q][0] = q < cnt ? src.at(item, r, c) : 0.0;
        acc[q][1] = q < cnt ? src.at(item, r, c + 1) : 0.0;
      }
      const double* pa = panel_u(tiles, CT, tg, 8 * ti + gr);            // + kk * (4 rows) below
      const double* pb = panel_u(tiles, CT, tg ^ 1, 8 * tj0 + gr);
#pragma unroll 4
      for (int kk = 0; kk < kPanelH / 4; ++kk) {
        // rows 4kk + tg: tile row kk >> 1 (stride CT tiles), in-tile row 4 (kk & 1) + tg (stride 4 double2 = 8 doubles)
        const int off = ((kk >> 1) * CT * 32 + (kk & 1) * 16) * 2;
        const double a = pa[off] * (sgn * pinv[2 * kk + (tg >> 1)]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q < cnt) dmma_acc(acc[q], a, pb[off + q * 64]);
        }
      }
      // S tile (ti, tj0 + q): the A11 tiles are dead, but other warps may still be reading U -> S goes to its own region
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (q < cnt) tiles[bt_idx(CT, ti, tj0 + q) * 32 + lane] = make_double2(acc[q][0], acc[q][1]);
    }
  }
}


// Signed Pfaffian of matrix `item` (128 < m <= 256) by the whole block; valid in every thread.
// smem: tiles (136 + 16 CT) * 32 double2 | vec 8 * kPanelVS doubles | pinv 64 doubles.  slab: fallback storage.
template <class Src>
__device__ double panel_pfaffian(const Src& src, long long item, int m, double2* tiles, double* vec, double* pinv,
                                 double2* slab) {
  const int MT = (m + 7) >> 3, CT = MT - kPanelHT;
  const int lane = lane_id(), w = warp_id(), gr = lane >> 2, tg = lane & 3, nw = (int)(blockDim.x >> 5);
  // ---- phase 1: panel
  for (int ti = w; ti < kPanelHT; ti += nw)
    for (int tj = ti; tj < MT; ++tj)
      tiles[panel_tile(CT, ti, tj) * 32 + lane] = make_double2(src.at(item, 8 * ti + gr, 8 * tj + 2 * tg),
                                                               src.at(item, 8 * ti + gr, 8 * tj + 2 * tg + 1));
  __syncthreads();
  double pf = 1.0;
  const PanelLayout<true> L1{CT, MT};
  const int done1 = panel_static<true>(tiles, L1, m, kPanelH, vec, pinv, pf);
  if (done1 < kPanelH) {  // block-uniform: every thread evaluates the same guards on the same shared vectors
    __syncthreads();
    return block_tiles_pfaffian<256>(src, item, m, slab, vec);
  }
  // ---- phase 2: Schur complement
  panel_schur(src, item, m, tiles, pinv);
  __syncthreads();
  // ---- phase 3: S, in place
  const int m2 = m - kPanelH;
  const PanelLayout<false> L2{CT, CT};
  const int done2 = panel_static<false>(tiles, L2, m2, m2, vec, nullptr, pf);
  if (done2 < m2) {
    __syncthreads();
    return bt_pivoted_tail<256>(tiles, CT, m2, done2, vec, pf);
  }
  return pf;
}

__host__ __device__ constexpr size_t panel_smem_bytes(int m) {
  return ((size_t)(kPanelA11 + kPanelHT * (((m + 7) >> 3) - kPanelHT)) * 32) * sizeof(double2) +
         (size_t)(8 * kPanelVS + 64) * sizeof(double);
}

template <class Src>
__global__ void __launch_bounds__(kPanelThreads) pfaffian_panel_kernel(Src src, long long n_items, int m, int out_mode,
                                                                       double scale, double floor2, double* out,
                                                                       double2* slabs) {
  extern __shared__ __align__(16) unsigned char pn_smem[];
  const int MT = (m + 7) >> 3, CT = MT - kPanelHT, NT = MT * (MT + 1) / 2;
  double2* tiles = reinterpret_cast<double2*>(pn_smem);
  double* vec = reinterpret_cast<double*>(tiles + (kPanelA11 + kPanelHT * CT) * 32);
  double* pinv = vec + 8 * kPanelVS;
  double2* slab = slabs + (size_t)blockIdx.x * NT * 32;
  for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
    const double pf = panel_pfaffian(src, item, m, tiles, vec, pinv, slab);
    if (threadIdx.x == 0) out[item] = out_mode == MCB200_PF_SIGNED ? scale * pf : abs_floor(scale * pf, floor2);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kPanelThreads) gram_panel_kernel(GramBtArgs g, double2* slabs) {
  extern __shared__ __align__(16) unsigned char pn_smem[];
  const int m = g.d, MT = (m + 7) >> 3, CT = MT - kPanelHT, NT = MT * (MT + 1) / 2;
  double2* tiles = reinterpret_cast<double2*>(pn_smem);
  double* vec = reinterpret_cast<double*>(tiles + (kPanelA11 + kPanelHT * CT) * 32);
  double* pinv = vec + 8 * kPanelVS;
  double2* slab = slabs + (size_t)blockIdx.x * NT * 32;
  const bool upper = g.n0 < g.n1;
  for (long long cell = blockIdx.x; cell < g.rows * g.n1; cell += gridDim.x) {
    const long long i = g.row_begin + cell / g.n1, j = cell % g.n1;
    if (g.mode != MCB200_GRAM_FULL && !(upper ? (j > i) : (i > j))) continue;
    GramPairSrc src{g.cov0 + i * (long long)m * m, g.cov1 + j * (long long)m * m, m};
    const double pf = panel_pfaffian(src, 0, m, tiles, vec, pinv, slab);
    if (threadIdx.x == 0) {
      const double val = abs_floor(g.scale * pf, g.floor2);
      g.K[i * g.ldk + j] = val;
      if (g.mode == MCB200_GRAM_REFERENCE && j < g.n0 && i < g.n1) g.K[j * g.ldk + i] = val;  // gram_matrix.py:139-143
    }
    __syncthreads();
  }
}

// one block per SM (200 KB of shared memory); the slabs are only touched by the fallback
constexpr long long kPanelCtas = kNumSMs;

long long panel_workspace_bytes(long long n_items, int m) {
  const int MT = (m + 7) / 8, NT = MT * (MT + 1) / 2;
  const long long ctas = n_items < kPanelCtas ? n_items : kPanelCtas;
  return ctas * (long long)NT * 32 * (long long)sizeof(double2);
}

int panel_pfaffian_plain(const double* A, long long n, int m, int mode, double scale, double* out, void* workspace,
                         cudaStream_t st) {
  if (n == 0) return 0;
  if (workspace == nullptr) return set_error("pfaffian: m=%d needs a workspace (see *_workspace_bytes)", m);
  PlainSrc s{A, m};
  const size_t smem = panel_smem_bytes(m);
  auto kernel = pfaffian_panel_kernel<PlainSrc>;
  if (int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                          "cudaFuncSetAttribute(pfaffian_panel)"))
    return rc;
  const long long ctas = n < kPanelCtas ? n : kPanelCtas;
  kernel<<<(unsigned)ctas, kPanelThreads, smem, st>>>(s, n, m, mode, scale, g_abs_floor2, out, (double2*)workspace);
  return check_launch("pfaffian_panel_kernel");
}

int panel_gram(const double* cov0, const double* cov1, long long n0, long long n1, int d, long long row_begin,
               long long row_end, int mode, double scale, double* K, long long ldk, void* workspace, cudaStream_t st) {
  const long long rows = row_end - row_begin;
  if (rows <= 0 || n1 <= 0) return 0;
  if (workspace == nullptr) return set_error("gram: d=%d needs a workspace (mcb200_gram_workspace_bytes)", d);
  GramBtArgs g{cov0, cov1, n0, n1, d, row_begin, rows, mode, scale, K, ldk, g_abs_floor2};
  const size_t smem = panel_smem_bytes(d);
  if (int rc = check_cuda(cudaFuncSetAttribute(gram_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                          "cudaFuncSetAttribute(gram_panel)"))
    return rc;
  const long long cells = rows * n1;
  const long long ctas = cells < kPanelCtas ? cells : kPanelCtas;
  gram_panel_kernel<<<(unsigned)ctas, kPanelThreads, smem, st>>>(g, (double2*)workspace);
  return check_launch("gram_panel_kernel");
}

}  // namespace mcb200
