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

#ifdef MCB200_PANEL_PROFILE  // phase timing of block 0 (experiments only): cycles per phase, printed per matrix
__device__ long long g_pp[8];
#define PP_T0() long long pp_t = clock64()
#define PP_RESET() pp_t = clock64()
#define PP_ADD(k)                                  \
  do {                                             \
    const long long pp_n = clock64();              \
    if (blockIdx.x == 0 && threadIdx.x == 0) g_pp[k] += pp_n - pp_t; \
    pp_t = pp_n;                                   \
  } while (0)
#else
#define PP_T0()
#define PP_RESET()
#define PP_ADD(k)
#endif

constexpr int kPanelH = 128;               // indices eliminated by the panel phase
constexpr int kPanelHT = kPanelH / 8;      // its tile rows
constexpr int kPanelA11 = kPanelHT * (kPanelHT + 1) / 2;  // 136 upper tiles of A11 (= room for S afterwards)
constexpr int kPanelThreads = 512;
constexpr int kPanelVS = 256 + 4;          // stride of the four row vectors
constexpr int kPanelMaxUnits = 40;         // 4-tile units of the Schur product at CT = 16

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

// Row i of the upper triangle as a function of the column: the tile index is linear in the tile column on either side of
// tile column 16, so a row costs two constants and an element ~8 instructions.
template <bool TWO>
struct PanelRow {
  const double* base;  // tiles as doubles, advanced to the in-tile row
  int ca, cb;          // tile index = tj + (tj < 16 ? ca : cb)
  __device__ __forceinline__ PanelRow(const double2* tiles, const PanelLayout<TWO>& L, int i) {
    const int ti = i >> 3;
    base = reinterpret_cast<const double*>(tiles) + (i & 7) * 8;
    if (TWO) {
      ca = bt_idx(kPanelHT, ti, ti) - ti;
      cb = kPanelA11 + ti * L.CT - kPanelHT;
    } else {
      ca = cb = bt_idx(L.cols, ti, ti) - ti;
    }
  }
  __device__ __forceinline__ double at(int j) const {  // A[i][j], j > i
    const int tj = j >> 3;
    return base[(tj + ((TWO && tj >= kPanelHT) ? cb : ca)) * 64 + (j & 7)];
  }
};

// Chunk tables of a triangle of tile rows (rows ti < n, tiles tj in [ti, n), units of 4 tiles): pre[ti] = chunks of the
// rows before ti, row[g] = tile row of flat chunk g.  Built once per kernel for n = 16 (panel, A11 part) and n = CT (S).
struct PanelTri {
  unsigned char pre[kPanelHT + 1];
  unsigned char row[kPanelMaxUnits];
};
__device__ __forceinline__ void panel_build_tri(PanelTri* t, int n) {
  int g = 0;
  for (int ti = 0; ti < n; ++ti) {
    t->pre[ti] = (unsigned char)g;
    for (int tj0 = ti; tj0 < n; tj0 += 4) t->row[g++] = (unsigned char)ti;
  }
  for (int ti = n; ti <= kPanelHT; ++ti) t->pre[ti] = (unsigned char)g;
}

// Rank-4 update of the tile rows ti0 .. rows-1, in units of 4 consecutive tiles of one row dealt round-robin to the warps
// (the rectangular A12 part first, then the triangle through the chunk table): the four 16-byte tile loads and the four
// B-fragment loads are issued together, then four DMMAs, then four stores -- no per-tile multiplies, masks or index
// arithmetic (the vectors are pre-masked, the A-side scale is applied once per unit).
template <bool TWO>
__device__ __forceinline__ void panel_update(double2* tiles, const PanelLayout<TWO>& L, const PanelTri& tri, int ti0,
                                             const double* avec, double sa, const double* bvec) {
  const int lane = lane_id(), w = warp_id(), nw = (int)(blockDim.x >> 5);
  const int rows = L.rows();
  if (ti0 >= rows) return;
  const int tcols = TWO ? kPanelHT : L.cols;                         // tile columns of the triangle part
  const int nch1 = TWO ? (L.cols - kPanelHT + 3) >> 2 : 0;           // chunks per row of the rectangular part
  const int nrect = (rows - ti0) * nch1, pre0 = tri.pre[ti0];
  const int total = nrect + (tri.pre[rows] - pre0);
  for (int f = w; f < total; f += nw) {
    int ti, tj0, hi;
    if (f < nrect) {
      const int r = f / nch1;
      ti = ti0 + r;
      tj0 = kPanelHT + 4 * (f - r * nch1);
      hi = L.cols;
    } else {
      const int g = f - nrect + pre0;
      ti = tri.row[g];
      tj0 = ti + 4 * (g - tri.pre[ti]);
      hi = tcols;
    }
    const double af = sa * avec[8 * ti];
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
// NC = column capacity (256: blocks of 512 threads; 128: blocks of 256 threads): threads [0, NC) own one column each,
// threads [NC, 2 NC) extract the second pair's rows.
template <bool TWO, int NC = 256>
__device__ int panel_static(double2* tiles, const PanelLayout<TWO>& L, const PanelTri& tri, int m, int stop, double* vec,
                            double* pinv, unsigned* kmx, double& pf) {
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
  PP_T0();
  while (R < stop) {
    const bool has2 = R + 3 < stop;  // a second pair exists inside [0, stop)
    {
      // rows R .. R+3 -> vectors; kmx[k] = largest |entry| of row R+k right of the pair (high word), for the guards
      const int c = tid & (NC - 1);
      if (tid < NC) {
        const PanelRow<TWO> r0(tiles, L, R), r1(tiles, L, R + 1);
        const double x = (c > R && c < m) ? r0.at(c) : 0.0, y = (c > R + 1 && c < m) ? r1.at(c) : 0.0;
        u[c] = x;
        v[c] = y;
        const unsigned kx = __reduce_max_sync(0xffffffffu, c > R + 1 ? ((unsigned)__double2hiint(x) & 0x7fffffffu) : 0u);
        const unsigned ky = __reduce_max_sync(0xffffffffu, c > R + 3 ? ((unsigned)__double2hiint(y) & 0x7fffffffu) : 0u);
        if (lane == 0) {
          if (kx) atomicMax(&kmx[0], kx);
          if (ky) atomicMax(&kmx[1], ky);
        }
      } else if (tid < 2 * NC && has2) {
        const PanelRow<TWO> r2(tiles, L, R + 2), r3(tiles, L, R + 3);
        const double x = (c > R + 2 && c < m) ? r2.at(c) : 0.0;
        u2[c] = x;
        v2[c] = (c > R + 3 && c < m) ? r3.at(c) : 0.0;
        const unsigned kx = __reduce_max_sync(0xffffffffu, c > R + 3 ? ((unsigned)__double2hiint(x) & 0x7fffffffu) : 0u);
        if (lane == 0 && kx) atomicMax(&kmx[2], kx);
      }
    }
    __syncthreads();
    PP_ADD(0);
    double piv1 = u[R + 1];
    {
      const double thr = kStaticPivotRatio * __hiloint2double((int)kmx[0], 0);
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
          kmx[1] = kmx[2] = 0u;
        }
        __syncthreads();
        {
          const int c = tid & (NC - 1);
          if (tid < NC) {
            const PanelRow<TWO> r1(tiles, L, R + 1);
            const double y = (c > R + 1 && c < m) ? r1.at(c) : 0.0;
            v[c] = y;
            const unsigned ky = __reduce_max_sync(0xffffffffu, c > R + 3 ? ((unsigned)__double2hiint(y) & 0x7fffffffu) : 0u);
            if (lane == 0 && ky) atomicMax(&kmx[1], ky);
          } else if (tid < 2 * NC && has2) {
            const PanelRow<TWO> r2(tiles, L, R + 2), r3(tiles, L, R + 3);
            const double x = (c > R + 2 && c < m) ? r2.at(c) : 0.0;
            u2[c] = x;
            v2[c] = (c > R + 3 && c < m) ? r3.at(c) : 0.0;
            const unsigned kx = __reduce_max_sync(0xffffffffu, c > R + 3 ? ((unsigned)__double2hiint(x) & 0x7fffffffu) : 0u);
            if (lane == 0 && kx) atomicMax(&kmx[2], kx);
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
    // second pair: its pivot is the entry (R+2, R+3) after the first pair's update -- every thread evaluates it, and the
    // guard compares it with a BOUND on the updated row R+2 (|row R+2| + |w2| |row R| / |piv1| + |t2| |row R+1|), so no
    // reduction over the updated row and no barrier is needed before the decision
    double t2 = 0.0, t3 = 0.0, w2 = 0.0, w3 = 0.0, piv2 = 0.0, inv2 = 0.0;
    bool quad = false;
    if (has2) {
      t2 = u[R + 2] * inv1;
      t3 = u[R + 3] * inv1;
      w2 = v[R + 2];
      w3 = v[R + 3];
      piv2 = u2[R + 3] + (w2 * t3 - t2 * w3);
      const double bound = __hiloint2double((int)kmx[2], 0) + fabs(w2 * inv1) * __hiloint2double((int)kmx[0], 0) +
                           fabs(t2) * __hiloint2double((int)kmx[1], 0);
      quad = fabs(piv2) >= kStaticPivotRatio * bound && piv2 != 0.0;
      if (quad) inv2 = fast_rcp(piv2);
    }
    pf *= quad ? piv1 * piv2 : piv1;
    const int dead = quad ? R + 3 : R + 1;  // indices <= dead leave the matrix with this pass
    if (tid < NC) {
      const int c = tid;
      const bool live = c > dead;
      const double uc = u[c], wc = v[c];
      double nu = 0.0, nv = 0.0;
      if (quad) {
        const double tc = uc * inv1;
        nu = u2[c] + (w2 * tc - t2 * wc);
        nv = v2[c] + (w3 * tc - t3 * wc);
      }
      zv[c] = live ? uc : 0.0;
      zv[2 * kPanelVS + c] = live ? wc : 0.0;
      zv[kPanelVS + c] = live ? nu : 0.0;
      zv[3 * kPanelVS + c] = live ? nv : 0.0;
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
    PP_ADD(1);
    if (tid == 0) kmx[0] = kmx[1] = kmx[2] = 0u;  // every reader is past its use; the next pass's atomics follow the barrier below
    if (dead + 1 < m)
      panel_update<TWO>(tiles, L, tri, (dead + 1) >> 3, avec,
                        (tg & 2) ? ((tg & 1) ? -inv2 : inv2) : ((tg & 1) ? -inv1 : inv1), bvec);
    __syncthreads();
    PP_ADD(2);
    R = dead + 1;
  }
  return stop;
}

// Phase 2: S = A22 + U^T (J U) into the packed tiles [0, CT (CT + 1) / 2) (layout of a CT-tile-row matrix).
// Work unit = up to 4 consecutive tiles of one tile row (they share the A fragment); the unit list is walked round-robin.
template <class Src>
__device__ void panel_schur(const Src& src, long long item, int m, double2* tiles, const double* pinv,
                            const short* units) {
  const int MT = (m + 7) >> 3, CT = MT - kPanelHT;
  const int lane = lane_id(), w = warp_id(), gr = lane >> 2, tg = lane & 3, nw = (int)(blockDim.x >> 5);
  const double sgn = (tg & 1) ? 1.0 : -1.0;   // V[2s] = -U[2s+1] / p_s,  V[2s+1] = +U[2s] / p_s, folded into the A fragment
  const int nunits = units[kPanelMaxUnits * 2];
  for (int un = w; un < nunits; un += nw) {
    const int ti = units[2 * un], tj0 = units[2 * un + 1];
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
    const double* pv = pinv + (tg >> 1);
    if (cnt == 4) {
#pragma unroll 4
      for (int kk = 0; kk < kPanelH / 4; ++kk) {
        // rows 4kk + tg: tile row kk >> 1 (stride CT tiles), in-tile row 4 (kk & 1) + tg (stride 4 double2 = 8 doubles)
        const int off = ((kk >> 1) * CT * 32 + (kk & 1) * 16) * 2;
        const double a = pa[off] * (sgn * pv[2 * kk]);
        const double b0 = pb[off], b1 = pb[off + 64], b2 = pb[off + 128], b3 = pb[off + 192];
        dmma_acc(acc[0], a, b0);
        dmma_acc(acc[1], a, b1);
        dmma_acc(acc[2], a, b2);
        dmma_acc(acc[3], a, b3);
      }
    } else {
      for (int kk = 0; kk < kPanelH / 4; ++kk) {
        const int off = ((kk >> 1) * CT * 32 + (kk & 1) * 16) * 2;
        const double a = pa[off] * (sgn * pv[2 * kk]);
#pragma unroll
        for (int q = 0; q < 3; ++q)
          if (q < cnt) dmma_acc(acc[q], a, pb[off + q * 64]);
      }
    }
    // S tile (ti, tj0 + q): the A11 tiles are dead, but other warps may still be reading U -> S goes to its own region
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (q < cnt) tiles[bt_idx(CT, ti, tj0 + q) * 32 + lane] = make_double2(acc[q][0], acc[q][1]);
  }
}

// the units of panel_schur for CT tile rows, longest rows first: (ti, tj0) pairs, count at [2 * kPanelMaxUnits]
__device__ __forceinline__ void panel_build_units(short* units, int CT) {
  int n = 0;
  for (int ti = 0; ti < CT; ++ti)
    for (int tj0 = ti; tj0 < CT; tj0 += 4) {
      units[2 * n] = (short)ti;
      units[2 * n + 1] = (short)tj0;
      ++n;
    }
  units[2 * kPanelMaxUnits] = (short)n;
}

// Signed Pfaffian of matrix `item` (128 < m <= 256) by the whole block; valid in every thread.
// smem: tiles (136 + 16 CT) * 32 double2 | vec 8 * kPanelVS doubles | pinv 64 doubles.  slab: fallback storage.
template <class Src>
__device__ double panel_pfaffian(const Src& src, long long item, int m, double2* tiles, double* vec, double* pinv,
                                 double2* slab) {
  unsigned* kmx = reinterpret_cast<unsigned*>(pinv + 64);           // 4 words
  const short* units = reinterpret_cast<const short*>(pinv + 66);   // 2 * 40 + 1 shorts
  const PanelTri* tri = reinterpret_cast<const PanelTri*>(pinv + 90);  // [0]: panel (n = 16), [1]: S (n = CT)
  if (threadIdx.x == 0) kmx[0] = kmx[1] = kmx[2] = 0u;
  const int MT = (m + 7) >> 3, CT = MT - kPanelHT;
  const int lane = lane_id(), w = warp_id(), gr = lane >> 2, tg = lane & 3, nw = (int)(blockDim.x >> 5);
  PP_T0();
  // ---- phase 1: panel
  for (int ti = w; ti < kPanelHT; ti += nw) {
#pragma unroll 4
    for (int tj = ti; tj < MT; ++tj)
      tiles[panel_tile(CT, ti, tj) * 32 + lane] = make_double2(src.at(item, 8 * ti + gr, 8 * tj + 2 * tg),
                                                               src.at(item, 8 * ti + gr, 8 * tj + 2 * tg + 1));
  }
  __syncthreads();
  PP_ADD(3);
  double pf = 1.0;
  const PanelLayout<true> L1{CT, MT};
  const int done1 = panel_static<true>(tiles, L1, tri[0], m, kPanelH, vec, pinv, kmx, pf);
  if (done1 < kPanelH) {  // block-uniform: every thread evaluates the same guards on the same shared vectors
    __syncthreads();
    return block_tiles_pfaffian<256>(src, item, m, slab, vec);
  }
  // ---- phase 2: Schur complement
  PP_RESET();
  panel_schur(src, item, m, tiles, pinv, units);
  __syncthreads();
  PP_ADD(4);
  // ---- phase 3: S, in place
  const int m2 = m - kPanelH;
  const PanelLayout<false> L2{CT, CT};
  const int done2 = panel_static<false>(tiles, L2, tri[1], m2, m2, vec, nullptr, kmx, pf);
  if (done2 < m2) {
    __syncthreads();
    return bt_pivoted_tail<256>(tiles, CT, m2, done2, vec, pf);
  }
  return pf;
}

__host__ __device__ constexpr size_t panel_smem_bytes(int m) {
  return ((size_t)(kPanelA11 + kPanelHT * (((m + 7) >> 3) - kPanelHT)) * 32) * sizeof(double2) +
         (size_t)(8 * kPanelVS + 64 + 2 + 24 + 16) * sizeof(double);
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
  if (threadIdx.x == 0) {
    panel_build_units(reinterpret_cast<short*>(pinv + 66), CT);
    panel_build_tri(reinterpret_cast<PanelTri*>(pinv + 90), kPanelHT);
    panel_build_tri(reinterpret_cast<PanelTri*>(pinv + 90) + 1, CT);
  }
  __syncthreads();
  for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
    const double pf = panel_pfaffian(src, item, m, tiles, vec, pinv, slab);
    if (threadIdx.x == 0) out[item] = out_mode == MCB200_PF_SIGNED ? scale * pf : abs_floor(scale * pf, floor2);
#ifdef MCB200_PANEL_PROFILE
    if (blockIdx.x == 0 && threadIdx.x == 0 && item == (long long)gridDim.x * 3) {
      printf("panel phases (cycles, 4 matrices): extract %lld  vectors %lld  update %lld  load %lld  schur %lld\n", g_pp[0],
             g_pp[1], g_pp[2], g_pp[3], g_pp[4]);
      for (int k = 0; k < 8; ++k) g_pp[k] = 0;
    }
#endif
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
  if (threadIdx.x == 0) {
    panel_build_units(reinterpret_cast<short*>(pinv + 66), CT);
    panel_build_tri(reinterpret_cast<PanelTri*>(pinv + 90), kPanelHT);
    panel_build_tri(reinterpret_cast<PanelTri*>(pinv + 90) + 1, CT);
  }
  __syncthreads();
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

// ---------------------------------------------------------------------------------------------------------------------
// 64 < m <= 128: the same pass structure on the whole matrix (no panel split): 136 tiles in shared memory, blocks of 256
// threads, two blocks per SM; threshold pivoting in place, the pivoted tail of pfaffian_block_tiles.cuh only for rows that
// have no acceptable pivot at all.
constexpr int kStatic128Threads = 256;

template <class Src>
__device__ double static128_pfaffian(const Src& src, long long item, int m, double2* tiles, double* vec, unsigned* kmx,
                                     const PanelTri& tri) {
  const int MT = (m + 7) >> 3;
  const int lane = lane_id(), w = warp_id(), gr = lane >> 2, tg = lane & 3, nw = (int)(blockDim.x >> 5);
  if (threadIdx.x == 0) kmx[0] = kmx[1] = kmx[2] = 0u;
  for (int ti = w; ti < MT; ti += nw) {
#pragma unroll 4
    for (int tj = ti; tj < MT; ++tj)
      tiles[bt_idx(MT, ti, tj) * 32 + lane] = make_double2(src.at(item, 8 * ti + gr, 8 * tj + 2 * tg),
                                                           src.at(item, 8 * ti + gr, 8 * tj + 2 * tg + 1));
  }
  __syncthreads();
  double pf = 1.0;
  const PanelLayout<false> L{MT, MT};
  const int done = panel_static<false, 128>(tiles, L, tri, m, m, vec, nullptr, kmx, pf);
  if (done < m) {
    __syncthreads();
    return bt_pivoted_tail<128>(tiles, MT, m, done, vec, pf);
  }
  return pf;
}

__host__ __device__ constexpr size_t static128_smem_bytes(int m) {
  return (size_t)(((m + 7) >> 3) * (((m + 7) >> 3) + 1) / 2) * 32 * sizeof(double2) +
         (size_t)(8 * kPanelVS + 2 + 8) * sizeof(double);
}

template <class Src>
__global__ void __launch_bounds__(kStatic128Threads, 2) pfaffian_static128_kernel(Src src, long long n_items, int m,
                                                                                 int out_mode, double scale,
                                                                                 double floor2, double* out) {
  extern __shared__ __align__(16) unsigned char pn_smem[];
  const int MT = (m + 7) >> 3, NT = MT * (MT + 1) / 2;
  double2* tiles = reinterpret_cast<double2*>(pn_smem);
  double* vec = reinterpret_cast<double*>(tiles + NT * 32);
  unsigned* kmx = reinterpret_cast<unsigned*>(vec + 8 * kPanelVS);
  PanelTri* tri = reinterpret_cast<PanelTri*>(vec + 8 * kPanelVS + 2);
  if (threadIdx.x == 0) panel_build_tri(tri, MT);
  __syncthreads();
  for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
    const double pf = static128_pfaffian(src, item, m, tiles, vec, kmx, *tri);
    if (threadIdx.x == 0) out[item] = out_mode == MCB200_PF_SIGNED ? scale * pf : abs_floor(scale * pf, floor2);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kStatic128Threads, 2) gram_static128_kernel(GramBtArgs g) {
  extern __shared__ __align__(16) unsigned char pn_smem[];
  const int m = g.d, MT = (m + 7) >> 3, NT = MT * (MT + 1) / 2;
  double2* tiles = reinterpret_cast<double2*>(pn_smem);
  double* vec = reinterpret_cast<double*>(tiles + NT * 32);
  unsigned* kmx = reinterpret_cast<unsigned*>(vec + 8 * kPanelVS);
  PanelTri* tri = reinterpret_cast<PanelTri*>(vec + 8 * kPanelVS + 2);
  if (threadIdx.x == 0) panel_build_tri(tri, MT);
  __syncthreads();
  const bool upper = g.n0 < g.n1;
  for (long long cell = blockIdx.x; cell < g.rows * g.n1; cell += gridDim.x) {
    const long long i = g.row_begin + cell / g.n1, j = cell % g.n1;
    if (g.mode != MCB200_GRAM_FULL && !(upper ? (j > i) : (i > j))) continue;
    GramPairSrc src{g.cov0 + i * (long long)m * m, g.cov1 + j * (long long)m * m, m};
    const double pf = static128_pfaffian(src, 0, m, tiles, vec, kmx, *tri);
    if (threadIdx.x == 0) {
      const double val = abs_floor(g.scale * pf, g.floor2);
      g.K[i * g.ldk + j] = val;
      if (g.mode == MCB200_GRAM_REFERENCE && j < g.n0 && i < g.n1) g.K[j * g.ldk + i] = val;  // gram_matrix.py:139-143
    }
    __syncthreads();
  }
}

template <class Src>
static int static128_launch(const Src& src, long long n_items, int m, int out_mode, double scale, double* out,
                            cudaStream_t st, const char* what) {
  if (n_items == 0) return 0;
  const size_t smem = static128_smem_bytes(m);
  auto kernel = pfaffian_static128_kernel<Src>;
  if (int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), what)) return rc;
  const long long per_sm = (long long)(kMaxDynSmem / (smem + 1024)) < 4 ? (long long)(kMaxDynSmem / (smem + 1024)) : 4;
  const long long cap = per_sm * kNumSMs;
  kernel<<<(unsigned)(n_items < cap ? n_items : cap), kStatic128Threads, smem, st>>>(src, n_items, m, out_mode, scale,
                                                                                    g_abs_floor2, out);
  return check_launch(what);
}

int static128_pfaffian_plain(const double* A, long long n, int m, int mode, double scale, double* out, cudaStream_t st) {
  PlainSrc s{A, m};
  return static128_launch(s, n, m, mode, scale, out, st, "pfaffian_static128_kernel");
}
int static128_pfaffian_probs(const double* cov, const int8_t* outcomes, long long B, int Q, int k, double* out,
                             cudaStream_t st) {
  ProbsSrc s{cov, outcomes, Q, k, 2 * k};
  return static128_launch(s, B * Q, 2 * k, MCB200_PF_ABS, ldexp(1.0, -k), out, st, "probs_static128_kernel");
}
int static128_pfaffian_sector(const double* cov, long long B, int D, const int32_t* index_sets, int T, int sz,
                              double* feats, cudaStream_t st) {
  SectorSrc s{cov, index_sets, D, T, sz};
  return static128_launch(s, B * T, sz, MCB200_PF_SIGNED, 1.0, feats, st, "sector_static128_kernel");
}
int static128_gram(const double* cov0, const double* cov1, long long n0, long long n1, int d, long long row_begin,
                   long long row_end, int mode, double scale, double* K, long long ldk, cudaStream_t st) {
  const long long rows = row_end - row_begin;
  if (rows <= 0 || n1 <= 0) return 0;
  GramBtArgs g{cov0, cov1, n0, n1, d, row_begin, rows, mode, scale, K, ldk, g_abs_floor2};
  const size_t smem = static128_smem_bytes(d);
  if (int rc = check_cuda(cudaFuncSetAttribute(gram_static128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)smem), "cudaFuncSetAttribute(gram_static128)"))
    return rc;
  const long long per_sm = (long long)(kMaxDynSmem / (smem + 1024)) < 4 ? (long long)(kMaxDynSmem / (smem + 1024)) : 4;
  const long long cells = rows * n1, cap = per_sm * kNumSMs;
  gram_static128_kernel<<<(unsigned)(cells < cap ? cells : cap), kStatic128Threads, smem, st>>>(g);
  return check_launch("gram_static128_kernel");
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
