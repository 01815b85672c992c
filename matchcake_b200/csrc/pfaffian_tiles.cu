// K3 fast path for m <= 64: warp-per-matrix Pfaffians with the matrix held as DMMA accumulator fragments
// (pfaffian_tiles.cuh).  Entry points are internal (declared in tiles_api.cuh) and are dispatched to from the
// public C ABI in pfaffian.cu.
#include <stdlib.h>

#include "common.cuh"
#include "pfaffian_src.cuh"
#include "pfaffian_tiles.cuh"
#include "pfaffian_flat.cuh"
#include "tiles_api.cuh"

namespace mcb200 {

template <int MT, class Src>
__device__ __forceinline__ void load_fragments(const Src& src, long long item, double (&c)[Tiles<MT>::NT][2]) {
  const int lane = lane_id(), gr = lane >> 2, tg = lane & 3;
#pragma unroll
  for (int ti = 0; ti < MT; ++ti)
#pragma unroll
    for (int tj = ti; tj < MT; ++tj) {
      c[Tiles<MT>::idx(ti, tj)][0] = src.at(item, 8 * ti + gr, 8 * tj + 2 * tg);
      c[Tiles<MT>::idx(ti, tj)][1] = src.at(item, 8 * ti + gr, 8 * tj + 2 * tg + 1);
    }
}

template <int MT, class Src>
__global__ void __launch_bounds__(128) pfaffian_tiles_kernel(Src src, long long n_items, int m, int out_mode,
                                                             double scale, double floor2, double* out) {
  extern __shared__ double smem[];
  double* scratch = smem + (size_t)warp_id() * Tiles<MT>::kScratch;
  const int warps = blockDim.x >> 5;
  for (long long item = (long long)blockIdx.x * warps + warp_id(); item < n_items;
       item += (long long)gridDim.x * warps) {
    double c[Tiles<MT>::NT][2];
    load_fragments<MT>(src, item, c);
    double pf = warp_pfaffian_tiles<MT>(c, m, scratch);
    if (lane_id() == 0) out[item] = out_mode == MCB200_PF_SIGNED ? scale * pf : abs_floor(scale * pf, floor2);
    __syncwarp();
  }
}

template <class Src>
static int launch_tiles(const Src& src, long long n_items, int m, int out_mode, double scale, double* out,
                        cudaStream_t st, const char* what) {
  if (n_items == 0) return 0;
  const int warps = 4;
  long long blocks = (n_items + warps - 1) / warps;
  const long long cap = 16LL * kNumSMs;
  if (blocks > cap) blocks = cap;
  auto go = [&](auto kernel, int mt) -> int {
    const size_t smem = (size_t)warps * (2 * 8 * mt + 4 * (8 * mt + 8)) * sizeof(double);
    kernel<<<(unsigned)blocks, warps * 32, smem, st>>>(src, n_items, m, out_mode, scale, g_abs_floor2, out);
    return check_launch(what);
  };
  if (m <= 8) return go(pfaffian_tiles_kernel<1, Src>, 1);
  if (m <= 16) return go(pfaffian_tiles_kernel<2, Src>, 2);
  if (m <= 24) return go(pfaffian_tiles_kernel<3, Src>, 3);
  if (m <= 32) return go(pfaffian_tiles_kernel<4, Src>, 4);
  if (m <= 48) return go(pfaffian_tiles_kernel<6, Src>, 6);
  return go(pfaffian_tiles_kernel<8, Src>, 8);
}

int tiles_pfaffian_plain(const double* A, long long n, int m, int mode, double scale, double* out, cudaStream_t st) {
  PlainSrc s{A, m};
  return launch_tiles(s, n, m, mode, scale, out, st, "pfaffian_tiles_kernel");
}

int tiles_pfaffian_probs(const double* cov, const int8_t* outcomes, long long B, int Q, int k, double* out,
                         cudaStream_t st) {
  ProbsSrc s{cov, outcomes, Q, k, 2 * k};
  return launch_tiles(s, B * Q, 2 * k, MCB200_PF_ABS, ldexp(1.0, -k), out, st, "probs_tiles_kernel");
}

int tiles_pfaffian_sector(const double* cov, long long B, int D, const int32_t* index_sets, int T, int s,
                          double* feats, cudaStream_t st) {
  SectorSrc src{cov, index_sets, D, T, s};
  return launch_tiles(src, B * T, s, MCB200_PF_SIGNED, 1.0, feats, st, "sector_tiles_kernel");
}

// ------------------------------------------------------------------------------------------- Gram
// Packed covariance: for every sample the NT upper tiles in fragment order, [sample][tile][lane] as double2
// (c0, c1), zero beyond the real matrix -> a warp reads one tile as 32 coalesced 16-byte loads.
template <int MT>
__global__ void pack_tiles_kernel(const double* __restrict__ cov, long long n, int m, double2* __restrict__ packed) {
  constexpr int NT = Tiles<MT>::NT;
  const long long total = n * NT * 32;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int lane = (int)(e & 31);
    const long long rest = e >> 5;
    const int t = (int)(rest % NT);
    const long long smp = rest / NT;
    // invert t -> (ti, tj)
    int ti = 0, base = 0;
    while (t >= base + (MT - ti)) {
      base += MT - ti;
      ++ti;
    }
    const int tj = ti + (t - base);
    const int i = 8 * ti + (lane >> 2), j = 8 * tj + 2 * (lane & 3);
    const double* a = cov + smp * (long long)m * m;
    auto at = [&](int r, int cc) -> double {
      if (r >= m || cc >= m || r == cc) return 0.0;
      return r < cc ? a[r * m + cc] : -a[cc * m + r];
    };
    packed[e] = make_double2(at(i, j), at(i, j + 1));
  }
}

struct GramTileArgs {
  const double2* p0;
  const double2* p1;
  long long n0, n1;
  int m;
  long long row_begin, row_end;
  int mode;
  double scale;
  double* K;
  long long ldk;
  double floor2;
  double growth;  // bound of the static block pivots (mcb200_set_gram_pivot_growth)
};

// One warp per (i, j) cell.  A block of TI x TJ warps walks (row-block, column-block) tiles so that the fragments
// of the TI + TJ covariances it touches stay in L1.  SR tile-rows of the working matrix live in shared memory.
template <int MT, int SR>
struct GramSmem {
  static constexpr int kSmTiles = SR > 0 ? Tiles<MT>::idx(SR - 1, MT - 1) + 1 : 0;     // tiles with ti < SR
  static constexpr int kPerWarp = Tiles<MT>::kScratch + kSmTiles * 64;                  // doubles
};

template <int MT, int SR, int TI, int TJ, int MINB, bool PHYS = true>
__global__ void __launch_bounds__(32 * TI * TJ, MINB) gram_tiles_kernel(GramTileArgs g) {
  constexpr int NT = Tiles<MT>::NT;
  constexpr int SMT = GramSmem<MT, SR>::kSmTiles;
  extern __shared__ __align__(16) double smem[];
  double* scratch = smem + (size_t)warp_id() * GramSmem<MT, SR>::kPerWarp;
  const int lane = lane_id();
  double2* sm = reinterpret_cast<double2*>(scratch + Tiles<MT>::kScratch) + lane;
  const int wi = warp_id() / TJ, wj = warp_id() % TJ;
  const long long rows = g.row_end - g.row_begin;
  const long long rb_n = (rows + TI - 1) / TI, cb_n = (g.n1 + TJ - 1) / TJ;
  const bool upper = g.n0 < g.n1;  // reference rule: evaluate j > i, else i > j (gram_matrix.py:191-195)
  for (long long t = blockIdx.x; t < rb_n * cb_n; t += gridDim.x) {
    const long long rb = t / cb_n, cb = t - rb * cb_n;
    const long long i_lo = g.row_begin + rb * TI, j_lo = cb * TJ;
    if (g.mode != MCB200_GRAM_FULL) {  // skip tiles that hold no evaluated cell (uniform over the block)
      if (!upper && j_lo >= i_lo + TI) continue;              // all j >= i
      if (upper && j_lo + TJ - 1 <= i_lo) continue;           // all j <= i
    }
    const long long i = i_lo + wi, j = j_lo + wj;
    bool on = i < g.row_end && j < g.n1;
    if (g.mode != MCB200_GRAM_FULL) on = on && (upper ? (j > i) : (i > j));
    if (!on) continue;
    const double2* a = g.p0 + i * (long long)(NT * 32) + lane;
    const double2* b = g.p1 + j * (long long)(NT * 32) + lane;
    double c[NT][2];
#pragma unroll
    for (int q = 0; q < NT; ++q) {
      const double2 x = __ldg(a + q * 32), y = __ldg(b + q * 32);
      if (q < SMT) {
        sm[q * 32] = make_double2(x.x + y.x, x.y + y.y);
      } else {
        c[q][0] = x.x + y.x;
        c[q][1] = x.y + y.y;
      }
    }
    // Lambda_i + Lambda_j: growth bound 1e5 (PHYS, calibrated on the kernel ansatz) or 1e3 (strict)
    const double pf = warp_pfaffian_tiles<MT, SR, false, PHYS>(c, g.m, scratch, sm);
    if (lane == 0) {
      const double v = abs_floor(g.scale * pf, g.floor2);
      g.K[i * g.ldk + j] = v;
      if (g.mode == MCB200_GRAM_REFERENCE && j < g.n0 && i < g.n1) g.K[j * g.ldk + i] = v;  // gram_matrix.py:139-143
    }
    __syncwarp();
  }
}

// ---- full-size matrices (m = 8 MT): branch-free static elimination (pfaffian_flat.cuh) -------------------------------
// A cell whose guard failed somewhere is redone from its inputs with partial pivoting (out of line: rare).
template <int MT, int SR>
__device__ __noinline__ double gram_cell_pivoted(const double2* a, const double2* b, int m, double* scratch, double2* sm) {
  constexpr int NT = Tiles<MT>::NT;
  constexpr int SMT = GramSmem<MT, SR>::kSmTiles;
  double c[NT][2];
#pragma unroll
  for (int q = 0; q < NT; ++q) {
    const double2 x = __ldg(a + q * 32), y = __ldg(b + q * 32);
    if (q < SMT) {
      sm[q * 32] = make_double2(x.x + y.x, x.y + y.y);
    } else {
      c[q][0] = x.x + y.x;
      c[q][1] = x.y + y.y;
    }
  }
  __syncwarp();
  return warp_pfaffian_tiles_pivoted<MT, SR>(c, m, scratch, sm, 0);
}

struct GramCellWalk {  // the (i, j) cells of one warp, in the order of gram_tiles_kernel
  long long rb_n, cb_n, total;
  bool upper;
  template <int TI, int TJ>
  __device__ __forceinline__ void init(const GramTileArgs& g) {
    const long long rows = g.row_end - g.row_begin;
    rb_n = (rows + TI - 1) / TI;
    cb_n = (g.n1 + TJ - 1) / TJ;
    total = rb_n * cb_n;
    upper = g.n0 < g.n1;  // reference rule: evaluate j > i, else i > j (gram_matrix.py:191-195)
  }
  template <int TI, int TJ>
  __device__ __forceinline__ bool cell(const GramTileArgs& g, long long t, long long& i, long long& j) const {
    const long long rb = t / cb_n, cb = t - rb * cb_n;
    const long long i_lo = g.row_begin + rb * TI, j_lo = cb * TJ;
    if (g.mode != MCB200_GRAM_FULL) {
      if (!upper && j_lo >= i_lo + TI) return false;
      if (upper && j_lo + TJ - 1 <= i_lo) return false;
    }
    i = i_lo + warp_id() / TJ;
    j = j_lo + warp_id() % TJ;
    bool on = i < g.row_end && j < g.n1;
    if (g.mode != MCB200_GRAM_FULL) on = on && (upper ? (j > i) : (i > j));
    return on;
  }
};

__device__ __forceinline__ void gram_emit(const GramTileArgs& g, long long i, long long j, double pf) {
  if (lane_id() == 0) {
    const double v = abs_floor(g.scale * pf, g.floor2);
    g.K[i * g.ldk + j] = v;
    if (g.mode == MCB200_GRAM_REFERENCE && j < g.n0 && i < g.n1) g.K[j * g.ldk + i] = v;  // gram_matrix.py:139-143
  }
}

// One cell at a time, branch-free steps.
template <int MT, int SR, int TI, int TJ>
__global__ void __launch_bounds__(32 * TI * TJ, 1) gram_tiles_flat_kernel(GramTileArgs g) {
  constexpr int NT = Tiles<MT>::NT;
  constexpr int SMT = GramSmem<MT, SR>::kSmTiles;
  extern __shared__ __align__(16) double smem[];
  double* scratch = smem + (size_t)warp_id() * GramSmem<MT, SR>::kPerWarp;
  const int lane = lane_id();
  double2* sm = reinterpret_cast<double2*>(scratch + Tiles<MT>::kScratch) + lane;
  const StaticLane<MT> L = static_lane_setup<MT>(scratch);
  GramCellWalk w;
  w.init<TI, TJ>(g);
  for (long long t = blockIdx.x; t < w.total; t += gridDim.x) {
    long long i, j;
    if (!w.cell<TI, TJ>(g, t, i, j)) continue;
    const double2* a = g.p0 + i * (long long)(NT * 32) + lane;
    const double2* b = g.p1 + j * (long long)(NT * 32) + lane;
    double c[NT][2];
#pragma unroll
    for (int q = 0; q < NT; ++q) {
      const double2 x = __ldg(a + q * 32), y = __ldg(b + q * 32);
      if (q < SMT) {
        sm[q * 32] = make_double2(x.x + y.x, x.y + y.y);
      } else {
        c[q][0] = x.x + y.x;
        c[q][1] = x.y + y.y;
      }
    }
    double pf = 1.0;
    if (!warp_pfaffian_flat_steps<MT, SR>(c, scratch, sm, pf, L, g.growth)) pf = gram_cell_pivoted<MT, SR>(a, b, g.m, scratch, sm);
    gram_emit(g, i, j, pf);
    __syncwarp();
  }
}

long long tiles_gram_workspace_bytes(long long n0, long long n1, int d) {
  if (d > 64) return 0;
  const int mt = d <= 8 ? 1 : d <= 16 ? 2 : d <= 24 ? 3 : d <= 32 ? 4 : d <= 48 ? 6 : 8;
  return (n0 + n1) * (long long)(mt * (mt + 1) / 2) * 32 * (long long)sizeof(double2);
}

template <int MT>
static int gram_tiles_launch(const double* cov0, const double* cov1, long long n0, long long n1, int d,
                             long long row_begin, long long row_end, int mode, double scale, double* K,
                             long long ldk, void* workspace, cudaStream_t st) {
  constexpr int NT = Tiles<MT>::NT;
  double2* p0 = (double2*)workspace;
  double2* p1 = p0 + n0 * (long long)NT * 32;
  const bool same = (cov0 == cov1 && n0 == n1);
  auto pack = [&](const double* cov, long long n, double2* dst) -> int {
    long long blocks = (n * NT * 32 + 255) / 256;
    if (blocks > 32LL * kNumSMs) blocks = 32LL * kNumSMs;
    pack_tiles_kernel<MT><<<(unsigned)blocks, 256, 0, st>>>(cov, n, d, dst);
    return check_launch("pack_tiles_kernel");
  };
  if (int rc = pack(cov0, n0, p0)) return rc;
  if (same) p1 = p0;
  else if (int rc = pack(cov1, n1, p1)) return rc;
  GramTileArgs g{p0, p1, n0, n1, d, row_begin, row_end, mode, scale, K, ldk, g_abs_floor2, g_gram_growth};
  const bool phys = g_gram_growth > kStaticBlockGrowth;  // compile-time bounds of the kernels with a branch per quadruple
  const long long rows = row_end - row_begin;
  auto launch = [&](auto kernel, int ti, int tj, int per_sm, size_t per_warp_doubles) -> int {
    const size_t smem = (size_t)ti * tj * per_warp_doubles * sizeof(double);
    const long long tiles = ((rows + ti - 1) / ti) * ((n1 + tj - 1) / tj);
    long long blocks = tiles < (long long)per_sm * kNumSMs ? tiles : (long long)per_sm * kNumSMs;
    if (int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(gram_tiles)"))
      return rc;
    kernel<<<(unsigned)blocks, 32 * ti * tj, smem, st>>>(g);
    return check_launch("gram_tiles_kernel");
  };
  if constexpr (MT >= 8) {
    // 64 x 64: tile-row 0 of the working matrix in shared memory, one 12-warp block per SM at 168 registers.  Measured
    // on C3 (round 2, 4 x 4 block pivots): 249.5 ms against 281 ms (two tile-rows in shared memory, 16 warps at 128
    // registers, spills), 270 ms (three tile-rows, 14 warps) and 259 ms (three tile-rows, 16 warps): more resident warps
    // do not pay for the shared-memory traffic of the tiles they evict from the registers.
    // full 64 x 64 matrices: branch-free steps, 8 warps per SM with every tile in registers (255): 57.3 ms against
    // 61.9 ms on a 4096-sample Gram of config C3 (59.6 ms at 12 warps / 168 registers).  MCB200_GRAM_FLAT=0 selects the
    // one-branch-per-quadruple kernel below (A/B measurements, scripts/gram_flat_check.py).
    const char* env = getenv("MCB200_GRAM_FLAT");
    if (d == 8 * MT && (env == nullptr || atoi(env) != 0))
      return launch(gram_tiles_flat_kernel<MT, 0, 2, 4>, 2, 4, 1, GramSmem<MT, 0>::kPerWarp);
    if (!phys) return launch(gram_tiles_kernel<MT, 1, 3, 4, 1, false>, 3, 4, 1, GramSmem<MT, 1>::kPerWarp);
    return launch(gram_tiles_kernel<MT, 1, 3, 4, 1>, 3, 4, 1, GramSmem<MT, 1>::kPerWarp);
  } else {
    if (!phys) return launch(gram_tiles_kernel<MT, 0, 2, 8, 1, false>, 2, 8, 2, GramSmem<MT, 0>::kPerWarp);
    return launch(gram_tiles_kernel<MT, 0, 2, 8, 1>, 2, 8, 2, GramSmem<MT, 0>::kPerWarp);
  }
}

int tiles_gram(const double* cov0, const double* cov1, long long n0, long long n1, int d, long long row_begin,
               long long row_end, int mode, double scale, double* K, long long ldk, void* workspace,
               cudaStream_t st) {
  if (workspace == nullptr) return set_error("gram: workspace is NULL (see mcb200_gram_workspace_bytes)");
  if (d <= 8) return gram_tiles_launch<1>(cov0, cov1, n0, n1, d, row_begin, row_end, mode, scale, K, ldk, workspace, st);
  if (d <= 16) return gram_tiles_launch<2>(cov0, cov1, n0, n1, d, row_begin, row_end, mode, scale, K, ldk, workspace, st);
  if (d <= 24) return gram_tiles_launch<3>(cov0, cov1, n0, n1, d, row_begin, row_end, mode, scale, K, ldk, workspace, st);
  if (d <= 32) return gram_tiles_launch<4>(cov0, cov1, n0, n1, d, row_begin, row_end, mode, scale, K, ldk, workspace, st);
  if (d <= 48) return gram_tiles_launch<6>(cov0, cov1, n0, n1, d, row_begin, row_end, mode, scale, K, ldk, workspace, st);
  return gram_tiles_launch<8>(cov0, cov1, n0, n1, d, row_begin, row_end, mode, scale, K, ldk, workspace, st);
}

}  // namespace mcb200

// ------------------------------------------------------------------------------------------- probs over ALL outcomes
// qml.probs(wires) asks for p(y) = 2^-k |Pf(Lambda|_W + Lambda_y)| for every y in {0,1}^k (nif_device.py:430-479).
// Eliminating the wire pairs in order WITHOUT pivoting, the pivot of wire j is
//     piv_j = A^(j)[0][1] -+ 1 = 2 p(y_j | y_0..y_{j-1}),
// and it only depends on the outcome prefix, so the 2^k Pfaffians share a binary tree of partial eliminations:
// 2^(j+1) rank-2 updates of size (2k-2j-2)^2 at level j instead of 2^k full Pfaffians (70x less work at k = 8).
// A vanishing pivot means the whole subtree has probability 0, so no pivoting is needed: sub-trees whose
// conditional probability is below 1e-14 are set to exactly 0 (absolute error <= 5e-15 on outcomes that the
// reference returns as ~1e-16 noise).  One warp per circuit instance, nodes ping-pong through shared memory.
namespace mcb200 {

__global__ void __launch_bounds__(128) probs_tree_kernel(const double* __restrict__ cov, long long B, int k,
                                                         int normalize, int buf_doubles, double floor2,
                                                         double* __restrict__ out) {
  extern __shared__ double smem[];
  const int m = 2 * k, Q = 1 << k;
  const int lane = lane_id(), warps = blockDim.x >> 5;
  double* base = smem + (size_t)warp_id() * (2 * (size_t)buf_doubles + 2 * (size_t)Q);
  double* bufA = base;
  double* bufB = base + buf_doubles;
  double* fA = base + 2 * (size_t)buf_doubles;
  double* fB = fA + Q;
  for (long long b = (long long)blockIdx.x * warps + warp_id(); b < B; b += (long long)gridDim.x * warps) {
    const double* a = cov + b * (long long)m * m;
    for (int e = lane; e < m * m; e += 32) bufA[e] = a[e];
    if (lane == 0) fA[0] = 1.0;
    __syncwarp();
    double* cur = bufA; double* nxt = bufB; double* fc = fA; double* fn = fB;
    for (int j = 0; j < k; ++j) {
      const int s = m - 2 * j, s2 = s - 2, nodes = 1 << j;
      // node strides padded to odd: with lanes over the children (deep levels) consecutive parents / children then fall
      // into different banks (unpadded, 16 x 16 or 64 doubles apart, they all hit one: ncu counted 61 % of the kernel's
      // shared-memory wavefronts as conflicts)
      const int ss = (s * s) | 1, ss2 = (s2 * s2) | 1, ne2 = s2 * s2;
      // children c = 2 * node + y
      for (int c = lane; c < 2 * nodes; c += 32) {
        const double* T = cur + (size_t)(c >> 1) * ss;
        const double piv = T[1] + ((c & 1) ? 1.0 : -1.0);
        const double f = fc[c >> 1];
        fn[c] = (fabs(piv) < 1e-14 || f == 0.0) ? 0.0 : f * piv;
      }
      __syncwarp();
      if (s2 > 0) {
        if (ne2 >= 64) {  // lanes over the elements of one child
          for (int c = 0; c < 2 * nodes; ++c) {
            const double* T = cur + (size_t)(c >> 1) * ss;
            double* U = nxt + (size_t)c * ss2;
            const double piv = T[1] + ((c & 1) ? 1.0 : -1.0);
            const bool dead = fn[c] == 0.0;
            const double inv = dead ? 0.0 : 1.0 / piv;
            for (int e = lane; e < ne2; e += 32) {
              const int ra = e / s2, rb = e - ra * s2;
              if (rb > ra) {
                const double ta = T[ra + 2] * inv, tb = T[rb + 2] * inv;          // tau = T[0][.] / piv
                U[e] = T[(ra + 2) * s + rb + 2] - ta * T[s + rb + 2] + T[s + ra + 2] * tb;
              }
            }
          }
        } else {          // lanes over the children
          for (int c = lane; c < 2 * nodes; c += 32) {
            const double* T = cur + (size_t)(c >> 1) * ss;
            double* U = nxt + (size_t)c * ss2;
            const double piv = T[1] + ((c & 1) ? 1.0 : -1.0);
            const bool dead = fn[c] == 0.0;
            const double inv = dead ? 0.0 : 1.0 / piv;
            for (int ra = 0; ra < s2 - 1; ++ra) {
              const double ta = T[ra + 2] * inv, wa = T[s + ra + 2];
              for (int rb = ra + 1; rb < s2; ++rb)
                U[ra * s2 + rb] = T[(ra + 2) * s + rb + 2] - ta * T[s + rb + 2] + wa * (T[rb + 2] * inv);
            }
          }
        }
      }
      __syncwarp();
      double* t = cur; cur = nxt; nxt = t;
      t = fc; fc = fn; fn = t;
    }
    // leaves: fc[q] = prod of pivots = Pf; p = 2^-k |Pf|
    double sum = 0.0;
    for (int q = lane; q < Q; q += 32) {
      const double p = abs_floor(ldexp(fc[q], -k), floor2);
      fc[q] = p;
      sum += p;
    }
    if (normalize) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    double* o = out + b * (long long)Q;
    for (int q = lane; q < Q; q += 32) o[q] = normalize ? fc[q] / sum : fc[q];
    __syncwarp();
  }
}

int tiles_probs_all(const double* cov, long long B, int k, int normalize, double* out, cudaStream_t st) {
  const int m = 2 * k;
  size_t buf = (size_t)m * m + 1;
  for (int j = 0; j <= k; ++j) {
    const size_t w = ((size_t)1 << j) * (((size_t)(m - 2 * j) * (m - 2 * j)) | 1);  // padded node stride, see the kernel
    if (w > buf) buf = w;
  }
  const size_t per_warp = (2 * buf + 2 * ((size_t)1 << k)) * sizeof(double);
  int warps = 4;
  while (warps > 1 && warps * per_warp > 96 * 1024) warps >>= 1;
  const size_t smem = warps * per_warp;
  if (smem > (size_t)kMaxDynSmem) return -1;
  if (int rc = check_cuda(cudaFuncSetAttribute(probs_tree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                          "cudaFuncSetAttribute(probs_tree)"))
    return rc;
  long long blocks = (B + warps - 1) / warps;
  if (blocks > 16LL * kNumSMs) blocks = 16LL * kNumSMs;
  probs_tree_kernel<<<(unsigned)blocks, warps * 32, smem, st>>>(cov, B, k, normalize, (int)buf, g_abs_floor2, out);
  return check_launch("probs_tree_kernel");
}

}  // namespace mcb200

