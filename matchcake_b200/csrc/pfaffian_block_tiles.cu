// K3 for 64 < m <= 256: one block per matrix, the upper 8x8 tiles held in DMMA fragment layout (tile t, lane l ->
// double2 at tiles[t * 32 + l], i.e. conflict-free / fully coalesced 16-byte accesses) in SHARED memory for m <= 128
// (4 warps) and in an L2-resident slab of the caller's workspace for 128 < m <= 256 (8 warps, no shared-memory limit
// on the number of resident blocks).  Every index is a run-time value: the same code serves any m and both orders.
//
//  * static order (fast path): double step D eliminates the pairs (4D, 4D+1), (4D+2, 4D+3) with their natural pivots
//    and applies both rank-2 updates as ONE K = 4 DMMA per live tile (see pfaffian_tiles.cuh for the algebra and
//    the guard |pivot| >= 1e-3 max_j |row_j|);
//  * pivoted continuation: when a guard fails the remaining Schur complement is finished by Parlett-Reid with logical
//    pivoting (active bit-set, no swaps, sign from population counts), one rank-2 DMMA per tile that still holds an
//    active row.
// The warps of the block split the tile rows; vectors (rows being eliminated, tau / v) go through a few hundred bytes of
// shared memory and a block barrier.  Replaces the shared-memory scalar kernels for this size range (utils/_pfaffian.py
// semantics as in pfaffian_tiles.cuh).
#include <stdlib.h>

#include "common.cuh"
#include "pfaffian_block_tiles.cuh"
#include "tiles_api.cuh"

namespace mcb200 {

// GLOBAL = false: tiles in dynamic shared memory (m <= 128).  GLOBAL = true: tiles in the block's slab of `slabs`
// (gridDim.x slabs of MT(MT+1)/2 * 512 bytes, L2 resident), only the four vectors in shared memory.
template <class Src, int MAXM, bool GLOBAL>
__global__ void __launch_bounds__(MAXM) pfaffian_block_tiles_kernel(Src src, long long n_items, int m, int out_mode,
                                                                     double scale, double floor2, double* out,
                                                                     double2* slabs) {
  extern __shared__ __align__(16) unsigned char bt_smem[];
  const int MT = (m + 7) >> 3, NT = MT * (MT + 1) / 2;
  double2* tiles = GLOBAL ? slabs + (size_t)blockIdx.x * NT * 32 : reinterpret_cast<double2*>(bt_smem);
  double* vec = GLOBAL ? reinterpret_cast<double*>(bt_smem)
                       : reinterpret_cast<double*>(reinterpret_cast<double2*>(bt_smem) + NT * 32);
  for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
    const double pf = block_tiles_pfaffian<MAXM>(src, item, m, tiles, vec);
    if (threadIdx.x == 0) out[item] = out_mode == MCB200_PF_SIGNED ? scale * pf : abs_floor(scale * pf, floor2);
  }
}

template <int MAXM, bool GLOBAL>
__global__ void __launch_bounds__(MAXM) gram_block_tiles_kernel(GramBtArgs g, double2* slabs) {
  extern __shared__ __align__(16) unsigned char bt_smem[];
  const int m = g.d, MT = (m + 7) >> 3, NT = MT * (MT + 1) / 2;
  double2* tiles = GLOBAL ? slabs + (size_t)blockIdx.x * NT * 32 : reinterpret_cast<double2*>(bt_smem);
  double* vec = GLOBAL ? reinterpret_cast<double*>(bt_smem)
                       : reinterpret_cast<double*>(reinterpret_cast<double2*>(bt_smem) + NT * 32);
  const bool upper = g.n0 < g.n1;
  for (long long cell = blockIdx.x; cell < g.rows * g.n1; cell += gridDim.x) {
    const long long i = g.row_begin + cell / g.n1, j = cell % g.n1;
    if (g.mode != MCB200_GRAM_FULL && !(upper ? (j > i) : (i > j))) continue;
    GramPairSrc src{g.cov0 + i * (long long)m * m, g.cov1 + j * (long long)m * m, m};
    const double pf = block_tiles_pfaffian<MAXM>(src, 0, m, tiles, vec);
    if (threadIdx.x == 0) {
      const double val = abs_floor(g.scale * pf, g.floor2);
      g.K[i * g.ldk + j] = val;
      if (g.mode == MCB200_GRAM_REFERENCE && j < g.n0 && i < g.n1) g.K[j * g.ldk + i] = val;  // gram_matrix.py:139-143
    }
  }
}

// blocks that share the L2-resident slabs of the large variant (4 per SM = 160 MB at m = 256; the variant is bound by L2
// bandwidth -- ~11 MB of tile traffic per 256 x 256 matrix -- so more resident blocks only add 7 %)
constexpr long long kBtGlobalCtas = 4LL * kNumSMs;

// 128 < m <= 256: the panel-blocked kernel (pfaffian_panel.cu) unless MCB200_PF256_SLAB=1 asks for the slab variant
// (kept for A/B measurements and as the panel kernel's pivoted fallback)
static bool use_slab_variant() {
  static const bool v = [] {
    const char* e = getenv("MCB200_PF256_SLAB");
    return e != nullptr && e[0] == '1';
  }();
  return v;
}

// 64 < m <= 128: the pass kernels of pfaffian_panel.cu (5.8 against 4.3 M Pfaffians/s at m = 128, 9.0 against 8.1 at m = 96,
// threshold pivoting in place) unless MCB200_PF128_OLD=1 asks for the round-1 kernels of this file (A/B measurements)
static bool use_static128() {
  static const bool v = [] {
    const char* e = getenv("MCB200_PF128_OLD");
    return !(e != nullptr && e[0] == '1');
  }();
  return v;
}

long long block_tiles_workspace_bytes(long long n_items, int m) {
  if (m <= 128) return 0;
  if (!use_slab_variant()) return panel_workspace_bytes(n_items, m);
  const int MT = (m + 7) / 8, NT = MT * (MT + 1) / 2;
  const long long ctas = n_items < kBtGlobalCtas ? n_items : kBtGlobalCtas;
  return ctas * (long long)NT * 32 * (long long)sizeof(double2);
}

template <class Src>
static int launch_block_tiles(const Src& src, long long n_items, int m, int out_mode, double scale, double* out,
                              void* workspace, cudaStream_t st, const char* what) {
  if (n_items == 0) return 0;
  const int MT = (m + 7) / 8, NT = MT * (MT + 1) / 2;
  if (m <= 128) {
    const size_t smem = (size_t)NT * 32 * sizeof(double2) + 4 * (size_t)(128 + 4) * sizeof(double);
    auto kernel = pfaffian_block_tiles_kernel<Src, 128, false>;
    if (int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(pfaffian_block_tiles)"))
      return rc;
    const long long cap = 6LL * kNumSMs;
    kernel<<<(unsigned)(n_items < cap ? n_items : cap), 128, smem, st>>>(src, n_items, m, out_mode, scale, g_abs_floor2, out, nullptr);
    return check_launch(what);
  }
  if (workspace == nullptr) return set_error("%s: m=%d needs a workspace (see *_workspace_bytes)", what, m);
  const size_t smem = 4 * (size_t)(256 + 4) * sizeof(double);
  const long long ctas = n_items < kBtGlobalCtas ? n_items : kBtGlobalCtas;
  pfaffian_block_tiles_kernel<Src, 256, true><<<(unsigned)ctas, 256, smem, st>>>(src, n_items, m, out_mode, scale, g_abs_floor2, out,
                                                                                  (double2*)workspace);
  return check_launch(what);
}

int block_tiles_gram(const double* cov0, const double* cov1, long long n0, long long n1, int d, long long row_begin,
                     long long row_end, int mode, double scale, double* K, long long ldk, void* workspace,
                     cudaStream_t st) {
  const long long rows = row_end - row_begin;
  if (rows <= 0 || n1 <= 0) return 0;
  const int MT = (d + 7) / 8, NT = MT * (MT + 1) / 2;
  if (d > 128 && !use_slab_variant())
    return panel_gram(cov0, cov1, n0, n1, d, row_begin, row_end, mode, scale, K, ldk, workspace, st);
  if (d <= 128 && use_static128()) return static128_gram(cov0, cov1, n0, n1, d, row_begin, row_end, mode, scale, K, ldk, st);
  GramBtArgs g{cov0, cov1, n0, n1, d, row_begin, rows, mode, scale, K, ldk, g_abs_floor2};
  const long long cells = rows * n1;
  if (d <= 128) {
    const size_t smem = (size_t)NT * 32 * sizeof(double2) + 4 * (size_t)(128 + 4) * sizeof(double);
    auto kernel = gram_block_tiles_kernel<128, false>;
    if (int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(gram_block_tiles)"))
      return rc;
    const long long cap = 6LL * kNumSMs;
    kernel<<<(unsigned)(cells < cap ? cells : cap), 128, smem, st>>>(g, nullptr);
    return check_launch("gram_block_tiles_kernel");
  }
  if (workspace == nullptr) return set_error("gram: d=%d needs a workspace (mcb200_gram_workspace_bytes)", d);
  const size_t smem = 4 * (size_t)(256 + 4) * sizeof(double);
  const long long ctas = cells < kBtGlobalCtas ? cells : kBtGlobalCtas;
  gram_block_tiles_kernel<256, true><<<(unsigned)ctas, 256, smem, st>>>(g, (double2*)workspace);
  return check_launch("gram_block_tiles_kernel");
}

int block_tiles_pfaffian_plain(const double* A, long long n, int m, int mode, double scale, double* out, void* workspace,
                               cudaStream_t st) {
  if (m > 128 && !use_slab_variant()) return panel_pfaffian_plain(A, n, m, mode, scale, out, workspace, st);
  if (m <= 128 && use_static128()) return static128_pfaffian_plain(A, n, m, mode, scale, out, st);
  PlainSrc s{A, m};
  return launch_block_tiles(s, n, m, mode, scale, out, workspace, st, "pfaffian_block_tiles_kernel");
}

int block_tiles_pfaffian_probs(const double* cov, const int8_t* outcomes, long long B, int Q, int k, double* out,
                               cudaStream_t st) {
  if (use_static128()) return static128_pfaffian_probs(cov, outcomes, B, Q, k, out, st);
  ProbsSrc s{cov, outcomes, Q, k, 2 * k};
  return launch_block_tiles(s, B * Q, 2 * k, MCB200_PF_ABS, ldexp(1.0, -k), out, nullptr, st, "probs_block_tiles_kernel");
}

int block_tiles_pfaffian_sector(const double* cov, long long B, int D, const int32_t* index_sets, int T, int s,
                                double* feats, cudaStream_t st) {
  if (use_static128()) return static128_pfaffian_sector(cov, B, D, index_sets, T, s, feats, st);
  SectorSrc src{cov, index_sets, D, T, s};
  return launch_block_tiles(src, B * T, s, MCB200_PF_SIGNED, 1.0, feats, nullptr, st, "sector_block_tiles_kernel");
}

}  // namespace mcb200
