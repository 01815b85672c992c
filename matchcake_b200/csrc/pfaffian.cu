// K3: the C ABI of the batched Pfaffians (plain, projector / probs read-out, Pauli-sector gather, Gram pairs) and the
// small kernels around them.  Dispatch by matrix size:
//   m <= 64          register-tile kernels, one warp per matrix          (pfaffian_tiles.cu / .cuh)
//   64 < m <= 128    shared-memory-tile kernels, one block per matrix    (pfaffian_block_tiles.cu)
//   128 < m <= 256   the same kernels with the tiles in an L2-resident slab of the caller's workspace
//                    (plain matrices and Gram cells only)
//   m > 256          scalar Parlett-Reid, one block per matrix, working copy in the caller's workspace
//                    (plain matrices and Gram cells only; pfaffian_dev.cuh)
// Matrices are formed on the fly from their sources so that sums / gathers never round-trip through HBM:
//   Plain   A[i]                              utils/_pfaffian.py:78-119
//   Probs   cov[b] + Lambda_y(q)              product_state_strategy.py:196-213
//   Sector  cov[b][S_t, S_t]                  utils/_pfaffian.py:39-75
//   Gram    cov0[i] + cov1[j]                 nif_kernel.py:210-221 via SURVEY.md A.5
#include "common.cuh"
#include "pfaffian_dev.cuh"
#include "tiles_api.cuh"

namespace mcb200 {

// ------------------------------------------------------------------------------------------- loaders
struct PlainLoader {
  const double* A;
  double* out;
  double scale;
  int mode;
  double floor2;
  __device__ void load(long long item, double* S, int lds, int m, int tid, int nthr) const {
    const double* src = A + item * (long long)m * m;
    for (int idx = tid; idx < m * m; idx += nthr) {
      int i = idx / m, j = idx - i * m;
      if (i < j) S[i * lds + j] = 0.5 * (src[idx] - src[j * m + i]);  // skew part, see pfaffian_src.cuh
    }
  }
  __device__ void store(long long item, double pf) const {
    out[item] = mode == MCB200_PF_SIGNED ? scale * pf : abs_floor(scale * pf, floor2);
  }
};

// ------------------------------------------------------------------------------------------- kernels
template <class Loader>
__global__ void pfaffian_block_global_kernel(Loader loader, long long n_items, int m, double* workspace) {
  __shared__ double red[8];
  const int lds = m + 1;
  double* S = workspace + (size_t)blockIdx.x * m * lds;
  for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
    loader.load(item, S, lds, m, threadIdx.x, blockDim.x);
    __syncthreads();
    double pf = block_pfaffian_upper(S, m, lds, red);
    if (threadIdx.x == 0) loader.store(item, pf);
    __syncthreads();
  }
}

__global__ void fill_kernel(double* out, long long n, double v) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = v;
}

__global__ void normalize_rows_kernel(double* p, long long B, int Q) {
  // one warp per row (nif_device.py:478)
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp_id();
  if (row >= B) return;
  double* r = p + row * Q;
  double s = 0.0;
  for (int q = lane_id(); q < Q; q += 32) s += r[q];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  for (int q = lane_id(); q < Q; q += 32) r[q] = r[q] / s;
}

__global__ void sector_gather2_kernel(const double* cov, long long B, int D, const int32_t* index_sets, int T,
                                      double* feats) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * T) return;
  long long b = e / T;
  int t = (int)(e - b * T);
  feats[e] = cov[b * (long long)D * D + (long long)index_sets[2 * t] * D + index_sets[2 * t + 1]];
}

// out[b] (+)= sum_t feats[b, t] * coeffs[t]   (deterministic order; m_pfaffian_expval_strategy.py:304)
__global__ void rowdot_kernel(const double* feats, const double* coeffs, long long B, int T, int accumulate,
                              double* out) {
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp_id();
  if (row >= B) return;
  double s = 0.0;
  for (int t = lane_id(); t < T; t += 32) s += feats[row * T + t] * coeffs[t];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane_id() == 0) out[row] = accumulate ? out[row] + s : s;
}

// Fused Hamiltonian read-out: out[b] (+)= sum_t coeffs[t] * Pf(cov[b][S_t, S_t]) over index sets
// of MIXED sizes 0, 2, 4, 6 in one launch (m_pfaffian_expval_strategy.py:270-306 loops over sectors and calls
// sector_pfaffian_features + a tensordot per sector).  One block per instance, threads over terms, closed-form
// Pfaffians, deterministic tree reduction.  Larger sectors go through mcb200_sector_features + mcb200_rowdot.
__device__ __forceinline__ double pf4(double a01, double a02, double a03, double a12, double a13, double a23) {
  return fma(a03, a12, fma(-a02, a13, a01 * a23));
}

__global__ void pauli_sum_small_kernel(const double* __restrict__ cov, long long B, int D,
                                       const int32_t* __restrict__ term_ptr, const int32_t* __restrict__ indices,
                                       const double* __restrict__ coeffs, int T, int accumulate, double* out) {
  __shared__ double partial[8];
  for (long long b = blockIdx.x; b < B; b += gridDim.x) {
    const double* a = cov + b * (long long)D * D;
    double s = 0.0;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      const int o = term_ptr[t], sz = term_ptr[t + 1] - o;
      const int32_t* ix = indices + o;
      double pf = 0.0;
      if (sz == 0) {
        pf = 1.0;
      } else if (sz == 2) {
        pf = a[(long long)ix[0] * D + ix[1]];
      } else if (sz == 4) {
        const long long r0 = (long long)ix[0] * D, r1 = (long long)ix[1] * D, r2 = (long long)ix[2] * D;
        pf = pf4(a[r0 + ix[1]], a[r0 + ix[2]], a[r0 + ix[3]], a[r1 + ix[2]], a[r1 + ix[3]], a[r2 + ix[3]]);
      } else if (sz == 6) {
        double u[6][6];
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = i + 1; j < 6; ++j) u[i][j] = a[(long long)ix[i] * D + ix[j]];
        // expansion along row 0: Pf = sum_j (-1)^(j+1) a_0j Pf(A without rows/cols 0 and j)
        pf = u[0][1] * pf4(u[2][3], u[2][4], u[2][5], u[3][4], u[3][5], u[4][5]);
        pf = fma(-u[0][2], pf4(u[1][3], u[1][4], u[1][5], u[3][4], u[3][5], u[4][5]), pf);
        pf = fma(u[0][3], pf4(u[1][2], u[1][4], u[1][5], u[2][4], u[2][5], u[4][5]), pf);
        pf = fma(-u[0][4], pf4(u[1][2], u[1][3], u[1][5], u[2][3], u[2][5], u[3][5]), pf);
        pf = fma(u[0][5], pf4(u[1][2], u[1][3], u[1][4], u[2][3], u[2][4], u[3][4]), pf);
      }  // odd sizes: Pfaffian 0 (utils/_pfaffian.py:64-66)
      s = fma(coeffs[t], pf, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const int warps = blockDim.x >> 5;
    if (lane_id() == 0) partial[warp_id()] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < warps; ++w) tot += partial[w];
      out[b] = accumulate ? out[b] + tot : tot;
    }
    __syncthreads();
  }
}

__global__ void enumerate_outcomes_kernel(int8_t* bits, int k) {
  // states_to_binary (nif_device.py:151-165): outcome index -> bits, MSB first
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ((long long)k << k)) return;
  long long q = e / k;
  int j = (int)(e - q * k);
  bits[e] = (int8_t)((q >> (k - 1 - j)) & 1);
}

constexpr int kWarpKernelMaxM = kBlockTilesMaxM;  // largest matrix of the tile kernels
constexpr int kBlockGlobalCtas = 4 * kNumSMs;

template <class Loader>
static int launch_pfaffian(const Loader& loader, long long n_items, int m, double* workspace, cudaStream_t st,
                           const char* what) {
  if (n_items == 0) return 0;
  if (workspace == nullptr) return set_error("%s: m=%d needs a workspace (see *_workspace_bytes)", what, m);
  long long blocks = n_items < kBlockGlobalCtas ? n_items : kBlockGlobalCtas;
  pfaffian_block_global_kernel<Loader><<<(unsigned)blocks, 256, 0, st>>>(loader, n_items, m, workspace);
  return check_launch(what);
}

static long long pfaffian_ws_bytes(long long n_items, int m) {
  if (m <= kBlockTilesGlobalMaxM) return block_tiles_workspace_bytes(n_items, m);
  long long blocks = n_items < kBlockGlobalCtas ? n_items : kBlockGlobalCtas;
  return blocks * (long long)m * (m + 1) * (long long)sizeof(double);
}

// ------------------------------------------------------------------------------------------- Gram
// K[i][j] = scale |Pf(cov0[i] + cov1[j])|.  blockIdx.y = row (relative to row_begin), the warps of a block take
// consecutive columns so that cov0[i] stays hot in L1/L2.
struct GramArgs {
  const double* cov0;
  const double* cov1;
  long long n0, n1;
  int d;
  long long row_begin;
  int mode;
  double scale;
  double* K;
  long long ldk;
  double floor2;
};

__device__ __forceinline__ bool gram_cell_active(const GramArgs& g, long long i, long long j) {
  if (j >= g.n1) return false;
  if (g.mode == MCB200_GRAM_FULL) return true;
  return (g.n0 < g.n1) ? (j > i) : (i > j);  // gram_matrix.py:191-195 (REFERENCE and TRIANGLE)
}

__device__ __forceinline__ void gram_store(const GramArgs& g, long long i, long long j, double v) {
  g.K[i * g.ldk + j] = v;
  if (g.mode == MCB200_GRAM_REFERENCE && j < g.n0 && i < g.n1) g.K[j * g.ldk + i] = v;  // gram_matrix.py:139-143
}

__global__ void gram_block_global_kernel(GramArgs g, long long rows, double* workspace) {
  __shared__ double red[8];
  const int m = g.d, lds = m + 1;
  double* S = workspace + (size_t)blockIdx.x * m * lds;
  for (long long cell = blockIdx.x; cell < rows * g.n1; cell += gridDim.x) {
    const long long i = g.row_begin + cell / g.n1;
    const long long j = cell % g.n1;
    if (!gram_cell_active(g, i, j)) continue;
    const double* a = g.cov0 + i * (long long)m * m;
    const double* b = g.cov1 + j * (long long)m * m;
    for (int idx = threadIdx.x; idx < m * m; idx += blockDim.x) {
      int r = idx / m, c = idx - r * m;
      if (r < c) S[r * lds + c] = a[idx] + b[idx];
    }
    __syncthreads();
    double pf = block_pfaffian_upper(S, m, lds, red);
    if (threadIdx.x == 0) gram_store(g, i, j, abs_floor(g.scale * pf, g.floor2));
    __syncthreads();
  }
}

// gram_matrix.py:127-161: mirror the computed triangle inside the leading min(n0,n1)^2 block, diagonal := 1.
// One 32 x 32 tile pair per block: the source tile (bj, bi) is read row-wise (coalesced), transposed through shared
// memory and written row-wise into the target tile (bi, bj); 8 warps, 4 rows each.
__global__ void __launch_bounds__(256) gram_symmetrize_kernel(double* K, long long q, long long ldk, int fill_lower) {
  __shared__ double tile[32][33];
  // block index -> (a, b) with a <= b over the tiles of the q x q block
  const long long nt = (q + 31) / 32;
  long long t = blockIdx.x;
  // row a of the upper-triangular tile enumeration holds nt - a tiles
  long long a = (long long)((2.0 * nt + 1.0 - sqrt((2.0 * nt + 1.0) * (2.0 * nt + 1.0) - 8.0 * (double)t)) * 0.5);
  while (a > 0 && a * nt - a * (a - 1) / 2 > t) --a;
  while ((a + 1) * nt - (a + 1) * a / 2 <= t) ++a;
  const long long b = a + (t - (a * nt - a * (a - 1) / 2));
  // target tile = the one in the triangle that is NOT evaluated: upper (a, b) when the evaluated cells are i > j
  // (n0 >= n1), lower (b, a) when they are j > i (n0 < n1)
  const long long ti = fill_lower ? b : a, tj = fill_lower ? a : b;   // target tile (rows ti, cols tj)
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {   // source tile (tj, ti): rows 32 tj + r, cols 32 ti + tx
    const long long i = 32 * tj + r, j = 32 * ti + tx;
    tile[r][tx] = (i < q && j < q) ? K[i * ldk + j] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {   // target rows 32 ti + r, cols 32 tj + tx  <-  source[tx][r]
    const long long i = 32 * ti + r, j = 32 * tj + tx;
    if (i >= q || j >= q) continue;
    if (i == j) K[i * ldk + j] = 1.0;
    else if (fill_lower ? (i > j) : (j > i)) K[i * ldk + j] = tile[tx][r];
  }
}

__global__ void gram_diag_kernel(GramArgs g, long long row_end) {
  long long i = g.row_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < row_end && i < g.n1) g.K[i * g.ldk + i] = 1.0;  // gram_matrix.py:147-161
}

}  // namespace mcb200

// ---------------------------------------------------------------------------------------------- C ABI
using namespace mcb200;

extern "C" int64_t mcb200_pfaffian_workspace_bytes(int64_t n, int32_t m) { return pfaffian_ws_bytes(n, m); }

extern "C" int mcb200_pfaffian(const double* A, int64_t n, int32_t m, int32_t mode, double scale, double* out,
                               void* workspace, void* stream) {
  if (n < 0 || m < 0) return set_error("pfaffian: negative size");
  if (mode != MCB200_PF_ABS && mode != MCB200_PF_SIGNED) return set_error("pfaffian: unknown mode %d", mode);
  if (n == 0) return 0;
  if (out == nullptr) return set_error("pfaffian: out is NULL");
  cudaStream_t st = as_stream(stream);
  if (m == 0 || (m & 1)) {  // empty -> 1, odd -> 0 (tests/test_utils/test_pfaffian.py:66-75,121-127)
    fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out, n, m == 0 ? scale : 0.0);
    return check_launch("fill_kernel");
  }
  if (A == nullptr) return set_error("pfaffian: A is NULL");
  if (m <= kTilesMaxM) return tiles_pfaffian_plain(A, n, m, mode, scale, out, st);
  if (m <= kBlockTilesGlobalMaxM) return block_tiles_pfaffian_plain(A, n, m, mode, scale, out, workspace, st);
  PlainLoader ld{A, out, scale, mode, g_abs_floor2};
  return launch_pfaffian(ld, n, m, (double*)workspace, st, "pfaffian_kernel");
}

extern "C" int mcb200_probs(const double* cov, const int8_t* outcomes, int64_t B, int32_t Q, int32_t k,
                            int32_t normalize, double* out, void* stream) {
  if (B < 0 || Q < 0 || k < 1) return set_error("probs: bad sizes");
  if (B == 0 || Q == 0) return 0;
  if (cov == nullptr || outcomes == nullptr || out == nullptr) return set_error("probs: NULL pointer");
  if (2 * k > kWarpKernelMaxM) return set_error("probs: k=%d measured wires exceed the fused limit of %d", k,
                                                 kWarpKernelMaxM / 2);
  cudaStream_t st = as_stream(stream);
  if (2 * k <= kTilesMaxM) {
    if (int rc = tiles_pfaffian_probs(cov, outcomes, B, Q, k, out, st)) return rc;
  } else {
    if (int rc = block_tiles_pfaffian_probs(cov, outcomes, B, Q, k, out, st)) return rc;
  }
  if (normalize) {
    normalize_rows_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(out, B, Q);
    return check_launch("normalize_rows_kernel");
  }
  return 0;
}

extern "C" int mcb200_sector_features(const double* cov, int64_t B, int32_t D, const int32_t* index_sets,
                                      int32_t T, int32_t s, double* feats, void* stream) {
  if (B < 0 || T < 0 || s < 0 || D < 1) return set_error("sector_features: bad sizes");
  if (B == 0 || T == 0) return 0;
  if (feats == nullptr) return set_error("sector_features: feats is NULL");
  cudaStream_t st = as_stream(stream);
  const long long n = (long long)B * T;
  if (s == 0 || (s & 1)) {
    fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(feats, n, s == 0 ? 1.0 : 0.0);
    return check_launch("fill_kernel");
  }
  if (cov == nullptr || index_sets == nullptr) return set_error("sector_features: NULL pointer");
  if (s == 2) {  // fast path utils/_pfaffian.py:70-72
    sector_gather2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cov, B, D, index_sets, T, feats);
    return check_launch("sector_gather2_kernel");
  }
  if (s > kWarpKernelMaxM) return set_error("sector_features: sector size %d > %d", s, kWarpKernelMaxM);
  if (s <= kTilesMaxM) return tiles_pfaffian_sector(cov, B, D, index_sets, T, s, feats, st);
  return block_tiles_pfaffian_sector(cov, B, D, index_sets, T, s, feats, st);
}

extern "C" int mcb200_rowdot(const double* feats, const double* coeffs, int64_t B, int32_t T, int32_t accumulate,
                             double* out, void* stream) {
  if (B < 0 || T < 0) return set_error("rowdot: bad sizes");
  if (B == 0) return 0;
  if (out == nullptr || (T > 0 && (feats == nullptr || coeffs == nullptr))) return set_error("rowdot: NULL pointer");
  rowdot_kernel<<<(unsigned)((B + 7) / 8), 256, 0, as_stream(stream)>>>(feats, coeffs, B, T, accumulate, out);
  return check_launch("rowdot_kernel");
}

extern "C" int mcb200_pauli_sum_small(const double* cov, int64_t B, int32_t D, const int32_t* term_ptr,
                                      const int32_t* indices, const double* coeffs, int32_t T, int32_t accumulate,
                                      double* out, void* stream) {
  if (B < 0 || T < 0 || D < 1) return set_error("pauli_sum_small: bad sizes");
  if (B == 0) return 0;
  if (out == nullptr) return set_error("pauli_sum_small: out is NULL");
  if (T > 0 && (cov == nullptr || term_ptr == nullptr || coeffs == nullptr))
    return set_error("pauli_sum_small: NULL pointer");
  const int threads = T <= 32 ? 32 : T <= 64 ? 64 : T <= 128 ? 128 : 256;
  const long long cap = 32LL * kNumSMs;
  pauli_sum_small_kernel<<<(unsigned)(B < cap ? B : cap), threads, 0, as_stream(stream)>>>(
      cov, B, D, term_ptr, indices, coeffs, T, accumulate, out);
  return check_launch("pauli_sum_small_kernel");
}

extern "C" int64_t mcb200_gram_workspace_bytes(int64_t n0, int64_t n1, int32_t d) {
  if (d <= kTilesMaxM) return tiles_gram_workspace_bytes(n0, n1, d);
  if (d <= kBlockTilesGlobalMaxM) return block_tiles_workspace_bytes(n0 * n1, d);
  return (int64_t)kBlockGlobalCtas * d * (d + 1) * (int64_t)sizeof(double);
}

extern "C" int mcb200_gram(const double* cov0, const double* cov1, int64_t n0, int64_t n1, int32_t d,
                           int64_t row_begin, int64_t row_end, int32_t mode, double scale, double* K, int64_t ldk,
                           void* workspace, void* stream) {
  if (n0 < 0 || n1 < 0 || d < 2 || (d & 1)) return set_error("gram: bad sizes (n0=%lld n1=%lld d=%d)",
                                                               (long long)n0, (long long)n1, d);
  if (row_begin < 0 || row_end > n0 || row_begin > row_end) return set_error("gram: bad row range");
  if (mode != MCB200_GRAM_REFERENCE && mode != MCB200_GRAM_FULL && mode != MCB200_GRAM_TRIANGLE)
    return set_error("gram: unknown mode %d", mode);
  if (ldk < n1) return set_error("gram: ldk < n1");
  const long long rows = row_end - row_begin;
  if (rows == 0 || n1 == 0) return 0;
  if (cov0 == nullptr || cov1 == nullptr || K == nullptr) return set_error("gram: NULL pointer");
  cudaStream_t st = as_stream(stream);
  GramArgs g{cov0, cov1, n0, n1, d, row_begin, mode, scale, K, ldk, g_abs_floor2};
  if (d <= kTilesMaxM) {
    if (int rc = tiles_gram(cov0, cov1, n0, n1, d, row_begin, row_end, mode, scale, K, ldk, workspace, st)) return rc;
  } else if (d <= kBlockTilesGlobalMaxM) {
    if (int rc = block_tiles_gram(cov0, cov1, n0, n1, d, row_begin, row_end, mode, scale, K, ldk, workspace, st))
      return rc;
  } else {
    if (workspace == nullptr) return set_error("gram: d=%d needs a workspace (mcb200_gram_workspace_bytes)", d);
    long long blocks = rows * n1 < kBlockGlobalCtas ? rows * n1 : kBlockGlobalCtas;
    gram_block_global_kernel<<<(unsigned)blocks, 256, 0, st>>>(g, rows, (double*)workspace);
    if (int rc = check_launch("gram_block_global_kernel")) return rc;
  }
  if (mode == MCB200_GRAM_REFERENCE) {
    gram_diag_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(g, row_end);
    return check_launch("gram_diag_kernel");
  }
  return 0;
}

extern "C" int mcb200_gram_symmetrize(double* K, int64_t n0, int64_t n1, int64_t ldk, void* stream) {
  if (n0 < 0 || n1 < 0 || ldk < n1) return set_error("gram_symmetrize: bad sizes");
  const long long q = n0 < n1 ? n0 : n1;
  if (q == 0) return 0;
  if (K == nullptr) return set_error("gram_symmetrize: K is NULL");
  const long long nt = (q + 31) / 32;
  gram_symmetrize_kernel<<<(unsigned)(nt * (nt + 1) / 2), 256, 0, as_stream(stream)>>>(K, q, ldk, n0 < n1 ? 1 : 0);
  return check_launch("gram_symmetrize_kernel");
}

extern "C" int mcb200_probs_all(const double* cov, int64_t B, int32_t k, int32_t normalize, double* out,
                                void* workspace, void* stream) {
  if (B < 0 || k < 1 || k > 20) return set_error("probs_all: bad sizes (k=%d)", k);
  if (B == 0) return 0;
  if (cov == nullptr || out == nullptr) return set_error("probs_all: NULL pointer");
  cudaStream_t st = as_stream(stream);
  const int rc = tiles_probs_all(cov, B, k, normalize, out, st);
  if (rc >= 0) return rc;
  // wide marginals: enumerate the outcomes (workspace holds the 2^k x k bit table) and use the per-outcome path
  if (workspace == nullptr) return set_error("probs_all: k=%d needs a workspace (mcb200_probs_all_workspace_bytes)", k);
  const long long Q = 1LL << k;
  enumerate_outcomes_kernel<<<(unsigned)((Q * k + 255) / 256), 256, 0, st>>>((int8_t*)workspace, k);
  if (int rc2 = check_launch("enumerate_outcomes_kernel")) return rc2;
  return mcb200_probs(cov, (const int8_t*)workspace, B, (int32_t)Q, k, normalize, out, stream);
}

extern "C" int64_t mcb200_probs_all_workspace_bytes(int32_t k) { return k <= 10 ? 0 : ((int64_t)k << k); }
