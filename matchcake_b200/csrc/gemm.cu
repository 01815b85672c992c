// K1 (dense chain step) and K2 (dense covariance conjugation): strided-batched FP64 GEMM on the FP64 tensor
// pipe (mma.sync.m8n8k4.f64 -> DMMA), tiles staged through shared memory -- by the TMA engine (cp.async.bulk +
// mbarrier) for the per-instance SPTM products of up to 32 qubits, by register-prefetched loads for larger shapes.
//
//   C[b] (M x N) = op(A[b]) (M x K) @ op(B[b]) (K x N),   row-major, op = identity or transpose
//
// Block tile 64 x 64 x 16, 8 warps as 4 (M) x 2 (N); each warp owns a 16 x 32 patch = 2 x 4 DMMA tiles.
// Shared-memory leading dimensions are chosen so that the 8-byte fragment reads of a half-warp hit 16
// distinct bank pairs (ld = 4 mod 16).
#include "common.cuh"
#include "tiles_api.cuh"

namespace mcb200 {

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int LDA_S = BK + 4;   // As[m][k]
constexpr int LDB_S = BN + 4;   // Bs[k][n]

__device__ __forceinline__ void dmma8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

struct GemmArgs {
  const double* A;
  const double* B;
  double* C;
  long long sA, sB, sC;  // batch strides (elements), 0 = shared
  int M, N, K;
  int lda, ldb, ldc;
  long long batch;
  // optional sign per K index applied to op(A): -1, or +1 where kbits[k] != 0 (basis-state Lambda_0, cov_basis_gemm)
  int use_ksign = 0;
  const int8_t* kbits = nullptr;
};

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) bgemm_dmma_kernel(GemmArgs g) {
  __shared__ double As[BM * LDA_S];
  __shared__ double Bs[BK * LDB_S];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const long long b = blockIdx.z;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const double* A = g.A + b * g.sA;
  const double* B = g.B + b * g.sB;
  double* C = g.C + b * g.sC;

  double acc[2][4][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // global -> register staging: 4 elements of each operand tile per thread
  double ra[4], rb[4];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      int e = tid + 256 * t;
      // A tile element (mm, kk)
      int mm, kk;
      if (!TA) { mm = e / BK; kk = e % BK; } else { kk = e / BM; mm = e % BM; }
      int gm = m0 + mm, gk = k0 + kk;
      ra[t] = (gm < g.M && gk < g.K) ? (TA ? A[(long long)gk * g.lda + gm] : A[(long long)gm * g.lda + gk]) : 0.0;
      if (g.use_ksign && gk < g.K && !(g.kbits != nullptr && g.kbits[gk])) ra[t] = -ra[t];
      // B tile element (kk, nn)
      int nn;
      if (!TB) { kk = e / BN; nn = e % BN; } else { nn = e / BK; kk = e % BK; }
      int gn = n0 + nn;
      gk = k0 + kk;
      rb[t] = (gn < g.N && gk < g.K) ? (TB ? B[(long long)gn * g.ldb + gk] : B[(long long)gk * g.ldb + gn]) : 0.0;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      int e = tid + 256 * t;
      int mm, kk, nn;
      if (!TA) { mm = e / BK; kk = e % BK; } else { kk = e / BM; mm = e % BM; }
      As[mm * LDA_S + kk] = ra[t];
      if (!TB) { kk = e / BN; nn = e % BN; } else { nn = e / BK; kk = e % BK; }
      Bs[kk * LDB_S + nn] = rb[t];
    }
  };

  load_tiles(0);
  for (int k0 = 0; k0 < g.K; k0 += BK) {
    store_tiles();
    __syncthreads();
    if (k0 + BK < g.K) load_tiles(k0 + BK);
    const int gr = lane >> 2, tg = lane & 3;
#pragma unroll
    for (int ks = 0; ks < BK; ks += 4) {
      double af[2], bf[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) af[i] = As[(wm * 16 + i * 8 + gr) * LDA_S + ks + tg];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Bs[(ks + tg) * LDB_S + wn * 32 + j * 8 + gr];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncthreads();
  }
  const int gr = lane >> 2, tg = lane & 3;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int r = m0 + wm * 16 + i * 8 + gr;
      int c = n0 + wn * 32 + j * 8 + tg * 2;
      if (r < g.M) {
        if (c < g.N) C[(long long)r * g.ldc + c] = acc[i][j][0];
        if (c + 1 < g.N) C[(long long)r * g.ldc + c + 1] = acc[i][j][1];
      }
    }
}

// ---------------------------------------------------------------------------------------------- TMA-staged chain step
// One block per circuit instance for n <= 64 (<= 32 qubits): both operands are brought into shared memory by the TMA
// engine -- one cp.async.bulk per matrix row, landing in rows padded to n + 4 doubles so that the DMMA fragment reads
// stay conflict free -- and completion is signalled on an mbarrier; the 8 warps then compute the 64 x 64 product out
// of shared memory with m8n8k4 DMMA tiles.  No thread touches global memory for the operands.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

constexpr int TMA_N = 64;  // largest SPTM side of the TMA-staged path

// NT = padded matrix side (16, 32 or 64).  A block stages G = 8, 4 or 1 instances at a time; WPI = 8 / G warps share
// one instance, each owning a 16 x min(32, NT) patch of the product.
template <int NT, bool TA, bool TB>
__global__ void __launch_bounds__(256) sptm_matmul_tma_kernel(GemmArgs g) {
  constexpr int LD = NT + 4;
  constexpr int G = NT == 64 ? 1 : (NT == 32 ? 4 : 8);
  constexpr int WPI = 8 / G;
  constexpr int RB = NT / 16;            // 16-row blocks per instance
  constexpr int JT = NT >= 32 ? 4 : 2;   // 8-column tiles per warp patch
  constexpr int SLOT = 2 * NT * LD;      // doubles per instance: A then B
  extern __shared__ __align__(128) unsigned char tma_smem[];
  double* S = reinterpret_cast<double*>(tma_smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(S + G * SLOT);
  const int n = g.M, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slot = warp / WPI, wl = warp % WPI, wm = wl % RB, wn = wl / RB;
  const int gr = lane >> 2, tg = lane & 3;
  if (tid == 0) mbar_init(bar, 1);
  // rows/columns beyond n are read by the fragment loads of the last tiles: keep them finite
  for (int e = tid; e < G * SLOT; e += 256) S[e] = 0.0;
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy zero fill before async-proxy writes
  uint32_t phase = 0;
  const double* As = S + slot * SLOT;
  const double* Bs = As + NT * LD;
  for (long long b0 = (long long)blockIdx.x * G; b0 < g.batch; b0 += (long long)gridDim.x * G) {
    const int act = (int)min((long long)G, g.batch - b0);
    if (warp == 0) {
      if (lane == 0) mbar_expect_tx(bar, (uint32_t)act * 2u * (uint32_t)n * (uint32_t)n * 8u);
      __syncwarp();
      for (int idx = lane; idx < act * n; idx += 32) {
        const int sl = idx / n, r = idx - sl * n;
        double* dst = S + sl * SLOT + r * LD;
        tma_load_1d(dst, g.A + (b0 + sl) * g.sA + (long long)r * n, (uint32_t)n * 8u, bar);
        tma_load_1d(dst + NT * LD, g.B + (b0 + sl) * g.sB + (long long)r * n, (uint32_t)n * 8u, bar);
      }
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    double acc[2][JT][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < JT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int k0 = 0; k0 < n; k0 += 4) {
      double af[2], bf[JT];
      const int kk = k0 + tg;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int m = wm * 16 + i * 8 + gr;
        af[i] = TA ? As[kk * LD + m] : As[m * LD + kk];
      }
#pragma unroll
      for (int j = 0; j < JT; ++j) {
        const int c = wn * 32 + j * 8 + gr;
        bf[j] = TB ? Bs[c * LD + kk] : Bs[kk * LD + c];
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < JT; ++j) dmma8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    if (slot < act) {
      double* C = g.C + (b0 + slot) * g.sC;
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < JT; ++j) {
          const int r = wm * 16 + i * 8 + gr, c = wn * 32 + j * 8 + tg * 2;
          if (r < n && c < n) {  // n is even on this path, so c + 1 < n as well
            *reinterpret_cast<double2*>(C + (long long)r * g.ldc + c) = make_double2(acc[i][j][0], acc[i][j][1]);
          }
        }
    }
    __syncthreads();  // every warp is done reading the slots before the next stage's TMA writes land
  }
}

template <int NT>
static int launch_sptm_tma_nt(const GemmArgs& g, int transA, int transB, cudaStream_t st) {
  constexpr int G = NT == 64 ? 1 : (NT == 32 ? 4 : 8);
  const size_t smem = (size_t)G * 2 * NT * (NT + 4) * sizeof(double) + 16;
  const long long stages = (g.batch + G - 1) / G;
  const long long resident = (NT == 16 ? 5LL : 3LL) * kNumSMs;
  long long blocks = stages < resident ? stages : resident;
  auto go = [&](auto kernel) -> int {
    if (int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(sptm_matmul_tma)"))
      return rc;
    kernel<<<(unsigned)blocks, 256, smem, st>>>(g);
    return check_launch("sptm_matmul_tma_kernel");
  };
  if (!transA && !transB) return go(sptm_matmul_tma_kernel<NT, false, false>);
  if (transA && !transB) return go(sptm_matmul_tma_kernel<NT, true, false>);
  if (!transA && transB) return go(sptm_matmul_tma_kernel<NT, false, true>);
  return go(sptm_matmul_tma_kernel<NT, true, true>);
}

static int launch_sptm_tma(const GemmArgs& g, int transA, int transB, cudaStream_t st) {
  if (g.M <= 16) return launch_sptm_tma_nt<16>(g, transA, transB, st);
  if (g.M <= 32) return launch_sptm_tma_nt<32>(g, transA, transB, st);
  return launch_sptm_tma_nt<64>(g, transA, transB, st);
}

int launch_bgemm(const GemmArgs& g, int transA, int transB, long long batch, cudaStream_t st) {
  if (batch == 0 || g.M == 0 || g.N == 0) return 0;
  for (long long b0 = 0; b0 < batch; b0 += 65535) {
    long long nb = batch - b0 < 65535 ? batch - b0 : 65535;
    GemmArgs h = g;
    h.A += b0 * g.sA;
    h.B += b0 * g.sB;
    h.C += b0 * g.sC;
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, (unsigned)nb);
    if (!transA && !transB) bgemm_dmma_kernel<false, false><<<grid, 256, 0, st>>>(h);
    else if (transA && !transB) bgemm_dmma_kernel<true, false><<<grid, 256, 0, st>>>(h);
    else if (!transA && transB) bgemm_dmma_kernel<false, true><<<grid, 256, 0, st>>>(h);
    else bgemm_dmma_kernel<true, true><<<grid, 256, 0, st>>>(h);
    if (int rc = check_launch("bgemm_dmma_kernel")) return rc;
  }
  return 0;
}

// X[b] (D x m) = Rlift[b][:, idx],  Rlift = R (+) 1 when D == d + 1  (_extended_covariance.py:101-126)
__global__ void gather_columns_kernel(const double* R, long long sR, int d, int D, const int32_t* idx, int m,
                                      double* X) {
  const long long b = blockIdx.y;
  const double* r = R + b * sR;
  double* x = X + b * (long long)D * m;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < D * m; e += gridDim.x * blockDim.x) {
    int row = e / m, c = e - row * m;
    int col = idx ? idx[c] : c;
    double v;
    if (row < d && col < d) v = r[(long long)row * d + col];
    else v = (row == col) ? 1.0 : 0.0;
    x[e] = v;
  }
}

// cov <- cov - cov^T in place (one thread per pair i < j; the diagonal becomes 0)
__global__ void antisymmetrize_kernel(double* cov, long long B, int m) {
  const long long pairs = (long long)m * (m + 1) / 2;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < B * pairs;
       e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / pairs;
    long long t = e - b * pairs;
    int i = 0;
    while (t >= m - i) {  // row i holds m - i pairs (j = i .. m-1)
      t -= m - i;
      ++i;
    }
    const int j = i + (int)t;
    double* c = cov + b * (long long)m * m;
    if (i == j) {
      c[i * m + i] = 0.0;
    } else {
      const double x = c[i * m + j], y = c[j * m + i];
      c[i * m + j] = x - y;
      c[j * m + i] = y - x;
    }
  }
}

// K2 for large panels: Lambda = X^T Lambda_0 X with the 2x2-block-diagonal Lambda_0 of a basis state is P - P^T,
// P = E^T diag(s) O, E / O = even / odd rows of X (d x m, row stride 2m): one (m x N x m) DMMA GEMM, N = d / 2, instead
// of the two (d x d x m) products of the dense formulation (nif_device.py:1118-1131).
int cov_basis_gemm(const double* X, const int8_t* bits, long long B, int n_qubits, int m, double* cov_out,
                   cudaStream_t st) {
  const int d = 2 * n_qubits;
  GemmArgs g{X, X + m, cov_out, (long long)d * m, (long long)d * m, (long long)m * m, m, m, n_qubits, 2 * m, 2 * m, m, B};
  g.use_ksign = 1;
  g.kbits = bits;
  if (int rc = launch_bgemm(g, 1, 0, B, st)) return rc;
  const long long work = B * ((long long)m * (m + 1) / 2);
  long long blocks = (work + 255) / 256;
  if (blocks > 64LL * kNumSMs) blocks = 64LL * kNumSMs;
  antisymmetrize_kernel<<<(unsigned)blocks, 256, 0, st>>>(cov_out, B, m);
  return check_launch("antisymmetrize_kernel");
}

}  // namespace mcb200

using namespace mcb200;

extern "C" int mcb200_sptm_matmul(const double* A, int64_t strideA, int transA, const double* Bm, int64_t strideB,
                                  int transB, double* C, int64_t strideC, int32_t n, int64_t batch, void* stream) {
  if (n < 0 || batch < 0) return set_error("sptm_matmul: negative size");
  if (n == 0 || batch == 0) return 0;
  if (A == nullptr || Bm == nullptr || C == nullptr) return set_error("sptm_matmul: NULL pointer");
  GemmArgs g{A, Bm, C, strideA, strideB, strideC, n, n, n, n, n, n, batch};
  // TMA bulk copies need 16-byte aligned rows: even n, 16-byte aligned operands and batch strides
  const bool aligned = (n % 2 == 0) && (((uintptr_t)A | (uintptr_t)Bm | (uintptr_t)C) % 16 == 0) && (strideA % 2 == 0) &&
                       (strideB % 2 == 0) && (strideC % 2 == 0);
  if (n <= TMA_N && aligned) return launch_sptm_tma(g, transA, transB, as_stream(stream));
  return launch_bgemm(g, transA, transB, batch, as_stream(stream));
}

// Z[b] = G[b]^T - G[b]  (m x m)
__global__ void transpose_minus_kernel(const double* __restrict__ G, long long B, int m, double* __restrict__ Z) {
  const long long total = B * (long long)m * m;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / ((long long)m * m);
    const int r = (int)((e - b * (long long)m * m) / m), c = (int)(e % m);
    const double* g = G + b * (long long)m * m;
    Z[e] = g[c * m + r] - g[r * m + c];
  }
}

// grad_R[b][row][idx[c]] += Xbar[b][row][c] for row < d (the lifted row d of the extended covariance is constant)
__global__ void scatter_columns_add_kernel(const double* __restrict__ Xbar, long long B, int d, int D, const int32_t* idx,
                                           int m, double* grad_R, long long sR) {
  const long long total = B * (long long)d * m;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / ((long long)d * m);
    const int row = (int)((e - b * (long long)d * m) / m), c = (int)(e % m);
    const int col = idx ? idx[c] : c;
    if (col < d) atomicAdd(grad_R + b * sR + (long long)row * d + col, Xbar[b * (long long)D * m + (long long)row * m + c]);
  }
}

extern "C" int64_t mcb200_cov_dense_backward_workspace_bytes(int64_t B, int32_t D, int32_t m) {
  return (2 * B * (int64_t)D * m + B * (int64_t)m * m) * (int64_t)sizeof(double);
}

extern "C" int mcb200_cov_dense_backward(const double* R, int64_t strideR, const double* L0, int64_t strideL0, int64_t B,
                                         int32_t d, int32_t D, const int32_t* idx, int32_t m, const double* grad_cov,
                                         double* grad_R, void* workspace, void* stream) {
  if (B < 0 || d < 1 || m < 0) return set_error("cov_dense_backward: bad sizes");
  if (D != d && D != d + 1) return set_error("cov_dense_backward: D must be d or d+1 (got d=%d D=%d)", d, D);
  if (B == 0) return 0;
  if (grad_R == nullptr) return set_error("cov_dense_backward: grad_R is NULL");
  cudaStream_t st = as_stream(stream);
  const long long nR = (strideR == 0 ? 1 : B) * (long long)d * d;
  if (int rc = check_cuda(cudaMemsetAsync(grad_R, 0, (size_t)nR * sizeof(double), st), "cudaMemsetAsync")) return rc;
  if (m == 0) return 0;
  if (R == nullptr || L0 == nullptr || grad_cov == nullptr || workspace == nullptr)
    return set_error("cov_dense_backward: NULL pointer");
  double* X = (double*)workspace;              // (B, D, m)   R~[:, idx], later Xbar
  double* Y = X + B * (int64_t)D * m;          // (B, D, m)   L0 X
  double* Z = Y + B * (int64_t)D * m;          // (B, m, m)   G^T - G
  for (int64_t b0 = 0; b0 < B; b0 += 65535) {
    int64_t nb = B - b0 < 65535 ? B - b0 : 65535;
    dim3 grid((unsigned)((D * m + 255) / 256 < 64 ? (D * m + 255) / 256 : 64), (unsigned)nb);
    gather_columns_kernel<<<grid, 256, 0, st>>>(R + b0 * strideR, strideR, d, D, idx, m, X + b0 * (int64_t)D * m);
    if (int rc = check_launch("gather_columns_kernel")) return rc;
  }
  GemmArgs g1{L0, X, Y, strideL0, (long long)D * m, (long long)D * m, D, m, D, D, m, m, B};
  if (int rc = launch_bgemm(g1, 0, 0, B, st)) return rc;
  {
    const long long total = B * (long long)m * m;
    long long blocks = (total + 255) / 256;
    if (blocks > 64LL * kNumSMs) blocks = 64LL * kNumSMs;
    transpose_minus_kernel<<<(unsigned)blocks, 256, 0, st>>>(grad_cov, B, m, Z);
    if (int rc = check_launch("transpose_minus_kernel")) return rc;
  }
  // Xbar = Y (G^T - G)   (L0 is skew: L0^T X = -Y)
  GemmArgs g2{Y, Z, X, (long long)D * m, (long long)m * m, (long long)D * m, D, m, m, m, m, m, B};
  if (int rc = launch_bgemm(g2, 0, 0, B, st)) return rc;
  {
    const long long total = B * (long long)d * m;
    long long blocks = (total + 255) / 256;
    if (blocks > 64LL * kNumSMs) blocks = 64LL * kNumSMs;
    scatter_columns_add_kernel<<<(unsigned)blocks, 256, 0, st>>>(X, B, d, D, idx, m, grad_R, strideR);
    return check_launch("scatter_columns_add_kernel");
  }
}

extern "C" int64_t mcb200_cov_dense_workspace_bytes(int64_t B, int32_t D, int32_t m) {
  return 2 * B * (int64_t)D * m * (int64_t)sizeof(double);
}

extern "C" int mcb200_cov_dense(const double* R, int64_t strideR, const double* L0, int64_t strideL0, int64_t B,
                                int32_t d, int32_t D, const int32_t* idx, int32_t m, double* cov_out,
                                void* workspace, void* stream) {
  if (B < 0 || d < 1 || m < 0) return set_error("cov_dense: bad sizes");
  if (D != d && D != d + 1) return set_error("cov_dense: D must be d or d+1 (got d=%d D=%d)", d, D);
  if (B == 0 || m == 0) return 0;
  if (R == nullptr || L0 == nullptr || cov_out == nullptr || workspace == nullptr)
    return set_error("cov_dense: NULL pointer");
  cudaStream_t st = as_stream(stream);
  double* X = (double*)workspace;
  double* Y = X + B * (int64_t)D * m;
  for (int64_t b0 = 0; b0 < B; b0 += 65535) {
    int64_t nb = B - b0 < 65535 ? B - b0 : 65535;
    dim3 grid((unsigned)((D * m + 255) / 256 < 64 ? (D * m + 255) / 256 : 64), (unsigned)nb);
    gather_columns_kernel<<<grid, 256, 0, st>>>(R + b0 * strideR, strideR, d, D, idx, m, X + b0 * (int64_t)D * m);
    if (int rc = check_launch("gather_columns_kernel")) return rc;
  }
  // Y = L0 @ X  (D x D @ D x m)
  GemmArgs g1{L0, X, Y, strideL0, (long long)D * m, (long long)D * m, D, m, D, D, m, m, B};
  if (int rc = launch_bgemm(g1, 0, 0, B, st)) return rc;
  // cov = X^T @ Y  (m x D @ D x m)
  GemmArgs g2{X, Y, cov_out, (long long)D * m, (long long)D * m, (long long)m * m, m, m, D, m, m, m, B};
  return launch_bgemm(g2, 1, 0, B, st);
}
