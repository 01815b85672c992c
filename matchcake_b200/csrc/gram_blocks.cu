// Gram builder for circuits whose wires split into independent 2-qubit groups (block-diagonal SPTM).
//
// If no gate ever couples two groups of wires, every per-sample covariance Lambda_i = R_i^T Lambda_0 R_i is block
// diagonal with the same 4x4 blocks pattern for all samples, and
//     K_ij = 2^-N |Pf(Lambda_i + Lambda_j)| = prod_c  2^-2 |Pf_4(Lambda_i^(c) + Lambda_j^(c))|,
//     Pf_4(S) = s01 s23 - s02 s13 + s03 s12.
// This is exactly what FermionicPQCKernel builds when n_features <= n_qubits (depth_ = 1: rotations and fSWAPs act
// on the same pairs, fermionic_pqc_kernel.py:183-192 -- config C5, 128 qubits = 64 blocks).  The algorithmic work
// per Gram entry drops from d^3/3 = 5.6e6 flop to 64 x ~16 flop, and the per-sample state from 512 KB to 3 KB.
// One thread computes 4 x 4 cells (rows ty + 16a, columns tx + 16b); a block stages 64 + 64 samples x 8 blocks through shared memory.
#include "common.cuh"

namespace mcb200 {

struct GramBlocksArgs {
  const double* b0;  // (n0, C, 6): upper triangle (01, 02, 03, 12, 13, 23) of every 4x4 diagonal block
  const double* b1;  // (n1, C, 6)
  long long n0, n1;
  int C;
  long long row_begin, row_end;
  int mode;
  double scale;      // global factor, normally 2^-(2C) = 2^-N
  double* K;
  long long ldk;
};

constexpr int GB_T = 64;   // cells per tile side
constexpr int GB_CC = 8;   // blocks (components) per shared-memory stage
constexpr int GB_LD = GB_CC * 6 + 2;  // sample stride 400 B = 16 mod 128: the 8 lanes of a 128-bit phase hit 8 different 16-byte slots

__global__ void __launch_bounds__(256) gram_blocks4_kernel(GramBlocksArgs g) {
  extern __shared__ __align__(16) double gb_smem[];
  double* sr = gb_smem;
  double* sc = gb_smem + GB_T * GB_LD;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long rows = g.row_end - g.row_begin;
  const long long rt = (rows + GB_T - 1) / GB_T, ct = (g.n1 + GB_T - 1) / GB_T;
  const bool upper = g.n0 < g.n1;
  for (long long t = blockIdx.x; t < rt * ct; t += gridDim.x) {
    const long long tr = t / ct, tc = t - tr * ct;
    const long long i0 = g.row_begin + tr * GB_T, j0 = tc * GB_T;
    if (g.mode != MCB200_GRAM_FULL) {  // tile without any evaluated cell (gram_matrix.py:191-195)
      if (!upper && j0 >= i0 + GB_T) continue;
      if (upper && j0 + GB_T - 1 <= i0) continue;
    }
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 1.0;
    for (int c0 = 0; c0 < g.C; c0 += GB_CC) {
      const int cc = min(GB_CC, g.C - c0);
      __syncthreads();
      {  // 64 samples x (cc * 6) doubles per operand: 4 threads per sample, 2 * cc... contiguous doubles each
        const int s = tid >> 2, seg = tid & 3;
        const int per = (cc * 6 + 3) >> 2;  // doubles per thread (12 for a full stage)
        const int w0 = seg * per, w1 = min(w0 + per, cc * 6);
        const long long i = i0 + s, j = j0 + s;
        const double* pr = g.b0 + (i * g.C + c0) * 6;
        const double* pc = g.b1 + (j * g.C + c0) * 6;
        for (int w = w0; w < w1; ++w) {
          sr[s * GB_LD + w] = (i < g.row_end) ? __ldg(pr + w) : 0.0;
          sc[s * GB_LD + w] = (j < g.n1) ? __ldg(pc + w) : 0.0;
        }
      }
      __syncthreads();
      for (int c = 0; c < cc; ++c) {
        double r[4][6], q[4][6];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          // rows: one address per 16 lanes (broadcast); columns: consecutive samples (conflict free)
          const double2* pr = reinterpret_cast<const double2*>(sr + (ty + 16 * a) * GB_LD + c * 6);
          const double2* pq = reinterpret_cast<const double2*>(sc + (tx + 16 * a) * GB_LD + c * 6);
#pragma unroll
          for (int w = 0; w < 3; ++w) {
            const double2 u = pr[w], v = pq[w];
            r[a][2 * w] = u.x;
            r[a][2 * w + 1] = u.y;
            q[a][2 * w] = v.x;
            q[a][2 * w + 1] = v.y;
          }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const double s01 = r[a][0] + q[b][0], s02 = r[a][1] + q[b][1], s03 = r[a][2] + q[b][2];
            const double s12 = r[a][3] + q[b][3], s13 = r[a][4] + q[b][4], s23 = r[a][5] + q[b][5];
            const double pf = fma(s03, s12, fma(-s02, s13, s01 * s23));
            acc[a][b] *= fabs(pf);
          }
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const long long i = i0 + ty + 16 * a;
      if (i >= g.row_end) continue;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const long long j = j0 + tx + 16 * b;
        if (j >= g.n1) continue;
        if (g.mode != MCB200_GRAM_FULL && !(upper ? (j > i) : (i > j))) continue;
        const double v = g.scale * acc[a][b];
        g.K[i * g.ldk + j] = v;
        if (g.mode == MCB200_GRAM_REFERENCE && j < g.n0 && i < g.n1) g.K[j * g.ldk + i] = v;  // gram_matrix.py:139-143
      }
    }
  }
}

__global__ void gram_blocks_diag_kernel(double* K, long long ldk, long long row_begin, long long row_end, long long n1) {
  long long i = row_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < row_end && i < n1) K[i * ldk + i] = 1.0;  // gram_matrix.py:147-161
}

}  // namespace mcb200

using namespace mcb200;

extern "C" int mcb200_gram_blocks4(const double* blocks0, const double* blocks1, int64_t n0, int64_t n1, int32_t C,
                                   int64_t row_begin, int64_t row_end, int32_t mode, double scale, double* K,
                                   int64_t ldk, void* stream) {
  if (n0 < 0 || n1 < 0 || C < 1) return set_error("gram_blocks4: bad sizes");
  if (row_begin < 0 || row_end > n0 || row_begin > row_end) return set_error("gram_blocks4: bad row range");
  if (mode != MCB200_GRAM_REFERENCE && mode != MCB200_GRAM_FULL && mode != MCB200_GRAM_TRIANGLE)
    return set_error("gram_blocks4: unknown mode %d", mode);
  if (ldk < n1) return set_error("gram_blocks4: ldk < n1");
  const long long rows = row_end - row_begin;
  if (rows == 0 || n1 == 0) return 0;
  if (blocks0 == nullptr || blocks1 == nullptr || K == nullptr) return set_error("gram_blocks4: NULL pointer");
  cudaStream_t st = as_stream(stream);
  GramBlocksArgs g{blocks0, blocks1, n0, n1, C, row_begin, row_end, mode, scale, K, ldk};
  const long long tiles = ((rows + GB_T - 1) / GB_T) * ((n1 + GB_T - 1) / GB_T);
  long long blocks = tiles < 8LL * kNumSMs ? tiles : 8LL * kNumSMs;
  const size_t smem = 2 * (size_t)GB_T * GB_LD * sizeof(double);
  if (int rc = check_cuda(cudaFuncSetAttribute(gram_blocks4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                          "cudaFuncSetAttribute(gram_blocks4)"))
    return rc;
  gram_blocks4_kernel<<<(unsigned)blocks, 256, smem, st>>>(g);
  if (int rc = check_launch("gram_blocks4_kernel")) return rc;
  if (mode == MCB200_GRAM_REFERENCE) {
    gram_blocks_diag_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(K, ldk, row_begin, row_end, n1);
    return check_launch("gram_blocks_diag_kernel");
  }
  return 0;
}
