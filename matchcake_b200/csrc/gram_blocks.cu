// Gram builder for circuits whose wires split into independent 2-qubit groups (block-diagonal SPTM).
//
// If no gate ever couples two groups of wires, every per-sample covariance Lambda_i = R_i^T Lambda_0 R_i is block
// diagonal with the same 4x4 blocks pattern for all samples, and
//     K_ij = 2^-N |Pf(Lambda_i + Lambda_j)| = prod_c  2^-2 |Pf_4(Lambda_i^(c) + Lambda_j^(c))|,
//     Pf_4(S) = s01 s23 - s02 s13 + s03 s12.
// This is exactly what FermionicPQCKernel builds when n_features <= n_qubits (depth_ = 1: rotations and fSWAPs act
// on the same pairs, fermionic_pqc_kernel.py:183-192 -- config C5, 128 qubits = 64 blocks).  The algorithmic work
// per Gram entry drops from d^3/3 = 5.6e6 flop to 64 x ~16 flop, and the per-sample state from 512 KB to 3 KB.
// A block computes 64 x 64 cells, 8 components per shared-memory stage, on the FP64 tensor pipe (see the kernel).
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
  double floor2;
};

constexpr int GB_TM = 128;          // row samples per block tile
constexpr int GB_TN = 64;           // column samples per block tile
constexpr int GB_CC = 8;            // blocks (components) per shared-memory stage
constexpr int GB_LD = GB_CC * 8 + 4;  // doubles per staged sample: 8 per component, +4 so that the 8-byte fragment reads
                                      // of a half warp (4 rows x 4 consecutive doubles) fall on 16 distinct bank pairs

__device__ __forceinline__ void gb_dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// Pf_4(r + q) is bilinear in the two blocks up to their own Pfaffians:
//   Pf_4(r + q) = Pf_4(r) + Pf_4(q) + r01 q23 + r23 q01 - r02 q13 - r13 q02 + r03 q12 + r12 q03
//               = < (r01, r02, r03, r12, r13, r23, Pf_4(r), 1), (q23, -q13, q12, q03, -q02, q01, 1, Pf_4(q)) >,
// an inner product of length 8.  For one component the 64 x 64 cells of a tile are therefore a (64 x 8)(8 x 64) matrix
// product: two K = 4 DMMAs per 8 x 8 cells instead of ten FP64 instructions per cell, and the operands are read as MMA
// fragments (16 eight-byte loads per warp, component and 1024 cells instead of 48 sixteen-byte ones).  The running
// product over the components is kept in the accumulator layout.  Block tile 128 x 64 cells, 8 warps as 4 (rows) x 2
// (columns), 32 x 32 cells each.
__global__ void __launch_bounds__(256, 2) gram_blocks4_kernel(GramBlocksArgs g) {
  extern __shared__ __align__(16) double gb_smem[];
  double* As = gb_smem;                   // GB_TM row samples    x GB_LD
  double* Bs = gb_smem + GB_TM * GB_LD;   // GB_TN column samples x GB_LD
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2, gr = lane >> 2, tg = lane & 3;
  const long long rows = g.row_end - g.row_begin;
  const long long rt = (rows + GB_TM - 1) / GB_TM, ct = (g.n1 + GB_TN - 1) / GB_TN;
  const bool upper = g.n0 < g.n1;
  for (long long t = blockIdx.x; t < rt * ct; t += gridDim.x) {
    const long long tr = t / ct, tc = t - tr * ct;
    const long long i0 = g.row_begin + tr * GB_TM, j0 = tc * GB_TN;
    if (g.mode != MCB200_GRAM_FULL) {  // tile without any evaluated cell (gram_matrix.py:191-195)
      if (!upper && j0 >= i0 + GB_TM) continue;
      if (upper && j0 + GB_TN - 1 <= i0) continue;
    }
    double acc[4][4][2], pend[4][2];  // running products; results of the unit whose multiply is still outstanding
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 1.0;
#pragma unroll
    for (int b = 0; b < 4; ++b) pend[b][0] = pend[b][1] = 1.0;
    for (int c0 = 0; c0 < g.C; c0 += GB_CC) {
      const int cc = min(GB_CC, g.C - c0);
      __syncthreads();
      // stage: 192 samples x cc components, 6 (sample, component) items per thread; the 6 block entries
      // (x01, x02 | x03, x12 | x13, x23) become the 8-vector of the row side or of the column side.  (A register
      // prefetch of the next stage was measured: it costs the third resident block and is slower, 55 vs 52 ms on C5.)
      for (int it = tid; it < (GB_TM + GB_TN) * GB_CC; it += 256) {
        const int smp = it / GB_CC, c = it - smp * GB_CC;
        const bool is_row = smp < GB_TM;
        const long long idx = is_row ? i0 + smp : j0 + (smp - GB_TM);
        const bool live = c < cc && (is_row ? idx < g.row_end : idx < g.n1);
        double2 e0 = make_double2(0.0, 0.0), e1 = e0, e2 = e0;
        if (live) {
          const double2* src = reinterpret_cast<const double2*>((is_row ? g.b0 : g.b1) + (idx * g.C + c0 + c) * 6);
          e0 = __ldg(src);
          e1 = __ldg(src + 1);
          e2 = __ldg(src + 2);
        }
        const double pf = fma(e1.x, e1.y, fma(-e0.y, e2.x, e0.x * e2.y));
        double2* dst = reinterpret_cast<double2*>((is_row ? As + smp * GB_LD : Bs + (smp - GB_TM) * GB_LD) + c * 8);
        if (is_row) {
          dst[0] = e0;
          dst[1] = e1;
          dst[2] = e2;
          dst[3] = make_double2(pf, live ? 1.0 : 0.0);
        } else {
          dst[0] = make_double2(e2.y, -e2.x);
          dst[1] = make_double2(e1.y, e1.x);
          dst[2] = make_double2(-e0.y, e0.x);
          dst[3] = make_double2(live ? 1.0 : 0.0, pf);
        }
      }
      __syncthreads();
      for (int c = 0; c < cc; ++c) {
        double af[4][2], bf[4][2];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) af[a][ks] = As[(wm * 32 + a * 8 + gr) * GB_LD + c * 8 + ks * 4 + tg];
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) bf[b][ks] = Bs[(wn * 32 + b * 8 + gr) * GB_LD + c * 8 + ks * 4 + tg];
        // Software pipeline over the four 8-row units of the patch: the DMMAs of unit a are issued before the
        // running products are multiplied by the results of unit a - 1 (unit 3 of the previous component for a = 0),
        // so no multiply waits on a tensor instruction that has just been issued.
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          double p[4][2];
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            p[b][0] = p[b][1] = 0.0;
            gb_dmma(p[b][0], p[b][1], af[a][0], bf[b][0]);
          }
#pragma unroll
          for (int b = 0; b < 4; ++b) gb_dmma(p[b][0], p[b][1], af[a][1], bf[b][1]);
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            acc[(a + 3) & 3][b][0] *= fabs(pend[b][0]);
            acc[(a + 3) & 3][b][1] *= fabs(pend[b][1]);
            pend[b][0] = p[b][0];
            pend[b][1] = p[b][1];
          }
        }
      }
    }
    // drain the pipeline: the last unit (a = 3 of the last component) is still pending
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      acc[3][b][0] *= fabs(pend[b][0]);
      acc[3][b][1] *= fabs(pend[b][1]);
    }
    // accumulator layout: lane (gr, tg) holds cells (row gr, columns 2 tg, 2 tg + 1) of every 8 x 8 tile
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const long long i = i0 + wm * 32 + a * 8 + gr;
      if (i >= g.row_end) continue;
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const long long j = j0 + wn * 32 + b * 8 + 2 * tg + e;
          if (j >= g.n1) continue;
          if (g.mode != MCB200_GRAM_FULL && !(upper ? (j > i) : (i > j))) continue;
          const double v = abs_floor(g.scale * acc[a][b][e], g.floor2);
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
  GramBlocksArgs g{blocks0, blocks1, n0, n1, C, row_begin, row_end, mode, scale, K, ldk, g_abs_floor2};
  const long long tiles = ((rows + GB_TM - 1) / GB_TM) * ((n1 + GB_TN - 1) / GB_TN);
  long long blocks = tiles < 8LL * kNumSMs ? tiles : 8LL * kNumSMs;
  const size_t smem = (size_t)(GB_TM + GB_TN) * GB_LD * sizeof(double);
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
