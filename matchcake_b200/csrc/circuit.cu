// K1 (structured SPTM chain) + K2 (basis-state covariance conjugation) + fused projector expval.
//
// One thread block per circuit instance.  The block keeps X = R[:, cols] (d x m, FP64) in shared memory and
// applies the gate list in REVERSE application order from the left:
//     R = G_0 G_1 ... G_{n-1}   =>   R[:, cols] = G_0 (G_1 ( ... (G_{n-1} I[:, cols])))
// Every gate is a 4x4 block on rows 2*wire0 .. 2*wire0+3, so one gate costs 8-16 FMA per column instead of
// the 2 d^2 per column of the reference's dense padded product (devices/nif_device.py:195,
// single_particle_transition_matrix.py:379-402).  Gates of one layer act on disjoint rows and run in parallel.
#include <stdlib.h>

#include "common.cuh"
#include "circuit_dev.cuh"
#include "pfaffian_dev.cuh"
#include "tiles_api.cuh"

namespace mcb200 {

// Rotation gate on nk columns cc = first + step * j, visited in the rotated order j = rot, rot + 1, ... (mod nk): two
// Givens pairs (i1, j1), (i2, j2) of the 4-row block; the pair layout is the only thing that depends on the kind
// (RXRX (0,3)(1,2), RYRY (0,2)(1,3), RZRZ (0,1)(2,3)).  The rotation staggers the columns touched by neighbouring
// gates at the same time, which is what keeps lane = gate accesses off the same shared-memory banks.
template <bool UNIT>
__device__ __forceinline__ void apply_rotation_columns(int kind, int wire0, double2 a, double2 b, double* X, int ld,
                                                       int first, int step, int nk, int rot) {
  const int j1 = 3 - kind, i2 = (kind == MCB200_GATE_RZRZ) ? 2 : 1, j2 = (kind == MCB200_GATE_RXRX) ? 2 : 3;
  double* x0 = X + (2 * wire0) * ld + first;
  double* x1 = x0 + j1 * ld;
  double* x2 = x0 + i2 * ld;
  double* x3 = x0 + j2 * ld;
  int j = rot;
#pragma unroll 4
  for (int k = 0; k < nk; ++k) {
    const int cc = UNIT ? j : j * step;
    const double p = x0[cc], q = x1[cc], r = x2[cc], t = x3[cc];
    x0[cc] = a.x * p + a.y * q;
    x1[cc] = a.x * q - a.y * p;
    x2[cc] = b.x * r + b.y * t;
    x3[cc] = b.x * t - b.y * r;
    if (++j == nk) j = 0;
  }
}

// Build X = R[:, cols[c0 .. c0+mc)] in shared memory (d x mc, leading dimension ld).  coef holds the rotation
// coefficients (4 doubles per gate) of up to coef_cap gates: whole layers are batched so that the sincos phase keeps
// every thread busy and costs one block barrier per batch.  Columns never mix, so every WARP owns a fixed subset of
// the columns (cc = warp, warp + n_warps, ...) and walks the layers on its own: a layer boundary is a __syncwarp, not
// a block barrier.  Inside a warp the lanes are spread over (gate, column sub-group) items.  All threads participate.
__device__ void build_columns(const CircuitDev& c, const double* prow, const int32_t* cols, int c0, int mc,
                              double* X, int ld, double* coef, int coef_cap) {
  const int d = 2 * c.n_qubits;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int lane = tid & 31, w = tid >> 5, nw = nthr >> 5;
  const int shift = ((mc & (mc - 1)) == 0) ? (31 - __clz(mc)) : -1;  // mc power of two: index split by shift
  for (int idx = tid; idx < d * mc; idx += nthr) {
    int r = shift >= 0 ? (idx >> shift) : idx / mc, cc = idx - r * mc;
    int col = cols ? cols[c0 + cc] : c0 + cc;
    X[r * ld + cc] = (r == col) ? 1.0 : 0.0;
  }
  // warp w owns the contiguous columns [base, base + wc)
  const int cpw = (mc + nw - 1) / nw;
  const int base = w * cpw;
  const int wc = min(cpw, mc - base);  // <= 0: none
  int l = c.n_layers - 1;
  while (l >= 0) {
    const int ge = c.layer_ptr[l + 1];
    int lb = l;
    while (lb > 0 && ge - c.layer_ptr[lb - 1] <= coef_cap) --lb;
    const int gb = c.layer_ptr[lb];
    __syncthreads();  // X initialised / every warp is done with the previous batch's coefficients
    // coefficients of the batch, structure of arrays: (c1, t1) of gate k at coefA[k], (c2, t2) at coefB[k]
    double2* coefA = reinterpret_cast<double2*>(coef);
    double2* coefB = coefA + coef_cap;
    for (int g = gb + tid; g < ge; g += nthr) {
      const mcb200_gate_t gt = c.gates[g];
      // (a layer of more than coef_cap gates -- impossible for gates on disjoint wires, possible in a malformed
      // descriptor handed over the C ABI -- keeps its surplus gates out of the buffer: they are applied below with
      // coefficients computed on the fly)
      if (gt.kind <= MCB200_GATE_RZRZ && g - gb < coef_cap) {
        double cf[4];
        gate_coefficients(c, gt, prow, cf);
        coefA[g - gb] = make_double2(cf[0], cf[1]);
        coefB[g - gb] = make_double2(cf[2], cf[3]);
      }
    }
    __syncthreads();
    if (wc > 0) {
      const bool wc_pow2 = (wc & (wc - 1)) == 0;
      for (int ll = l; ll >= lb; --ll) {
        const int g0 = c.layer_ptr[ll], ng = c.layer_ptr[ll + 1] - g0;
        if (ng > 16 || wc == 1) {
          // lane = gate, all of the warp's columns; neighbouring gates start at staggered columns (bank spread)
          for (int g = lane; g < ng; g += 32) {
            const mcb200_gate_t gt = c.gates[g0 + g];
            const int k = g0 + g - gb;
            if (gt.kind <= MCB200_GATE_RZRZ && k < coef_cap) {
              apply_rotation_columns<true>(gt.kind, gt.wire0, coefA[k], coefB[k], X, ld, base, 1, wc,
                                           wc_pow2 ? ((g >> 2) & (wc - 1)) : 0);
            } else if (gt.kind <= MCB200_GATE_RZRZ) {
              double cf[4];
              gate_coefficients(c, gt, prow, cf);
              apply_gate_columns(c, gt, cf, prow, X, ld, base, 1, base + wc);
            } else {
              apply_gate_columns(c, gt, nullptr, prow, X, ld, base, 1, base + wc);
            }
          }
        } else {
          // few gates per layer: S column sub-groups per gate so that ng * S items keep the lanes busy
          int S = ng > 8 ? 2 : (ng > 4 ? 4 : 8);
          if (S > wc) S = wc;
          for (int it = lane; it < ng * S; it += 32) {
            const int g = it / S, sub = it - g * S;
            const mcb200_gate_t gt = c.gates[g0 + g];
            if (gt.kind <= MCB200_GATE_RZRZ) {
              const int k = g0 + g - gb;
              apply_rotation_columns<false>(gt.kind, gt.wire0, coefA[k], coefB[k], X, ld, base + sub, S,
                                            (wc - sub + S - 1) / S, 0);
            } else {
              apply_gate_columns(c, gt, nullptr, prow, X, ld, base + sub, S, base + wc);
            }
          }
        }
        __syncwarp();
      }
    }
    l = lb - 1;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------- kernels
__global__ void circuit_columns_kernel(CircuitDev c, const double* __restrict__ params, const int32_t* cols,
                                       int m, int mc, int max_layer_gates, double* __restrict__ X_out) {
  extern __shared__ double smem[];
  const int d = 2 * c.n_qubits;
  const long long b = blockIdx.x;
  const int c0 = blockIdx.y * mc;
  const int mcur = min(mc, m - c0);
  const int ld = mc | 1;  // odd leading dimension: the 4-row blocks of different gates fall into different banks
  double* X = smem;
  double* coef = X + (size_t)d * ld;
  build_columns(c, params + b * c.n_param_cols, cols, c0, mcur, X, ld, coef, max_layer_gates);
  double* out = X_out + b * (long long)d * m;
  for (int idx = threadIdx.x; idx < d * mcur; idx += blockDim.x) {
    int r = idx / mcur, cc = idx - r * mcur;
    out[(long long)r * m + c0 + cc] = X[r * ld + cc];
  }
}

// cov[a][b] = sum_k s_k (X[2k][a] X[2k+1][b] - X[2k+1][a] X[2k][b]),  s_k = -z_k = -(-1)^{bits[k]}
// (product_state.py:203-205 + nif_device.py:1131).  Strict upper triangle a < b only, computed as 2x2 tiles.
__device__ void cov_basis_upper(const double* X, int ldx, int n_qubits, int m, const int8_t* bits, double* A,
                                int lda) {
  const int tiles = (m + 1) / 2;
  const int items = tiles * tiles;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    int ta = it / tiles, tb = it - ta * tiles;
    if (tb < ta) continue;
    int a0 = 2 * ta, b0 = 2 * tb;
    int a1 = min(a0 + 1, m - 1), b1 = min(b0 + 1, m - 1);
    double s00 = 0, s01 = 0, s10 = 0, s11 = 0;
    for (int k = 0; k < n_qubits; ++k) {
      double sg = (bits != nullptr && bits[k]) ? 1.0 : -1.0;
      const double* xe = X + (2 * k) * ldx;
      const double* xo = xe + ldx;
      double ea0 = sg * xe[a0], ea1 = sg * xe[a1], oa0 = sg * xo[a0], oa1 = sg * xo[a1];
      double eb0 = xe[b0], eb1 = xe[b1], ob0 = xo[b0], ob1 = xo[b1];
      s00 += ea0 * ob0 - oa0 * eb0;
      s01 += ea0 * ob1 - oa0 * eb1;
      s10 += ea1 * ob0 - oa1 * eb0;
      s11 += ea1 * ob1 - oa1 * eb1;
    }
    if (a0 < b0) A[a0 * lda + b0] = s00;
    if (a0 < b0 + 1 && b0 + 1 < m) A[a0 * lda + b0 + 1] = s01;
    if (a0 + 1 < b0 && a0 + 1 < m) A[(a0 + 1) * lda + b0] = s10;
    if (a0 + 1 < b0 + 1 && a0 + 1 < m && b0 + 1 < m) A[(a0 + 1) * lda + b0 + 1] = s11;
  }
}

// K2 standalone: X (B, d, m) in global memory -> cov (B, m, m) full skew-symmetric matrix.
__global__ void cov_basis_kernel(const double* __restrict__ X, const int8_t* bits, int n_qubits, int m,
                                 double* __restrict__ cov_out) {
  extern __shared__ double smem[];
  const int d = 2 * n_qubits;
  const long long b = blockIdx.x;
  double* Xs = smem;            // d x m
  double* A = Xs + (size_t)d * m;  // m x (m+1)
  const int lda = m + 1;
  const double* src = X + b * (long long)d * m;
  for (int idx = threadIdx.x; idx < d * m; idx += blockDim.x) Xs[idx] = src[idx];
  __syncthreads();
  cov_basis_upper(Xs, m, n_qubits, m, bits, A, lda);
  __syncthreads();
  double* out = cov_out + b * (long long)m * m;
  for (int idx = threadIdx.x; idx < m * m; idx += blockDim.x) {
    int i = idx / m, j = idx - i * m;
    out[idx] = (i < j) ? A[i * lda + j] : (i > j ? -A[j * lda + i] : 0.0);
  }
}

// Fused: circuit -> X (d x d) -> Lambda = X^T Lambda_0 X -> + Lambda_y -> 2^-N |Pf|.   One block per instance.
__global__ void circuit_expval_projector_kernel(CircuitDev c, const double* __restrict__ params,
                                                const int8_t* init_bits, const int8_t* target_bits,
                                                double floor2, double* __restrict__ out) {
  extern __shared__ double smem[];
  const int n = c.n_qubits, d = 2 * n;
  const long long b = blockIdx.x;
  double* X = smem;                    // d x d
  double* A = X + (size_t)d * d;       // d x (d+1)
  double* coef = A + (size_t)d * (d + 1);
  double* red = coef;                  // reused after the circuit phase
  const int lda = d + 1;
  build_columns(c, params + b * c.n_param_cols, nullptr, 0, d, X, d, coef, n > 16 ? n : 16);
  cov_basis_upper(X, d, n, d, init_bits, A, lda);
  __syncthreads();
  // + Lambda_y: (Lambda_y)[2j][2j+1] = -(-1)^{y_j}   (product_state_strategy.py:89-91)
  for (int j = threadIdx.x; j < n; j += blockDim.x)
    A[(2 * j) * lda + 2 * j + 1] += (target_bits != nullptr && target_bits[j]) ? 1.0 : -1.0;
  __syncthreads();
  double pf = block_pfaffian_upper(A, d, lda, red);
  if (threadIdx.x == 0) out[b] = abs_floor(ldexp(pf, -n), floor2);
}

// Block-diagonal circuits: one THREAD per (instance, wire pair).  The pair's 4x4 SPTM block is built in registers
// from the gates that act on it (reverse application order, left multiplication), then its covariance block
// Lambda^(c) = X^T Lambda_0^(c) X is written as the 6 upper-triangular entries (01,02,03,12,13,23).
__global__ void __launch_bounds__(256) circuit_pair_blocks_kernel(CircuitDev c, const double* __restrict__ params,
                                                                  long long B, const int32_t* __restrict__ pair_wire0,
                                                                  int C, const int8_t* init_bits,
                                                                  double* __restrict__ out) {
  extern __shared__ mcb200_gate_t sgates[];
  for (int i = threadIdx.x; i < c.n_gates * 6; i += blockDim.x)
    reinterpret_cast<int32_t*>(sgates)[i] = reinterpret_cast<const int32_t*>(c.gates)[i];
  __syncthreads();
  const long long total = B * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / C;
    const int pc = (int)(e - b * C);
    const int w0 = pair_wire0[pc];
    const double* prow = params + b * c.n_param_cols;
    double x[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) x[i][j] = (i == j) ? 1.0 : 0.0;
    for (int gi = c.n_gates - 1; gi >= 0; --gi) {
      const mcb200_gate_t g = sgates[gi];
      if (g.wire0 != w0) continue;
      if (g.kind <= MCB200_GATE_RZRZ) {
        double cf[4];
        gate_coefficients(c, g, prow, cf);
        // row pairs (i0, i1), (j0, j1): RXRX (0,3),(1,2); RYRY (0,2),(1,3); RZRZ (0,1),(2,3)
#pragma unroll
        for (int col = 0; col < 4; ++col) {
          const double x0 = x[0][col], x1 = x[1][col], x2 = x[2][col], x3 = x[3][col];
          if (g.kind == MCB200_GATE_RXRX) {
            x[0][col] = cf[0] * x0 + cf[1] * x3;  x[3][col] = cf[0] * x3 - cf[1] * x0;
            x[1][col] = cf[2] * x1 + cf[3] * x2;  x[2][col] = cf[2] * x2 - cf[3] * x1;
          } else if (g.kind == MCB200_GATE_RYRY) {
            x[0][col] = cf[0] * x0 + cf[1] * x2;  x[2][col] = cf[0] * x2 - cf[1] * x0;
            x[1][col] = cf[2] * x1 + cf[3] * x3;  x[3][col] = cf[2] * x3 - cf[3] * x1;
          } else {
            x[0][col] = cf[0] * x0 + cf[1] * x1;  x[1][col] = cf[0] * x1 - cf[1] * x0;
            x[2][col] = cf[2] * x2 + cf[3] * x3;  x[3][col] = cf[2] * x3 - cf[3] * x2;
          }
        }
      } else if (g.kind == MCB200_GATE_FSWAP) {  // adjacent fSWAP: rows (0,1) <-> (2,3)
#pragma unroll
        for (int col = 0; col < 4; ++col) {
          const double x0 = x[0][col], x1 = x[1][col];
          x[0][col] = x[2][col];  x[1][col] = x[3][col];  x[2][col] = x0;  x[3][col] = x1;
        }
      } else {
        const double* m4 = (g.kind == MCB200_GATE_BLOCK4_CONST) ? c.consts + g.off : prow + g.off;
        const bool tr = g.flags & MCB200_GATE_FLAG_TRANSPOSE;
#pragma unroll
        for (int col = 0; col < 4; ++col) {
          const double x0 = x[0][col], x1 = x[1][col], x2 = x[2][col], x3 = x[3][col];
#pragma unroll
          for (int a = 0; a < 4; ++a)
            x[a][col] = (tr ? m4[a] : m4[4 * a]) * x0 + (tr ? m4[4 + a] : m4[4 * a + 1]) * x1 +
                        (tr ? m4[8 + a] : m4[4 * a + 2]) * x2 + (tr ? m4[12 + a] : m4[4 * a + 3]) * x3;
        }
      }
    }
    const double s0 = (init_bits != nullptr && init_bits[w0]) ? 1.0 : -1.0;
    const double s1 = (init_bits != nullptr && init_bits[w0 + 1]) ? 1.0 : -1.0;
    double* o = out + e * 6;
    int q = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b2 = a + 1; b2 < 4; ++b2)
        o[q++] = s0 * (x[0][a] * x[1][b2] - x[1][a] * x[0][b2]) + s1 * (x[2][a] * x[3][b2] - x[3][a] * x[2][b2]);
  }
}

}  // namespace mcb200

// ---------------------------------------------------------------------------------------------- C ABI
using namespace mcb200;

extern "C" int mcb200_circuit_columns(const mcb200_circuit_t* circuit_host, const double* params, int64_t B,
                                      const int32_t* cols, int32_t m, double* X_out, void* stream) {
  CircuitDev c;
  if (int rc = circuit_to_dev(circuit_host, &c)) return rc;
  if (B < 0 || m < 0) return set_error("negative batch or column count");
  if (B == 0 || m == 0) return 0;
  if (X_out == nullptr) return set_error("X_out is NULL");
  if (c.n_param_cols > 0 && params == nullptr) return set_error("params is NULL");
  const int d = 2 * c.n_qubits;
  // narrow panels (marginals over <= 4 wires) of nearest-neighbour circuits: register-resident kernel; wider panels
  // stay on the shared-memory kernel, which shares the trigonometry between the columns (see circuit_brick.cu)
  if (c.nearest_neighbour && c.n_qubits <= 64 && m <= 8 && c.n_gates > 0) {
    const int rc = circuit_columns_brick_launch(c, params, B, cols, m, X_out, as_stream(stream));
    if (rc >= 0) return rc;
  }
  // (the column-major panel kernel of circuit_panel.cu is used by mcb200_circuit_cov_basis only: for X itself it measures
  // 1.60 ms against 1.37 ms of the row-major kernel below on config C4's pruned gate list, MCB200_PANEL_K1=1 selects it)
  static const bool panel_k1 = [] {
    const char* e = getenv("MCB200_PANEL_K1");
    return e != nullptr && e[0] == '1';
  }();
  if (panel_k1 && c.nearest_neighbour && m > 8 && m <= 32 && cols != nullptr) {
    const int rc = circuit_panel_launch(c, params, B, cols, m, nullptr, X_out, nullptr, as_stream(stream));
    if (rc >= 0) return rc;
  }
  // coefficient buffer: at least one full layer (<= N gates on disjoint wires), normally 256 gates per batch
  const int max_layer_gates = c.n_qubits > 256 ? c.n_qubits : 256;
  // column chunk so that X (d x mc) + coefficients fit in shared memory
  const size_t coef_bytes = (size_t)max_layer_gates * 4 * sizeof(double);
  int mc = m;
  while ((size_t)d * (mc | 1) * sizeof(double) + coef_bytes > 160 * 1024 && mc > 1) mc = (mc + 1) / 2;
  const size_t smem = (size_t)d * (mc | 1) * sizeof(double) + coef_bytes;
  if (smem > (size_t)kMaxDynSmem) return set_error("circuit_columns: d=%d does not fit shared memory", d);
  if (int rc = check_cuda(cudaFuncSetAttribute(circuit_columns_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)smem), "cudaFuncSetAttribute(circuit_columns)"))
    return rc;
  const int chunks = (m + mc - 1) / mc;
  const int threads = (d * mc >= 4096) ? 256 : 128;
  for (int64_t b0 = 0; b0 < B; b0 += 2147483647LL) {  // gridDim.x limit
    int64_t nb = B - b0 < 2147483647LL ? B - b0 : 2147483647LL;
    dim3 grid((unsigned)nb, (unsigned)chunks);
    circuit_columns_kernel<<<grid, threads, smem, as_stream(stream)>>>(
        c, params + b0 * c.n_param_cols, cols, m, mc, max_layer_gates, X_out + b0 * (int64_t)d * m);
    if (int rc = check_launch("circuit_columns_kernel")) return rc;
  }
  return 0;
}

extern "C" int mcb200_cov_basis(const double* X, const int8_t* bits, int64_t B, int32_t n_qubits, int32_t m,
                                double* cov_out, void* stream) {
  if (B < 0 || n_qubits < 1 || m < 0) return set_error("cov_basis: bad sizes");
  if (B == 0 || m == 0) return 0;
  if (X == nullptr || cov_out == nullptr) return set_error("cov_basis: NULL pointer");
  const int d = 2 * n_qubits;
  const size_t smem = ((size_t)d * m + (size_t)m * (m + 1)) * sizeof(double);
  // wide panels: one DMMA GEMM (P = E^T diag(s) O) + antisymmetrisation; the shared-memory kernel serves the narrow ones
  if (m > 64 || smem > 96 * 1024) return cov_basis_gemm(X, bits, B, n_qubits, m, cov_out, as_stream(stream));
  if (int rc = check_cuda(cudaFuncSetAttribute(cov_basis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)smem), "cudaFuncSetAttribute(cov_basis)"))
    return rc;
  cov_basis_kernel<<<(unsigned)B, 256, smem, as_stream(stream)>>>(X, bits, n_qubits, m, cov_out);
  return check_launch("cov_basis_kernel");
}

// K1 + K2 in one call.  Fused (the column panel never leaves shared memory) for panels of up to 32 columns of
// nearest-neighbour circuits; otherwise mcb200_circuit_columns into the workspace, then mcb200_cov_basis.
static bool circuit_cov_fused(const CircuitDev& c, int m) {
  return c.nearest_neighbour && c.n_gates > 0 && m >= 2 && m <= 32 && c.n_qubits <= 256 &&
         (size_t)m * (2 * c.n_qubits + 2) * sizeof(double) <= 80 * 1024;
}

extern "C" int64_t mcb200_circuit_cov_basis_workspace_bytes(const mcb200_circuit_t* circuit_host, int64_t B, int32_t m) {
  CircuitDev c;
  if (circuit_to_dev(circuit_host, &c)) return -1;
  if (B <= 0 || m <= 0 || circuit_cov_fused(c, m)) return 0;
  return B * 2LL * c.n_qubits * m * (int64_t)sizeof(double);
}

extern "C" int mcb200_circuit_cov_basis(const mcb200_circuit_t* circuit_host, const double* params, int64_t B,
                                        const int32_t* cols, int32_t m, const int8_t* bits, double* cov_out,
                                        void* workspace, void* stream) {
  CircuitDev c;
  if (int rc = circuit_to_dev(circuit_host, &c)) return rc;
  if (B < 0 || m < 0) return set_error("circuit_cov_basis: negative batch or column count");
  if (B == 0 || m == 0) return 0;
  if (cov_out == nullptr || cols == nullptr) return set_error("circuit_cov_basis: NULL pointer");
  if (c.n_param_cols > 0 && params == nullptr) return set_error("circuit_cov_basis: params is NULL");
  if (circuit_cov_fused(c, m)) {
    const int rc = circuit_panel_launch(c, params, B, cols, m, bits, nullptr, cov_out, as_stream(stream));
    if (rc >= 0) return rc;
  }
  if (workspace == nullptr) return set_error("circuit_cov_basis: this shape needs a workspace (see *_workspace_bytes)");
  if (int rc = mcb200_circuit_columns(circuit_host, params, B, cols, m, (double*)workspace, stream)) return rc;
  return mcb200_cov_basis((const double*)workspace, bits, B, c.n_qubits, m, cov_out, stream);
}

extern "C" int mcb200_circuit_expval_projector(const mcb200_circuit_t* circuit_host, const double* params,
                                               int64_t B, const int8_t* init_bits, const int8_t* target_bits,
                                               double* out, void* stream) {
  CircuitDev c;
  if (int rc = circuit_to_dev(circuit_host, &c)) return rc;
  if (B < 0) return set_error("negative batch");
  if (B == 0) return 0;
  if (out == nullptr) return set_error("out is NULL");
  if (c.n_param_cols > 0 && params == nullptr) return set_error("params is NULL");
  const int d = 2 * c.n_qubits;
  if (d <= kTilesMaxM) {  // one warp per instance, Lambda and the Pfaffian stay in registers (expval_warp.cu)
    const int rc = expval_warp_launch(circuit_host, params, B, init_bits, target_bits, out, as_stream(stream));
    if (rc >= 0) return rc;  // -1: gate program too large for shared memory -> block-per-instance kernel below
  }
  const size_t coef_bytes = (size_t)(c.n_qubits > 16 ? c.n_qubits : 16) * 4 * sizeof(double);
  const size_t smem = ((size_t)d * d + (size_t)d * (d + 1)) * sizeof(double) + coef_bytes;
  if (smem > (size_t)kMaxDynSmem)
    return set_error("circuit_expval_projector: %d qubits exceed the fused kernel's shared memory (max 59)",
                     c.n_qubits);
  if (int rc = check_cuda(cudaFuncSetAttribute(circuit_expval_projector_kernel,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                          "cudaFuncSetAttribute(circuit_expval_projector)"))
    return rc;
  const int threads = d >= 64 ? 256 : 128;
  circuit_expval_projector_kernel<<<(unsigned)B, threads, smem, as_stream(stream)>>>(c, params, init_bits,
                                                                                   target_bits, g_abs_floor2, out);
  return check_launch("circuit_expval_projector_kernel");
}

extern "C" int mcb200_circuit_pair_blocks(const mcb200_circuit_t* circuit_host, const double* params, int64_t B,
                                          const int32_t* pair_wire0, int32_t C, const int8_t* init_bits,
                                          double* blocks_out, void* stream) {
  CircuitDev c;
  if (int rc = circuit_to_dev(circuit_host, &c)) return rc;
  if (B < 0 || C < 0) return set_error("circuit_pair_blocks: negative size");
  if (B == 0 || C == 0) return 0;
  if (pair_wire0 == nullptr || blocks_out == nullptr) return set_error("circuit_pair_blocks: NULL pointer");
  if (c.n_param_cols > 0 && params == nullptr) return set_error("params is NULL");
  const size_t smem = (size_t)c.n_gates * sizeof(mcb200_gate_t);
  if (smem > 96 * 1024) return set_error("circuit_pair_blocks: %d gates exceed the shared-memory gate table", c.n_gates);
  if (int rc = check_cuda(cudaFuncSetAttribute(circuit_pair_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)smem), "cudaFuncSetAttribute(circuit_pair_blocks)"))
    return rc;
  long long blocks = (B * C + 255) / 256;
  if (blocks > 64LL * kNumSMs) blocks = 64LL * kNumSMs;
  circuit_pair_blocks_kernel<<<(unsigned)blocks, 256, smem, as_stream(stream)>>>(c, params, B, pair_wire0, C, init_bits,
                                                                               blocks_out);
  return check_launch("circuit_pair_blocks_kernel");
}
