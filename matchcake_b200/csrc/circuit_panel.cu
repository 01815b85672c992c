// K1 (+ K2) for column panels of nearest-neighbour circuits: X = R[:, cols] built in a COLUMN-MAJOR shared-memory panel,
// optionally followed by the basis-state conjugation cov = X^T Lambda_0 X in the same block (the d x m panel then never
// goes to global memory: config C4 saves its 268 MB round trip and a launch).
//
// circuit_columns_kernel (circuit.cu) keeps X row-major: a gate's four rows are four 8-byte accesses per column, lane =
// gate walks its columns one after the other with run-time indices -- 68 k warp instructions per C4 instance of which 28 %
// are IMAD / ISETP (profiles/r02/ncu_c4_circuit_columns.txt).  Here
//   * Xt[c][r] (leading dimension d + 2): the four rows of a gate on wires (w, w+1) are 32 contiguous bytes of column c,
//     i.e. two 16-byte shared-memory accesses in, two out;
//   * lane = gate of the layer, and the lane applies its gate to ALL the columns its warp owns, two columns per step: kind,
//     wire and coefficients are decoded once per gate and layer, consecutive lanes touch consecutive 32-byte groups
//     (conflict free); the host prunes the gate list to the backward light cone of the panel
//     (ops.CompiledCircuit.light_cone), so late layers of a deep circuit hold few gates;
//   * the trigonometry of a whole batch of layers is evaluated by all threads at once (structure of arrays in shared
//     memory, gate kind and wire packed next to it), then every warp walks the layers of the batch on its own columns
//     with warp barriers only.
// Same arithmetic per element as build_columns / brick_apply (results agree to the rounding of a fused multiply-add).
// Measured on config C4 (pruned gate list): 1.60 ms including the covariance = circuit_columns 1.37 + cov_basis 0.21 ms;
// what the fusion buys today is the 268 MB panel that is never allocated, not time -- the trigonometry (2 sincos per
// gate and instance, 21 % of the instructions are their IMADs) and the per-layer bookkeeping bound both kernels.
// Replaces the same reference code as mcb200_circuit_columns (+ mcb200_cov_basis): devices/nif_device.py:281-347,
// 168-195, 1118-1131, product_state_strategy.py:95-110.
#include "common.cuh"
#include "circuit_dev.cuh"
#include "tiles_api.cuh"

namespace mcb200 {

constexpr int kPanelThreadsK1 = 128;
constexpr int kPanelCoefCap = 256;  // gates per trigonometry batch

// one (gate, column pair) item: rows 2 w0 .. 2 w0 + 3 of the columns ca, ca + 1 (ca + 1 may be absent: `two`)
__device__ __forceinline__ void panel_apply(int kind, double* x, int ldr, bool two, double2 a, double2 b) {
  double2* p0 = reinterpret_cast<double2*>(x);
  double2* p1 = reinterpret_cast<double2*>(x + ldr);
  double2 u0 = p0[0], u1 = p0[1];                       // rows 0,1 | 2,3 of the first column
  double2 v0 = two ? p1[0] : make_double2(0.0, 0.0), v1 = two ? p1[1] : make_double2(0.0, 0.0);
  if (kind == MCB200_GATE_RXRX) {  // pairs (0,3): a, (1,2): b
    const double x0 = u0.x, x1 = u0.y, x2 = u1.x, x3 = u1.y, y0 = v0.x, y1 = v0.y, y2 = v1.x, y3 = v1.y;
    u0.x = a.x * x0 + a.y * x3;  u1.y = a.x * x3 - a.y * x0;  u0.y = b.x * x1 + b.y * x2;  u1.x = b.x * x2 - b.y * x1;
    v0.x = a.x * y0 + a.y * y3;  v1.y = a.x * y3 - a.y * y0;  v0.y = b.x * y1 + b.y * y2;  v1.x = b.x * y2 - b.y * y1;
  } else if (kind == MCB200_GATE_RYRY) {  // pairs (0,2): a, (1,3): b
    const double x0 = u0.x, x1 = u0.y, x2 = u1.x, x3 = u1.y, y0 = v0.x, y1 = v0.y, y2 = v1.x, y3 = v1.y;
    u0.x = a.x * x0 + a.y * x2;  u1.x = a.x * x2 - a.y * x0;  u0.y = b.x * x1 + b.y * x3;  u1.y = b.x * x3 - b.y * x1;
    v0.x = a.x * y0 + a.y * y2;  v1.x = a.x * y2 - a.y * y0;  v0.y = b.x * y1 + b.y * y3;  v1.y = b.x * y3 - b.y * y1;
  } else if (kind == MCB200_GATE_RZRZ) {  // pairs (0,1): a, (2,3): b
    const double x0 = u0.x, x1 = u0.y, x2 = u1.x, x3 = u1.y, y0 = v0.x, y1 = v0.y, y2 = v1.x, y3 = v1.y;
    u0.x = a.x * x0 + a.y * x1;  u0.y = a.x * x1 - a.y * x0;  u1.x = b.x * x2 + b.y * x3;  u1.y = b.x * x3 - b.y * x2;
    v0.x = a.x * y0 + a.y * y1;  v0.y = a.x * y1 - a.y * y0;  v1.x = b.x * y2 + b.y * y3;  v1.y = b.x * y3 - b.y * y2;
  } else {  // adjacent fSWAP: rows (0,1) <-> (2,3)   (sptm_fswap.py:12-27)
    const double2 t = u0, s = v0;
    u0 = u1;  u1 = t;  v0 = v1;  v1 = s;
  }
  p0[0] = u0;  p0[1] = u1;
  if (two) { p1[0] = v0;  p1[1] = v1; }
}

// explicit 4x4 block (rare): one column
__device__ __forceinline__ void panel_apply_block(const CircuitDev& c, const mcb200_gate_t& g, const double* prow, double* x) {
  const double* m4 = (g.kind == MCB200_GATE_BLOCK4_CONST) ? c.consts + g.off : prow + g.off;
  const bool tr = g.flags & MCB200_GATE_FLAG_TRANSPOSE;
  const double x0 = x[0], x1 = x[1], x2 = x[2], x3 = x[3];
#pragma unroll
  for (int a = 0; a < 4; ++a)
    x[a] = (tr ? m4[a] : m4[4 * a]) * x0 + (tr ? m4[4 + a] : m4[4 * a + 1]) * x1 + (tr ? m4[8 + a] : m4[4 * a + 2]) * x2 +
           (tr ? m4[12 + a] : m4[4 * a + 3]) * x3;
}

// COV: write cov (m x m, full skew-symmetric) instead of X.
template <bool COV>
__global__ void __launch_bounds__(kPanelThreadsK1) circuit_panel_kernel(CircuitDev c, const double* __restrict__ params,
                                                                       const int32_t* __restrict__ cols, int m,
                                                                       const int8_t* __restrict__ bits,
                                                                       double* __restrict__ out) {
  extern __shared__ __align__(16) double cp_smem[];
  const int n = c.n_qubits, d = 2 * n, ldr = d + 2;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = kPanelThreadsK1 / 32;
  double* Xt = cp_smem;                                             // m x ldr
  double2* coefA = reinterpret_cast<double2*>(Xt + (size_t)m * ldr);  // kPanelCoefCap
  double2* coefB = coefA + kPanelCoefCap;
  int* ginfo = reinterpret_cast<int*>(coefB + kPanelCoefCap);       // wire0 | kind << 16
  const long long b = blockIdx.x;
  const double* prow = params + b * c.n_param_cols;
  for (int idx = tid; idx < m * ldr; idx += kPanelThreadsK1) {
    const int cc = idx / ldr, r = idx - cc * ldr;
    Xt[idx] = (r == (cols ? cols[cc] : cc)) ? 1.0 : 0.0;
  }
  // warp w owns the column pairs [pb, pe)
  const int npairs = (m + 1) >> 1, ppw = (npairs + nw - 1) / nw;
  const int pb = w * ppw, pe = min(npairs, pb + ppw), wp = max(0, pe - pb);
  int l = c.n_layers - 1;
  while (l >= 0) {
    const int ge = c.layer_ptr[l + 1];
    int lb = l;
    while (lb > 0 && ge - c.layer_ptr[lb - 1] <= kPanelCoefCap) --lb;
    const int gb = c.layer_ptr[lb];
    __syncthreads();  // panel initialised / every warp is done with the previous batch's tables
    for (int g = gb + tid; g < ge; g += kPanelThreadsK1) {
      const mcb200_gate_t gt = c.gates[g];
      if (gt.kind <= MCB200_GATE_RZRZ) {
        double cf[4];
        gate_coefficients(c, gt, prow, cf);
        coefA[g - gb] = make_double2(cf[0], cf[1]);
        coefB[g - gb] = make_double2(cf[2], cf[3]);
      }
      ginfo[g - gb] = gt.wire0 | (gt.kind << 16);
    }
    __syncthreads();
    if (wp > 0) {
      for (int ll = l; ll >= lb; --ll) {
        const int g0 = c.layer_ptr[ll] - gb, ng = c.layer_ptr[ll + 1] - gb - g0;
        for (int g = lane; g < ng; g += 32) {   // lane = gate: the four rows of consecutive gates are consecutive 32-byte groups
          const int gi = ginfo[g0 + g], kind = gi >> 16, w0 = gi & 0xffff;
          double* x = Xt + (size_t)(2 * pb) * ldr + 2 * w0;
          if (kind <= MCB200_GATE_FSWAP) {
            const double2 a = coefA[g0 + g], bq = coefB[g0 + g];
            for (int cp = pb; cp < pe; ++cp, x += 2 * ldr) panel_apply(kind, x, ldr, 2 * cp + 1 < m, a, bq);
          } else {
            const mcb200_gate_t gt = c.gates[gb + g0 + g];
            for (int cp = pb; cp < pe; ++cp, x += 2 * ldr) {
              panel_apply_block(c, gt, prow, x);
              if (2 * cp + 1 < m) panel_apply_block(c, gt, prow, x + ldr);
            }
          }
        }
        __syncwarp();
      }
    }
    l = lb - 1;
  }
  __syncthreads();
  if (!COV) {
    double* o = out + b * (long long)d * m;
    for (int idx = tid; idx < d * m; idx += kPanelThreadsK1) {
      const int r = idx / m, cc = idx - r * m;
      o[idx] = Xt[(size_t)cc * ldr + r];
    }
  } else {
    // cov[a][b2] = sum_k s_k (X[2k][a] X[2k+1][b2] - X[2k+1][a] X[2k][b2]),  s_k = -(-1)^{bits[k]}; one pair (a < b2) per thread
    double* o = out + b * (long long)m * m;
    const int npr = m * (m - 1) / 2;
    for (int pi = tid; pi < npr; pi += kPanelThreadsK1) {
      int a = 0, rem = pi;
      while (rem >= m - 1 - a) {
        rem -= m - 1 - a;
        ++a;
      }
      const int b2 = a + 1 + rem;
      const double2* xa = reinterpret_cast<const double2*>(Xt + (size_t)a * ldr);
      const double2* xb = reinterpret_cast<const double2*>(Xt + (size_t)b2 * ldr);
      double s = 0.0;
      for (int k = 0; k < n; ++k) {
        const double2 ea = xa[k], eb = xb[k];
        const double t = ea.x * eb.y - ea.y * eb.x;
        s += (bits != nullptr && bits[k]) ? t : -t;
      }
      o[a * m + b2] = s;
      o[b2 * m + a] = -s;
    }
    for (int a = tid; a < m; a += kPanelThreadsK1) o[a * m + a] = 0.0;
  }
}

static size_t circuit_panel_smem(int d, int m) {
  return (size_t)m * (d + 2) * sizeof(double) + 2 * (size_t)kPanelCoefCap * sizeof(double2) + (size_t)kPanelCoefCap * sizeof(int);
}

// -1: not applicable (the caller takes the general kernels)
int circuit_panel_launch(const CircuitDev& c, const double* params, long long B, const int32_t* cols, int m,
                         const int8_t* bits, double* X_out, double* cov_out, cudaStream_t st) {
  const int d = 2 * c.n_qubits;
  if (!c.nearest_neighbour || c.n_gates == 0 || c.n_qubits > 0xffff || 2 * c.n_qubits > 2 * kPanelCoefCap) return -1;
  const size_t smem = circuit_panel_smem(d, m);
  if (smem > 100 * 1024) return -1;  // keep >= 2 blocks per SM; wide panels stay on the row-major kernel
  auto go = [&](auto kernel, double* out) -> int {
    if (int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(circuit_panel)"))
      return rc;
    for (long long b0 = 0; b0 < B; b0 += 2147483647LL) {
      const long long nb = B - b0 < 2147483647LL ? B - b0 : 2147483647LL;
      kernel<<<(unsigned)nb, kPanelThreadsK1, smem, st>>>(
          c, params + b0 * c.n_param_cols, cols, m, bits,
          out + b0 * (cov_out != nullptr ? (long long)m * m : (long long)d * m));
      if (int rc = check_launch("circuit_panel_kernel")) return rc;
    }
    return 0;
  };
  return cov_out != nullptr ? go(circuit_panel_kernel<true>, cov_out) : go(circuit_panel_kernel<false>, X_out);
}

}  // namespace mcb200
