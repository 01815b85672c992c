// K1 for nearest-neighbour circuits on up to 64 qubits: the column panel X = R[:, cols] never leaves the registers.
//
// One warp per (circuit instance, group of C columns).  Lane g owns the four rows 4g .. 4g+3 of its C columns
// (32 lanes x 4 rows = 128 = 2 N_max rows).  A gate on wires (w, w+1) acts on rows 2w .. 2w+3:
//   * even w: exactly the rows of lane w/2 -> pure register arithmetic, no memory traffic at all;
//   * odd w : rows 2,3 of lane (w-1)/2 and rows 0,1 of its neighbour -> the warp first SHIFTS its ownership by two rows
//             (one cyclic shuffle of two rows per column), after which lane g owns rows 4g+2 .. 4g+5 and odd gates are
//             register-local too.  Brickwork circuits alternate parity every layer: one shift per layer.
// Lane g also evaluates the trigonometry of "its" gate itself, so no coefficient buffer, no shared-memory round trip
// and no barrier exists on this path; per layer a lane does 2 sincos + 12 C flop + 4 C shuffles.  The shared-memory
// kernel of circuit.cu (8-byte loads/stores for every row touched: LSU bound, profiles/r01) remains the general path:
// long-range fSWAPs, more than 64 qubits, and panels wider than 8 columns -- measured (scripts/columns_bench.py, N = 64,
// depth 64): 0.21 vs 0.37 ms at m = 4, 0.33 vs 0.38 ms at m = 8, but 0.54 vs 0.42 ms at m = 16, where 16 columns per
// lane cost 255 registers (8 resident warps) and two warps per instance would each redo the trigonometry.
// Same arithmetic as build_columns: results are bit-identical for rotation gates.
#include "circuit_dev.cuh"
#include "tiles_api.cuh"

namespace mcb200 {

constexpr int kBrickWarps = 4;

template <int C>
__device__ __forceinline__ void brick_shift_up(double (&x)[4][C]) {  // rows (4g .. 4g+3) -> (4g+2 .. 4g+5)
  const int src = (lane_id() + 1) & 31;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const double t0 = x[0][c], t1 = x[1][c];
    x[0][c] = x[2][c];
    x[1][c] = x[3][c];
    x[2][c] = __shfl_sync(0xffffffffu, t0, src);
    x[3][c] = __shfl_sync(0xffffffffu, t1, src);
  }
}

template <int C>
__device__ __forceinline__ void brick_shift_down(double (&x)[4][C]) {  // inverse of brick_shift_up
  const int src = (lane_id() + 31) & 31;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const double t2 = x[2][c], t3 = x[3][c];
    x[2][c] = x[0][c];
    x[3][c] = x[1][c];
    x[0][c] = __shfl_sync(0xffffffffu, t2, src);
    x[1][c] = __shfl_sync(0xffffffffu, t3, src);
  }
}

template <int C>
__device__ __forceinline__ void brick_apply(const CircuitDev& c, const mcb200_gate_t& g, const double* prow,
                                            double (&x)[4][C]) {
  if (g.kind <= MCB200_GATE_RZRZ) {
    double cf[4];
    gate_coefficients(c, g, prow, cf);
    const double c1 = cf[0], t1 = cf[1], c2 = cf[2], t2 = cf[3];
    if (g.kind == MCB200_GATE_RXRX) {  // pairs (0,3), (1,2)
#pragma unroll
      for (int k = 0; k < C; ++k) {
        const double p = x[0][k], q = x[3][k], r = x[1][k], s = x[2][k];
        x[0][k] = c1 * p + t1 * q;
        x[3][k] = c1 * q - t1 * p;
        x[1][k] = c2 * r + t2 * s;
        x[2][k] = c2 * s - t2 * r;
      }
    } else if (g.kind == MCB200_GATE_RYRY) {  // pairs (0,2), (1,3)
#pragma unroll
      for (int k = 0; k < C; ++k) {
        const double p = x[0][k], q = x[2][k], r = x[1][k], s = x[3][k];
        x[0][k] = c1 * p + t1 * q;
        x[2][k] = c1 * q - t1 * p;
        x[1][k] = c2 * r + t2 * s;
        x[3][k] = c2 * s - t2 * r;
      }
    } else {  // pairs (0,1), (2,3)
#pragma unroll
      for (int k = 0; k < C; ++k) {
        const double p = x[0][k], q = x[1][k], r = x[2][k], s = x[3][k];
        x[0][k] = c1 * p + t1 * q;
        x[1][k] = c1 * q - t1 * p;
        x[2][k] = c2 * r + t2 * s;
        x[3][k] = c2 * s - t2 * r;
      }
    }
  } else if (g.kind == MCB200_GATE_FSWAP) {  // adjacent wires: rows (0,1) <-> (2,3)   (sptm_fswap.py:12-27)
#pragma unroll
    for (int k = 0; k < C; ++k) {
      const double p = x[0][k], q = x[1][k];
      x[0][k] = x[2][k];
      x[1][k] = x[3][k];
      x[2][k] = p;
      x[3][k] = q;
    }
  } else {  // explicit 4x4 block
    const double* m4 = (g.kind == MCB200_GATE_BLOCK4_CONST) ? c.consts + g.off : prow + g.off;
    const bool tr = g.flags & MCB200_GATE_FLAG_TRANSPOSE;
    double mm[16];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) mm[4 * a + b] = tr ? m4[4 * b + a] : m4[4 * a + b];
#pragma unroll
    for (int k = 0; k < C; ++k) {
      const double p = x[0][k], q = x[1][k], r = x[2][k], s = x[3][k];
#pragma unroll
      for (int a = 0; a < 4; ++a) x[a][k] = mm[4 * a] * p + mm[4 * a + 1] * q + mm[4 * a + 2] * r + mm[4 * a + 3] * s;
    }
  }
}

// slot[(l * 2 + parity) * 32 + lane] = index of the layer-l gate whose wire0 is 2 * lane + parity, or -1
template <int C>
__global__ void __launch_bounds__(32 * kBrickWarps) circuit_columns_brick_kernel(CircuitDev c,
                                                                                  const double* __restrict__ params,
                                                                                  long long B, const int32_t* cols,
                                                                                  int m, double* __restrict__ X_out) {
  extern __shared__ int brick_smem[];
  int* slot = brick_smem;                       // n_layers * 64
  int* has = slot + (size_t)c.n_layers * 64;    // n_layers * 2
  for (int e = threadIdx.x; e < c.n_layers * 64; e += blockDim.x) slot[e] = -1;
  for (int e = threadIdx.x; e < c.n_layers * 2; e += blockDim.x) has[e] = 0;
  __syncthreads();
  for (int l = 0; l < c.n_layers; ++l)
    for (int g = c.layer_ptr[l] + threadIdx.x; g < c.layer_ptr[l + 1]; g += blockDim.x) {
      const int w0 = c.gates[g].wire0;
      slot[(l * 2 + (w0 & 1)) * 32 + (w0 >> 1)] = g;
      has[l * 2 + (w0 & 1)] = 1;
    }
  __syncthreads();
  const int d = 2 * c.n_qubits, lane = lane_id();
  const int chunks = (m + C - 1) / C;
  const long long items = B * chunks;
  for (long long it = (long long)blockIdx.x * kBrickWarps + warp_id(); it < items;
       it += (long long)gridDim.x * kBrickWarps) {
    const long long b = it / chunks;
    const int c0 = (int)(it - b * chunks) * C;
    const double* prow = params + b * c.n_param_cols;
    double x[4][C];
#pragma unroll
    for (int k = 0; k < C; ++k) {
      const int col = (c0 + k < m) ? (cols ? cols[c0 + k] : c0 + k) : -1;
#pragma unroll
      for (int r = 0; r < 4; ++r) x[r][k] = (4 * lane + r == col) ? 1.0 : 0.0;
    }
    int shifted = 0;
    for (int l = c.n_layers - 1; l >= 0; --l) {
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        if (!has[l * 2 + par]) continue;
        if (shifted != par) {
          if (par) brick_shift_up<C>(x); else brick_shift_down<C>(x);
          shifted = par;
        }
        const int gi = slot[(l * 2 + par) * 32 + lane];
        if (gi >= 0) brick_apply<C>(c, c.gates[gi], prow, x);
      }
    }
    if (shifted) brick_shift_down<C>(x);
    double* out = X_out + b * (long long)d * m;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = 4 * lane + r;
      if (row < d) {
#pragma unroll
        for (int k = 0; k < C; ++k)
          if (c0 + k < m) out[(long long)row * m + c0 + k] = x[r][k];
      }
    }
  }
}

int circuit_columns_brick_launch(const CircuitDev& c, const double* params, long long B, const int32_t* cols, int m,
                                 double* X_out, cudaStream_t st) {
  const size_t smem = (size_t)c.n_layers * 66 * sizeof(int);
  if (smem > 200 * 1024) return -1;  // too many layers for the slot table: general path
  auto go = [&](auto kernel, int C) -> int {
    if (int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(circuit_columns_brick)"))
      return rc;
    const long long items = B * ((m + C - 1) / C);
    long long blocks = (items + kBrickWarps - 1) / kBrickWarps;
    const long long cap = 8LL * kNumSMs;
    if (blocks > cap) blocks = cap;
    kernel<<<(unsigned)blocks, 32 * kBrickWarps, smem, st>>>(c, params, B, cols, m, X_out);
    return check_launch("circuit_columns_brick_kernel");
  };
  if (m <= 4) return go(circuit_columns_brick_kernel<4>, 4);
  return go(circuit_columns_brick_kernel<8>, 8);
}

}  // namespace mcb200
