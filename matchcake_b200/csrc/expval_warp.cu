// Fused projector expectation value, one WARP per circuit instance (2N <= 64):
//   gate list -> X = R (d x d, shared memory, lane = column) -> Lambda = X^T Lambda_0 X as DMMA accumulator
//   fragments -> + Lambda_y -> register-tile Pfaffian (pfaffian_tiles.cuh) -> 2^-N |Pf|.
// Nothing but the parameter row (16 B for config C2) is read from HBM and 8 B are written per instance.
//
//  * Gates are applied in reverse application order from the left (see circuit.cu); consecutive gates on the
//    same wire pair are applied while the four rows are still in registers.
//  * fSWAPs (adjacent or long range) move no data: they permute a logical->physical row table.
//  * The conjugation with a basis-state Lambda_0 is a skew rank-2 update per wire,
//        Lambda += s_k (e_k^T o_k - o_k^T e_k),  e_k = X[2k,:], o_k = X[2k+1,:], s_k = -(-1)^{bit_k},
//    i.e. one DMMA (K = 4) per wire PAIR and upper tile: A-fragment columns (s e_k, -s o_k, s' e_k', -s' o_k'),
//    B-fragment rows (o_k; e_k; o_k'; e_k').  The result lands directly in the Pfaffian's register layout.
// Replaces the QNode execution path of SURVEY.md section 3.1 (devices/nif_device.py:562 -> :281 -> :506 -> :742,
// lookup_table_strategy.py:19-71, utils/_pfaffian.py:78-119).
#include "common.cuh"
#include "pfaffian_tiles.cuh"
#include "tiles_api.cuh"

namespace mcb200 {

struct WarpCircuit {
  const mcb200_gate_t* gates;
  int n_gates, n_qubits, n_param_cols;
  const double* scale;
  const double* bias;
  const double* consts;
};

__device__ __forceinline__ double wc_angle(const WarpCircuit& c, const double* prow, int col) {
  double x = prow[col];
  if (c.scale != nullptr) x = c.bias[col] + c.scale[col] * x;
  return x;
}

// {c1, t1, c2, t2} of a rotation gate: two Givens pairs new_i = c x_i + t x_j, new_j = c x_j - t x_i
__device__ __forceinline__ void wc_rotation_coef(const WarpCircuit& c, const mcb200_gate_t& g, const double* prow,
                                                 double* out) {
  const double p0 = wc_angle(c, prow, g.off), p1 = wc_angle(c, prow, g.off + 1);
  // every kind needs exactly cos/sin of the half sum and the half difference (no divergent trig):
  //   RXRX pairs (0,3): cos/-sin((p1-p0)/2), (1,2): cos/sin((p1+p0)/2)      sptm_comp_rxrx.py:48-57
  //   RYRY pairs (0,2): cos/-sin((p0+p1)/2), (1,3): cos/sin((p0-p1)/2)      sptm_comp_ryry.py:53-67
  //   RZRZ pairs (0,1): cos/ sin((p0+p1)/2), (2,3): cos/sin((p0-p1)/2)      sptm_comp_rzrz.py:51-67
  double ss, cs, sd, cd;
  sincos(p0 / 2 + p1 / 2, &ss, &cs);
  sincos(p0 / 2 - p1 / 2, &sd, &cd);
  const bool x = g.kind == MCB200_GATE_RXRX, y = g.kind == MCB200_GATE_RYRY;
  out[0] = x ? cd : cs;
  out[1] = x ? sd : (y ? -ss : ss);
  out[2] = x ? cs : cd;
  out[3] = x ? ss : sd;
}

// Compact per-gate code built once per block: kind | wire0 << 3 | wire1 << 11 | transpose << 19 | run << 20, where
// `run` counts the consecutive non-fSWAP gates (walking DOWN the list, i.e. in the order they are applied here)
// that act on the same wire pair, this gate included.
__device__ __forceinline__ int wc_kind(int code) { return code & 7; }
__device__ __forceinline__ int wc_wire0(int code) { return (code >> 3) & 0xff; }
__device__ __forceinline__ int wc_wire1(int code) { return (code >> 11) & 0xff; }
__device__ __forceinline__ int wc_run(int code) { return (code >> 20) & 0xfff; }

template <int MT>
__global__ void __launch_bounds__(192, (MT <= 4 ? 3 : 1)) expval_warp_kernel(WarpCircuit wc, const double* __restrict__ params,
                                                          long long B, const int8_t* init_bits,
                                                          const int8_t* target_bits, double floor2,
                                                          double* __restrict__ out) {
  using T = Tiles<MT>;
  constexpr int M = T::M;
  constexpr int LDX = M + 4;                 // = 4 mod 16 (in doubles) for M = 16, 32, 48, 64: spreads fragment reads
  constexpr int H = (M + 31) / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = wc.n_qubits, d = 2 * n;
  const int warps = blockDim.x >> 5, w = warp_id(), lane = lane_id(), gr = lane >> 2, tg = lane & 3;
  // shared layout: [code | off (block)] [per warp: X (d x LDX) | coef 32x4 | scratch | phys]
  int* scode = reinterpret_cast<int*>(smem_raw);
  int* soff = scode + wc.n_gates;
  const size_t gate_bytes = (((size_t)wc.n_gates * 2 * sizeof(int)) + 15) & ~(size_t)15;
  const size_t per_warp = ((((size_t)d * LDX + 128 + T::kScratch) * sizeof(double) + 64) + 15) & ~(size_t)15;
  unsigned char* wbase = smem_raw + gate_bytes + (size_t)w * per_warp;
  double* X = reinterpret_cast<double*>(wbase);
  double* coef = X + (size_t)d * LDX;
  double* scratch = coef + 128;
  unsigned char* phys = reinterpret_cast<unsigned char*>(scratch + T::kScratch);
  for (int i = threadIdx.x; i < wc.n_gates; i += blockDim.x) {
    const mcb200_gate_t g = wc.gates[i];
    int run = 0;
    if (g.kind != MCB200_GATE_FSWAP) {
      run = 1;
      for (int j = i - 1; j >= 0 && run < 4000; --j) {
        const mcb200_gate_t g2 = wc.gates[j];
        if (g2.kind == MCB200_GATE_FSWAP || g2.wire0 != g.wire0) break;
        ++run;
      }
    }
    scode[i] = g.kind | (g.wire0 << 3) | (g.wire1 << 11) | ((g.flags & MCB200_GATE_FLAG_TRANSPOSE) << 19) | (run << 20);
    soff[i] = g.off;
  }
  __syncthreads();

  // initial-state bits as a mask (wire k -> bit k), built once; s_k = -(-1)^bit_k
  unsigned long long bitmask = 0ull;
  if (init_bits != nullptr) {
    const unsigned lo = __ballot_sync(0xffffffffu, lane < n && init_bits[lane]);
    const unsigned hi = __ballot_sync(0xffffffffu, lane + 32 < n && init_bits[lane + 32 < n ? lane + 32 : 0]);
    bitmask = ((unsigned long long)hi << 32) | lo;
  }
  const double sgn_lane = (tg & 1) ? 1.0 : -1.0;  // bit 0: (+s e, -s o) with s = -1

  for (long long b = (long long)blockIdx.x * warps + w; b < B; b += (long long)gridDim.x * warps) {
    const double* prow = params + b * wc.n_param_cols;
    // X = I (lane owns column(s) lane, lane+32), phys = identity
    for (int e = lane; e < (d * LDX) / 2; e += 32) reinterpret_cast<double2*>(X)[e] = make_double2(0.0, 0.0);
    __syncwarp();
    for (int r = lane; r < d; r += 32) {
      X[r * LDX + r] = 1.0;
      phys[r] = (unsigned char)r;
    }
    __syncwarp();
    // gates in reverse application order, 32 at a time
    for (int top = wc.n_gates - 1; top >= 0; top -= 32) {
      {
        const int gi = top - lane;
        if (gi >= 0) {
          const int code = scode[gi];
          if (wc_kind(code) <= MCB200_GATE_RZRZ) {
            mcb200_gate_t g;
            g.kind = wc_kind(code);
            g.off = soff[gi];
            wc_rotation_coef(wc, g, prow, coef + 4 * lane);
          }
        }
      }
      __syncwarp();
      const int cnt = top + 1 < 32 ? top + 1 : 32;
      int q = 0;
      while (q < cnt) {
        const int code = scode[top - q];
        const int r0 = 2 * wc_wire0(code);
        if (wc_kind(code) == MCB200_GATE_FSWAP) {
          // no data moves: cyclic shift of the logical->physical row table over [r0, r0 + s)
          //   forward block: new[0:2] = old[s-2:s], new[i] = old[i-2];   transposed block: new[i] = old[i+2 mod s]
          const int s = 2 * (wc_wire1(code) - wc_wire0(code) + 1);
          if (s == 4) {  // adjacent wires: rows (0,1) <-> (2,3)
            unsigned char v = 0;
            if (lane < 4) v = phys[r0 + (lane ^ 2)];
            __syncwarp();
            if (lane < 4) phys[r0 + lane] = v;
            __syncwarp();
            ++q;
            continue;
          }
          const bool fwd = !((code >> 19) & 1);
          unsigned char v0 = 0, v1 = 0;
          const int i0 = lane, i1 = lane + 32;
          if (i0 < s) {
            int src = fwd ? i0 - 2 : i0 + 2;
            src += (src < 0) ? s : ((src >= s) ? -s : 0);
            v0 = phys[r0 + src];
          }
          if (H > 1 && i1 < s) {
            int src = fwd ? i1 - 2 : i1 + 2;
            src += (src >= s) ? -s : 0;
            v1 = phys[r0 + src];
          }
          __syncwarp();
          if (i0 < s) phys[r0 + i0] = v0;
          if (H > 1 && i1 < s) phys[r0 + i1] = v1;
          __syncwarp();
          ++q;
          continue;
        }
        const int p0 = phys[r0], p1 = phys[r0 + 1], p2 = phys[r0 + 2], p3 = phys[r0 + 3];
        int run = wc_run(code);
        if (run > cnt - q) run = cnt - q;  // the rest of the run continues in the next chunk of 32
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const int col = lane + 32 * h;
          if (col < M) {
            double x0 = X[p0 * LDX + col], x1 = X[p1 * LDX + col], x2 = X[p2 * LDX + col], x3 = X[p3 * LDX + col];
            for (int u = 0; u < run; ++u) {
              const int kind = wc_kind(scode[top - q - u]);
              if (kind <= MCB200_GATE_RZRZ) {
                const double2 ca = *reinterpret_cast<const double2*>(coef + 4 * (q + u));
                const double2 cb = *reinterpret_cast<const double2*>(coef + 4 * (q + u) + 2);
                double y0, y1, y2, y3;
                if (kind == MCB200_GATE_RXRX) {        // pairs (0,3), (1,2)
                  y0 = ca.x * x0 + ca.y * x3;  y3 = ca.x * x3 - ca.y * x0;
                  y1 = cb.x * x1 + cb.y * x2;  y2 = cb.x * x2 - cb.y * x1;
                } else if (kind == MCB200_GATE_RYRY) { // pairs (0,2), (1,3)
                  y0 = ca.x * x0 + ca.y * x2;  y2 = ca.x * x2 - ca.y * x0;
                  y1 = cb.x * x1 + cb.y * x3;  y3 = cb.x * x3 - cb.y * x1;
                } else {                                // pairs (0,1), (2,3)
                  y0 = ca.x * x0 + ca.y * x1;  y1 = ca.x * x1 - ca.y * x0;
                  y2 = cb.x * x2 + cb.y * x3;  y3 = cb.x * x3 - cb.y * x2;
                }
                x0 = y0; x1 = y1; x2 = y2; x3 = y3;
              } else {  // explicit 4x4 block
                const int gcode = scode[top - q - u];
                const int goff = soff[top - q - u];
                const double* m4 = (kind == MCB200_GATE_BLOCK4_CONST) ? wc.consts + goff : prow + goff;
                const bool tr = (gcode >> 19) & 1;
                double y[4];
#pragma unroll
                for (int a = 0; a < 4; ++a)
                  y[a] = (tr ? m4[a] : m4[4 * a]) * x0 + (tr ? m4[4 + a] : m4[4 * a + 1]) * x1 +
                         (tr ? m4[8 + a] : m4[4 * a + 2]) * x2 + (tr ? m4[12 + a] : m4[4 * a + 3]) * x3;
                x0 = y[0]; x1 = y[1]; x2 = y[2]; x3 = y[3];
              }
            }
            X[p0 * LDX + col] = x0; X[p1 * LDX + col] = x1; X[p2 * LDX + col] = x2; X[p3 * LDX + col] = x3;
          }
        }
        q += run;
      }
      __syncwarp();
    }
    // Lambda = X^T Lambda_0 X into the accumulator fragments of the upper tiles
    double c[T::NT][2];
#pragma unroll
    for (int t = 0; t < T::NT; ++t) c[t][0] = c[t][1] = 0.0;
    for (int g4 = 0; g4 < d; g4 += 4) {
      const int ra = g4 + tg, rb = g4 + (tg ^ 1);
      const bool live = ra < d;
      const double sgn = ((bitmask >> (ra >> 1)) & 1ull) ? -sgn_lane : sgn_lane;
      const double* xa = X + (live ? phys[ra] : 0) * LDX + gr;
      const double* xb = X + (live ? phys[rb] : 0) * LDX + gr;
      double af[MT], bf[MT];
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        af[t] = live ? sgn * xa[8 * t] : 0.0;
        bf[t] = live ? xb[8 * t] : 0.0;
      }
#pragma unroll
      for (int ti = 0; ti < MT; ++ti)
#pragma unroll
        for (int tj = ti; tj < MT; ++tj) dmma_acc(c[T::idx(ti, tj)], af[ti], bf[tj]);
    }
    // + Lambda_y on the 2x2 diagonal blocks (both triangles of the diagonal tiles)
    if (tg == (gr >> 1)) {
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        const int row = 8 * t + gr;
        if (row < d) {
          const double val = (target_bits != nullptr && target_bits[row >> 1]) ? 1.0 : -1.0;
          if (gr & 1) c[T::idx(t, t)][0] -= val;
          else c[T::idx(t, t)][1] += val;
        }
      }
    }
    const double pf = warp_pfaffian_tiles<MT, 0, true>(c, d, scratch);  // Lambda|_W + Lambda_y of a physical state: bounded
    if (lane == 0) out[b] = abs_floor(ldexp(pf, -n), floor2);
    __syncwarp();
  }
}

int expval_warp_launch(const mcb200_circuit_t* h, const double* params, long long B, const int8_t* init_bits,
                       const int8_t* target_bits, double* out, cudaStream_t st) {
  WarpCircuit wc{h->gates, h->n_gates, h->n_qubits, h->n_param_cols, h->scale, h->bias, h->consts};
  const int d = 2 * h->n_qubits;
  // Route: evolving the covariance gate by gate (expval_cov.cu, 24 d flop per rotation) beats building X = R and
  // conjugating (12 d per rotation + d^3) while the circuit is shallow: n_gates (an upper bound of the number of
  // rotations; the gate list lives in device memory) below d^2 / 12.
  const int route = g_expval_route;
  if (route == 2 || (route == 0 && 12LL * h->n_gates <= (long long)d * d)) {
    const int rc = expval_cov_launch(h, params, B, init_bits, target_bits, out, st);
    if (rc >= 0) return rc;  // -1: tables do not fit shared memory -> column route
  }
  const int mt = d <= 8 ? 1 : d <= 16 ? 2 : d <= 24 ? 3 : d <= 32 ? 4 : d <= 48 ? 6 : 8;
  const int M = 8 * mt, LDX = M + 4;
  const int warps = 6;
  const size_t gate_bytes = (((size_t)h->n_gates * 2 * sizeof(int)) + 15) & ~(size_t)15;
  const size_t per_warp = ((((size_t)d * LDX + 128 + (2 * M + 4 * (M + 8))) * sizeof(double) + 64) + 15) & ~(size_t)15;
  const size_t smem = gate_bytes + warps * per_warp;
  if (smem > (size_t)kMaxDynSmem) return -1;  // caller falls back to the block-per-instance kernel
  long long blocks = (B + warps - 1) / warps;
  const long long cap = 32LL * kNumSMs;
  if (blocks > cap) blocks = cap;
  auto go = [&](auto kernel) -> int {
    if (int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(expval_warp)"))
      return rc;
    kernel<<<(unsigned)blocks, warps * 32, smem, st>>>(wc, params, B, init_bits, target_bits, g_abs_floor2, out);
    return check_launch("expval_warp_kernel");
  };
  switch (mt) {
    case 1: return go(expval_warp_kernel<1>);
    case 2: return go(expval_warp_kernel<2>);
    case 3: return go(expval_warp_kernel<3>);
    case 4: return go(expval_warp_kernel<4>);
    case 6: return go(expval_warp_kernel<6>);
    default: return go(expval_warp_kernel<8>);
  }
}

}  // namespace mcb200
