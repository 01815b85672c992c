// Circuit descriptor on the device and the per-gate primitives shared by the forward (circuit.cu) and backward
// (backward.cu) chain kernels.
#pragma once
#include "common.cuh"

namespace mcb200 {

struct CircuitDev {
  const mcb200_gate_t* gates;
  const int32_t* layer_ptr;
  int n_gates, n_layers, n_qubits, n_param_cols;
  const double* scale;
  const double* bias;
  const double* consts;
  int nearest_neighbour;
};

// Coefficients of one rotation gate for one circuit instance: {c1, t1, c2, t2} (two independent Givens pairs).
// Explicit 4x4 blocks are read from consts / the parameter row when they are applied; fSWAPs need none.
__device__ __forceinline__ double load_angle(const CircuitDev& c, const double* prow, int col) {
  double x = prow[col];
  if (c.scale != nullptr) x = c.bias[col] + c.scale[col] * x;
  return x;
}

__device__ __forceinline__ void gate_coefficients(const CircuitDev& c, const mcb200_gate_t& g, const double* prow, double* out) {
  if (g.kind > MCB200_GATE_RZRZ) return;
  // every kind needs exactly cos/sin of the half sum and the half difference of its two angles (no divergent trig):
  //   RXRX pairs (0,3): cos/-sin((p1-p0)/2), (1,2): cos/sin((p1+p0)/2)      sptm_comp_rxrx.py:48-57
  //   RYRY pairs (0,2): cos/-sin((p0+p1)/2), (1,3): cos/sin((p0-p1)/2)      sptm_comp_ryry.py:53-67
  //   RZRZ pairs (0,1): cos/ sin((p0+p1)/2), (2,3): cos/sin((p0-p1)/2)      sptm_comp_rzrz.py:51-67 (imaginary parts ~0)
  const double p0 = load_angle(c, prow, g.off), p1 = load_angle(c, prow, g.off + 1);
  double ss, cs, sd, cd;
  sincos(p0 / 2 + p1 / 2, &ss, &cs);
  sincos(p0 / 2 - p1 / 2, &sd, &cd);
  const bool x = g.kind == MCB200_GATE_RXRX, y = g.kind == MCB200_GATE_RYRY;
  out[0] = x ? cd : cs;
  out[1] = x ? sd : (y ? -ss : ss);
  out[2] = x ? cs : cd;
  out[3] = x ? ss : sd;
}

// x <- G x on rows 2*wire0 .. of the columns cc = first, first + step, ... < mc of X (one gate, many columns).
__device__ __forceinline__ void apply_gate_columns(const CircuitDev& c, const mcb200_gate_t& g, const double* cf,
                                                   const double* prow, double* X, int ld, int first, int step,
                                                   int mc) {
  double* xb = X + (2 * g.wire0) * ld;
  switch (g.kind) {
    case MCB200_GATE_RXRX: {  // Givens pairs (0,3), (1,2)
      const double c1 = cf[0], t1 = cf[1], c2 = cf[2], t2 = cf[3];
      for (int cc = first; cc < mc; cc += step) {
        double* x = xb + cc;
        const double x0 = x[0], x1 = x[ld], x2 = x[2 * ld], x3 = x[3 * ld];
        x[0] = c1 * x0 + t1 * x3;
        x[3 * ld] = c1 * x3 - t1 * x0;
        x[ld] = c2 * x1 + t2 * x2;
        x[2 * ld] = c2 * x2 - t2 * x1;
      }
      break;
    }
    case MCB200_GATE_RYRY: {  // pairs (0,2), (1,3)
      const double c1 = cf[0], t1 = cf[1], c2 = cf[2], t2 = cf[3];
      for (int cc = first; cc < mc; cc += step) {
        double* x = xb + cc;
        const double x0 = x[0], x1 = x[ld], x2 = x[2 * ld], x3 = x[3 * ld];
        x[0] = c1 * x0 + t1 * x2;
        x[2 * ld] = c1 * x2 - t1 * x0;
        x[ld] = c2 * x1 + t2 * x3;
        x[3 * ld] = c2 * x3 - t2 * x1;
      }
      break;
    }
    case MCB200_GATE_RZRZ: {  // pairs (0,1), (2,3)
      const double c1 = cf[0], t1 = cf[1], c2 = cf[2], t2 = cf[3];
      for (int cc = first; cc < mc; cc += step) {
        double* x = xb + cc;
        const double x0 = x[0], x1 = x[ld], x2 = x[2 * ld], x3 = x[3 * ld];
        x[0] = c1 * x0 + t1 * x1;
        x[ld] = c1 * x1 - t1 * x0;
        x[2 * ld] = c2 * x2 + t2 * x3;
        x[3 * ld] = c2 * x3 - t2 * x2;
      }
      break;
    }
    case MCB200_GATE_FSWAP: {  // sptm_fswap.py:12-27
      const int s = 2 * (g.wire1 - g.wire0 + 1);
      for (int cc = first; cc < mc; cc += step) {
        double* x = xb + cc;
        if (s == 4) {
          const double x0 = x[0], x1 = x[ld];
          x[0] = x[2 * ld];
          x[ld] = x[3 * ld];
          x[2 * ld] = x0;
          x[3 * ld] = x1;
        } else if (!(g.flags & MCB200_GATE_FLAG_TRANSPOSE)) {  // new[0:2] = old[s-2:s], new[i] = old[i-2]
          const double a = x[(s - 2) * ld], b = x[(s - 1) * ld];
          for (int i = s - 1; i >= 2; --i) x[i * ld] = x[(i - 2) * ld];
          x[0] = a;
          x[ld] = b;
        } else {                                                // new[i] = old[i+2], new[s-2:s] = old[0:2]
          const double a = x[0], b = x[ld];
          for (int i = 0; i < s - 2; ++i) x[i * ld] = x[(i + 2) * ld];
          x[(s - 2) * ld] = a;
          x[(s - 1) * ld] = b;
        }
      }
      break;
    }
    default: {  // BLOCK4_*: explicit real 4x4 block (optionally transposed)
      const double* m4 = (g.kind == MCB200_GATE_BLOCK4_CONST) ? c.consts + g.off : prow + g.off;
      const bool tr = g.flags & MCB200_GATE_FLAG_TRANSPOSE;
      double mm[16];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b2 = 0; b2 < 4; ++b2) mm[4 * a + b2] = tr ? m4[4 * b2 + a] : m4[4 * a + b2];
      for (int cc = first; cc < mc; cc += step) {
        double* x = xb + cc;
        const double x0 = x[0], x1 = x[ld], x2 = x[2 * ld], x3 = x[3 * ld];
#pragma unroll
        for (int a = 0; a < 4; ++a) x[a * ld] = mm[4 * a] * x0 + mm[4 * a + 1] * x1 + mm[4 * a + 2] * x2 + mm[4 * a + 3] * x3;
      }
      break;
    }
  }
}

}  // namespace mcb200

static inline int circuit_to_dev(const mcb200_circuit_t* h, mcb200::CircuitDev* c) {
  using namespace mcb200;
  if (h == nullptr) return set_error("circuit descriptor is NULL");
  if (h->n_qubits < 1) return set_error("n_qubits must be >= 1, got %d", h->n_qubits);
  if (h->n_gates < 0 || h->n_layers < 0) return set_error("negative gate/layer count");
  if (h->n_gates > 0 && (h->gates == nullptr || h->layer_ptr == nullptr))
    return set_error("gates/layer_ptr must be device pointers when n_gates > 0");
  if ((h->scale == nullptr) != (h->bias == nullptr)) return set_error("scale and bias must be given together");
  c->gates = h->gates;
  c->layer_ptr = h->layer_ptr;
  c->n_gates = h->n_gates;
  c->n_layers = h->n_layers;
  c->n_qubits = h->n_qubits;
  c->n_param_cols = h->n_param_cols;
  c->scale = h->scale;
  c->bias = h->bias;
  c->consts = h->consts;
  c->nearest_neighbour = h->nearest_neighbour;
  return 0;
}

