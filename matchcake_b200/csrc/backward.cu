// Reverse-mode derivatives of K1 / K2 / K3.
//
// The reference gets its gradients from torch autograd over its op sequence (device advertises backprop,
// nif_device.py:949-989; torch_pfaffian's PfaffianDet / signed Pfaffian define d Pf = Pf/2 tr(A^-1 dA),
// tests/test_utils/test_pfaffian.py:86-115; kernel alignment trains through the Gram, ml/kernels/kernel.py:200-283).
// Here every forward kernel gets a hand-written adjoint:
//
//  * K3  out = s |Pf(M)| or s Pf(M):  dL/dM = g * out / 2 * M^-T, M^-1 by in-place Gauss-Jordan with partial pivoting,
//    one warp per matrix in shared memory; the result is scattered back through the same "source" the forward pass
//    gathered M with (plain matrix, cov + Lambda_y per outcome, Pauli sector sub-block).  A singular M (out == 0,
//    e.g. parity-forbidden outcomes; detected by a pivot ratio below 1e-12) contributes 0.
//  * K2  cov = X^T Lambda_0 X (basis state):  dL/dX = Lambda_0 X (G^T - G).
//  * K1  X = G_0 ... G_{n-1} I[:, cols]:  the gates are orthogonal, so the adjoint sweep UN-applies them in application
//    order to X and to dL/dX together (nothing is stored in the forward pass); a rotation gate is two Givens pairs
//    G(alpha) with dG/dalpha = K G, K = [[0,1],[-1,0]], so d L/d alpha = sum_columns (xbar_i x_j - xbar_j x_i)
//    read off the post-gate state.  One warp per gate, lanes over columns, shuffle reduction, atomics only where
//    several gates share a parameter column.
#include "circuit_dev.cuh"
#include "pfaffian_src.cuh"

namespace mcb200 {

// ---------------------------------------------------------------------------------------------- K3 adjoint
// In-place inverse of the m x m matrix A (m <= 128, leading dimension ld odd) by one warp.  False if singular.
__device__ bool warp_inverse_inplace(double* __restrict__ A, int m, int ld, int* __restrict__ perm,
                                     double* __restrict__ colk) {
  const int lane = lane_id();
  double pmax = 0.0, pmin = 1e300;
  for (int k = 0; k < m; ++k) {
    // pivot search in column k, rows k..m-1
    double best = -1.0;
    int bi = k;
    for (int i = k + lane; i < m; i += 32) {
      const double v = fabs(A[i * ld + k]);
      if (v > best) {
        best = v;
        bi = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    // numerically singular (a Pfaffian that is zero up to rounding, e.g. a parity-forbidden outcome): |Pf| has a kink
    // there and out * M^-1 would be rounding noise times 1e16 -- the caller uses the zero subgradient instead
    pmax = fmax(pmax, best);
    pmin = fmin(pmin, best);
    if (!(pmin > 1e-12 * pmax)) return false;
    if (lane == 0) perm[k] = bi;
    if (bi != k) {
      for (int c = lane; c < m; c += 32) {
        const double t = A[k * ld + c];
        A[k * ld + c] = A[bi * ld + c];
        A[bi * ld + c] = t;
      }
    }
    __syncwarp();
    const double inv = 1.0 / A[k * ld + k];
    __syncwarp();
    // row k <- row k / pivot, with the pivot slot holding 1 / pivot; each lane keeps its (<= 4) entries of row k
    double rk[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = lane + 32 * q;
      rk[q] = 0.0;
      if (c < m) {
        rk[q] = (c == k) ? inv : A[k * ld + c] * inv;
        A[k * ld + c] = rk[q];
      }
    }
    // the multipliers (column k) go to a separate vector first, so that the row updates below carry no hazard between
    // iterations and need no barrier: the rows are independent and the loop pipelines
    for (int i = lane; i < m; i += 32) colk[i] = A[i * ld + k];
    __syncwarp();
#pragma unroll 4
    for (int i = 0; i < m; ++i) {
      if (i == k) continue;
      const double f = colk[i];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = lane + 32 * q;
        if (c < m) A[i * ld + c] = (c == k) ? -f * rk[q] : fma(-f, rk[q], A[i * ld + c]);
      }
    }
    __syncwarp();
  }
  // undo the row exchanges as column exchanges, last first
  for (int k = m - 1; k >= 0; --k) {
    const int p = perm[k];
    if (p != k) {
      for (int r = lane; r < m; r += 32) {
        const double t = A[r * ld + k];
        A[r * ld + k] = A[r * ld + p];
        A[r * ld + p] = t;
      }
    }
    __syncwarp();
  }
  return true;
}

struct PlainSink {
  static constexpr bool kOverwrite = true;
  double* gA;
  int m;
  __device__ __forceinline__ void add(long long item, int i, int j, double w) const {
    gA[item * (long long)m * m + i * m + j] = w;
  }
};
struct ProbsSink {
  static constexpr bool kOverwrite = false;
  double* gcov;
  int Q, m;
  __device__ __forceinline__ void add(long long item, int i, int j, double w) const {
    atomicAdd(gcov + (item / Q) * (long long)m * m + i * m + j, w);
  }
};
struct SectorSink {
  static constexpr bool kOverwrite = false;
  double* gcov;
  const int32_t* index_sets;
  int D, T, m;
  __device__ __forceinline__ void add(long long item, int i, int j, double w) const {
    const long long b = item / T;
    const int32_t* idx = index_sets + (item - b * T) * (long long)m;
    atomicAdd(gcov + b * (long long)D * D + (long long)idx[i] * D + idx[j], w);
  }
};

// grad_M = g[item] * out[item] / 2 * M^-T, scattered through the sink.
template <class Src, class Sink>
__global__ void __launch_bounds__(128) pfaffian_backward_kernel(Src src, Sink sink, long long n_items, int m,
                                                                const double* __restrict__ out,
                                                                const double* __restrict__ gout) {
  extern __shared__ double smem[];
  const int ld = m | 1;
  const int warps = blockDim.x >> 5;
  double* A = smem + (size_t)warp_id() * ((size_t)m * ld + 64 + 128);  // + 128 ints of pivot bookkeeping + column vector
  int* perm = reinterpret_cast<int*>(A + (size_t)m * ld);
  double* colk = A + (size_t)m * ld + 64;
  const int lane = lane_id();
  for (long long item = (long long)blockIdx.x * warps + warp_id(); item < n_items;
       item += (long long)gridDim.x * warps) {
    const double coef = 0.5 * gout[item] * out[item];
    if (coef == 0.0) {
      if (Sink::kOverwrite)
        for (int e = lane; e < m * m; e += 32) sink.add(item, e / m, e % m, 0.0);
      continue;
    }
    for (int e = lane; e < m * m; e += 32) {
      const int i = e / m, j = e - i * m;
      A[i * ld + j] = src.at(item, i, j);
    }
    __syncwarp();
    const bool ok = warp_inverse_inplace(A, m, ld, perm, colk);
    __syncwarp();
    for (int e = lane; e < m * m; e += 32) {
      const int i = e / m, j = e - i * m;
      const double w = ok ? coef * A[j * ld + i] : 0.0;  // (M^-1)^T
      if (Sink::kOverwrite || w != 0.0) sink.add(item, i, j, w);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------- K2 adjoint
// grad_X[2k][c] = s_k sum_a X[2k+1][a] Y[a][c],  grad_X[2k+1][c] = -s_k sum_a X[2k][a] Y[a][c],  Y = G^T - G.
__global__ void cov_basis_backward_kernel(const double* __restrict__ X, const int8_t* bits, int n_qubits, int m,
                                          const double* __restrict__ gcov, double* __restrict__ gX) {
  extern __shared__ double smem[];
  const int d = 2 * n_qubits;
  const long long b = blockIdx.x;
  double* Xs = smem;                  // d x m
  double* Y = Xs + (size_t)d * m;     // m x (m | 1)
  const int ldy = m | 1;
  const double* src = X + b * (long long)d * m;
  const double* g = gcov + b * (long long)m * m;
  for (int e = threadIdx.x; e < d * m; e += blockDim.x) Xs[e] = src[e];
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    const int i = e / m, j = e - i * m;
    Y[i * ldy + j] = g[j * m + i] - g[i * m + j];
  }
  __syncthreads();
  double* out = gX + b * (long long)d * m;
  for (int e = threadIdx.x; e < d * m; e += blockDim.x) {
    const int r = e / m, c = e - r * m, k = r >> 1;
    const double sg = (bits != nullptr && bits[k]) ? 1.0 : -1.0;
    const double* xo = Xs + (r ^ 1) * m;  // partner row of the Majorana pair
    double s = 0.0;
    for (int a = 0; a < m; ++a) s = fma(xo[a], Y[a * ldy + c], s);
    out[e] = (r & 1) ? -sg * s : sg * s;
  }
}

// Shapes whose panels exceed shared memory (full-register gradients beyond ~60 qubits): same arithmetic straight from
// global memory / L2, Y = G^T - G formed on the fly.  One block per instance.
__global__ void cov_basis_backward_global_kernel(const double* __restrict__ X, const int8_t* bits, int n_qubits, int m,
                                                 const double* __restrict__ gcov, double* __restrict__ gX) {
  const int d = 2 * n_qubits;
  const long long b = blockIdx.x;
  const double* src = X + b * (long long)d * m;
  const double* g = gcov + b * (long long)m * m;
  double* out = gX + b * (long long)d * m;
  for (int e = threadIdx.x; e < d * m; e += blockDim.x) {
    const int r = e / m, c = e - r * m, k = r >> 1;
    const double sg = (bits != nullptr && bits[k]) ? 1.0 : -1.0;
    const double* xo = src + (long long)(r ^ 1) * m;
    double s = 0.0;
    for (int a = 0; a < m; ++a) s = fma(xo[a], g[(long long)c * m + a] - g[(long long)a * m + c], s);
    out[e] = (r & 1) ? -sg * s : sg * s;
  }
}

// ---------------------------------------------------------------------------------------------- K1 adjoint
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One block per (instance, column chunk): X and grad_X chunks (d x mc) live in shared memory.  Layers are visited in
// application order; within a layer one warp handles one gate.
__global__ void circuit_columns_backward_kernel(CircuitDev c, const double* __restrict__ params, int m, int mc,
                                                const double* __restrict__ X_in, const double* __restrict__ gX_in,
                                                double* __restrict__ gangles) {
  extern __shared__ double smem[];
  const int d = 2 * c.n_qubits;
  const long long b = blockIdx.x;
  const int c0 = blockIdx.y * mc;
  const int mcur = min(mc, m - c0);
  const int ld = mc | 1;
  double* X = smem;
  double* Gx = X + (size_t)d * ld;
  const double* prow = params + b * c.n_param_cols;
  double* grow = gangles + b * c.n_param_cols;
  for (int e = threadIdx.x; e < d * mcur; e += blockDim.x) {
    const int r = e / mcur, cc = e - r * mcur;
    X[r * ld + cc] = X_in[b * (long long)d * m + (long long)r * m + c0 + cc];
    Gx[r * ld + cc] = gX_in[b * (long long)d * m + (long long)r * m + c0 + cc];
  }
  __syncthreads();
  const int lane = lane_id(), warps = blockDim.x >> 5;
  for (int l = 0; l < c.n_layers; ++l) {
    const int g0 = c.layer_ptr[l], g1 = c.layer_ptr[l + 1];
    for (int g = g0 + warp_id(); g < g1; g += warps) {
      mcb200_gate_t gt = c.gates[g];
      if (gt.kind <= MCB200_GATE_RZRZ) {
        double cf[4];
        gate_coefficients(c, gt, prow, cf);
        // row offsets of the two Givens pairs (i1, j1), (i2, j2) inside the 4-row block
        int i1, j1, i2, j2;
        if (gt.kind == MCB200_GATE_RXRX) { i1 = 0; j1 = 3; i2 = 1; j2 = 2; }
        else if (gt.kind == MCB200_GATE_RYRY) { i1 = 0; j1 = 2; i2 = 1; j2 = 3; }
        else { i1 = 0; j1 = 1; i2 = 2; j2 = 3; }
        double* xb = X + (2 * gt.wire0) * ld;
        double* gb = Gx + (2 * gt.wire0) * ld;
        double a1 = 0.0, a2 = 0.0;
        for (int cc = lane; cc < mcur; cc += 32) {
          double xi = xb[i1 * ld + cc], xj = xb[j1 * ld + cc], gi = gb[i1 * ld + cc], gj = gb[j1 * ld + cc];
          a1 += gi * xj - gj * xi;
          // un-apply: [x_i; x_j] <- [[c, -t], [t, c]] [x_i; x_j]
          xb[i1 * ld + cc] = cf[0] * xi - cf[1] * xj;
          xb[j1 * ld + cc] = cf[0] * xj + cf[1] * xi;
          gb[i1 * ld + cc] = cf[0] * gi - cf[1] * gj;
          gb[j1 * ld + cc] = cf[0] * gj + cf[1] * gi;
          xi = xb[i2 * ld + cc]; xj = xb[j2 * ld + cc]; gi = gb[i2 * ld + cc]; gj = gb[j2 * ld + cc];
          a2 += gi * xj - gj * xi;
          xb[i2 * ld + cc] = cf[2] * xi - cf[3] * xj;
          xb[j2 * ld + cc] = cf[2] * xj + cf[3] * xi;
          gb[i2 * ld + cc] = cf[2] * gi - cf[3] * gj;
          gb[j2 * ld + cc] = cf[2] * gj + cf[3] * gi;
        }
        a1 = warp_sum(a1);
        a2 = warp_sum(a2);
        if (lane == 0) {
          // effective Givens angles (alpha1, alpha2) as functions of the two loaded angles (a0, a1):
          //   RXRX: ((a0 - a1)/2, (a0 + a1)/2)   RYRY: (-(a0 + a1)/2, (a0 - a1)/2)   RZRZ: ((a0 + a1)/2, (a0 - a1)/2)
          double d0, d1;
          if (gt.kind == MCB200_GATE_RXRX) { d0 = 0.5 * (a1 + a2); d1 = 0.5 * (a2 - a1); }
          else if (gt.kind == MCB200_GATE_RYRY) { d0 = 0.5 * (a2 - a1); d1 = -0.5 * (a1 + a2); }
          else { d0 = 0.5 * (a1 + a2); d1 = 0.5 * (a1 - a2); }
          atomicAdd(grow + gt.off, d0);
          atomicAdd(grow + gt.off + 1, d1);
        }
      } else if (gt.kind == MCB200_GATE_BLOCK4_PARAM) {
        // x_t = M x_{t+1}: grad_M[a][b] = sum_cols gx_t[a] x_{t+1}[b], x_{t+1} = M^T x_t (M orthogonal)
        const double* m4 = prow + gt.off;
        const bool tr = gt.flags & MCB200_GATE_FLAG_TRANSPOSE;
        double mm[16];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b2 = 0; b2 < 4; ++b2) mm[4 * a + b2] = tr ? m4[4 * b2 + a] : m4[4 * a + b2];
        double* xb = X + (2 * gt.wire0) * ld;
        double* gb = Gx + (2 * gt.wire0) * ld;
        double gm[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) gm[e] = 0.0;
        for (int cc = lane; cc < mcur; cc += 32) {
          double x[4], gx[4], xn[4], gn[4];
#pragma unroll
          for (int a = 0; a < 4; ++a) { x[a] = xb[a * ld + cc]; gx[a] = gb[a * ld + cc]; }
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            xn[a] = mm[a] * x[0] + mm[4 + a] * x[1] + mm[8 + a] * x[2] + mm[12 + a] * x[3];
            gn[a] = mm[a] * gx[0] + mm[4 + a] * gx[1] + mm[8 + a] * gx[2] + mm[12 + a] * gx[3];
          }
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b2 = 0; b2 < 4; ++b2) gm[4 * a + b2] += gx[a] * xn[b2];
#pragma unroll
          for (int a = 0; a < 4; ++a) { xb[a * ld + cc] = xn[a]; gb[a * ld + cc] = gn[a]; }
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const double s = warp_sum(gm[e]);
          if (lane == 0) atomicAdd(grow + gt.off + (tr ? 4 * (e & 3) + (e >> 2) : e), s);
        }
      } else {
        // fSWAP / constant block: orthogonal, parameter free -> apply the transpose to both panels
        gt.flags ^= MCB200_GATE_FLAG_TRANSPOSE;
        apply_gate_columns(c, gt, nullptr, prow, X, ld, lane, 32, mcur);
        apply_gate_columns(c, gt, nullptr, prow, Gx, ld, lane, 32, mcur);
      }
    }
    __syncthreads();
  }
}

}  // namespace mcb200

using namespace mcb200;

namespace mcb200 {
template <class Src, class Sink>
static int launch_pf_backward(const Src& src, const Sink& sink, long long n_items, int m, const double* out,
                              const double* gout, cudaStream_t st, const char* what) {
  if (n_items == 0) return 0;
  const size_t per_warp = ((size_t)m * (m | 1) + 64 + 128) * sizeof(double);
  int warps = (int)((200 * 1024) / per_warp);
  if (warps > 4) warps = 4;
  if (warps < 1) return set_error("%s: m=%d does not fit shared memory", what, m);
  const size_t smem = per_warp * warps;
  auto kernel = pfaffian_backward_kernel<Src, Sink>;
  if (int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                          "cudaFuncSetAttribute(pfaffian_backward)"))
    return rc;
  long long blocks = (n_items + warps - 1) / warps;
  const long long cap = 8LL * kNumSMs;
  if (blocks > cap) blocks = cap;
  kernel<<<(unsigned)blocks, warps * 32, smem, st>>>(src, sink, n_items, m, out, gout);
  return check_launch(what);
}
}  // namespace mcb200

constexpr int kBackwardMaxM = 128;

extern "C" int mcb200_pfaffian_backward(const double* A, int64_t n, int32_t m, const double* out,
                                        const double* grad_out, double* grad_A, void* stream) {
  if (n < 0 || m < 0) return set_error("pfaffian_backward: bad sizes");
  if (n == 0 || m == 0) return 0;
  if (A == nullptr || out == nullptr || grad_out == nullptr || grad_A == nullptr)
    return set_error("pfaffian_backward: NULL pointer");
  cudaStream_t st = as_stream(stream);
  if (m & 1) return check_cuda(cudaMemsetAsync(grad_A, 0, (size_t)n * m * m * sizeof(double), st), "cudaMemsetAsync");
  if (m > kBackwardMaxM) return set_error("pfaffian_backward: m=%d > %d", m, kBackwardMaxM);
  PlainSrc src{A, m};
  PlainSink sink{grad_A, m};
  return launch_pf_backward(src, sink, n, m, out, grad_out, st, "pfaffian_backward_kernel");
}

extern "C" int mcb200_probs_backward(const double* cov, const int8_t* outcomes, int64_t B, int32_t Q, int32_t k,
                                     const double* p, const double* grad_p, double* grad_cov, void* stream) {
  if (B < 0 || Q < 0 || k < 1) return set_error("probs_backward: bad sizes");
  if (B == 0) return 0;
  const int m = 2 * k;
  if (m > kBackwardMaxM) return set_error("probs_backward: 2k=%d > %d", m, kBackwardMaxM);
  if (grad_cov == nullptr) return set_error("probs_backward: grad_cov is NULL");
  cudaStream_t st = as_stream(stream);
  if (int rc = check_cuda(cudaMemsetAsync(grad_cov, 0, (size_t)B * m * m * sizeof(double), st), "cudaMemsetAsync"))
    return rc;
  if (Q == 0) return 0;
  if (cov == nullptr || outcomes == nullptr || p == nullptr || grad_p == nullptr)
    return set_error("probs_backward: NULL pointer");
  ProbsSrc src{cov, outcomes, Q, k, m};
  ProbsSink sink{grad_cov, Q, m};
  return launch_pf_backward(src, sink, (long long)B * Q, m, p, grad_p, st, "probs_backward_kernel");
}

extern "C" int mcb200_sector_features_backward(const double* cov, int64_t B, int32_t D, const int32_t* index_sets,
                                               int32_t T, int32_t s, const double* feats, const double* grad_feats,
                                               double* grad_cov, void* stream) {
  if (B < 0 || T < 0 || s < 0 || D < 1) return set_error("sector_features_backward: bad sizes");
  if (B == 0) return 0;
  if (grad_cov == nullptr) return set_error("sector_features_backward: grad_cov is NULL");
  cudaStream_t st = as_stream(stream);
  if (int rc = check_cuda(cudaMemsetAsync(grad_cov, 0, (size_t)B * D * D * sizeof(double), st), "cudaMemsetAsync"))
    return rc;
  if (T == 0 || s == 0 || (s & 1)) return 0;
  if (s > kBackwardMaxM) return set_error("sector_features_backward: sector size %d > %d", s, kBackwardMaxM);
  if (cov == nullptr || index_sets == nullptr || feats == nullptr || grad_feats == nullptr)
    return set_error("sector_features_backward: NULL pointer");
  SectorSrc src{cov, index_sets, D, T, s};
  SectorSink sink{grad_cov, index_sets, D, T, s};
  return launch_pf_backward(src, sink, (long long)B * T, s, feats, grad_feats, st, "sector_backward_kernel");
}

extern "C" int mcb200_cov_basis_backward(const double* X, const int8_t* bits, int64_t B, int32_t n_qubits,
                                         int32_t m, const double* grad_cov, double* grad_X, void* stream) {
  if (B < 0 || n_qubits < 1 || m < 0) return set_error("cov_basis_backward: bad sizes");
  if (B == 0 || m == 0) return 0;
  if (X == nullptr || grad_cov == nullptr || grad_X == nullptr) return set_error("cov_basis_backward: NULL pointer");
  const int d = 2 * n_qubits;
  const size_t smem = ((size_t)d * m + (size_t)m * (m | 1)) * sizeof(double);
  if (smem > (size_t)kMaxDynSmem) {
    cov_basis_backward_global_kernel<<<(unsigned)B, 256, 0, as_stream(stream)>>>(X, bits, n_qubits, m, grad_cov, grad_X);
    return check_launch("cov_basis_backward_global_kernel");
  }
  if (int rc = check_cuda(cudaFuncSetAttribute(cov_basis_backward_kernel,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                          "cudaFuncSetAttribute(cov_basis_backward)"))
    return rc;
  cov_basis_backward_kernel<<<(unsigned)B, 256, smem, as_stream(stream)>>>(X, bits, n_qubits, m, grad_cov, grad_X);
  return check_launch("cov_basis_backward_kernel");
}

extern "C" int mcb200_circuit_columns_backward(const mcb200_circuit_t* circuit_host, const double* params, int64_t B,
                                               int32_t m, const double* X, const double* grad_X,
                                               double* grad_angles, void* stream) {
  CircuitDev c;
  if (int rc = circuit_to_dev(circuit_host, &c)) return rc;
  if (B < 0 || m < 0) return set_error("circuit_columns_backward: bad sizes");
  if (B == 0 || c.n_param_cols == 0) return 0;
  if (grad_angles == nullptr || params == nullptr) return set_error("circuit_columns_backward: NULL pointer");
  cudaStream_t st = as_stream(stream);
  if (int rc = check_cuda(cudaMemsetAsync(grad_angles, 0, (size_t)B * c.n_param_cols * sizeof(double), st),
                          "cudaMemsetAsync"))
    return rc;
  if (m == 0) return 0;
  if (X == nullptr || grad_X == nullptr) return set_error("circuit_columns_backward: NULL pointer");
  if (B > 2147483647LL) return set_error("circuit_columns_backward: batch too large");
  const int d = 2 * c.n_qubits;
  int mc = m;
  while ((size_t)2 * d * (mc | 1) * sizeof(double) > 160 * 1024 && mc > 1) mc = (mc + 1) / 2;
  const size_t smem = (size_t)2 * d * (mc | 1) * sizeof(double);
  if (smem > (size_t)kMaxDynSmem) return set_error("circuit_columns_backward: d=%d does not fit shared memory", d);
  if (int rc = check_cuda(cudaFuncSetAttribute(circuit_columns_backward_kernel,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                          "cudaFuncSetAttribute(circuit_columns_backward)"))
    return rc;
  dim3 grid((unsigned)B, (unsigned)((m + mc - 1) / mc));
  circuit_columns_backward_kernel<<<grid, 256, smem, st>>>(c, params, m, mc, X, grad_X, grad_angles);
  return check_launch("circuit_columns_backward_kernel");
}
