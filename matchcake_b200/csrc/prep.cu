// Input preparation kernels that keep the remaining small pieces of the path off the CPU:
//  * generic matchgate (4x4 complex unitary) -> real 4x4 SPTM block, R_{mu nu} = 1/4 Tr(U c_mu U^dag c_nu)
//    (utils/__init__.py:561-596, utils/majorana.py:15-47; imaginary part checked against REAL_PART_ATOL,
//    single_particle_transition_matrix.py:87-113);
//  * product-state Majorana covariance Lambda_0 and its (2n+1) extension with the displacement vector
//    (operations/state_preparation/product_state.py:141-245,
//     devices/expval_strategies/m_pfaffian/_extended_covariance.py:12-98).
#include "common.cuh"

namespace mcb200 {

struct cplx {
  double re, im;
};
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }

// Two-qubit Majoranas as signed/phased permutations: (c_mu)[i][perm[mu][i]] = ph[mu][i]  (row i has one entry).
//   c0 = X (x) I, c1 = Y (x) I, c2 = Z (x) X, c3 = Z (x) Y
__constant__ int kMajPerm[4][4] = {{2, 3, 0, 1}, {2, 3, 0, 1}, {1, 0, 3, 2}, {1, 0, 3, 2}};
// phases as (re, im)
__constant__ double kMajPh[4][4][2] = {
    {{1, 0}, {1, 0}, {1, 0}, {1, 0}},
    {{0, -1}, {0, -1}, {0, 1}, {0, 1}},
    {{1, 0}, {1, 0}, {-1, 0}, {-1, 0}},
    {{0, -1}, {0, 1}, {0, 1}, {0, -1}}};

__global__ void matchgate_sptm_kernel(const double* __restrict__ U, long long B, double* __restrict__ out,
                                      unsigned long long* max_imag_bits) {
  long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double worst = 0.0;
  if (b < B) {
    cplx u[4][4];
    const double* p = U + b * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) u[i][j] = {p[(i * 4 + j) * 2], p[(i * 4 + j) * 2 + 1]};
#pragma unroll
    for (int mu = 0; mu < 4; ++mu) {
      // M = U c_mu : M[i][perm[k]] = U[i][k] ph[k];   N = M U^dag : N[i][j] = sum_l M[i][l] conj(U[j][l])
      cplx m[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) m[i][kMajPerm[mu][k]] = cmul(u[i][k], {kMajPh[mu][k][0], kMajPh[mu][k][1]});
      cplx nn[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          cplx acc = {0.0, 0.0};
#pragma unroll
          for (int l = 0; l < 4; ++l) acc = cadd(acc, cmul(m[i][l], {u[j][l].re, -u[j][l].im}));
          nn[i][j] = acc;
        }
#pragma unroll
      for (int nu = 0; nu < 4; ++nu) {
        // Tr(N c_nu) = sum_j N[perm[j]][j]... c_nu[j][perm[j]] = ph[j]  =>  Tr = sum_j N[perm[nu][j]][j] ph[nu][j]
        cplx tr = {0.0, 0.0};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          tr = cadd(tr, cmul(nn[kMajPerm[nu][j]][j], {kMajPh[nu][j][0], kMajPh[nu][j][1]}));
        out[b * 16 + mu * 4 + nu] = 0.25 * tr.re;
        worst = fmax(worst, fabs(0.25 * tr.im));
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
  if ((threadIdx.x & 31) == 0 && worst > 0.0)
    atomicMax(max_imag_bits, (unsigned long long)__double_as_longlong(worst));  // non-negative doubles order like ints
}

// Adjoint of matchgate_sptm_kernel: with g = dL/dR (real 4 x 4),  dL/dU = 1/2 sum_{mu nu} g[mu][nu] c_nu U c_mu  (torch's
// convention for a real loss of a complex tensor: dL/dRe U + i dL/dIm U).  (c_nu U c_mu)[i][j] = ph_nu[i] U[perm_nu[i]][perm_mu[j]]
// ph_mu[perm_mu[j]] because the two-qubit Majoranas are phased involutive permutations.
__global__ void matchgate_sptm_backward_kernel(const double* __restrict__ U, long long B, const double* __restrict__ gR,
                                               double* __restrict__ gU) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  cplx u[4][4];
  const double* p = U + b * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) u[i][j] = {p[(i * 4 + j) * 2], p[(i * 4 + j) * 2 + 1]};
  cplx acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = {0.0, 0.0};
  const double* g = gR + b * 16;
#pragma unroll
  for (int mu = 0; mu < 4; ++mu)
#pragma unroll
    for (int nu = 0; nu < 4; ++nu) {
      const double w = 0.5 * g[mu * 4 + nu];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = kMajPerm[nu][i], l = kMajPerm[mu][j];
          const cplx t = cmul(cmul({kMajPh[nu][i][0], kMajPh[nu][i][1]}, u[k][l]), {kMajPh[mu][l][0], kMajPh[mu][l][1]});
          acc[i][j].re += w * t.re;
          acc[i][j].im += w * t.im;
        }
    }
  double* o = gU + b * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[(i * 4 + j) * 2] = acc[i][j].re;
      o[(i * 4 + j) * 2 + 1] = acc[i][j].im;
    }
}

// One block per product state: amplitudes (n, 2) complex as (re, im) pairs -> cov (D x D), D = 2n (+1 if extended).
__global__ void product_state_cov_kernel(const double* __restrict__ amp, int n, int extended, double* __restrict__ cov) {
  extern __shared__ double sh[];  // x[n] | y[n] | z[n]
  double* x = sh;
  double* y = sh + n;
  double* z = sh + 2 * n;
  const long long b = blockIdx.x;
  const int D = 2 * n + (extended ? 1 : 0);
  const double* a = amp + b * (long long)n * 4;
  double* out = cov + b * (long long)D * D;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const double ar = a[4 * k], ai = a[4 * k + 1], br = a[4 * k + 2], bi = a[4 * k + 3];
    x[k] = 2.0 * (ar * br + ai * bi);       // 2 Re(conj(alpha) beta)
    y[k] = 2.0 * (ar * bi - ai * br);       // 2 Im(conj(alpha) beta)
    z[k] = (ar * ar + ai * ai) - (br * br + bi * bi);
  }
  for (int e = threadIdx.x; e < D * D; e += blockDim.x) out[e] = 0.0;
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    out[(2 * j) * D + 2 * j + 1] = -z[j];
    out[(2 * j + 1) * D + 2 * j] = z[j];
    double p = 1.0;  // parity string prod_{j<l<k} z_l
    for (int k = j + 1; k < n; ++k) {
      const double v00 = y[j] * p * x[k], v01 = y[j] * p * y[k], v10 = -x[j] * p * x[k], v11 = -x[j] * p * y[k];
      out[(2 * j) * D + 2 * k] = v00;          out[(2 * k) * D + 2 * j] = -v00;
      out[(2 * j) * D + 2 * k + 1] = v01;      out[(2 * k + 1) * D + 2 * j] = -v01;
      out[(2 * j + 1) * D + 2 * k] = v10;      out[(2 * k) * D + 2 * j + 1] = -v10;
      out[(2 * j + 1) * D + 2 * k + 1] = v11;  out[(2 * k + 1) * D + 2 * j + 1] = -v11;
      p *= z[k];
    }
    if (extended) {  // displacement d[2j] = (prod_{l<j} z_l) x_j, d[2j+1] = (prod_{l<j} z_l) y_j
      double zp = 1.0;
      for (int l = 0; l < j; ++l) zp *= z[l];
      out[(2 * j) * D + 2 * n] = zp * x[j];
      out[(2 * j + 1) * D + 2 * n] = zp * y[j];
      out[(2 * n) * D + 2 * j] = -zp * x[j];
      out[(2 * n) * D + 2 * j + 1] = -zp * y[j];
    }
  }
}

}  // namespace mcb200

using namespace mcb200;

extern "C" int mcb200_matchgate_sptm(const double* U_re_im, int64_t B, double* blocks_out, double* max_imag_out,
                                     void* stream) {
  if (B < 0) return set_error("matchgate_sptm: negative batch");
  if (B == 0) return 0;
  if (U_re_im == nullptr || blocks_out == nullptr || max_imag_out == nullptr)
    return set_error("matchgate_sptm: NULL pointer");
  cudaStream_t st = as_stream(stream);
  if (int rc = check_cuda(cudaMemsetAsync(max_imag_out, 0, sizeof(double), st), "cudaMemsetAsync")) return rc;
  matchgate_sptm_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(U_re_im, B, blocks_out,
                                                                    reinterpret_cast<unsigned long long*>(max_imag_out));
  return check_launch("matchgate_sptm_kernel");
}

extern "C" int mcb200_matchgate_sptm_backward(const double* U_re_im, int64_t B, const double* grad_blocks,
                                              double* grad_U_re_im, void* stream) {
  if (B < 0) return set_error("matchgate_sptm_backward: negative batch");
  if (B == 0) return 0;
  if (U_re_im == nullptr || grad_blocks == nullptr || grad_U_re_im == nullptr)
    return set_error("matchgate_sptm_backward: NULL pointer");
  matchgate_sptm_backward_kernel<<<(unsigned)((B + 127) / 128), 128, 0, as_stream(stream)>>>(U_re_im, B, grad_blocks,
                                                                                             grad_U_re_im);
  return check_launch("matchgate_sptm_backward_kernel");
}

extern "C" int mcb200_product_state_cov(const double* amplitudes_re_im, int64_t B, int32_t n_qubits, int32_t extended,
                                        double* cov_out, void* stream) {
  if (B < 0 || n_qubits < 1) return set_error("product_state_cov: bad sizes");
  if (B == 0) return 0;
  if (amplitudes_re_im == nullptr || cov_out == nullptr) return set_error("product_state_cov: NULL pointer");
  const size_t smem = 3 * (size_t)n_qubits * sizeof(double);
  product_state_cov_kernel<<<(unsigned)B, 128, smem, as_stream(stream)>>>(amplitudes_re_im, n_qubits, extended, cov_out);
  return check_launch("product_state_cov_kernel");
}
