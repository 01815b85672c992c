// Computational-basis sampling kept entirely on the device.
//
// The reference samples autoregressively (devices/sampling_strategies/k_qubits_by_k_qubits_sampling.py:121-176): at step
// j it asks the device for the probabilities of every prefix extended by the 2^k candidate blocks
// (get_states_probability -> one complex determinant of size 2(j+1) per candidate), normalises and draws.  The chain
// rule it implements, p(y) = prod_j p(y_j | y_<j), is exactly what a pivot-free Parlett-Reid elimination of
// M = Lambda|_W + Lambda_y produces:   2 p(y_j | y_<j) = |M^(j)[2j][2j+1]|,   M^(j) = Schur complement after j pairs,
// and only the pivot entry depends on y_j (Lambda_y touches nothing else), so
//     p(y_j = 1 | y_<j) = |a + 1| / 2,   p(y_j = 0 | y_<j) = |a - 1| / 2,   a = (Schur complement of Lambda|_W)[2j][2j+1].
// One warp draws one bit string: k dependent steps of a rank-2 update on a shrinking trailing block, m^3/6 flop per
// sample instead of sum_j 2^k (2j)^3 of the reference's formulation.  A small pivot is chosen with small probability, so
// no pivoting is needed.  The product of the conditionals is returned as well (it equals mcb200_probs of the drawn
// string -- that identity is the parity test).
#include "common.cuh"

namespace mcb200 {

// counter-based generator: splitmix64 finaliser of (seed, sample index, step) -> uniform double in [0, 1)
__device__ __forceinline__ double uniform01(unsigned long long seed, unsigned long long sample, unsigned step) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (sample * 131ull + step + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// cov (B, m, m) skew, m = 2k.  Sample index s * B + b (shots leading, like the reference's (shots, *batch, wires)).
// forced != nullptr: do not draw, follow the given bits (shots, B, k) -- used to evaluate prod_j p(y_j | y_<j).
__global__ void __launch_bounds__(256) sample_kernel(const double* __restrict__ cov, long long B, int k,
                                                     long long shots, unsigned long long seed,
                                                     const int8_t* __restrict__ forced, int8_t* __restrict__ bits,
                                                     double* __restrict__ prob) {
  extern __shared__ double smem[];
  const int m = 2 * k, ld = m | 1;
  const int warps = blockDim.x >> 5, lane = lane_id();
  double* A = smem + (size_t)warp_id() * m * ld;
  const long long total = B * shots;
  for (long long smp = (long long)blockIdx.x * warps + warp_id(); smp < total; smp += (long long)gridDim.x * warps) {
    const long long b = smp % B;
    const double* src = cov + b * (long long)m * m;
    for (int e = lane; e < m * m; e += 32) {
      const int i = e / m, j = e - i * m;
      if (i < j) A[i * ld + j] = src[e];
    }
    __syncwarp();
    double p = 1.0;
    for (int j = 0; j < k; ++j) {
      const int r0 = 2 * j, r1 = r0 + 1;
      const double a = A[r0 * ld + r1];
      const double p1 = 0.5 * fabs(a + 1.0), p0 = 0.5 * fabs(a - 1.0);
      int bit;
      if (forced != nullptr) {
        bit = forced[smp * k + j];
      } else {
        // normalise by p0 + p1 (= 1 for a physical covariance) like random_index does (utils/math.py:230-300)
        bit = uniform01(seed, (unsigned long long)smp, (unsigned)j) * (p0 + p1) < p1 ? 1 : 0;
      }
      p *= bit ? p1 : p0;
      if (lane == 0 && bits != nullptr) bits[smp * k + j] = (int8_t)bit;
      const double piv = a + (bit ? 1.0 : -1.0);
      if (piv == 0.0) {  // only reachable in forced mode: impossible string
        p = 0.0;
        break;
      }
      const double inv = 1.0 / piv;
      // trailing block: A[r][c] += (v_r u_c - u_r v_c) / piv,  u = row r0, v = row r1  (r1 < r < c)
      // (row by row, lanes over the columns to the right of the diagonal: no index arithmetic beyond an add)
      for (int r = r1 + 1; r < m - 1; ++r) {
        const double vr = A[r1 * ld + r] * inv, ur = A[r0 * ld + r] * inv;
        for (int c = r + 1 + lane; c < m; c += 32) A[r * ld + c] += vr * A[r0 * ld + c] - ur * A[r1 * ld + c];
      }
      __syncwarp();
    }
    if (lane == 0 && prob != nullptr) prob[smp] = p;
    __syncwarp();
  }
}

}  // namespace mcb200

using namespace mcb200;

extern "C" int mcb200_sample(const double* cov, int64_t B, int32_t k, int64_t shots, uint64_t seed,
                             const int8_t* forced_bits, int8_t* out_bits, double* out_prob, void* stream) {
  if (B < 0 || shots < 0 || k < 1) return set_error("sample: bad sizes");
  if (B == 0 || shots == 0) return 0;
  if (cov == nullptr) return set_error("sample: cov is NULL");
  if (out_bits == nullptr && out_prob == nullptr) return set_error("sample: no output requested");
  const int m = 2 * k;
  const size_t per_warp = (size_t)m * (m | 1) * sizeof(double);
  int warps = (int)((200 * 1024) / per_warp);
  if (warps > 8) warps = 8;
  if (warps < 1) return set_error("sample: k=%d does not fit shared memory", k);
  const size_t smem = per_warp * warps;
  if (int rc = check_cuda(cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                          "cudaFuncSetAttribute(sample)"))
    return rc;
  const long long total = (long long)B * shots;
  long long blocks = (total + warps - 1) / warps;
  const long long cap = 16LL * kNumSMs;
  if (blocks > cap) blocks = cap;
  sample_kernel<<<(unsigned)blocks, warps * 32, smem, as_stream(stream)>>>(cov, B, k, shots, seed, forced_bits, out_bits,
                                                                          out_prob);
  return check_launch("sample_kernel");
}
