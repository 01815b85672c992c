// Library plumbing: error reporting, launch accounting and FP64 peak probes.
#include <stdarg.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "tiles_api.cuh"

namespace mcb200 {

thread_local std::string g_last_error;
std::atomic<long long> g_launch_count{0};
thread_local double g_abs_floor2 = 0.0;
thread_local double g_gram_growth = 1e5;  // kStaticBlockGrowthPhysical
thread_local int g_expval_route = 0;

int set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return 1;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  return set_error("%s: %s", what, cudaGetErrorString(e));
}

int check_launch(const char* what) {
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return check_cuda(cudaGetLastError(), what);
}

// ---- FP64 peak probes (roofline denominators; MEASURED_PEAKS.json only has HBM and bf16) ----------------
__global__ void probe_dfma_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double x = 1.0000001, y = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
      a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void probe_dmma_kernel(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int t = 0; t < 8; ++t) c[t][0] = c[t][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * (threadIdx.x + 1);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int t = 0; t < 8; ++t)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[t][0]), "+d"(c[t][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < 8; ++t) s += c[t][0] + c[t][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace mcb200

using namespace mcb200;

extern "C" int mcb200_version(void) { return MCB200_VERSION; }
extern "C" const char* mcb200_last_error(void) { return g_last_error.c_str(); }
extern "C" int mcb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
extern "C" int mcb200_set_abs_floor(double floor2) {
  if (!(floor2 >= 0.0)) return set_error("set_abs_floor: floor2 must be >= 0");
  g_abs_floor2 = floor2;
  return 0;
}
extern "C" double mcb200_get_abs_floor(void) { return g_abs_floor2; }

extern "C" int mcb200_set_gram_pivot_growth(double growth) {
  if (!(growth >= 1.0)) return set_error("set_gram_pivot_growth: growth must be >= 1");
  g_gram_growth = growth;
  return 0;
}

extern "C" double mcb200_get_gram_pivot_growth(void) { return g_gram_growth; }
// Host-side validation of a circuit descriptor (synchronous: copies the gate list and the layer table to the host).  The
// kernels trust the descriptor -- layers of gates on DISJOINT wires (so at most N/2 gates per layer: the shared-memory
// coefficient buffers are sized for that), wires in range, parameter columns inside the parameter row.
extern "C" int mcb200_circuit_validate(const mcb200_circuit_t* c) {
  if (c == nullptr) return set_error("circuit_validate: descriptor is NULL");
  if (c->n_qubits < 1) return set_error("circuit_validate: n_qubits must be >= 1, got %d", c->n_qubits);
  if (c->n_gates < 0 || c->n_layers < 0 || c->n_param_cols < 0) return set_error("circuit_validate: negative count");
  if ((c->scale == nullptr) != (c->bias == nullptr)) return set_error("circuit_validate: scale and bias must be given together");
  if (c->n_gates == 0) return 0;
  if (c->gates == nullptr || c->layer_ptr == nullptr) return set_error("circuit_validate: gates/layer_ptr are NULL");
  std::vector<mcb200_gate_t> gates((size_t)c->n_gates);
  std::vector<int32_t> ptr((size_t)c->n_layers + 1);
  if (int rc = check_cuda(cudaMemcpy(gates.data(), c->gates, gates.size() * sizeof(mcb200_gate_t), cudaMemcpyDeviceToHost),
                          "circuit_validate: copy of the gate list"))
    return rc;
  if (int rc = check_cuda(cudaMemcpy(ptr.data(), c->layer_ptr, ptr.size() * sizeof(int32_t), cudaMemcpyDeviceToHost),
                          "circuit_validate: copy of layer_ptr"))
    return rc;
  if (ptr.front() != 0 || ptr.back() != c->n_gates) return set_error("circuit_validate: layer_ptr must run from 0 to n_gates");
  std::vector<int> owner((size_t)c->n_qubits);
  bool nn = true;
  for (int l = 0; l < c->n_layers; ++l) {
    if (ptr[l + 1] < ptr[l]) return set_error("circuit_validate: layer_ptr is not monotone at layer %d", l);
    std::fill(owner.begin(), owner.end(), -1);
    for (int g = ptr[l]; g < ptr[l + 1]; ++g) {
      const mcb200_gate_t& t = gates[(size_t)g];
      if (t.kind < MCB200_GATE_RXRX || t.kind > MCB200_GATE_BLOCK4_PARAM) return set_error("circuit_validate: gate %d has unknown kind %d", g, t.kind);
      if (t.wire0 < 0 || t.wire1 <= t.wire0 || t.wire1 >= c->n_qubits)
        return set_error("circuit_validate: gate %d acts on wires (%d, %d) outside 0 <= wire0 < wire1 < %d", g, t.wire0, t.wire1, c->n_qubits);
      if (t.kind != MCB200_GATE_FSWAP && t.wire1 != t.wire0 + 1)
        return set_error("circuit_validate: gate %d: two-qubit blocks act on consecutive wires", g);
      if (t.wire1 != t.wire0 + 1) nn = false;
      const int width = t.kind <= MCB200_GATE_RZRZ ? 2 : (t.kind == MCB200_GATE_BLOCK4_PARAM ? 16 : 0);
      if (t.off < 0 || (width && t.off + width > c->n_param_cols))
        return set_error("circuit_validate: gate %d reads parameter columns [%d, %d) of %d", g, t.off, t.off + width, c->n_param_cols);
      if (t.kind == MCB200_GATE_BLOCK4_CONST && c->consts == nullptr) return set_error("circuit_validate: gate %d needs consts", g);
      for (int w = t.wire0; w <= t.wire1; ++w) {
        if (owner[(size_t)w] >= 0)
          return set_error("circuit_validate: gates %d and %d of layer %d share wire %d (layers hold gates on disjoint wires)", owner[(size_t)w], g, l, w);
        owner[(size_t)w] = g;
      }
    }
  }
  if (c->nearest_neighbour && !nn) return set_error("circuit_validate: nearest_neighbour is set but a gate acts on distant wires");
  return 0;
}

extern "C" int mcb200_set_expval_route(int32_t route) {
  if (route < 0 || route > 2) return set_error("set_expval_route: route must be 0 (auto), 1 (columns) or 2 (covariance)");
  g_expval_route = route;
  return 0;
}
extern "C" int64_t mcb200_launch_count(void) { return g_launch_count.load(); }
extern "C" void mcb200_reset_launch_count(void) { g_launch_count.store(0); }

extern "C" double mcb200_probe_fp64_tflops(int32_t kind, int32_t iters, void* stream) {
  cudaStream_t st = as_stream(stream);
  const int blocks = kNumSMs * 8, threads = 256;
  double* buf = nullptr;
  if (cudaMalloc(&buf, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) {
    set_error("probe: cudaMalloc failed");
    return -1.0;
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best_ms = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, st);
    if (kind == 0) probe_dfma_kernel<<<blocks, threads, 0, st>>>(buf, iters);
    else probe_dmma_kernel<<<blocks, threads, 0, st>>>(buf, iters);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best_ms) best_ms = ms;
  }
  cudaError_t err = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  if (err != cudaSuccess) {
    set_error("probe: %s", cudaGetErrorString(err));
    return -1.0;
  }
  double flops;
  if (kind == 0) flops = 2.0 * 64.0 * iters * (double)blocks * threads;            // 64 FMA / thread / iter
  else flops = 2.0 * 8.0 * 256.0 * iters * (double)blocks * (threads / 32);         // 8 DMMA (8x8x4) / warp / iter
  return flops / (best_ms * 1e-3) / 1e12;
}
