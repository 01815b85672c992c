/* Plain C99 client of libmcb200.so: what a non-Python host (cgo / JNI / FFI) sees of the drop-in boundary.
 * Uses only include/mcb200.h and the CUDA runtime C API for device memory.  Exit code 0 = every check passed.
 *
 *   gcc -std=c99 -I include -I /usr/local/cuda/include tests/c/abi_client.c -o abi_client \
 *       -L matchcake_b200 -lmcb200 -L /usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/matchcake_b200
 *
 * Checks (values from the reference's own tests / paper):
 *   1. Pfaffian KATs of tests/test_utils/test_pfaffian.py:82-84,129-135: [[0,3],[-3,0]] -> 3, (a..f) = (1..6) -> 8;
 *   2. the JOSS / README circuit (4 wires, params (0.1, 0.2), projector |0000>) -> 0.9901 (joss/paper.md:186-232)
 *      through mcb200_circuit_expval_projector AND through the three-call sequence columns -> cov_basis -> probs;
 *   3. error reporting: a NULL output pointer returns non-zero and sets mcb200_last_error().
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mcb200.h"

#define CHECK_CUDA(x)                                                            \
  do {                                                                           \
    cudaError_t e_ = (x);                                                        \
    if (e_ != cudaSuccess) {                                                     \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 2;                                                                  \
    }                                                                            \
  } while (0)
#define CHECK_MCB(x)                                                             \
  do {                                                                           \
    if ((x) != 0) {                                                              \
      fprintf(stderr, "mcb200 error: %s (%s:%d)\n", mcb200_last_error(), __FILE__, __LINE__); \
      return 3;                                                                  \
    }                                                                            \
  } while (0)

static void* to_device(const void* host, size_t bytes) {
  void* d = NULL;
  if (cudaMalloc(&d, bytes) != cudaSuccess) return NULL;
  if (cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return NULL;
  return d;
}

int main(void) {
  int failures = 0;
  if (mcb200_device_count() < 1) {
    fprintf(stderr, "no CUDA device\n");
    return 4;
  }
  /* ---- 1. Pfaffian KATs */
  {
    const double a2[4] = {0, 3, -3, 0};
    const double a4[16] = {0, 1, 2, 3, -1, 0, 4, 5, -2, -4, 0, 6, -3, -5, -6, 0};
    double out[1];
    double *d2 = (double*)to_device(a2, sizeof a2), *d4 = (double*)to_device(a4, sizeof a4), *dout = NULL;
    CHECK_CUDA(cudaMalloc((void**)&dout, sizeof(double)));
    CHECK_MCB(mcb200_pfaffian(d2, 1, 2, MCB200_PF_SIGNED, 1.0, dout, NULL, NULL));
    CHECK_CUDA(cudaMemcpy(out, dout, sizeof out, cudaMemcpyDeviceToHost));
    if (fabs(out[0] - 3.0) > 1e-12) { fprintf(stderr, "Pf 2x2 = %.17g\n", out[0]); ++failures; }
    CHECK_MCB(mcb200_pfaffian(d4, 1, 4, MCB200_PF_SIGNED, 1.0, dout, NULL, NULL));
    CHECK_CUDA(cudaMemcpy(out, dout, sizeof out, cudaMemcpyDeviceToHost));
    if (fabs(out[0] - 8.0) > 1e-12) { fprintf(stderr, "Pf 4x4 = %.17g\n", out[0]); ++failures; }
    /* ---- 3. loud errors */
    if (mcb200_pfaffian(d4, 1, 4, MCB200_PF_SIGNED, 1.0, NULL, NULL, NULL) == 0 || strlen(mcb200_last_error()) == 0) {
      fprintf(stderr, "NULL output was not rejected\n");
      ++failures;
    }
    cudaFree(d2); cudaFree(d4); cudaFree(dout);
  }
  /* ---- 2. README circuit on 4 wires */
  {
    const int n = 4, d = 8;
    mcb200_gate_t gates[7];
    int32_t layer_ptr[5] = {0, 2, 4, 6, 7};
    const int kinds[3] = {MCB200_GATE_RXRX, MCB200_GATE_RYRY, MCB200_GATE_RZRZ};
    int g = 0, k, w;
    memset(gates, 0, sizeof gates);
    for (k = 0; k < 3; ++k)          /* one layer per rotation kind, pairs (0,1) and (2,3), same two parameters */
      for (w = 0; w < n - 1; w += 2) {
        gates[g].kind = kinds[k]; gates[g].wire0 = w; gates[g].wire1 = w + 1; gates[g].off = 0; ++g;
      }
    gates[g].kind = MCB200_GATE_FSWAP; gates[g].wire0 = 1; gates[g].wire1 = 2; ++g;
    {
      const double params[2] = {0.1, 0.2};
      double out[1], p[1];
      mcb200_circuit_t c;
      double *dparams = (double*)to_device(params, sizeof params), *dout = NULL, *dX = NULL, *dcov = NULL;
      int8_t zeros[4] = {0, 0, 0, 0}, *dzeros = (int8_t*)to_device(zeros, sizeof zeros);
      memset(&c, 0, sizeof c);
      c.gates = (const mcb200_gate_t*)to_device(gates, sizeof gates);
      c.layer_ptr = (const int32_t*)to_device(layer_ptr, sizeof layer_ptr);
      c.n_gates = 7; c.n_layers = 4; c.n_qubits = n; c.n_param_cols = 2; c.nearest_neighbour = 1;
      CHECK_CUDA(cudaMalloc((void**)&dout, sizeof(double)));
      CHECK_CUDA(cudaMalloc((void**)&dX, sizeof(double) * d * d));
      CHECK_CUDA(cudaMalloc((void**)&dcov, sizeof(double) * d * d));
      CHECK_MCB(mcb200_circuit_expval_projector(&c, dparams, 1, NULL, NULL, dout, NULL));
      CHECK_CUDA(cudaMemcpy(out, dout, sizeof out, cudaMemcpyDeviceToHost));
      CHECK_MCB(mcb200_circuit_columns(&c, dparams, 1, NULL, d, dX, NULL));
      CHECK_MCB(mcb200_cov_basis(dX, NULL, 1, n, d, dcov, NULL));
      CHECK_MCB(mcb200_probs(dcov, dzeros, 1, 1, n, 0, dout, NULL));
      CHECK_CUDA(cudaMemcpy(p, dout, sizeof p, cudaMemcpyDeviceToHost));
      printf("README circuit: fused %.10f, three calls %.10f (paper: 0.9901)\n", out[0], p[0]);
      if (fabs(out[0] - 0.9901) > 5e-5) ++failures;
      if (fabs(out[0] - p[0]) > 1e-13) ++failures;
      if (mcb200_launch_count() < 4) ++failures;
    }
  }
  printf(failures ? "FAILED (%d)\n" : "abi_client ok\n", failures);
  return failures ? 1 : 0;
}
