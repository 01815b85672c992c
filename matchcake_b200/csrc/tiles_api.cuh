// Internal entry points of the register-tile (DMMA) Pfaffian path, m <= 64 (pfaffian_tiles.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mcb200.h"

namespace mcb200 {
constexpr int kTilesMaxM = 64;
int tiles_pfaffian_plain(const double* A, long long n, int m, int mode, double scale, double* out, cudaStream_t st);
int tiles_pfaffian_probs(const double* cov, const int8_t* outcomes, long long B, int Q, int k, double* out,
                         cudaStream_t st);
int tiles_pfaffian_sector(const double* cov, long long B, int D, const int32_t* index_sets, int T, int s,
                          double* feats, cudaStream_t st);
// 64 < m <= 128: block-per-matrix variant with the tiles in shared memory; 128 < m <= 256: tiles in an L2-resident slab
// of the caller's workspace, plain matrices and Gram cells only (pfaffian_block_tiles.cu)
constexpr int kBlockTilesMaxM = 128;
constexpr int kBlockTilesGlobalMaxM = 256;
long long block_tiles_workspace_bytes(long long n_items, int m);
int block_tiles_pfaffian_plain(const double* A, long long n, int m, int mode, double scale, double* out, void* workspace,
                               cudaStream_t st);
int block_tiles_pfaffian_probs(const double* cov, const int8_t* outcomes, long long B, int Q, int k, double* out,
                               cudaStream_t st);
int block_tiles_pfaffian_sector(const double* cov, long long B, int D, const int32_t* index_sets, int T, int s,
                                double* feats, cudaStream_t st);
int block_tiles_gram(const double* cov0, const double* cov1, long long n0, long long n1, int d, long long row_begin,
                     long long row_end, int mode, double scale, double* K, long long ldk, void* workspace,
                     cudaStream_t st);
// 128 < m <= 256, panel-blocked (pfaffian_panel.cu): first 128 indices eliminated in shared memory, Schur complement as
// one K = 128 DMMA product, remainder by the m <= 128 routine; the workspace only backs the pivoted fallback
long long panel_workspace_bytes(long long n_items, int m);
int panel_pfaffian_plain(const double* A, long long n, int m, int mode, double scale, double* out, void* workspace,
                         cudaStream_t st);
int panel_gram(const double* cov0, const double* cov1, long long n0, long long n1, int d, long long row_begin,
               long long row_end, int mode, double scale, double* K, long long ldk, void* workspace, cudaStream_t st);
// 64 < m <= 128 with the pass structure of the panel kernel (threshold pivoting in place, 256-thread blocks)
int static128_pfaffian_plain(const double* A, long long n, int m, int mode, double scale, double* out, cudaStream_t st);
int static128_pfaffian_probs(const double* cov, const int8_t* outcomes, long long B, int Q, int k, double* out,
                             cudaStream_t st);
int static128_pfaffian_sector(const double* cov, long long B, int D, const int32_t* index_sets, int T, int s,
                              double* feats, cudaStream_t st);
int static128_gram(const double* cov0, const double* cov1, long long n0, long long n1, int d, long long row_begin,
                   long long row_end, int mode, double scale, double* K, long long ldk, cudaStream_t st);
long long tiles_gram_workspace_bytes(long long n0, long long n1, int d);
int tiles_gram(const double* cov0, const double* cov1, long long n0, long long n1, int d, long long row_begin,
               long long row_end, int mode, double scale, double* K, long long ldk, void* workspace,
               cudaStream_t st);
// every outcome of k measured wires at once (outcome index MSB first); returns -1 if k is too large for the tree kernel
int tiles_probs_all(const double* cov, long long B, int k, int normalize, double* out, cudaStream_t st);
// K2 as one DMMA GEMM + antisymmetrisation (gemm.cu), for panels too large for the shared-memory kernel
int cov_basis_gemm(const double* X, const int8_t* bits, long long B, int n_qubits, int m, double* cov_out,
                   cudaStream_t st);
// K1 with the panel in registers (circuit_brick.cu): nearest-neighbour circuits, <= 64 qubits; -1 = not applicable
struct CircuitDev;
int circuit_columns_brick_launch(const CircuitDev& c, const double* params, long long B, const int32_t* cols, int m,
                                 double* X_out, cudaStream_t st);
// K1 (+ K2) with a column-major shared-memory panel (circuit_panel.cu): nearest-neighbour circuits, narrow panels;
// cov_out != NULL: write cov = X^T Lambda_0(bits) X (m x m) instead of X.  -1 = not applicable
int circuit_panel_launch(const CircuitDev& c, const double* params, long long B, const int32_t* cols, int m,
                         const int8_t* bits, double* X_out, double* cov_out, cudaStream_t st);
// covariance route of the fused projector kernel (expval_cov.cu); -1 = tables do not fit shared memory
int expval_cov_launch(const mcb200_circuit_t* h, const double* params, long long B, const int8_t* init_bits,
                      const int8_t* target_bits, double* out, cudaStream_t st);
extern thread_local int g_expval_route;  // 0 = automatic, 1 = column route, 2 = covariance route (mcb200_set_expval_route)
int expval_warp_launch(const mcb200_circuit_t* h, const double* params, long long B, const int8_t* init_bits,
                       const int8_t* target_bits, double* out, cudaStream_t st);
}  // namespace mcb200
