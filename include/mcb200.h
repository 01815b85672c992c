/*
 * mcb200 -- C ABI of the B200-native matchgate-circuit contraction library (libmcb200.so).
 *
 * This is the drop-in boundary for ONE hot path of MatchCake (reference: /root/reference, v0.2.3):
 * SPTM chaining -> covariance conjugation + Majorana sub-block gather -> Pfaffian, and the kernel
 * Gram-matrix builder on top of it.  The reference has no native code of its own for this path: its
 * arithmetic runs in PyTorch (einsum/linalg.det) and in the third-party wheel torchpfaffian; each entry
 * point below cites the reference Python interface it replaces (paths relative to
 * /root/reference/src/matchcake).  INTEGRATION.md shows the ctypes / torch.library binding a MatchCake
 * maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer on the current CUDA device unless its name ends in _host;
 *  - matrices are row-major FP64, batch axis leading;  N = qubits, d = 2N Majorana modes;
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are asynchronous
 *    and stream-ordered, the library allocates nothing (scratch is passed in, see *_workspace_bytes);
 *  - return value 0 = success; non-zero = error, message via mcb200_last_error() (thread local);
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails with an error.
 */
#ifndef MCB200_H
#define MCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCB200_VERSION 100

/* ------------------------------------------------------------------ circuit descriptor */
/* Gate kinds.  Angle conventions follow the closed forms of
 * operations/single_particle_transition_matrices/sptm_comp_rxrx.py:48-57, sptm_comp_ryry.py:53-67,
 * sptm_comp_rzrz.py:51-67, sptm_fswap.py:12-27. */
enum {
  MCB200_GATE_RXRX = 0,   /* theta = params[b, off], phi = params[b, off+1]                         */
  MCB200_GATE_RYRY = 1,
  MCB200_GATE_RZRZ = 2,
  MCB200_GATE_FSWAP = 3,  /* wires (wire0 < wire1), long range allowed; flag TRANSPOSE if descending */
  MCB200_GATE_BLOCK4_CONST = 4, /* explicit real 4x4 block, 16 doubles at consts[off ..], shared by batch */
  MCB200_GATE_BLOCK4_PARAM = 5  /* explicit real 4x4 block, 16 doubles at params[b, off ..]            */
};
#define MCB200_GATE_FLAG_TRANSPOSE 1

typedef struct mcb200_gate {
  int32_t kind;
  int32_t wire0;     /* lowest wire index the block acts on; the 4x4 block sits at rows/cols 2*wire0   */
  int32_t wire1;     /* FSWAP: the other wire (> wire0); otherwise wire0 + 1                          */
  int32_t off;       /* column offset into params (or consts)                                        */
  int32_t flags;
  int32_t reserved;
} mcb200_gate_t;

/* A circuit = gates in APPLICATION order, grouped into layers of gates acting on disjoint wires:
 * layer l = gates[layer_ptr[l] .. layer_ptr[l+1]).  The global SPTM is R = G_0 G_1 ... G_{n-1}
 * (devices/nif_device.py:168-195, constants/__init__.py:33).  Optional affine map applied to every
 * parameter column: angle = bias[c] + scale[c] * params[b, c]  (ml/kernels/fermionic_pqc_kernel.py:181). */
typedef struct mcb200_circuit {
  const mcb200_gate_t* gates;   /* device, n_gates                                                   */
  const int32_t* layer_ptr;     /* device, n_layers + 1                                              */
  int32_t n_gates;
  int32_t n_layers;
  int32_t n_qubits;
  int32_t n_param_cols;         /* P: columns of params                                              */
  const double* scale;          /* device, P or NULL                                                 */
  const double* bias;           /* device, P or NULL                                                 */
  const double* consts;         /* device or NULL                                                    */
  int32_t nearest_neighbour;    /* 1: EVERY gate (fSWAPs included) acts on adjacent wires, wire1 == wire0 + 1 --
                                 * enables the register-resident chain kernel (<= 64 qubits, narrow panels);
                                 * 0: unknown / general                                               */
  int32_t reserved;
} mcb200_circuit_t;

/* ------------------------------------------------------------------ library */
int mcb200_version(void);
const char* mcb200_last_error(void);
int mcb200_device_count(void);

/* Host-side check of a circuit descriptor (synchronous: the gate list and the layer table are copied to the host).
 * The compute entry points TRUST the descriptor: every layer holds gates on disjoint wires (hence at most N/2 gates, which
 * is what the kernels' shared-memory coefficient buffers are sized for), wires are in range, two-qubit blocks act on
 * consecutive wires, parameter columns lie inside the parameter row, nearest_neighbour is only set when it is true.  A
 * host that builds descriptors itself calls this once per circuit structure (matchcake_b200.ops.CompiledCircuit
 * enforces the same rules when it builds one). */
int mcb200_circuit_validate(const mcb200_circuit_t* circuit_host);

/* Magnitude floor, sticky per calling thread.  The reference's probability read-outs are
 * utils.pfaffian(sign=False) = torch_pfaffian.PfaffianDet = sqrt(|det A| + EPSILON), EPSILON = 1e-32
 * (utils/_pfaffian.py:81,110-118; a process-wide global the reference mutates under a lock before every call).
 * With floor2 > 0 every magnitude-type output v of this library -- mcb200_pfaffian (mode ABS, scale included),
 * mcb200_probs / mcb200_probs_all (before the optional normalisation), mcb200_circuit_expval_projector,
 * mcb200_gram / mcb200_gram_blocks4 entries -- is returned as sqrt(v^2 + floor2).  The caller converts the reference's
 * EPSILON to output units: LookupTableStrategy (lookup_table_strategy.py:70-71, basis-state inputs, Gram) floors the
 * probability itself (floor2 = EPSILON); ProductStateProbabilityStrategy (product_state_strategy.py:211-212) floors
 * Pf before the 2^-k factor (floor2 = 4^-k EPSILON).  0 (the initial value) switches the floor off. */
int mcb200_set_abs_floor(double floor2);
double mcb200_get_abs_floor(void);

/* Accuracy / speed of mcb200_gram for d <= 64, sticky per calling thread.  A Gram cell |Pf(cov0[i] + cov1[j])| is
 * eliminated in the natural index order with 4 x 4 block pivots as long as a guard bounds the growth of every update by
 * `growth`; a cell whose guard fails is computed with partial pivoting instead (slower, accuracy of the reference's LU).
 * The initial value 1e5 is calibrated on the covariances of the reference's kernel ansatz (FermionicPQCKernel,
 * fermionic_pqc_kernel.py:166-199: shallow light cones; worst relative error measured 1e-14).  For covariances of generic
 * (Haar-like) Gaussian states the same bound leaves about 2 % of the cells off by more than 1e-10 (worst 1e-8, profiles/r02/guard_study.txt); 1e3 keeps
 * every cell within 1e-11 of the reference at the price of 10 - 40 % of the cells taking the pivoted route.
 * matchcake_b200.ops.gram(pivoting="auto") decides per call by evaluating a sample of the cells both ways. */
int mcb200_set_gram_pivot_growth(double growth);
double mcb200_get_gram_pivot_growth(void);

/* ------------------------------------------------------------------ K1: SPTM chain */
/* Columns of the global SPTM of a gate-list circuit: X[b] = R_b[:, cols] (d x m), one instance per
 * batch row of params (B, P).  cols = NULL -> all d columns (X = R).  Replaces the per-gate Python
 * loop NonInteractingFermionicDevice.apply_generator (devices/nif_device.py:281-347) ->
 * contraction container (contraction_strategies/neighbours_strategy.py:13-26) -> Sptm.pad
 * (single_particle_transition_matrix.py:379-402) -> update_single_particle_transition_matrix
 * (nif_device.py:168-195), and NIFKernel._x_to_sptm (ml/kernels/nif_kernel.py:179-208). */
int mcb200_circuit_columns(const mcb200_circuit_t* circuit_host, const double* params, int64_t B,
                           const int32_t* cols, int32_t m, double* X_out, void* stream);

/* Dense chain step: C[b] = op(A[b]) @ op(B[b]), (n x n), FP64 DMMA tiles.  Batch strides in elements
 * (0 = operand shared by the whole batch).  Replaces utils/math.py:370-444 fermionic_operator_matmul /
 * einsum("...ij,...jk->...ik") (nif_device.py:195, single_particle_transition_matrix.py:404-418) and
 * the adjoint-then-multiply of NIFKernel.circuit (nif_kernel.py:159-177: R = b0 @ b1^T). */
int mcb200_sptm_matmul(const double* A, int64_t strideA, int transA, const double* Bm, int64_t strideB,
                       int transB, double* C, int64_t strideC, int32_t n, int64_t batch, void* stream);

/* Generic matchgate -> SPTM block: U (B, 4, 4) complex as interleaved (re, im) doubles -> blocks_out (B, 4, 4)
 * real, R_{mu nu} = 1/4 Tr(U c_mu U^dag c_nu) (utils/__init__.py:561-596).  max_imag_out (device double) receives
 * max |Im R| so that the caller can enforce REAL_PART_ATOL (single_particle_transition_matrix.py:87-113). */
int mcb200_matchgate_sptm(const double* U_re_im, int64_t B, double* blocks_out, double* max_imag_out, void* stream);

/* ------------------------------------------------------------------ K2: covariance conjugation + gather */
/* Majorana covariance of product states: amplitudes (B, n, 2) complex as interleaved (re, im) -> cov_out
 * (B, D, D), D = 2n, or D = 2n + 1 with the displacement column/row when extended != 0
 * (operations/state_preparation/product_state.py:141-245, _extended_covariance.py:12-98). */
int mcb200_product_state_cov(const double* amplitudes_re_im, int64_t B, int32_t n_qubits, int32_t extended,
                             double* cov_out, void* stream);

/* Basis-state input |x>: cov[b] = (R_b^T Lambda_0 R_b)[maj, maj] from X[b] = R_b[:, maj] (d x m);
 * Lambda_0 = (+)_k (-z_k) [[0,1],[-1,0]], z_k = (-1)^{bits[k]} (bits = NULL -> |0..0>).
 * Replaces NonInteractingFermionicDevice.covariance_matrix (nif_device.py:1118-1131) +
 * ProductStateProbabilityStrategy.extract_majorana_submatrix (product_state_strategy.py:95-110) for
 * ProductState.from_basis_state inputs (product_state.py:46-79,203-205). */
int mcb200_cov_basis(const double* X, const int8_t* bits, int64_t B, int32_t n_qubits, int32_t m,
                     double* cov_out, void* stream);

/* K1 + K2 in one call for basis-state inputs: cov[b] = (R_b^T Lambda_0(bits) R_b)[cols, cols] (m x m, full skew matrix)
 * straight from the gate list -- mcb200_circuit_columns followed by mcb200_cov_basis without the (B, d, m) panel in
 * global memory when the circuit is nearest-neighbour and the panel has at most 32 columns (marginals over <= 16 wires:
 * qml.probs(wires), nif_device.py:430-479); other shapes run the two kernels through `workspace`
 * (mcb200_circuit_cov_basis_workspace_bytes, 0 when fused). */
int64_t mcb200_circuit_cov_basis_workspace_bytes(const mcb200_circuit_t* circuit_host, int64_t B, int32_t m);
int mcb200_circuit_cov_basis(const mcb200_circuit_t* circuit_host, const double* params, int64_t B,
                             const int32_t* cols, int32_t m, const int8_t* bits, double* cov_out,
                             void* workspace, void* stream);

/* General input: cov[b] = (R~^T L0 R~)[idx, idx] for a dense initial (extended) covariance L0 (D x D,
 * D = d or d+1; batch stride strideL0 = 0 if shared) and R (B, d, d) lifted by R (+) 1 when D = d+1
 * (devices/expval_strategies/m_pfaffian/_extended_covariance.py:101-126, nif_device.py:1134-1164).
 * idx = NULL -> all D indices.  workspace: mcb200_cov_dense_workspace_bytes. */
int64_t mcb200_cov_dense_workspace_bytes(int64_t B, int32_t D, int32_t m);
int mcb200_cov_dense(const double* R, int64_t strideR, const double* L0, int64_t strideL0, int64_t B,
                     int32_t d, int32_t D, const int32_t* idx, int32_t m, double* cov_out,
                     void* workspace, void* stream);

/* ------------------------------------------------------------------ K3: Pfaffians */
enum { MCB200_PF_ABS = 0, MCB200_PF_SIGNED = 1 };

/* out[i] = scale * Pf(A[i]) (signed) or scale * |Pf(A[i])|, A (n, m, m) real skew-symmetric (only the
 * strict upper triangle is read).  m = 0 -> 1, odd m -> 0.  Replaces utils.pfaffian / signed_pfaffian
 * (utils/_pfaffian.py:17-36,78-119), i.e. torch_pfaffian.pfaffian(sign=True) and PfaffianDet.
 * workspace: mcb200_pfaffian_workspace_bytes (0 for m <= 128). */
int64_t mcb200_pfaffian_workspace_bytes(int64_t n, int32_t m);
int mcb200_pfaffian(const double* A, int64_t n, int32_t m, int32_t mode, double scale, double* out,
                    void* workspace, void* stream);

/* out[b, q] = 2^-k |Pf(cov[b] + Lambda_y(outcomes[q]))|, cov (B, 2k, 2k), outcomes (Q, k) bits.
 * normalize != 0 divides each row by its sum (qml.probs, nif_device.py:476-479).
 * Replaces ProductStateProbabilityStrategy.__call__ (product_state_strategy.py:138-213) and
 * LookupTableStrategy.__call__ (lookup_table_strategy.py:19-71) + base/lookup_table.py:132-205. */
int mcb200_probs(const double* cov, const int8_t* outcomes, int64_t B, int32_t Q, int32_t k,
                 int32_t normalize, double* out, void* stream);

/* All 2^k outcomes of k measured wires at once (qml.probs, nif_device.py:430-479): out[b, q], outcome index q
 * MSB first (states_to_binary, nif_device.py:151-165).  The 2^k Pfaffians share a binary tree of partial
 * eliminations (k <= 10); wider marginals enumerate the outcomes into `workspace` and call mcb200_probs. */
int64_t mcb200_probs_all_workspace_bytes(int32_t k);
int mcb200_probs_all(const double* cov, int64_t B, int32_t k, int32_t normalize, double* out,
                     void* workspace, void* stream);

/* Pauli-sum read-out, two steps so that the summation order is deterministic:
 *   feats[b, t] = Pf(cov[b][S_t, S_t])  for index_sets (T, s) int32 (sorted Majorana index sets of ONE parity
 *   sector of size s; s = 0 -> 1, s = 2 -> direct gather), cov (B, D, D);
 *   out[b] (+)= sum_t feats[b, t] * coeffs[t].
 * Replaces sector_pfaffian_features (utils/_pfaffian.py:39-75) and the sector loop of
 * MPfaffianExpvalStrategy.__call__ (m_pfaffian_expval_strategy.py:291-304). */
int mcb200_sector_features(const double* cov, int64_t B, int32_t D, const int32_t* index_sets,
                           int32_t T, int32_t s, double* feats, void* stream);
int mcb200_rowdot(const double* feats, const double* coeffs, int64_t B, int32_t T, int32_t accumulate,
                  double* out, void* stream);

/* Fused Hamiltonian read-out for the small sectors:
 *   out[b] (+)= sum_t coeffs[t] * Pf(cov[b][S_t, S_t]),  S_t = indices[term_ptr[t] .. term_ptr[t+1])
 * with index sets of MIXED sizes 0 (-> 1), 2, 4 and 6 in one launch (odd sizes contribute 0); cov (B, D, D).
 * The caller guarantees every set has size <= 6; larger sectors go through mcb200_sector_features + mcb200_rowdot
 * with accumulate = 1.  Deterministic summation order.
 * Replaces the sector loop of MPfaffianExpvalStrategy.__call__ (m_pfaffian_expval_strategy.py:270-306) and its
 * sector_pfaffian_features / tensordot calls (utils/_pfaffian.py:39-75) for Pauli sums whose Majorana strings are
 * short (nearest-neighbour XX/YY/ZZ/Z Hamiltonians: sectors 2 and 4). */
int mcb200_pauli_sum_small(const double* cov, int64_t B, int32_t D, const int32_t* term_ptr,
                           const int32_t* indices, const double* coeffs, int32_t T, int32_t accumulate,
                           double* out, void* stream);

/* ------------------------------------------------------------------ fused projector expval */
/* out[b] = <y| rho_b |y> on ALL wires for the circuit applied to the basis state |bits>:
 * one kernel, SPTM columns + covariance + Pfaffian stay on chip (config C2 of BASELINE.json).
 * Replaces the whole QNode execution path of SURVEY.md section 3.1.
 * Up to 32 qubits this is the warp-per-instance kernel the benchmark measures.  For 33 .. 59 qubits the entry point
 * still works (block-per-instance kernel, scalar Pfaffian) but the three-call sequence mcb200_circuit_columns ->
 * mcb200_cov_basis -> mcb200_probs is 1.6-2.5x faster (tile Pfaffians, DMMA conjugation) and has no size limit below
 * 64 qubits; it needs caller-provided buffers for X and cov, which is why it is not hidden behind this signature. */
int mcb200_circuit_expval_projector(const mcb200_circuit_t* circuit_host, const double* params,
                                    int64_t B, const int8_t* init_bits, const int8_t* target_bits,
                                    double* out, void* stream);

/* Route of mcb200_circuit_expval_projector for <= 32 qubits, sticky per calling thread (tests and benchmarks pin it;
 * 0 = automatic is the initial value):  1 = column route (X = R[:, :] built gate by gate from the right, then
 * Lambda = X^T Lambda_0 X on the FP64 tensor pipe: deep circuits),  2 = covariance route (Lambda <- G^T Lambda G gate by
 * gate, no conjugation: shallow circuits, config C2).  Both replace nif_device.py:168-195 + :1118-1131. */
int mcb200_set_expval_route(int32_t route);

/* ------------------------------------------------------------------ Gram builder */
/* K[i, j] = scale * |Pf(cov0[i] + cov1[j])| for rows row_begin <= i < row_end (SURVEY.md A.5:
 * equals P(0..0 | R_i R_j^T) of NIFKernel._compute_similarities_func, nif_kernel.py:210-221).
 * mode MCB200_GRAM_REFERENCE reproduces GramMatrix.apply_/symmetrize_ (gram_matrix.py:84-161,177-201):
 * only cells i>j (n0 >= n1) or j>i (n0 < n1) are evaluated, the other triangle of the leading
 * min(n0,n1)^2 block is mirrored and the diagonal is 1.  MCB200_GRAM_FULL evaluates every cell.
 * MCB200_GRAM_TRIANGLE evaluates the reference's cells only and writes nothing else (row-sharded multi-GPU
 * builds: all-gather the row blocks, then call mcb200_gram_symmetrize once).
 * K is (n0, n1) row-major with leading dimension ldk; only rows [row_begin,row_end) (and, in
 * REFERENCE mode, their mirror cells) are written. */
enum { MCB200_GRAM_REFERENCE = 0, MCB200_GRAM_FULL = 1, MCB200_GRAM_TRIANGLE = 2 };
int64_t mcb200_gram_workspace_bytes(int64_t n0, int64_t n1, int32_t d);
int mcb200_gram(const double* cov0, const double* cov1, int64_t n0, int64_t n1, int32_t d,
                int64_t row_begin, int64_t row_end, int32_t mode, double scale, double* K,
                int64_t ldk, void* workspace, void* stream);

/* Gram for circuits whose wires split into C independent 2-qubit groups (block-diagonal SPTM, e.g.
 * FermionicPQCKernel with depth_ = 1, fermionic_pqc_kernel.py:183-192): blocks (n, C, 6) hold the upper triangles
 * (01,02,03,12,13,23) of the 4x4 diagonal blocks of the per-sample covariances and
 * K[i, j] = scale * prod_c |Pf_4(blocks0[i, c] + blocks1[j, c])|  (scale = 2^-2C).  Same modes / row range as mcb200_gram. */
int mcb200_gram_blocks4(const double* blocks0, const double* blocks1, int64_t n0, int64_t n1, int32_t C,
                        int64_t row_begin, int64_t row_end, int32_t mode, double scale, double* K, int64_t ldk,
                        void* stream);

/* Input of mcb200_gram_blocks4 straight from the circuit: for every instance b and wire pair (pair_wire0[c],
 * pair_wire0[c] + 1) the 6 upper-triangular entries of the pair's 4x4 covariance block, blocks_out (B, C, 6).
 * Only the gates whose wire0 equals pair_wire0[c] act on pair c (the caller guarantees that no gate couples two
 * pairs).  Replaces NIFKernel._x_to_sptm + covariance_matrix for block-diagonal circuits. */
int mcb200_circuit_pair_blocks(const mcb200_circuit_t* circuit_host, const double* params, int64_t B,
                               const int32_t* pair_wire0, int32_t C, const int8_t* init_bits,
                               double* blocks_out, void* stream);

/* GramMatrix.symmetrize_ + _fill_diagonal_ (gram_matrix.py:127-161) on a device (n0, n1) matrix. */
int mcb200_gram_symmetrize(double* K, int64_t n0, int64_t n1, int64_t ldk, void* stream);

/* ------------------------------------------------------------------ instrumentation */
/* Number of kernel launches issued by this library in the calling process since the last reset. */
int64_t mcb200_launch_count(void);
void mcb200_reset_launch_count(void);
/* FP64 peak probes used by bench.py for the roofline denominators: runs a register-resident DFMA loop
 * (kind 0) or a DMMA m8n8k4 loop (kind 1) and returns achieved TFLOP/s (< 0 on error). */
double mcb200_probe_fp64_tflops(int32_t kind, int32_t iters, void* stream);

/* Computational-basis sampling on the device.  cov (B, 2k, 2k) is the evolved covariance
 * restricted to the k sampled wires; every (shot, instance) pair draws y in {0,1}^k from p(y) = 2^-k |Pf(cov + Lambda_y)|
 * by the chain rule p(y) = prod_j p(y_j | y_<j), the conditionals being the pivots of a pivot-free elimination
 * (csrc/sampling.cu).  out_bits (shots, B, k) int8 and/or out_prob (shots, B) = prod_j p(y_j | y_<j) = p(y).
 * forced_bits (shots, B, k) != NULL: nothing is drawn, the given strings are followed (evaluates p of known strings).
 * The stream of uniforms is a counter-based hash of (seed, shot * B + b, j): results do not depend on the launch shape.
 * Replaces KQubitsByKQubitsSampling.batch_generate_samples_by_subsets_of_k
 * (devices/sampling_strategies/k_qubits_by_k_qubits_sampling.py:121-176) + random_index (utils/math.py:230-300) +
 * the get_states_probability calls they issue per step (devices/nif_device.py:705-723,742-810). */
int mcb200_sample(const double* cov, int64_t B, int32_t k, int64_t shots, uint64_t seed, const int8_t* forced_bits,
                  int8_t* out_bits, double* out_prob, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Reverse-mode derivatives.  The reference differentiates this path with torch autograd
 * (devices/nif_device.py:949-989 advertises backprop; ml/kernels/kernel.py:200-283 trains through the Gram; the
 * Pfaffian derivative d Pf = Pf/2 tr(A^-1 dA) is torch_pfaffian's, pinned by tests/test_utils/test_pfaffian.py:86-115).
 * Each call is the adjoint of the forward entry point of the same name; grad_* outputs are OVERWRITTEN.
 * Matrix sizes up to 128; a singular matrix (forward value exactly 0) contributes a zero gradient.
 * ------------------------------------------------------------------------------------------------------------------ */

/* grad_A[n] = grad_out[n] * out[n] / 2 * A[n]^-T   (out = the forward result of mcb200_pfaffian, either mode). */
int mcb200_pfaffian_backward(const double* A, int64_t n, int32_t m, const double* out, const double* grad_out,
                             double* grad_A, void* stream);

/* Adjoint of mcb200_probs with normalize = 0: grad_cov[b] = sum_q grad_p[b,q] p[b,q] / 2 (cov[b] + Lambda_y(q))^-T. */
int mcb200_probs_backward(const double* cov, const int8_t* outcomes, int64_t B, int32_t Q, int32_t k,
                          const double* p, const double* grad_p, double* grad_cov, void* stream);

/* Adjoint of mcb200_sector_features: scatters grad_feats[b,t] feats[b,t] / 2 (sub-block)^-T into grad_cov (B, D, D). */
int mcb200_sector_features_backward(const double* cov, int64_t B, int32_t D, const int32_t* index_sets, int32_t T,
                                    int32_t s, const double* feats, const double* grad_feats, double* grad_cov,
                                    void* stream);

/* Adjoint of mcb200_matchgate_sptm: grad_blocks (B, 4, 4) real -> grad_U (B, 4, 4) complex as interleaved (re, im),
 * dL/dU = 1/2 sum_{mu nu} grad[mu][nu] c_nu U c_mu (dL/dRe U + i dL/dIm U): gradients w.r.t. the entries of generic
 * matchgates (the reference differentiates make_single_particle_transition_matrix_from_gate, utils/__init__.py:561-596,
 * with torch autograd). */
int mcb200_matchgate_sptm_backward(const double* U_re_im, int64_t B, const double* grad_blocks, double* grad_U_re_im,
                                   void* stream);

/* Adjoint of mcb200_cov_basis: grad_X = Lambda_0 X (grad_cov^T - grad_cov), X and grad_X (B, 2N, m). */
int mcb200_cov_basis_backward(const double* X, const int8_t* bits, int64_t B, int32_t n_qubits, int32_t m,
                              const double* grad_cov, double* grad_X, void* stream);

/* Adjoint of mcb200_cov_dense with respect to R (L0 must be skew-symmetric, as every covariance is):
 * grad_R[:, idx] = (L0 R~[:, idx]) (grad_cov^T - grad_cov) restricted to the rows of R; grad_R is (B, d, d), or (d, d)
 * accumulated over the batch when strideR == 0.  Workspace as reported by the _workspace_bytes query. */
int64_t mcb200_cov_dense_backward_workspace_bytes(int64_t B, int32_t D, int32_t m);
int mcb200_cov_dense_backward(const double* R, int64_t strideR, const double* L0, int64_t strideL0, int64_t B,
                              int32_t d, int32_t D, const int32_t* idx, int32_t m, const double* grad_cov,
                              double* grad_R, void* workspace, void* stream);

/* Adjoint of mcb200_circuit_columns.  X is the forward OUTPUT (B, 2N, m); the gates are un-applied in application
 * order, nothing has to be saved by the forward pass.  grad_angles (B, n_param_cols) is the derivative with respect
 * to the angle each gate actually used, i.e. AFTER the optional affine map bias + scale * x of the descriptor
 * (the caller finishes the chain rule for x / bias / scale); explicit 4x4 parameter blocks receive the derivative
 * of their 16 entries. */
int mcb200_circuit_columns_backward(const mcb200_circuit_t* circuit, const double* params, int64_t B, int32_t m,
                                    const double* X, const double* grad_X, double* grad_angles, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MCB200_H */
