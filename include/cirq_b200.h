/*
 * cirq_b200 — C-ABI of the B200-native state evolution library for Cirq.
 *
 * The reference (quantumlib/Cirq 1.8.0.dev0) is pure Python; it has no FFI for
 * this path.  The boundary a Cirq maintainer binds is therefore defined here:
 * every entry point replaces one numpy-backed function of the reference and
 * cites it (paths relative to cirq-core/cirq/).  INTEGRATION.md shows the
 * ctypes stub for each.
 *
 * Conventions
 *   - All functions return 0 on success, non-zero on error; the message of the
 *     last error on the calling thread is returned by b2q_last_error().
 *   - Device buffers are owned by the caller (torch tensors in the Python
 *     host layer).  They are passed as raw device pointers.  Host arrays are
 *     read (or written) during the call only.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *     Kernels are asynchronous on that stream; functions that return scalars
 *     or fill host arrays synchronise the stream before returning.
 *   - dtype: B2Q_C64 = interleaved float (re,im), B2Q_C128 = interleaved double.
 *   - A state vector over n qubits is complex[2^n]; qubit "axis" a of Cirq
 *     (axis 0 = first qubit = most significant bit of the flat index,
 *     sim/state_vector_simulation_state.py:33-62) is BIT POSITION p = n-1-a.
 *     All `targets` / `bits` arguments are bit positions.
 *   - Gate matrices are row-major complex128[2^k * 2^k] on the host, with the
 *     FIRST target as the most significant bit of the row/column index — the
 *     convention of `U.reshape((2,)*2k)` in
 *     protocols/apply_unitary_protocol.py:440-466.  They are cast to the
 *     state dtype before use (same file, :450).
 *   - A density matrix over n qubits is stored as a 2n-qubit vector
 *     (row bits above column bits), sim/density_matrix_simulation_state.py:33-60.
 */
#ifndef CIRQ_B200_H_
#define CIRQ_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2Q_C64 0
#define B2Q_C128 1

#define B2Q_OK 0
#define B2Q_ERR_INVALID 1
#define B2Q_ERR_CUDA 2
#define B2Q_ERR_UNSUPPORTED 3

/* ---- library / device ---------------------------------------------------- */

/* ABI version (major*1000 + minor). */
int b2q_version(void);
/* Message of the last error raised on this thread ("" if none). */
const char* b2q_last_error(void);
/* Fills sm_count, total HBM bytes, compute capability (major*10+minor) of the
 * current device. */
int b2q_device_info(int* sm_count, uint64_t* hbm_bytes, int* cc);
/* Count of kernel launches issued by this library since load (all threads). */
uint64_t b2q_launch_count(void);

/* ---- state vector: initialise / copy ------------------------------------- */

/* |basis_index> in the computational basis (big-endian int of
 * qis/states.py:766-832 `to_valid_state_vector(int)` — one_hot on device). */
int b2q_sv_init_basis(void* state, int dtype, int n_qubits, uint64_t basis_index, void* stream);
/* state <- state * (re + i*im); replaces in-place scalar multiplies such as
 * sim/state_vector_simulation_state.py:371 (remove_qubits phase fold). */
int b2q_sv_scale(void* state, int dtype, int n_qubits, double re, double im, void* stream);

/* ---- state vector: gate application -------------------------------------- */

/* psi <- (M on `targets`) psi, in place, one streaming pass over HBM.
 * Replaces linalg/transformations.py:105-172 (targeted_left_multiply),
 * :310-380 (apply_matrix_to_slices) and the per-gate slice fast paths in
 * ops/*.py::_apply_unitary_.  M need not be unitary (Kraus operators, and the
 * superoperators of the density-matrix path use the same entry point).
 * k = number of targets (1 <= k <= 10).  complex64: k <= 4 register-tiled
 * streaming kernel, k = 5, 6 tcgen05/TMEM tensor-core kernel (3xTF32, fp32
 * accuracy) when n_qubits >= k + 7; complex128: k <= 4 register-tiled kernel.
 * Anything else takes a generic out-of-place kernel that needs `scratch`, a
 * device buffer of the state's size (NULL otherwise). */
int b2q_sv_apply_matrix(void* state, int dtype, int n_qubits, const double* matrix_c128,
                        const int* targets, int k, void* scratch, void* stream);

/* Applies `num_gates` gates in order with one call (one pass each).  ks[g]
 * targets for gate g are consecutive in `targets`; its matrix (complex128,
 * 4^k entries) is consecutive in `matrices_c128`.  Replaces the op loop of
 * sim/simulator_base.py:199-212 for runs of unitary ops. */
int b2q_sv_apply_batch(void* state, int dtype, int n_qubits, int num_gates, const int* ks,
                       const int* targets, const double* matrices_c128, void* scratch,
                       void* stream);

/* Applies `num_blocks` (1 or 2) fused blocks IN ORDER with ONE pass over HBM
 * (complex64, each k <= 5): a CTA stages a tile of 2^13 amplitudes spanning the
 * union of the blocks' targets in shared memory, applies every block to it on
 * the tensor cores (3xTF32, fp32 accuracy) and writes it back once — half the
 * HBM traffic per block of b2q_sv_apply_batch.  ks / targets / matrices_c128 as
 * in b2q_sv_apply_batch.  The union of all targets together with index bits 0,
 * 1 and 2 must not exceed 13 bits (always true for two blocks) and n_qubits >= 13: b2q_tile_blocks_feasible
 * (host-only, returns 1/0) tells; otherwise B2Q_ERR_INVALID.  Same reference
 * call sites as b2q_sv_apply_batch (sim/simulator_base.py:199-212 over
 * linalg/transformations.py:105-172). */
int b2q_sv_apply_tile_blocks(void* state, int dtype, int n_qubits, int num_blocks, const int* ks,
                             const int* targets, const double* matrices_c128, void* stream);
int b2q_tile_blocks_feasible(int dtype, int n_qubits, int num_blocks, const int* ks,
                             const int* targets);

/* psi[i] <- diag[bits of i at `targets`] * psi[i]  (diag: complex128[2^k],
 * first target = MSB).  Replaces the diagonal slice fast paths
 * ops/common_gates.py:658-669 (Z), :1072-1083 (CZ) and
 * ops/fourier_transform.py:137-146 (PhaseGradient).  1 <= k <= 16. */
int b2q_sv_apply_diagonal(void* state, int dtype, int n_qubits, const double* diag_c128,
                          const int* targets, int k, void* stream);

/* ---- state vector: reductions / read-out --------------------------------- */

/* *out = sum |psi_i|^2 (float64 accumulation). np.linalg.norm()**2 of
 * sim/state_vector_simulator.py:134 and state_vector_simulation_state.py:231. */
int b2q_sv_norm2(const void* state, int dtype, int n_qubits, double* out, void* stream);

/* out[j] = psi[indices[j]], out complex128[count].  Replaces the fancy-index
 * gather of sim/state_vector_simulator.py:95-98 (compute_amplitudes). */
int b2q_sv_gather(const void* state, int dtype, int n_qubits, const uint64_t* indices,
                  uint64_t count, double* out_c128, void* stream);

/* probs_out[v] = sum over i with (bits of i at `bits`, first = MSB of v) == v
 * of |psi_i|^2, float64[2^m], NOT normalised.  Replaces
 * sim/simulation_utils.py:24-65 (state_probabilities_by_indices) fused with
 * the |psi|^2 pass of sim/state_vector.py:220.  m <= 24. `probs_dev` is a
 * caller-owned device buffer of 2^m doubles; `probs_host` (may be NULL)
 * receives a copy. */
int b2q_sv_marginal_probs(const void* state, int dtype, int n_qubits, const int* bits, int m,
                          double* probs_dev, double* probs_host, void* stream);

/* Draws `reps` basis states from |psi|^2 by inverse CDF: sample j is the
 * smallest index i with cdf(i) > uniforms[j] * total, the rule of
 * numpy RandomState.choice (cdf.searchsorted(u, side='right')) used at
 * sim/state_vector.py:226.  `uniforms_dev`: device float64[reps] in [0,1).
 * `out_indices_dev`: device uint64[reps].  `workspace` from
 * b2q_sv_sample_workspace_bytes().  Replaces sim/state_vector.py:170-232 for
 * the all-qubits case; arbitrary subsets/orders go through
 * b2q_unpack_bits on the sampled indices. */
int b2q_sv_sample(const void* state, int dtype, int n_qubits, const double* uniforms_dev,
                  uint64_t reps, uint64_t* out_indices_dev, void* workspace,
                  uint64_t workspace_bytes, void* stream);
uint64_t b2q_sv_sample_workspace_bytes(int n_qubits, uint64_t reps);

/* Inverse-CDF draw from a small explicit distribution held on the device
 * (float64 probs_dev[count], need not be normalised): out_indices_dev[j] =
 * searchsorted(cumsum(p)/sum(p), uniforms[j], 'right').  One CTA. Used for
 * marginals (measurement of a few qubits, density-matrix diagonals). */
int b2q_cdf_sample(const double* probs_dev, uint64_t count, const double* uniforms_dev,
                   uint64_t reps, uint64_t* out_indices_dev, void* stream);

/* out[j*m + q] = bit `bits[q]` of indices[j] (uint8 0/1): the big-endian digit
 * loop of sim/state_vector.py:228-232 (value/digits.py:139-200) on device. */
int b2q_unpack_bits(const uint64_t* indices_dev, uint64_t reps, const int* bits, int m,
                    uint8_t* out_dev, void* stream);

/* Projects onto bits==values at `bits` and renormalises: amplitudes whose
 * bits differ are zeroed, the rest scaled by 1/sqrt(prob).  Replaces the mask
 * write + divide of sim/state_vector.py:300-318. */
int b2q_sv_collapse(void* state, int dtype, int n_qubits, const int* bits, const int* values,
                    int m, double prob, void* stream);

/* <psi| P |psi> for a Pauli string given as x_mask/z_mask over bit positions
 * (Y = both; the (-i)^{#Y} phase is applied inside).  Writes re, im.
 * Replaces ops/pauli_string.py:625-655. */
int b2q_sv_pauli_expectation(const void* state, int dtype, int n_qubits, uint64_t x_mask,
                             uint64_t z_mask, double* out_re_im, void* stream);

/* `count` Pauli strings that share one x_mask (they flip the same bits) in ONE pass
 * over the state, 16 strings per launch: out_re_im[2t], [2t+1] = <psi| P_t |psi>.
 * A PauliSum's terms grouped by x_mask (all Z-type terms of a cost Hamiltonian have
 * x_mask = 0) cost one read of the state per group instead of one per term
 * (sim/sparse_simulator.py:193-218 evaluates ops/pauli_string.py:625-655 per term).
 * Sums are float64; synchronises the stream. */
int b2q_sv_pauli_expectation_multi(const void* state, int dtype, int n_qubits, uint64_t x_mask,
                                   const uint64_t* z_masks, int count, double* out_re_im,
                                   void* stream);

/* Reduced density matrix of m <= 5 qubits, the rest traced out:
 * out[a * 2^m + b] = sum_rest psi[a, rest] conj(psi[b, rest]) as 4^m complex128
 * values on the host, bits[0] = most significant bit of a and b.  Replaces
 * qis/states.py:623-693 (density_matrix_from_state_vector) behind
 * StateVectorMixin.density_matrix_of / bloch_vector_of (sim/state_vector.py:109-167),
 * which the reference evaluates on a host copy of the state (and refuses above
 * 25 qubits).  One read of the state at any m (m >= 3 on >= 11 qubits: Gram
 * products of shared-memory tiles, partial sums added in a fixed order, so two
 * calls return identical bits).  Synchronises the stream. */
int b2q_sv_reduced_density_matrix(const void* state, int dtype, int n_qubits, const int* bits,
                                  int m, double* out_c128, void* stream);

/* ---- state layout: Kronecker product, axis permutation, factoring ---------- */

/* out[(i << nb) | j] = a[i] * b[j]: linalg/transformations.py:603-613
 * (state_vector_kronecker_product), used when unentangled sub-states are
 * joined (sim/simulation_product_state.py:110-117) and for ancilla qubits. */
int b2q_sv_kron(const void* a, int na, const void* b, int nb, int dtype, void* out, void* stream);
/* out[o] = in[i] with bit k of o == bit src_bit[k] of i (out-of-place):
 * np.moveaxis of linalg/transformations.py:743-754
 * (transpose_state_vector_to_axis_order) in bit-position form. */
int b2q_sv_permute_bits(const void* in, void* out, int dtype, int n_qubits, const int* src_bit,
                        void* stream);
/* The same permutation IN PLACE (no second buffer): state[o] <- state[i] with bit k
 * of o == bit src_bit[k] of i, as a short sequence of passes that each permute at
 * most 13 index bits (12 for complex128) inside 64 KB shared-memory tiles;
 * *passes_out (may be NULL) receives their number.  What lets states above 30
 * qubits keep the product-state form of sim/simulation_product_state.py:68-81
 * (the final merge transposes) and puts relabelled SWAP gates back. */
int b2q_sv_permute_bits_inplace(void* state, int dtype, int n_qubits, const int* src_bit,
                                int* passes_out, void* stream);
/* Index of the largest |amplitude| (first one on ties): the pivot of
 * factor_state_vector (linalg/transformations.py:677). */
int b2q_sv_argmax_abs(const void* state, int dtype, int n_qubits, uint64_t* index_out,
                      void* stream);
/* *ok_out = np.allclose(kron(a, b), t, atol, rtol): the validation step of
 * factor_state_vector (:684-690). */
int b2q_sv_kron_allclose(const void* a, int na, const void* b, int nb, const void* t, int dtype,
                         double atol, double rtol, int* ok_out, void* stream);

/* *ok_out = np.allclose(a, b, atol, rtol) for two states of n_qubits: true iff
 * |a[i] - b[i]| <= atol + rtol * |b[i]| for every i (compared in float64; NaNs
 * compare as close, unlike numpy).  Synchronises the stream. */
int b2q_sv_allclose(const void* a, const void* b, int dtype, int n_qubits, double atol,
                    double rtol, int* ok_out, void* stream);

/* ---- density matrix (rho as a 2n-qubit vector) ---------------------------- */

/* out = Tr_rest(rho) over every qubit whose column bit is not in `keep_bits`
 * (keep_bits[0] = most significant qubit of the k-qubit result): the
 * partial_trace calls of factor_density_matrix (linalg/transformations.py:694-727). */
int b2q_dm_partial_trace(const void* rho, int dtype, int n_qubits, const int* keep_bits, int k,
                         void* out, void* stream);

/* probs_dev[i] = Re rho[i,i], float64[2^n]: sim/density_matrix_utils.py:185-192. */
int b2q_dm_diagonal(const void* rho, int dtype, int n_qubits, double* probs_dev, void* stream);
/* out_dev[v] = sum of probs_dev[i] over i whose bits at `bits` (first = MSB of
 * v) equal v; probs_dev: float64[2^n], out_dev: float64[2^m].  The marginal
 * step of sim/density_matrix_utils.py:185-192 (`_probs` ->
 * state_probabilities_by_indices) on a diagonal already extracted. */
int b2q_probs_marginal(const double* probs_dev, int n_qubits, const int* bits, int m,
                       double* out_dev, void* stream);
/* *out = Re trace(rho). */
int b2q_dm_trace(const void* rho, int dtype, int n_qubits, double* out, void* stream);
/* tr(rho P) for a Pauli string given as x_mask/z_mask over the n_qubits bit
 * positions (same convention as b2q_sv_pauli_expectation); reads 2^n entries of
 * rho.  Replaces ops/pauli_string.py:734-770
 * (_expectation_from_density_matrix_no_validation), which
 * sim/density_matrix_simulator.py:204-235 evaluates on a host copy of rho. */
int b2q_dm_pauli_expectation(const void* rho, int dtype, int n_qubits, uint64_t x_mask,
                             uint64_t z_mask, double* out_re_im, void* stream);
/* Zeroes rows and columns whose measured bits differ from `values`, divides
 * by prob: sim/density_matrix_utils.py:167-180. */
int b2q_dm_collapse(void* rho, int dtype, int n_qubits, const int* bits, const int* values, int m,
                    double prob, void* stream);

/* ---- sharded state vector: global<->local qubit swap ---------------------- */

/* Packs, for each of the 2^g combinations of the `g` local bit positions
 * `local_bits`, the sub-block of the shard having those bits == combination
 * into a contiguous segment of `packed` (segment c holds 2^(n_local-g)
 * amplitudes in index order).  After an all-to-all of the segments,
 * b2q_dist_unpack writes segment c back to the positions with local_bits ==
 * c.  Together with the exchange this realises the global<->local qubit
 * swap of DESIGN.md §multi-GPU (no reference counterpart; closest is the
 * index-only SWAP of sim/simulation_product_state.py:95-108). */
int b2q_dist_pack(const void* shard, int dtype, int n_local, const int* local_bits, int g,
                  void* packed, void* stream);
int b2q_dist_unpack(void* shard, int dtype, int n_local, const int* local_bits, int g,
                    const void* packed, void* stream);

/* Shard memory shared between the per-GPU processes of one node (CUDA IPC):
 * b2q_dist_alloc = cudaMalloc; b2q_dist_ipc_get fills a 64-byte handle that the
 * other ranks turn into a peer pointer with b2q_dist_ipc_open. */
int b2q_dist_alloc(uint64_t bytes, void** out_ptr);
int b2q_dist_free(void* ptr);
int b2q_dist_ipc_get(void* ptr, unsigned char* handle64);
int b2q_dist_ipc_open(const unsigned char* handle64, void** out_ptr);
int b2q_dist_ipc_close(void* ptr);

/* Global<->local qubit swap, one kernel per rank over NVLink peer memory: the
 * caller's shard `mine` (global bit value `my_global_bit_value`) and the
 * partner's shard `peer` (the other value) exchange, for half of the index
 * range each, the amplitudes whose local bit `local_bit` differs from the
 * owner's global bit.  Both ranks of a pair must call it between two barriers.
 * local_bit >= 1 for complex64 (16-byte vectors). */
int b2q_dist_swap_bit(void* mine, void* peer, int dtype, int n_local, int local_bit,
                      int my_global_bit_value, void* stream);

/* Multi-bit exchange: the m (1..3) local bits `local_bits` (ascending) trade places
 * with m global bits in ONE kernel per rank.  `my_sub_rank` = this rank's values of
 * the exchanged global bits (bit i pairs with local_bits[i]); `peers[c]` = the shard
 * (peer-mapped) of the rank whose exchanged global bits read c and whose other rank
 * bits equal this rank's (entry [my_sub_rank] is ignored).  Each rank keeps the
 * sub-block whose local bits read my_sub_rank and swaps every other sub-block c
 * with sub-block my_sub_rank of peers[c]: (1 - 2^-m) of the shard out and in per
 * rank instead of m halves.  In place; all 2^m ranks of a group call it between
 * two barriers.  local bits >= 1 for complex64 (16-byte vectors). */
int b2q_dist_swap_bits(void* mine, void* const* peers, int dtype, int n_local,
                       const int* local_bits, int m, int my_sub_rank, void* stream);

/* One 4- or 5-qubit block (complex64) applied to `shard_in` and, in the same
 * kernel, the global<->local exchange of b2q_dist_swap_bit: results whose index
 * bit `local_bit` equals this rank's value of the global bit are written to
 * `out_local`, the others into the partner's buffer `out_peer` (peer memory, with
 * that bit set to this rank's value).  Out of place: both ranks of a pair call it
 * with their spare buffers, between two stream-ordered barriers; afterwards the
 * spare buffers hold the state.  The HBM pass hides behind the NVLink transfer. */
int b2q_dist_apply_exchange(const void* shard_in, void* out_local, void* out_peer, int dtype,
                            int n_local, const double* matrix_c128, const int* targets, int k,
                            int local_bit, int my_global_bit_value, void* stream);

/* ---- batched Monte-Carlo trajectories ---------------------------------------
 * 2^batch_bits independent n_qubits-qubit states stored back to back (trajectory
 * t = index bits [n_qubits, n_qubits + batch_bits)).  Unitary gates are ordinary
 * b2q_sv_apply_* calls on the (n_qubits + batch_bits)-bit array; the calls below
 * are the per-trajectory part of the stochastic operations that the reference
 * performs once per repetition (sim/simulator_base.py:249-264).  `*_dev` arrays
 * are device pointers with one entry per trajectory. */

/* psi_t <- scale_t * M[choice_t] psi_t on `targets` (k <= 4, count <= 65536): the chosen unitary
 * of a mixture (sim/state_vector_simulation_state.py:183-203) or the chosen Kraus
 * operator with its 1/sqrt(weight) (:205-257).  matrices_c128 = count row-major
 * 2^k x 2^k complex128 matrices on the host; scale_dev may be NULL (= 1);
 * trajectories whose choice == skip_index are not touched (pass -1 for none). */
int b2q_bsv_apply_select(void* state, int dtype, int n_qubits, int batch_bits,
                         const double* matrices_c128, int count, const int* targets, int k,
                         const int* choice_dev, const double* scale_dev, int skip_index,
                         void* stream);
/* m selections of 1-qubit operators in one launch: operator j acts on targets[j],
 * trajectory t applies matrices[choice_dev[j * 2^batch_bits + t]] there, in the
 * order j = 0..m-1 (m <= 32; targets may repeat).  A noise model's layer of
 * identical 1-qubit mixtures after a moment (sim/simulator_base.py:196 ->
 * noise_model.noisy_moments) is one call.  matrices_c128 = count row-major 2 x 2
 * complex128 matrices. */
int b2q_bsv_apply_select_multi(void* state, int dtype, int n_qubits, int batch_bits,
                               const double* matrices_c128, int count, const int* targets, int m,
                               const int* choice_dev, int skip_index, void* stream);
/* weights_dev[t * count + i] = || K_i psi_t ||^2 (float64): the trial weights of
 * the Kraus loop of sim/state_vector_simulation_state.py:228-245, all operators
 * and all trajectories in one read-only pass. */
int b2q_bsv_kraus_weights(const void* state, int dtype, int n_qubits, int batch_bits,
                          const double* matrices_c128, int count, const int* targets, int k,
                          double* weights_dev, void* stream);
/* psi_t[i] <- scale_t * psi_t[i] if (i & mask) == pattern_t else 0: the
 * measurement collapse of sim/state_vector.py:300-318 with one outcome per
 * trajectory. */
int b2q_bsv_collapse(void* state, int dtype, int n_qubits, int batch_bits, uint64_t mask,
                     const uint64_t* pattern_dev, const double* scale_dev, void* stream);

/* ---- host-side scheduler helper --------------------------------------------- */

/* block <- (matrix on row-index bits bitpos[]) . block on the HOST: block is a
 * 2^u x 2^u complex128 row-major matrix (u <= 6), matrix 2^k x 2^k complex128 with
 * its first wire as the most significant index bit, bitpos[j] the row-index bit of
 * wire j.  The gate fuser's block product (the host part of
 * transformers/merge_k_qubit_gates.py:70-114, whose merged unitary the reference
 * obtains from protocols.unitary(CircuitOperation)); no GPU involved. */
int b2q_host_left_apply(double* block, int u, const double* matrix, const int* bitpos, int k);

/* out <- M_{n-1} ... M_1 M_0: the product of `num_members` small matrices, each embedded
 * on its wires of a u-wire space (u <= 6; ks / bitpos / matrices_c128 consecutive per
 * member, conventions of b2q_host_left_apply), starting from the identity.  One call per
 * emitted block of the gate fuser; no GPU involved. */
int b2q_host_compose(double* out, int u, int num_members, const int* ks, const int* bitpos,
                     const double* matrices_c128);

/* ---- a recorded schedule executed by one call --------------------------------
 * The device operations a circuit's unitary prefix issues from |0...0> (basis states,
 * Kronecker joins, gate passes, in-place permutations, scalings: what
 * SimulationProductState + the scheduler do to the sub-states,
 * cirq-core/cirq/sim/simulation_product_state.py:83-139), recorded once per circuit
 * (cirq_b200/plan_cache.py) and replayed without the interpreter in the loop.
 * `slots[i]` = device buffer of state i (caller-allocated, 2^bits amplitudes each);
 * `ints` / `reals` hold the operations' arguments at ints_offset / reals_offset;
 * `permute_passes` (num_slots counters, or NULL) receives the passes the in-place
 * permutations took, per slot. */
enum {
  B2Q_OP_BASIS = 0,    /* slot <- |basis_index> of n_bits */
  B2Q_OP_KRON = 1,     /* slot <- slots[a] (x) slots[b]; ints: bits of a, bits of b */
  B2Q_OP_DENSE = 2,    /* b2q_sv_apply_batch: ints = ks[count] then the targets; reals = matrices */
  B2Q_OP_TILE = 3,     /* b2q_sv_apply_tile_blocks: same layout, count = 1 or 2 */
  B2Q_OP_DIAGONAL = 4, /* b2q_sv_apply_diagonal: ints = targets[count]; reals = 2^count entries */
  B2Q_OP_SCALE = 5,    /* reals = re, im */
  B2Q_OP_PERMUTE = 6   /* b2q_sv_permute_bits_inplace: ints = src_bit[n_bits] */
};
typedef struct {
  int32_t kind;
  int32_t slot;
  int32_t a, b;
  int32_t n_bits; /* bits of slots[slot] */
  int32_t count;
  int64_t ints_offset;
  int64_t reals_offset;
  uint64_t basis_index;
} b2q_schedule_op;
int b2q_schedule_op_bytes(void);
int b2q_run_schedule(int dtype, int num_ops, const b2q_schedule_op* ops, const int* ints,
                     const double* reals, int num_slots, void* const* slots, int* permute_passes,
                     void* stream);

/* out[i] = product over the members of diag_m[bits of i at member m's wires]: the table of
 * a diagonal block (u <= 16 wires) from the list of its diagonal gates; host only. */
int b2q_host_compose_diag(double* out, int u, int num_members, const int* ks, const int* bitpos,
                          const double* diags_c128);

/* ---- tuning knobs and host-only test hooks (not needed by a binding) -------- */

/* How target bits inside the 512-byte warp zone are handled by the register
 * kernel: 0 = warp shuffles, 1 = lane remap, 2 = measured per-target policy
 * (default). */
int b2q_set_lane_mode(int mode);
/* complex64 access width of the register kernel: 0 = policy (default), 1 = always
 * 16-byte vectors, 2 = always one amplitude per lane. */
int b2q_set_vec_mode(int mode);
/* Tensor-core (tcgen05) kernels for complex64 blocks: 0 = off, 1 = k = 5 and 6
 * (default: the scope allows tensor cores for k >= 5 only), 2 = also k = 4 (an
 * experiment: under the board's power cap the 4-qubit tensor-core pass measured
 * 2.84 ms at 30 qubits against 3.58 ms for the FP32 kernel, profiles/README.md r1x). */
int b2q_set_tc_mode(int mode);
/* Shared-memory staging of the tensor-core kernels (coalesced HBM access for any
 * target positions): 0 = never, 1 = when a target sits on index bit 0 or 1
 * (default), 2 = always. */
int b2q_set_tc_stage_mode(int mode);
/* Staged kernel pipeline experiments (profiles/README.md, r1p): early = 1 issues the
 * next region's copy at the top of an iteration instead of after the operands are in
 * TMEM (default 0), l2_ahead > 0 adds an L2 prefetch that many iterations ahead
 * (default 0; both measured slower under the power cap). */
int b2q_set_tc_stage_opts(int early, int l2_ahead);
/* Host-only: the register kernel's plan for a target set (see
 * tests/test_plan_host.py for the layout of `out`, 24 ints) and the matrix
 * permutation to sorted-target order; no GPU needed. */
int b2q_debug_plan(int dtype, int n_qubits, const int* targets, int k, int* out);
int b2q_debug_permute_matrix(const double* matrix_c128, const int* targets, int k, double* out);
/* Host-only: the passes of b2q_sv_permute_bits_inplace (per pass 27 ints: count |
 * tile bits[13] | local source bit of each local output bit[13]). */
int b2q_debug_permute_plan(int dtype, int n_qubits, const int* src_bit, int max_passes, int* out,
                           int* passes_out);
/* Host-only: address tables of the tile kernel (b2q_sv_apply_tile_blocks) for
 * `num_blocks` blocks of 5 ascending targets each (layout of `out`: see
 * tests/test_plan_host.py). */
int b2q_debug_tile_plan(int n_qubits, int num_blocks, const int* sorted_targets, int64_t* out);
/* Host-only: address tables of the staged tensor-core kernel for ascending
 * targets (layout of `out`, 86 int64: see tests/test_plan_host.py). */
int b2q_debug_tc_stage_plan(int n_qubits, const int* sorted_targets, int k, int64_t* out);
/* Host-only: tile plan of the one-read reduced-density-matrix kernel for 3-5 kept
 * bits of a register of >= 11 qubits (layout of `out`, 16 + 2 * 2048 int64: tile bits,
 * kept-bit ranks, then offset and shared-memory slot of every tile element). */
int b2q_debug_rdm_plan(int n_qubits, const int* bits, int m, int64_t* out);
/* Host-only: launch shape of b2q_sv_pauli_expectation_multi (out[4]: run kernel used,
 * virtual threads, runs per virtual thread, CTAs). */
int b2q_debug_pauli_plan(int n_qubits, int64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* CIRQ_B200_H_ */
