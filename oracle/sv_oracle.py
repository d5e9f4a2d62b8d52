"""CPU ORACLE — test infrastructure, NOT part of the product path.

A plain numpy restatement of the reference's state-evolution arithmetic
(quantumlib/Cirq 1.8.0.dev0, paths relative to ``cirq-core/cirq/``).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; ``cirq_b200`` never does.

The reference is pure Python; its arithmetic lives in numpy (``np.einsum``,
strided-slice ufuncs, ``RandomState.choice``; pinned ``numpy~=2.1`` in
``cirq-core/requirements.txt``), which is not vendored under
``/root/reference``.  Each function below restates what the cited reference
call site computes, in flat-index / bit-position form (the form the CUDA
kernels use) rather than the reference's tensor/einsum form, so that the
comparison is between two independently derived implementations.

Parity is PINNED: ``tests/test_oracle.py`` checks every function here against
golden vectors produced by the unmodified reference in this container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``) and, when the
reference is importable, against the live reference.

Conventions (verified against the reference): qubit axis ``a`` of
an n-qubit state is bit position ``p = n-1-a`` of the flat index (axis 0 = most
significant bit); a k-qubit matrix has its FIRST target as the most significant
bit of its row/column index.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np


# --------------------------------------------------------------------------- helpers


def axes_to_bits(n_qubits: int, axes: Sequence[int]) -> list[int]:
    """Cirq axis -> flat-index bit position (axis 0 is the MSB)."""
    return [n_qubits - 1 - int(a) for a in axes]


def _group_indices(n_qubits: int, targets: Sequence[int]) -> np.ndarray:
    """Index table idx[g, j]: flat index of member j of amplitude group g.

    Member j has the bit of ``targets[q]`` equal to bit ``k-1-q`` of j (first
    target = MSB), the other n-k bits enumerate g.
    """
    k = len(targets)
    total = 1 << n_qubits
    all_idx = np.arange(total, dtype=np.int64)
    mask = 0
    for t in targets:
        mask |= 1 << t
    bases = all_idx[(all_idx & mask) == 0]
    offs = np.zeros(1 << k, dtype=np.int64)
    for j in range(1 << k):
        o = 0
        for q, t in enumerate(targets):
            if (j >> (k - 1 - q)) & 1:
                o |= 1 << t
        offs[j] = o
    return bases[:, None] | offs[None, :]


# --------------------------------------------------------------------------- gates


def apply_matrix(
    state: np.ndarray, n_qubits: int, matrix: np.ndarray, targets: Sequence[int]
) -> np.ndarray:
    """psi <- (M on target BITS) psi; returns a new flat array.

    Restates ``linalg.targeted_left_multiply`` (linalg/transformations.py:105-172,
    ``einsum(M[out,in], psi[...in...]) -> psi[...out...]``) and
    ``apply_matrix_to_slices`` (:310-380) as reached from
    ``_apply_unitary_from_matrix`` (protocols/apply_unitary_protocol.py:440-466),
    including the cast of the matrix to the state dtype (:450).
    """
    state = np.asarray(state).reshape(-1)
    k = len(targets)
    m = np.asarray(matrix).reshape(1 << k, 1 << k).astype(state.dtype)
    idx = _group_indices(n_qubits, targets)
    out = np.empty_like(state)
    out[idx] = state[idx] @ m.T
    return out


def apply_diagonal(
    state: np.ndarray, n_qubits: int, diag: np.ndarray, targets: Sequence[int]
) -> np.ndarray:
    """psi[i] *= diag[bits of i at targets] (first target = MSB).

    Restates the diagonal fast paths ops/common_gates.py:658-669 (Z),
    :1072-1083 (CZ), ops/fourier_transform.py:137-146 (PhaseGradient).
    """
    state = np.asarray(state).reshape(-1)
    i = np.arange(state.size, dtype=np.int64)
    key = np.zeros_like(i)
    for t in targets:
        key = (key << 1) | ((i >> t) & 1)
    return state * np.asarray(diag).astype(state.dtype)[key]


# --------------------------------------------------------------------------- read-out


def norm2(state: np.ndarray) -> float:
    s = np.asarray(state).reshape(-1)
    return float(np.sum((s.real.astype(np.float64)) ** 2 + (s.imag.astype(np.float64)) ** 2))


def marginal_probs(state: np.ndarray, n_qubits: int, bits: Sequence[int]) -> np.ndarray:
    """Unnormalised float64 marginal over ``bits`` (bits[0] = MSB of the key).

    Restates ``|psi|^2`` (sim/state_vector.py:220) followed by
    ``state_probabilities_by_indices`` (sim/simulation_utils.py:24-65: transpose
    measured axes to the front, row-sum) without the final normalisation.
    """
    s = np.asarray(state).reshape(-1)
    p = (s * s.conj()).real.astype(np.float64)
    i = np.arange(s.size, dtype=np.int64)
    key = np.zeros_like(i)
    for b in bits:
        key = (key << 1) | ((i >> b) & 1)
    return np.bincount(key, weights=p, minlength=1 << len(bits))


def choice_indices(probs: np.ndarray, uniforms: np.ndarray) -> np.ndarray:
    """Inverse-CDF draw with numpy ``RandomState.choice`` semantics.

    ``choice(n, size, p)`` (used at sim/state_vector.py:226 and
    sim/density_matrix_utils.py:89) computes ``cdf = p.cumsum(); cdf /= cdf[-1];
    cdf.searchsorted(random_sample(size), side='right')``.
    """
    cdf = np.cumsum(np.asarray(probs, dtype=np.float64))
    cdf /= cdf[-1]
    return np.searchsorted(cdf, np.asarray(uniforms, dtype=np.float64), side='right')


def unpack_bits(indices: np.ndarray, bits: Sequence[int]) -> np.ndarray:
    """uint8[reps, m]: column q is bit ``bits[q]`` of each index.

    Restates the big-endian digit loop of sim/state_vector.py:228-232
    (value/digits.py:139-200).
    """
    idx = np.asarray(indices, dtype=np.uint64)
    out = np.zeros((idx.size, len(bits)), dtype=np.uint8)
    for q, b in enumerate(bits):
        out[:, q] = (idx >> np.uint64(b)) & np.uint64(1)
    return out


def sample(
    state: np.ndarray, n_qubits: int, bits: Sequence[int], uniforms: np.ndarray
) -> np.ndarray:
    """``sample_state_vector`` (sim/state_vector.py:170-232) given the uniforms
    ``RandomState.choice`` would draw: uint8[reps, m] ordered like ``bits``."""
    m = len(bits)
    probs = marginal_probs(state, n_qubits, bits)
    picks = choice_indices(probs, uniforms)
    return unpack_bits(picks, [m - 1 - q for q in range(m)])


def collapse(
    state: np.ndarray, n_qubits: int, bits: Sequence[int], values: Sequence[int], prob: float
) -> np.ndarray:
    """Projects onto bits==values and divides by sqrt(prob) in the state's real
    dtype (sim/state_vector.py:300-318)."""
    s = np.array(state).reshape(-1)
    i = np.arange(s.size, dtype=np.int64)
    keep = np.ones(s.size, dtype=bool)
    for b, v in zip(bits, values):
        keep &= ((i >> b) & 1) == int(v)
    s[~keep] = 0
    real = np.float32 if s.dtype == np.complex64 else np.float64
    s /= np.sqrt(real(prob))
    return s


def pauli_expectation(state: np.ndarray, n_qubits: int, x_mask: int, z_mask: int) -> complex:
    """<psi|P|psi>, P given by x/z bit masks (Y = both).

    Restates ops/pauli_string.py:625-655 (apply each Pauli to a copy, then
    tensordot with psi*), using P|i> = i^{nY} (-1)^{popcount(i&z)} |i^x>.
    """
    s = np.asarray(state).reshape(-1).astype(np.complex128)
    i = np.arange(s.size, dtype=np.int64)
    par = np.zeros(s.size, dtype=np.int64)
    zz = i & z_mask
    while np.any(zz):
        par ^= zz & 1
        zz >>= 1
    sign = 1.0 - 2.0 * par
    ny = bin(x_mask & z_mask).count('1')
    return complex((1j**ny) * np.sum(np.conj(s[i ^ x_mask]) * sign * s))


# --------------------------------------------------------------------------- density matrix


def superoperator(kraus_ops: Sequence[np.ndarray]) -> np.ndarray:
    """sum_i K_i (x) conj(K_i): the matrix acting on (row bits, column bits).

    Restates ``_apply_kraus`` (protocols/apply_channel_protocol.py:297-356:
    out = sum_i K_i rho K_i^dagger, left multiply on row axes by K_i, on column
    axes by conj(K_i)) and ``_apply_unitary`` (:274-294) as one linear map.
    """
    return sum(np.kron(np.asarray(k), np.conj(np.asarray(k))) for k in kraus_ops)


def dm_apply_channel(
    rho: np.ndarray, n_qubits: int, kraus_ops: Sequence[np.ndarray], targets: Sequence[int]
) -> np.ndarray:
    """rho <- sum_i K_i rho K_i^dagger on target BITS of an n-qubit density
    matrix stored flat (row bits above column bits)."""
    tg = [t + n_qubits for t in targets] + list(targets)
    out = apply_matrix(np.asarray(rho).reshape(-1), 2 * n_qubits, superoperator(kraus_ops), tg)
    return out


def dm_diagonal(rho: np.ndarray, n_qubits: int) -> np.ndarray:
    """Re diag(rho) as float64 (sim/density_matrix_utils.py:185-192)."""
    d = 1 << n_qubits
    return np.real(np.asarray(rho).reshape(d, d).diagonal()).astype(np.float64)


def dm_collapse(
    rho: np.ndarray, n_qubits: int, bits: Sequence[int], values: Sequence[int], prob: float
) -> np.ndarray:
    """Zeroes rows/columns off the outcome, divides by prob
    (sim/density_matrix_utils.py:167-180)."""
    d = 1 << n_qubits
    r = np.array(rho).reshape(d, d)
    i = np.arange(d, dtype=np.int64)
    keep = np.ones(d, dtype=bool)
    for b, v in zip(bits, values):
        keep &= ((i >> b) & 1) == int(v)
    r[~keep, :] = 0
    r[:, ~keep] = 0
    real = np.float32 if r.dtype == np.complex64 else np.float64
    r /= real(prob)
    return r.reshape(-1)


# --------------------------------------------------------------------------- sharding


def dist_pack(shard: np.ndarray, n_local: int, local_bits: Sequence[int]) -> np.ndarray:
    """Segment c (local_bits[0] = MSB of c) holds the sub-block of the shard whose
    ``local_bits`` equal c, in index order."""
    s = np.asarray(shard).reshape(-1)
    return s[_group_indices(n_local, local_bits).T.reshape(-1)]


def dist_unpack(packed: np.ndarray, n_local: int, local_bits: Sequence[int]) -> np.ndarray:
    p = np.asarray(packed).reshape(-1)
    out = np.empty_like(p)
    out[_group_indices(n_local, local_bits).T.reshape(-1)] = p
    return out


# --------------------------------------------------------------------------- circuits


def run_gate_list(n_qubits: int, gates, dtype=np.complex64, initial: int = 0) -> np.ndarray:
    """Applies ``gates`` = [(matrix, target_bits), ...] to |initial>.  The CPU
    stand-in for the reference's op loop (sim/simulator_base.py:199-212) used by
    bench.py's cpu_baseline when the reference itself is not importable."""
    psi = np.zeros(1 << n_qubits, dtype=dtype)
    psi[initial] = 1
    for m, tg in gates:
        psi = apply_matrix(psi, n_qubits, m, tg)
    return psi


# ---- batched trajectories (B states of n qubits back to back) ---------------------------
# Per trajectory these are the reference's per-repetition operations:
# apply_mixture (cirq-core/cirq/sim/state_vector_simulation_state.py:183-203),
# apply_channel (:205-257) and the collapse of measure_state_vector
# (cirq-core/cirq/sim/state_vector.py:300-318).


def bsv_apply_select(states, n, matrices, targets, choice, scale=None, skip=-1):
    """states: complex[B * 2^n]; returns the array with trajectory t replaced by
    scale[t] * matrices[choice[t]] applied on `targets` (choice == skip: untouched)."""
    out = np.array(states).reshape(-1, 1 << n)
    for t in range(out.shape[0]):
        c = int(choice[t])
        if c == skip:
            continue
        v = apply_matrix(out[t], n, np.asarray(matrices[c]), list(targets))
        out[t] = v * (1.0 if scale is None else scale[t])
    return out.reshape(-1)


def bsv_kraus_weights(states, n, matrices, targets):
    """float64[B, count]: squared norm of matrices[i] applied to trajectory t."""
    st = np.asarray(states).reshape(-1, 1 << n)
    out = np.zeros((st.shape[0], len(matrices)), dtype=np.float64)
    for t in range(st.shape[0]):
        for i, m in enumerate(matrices):
            v = apply_matrix(st[t].astype(np.complex128), n, np.asarray(m), list(targets))
            out[t, i] = float(np.sum(np.abs(v) ** 2))
    return out


def bsv_collapse(states, n, bits, values, scale):
    """Trajectory t keeps the amplitudes whose `bits` equal values[t] (times scale[t])."""
    out = np.array(states).reshape(-1, 1 << n)
    idx = np.arange(1 << n, dtype=np.int64)
    vals = np.asarray(values).reshape(out.shape[0], len(bits))
    for t in range(out.shape[0]):
        keep = np.ones(1 << n, dtype=bool)
        for i, b in enumerate(bits):
            keep &= ((idx >> int(b)) & 1) == int(vals[t, i])
        out[t] = np.where(keep, out[t] * scale[t], 0)
    return out.reshape(-1)


def reduced_density_matrix(state: np.ndarray, n_qubits: int, bits) -> np.ndarray:
    """rho[a, b] = sum_rest psi[a, rest] conj(psi[b, rest]) over the kept `bits`
    (bits[0] = MSB of a, b): cirq-core/cirq/qis/states.py:676-693 with
    indices = [n-1-b for b in bits]."""
    bits = [int(b) for b in bits]
    m = len(bits)
    psi = np.asarray(state, dtype=np.complex128).reshape((2,) * n_qubits)
    axes = [n_qubits - 1 - b for b in bits]
    rest = [a for a in range(n_qubits) if a not in axes]
    mat = np.transpose(psi, axes + rest).reshape(1 << m, -1)
    return mat @ mat.conj().T


def dm_pauli_expectation(rho: np.ndarray, n_qubits: int, x_mask: int, z_mask: int) -> complex:
    """tr(rho P) with P given by bit masks (Y = both), rho flat (row bits above
    column bits): cirq-core/cirq/ops/pauli_string.py:734-770 in index form,
    P|i> = i^{#Y} (-1)^{popcount(i & z)} |i ^ x>."""
    d = 1 << n_qubits
    m = np.asarray(rho, dtype=np.complex128).reshape(d, d)
    i = np.arange(d, dtype=np.int64)
    sign = 1 - 2 * (np.array([bin(int(v) & z_mask).count('1') for v in i]) & 1)
    ny = bin(x_mask & z_mask).count('1')
    return complex((1j ** ny) * np.sum(m[i, i ^ x_mask] * sign))
