"""Generates tests/golden/*.npz from the UNMODIFIED reference (quantumlib/Cirq).

Run in the build container, where the reference is importable:

    python tests/golden/make_golden.py

Every array is produced by calling the reference's own functions (cited below,
paths relative to cirq-core/cirq/); nothing from ``oracle/`` or ``cirq_b200`` is
used.  The fixtures are small and committed so that the GPU box (which has no
/root/reference) can check both the oracle and the CUDA path against them.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from cirq_b200._cirq_compat import import_cirq  # noqa: E402  (only locates/imports cirq)

cirq = import_cirq()


def rand_state(rng, n, dtype):
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    v /= np.linalg.norm(v)
    return v.astype(dtype)


def rand_matrix(rng, k):
    d = 1 << k
    return rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))


def golden_targeted_left_multiply():
    """linalg/transformations.py:105-172 on random states/matrices/axes."""
    rng = np.random.RandomState(20260101)
    out = {}
    case = 0
    for dtype in (np.complex64, np.complex128):
        for n in (3, 5, 7, 9):
            for k in (1, 2, 3, 4):
                if k > n:
                    continue
                for _ in range(3):
                    axes = rng.permutation(n)[:k].tolist()
                    state = rand_state(rng, n, dtype)
                    mat = rand_matrix(rng, k)
                    res = cirq.targeted_left_multiply(
                        mat.astype(dtype).reshape((2,) * (2 * k)), state.reshape((2,) * n), axes
                    )
                    out[f'c{case}_state'] = state
                    out[f'c{case}_matrix'] = mat
                    out[f'c{case}_axes'] = np.array(axes)
                    out[f'c{case}_n'] = np.array(n)
                    out[f'c{case}_out'] = res.reshape(-1)
                    case += 1
    out['num_cases'] = np.array(case)
    np.savez_compressed(os.path.join(HERE, 'targeted_left_multiply.npz'), **out)
    return case


def circuit_to_gate_list(circuit, qubits):
    """[(unitary complex128, axes)] for a unitary circuit, op by op
    (protocols/unitary_protocol.py:79-145)."""
    qmap = {q: i for i, q in enumerate(qubits)}
    gates = []
    for op in circuit.all_operations():
        gates.append((cirq.unitary(op), [qmap[q] for q in op.qubits]))
    return gates


def golden_simulator():
    """sim/sparse_simulator.py: Simulator.simulate final_state_vector on
    cirq.testing.random_circuit (testing/random_circuit.py:49-125)."""
    out = {}
    case = 0
    for dtype in (np.complex64, np.complex128):
        for n, depth, seed in ((4, 8, 1), (7, 10, 2), (10, 12, 3), (12, 20, 1234)):
            qubits = cirq.LineQubit.range(n)
            circuit = cirq.testing.random_circuit(qubits, depth, 0.9, random_state=seed)
            # every qubit must appear so the default order covers all of them
            circuit.append(cirq.I.on_each(*qubits))
            res = cirq.Simulator(dtype=dtype, seed=1).simulate(circuit, qubit_order=qubits)
            gates = circuit_to_gate_list(circuit, qubits)
            out[f'c{case}_n'] = np.array(n)
            out[f'c{case}_num_gates'] = np.array(len(gates))
            for g, (u, axes) in enumerate(gates):
                out[f'c{case}_g{g}_u'] = u
                out[f'c{case}_g{g}_axes'] = np.array(axes)
            out[f'c{case}_final'] = res.final_state_vector
            out[f'c{case}_json'] = np.array(cirq.to_json(circuit))
            case += 1
    out['num_cases'] = np.array(case)
    np.savez_compressed(os.path.join(HERE, 'simulator_final_states.npz'), **out)
    return case


def golden_sampling():
    """sim/state_vector.py:170-232 sample_state_vector and :235-322
    measure_state_vector with literal seeds; sim/simulation_utils.py:24-65."""
    rng = np.random.RandomState(77)
    out = {}
    case = 0
    for dtype in (np.complex64, np.complex128):
        for n in (1, 3, 6, 9):
            state = rand_state(rng, n, dtype)
            for m in sorted({1, min(2, n), n}):
                indices = rng.permutation(n)[:m].tolist()
                seed = int(rng.randint(1 << 30))
                reps = 64
                bits = cirq.sample_state_vector(state, indices, repetitions=reps, seed=seed)
                uniforms = np.random.RandomState(seed).random_sample(reps)
                probs = cirq.sim.simulation_utils.state_probabilities_by_indices(
                    (state * state.conj()).real, indices, (2,) * n
                )
                mseed = int(rng.randint(1 << 30))
                mbits, collapsed = cirq.measure_state_vector(state, indices, seed=mseed)
                out[f'c{case}_state'] = state
                out[f'c{case}_n'] = np.array(n)
                out[f'c{case}_indices'] = np.array(indices)
                out[f'c{case}_uniforms'] = uniforms
                out[f'c{case}_bits'] = bits
                out[f'c{case}_probs'] = probs
                out[f'c{case}_measure_uniform'] = np.array(
                    np.random.RandomState(mseed).random_sample()
                )
                out[f'c{case}_measure_bits'] = np.array(mbits)
                out[f'c{case}_measure_state'] = collapsed
                case += 1
    out['num_cases'] = np.array(case)
    np.savez_compressed(os.path.join(HERE, 'sampling.npz'), **out)
    return case


def golden_density_matrix():
    """sim/density_matrix_simulator.py final_density_matrix for noisy circuits
    (protocols/apply_channel_protocol.py:168-356), with the op list restated as
    (Kraus operators, axes) so it can be replayed without cirq."""
    out = {}
    case = 0
    for dtype in (np.complex64, np.complex128):
        for n, depth, seed in ((2, 4, 5), (4, 6, 6), (5, 8, 7)):
            qubits = cirq.LineQubit.range(n)
            circuit = cirq.testing.random_circuit(qubits, depth, 0.9, random_state=seed)
            circuit.append(cirq.I.on_each(*qubits))
            noise = cirq.ConstantQubitNoiseModel(cirq.depolarize(0.05))
            noisy = cirq.Circuit(noise.noisy_moments(circuit, sorted(circuit.all_qubits())))
            # add some other channels
            noisy.append(cirq.amplitude_damp(0.1).on(qubits[0]))
            noisy.append(cirq.phase_damp(0.2).on(qubits[-1]))
            if n >= 2:
                noisy.append(cirq.depolarize(0.1, n_qubits=2).on(qubits[0], qubits[1]))
            res = cirq.DensityMatrixSimulator(dtype=dtype, seed=1).simulate(
                noisy, qubit_order=qubits
            )
            qmap = {q: i for i, q in enumerate(qubits)}
            ops = list(noisy.all_operations())
            out[f'c{case}_n'] = np.array(n)
            out[f'c{case}_num_ops'] = np.array(len(ops))
            for g, op in enumerate(ops):
                ks = cirq.kraus(op)
                out[f'c{case}_g{g}_kraus'] = np.array(ks)
                out[f'c{case}_g{g}_axes'] = np.array([qmap[q] for q in op.qubits])
            out[f'c{case}_final'] = res.final_density_matrix
            case += 1
    out['num_cases'] = np.array(case)
    np.savez_compressed(os.path.join(HERE, 'density_matrix_final_states.npz'), **out)
    return case


def golden_pauli():
    """ops/pauli_string.py:548-655 expectation_from_state_vector."""
    rng = np.random.RandomState(5)
    out = {}
    case = 0
    paulis = [cirq.I, cirq.X, cirq.Y, cirq.Z]
    for n in (1, 3, 6, 8):
        qubits = cirq.LineQubit.range(n)
        state = rand_state(rng, n, np.complex128)
        for _ in range(6):
            codes = rng.randint(0, 4, size=n)
            ps = cirq.PauliString({q: paulis[c] for q, c in zip(qubits, codes) if c != 0})
            val = ps.expectation_from_state_vector(state, {q: i for i, q in enumerate(qubits)})
            out[f'c{case}_state'] = state
            out[f'c{case}_n'] = np.array(n)
            out[f'c{case}_codes'] = codes  # per axis: 0 I, 1 X, 2 Y, 3 Z
            out[f'c{case}_value'] = np.array(val)
            case += 1
    out['num_cases'] = np.array(case)
    np.savez_compressed(os.path.join(HERE, 'pauli_expectation.npz'), **out)
    return case


def golden_reference_test_vectors():
    """Known-answer vectors copied from the reference's own tests as data:
    sim/state_vector_test.py:62-94 (big-endian sampling over all index
    permutations) and linalg/transformations_test.py:256-305."""
    out = {}
    # test_sample_state_big_endian: basis state x of 3 qubits sampled on [2,1,0]
    results = []
    for x in range(8):
        state = cirq.to_valid_state_vector(x, 3)
        results.append(cirq.sample_state_vector(state, [2, 1, 0]))
    out['big_endian_samples'] = np.array(results)
    # every permutation of 3 indices on |110>
    import itertools

    perm_results = []
    perms = list(itertools.permutations([0, 1, 2]))
    state = cirq.to_valid_state_vector(6, 3)
    for perm in perms:
        perm_results.append(cirq.sample_state_vector(state, list(perm)))
    out['perms'] = np.array(perms)
    out['perm_samples'] = np.array(perm_results)
    np.savez_compressed(os.path.join(HERE, 'reference_test_vectors.npz'), **out)
    return 2


def golden_reduced_density_matrix():
    """qis/states.py:586-693 density_matrix_from_state_vector / bloch_vector_from_state_vector."""
    rng = np.random.RandomState(11)
    out = {}
    case = 0
    for n in (1, 2, 4, 7, 9):
        state = rand_state(rng, n, np.complex128)
        for m in range(1, min(n, 5) + 1):
            indices = rng.permutation(n)[:m]
            out[f'c{case}_state'] = state
            out[f'c{case}_n'] = np.array(n)
            out[f'c{case}_indices'] = indices  # cirq axes, result index big-endian over them
            out[f'c{case}_rho'] = cirq.density_matrix_from_state_vector(state, [int(i) for i in indices])
            out[f'c{case}_bloch'] = cirq.bloch_vector_from_state_vector(state, int(indices[0]))
            case += 1
    out['num_cases'] = np.array(case)
    np.savez_compressed(os.path.join(HERE, 'reduced_density_matrix.npz'), **out)
    return case


def golden_trajectory_ops():
    """Per-trajectory pieces of the reference's per-repetition loop
    (sim/simulator_base.py:249-264) on B independent states:
    the Kraus trial weights of sim/state_vector_simulation_state.py:228-245
    (targeted_left_multiply + squared norm), the state after the chosen operator
    and its renormalisation (:246-257), and measure_state_vector's collapse
    (sim/state_vector.py:235-322) with the outcome it drew."""
    rng = np.random.RandomState(13)
    out = {}
    case = 0
    channels = [cirq.amplitude_damp(0.3), cirq.depolarize(0.2), cirq.phase_damp(0.4),
                cirq.generalized_amplitude_damp(0.6, 0.25), cirq.depolarize(0.1, n_qubits=2)]
    for n, B in ((1, 4), (3, 8), (6, 4)):
        states = np.stack([rand_state(rng, n, np.complex128) for _ in range(B)])
        for ch in channels:
            k = cirq.num_qubits(ch)
            if k > n:
                continue
            kraus = np.stack(cirq.kraus(ch))
            axes = [int(a) for a in rng.permutation(n)[:k]]
            weights = np.zeros((B, len(kraus)))
            applied = np.zeros((B, len(kraus), 1 << n), dtype=np.complex128)
            for t in range(B):
                psi = states[t].reshape((2,) * n)
                for i, op in enumerate(kraus):
                    res = cirq.linalg.targeted_left_multiply(op.reshape((2,) * (2 * k)), psi, axes)
                    weights[t, i] = np.linalg.norm(res) ** 2
                    applied[t, i] = res.reshape(-1)
            out[f'c{case}_states'] = states
            out[f'c{case}_n'] = np.array(n)
            out[f'c{case}_axes'] = np.array(axes)
            out[f'c{case}_kraus'] = kraus
            out[f'c{case}_weights'] = weights
            out[f'c{case}_applied'] = applied
            # measurement collapse of every trajectory on the same axes
            results, collapsed = [], []
            for t in range(B):
                bits, post = cirq.measure_state_vector(states[t], axes, seed=int(rng.randint(1 << 30)))
                results.append(bits)
                collapsed.append(post)
            out[f'c{case}_results'] = np.array(results)
            out[f'c{case}_collapsed'] = np.stack(collapsed)
            case += 1
    out['num_cases'] = np.array(case)
    np.savez_compressed(os.path.join(HERE, 'trajectory_ops.npz'), **out)
    return case


def golden_dm_pauli():
    """ops/pauli_string.py:657-770 expectation_from_density_matrix on random
    (valid) density matrices."""
    rng = np.random.RandomState(17)
    out = {}
    case = 0
    paulis = [cirq.I, cirq.X, cirq.Y, cirq.Z]
    for n in (1, 2, 4, 6):
        qubits = cirq.LineQubit.range(n)
        a = rand_matrix(rng, n)
        rho = a @ a.conj().T
        rho /= np.trace(rho)
        for _ in range(6):
            codes = rng.randint(0, 4, size=n)
            ps = cirq.PauliString({q: paulis[c] for q, c in zip(qubits, codes) if c != 0})
            val = ps.expectation_from_density_matrix(rho, {q: i for i, q in enumerate(qubits)})
            out[f'c{case}_rho'] = rho
            out[f'c{case}_n'] = np.array(n)
            out[f'c{case}_codes'] = codes  # per axis: 0 I, 1 X, 2 Y, 3 Z
            out[f'c{case}_value'] = np.array(val)
            case += 1
    out['num_cases'] = np.array(case)
    np.savez_compressed(os.path.join(HERE, 'dm_pauli_expectation.npz'), **out)
    return case


if __name__ == '__main__':
    print('cirq', cirq.__version__, cirq.__file__)
    print('targeted_left_multiply cases:', golden_targeted_left_multiply())
    print('simulator cases:', golden_simulator())
    print('sampling cases:', golden_sampling())
    print('density matrix cases:', golden_density_matrix())
    print('pauli cases:', golden_pauli())
    print('reference test vectors:', golden_reference_test_vectors())
    print('reduced density matrix cases:', golden_reduced_density_matrix())
    print('trajectory op cases:', golden_trajectory_ops())
    print('dm pauli cases:', golden_dm_pauli())
