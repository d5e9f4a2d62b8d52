"""Synthetic workloads of BASELINE.json's configs, built with Cirq's own
generators (so the reference and this backend see identical circuits), plus
the conversion of a unitary circuit to the (matrix, bits) gate list the C-ABI
consumes."""
from __future__ import annotations

from typing import Sequence

import numpy as np


def rqc_circuit(rows: int, cols: int, depth: int, seed: int = 1):
    """Sycamore-style random circuit (config 2): random sqrt(X/Y/W) rotations
    between FSim(pi/2, pi/6) interaction layers on a rows x cols grid —
    cirq.experiments.random_rotations_between_grid_interaction_layers_circuit
    (experiments/random_quantum_circuit_generation.py:538-622) with the Sycamore
    gate (cirq-google/cirq_google/ops/sycamore_gate.py:27-50)."""
    from cirq_b200._cirq_compat import import_cirq

    cirq = import_cirq()
    qubits = cirq.GridQubit.rect(rows, cols)
    circuit = cirq.experiments.random_rotations_between_grid_interaction_layers_circuit(
        qubits,
        depth=depth,
        two_qubit_op_factory=lambda a, b, _: cirq.FSimGate(np.pi / 2, np.pi / 6)(a, b),
        seed=seed,
    )
    return circuit, sorted(qubits)


def qft_circuit(n: int):
    """n-qubit generalisation of examples/quantum_fourier_transform.py:44-72
    (config 3): H, then CZ**(2^-k) + SWAP ladders on a line."""
    from cirq_b200._cirq_compat import import_cirq

    cirq = import_cirq()
    q = cirq.LineQubit.range(n)
    ops = []
    for r in range(n - 1, 0, -1):
        ops.append(cirq.H(q[0]))
        for i in range(r):
            ops.append(cirq.CZ(q[i], q[i + 1]) ** (2.0 ** -(i + 1)))
            ops.append(cirq.SWAP(q[i], q[i + 1]))
    ops.append(cirq.H(q[0]))
    return cirq.Circuit(ops, strategy=cirq.InsertStrategy.EARLIEST), q


def random_circuit(n: int, depth: int = 20, seed: int = 1234):
    """cirq.testing.random_circuit (config 1 / 4), testing/random_circuit.py:49-125."""
    from cirq_b200._cirq_compat import import_cirq

    cirq = import_cirq()
    q = cirq.LineQubit.range(n)
    return cirq.testing.random_circuit(q, depth, 0.9, random_state=seed), q


def circuit_to_gates(circuit, qubit_order: Sequence) -> list[tuple[np.ndarray, list[int]]]:
    """[(unitary, bit positions)] of a unitary circuit; bit = n-1-axis."""
    from cirq_b200._cirq_compat import import_cirq

    cirq = import_cirq()
    n = len(qubit_order)
    axis = {q: i for i, q in enumerate(qubit_order)}
    gates = []
    for op in circuit.all_operations():
        if cirq.is_measurement(op):
            continue
        gates.append((cirq.unitary(op), [n - 1 - axis[q] for q in op.qubits]))
    return gates


def builtin_rqc_gates(rows: int, cols: int, depth: int, seed: int = 1):
    """Gate list with the structure of `rqc_circuit`, generated without Cirq
    (only used when Cirq cannot be imported on the box): sqrt(X), sqrt(Y),
    sqrt(W) rotations never repeated on a qubit in consecutive cycles, FSim(pi/2,
    pi/6) on the ABCDCDAB staggered grid pattern."""
    rng = np.random.RandomState(seed)
    n = rows * cols

    def bit(r, c):
        return n - 1 - (r * cols + c)

    sx = np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]]) / 2
    sy = np.array([[1 + 1j, -1 - 1j], [1 + 1j, 1 + 1j]]) / 2
    w = (np.array([[0, 1], [1, 0]]) + np.array([[0, -1j], [1j, 0]])) / np.sqrt(2)
    evals, evecs = np.linalg.eigh(w)
    sw = (evecs * np.sqrt(evals.astype(complex))) @ evecs.conj().T
    singles = [sx, sy, sw]
    theta, phi = np.pi / 2, np.pi / 6
    fsim = np.array(
        [
            [1, 0, 0, 0],
            [0, np.cos(theta), -1j * np.sin(theta), 0],
            [0, -1j * np.sin(theta), np.cos(theta), 0],
            [0, 0, 0, np.exp(-1j * phi)],
        ]
    )
    # (vertical?, offset parity, stagger) for A B C D, order ABCDCDAB
    layers = {
        'A': (True, 0, 0), 'B': (True, 1, 0), 'C': (False, 1, 0), 'D': (False, 0, 0),
    }
    order = 'ABCDCDAB'
    prev = [-1] * n
    gates = []
    for d in range(depth + 1):
        for r in range(rows):
            for c in range(cols):
                q = r * cols + c
                choices = [i for i in range(3) if i != prev[q]]
                pick = choices[rng.randint(len(choices))]
                prev[q] = pick
                gates.append((singles[pick], [bit(r, c)]))
        if d == depth:
            break
        vertical, parity, _ = layers[order[d % 8]]
        for r in range(rows):
            for c in range(cols):
                if vertical and r + 1 < rows and (r + c) % 2 == parity:
                    gates.append((fsim, [bit(r, c), bit(r + 1, c)]))
                if not vertical and c + 1 < cols and (r + c) % 2 == parity:
                    gates.append((fsim, [bit(r, c), bit(r, c + 1)]))
    return gates


def qaoa_circuit(n: int = 16, p: int = 2, graph_seed: int = 0):
    """Parameterised max-cut QAOA (config 5): examples/qaoa.py:128-158
    ``qaoa_max_cut_circuit`` on a random 3-regular graph, symbols beta{i}, gamma{i}.
    Returns (circuit, qubits, symbol names)."""
    import networkx
    import sympy

    from cirq_b200._cirq_compat import import_cirq

    cirq = import_cirq()
    qubits = cirq.LineQubit.range(n)
    graph = networkx.random_regular_graph(3, n, seed=graph_seed)
    betas = [sympy.Symbol(f'beta{i}') for i in range(p)]
    gammas = [sympy.Symbol(f'gamma{i}') for i in range(p)]

    def rzz(rads):
        return cirq.ZZPowGate(exponent=2 * rads / sympy.pi, global_shift=-0.5)

    ops = [cirq.H.on_each(*qubits)]
    for beta, gamma in zip(betas, gammas):
        ops.append([rzz(-0.5 * gamma).on(qubits[i], qubits[j]) for i, j in graph.edges])
        ops.append(cirq.rx(2 * beta).on_each(*qubits))
    ops.append(cirq.measure(*qubits, key='m'))
    names = [s.name for s in betas + gammas]
    return cirq.Circuit(ops), qubits, names


def qaoa_sweep(names, points: int = 256):
    """`points` resolvers zipping a Linspace per symbol (config 5's run_sweep)."""
    from cirq_b200._cirq_compat import import_cirq

    cirq = import_cirq()
    return cirq.Zip(*[cirq.Linspace(name, 0.1 + 0.05 * i, 1.1 + 0.05 * i, points)
                      for i, name in enumerate(names)])
