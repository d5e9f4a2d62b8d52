"""Small-n pass through the kernel families for compute-sanitizer:

    compute-sanitizer --tool memcheck python tools/sanitize_small.py

(register / tensor-core / diagonal gate passes, reductions, sampler, collapse,
layout, in-place permutation, two-block tile, batched-trajectory and
reduced-density-matrix kernels at 12-15 qubits).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from cirq_b200.device_state import DeviceState

    rng = np.random.RandomState(0)

    def unitary(k):
        d = 1 << k
        q, r = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
        return q * (np.diag(r) / np.abs(np.diag(r)))

    for dtype in (np.complex64, np.complex128):
        n = 13
        dev = DeviceState.basis(n, dtype, 0)
        for k, bits in ((1, [0]), (2, [0, 7]), (3, [1, 5, 12]), (4, [2, 3, 9, 11]), (5, [12, 4, 9, 0, 6]),
                        (5, [0, 1, 2, 3, 4])):
            dev.apply_matrix(unitary(k), bits)
        dev.apply_diagonal(np.exp(1j * rng.standard_normal(1 << 12)), [12, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0])
        dev.apply_diagonal(np.exp(1j * rng.standard_normal(4)), [3, 11])
        assert abs(dev.norm2() - 1) < 1e-3
        dev.marginal_probs([3, 11, 0])
        dev.sample_bits(list(range(n - 1, -1, -1)), rng.random_sample(500))
        dev.pauli_expectation(0b101, 0b110)
        dev.reduced_density_matrix([5, 0, 9])  # 13 qubits, 3 kept bits: the one-read Gram kernel
        dev.reduced_density_matrix([12, 1, 7, 3, 10])
        dev.reduced_density_matrix([4, 11])  # the row-tile kernel
        dev.pauli_expectations(0, [0b101, 1 << 12, (1 << 13) - 1])  # run kernel (>= 12 qubits), Z type
        dev.pauli_expectations((1 << 12) | 3, list(range(1, 20)))  # ... with an X mask, two launches
        other = dev.copy()
        assert dev.allclose(other, 1e-6)
        dev.amplitudes([0, 5, 77])
        p = dev.marginal_probs([5])
        dev.collapse([5], [1], p[1] / p.sum())
        # in-place bit permutation (shared-memory tiles) and, for complex64, two blocks in
        # one tile pass (tcgen05 + TMEM, cp.async double buffering)
        dev.permute_bits_inplace(list(range(n))[::-1])
        dev.permute_bits_inplace(rng.permutation(n).tolist())
        if dtype == np.complex64:
            dev.apply_tile_blocks([(unitary(5), [12, 4, 9, 0, 6]), (unitary(3), [7, 1, 11])])
            dev.apply_tile_blocks([(unitary(5), [8, 3, 10, 5, 2])])
            big = DeviceState.basis(15, dtype, 1)
            big.apply_tile_blocks([(unitary(5), [14, 13, 12, 11, 10]), (unitary(5), [9, 8, 7, 6, 5])])
            assert abs(big.norm2() - 1) < 1e-3
        a = DeviceState.basis(4, dtype, 3)
        b = a.kron(DeviceState.basis(3, dtype, 1))
        assert b.kron_allclose(a, DeviceState.basis(3, dtype, 1), 1e-6)
        b.permute_bits([6, 5, 4, 3, 2, 1, 0][::-1])
        # batched trajectories: 2^4 states of 9 qubits
        t = DeviceState.from_numpy(np.ones(16, dtype=dtype), dtype).kron(DeviceState.basis(9, dtype, 0))
        mats = np.stack([np.eye(2), unitary(1), unitary(1)])
        t.bsv_apply_select(9, mats, [4], rng.randint(0, 3, size=16), None, 0)
        t.bsv_apply_select_multi(9, mats, [0, 8, 3], rng.randint(0, 3, size=(3, 16)), 0)
        t.bsv_kraus_weights(9, mats, [2])
        t.bsv_collapse(9, [1, 7], rng.randint(0, 2, size=(16, 2)), np.ones(16))
        rho = DeviceState.basis(12, dtype, 0)  # 6-qubit density matrix
        rho.dm_pauli_expectation(0b11, 0b101)
        rho.dm_trace()
    print('sanitize_small: all kernel families ran')


if __name__ == '__main__':
    main()
