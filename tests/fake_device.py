"""TEST-ONLY stand-in for cirq_b200.device_state.DeviceState backed by the CPU
oracle.  It lets the host-side simulator logic (queueing, fusion, flush-on-
observe, measurement/sampling plumbing, strategy order) be exercised on a
machine without a GPU.  Never imported by the product."""
from __future__ import annotations

import numpy as np

from oracle import sv_oracle as orc


class FakeTensor(np.ndarray):
    def cpu(self):
        return self

    def numpy(self):
        return np.asarray(self)

    def numel(self):
        return self.size


def _ft(a):
    return np.asarray(a).view(FakeTensor)


class OracleDeviceState:
    def __init__(self, n_bits, dtype, array=None):
        self.n_bits = int(n_bits)
        self.dtype = np.dtype(dtype)
        if array is None and self.n_bits > 34:
            raise MemoryError(f'Unable to allocate a {self.n_bits}-bit state')
        self.array = (
            np.zeros(1 << self.n_bits, dtype=self.dtype) if array is None else np.array(array, dtype=self.dtype)
        )

    @classmethod
    def basis(cls, n_bits, dtype, index=0):
        st = cls(n_bits, dtype)
        st.array[int(index)] = 1
        return st

    @classmethod
    def from_numpy(cls, array, dtype=None):
        flat = np.asarray(array).reshape(-1)
        dtype = np.dtype(dtype or flat.dtype)
        return cls(int(flat.size).bit_length() - 1, dtype, flat.astype(dtype))

    def copy(self):
        return OracleDeviceState(self.n_bits, self.dtype, self.array.copy())

    def to_numpy(self):
        return self.array.copy()

    def synchronize(self):
        pass

    def apply_matrix(self, matrix, bits):
        self.array = orc.apply_matrix(self.array, self.n_bits, np.asarray(matrix), list(bits))

    # tile pairing, as in DeviceState but from 4 bits on so that the CPU tests
    # exercise the host logic around it (held-back blocks, pass grouping)
    TILE_MIN_BITS = 4
    tile_calls = 0

    def tile_pairing(self):
        return self.dtype == np.dtype(np.complex64) and self.n_bits >= self.TILE_MIN_BITS

    def _pairable(self, m, b):
        return np.ndim(m) == 2 and len(b) <= 5

    def plan_passes(self, gates):
        from cirq_b200.device_state import DeviceState

        return DeviceState.plan_passes(self, gates)

    def split_unpaired_tail(self, gates):
        from cirq_b200.device_state import DeviceState

        return DeviceState.split_unpaired_tail(self, gates)

    def apply_batch(self, gates):
        for group in self.plan_passes(gates):
            if len(group) == 2:
                type(self).tile_calls += 1
            for m, b in group:
                if np.ndim(m) == 1:  # a diagonal block
                    self.apply_diagonal(m, b)
                else:
                    self.apply_matrix(m, b)

    def apply_diagonal(self, diag, bits):
        self.array = orc.apply_diagonal(self.array, self.n_bits, diag, list(bits))

    def scale(self, factor):
        self.array = (self.array * self.dtype.type(factor)).astype(self.dtype)

    def norm2(self):
        return orc.norm2(self.array)

    def amplitudes(self, indices):
        return self.array[np.asarray(indices, dtype=np.int64)].astype(np.complex128)

    def marginal_probs_device(self, bits):
        return _ft(orc.marginal_probs(self.array, self.n_bits, list(bits)))

    def marginal_probs(self, bits):
        return orc.marginal_probs(self.array, self.n_bits, list(bits))

    def sample_indices_device(self, uniforms):
        probs = orc.marginal_probs(self.array, self.n_bits, list(range(self.n_bits - 1, -1, -1)))
        return _ft(orc.choice_indices(probs, uniforms).astype(np.int64))

    def sample_indices(self, uniforms):
        return np.asarray(self.sample_indices_device(uniforms)).astype(np.uint64)

    @staticmethod
    def cdf_sample_device(probs_dev, uniforms):
        return _ft(orc.choice_indices(np.asarray(probs_dev), uniforms).astype(np.int64))

    @staticmethod
    def unpack_bits_device(indices_dev, bits):
        return _ft(orc.unpack_bits(np.asarray(indices_dev), list(bits)))

    @staticmethod
    def probs_marginal_device(probs_dev, n_qubits, bits):
        p = np.asarray(probs_dev)
        i = np.arange(p.size, dtype=np.int64)
        key = np.zeros_like(i)
        for b in bits:
            key = (key << 1) | ((i >> b) & 1)
        return _ft(np.bincount(key, weights=p, minlength=1 << len(bits)))

    def sample_bits(self, bits, uniforms, out_columns=None):
        bits = [int(b) for b in bits]
        u = np.asarray(uniforms, dtype=np.float64).reshape(-1)
        if len(bits) == 0 or u.size == 0:
            return np.zeros((u.size, len(bits)), dtype=np.uint8)
        out = orc.sample(self.array, self.n_bits, bits, u)
        return out if out_columns is None else np.ascontiguousarray(out[:, list(out_columns)])

    def collapse(self, bits, values, prob):
        self.array = orc.collapse(self.array, self.n_bits, list(bits), list(values), prob)

    def pauli_expectation(self, x_mask, z_mask):
        return orc.pauli_expectation(self.array, self.n_bits, x_mask, z_mask)

    def pauli_expectations(self, x_mask, z_masks):
        return np.array([orc.pauli_expectation(self.array, self.n_bits, x_mask, int(z)) for z in z_masks],
                        dtype=np.complex128)

    def kron(self, other):
        return OracleDeviceState(self.n_bits + other.n_bits, self.dtype, np.kron(self.array, other.array))

    def bsv_apply_select(self, n_qubits, matrices, bits, choice, scale=None, skip=-1):
        self.array = orc.bsv_apply_select(self.array, n_qubits, np.asarray(matrices), list(bits), choice,
                                          scale, skip).astype(self.dtype)

    def reduced_density_matrix(self, bits):
        return orc.reduced_density_matrix(self.array, self.n_bits, list(bits))

    def bsv_apply_select_multi(self, n_qubits, matrices, bits, choices, skip=-1):
        choices = np.asarray(choices).reshape(len(bits), -1)
        for j, b in enumerate(bits):
            self.bsv_apply_select(n_qubits, matrices, [b], choices[j], None, skip)

    def bsv_kraus_weights(self, n_qubits, matrices, bits):
        return orc.bsv_kraus_weights(self.array, n_qubits, np.asarray(matrices), list(bits))

    def bsv_collapse(self, n_qubits, bits, values, scale):
        self.array = orc.bsv_collapse(self.array, n_qubits, list(bits), values, scale).astype(self.dtype)

    def kron_into(self, other, out):
        out.array[:] = np.kron(self.array, other.array)
        return out

    def copy_into(self, out):
        out.array[:] = self.array
        return out

    def permute_bits(self, src_bit):
        o = np.arange(1 << self.n_bits, dtype=np.int64)
        i = np.zeros_like(o)
        for k, sb in enumerate(src_bit):
            i |= ((o >> k) & 1) << sb
        return OracleDeviceState(self.n_bits, self.dtype, self.array[i])

    def permute_bits_inplace(self, src_bit):
        self.array = self.permute_bits(src_bit).array
        return 1

    def argmax_abs(self):
        return int(np.argmax(np.abs(self.array.astype(np.complex128)) ** 2))

    def slice_copy(self, start, n_bits):
        return OracleDeviceState(n_bits, self.dtype, self.array[start : start + (1 << n_bits)].copy())

    def kron_allclose(self, a, b, atol, rtol=1e-5):
        return bool(np.allclose(np.kron(a.array, b.array), self.array, atol=atol, rtol=rtol))

    def allclose(self, other, atol, rtol=1e-5):
        return bool(np.allclose(self.array, other.array, atol=atol, rtol=rtol))

    def dm_partial_trace(self, keep_bits):
        n = self.n_bits // 2
        k = len(keep_bits)
        rho = self.array.reshape((2,) * (2 * n))
        # axis of column bit b is (2n-1-b); row bit b+n is axis (n-1-b)
        keep_axes = [n - 1 - b for b in keep_bits]
        traced = [a for a in range(n) if a not in keep_axes]
        letters = 'abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOP'
        rows = [letters[a] for a in range(n)]
        cols = [letters[n + a] if a in keep_axes else letters[a] for a in range(n)]
        out = [letters[a] for a in keep_axes] + [letters[n + a] for a in keep_axes]
        res = np.einsum(''.join(rows + cols) + '->' + ''.join(out), rho)
        return OracleDeviceState(2 * k, self.dtype, res.reshape(-1))

    def dm_diagonal_device(self):
        return _ft(orc.dm_diagonal(self.array, self.n_bits // 2))

    def dm_pauli_expectation(self, x_mask, z_mask):
        return orc.dm_pauli_expectation(self.array, self.n_bits // 2, x_mask, z_mask)

    def dm_trace(self):
        return float(orc.dm_diagonal(self.array, self.n_bits // 2).sum())

    def dm_collapse(self, bits, values, prob):
        self.array = orc.dm_collapse(self.array, self.n_bits // 2, list(bits), list(values), prob)
