"""Device-resident dense quantum state: the thin host mirror of the C-ABI.

``DeviceState`` owns one torch tensor in HBM (torch is used for allocation,
streams and copies only) and forwards every numeric operation to the sm_100a
kernels in ``libcirq_b200.so`` through ctypes.  It knows nothing about Cirq:
qubits are BIT POSITIONS of the flat index (bit p = n-1-axis).  It replaces
the numpy buffers of the reference's ``_BufferedStateVector``
(cirq-core/cirq/sim/state_vector_simulation_state.py:33-307) and
``_BufferedDensityMatrix`` (sim/density_matrix_simulation_state.py:33-250);
unlike them it needs no second buffer: every kernel is in place.

There is no CPU fallback: without a CUDA device or the built library every
method raises ``B200Error``.
"""
from __future__ import annotations

import ctypes
import os
from typing import Sequence

import numpy as np

from cirq_b200 import _lib
from cirq_b200._lib import B200Error, check


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise B200Error('cirq_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    return torch


def _stream_ptr(torch):
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class DeviceState:
    """complex[2^n_bits] in HBM (a state vector, or a density matrix with
    n_bits = 2 * n_qubits, row bits above column bits)."""

    def __init__(self, n_bits: int, dtype, tensor=None):
        torch = _torch()
        self.n_bits = int(n_bits)
        self.dtype = np.dtype(dtype)
        self.code = _lib.dtype_code(self.dtype)
        self._real = torch.float32 if self.code == _lib.C64 else torch.float64
        if tensor is None:
            nbytes = (1 << self.n_bits) * (8 if self.code == _lib.C64 else 16)
            if self.n_bits > 40 or nbytes > torch.cuda.get_device_properties(0).total_memory:
                # same exception type and wording as numpy's allocation failure in the
                # reference (its tests match on it)
                raise MemoryError(
                    f'Unable to allocate {nbytes} bytes for a {self.n_bits}-bit state in HBM'
                )
            tensor = torch.empty((1 << self.n_bits, 2), dtype=self._real, device='cuda')
        self.tensor = tensor
        self._lib = _lib.load()
        self._scratch = None

    # ------------------------------------------------------------------ creation

    @classmethod
    def basis(cls, n_bits: int, dtype, index: int = 0) -> 'DeviceState':
        st = cls(n_bits, dtype)
        torch = _torch()
        check(
            st._lib.b2q_sv_init_basis(
                st.ptr, st.code, st.n_bits, ctypes.c_uint64(int(index)), _stream_ptr(torch)
            )
        )
        return st

    @classmethod
    def from_numpy(cls, array: np.ndarray, dtype=None) -> 'DeviceState':
        torch = _torch()
        flat = np.ascontiguousarray(np.asarray(array).reshape(-1))
        dtype = np.dtype(dtype or flat.dtype)
        flat = flat.astype(dtype, copy=False)
        n_bits = int(flat.size).bit_length() - 1
        if flat.size != 1 << n_bits:
            raise ValueError(f'state size {flat.size} is not a power of two')
        real = np.float32 if dtype == np.complex64 else np.float64
        host = torch.from_numpy(flat.view(real).reshape(-1, 2).copy())
        st = cls(n_bits, dtype, tensor=host.to('cuda'))
        return st

    def copy(self) -> 'DeviceState':
        return DeviceState(self.n_bits, self.dtype, tensor=self.tensor.clone())

    # ------------------------------------------------------------------ plumbing

    @property
    def ptr(self):
        return ctypes.c_void_p(self.tensor.data_ptr())

    @property
    def nbytes(self) -> int:
        return self.tensor.numel() * self.tensor.element_size()

    def to_numpy(self) -> np.ndarray:
        """Flat complex ndarray (device -> host copy of the whole state)."""
        host = self.tensor.cpu().numpy()
        return host.view(self.dtype).reshape(-1)

    def synchronize(self):
        _torch().cuda.current_stream().synchronize()

    # ------------------------------------------------------------------ gates

    def apply_matrix(self, matrix, bits: Sequence[int]) -> None:
        """psi <- (M on bits) psi.  bits[0] is the MSB of M's index."""
        torch = _torch()
        k = len(bits)
        m = _lib.as_c128_buffer(matrix)
        if m.size != (1 << k) ** 2:
            raise ValueError(f'matrix of size {m.size} does not act on {k} qubits')
        max_fast = 5 if self.code == _lib.C64 else 4
        scratch = ctypes.c_void_p(0)
        if k > max_fast:
            if self._scratch is None:
                self._scratch = torch.empty_like(self.tensor)
            scratch = ctypes.c_void_p(self._scratch.data_ptr())
        check(
            self._lib.b2q_sv_apply_matrix(
                self.ptr, self.code, self.n_bits, m.ctypes.data, _lib.int_array(bits), k, scratch,
                _stream_ptr(torch),
            )
        )

    # ---- tile passes: two consecutive blocks in ONE pass over HBM -------------------------
    # (b2q_sv_apply_tile_blocks; complex64 states of at least TILE_MIN_BITS bits —
    # below that a pass is not HBM-bound and pairing buys nothing)
    TILE_MIN_BITS = 22

    def tile_pairing(self) -> bool:
        return (self.code == _lib.C64 and self.n_bits >= self.TILE_MIN_BITS
                and os.environ.get('CIRQ_B200_TILE_PAIRS', '1') != '0'
                and os.environ.get('CIRQ_B200_TC_MODE', '1') != '0')  # (the tile pass is a tensor-core kernel)

    def _pairable(self, m, b) -> bool:
        return np.ndim(m) == 2 and len(b) <= 5

    def plan_passes(self, gates: Sequence[tuple]) -> list[list[tuple]]:
        """Groups [(matrix, bits), ...] into kernel launches, in order: each inner
        list is ONE pass over the state — a single block, or two consecutive dense
        blocks of <= 5 bits each (their targets plus index bits 0-2 always fit one
        13-bit tile)."""
        gates = list(gates)
        if not self.tile_pairing():
            return [[g] for g in gates]
        # index bits that must be tile bits besides the targets: 0-2 (runs of >= 64 bytes)
        # lets ANY two 5-bit blocks share a pass; asking for bits 0-3 (128-byte runs: 5.7
        # instead of 8.3 ms for the worst pair) pairs fewer blocks and measured slower on
        # the bench circuit (1881 against 1937 gates/s, profiles/README.md r2e)
        low = set(range(int(os.environ.get('CIRQ_B200_TILE_RUN_BITS', '3'))))
        out, i = [], 0
        while i < len(gates):
            if (i + 1 < len(gates) and self._pairable(*gates[i]) and self._pairable(*gates[i + 1])
                    and len(set(gates[i][1]) | set(gates[i + 1][1]) | low) <= 13):
                out.append([gates[i], gates[i + 1]])
                i += 2
            else:
                out.append([gates[i]])
                i += 1
        return out

    def split_unpaired_tail(self, gates: Sequence[tuple]):
        """(apply now, hold back): while more blocks are still being scheduled, a
        trailing dense block without a partner is worth keeping for the next batch."""
        gates = list(gates)
        if not gates or not self.tile_pairing():
            return gates, []
        passes = self.plan_passes(gates)
        if len(passes[-1]) == 1 and self._pairable(*passes[-1][0]):
            return gates[:-1], gates[-1:]
        return gates, []

    def apply_tile_blocks(self, group: Sequence[tuple]) -> None:
        """The blocks of `group` (1 or 2), in order, in one pass over HBM."""
        torch = _torch()
        ks = [len(b) for _, b in group]
        targets = [int(t) for _, b in group for t in b]
        mats = np.concatenate([_lib.as_c128_buffer(m).reshape(-1) for m, _ in group])
        check(
            self._lib.b2q_sv_apply_tile_blocks(
                self.ptr, self.code, self.n_bits, len(group), _lib.int_array(ks),
                _lib.int_array(targets), mats.ctypes.data, _stream_ptr(torch),
            )
        )

    def lower_batch(self, gates: Sequence[tuple]) -> list[tuple]:
        """The library calls `apply_batch` makes for [(matrix, bits), ...], in order:
        ('tile', [two dense blocks]) = one tile pass, ('dense', [blocks]) = one
        b2q_sv_apply_batch call (a pass per block), ('diag', (entries, bits)).  Needs
        only n_bits / dtype of `self` (cirq_b200/program.py lowers recorded schedules
        with it, so a compiled schedule launches exactly what a live call does)."""
        out: list[tuple] = []

        def singles(run):
            dense: list = []
            for m, b in run:
                if np.ndim(m) == 1:
                    if dense:
                        out.append(('dense', dense))
                        dense = []
                    out.append(('diag', (m, b)))
                else:
                    dense.append((m, b))
            if dense:
                out.append(('dense', dense))

        if self.tile_pairing() and len(gates) > 1:
            run: list = []
            for group in self.plan_passes(gates):
                if len(group) == 2:
                    singles(run)
                    run = []
                    out.append(('tile', group))
                else:
                    run.append(group[0])
            singles(run)
        else:
            singles(gates)
        return out

    def apply_batch(self, gates: Sequence[tuple]) -> None:
        """Applies [(matrix, bits), ...] in order; consecutive dense blocks travel
        two per pass where the tile kernel applies (`plan_passes`)."""
        if not gates:
            return
        for what, payload in self.lower_batch(gates):
            if what == 'tile':
                self.apply_tile_blocks(payload)
            elif what == 'dense':
                self._apply_dense_run(payload)
            else:
                self.apply_diagonal(*payload)

    def _apply_dense_run(self, gates: Sequence[tuple]) -> None:
        """Dense blocks, one pass each, with one library call."""
        torch = _torch()
        if not gates:
            return
        max_fast = 5 if self.code == _lib.C64 else 4
        if any(len(b) > max_fast for _, b in gates):
            for m, b in gates:
                self.apply_matrix(m, b)
            return
        ks = [len(b) for _, b in gates]
        targets = [int(t) for _, b in gates for t in b]
        mats = np.concatenate([_lib.as_c128_buffer(m).reshape(-1) for m, _ in gates])
        check(
            self._lib.b2q_sv_apply_batch(
                self.ptr, self.code, self.n_bits, len(gates), _lib.int_array(ks),
                _lib.int_array(targets), mats.ctypes.data, ctypes.c_void_p(0), _stream_ptr(torch),
            )
        )

    def apply_diagonal(self, diag, bits: Sequence[int]) -> None:
        torch = _torch()
        d = _lib.as_c128_buffer(diag).reshape(-1)
        check(
            self._lib.b2q_sv_apply_diagonal(
                self.ptr, self.code, self.n_bits, d.ctypes.data, _lib.int_array(bits), len(bits),
                _stream_ptr(torch),
            )
        )

    def scale(self, factor: complex) -> None:
        torch = _torch()
        factor = complex(factor)
        check(
            self._lib.b2q_sv_scale(
                self.ptr, self.code, self.n_bits, factor.real, factor.imag, _stream_ptr(torch)
            )
        )

    # ------------------------------------------------------------------ read-out

    def norm2(self) -> float:
        torch = _torch()
        out = ctypes.c_double(0.0)
        check(self._lib.b2q_sv_norm2(self.ptr, self.code, self.n_bits, ctypes.byref(out), _stream_ptr(torch)))
        return out.value

    def amplitudes(self, indices: Sequence[int]) -> np.ndarray:
        torch = _torch()
        idx = np.ascontiguousarray(np.asarray(indices, dtype=np.uint64).reshape(-1))
        out = np.empty(idx.size, dtype=np.complex128)
        check(
            self._lib.b2q_sv_gather(
                self.ptr, self.code, self.n_bits, idx.ctypes.data, idx.size, out.ctypes.data,
                _stream_ptr(torch),
            )
        )
        return out

    def marginal_probs_device(self, bits: Sequence[int]):
        """Unnormalised float64 marginal over `bits` (bits[0] = MSB), on device."""
        torch = _torch()
        m = len(bits)
        probs = torch.empty(1 << m, dtype=torch.float64, device='cuda')
        check(
            self._lib.b2q_sv_marginal_probs(
                self.ptr, self.code, self.n_bits, _lib.int_array(bits), m,
                ctypes.c_void_p(probs.data_ptr()), ctypes.c_void_p(0), _stream_ptr(torch),
            )
        )
        return probs

    def marginal_probs(self, bits: Sequence[int]) -> np.ndarray:
        return self.marginal_probs_device(bits).cpu().numpy()

    def sample_indices_device(self, uniforms: np.ndarray):
        """Inverse-CDF draw of full basis-state indices; uniforms in [0, 1)."""
        torch = _torch()
        u = np.ascontiguousarray(np.asarray(uniforms, dtype=np.float64).reshape(-1))
        reps = u.size
        u_dev = torch.from_numpy(u).to('cuda')
        out = torch.empty(max(reps, 1), dtype=torch.int64, device='cuda')
        ws_bytes = int(self._lib.b2q_sv_sample_workspace_bytes(self.n_bits, reps))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device='cuda')
        check(
            self._lib.b2q_sv_sample(
                self.ptr, self.code, self.n_bits, ctypes.c_void_p(u_dev.data_ptr()), reps,
                ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(ws.data_ptr()), ws_bytes,
                _stream_ptr(torch),
            )
        )
        return out[:reps]

    def sample_indices(self, uniforms: np.ndarray) -> np.ndarray:
        return self.sample_indices_device(uniforms).cpu().numpy().astype(np.uint64)

    @staticmethod
    def cdf_sample_device(probs_dev, uniforms: np.ndarray):
        """searchsorted(cumsum(p)/sum(p), u, 'right') on the device."""
        torch = _torch()
        lib = _lib.load()
        u = np.ascontiguousarray(np.asarray(uniforms, dtype=np.float64).reshape(-1))
        reps = u.size
        u_dev = torch.from_numpy(u).to('cuda')
        out = torch.empty(max(reps, 1), dtype=torch.int64, device='cuda')
        check(
            lib.b2q_cdf_sample(
                ctypes.c_void_p(probs_dev.data_ptr()), probs_dev.numel(),
                ctypes.c_void_p(u_dev.data_ptr()), reps, ctypes.c_void_p(out.data_ptr()),
                _stream_ptr(torch),
            )
        )
        return out[:reps]

    @staticmethod
    def unpack_bits_device(indices_dev, bits: Sequence[int]):
        """uint8[reps, m] on device: column q = bit bits[q] of each index."""
        torch = _torch()
        lib = _lib.load()
        reps = indices_dev.numel()
        m = len(bits)
        out = torch.empty((reps, m), dtype=torch.uint8, device='cuda')
        if reps and m:
            check(
                lib.b2q_unpack_bits(
                    ctypes.c_void_p(indices_dev.data_ptr()), reps, _lib.int_array(bits), m,
                    ctypes.c_void_p(out.data_ptr()), _stream_ptr(torch),
                )
            )
        return out

    def sample_bits(self, bits: Sequence[int], uniforms: np.ndarray,
                    out_columns: Sequence[int] | None = None) -> np.ndarray:
        """uint8[reps, len(bits)] drawn from |psi|^2; mirrors sample_state_vector.

        All bits in natural order -> hierarchical sampler on the full state
        (same CDF order as the reference); up to 24 bits -> marginal in the
        requested order (again the reference's CDF order); otherwise full-state
        draw followed by bit extraction (same distribution).

        `out_columns` (a permutation of range(len(bits))): column c of the result
        is the measured bit bits[out_columns[c]] — the draw itself (and hence a
        seeded result) is that of the order `bits`, only the columns are emitted
        in another order, on the device.
        """
        bits = [int(b) for b in bits]
        m = len(bits)
        u = np.asarray(uniforms, dtype=np.float64).reshape(-1)
        if m == 0 or u.size == 0:
            return np.zeros((u.size, m), dtype=np.uint8)
        cols = list(range(m)) if out_columns is None else [int(c) for c in out_columns]
        natural = bits == list(range(self.n_bits - 1, -1, -1))
        if natural or m > 24:
            idx = self.sample_indices_device(u)
            return self._download(self.unpack_bits_device(idx, [bits[c] for c in cols]))
        probs = self.marginal_probs_device(bits)
        idx = self.cdf_sample_device(probs, u)
        return self._download(self.unpack_bits_device(idx, [m - 1 - c for c in cols]))

    @staticmethod
    def _download(dev_tensor) -> np.ndarray:
        """Device -> host through pinned memory for large results (1M x 30 sample bits:
        a pageable copy runs at a third of the PCIe rate); small ones take the plain path."""
        torch = _torch()
        if dev_tensor.numel() * dev_tensor.element_size() < (1 << 20):
            return dev_tensor.cpu().numpy()
        host = torch.empty(dev_tensor.shape, dtype=dev_tensor.dtype, pin_memory=True)
        host.copy_(dev_tensor, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host.numpy()  # (a view: the pinned block lives as long as the result does)

    def collapse(self, bits: Sequence[int], values: Sequence[int], prob: float) -> None:
        torch = _torch()
        check(
            self._lib.b2q_sv_collapse(
                self.ptr, self.code, self.n_bits, _lib.int_array(bits), _lib.int_array(values),
                len(bits), float(prob), _stream_ptr(torch),
            )
        )

    def pauli_expectation(self, x_mask: int, z_mask: int) -> complex:
        torch = _torch()
        out = (ctypes.c_double * 2)()
        check(
            self._lib.b2q_sv_pauli_expectation(
                self.ptr, self.code, self.n_bits, ctypes.c_uint64(int(x_mask)),
                ctypes.c_uint64(int(z_mask)), out, _stream_ptr(torch),
            )
        )
        return complex(out[0], out[1])

    def pauli_expectations(self, x_mask: int, z_masks: Sequence[int]) -> np.ndarray:
        """complex128[len(z_masks)]: <psi|P_t|psi> for Pauli strings sharing `x_mask`,
        one pass over the state per 16 strings."""
        torch = _torch()
        z = np.ascontiguousarray(np.asarray([int(v) for v in z_masks], dtype=np.uint64))
        out = np.empty(2 * z.size, dtype=np.float64)
        check(
            self._lib.b2q_sv_pauli_expectation_multi(
                self.ptr, self.code, self.n_bits, ctypes.c_uint64(int(x_mask)), z.ctypes.data, int(z.size),
                out.ctypes.data, _stream_ptr(torch),
            )
        )
        return out[0::2] + 1j * out[1::2]

    def reduced_density_matrix(self, bits: Sequence[int]) -> np.ndarray:
        """complex128[2^m, 2^m]: the state with every bit not in `bits` traced out
        (bits[0] = most significant index bit of the result), m <= 5."""
        torch = _torch()
        m = len(bits)
        out = np.empty((1 << m, 1 << m), dtype=np.complex128)
        check(
            self._lib.b2q_sv_reduced_density_matrix(
                self.ptr, self.code, self.n_bits, _lib.int_array(list(bits)), m, out.ctypes.data,
                _stream_ptr(torch),
            )
        )
        return out

    # ------------------------------------------------------------------ layout

    def kron(self, other: 'DeviceState') -> 'DeviceState':
        """|self> (x) |other>: self's bits become the high bits."""
        torch = _torch()
        out = DeviceState(self.n_bits + other.n_bits, self.dtype)
        check(
            self._lib.b2q_sv_kron(
                self.ptr, self.n_bits, other.ptr, other.n_bits, self.code, out.ptr,
                _stream_ptr(torch),
            )
        )
        return out

    # ---- batched trajectories: 2^(n_bits - n_qubits) states of n_qubits back to back ----

    @staticmethod
    def _upload_async(array: np.ndarray):
        """Host array -> device tensor without waiting for the GPU: staged in
        pinned memory (torch's caching host allocator keeps it alive until the copy
        has run) and copied stream-ordered.  A pageable `.to('cuda')` would block
        the host until every queued pass has finished, once per stochastic layer."""
        torch = _torch()
        return torch.from_numpy(np.ascontiguousarray(array)).pin_memory().to('cuda', non_blocking=True)

    def bsv_apply_select(self, n_qubits: int, matrices: np.ndarray, bits: Sequence[int],
                         choice: np.ndarray, scale: np.ndarray | None = None, skip: int = -1) -> None:
        """psi_t <- scale[t] * matrices[choice[t]] psi_t on `bits` for every
        trajectory t (choice[t] == skip: untouched)."""
        torch = _torch()
        m = np.ascontiguousarray(matrices, dtype=np.complex128)
        k = len(bits)
        count = m.shape[0]
        assert m.shape == (count, 1 << k, 1 << k)
        batch = self.n_bits - n_qubits
        c = np.ascontiguousarray(np.asarray(choice, dtype=np.int32).reshape(-1))
        assert c.size == 1 << batch
        c_dev = self._upload_async(c)
        s_dev = None
        if scale is not None:
            s_dev = self._upload_async(np.asarray(scale, dtype=np.float64).reshape(-1))
        check(
            self._lib.b2q_bsv_apply_select(
                self.ptr, self.code, n_qubits, batch, m.ctypes.data, count, _lib.int_array(list(bits)), k,
                ctypes.c_void_p(c_dev.data_ptr()),
                ctypes.c_void_p(s_dev.data_ptr() if s_dev is not None else 0), int(skip), _stream_ptr(torch),
            )
        )

    def bsv_apply_select_multi(self, n_qubits: int, matrices: np.ndarray, bits: Sequence[int],
                               choices: np.ndarray, skip: int = -1) -> None:
        """For j in order: psi_t <- matrices[choices[j, t]] psi_t on bits[j] (1-qubit
        operators; choices[j, t] == skip: untouched), all in one launch."""
        torch = _torch()
        m = np.ascontiguousarray(matrices, dtype=np.complex128)
        count = m.shape[0]
        assert m.shape == (count, 2, 2)
        batch = self.n_bits - n_qubits
        c = np.ascontiguousarray(np.asarray(choices, dtype=np.int32).reshape(len(bits), 1 << batch))
        c_dev = self._upload_async(c)
        check(
            self._lib.b2q_bsv_apply_select_multi(
                self.ptr, self.code, n_qubits, batch, m.ctypes.data, count, _lib.int_array(list(bits)),
                len(bits), ctypes.c_void_p(c_dev.data_ptr()), int(skip), _stream_ptr(torch),
            )
        )

    def bsv_kraus_weights(self, n_qubits: int, matrices: np.ndarray, bits: Sequence[int]) -> np.ndarray:
        """float64[trajectories, count]: || matrices[i] psi_t ||^2."""
        torch = _torch()
        m = np.ascontiguousarray(matrices, dtype=np.complex128)
        k = len(bits)
        count = m.shape[0]
        assert m.shape == (count, 1 << k, 1 << k)
        batch = self.n_bits - n_qubits
        out = torch.empty((1 << batch, count), dtype=torch.float64, device='cuda')
        check(
            self._lib.b2q_bsv_kraus_weights(
                self.ptr, self.code, n_qubits, batch, m.ctypes.data, count, _lib.int_array(list(bits)), k,
                ctypes.c_void_p(out.data_ptr()), _stream_ptr(torch),
            )
        )
        return out.cpu().numpy()

    def bsv_collapse(self, n_qubits: int, bits: Sequence[int], values: np.ndarray, scale: np.ndarray) -> None:
        """Projects trajectory t onto bits == values[t] (values[t, i] = value of
        bits[i]) and multiplies it by scale[t]."""
        torch = _torch()
        batch = self.n_bits - n_qubits
        vals = np.asarray(values, dtype=np.uint64).reshape(1 << batch, len(bits))
        mask = 0
        pattern = np.zeros(1 << batch, dtype=np.uint64)
        for i, b in enumerate(bits):
            mask |= 1 << int(b)
            pattern |= vals[:, i] << np.uint64(int(b))
        p_dev = self._upload_async(pattern.view(np.int64))
        s_dev = self._upload_async(np.asarray(scale, dtype=np.float64).reshape(-1))
        check(
            self._lib.b2q_bsv_collapse(
                self.ptr, self.code, n_qubits, batch, ctypes.c_uint64(mask),
                ctypes.c_void_p(p_dev.data_ptr()), ctypes.c_void_p(s_dev.data_ptr()), _stream_ptr(torch),
            )
        )

    def kron_into(self, other: 'DeviceState', out: 'DeviceState') -> 'DeviceState':
        """out <- |self> (x) |other> for an existing state of the right size (the
        IPC-shared shard of the sharded path)."""
        if out.n_bits != self.n_bits + other.n_bits or out.dtype != self.dtype:
            raise ValueError('kron_into: output state has the wrong shape or dtype')
        torch = _torch()
        check(
            self._lib.b2q_sv_kron(
                self.ptr, self.n_bits, other.ptr, other.n_bits, self.code, out.ptr,
                _stream_ptr(torch),
            )
        )
        return out

    def copy_into(self, out: 'DeviceState') -> 'DeviceState':
        if out.n_bits != self.n_bits or out.dtype != self.dtype:
            raise ValueError('copy_into: output state has the wrong shape or dtype')
        out.tensor.copy_(self.tensor)
        return out

    def permute_bits(self, src_bit: Sequence[int]) -> 'DeviceState':
        """New state with out[o] = self[i], bit k of o == bit src_bit[k] of i."""
        torch = _torch()
        if self.n_bits >= 20:
            # A copy followed by one or two in-place tile passes (each at the copy rate)
            # beats the gather kernel (2.7 TB/s); the planner tells how many it takes.
            plan = (ctypes.c_int * 27)()
            count = ctypes.c_int(0)
            check(self._lib.b2q_debug_permute_plan(self.code, self.n_bits, _lib.int_array(src_bit), 1, plan,
                                                   ctypes.byref(count)))
            if count.value <= 2:
                out = self.copy()
                if count.value:
                    out.permute_bits_inplace(src_bit)
                return out
        out = DeviceState(self.n_bits, self.dtype)
        check(
            self._lib.b2q_sv_permute_bits(
                self.ptr, out.ptr, self.code, self.n_bits, _lib.int_array(src_bit),
                _stream_ptr(torch),
            )
        )
        return out

    def permute_bits_inplace(self, src_bit: Sequence[int]) -> int:
        """self[o] <- self[i], bit k of o == bit src_bit[k] of i, without a second
        buffer (tile passes in shared memory); returns the number of passes."""
        torch = _torch()
        passes = ctypes.c_int(0)
        check(
            self._lib.b2q_sv_permute_bits_inplace(
                self.ptr, self.code, self.n_bits, _lib.int_array(src_bit), ctypes.byref(passes),
                _stream_ptr(torch),
            )
        )
        return int(passes.value)

    def argmax_abs(self) -> int:
        torch = _torch()
        out = ctypes.c_uint64(0)
        check(
            self._lib.b2q_sv_argmax_abs(
                self.ptr, self.code, self.n_bits, ctypes.byref(out), _stream_ptr(torch)
            )
        )
        return int(out.value)

    def slice_copy(self, start: int, n_bits: int) -> 'DeviceState':
        """Copy of the 2^n_bits amplitudes starting at `start`."""
        return DeviceState(n_bits, self.dtype, tensor=self.tensor[start : start + (1 << n_bits)].clone())

    def kron_allclose(self, a: 'DeviceState', b: 'DeviceState', atol: float, rtol: float = 1e-5) -> bool:
        """np.allclose(kron(a, b), self, atol=atol, rtol=rtol) on the device."""
        torch = _torch()
        ok = ctypes.c_int(0)
        check(
            self._lib.b2q_sv_kron_allclose(
                a.ptr, a.n_bits, b.ptr, b.n_bits, self.ptr, self.code, float(atol), float(rtol),
                ctypes.byref(ok), _stream_ptr(torch),
            )
        )
        return bool(ok.value)

    def allclose(self, other: 'DeviceState', atol: float, rtol: float = 1e-5) -> bool:
        """np.allclose(self, other, atol=atol, rtol=rtol) on the device."""
        torch = _torch()
        ok = ctypes.c_int(0)
        check(
            self._lib.b2q_sv_allclose(
                self.ptr, other.ptr, self.code, self.n_bits, float(atol), float(rtol),
                ctypes.byref(ok), _stream_ptr(torch),
            )
        )
        return bool(ok.value)

    def dm_partial_trace(self, keep_bits: Sequence[int]) -> 'DeviceState':
        """Reduced density matrix on the qubits whose column bits are `keep_bits`
        (keep_bits[0] = most significant qubit of the result)."""
        torch = _torch()
        k = len(keep_bits)
        out = DeviceState(2 * k, self.dtype)
        check(
            self._lib.b2q_dm_partial_trace(
                self.ptr, self.code, self.n_bits // 2, _lib.int_array(keep_bits), k, out.ptr,
                _stream_ptr(torch),
            )
        )
        return out

    # ------------------------------------------------------------------ density matrix view

    def dm_diagonal_device(self):
        torch = _torch()
        n = self.n_bits // 2
        probs = torch.empty(1 << n, dtype=torch.float64, device='cuda')
        check(
            self._lib.b2q_dm_diagonal(
                self.ptr, self.code, n, ctypes.c_void_p(probs.data_ptr()), _stream_ptr(torch)
            )
        )
        return probs

    @staticmethod
    def probs_marginal_device(probs_dev, n_qubits: int, bits: Sequence[int]):
        """Marginal of a float64[2^n] device vector over `bits` (bits[0] = MSB)."""
        torch = _torch()
        lib = _lib.load()
        m = len(bits)
        out = torch.empty(1 << m, dtype=torch.float64, device='cuda')
        check(
            lib.b2q_probs_marginal(
                ctypes.c_void_p(probs_dev.data_ptr()), int(n_qubits), _lib.int_array(bits), m,
                ctypes.c_void_p(out.data_ptr()), _stream_ptr(torch),
            )
        )
        return out

    def dm_pauli_expectation(self, x_mask: int, z_mask: int) -> complex:
        """tr(rho P) for this 2n-bit array read as an n-qubit density matrix."""
        torch = _torch()
        out = (ctypes.c_double * 2)()
        check(
            self._lib.b2q_dm_pauli_expectation(
                self.ptr, self.code, self.n_bits // 2, ctypes.c_uint64(int(x_mask)),
                ctypes.c_uint64(int(z_mask)), out, _stream_ptr(torch),
            )
        )
        return complex(out[0], out[1])

    def dm_trace(self) -> float:
        torch = _torch()
        out = ctypes.c_double(0.0)
        check(self._lib.b2q_dm_trace(self.ptr, self.code, self.n_bits // 2, ctypes.byref(out), _stream_ptr(torch)))
        return out.value

    def dm_collapse(self, bits: Sequence[int], values: Sequence[int], prob: float) -> None:
        torch = _torch()
        check(
            self._lib.b2q_dm_collapse(
                self.ptr, self.code, self.n_bits // 2, _lib.int_array(bits),
                _lib.int_array(values), len(bits), float(prob), _stream_ptr(torch),
            )
        )

    def dm_apply_channel(self, kraus_ops, bits: Sequence[int]) -> None:
        """rho <- sum_i K_i rho K_i^dagger as ONE pass with the superoperator
        sum_i K_i (x) conj(K_i) on (row bits, column bits)."""
        n = self.n_bits // 2
        sup = None
        for k in kraus_ops:
            k = np.asarray(k, dtype=np.complex128)
            term = np.kron(k, np.conj(k))
            sup = term if sup is None else sup + term
        self.apply_matrix(sup, [b + n for b in bits] + list(bits))

    # ------------------------------------------------------------------ sharding

    def dist_pack(self, local_bits: Sequence[int], packed) -> None:
        torch = _torch()
        check(
            self._lib.b2q_dist_pack(
                self.ptr, self.code, self.n_bits, _lib.int_array(local_bits), len(local_bits),
                ctypes.c_void_p(packed.data_ptr()), _stream_ptr(torch),
            )
        )

    def dist_unpack(self, local_bits: Sequence[int], packed) -> None:
        torch = _torch()
        check(
            self._lib.b2q_dist_unpack(
                self.ptr, self.code, self.n_bits, _lib.int_array(local_bits), len(local_bits),
                ctypes.c_void_p(packed.data_ptr()), _stream_ptr(torch),
            )
        )
