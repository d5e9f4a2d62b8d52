"""B200 state-vector simulator: a drop-in for ``cirq.Simulator``.

Boundary (reference paths relative to cirq-core/cirq/): the driver classes
``SimulatorBase`` (sim/simulator_base.py:46-352), ``SimulationState``
(sim/simulation_state.py:35-339) and the simulator interfaces
(sim/simulator.py) are imported from Cirq and reused unchanged; this module
replaces what sits below them:

  reference (numpy)                                   here (HBM + sm_100a kernels)
  ------------------------------------------------    ------------------------------------
  _BufferedStateVector                                B200StateVector
    sim/state_vector_simulation_state.py:33-307
  StateVectorSimulationState  (same file :310-447)    B200StateVectorSimulationState
  Simulator / SparseSimulatorStep                     B200Simulator / B200SimulatorStep
    sim/sparse_simulator.py:31-290
  StateVectorTrialResult                              B200StateVectorTrialResult
    sim/state_vector_simulator.py:107-217

Gate application is LAZY: unitary operations are queued as (matrix, bits) and
fused by ``cirq_b200.fusion.GateFuser`` into blocks of up to
``max_fused_qubits`` qubits; the queue is flushed (one GPU pass per block)
whenever the state is observed — measurement, sampling, copy, read-out, a
non-unitary operation.  This keeps ``simulate_moment_steps`` semantics while
still fusing across moments.
"""
from __future__ import annotations

from typing import Any, Iterator, Sequence

import numpy as np

from cirq_b200._cirq_compat import import_cirq
from cirq_b200.device_state import DeviceState
from cirq_b200.fusion import GateFuser, fuser_for

cirq = import_cirq()

from cirq import circuits, ops, protocols, qis, study, value  # noqa: E402
from cirq.sim import simulator, state_vector, state_vector_simulator  # noqa: E402
from cirq.sim.simulation_product_state import SimulationProductState  # noqa: E402
from cirq.sim.simulation_state import SimulationState, strat_act_on_from_apply_decompose  # noqa: E402

# Widest gate whose full unitary is requested from Cirq before falling back to
# the decomposition strategy (mirrors the "<= 4 qubits" shortcut of
# protocols/apply_unitary_protocol.py:375-399, one wider because a 5-qubit
# matrix is still a single streaming pass here).
_MAX_DIRECT_UNITARY_QUBITS = 5
# Product-state form (split_untangled_states) at any size.  A join whose result would
# exceed _DENSE_JOIN_BITS qubits merges EVERY sub-state into one dense state when the
# whole register still fits a GPU (<= _MAX_DENSE_QUBITS: 137 GB complex64): the join
# is out of place, so growing a 33-qubit state by one qubit would need 137 + 69 GB,
# whereas (largest sub-state) x (everything else) needs 137 + 17 GB.  Larger
# registers stay in product form for good (a join that does not fit raises
# MemoryError, as the reference's numpy allocation would).  States above
# _MAX_FACTOR_BITS are not factored back after a measurement (factoring transposes out
# of place).
# from this size on, three or more displaced bits are put back by the in-place
# permutation kernel rather than by SWAP gates
_PERMUTE_RESTORE_MIN_BITS = 20
_DENSE_JOIN_BITS = 31
_MAX_DENSE_QUBITS = 34
_MAX_FACTOR_BITS = 30


class B200StateVector(qis.QuantumStateRepresentation):
    """Device-resident replacement of ``_BufferedStateVector``.

    Holds one complex[2^n] buffer in HBM (no second buffer: kernels are in
    place) plus the queue of not-yet-applied unitaries.
    """

    def __init__(self, dev: DeviceState, num_qubits: int, max_fused_qubits: int | None = None):
        self._dev = dev
        self._n = int(num_qubits)
        self._max_fused = max_fused_qubits
        self._fuser = fuser_for(dev.dtype, max_fused_qubits, self._n, state_vector=True)
        self._qid_shape = (2,) * self._n
        self.passes = 0  # GPU gate passes issued so far (for benchmarks)
        self._since_drain = 0
        self._drain_every = max(8, self._n)
        self._held: list = []  # final blocks kept back by _drain to be paired later
        self._host = None  # cached host copy of the state, dropped on every mutation
        # logical bit -> index bit holding it, after SWAP gates that were relabelled
        # instead of executed and not yet undone (None: every bit in its place)
        self._where: list[int] | None = None

    # ------------------------------------------------------------------ creation

    @classmethod
    def create(
        cls,
        *,
        initial_state: Any = 0,
        qid_shape: tuple[int, ...],
        dtype=np.complex64,
        max_fused_qubits: int | None = None,
    ) -> 'B200StateVector':
        if any(d != 2 for d in qid_shape):
            raise ValueError(
                f'cirq_b200 simulates qubits only (dimension 2); got qid_shape={qid_shape}'
            )
        n = len(qid_shape)
        if isinstance(initial_state, (int, np.integer)):
            # Basis state by big-endian integer, as qis.to_valid_state_vector
            # (qis/states.py:766-832), but built on the device: at 34 qubits the
            # reference's host-side one_hot would be a 137 GB array.
            index = int(initial_state)
            if index < 0 or index >= (1 << n):
                raise ValueError(
                    f'initial_state={index} was out of range for {n} qubits (qid_shape={qid_shape})'
                )
            dev = DeviceState.basis(n, dtype, index)
        else:
            vec = qis.to_valid_state_vector(initial_state, n, qid_shape=qid_shape, dtype=dtype)
            dev = DeviceState.from_numpy(np.asarray(vec).reshape(-1), dtype)
        return cls(dev, n, max_fused_qubits)

    # ------------------------------------------------------------------ queue

    def _bits(self, axes: Sequence[int]) -> list[int]:
        if self._where is None:
            return [self._n - 1 - int(a) for a in axes]
        return [self._where[self._n - 1 - int(a)] for a in axes]

    def queue_unitary(self, matrix: np.ndarray, axes: Sequence[int]) -> None:
        self._host = None
        if len(axes) == 0:
            # a global phase / scalar: fold it into the state
            self.flush()
            self._dev.scale(complex(np.asarray(matrix).reshape(-1)[0]))
            return
        self._fuser.add(matrix, self._bits(axes))
        self._since_drain += 1
        if self._since_drain >= self._drain_every:
            self._drain()

    def _drain(self) -> None:
        """Launches the blocks that can no longer grow, so the GPU works while
        the host keeps scheduling (kernel launches are asynchronous)."""
        self._since_drain = 0
        ready = self._held + self._fuser.pop_final_blocks()
        # (a trailing block without a partner waits for the next batch: two blocks
        # share one pass over HBM, DeviceState.plan_passes)
        ready, self._held = self._dev.split_unpaired_tail(ready)
        if ready:
            self._dev.apply_batch(ready)
            self.passes += len(self._dev.plan_passes(ready))

    def flush(self, restore: bool = True) -> None:
        """Applies the queued gates.  SWAP gates the scheduler relabelled are undone
        with real swaps (`restore`, the default: every reader of the raw array
        expects the canonical bit order) or kept as a bit map (`restore=False`, for
        the readers that can translate bit positions: `mapped_device_state`)."""
        if restore and self._n >= _PERMUTE_RESTORE_MIN_BITS and self._displaced_bits() >= 3:
            # several relabelled SWAPs to put back: one or two in-place permutation
            # passes (pure copies) instead of real SWAP gates fused into dense passes
            self.flush(restore=False)
            if self._where is not None:
                self.passes += self._dev.permute_bits_inplace(self._where)
                self._where = None
                self._host = None
            return
        if restore:
            if self._where is not None:
                # fold the kept map into the fuser's, which then undoes both
                self._fuser.fold_permutation(self._where)
                self._where = None
            blocks = self._fuser.blocks()
        else:
            blocks = self._fuser.blocks(restore=False)
        self._fuser.clear()
        blocks = self._held + blocks
        self._held = []
        if blocks:
            self._host = None
            self._dev.apply_batch(blocks)
            self.passes += len(self._dev.plan_passes(blocks))
        if not restore:
            moved = self._fuser.take_permutation()
            if moved:
                where = self._where if self._where is not None else list(range(self._n))
                self._where = [moved.get(w, w) for w in where]

    def detach_pending(self) -> list:
        """The blocks `flush(restore=False)` would launch, taken off the queue WITHOUT
        launching them (the bit map is updated as if they had run).  For the plan
        cache (cirq_b200/plan_cache.py), which hands them to the replayed state."""
        blocks = self._held + self._fuser.blocks(restore=False)
        self._fuser.clear()
        self._held = []
        moved = self._fuser.take_permutation()
        if moved:
            where = self._where if self._where is not None else list(range(self._n))
            self._where = [moved.get(w, w) for w in where]
        return blocks

    def _displaced_bits(self) -> int:
        """Index bits not holding their own logical bit once the queue is applied
        (kept bit map composed with the fuser's pending SWAP relabelling)."""
        pending = self._fuser._map
        if not pending and self._where is None:
            return 0
        where = self._where if self._where is not None else range(self._n)
        return sum(1 for b, w in enumerate(where) if pending.get(w, w) != b)

    @property
    def device_state(self) -> DeviceState:
        self.flush()
        return self._dev

    def mapped_device_state(self) -> tuple[DeviceState, list[int]]:
        """(device array, where): logical bit b of the state is index bit where[b]
        of the array.  Spares the passes that would put relabelled SWAPs back —
        for readers that address the array by bit position (amplitude gathers,
        Pauli expectations, reduced density matrices)."""
        self.flush(restore=False)
        return self._dev, (self._where if self._where is not None else list(range(self._n)))

    # ------------------------------------------------------------------ QuantumStateRepresentation

    def copy(self, deep_copy_buffers: bool = True) -> 'B200StateVector':
        self.flush()
        out = B200StateVector(self._dev.copy(), self._n, self._max_fused)
        return out

    def measure(self, axes: Sequence[int], seed: 'cirq.RANDOM_STATE_OR_SEED_LIKE' = None) -> list[int]:
        """Projective measurement with collapse, as ``measure_state_vector``
        (sim/state_vector.py:235-322): one ``choice`` draw from the marginal."""
        axes = list(axes)
        if not axes:
            return []
        self.flush()
        self._host = None
        prng = value.parse_random_state(seed)
        bits = self._bits(axes)
        m = len(bits)
        u = float(prng.random_sample())
        if m <= 24:
            probs = self._dev.marginal_probs_device(bits)
            pick = int(DeviceState.cdf_sample_device(probs, np.array([u])).cpu()[0])
            values = [(pick >> (m - 1 - q)) & 1 for q in range(m)]
            p = probs.cpu().numpy()
            prob = float(p[pick] / p.sum())
            self._dev.collapse(bits, values, prob)
        else:
            idx = int(self._dev.sample_indices(np.array([u]))[0])
            values = [(idx >> b) & 1 for b in bits]
            before = self._dev.norm2()
            self._dev.collapse(bits, values, 1.0)
            prob = self._dev.norm2() / before
            self._dev.scale(1.0 / np.sqrt(prob))
        return values

    def sample(
        self,
        axes: Sequence[int],
        repetitions: int = 1,
        seed: 'cirq.RANDOM_STATE_OR_SEED_LIKE' = None,
        out_columns: Sequence[int] | None = None,
    ) -> np.ndarray:
        """``sample_state_vector`` (sim/state_vector.py:170-232) on the device.
        `out_columns` reorders the result's columns on the device (see
        ``DeviceState.sample_bits``)."""
        if repetitions < 0:
            raise ValueError(f'Number of repetitions cannot be negative. Was {repetitions}')
        axes = [int(a) for a in axes]
        for a in axes:
            if a < 0 or a >= self._n:
                raise IndexError(f'Out of range indices in {axes}, must be less than {self._n}')
        if repetitions == 0 or len(axes) == 0:
            return np.zeros(shape=(repetitions, len(axes)), dtype=np.uint8)
        self.flush()
        prng = value.parse_random_state(seed)
        uniforms = prng.random_sample(repetitions)
        return self._dev.sample_bits(self._bits(axes), uniforms, out_columns)

    @property
    def supports_factor(self) -> bool:
        return self._n <= _MAX_FACTOR_BITS

    # ---- layout: Kronecker product, factoring, axis order (split_untangled_states) ----------

    def kron(self, other: 'B200StateVector') -> 'B200StateVector':
        """``state_vector_kronecker_product`` (linalg/transformations.py:603-613)."""
        dev = self.device_state.kron(other.device_state)
        out = B200StateVector(dev, self._n + other._n, self._max_fused)
        out.passes = self.passes + other.passes
        return out

    def reindex(self, axes: Sequence[int]) -> 'B200StateVector':
        """``transpose_state_vector_to_axis_order``: new axis k = old axis axes[k]."""
        axes = [int(a) for a in axes]
        n = self._n
        if axes == list(range(n)):
            return B200StateVector(self.device_state.copy(), n, self._max_fused)
        src_bit = [0] * n
        for k, a in enumerate(axes):
            src_bit[n - 1 - k] = n - 1 - a
        out = B200StateVector(self.device_state.permute_bits(src_bit), n, self._max_fused)
        out.passes = self.passes
        return out

    def reindex_inplace(self, axes: Sequence[int]) -> None:
        """`reindex` without a second buffer and, at first, without moving anything: the
        new axis order is folded into the bit map that relabelled SWAPs already use
        (`_where`).  Readers that address the array by bit position (amplitude gathers,
        Pauli expectations, reduced density matrices) never pay for it; a reader of the
        raw array triggers ONE in-place permutation (b2q_sv_permute_bits_inplace) that
        settles the SWAPs and the axis order together.  For callers that drop the old
        order anyway (``transpose_to_qubit_order(inplace=True)`` at the end of
        ``create_merged_state``, sim/simulation_product_state.py:68-81)."""
        axes = [int(a) for a in axes]
        n = self._n
        if axes == list(range(n)):
            return
        self.flush(restore=False)
        where = self._where if self._where is not None else list(range(n))
        # new logical bit n-1-k is the old logical bit n-1-axes[k]
        self._where = [where[n - 1 - axes[n - 1 - b]] for b in range(n)]
        if self._where == list(range(n)):
            self._where = None
        self._host = None

    def factor(self, axes: Sequence[int], *, validate=True, atol=1e-07):
        """``factor_state_vector`` (linalg/transformations.py:647-691): pivot on the
        largest amplitude, slice the two factors through it, normalise."""
        axes = [int(a) for a in axes]
        n, k = self._n, len(axes)
        rest = [a for a in range(n) if a not in axes]
        t1 = self.reindex(axes + rest)._dev  # factored axes in front
        nr = n - k
        pivot = t1.argmax_abs()
        pf, pr = pivot >> nr, pivot & ((1 << nr) - 1)
        real = np.float32 if t1.dtype == np.complex64 else np.float64
        ext = t1.amplitudes([(e << nr) | pr for e in range(1 << k)]).astype(t1.dtype)
        pivot_amp = ext[pf]
        ext = ext / real(np.linalg.norm(ext))
        extracted = DeviceState.from_numpy(ext, t1.dtype)
        remainder = t1.slice_copy(pf << nr, nr)
        rnorm = np.sqrt(remainder.norm2())
        remainder.scale(1.0 / (complex(rnorm) * complex(pivot_amp) / abs(complex(pivot_amp))))
        if validate:
            if not t1.kron_allclose(extracted, remainder, atol):
                if not np.isclose(np.sqrt(t1.norm2()), 1):
                    raise ValueError('Input state must be normalized.')
                raise cirq.linalg.transformations.EntangledStateError(
                    'The tensor cannot be factored by the requested axes'
                )
        e_state = B200StateVector(extracted, k, self._max_fused)
        r_state = B200StateVector(remainder, nr, self._max_fused)
        r_state.passes = self.passes
        return e_state, r_state

    # ------------------------------------------------------------------ non-unitary helpers

    def apply_matrix_now(self, matrix: np.ndarray, axes: Sequence[int]) -> None:
        self.flush()
        self._host = None
        self._dev.apply_matrix(matrix, self._bits(axes))
        self.passes += 1

    def scale(self, factor: complex) -> None:
        self.flush()
        self._host = None
        self._dev.scale(factor)

    def replace_device_state(self, dev: DeviceState) -> None:
        self._fuser.clear()
        self._held = []
        self._host = None
        self._dev = dev

    def norm2(self) -> float:
        self.flush()
        return self._dev.norm2()

    def to_numpy_tensor(self) -> np.ndarray:
        """Host copy as a ``(2,)*n`` tensor, cached until the state changes."""
        if self._host is None:
            self.flush()
            self._host = self._dev.to_numpy().reshape(self._qid_shape)
        return self._host


class B200ProductState(SimulationProductState):
    """``SimulationProductState`` (sim/simulation_product_state.py:31-181) with a
    sampling fast path: when every requested qubit lives in ONE sub-state — the
    usual case once a circuit has entangled its qubits — the draw is made in that
    sub-state's own qubit order exactly as the reference does (so seeded results
    agree), but the columns are emitted in the requested order by the device
    kernel instead of two host-side shuffles of a (repetitions x qubits) array."""

    def copy(self, deep_copy_buffers: bool = True):
        base = super().copy(deep_copy_buffers)
        return B200ProductState(
            dict(base.sim_states), base.qubits, base.split_untangled_states,
            classical_data=base.classical_data,
        )

    def apply_unitary_op(self, op, unitary: np.ndarray) -> None:
        """``_act_on_fallback_`` (sim/simulation_product_state.py:83-139) for an
        operation already known to be a plain unitary gate: same joins and SWAP
        relabelling, but the matrix goes straight into the joined state's queue
        instead of a second trip through ``protocols.act_on`` (~35 us of dispatch
        per operation, which is what the GPU waits for at the start of a run)."""
        gate = op.gate
        if isinstance(gate, (ops.IdentityGate, ops.SwapPowGate)):
            self._act_on_fallback_(op, op.qubits)
            return
        target = self.join_for(op.qubits)
        target._state.queue_unitary(unitary, target.get_axes(op.qubits))

    def create_merged_state(self):
        """``SimulationProductState.create_merged_state``
        (sim/simulation_product_state.py:68-81).  The reference copies: it joins the
        empty [None] state with every sub-state out of place and transposes the result,
        i.e. needs the final state twice.  Registers above _MAX_FACTOR_BITS qubits are
        merged IN the product state instead — the sub-states are joined for good
        (`join_for`) and the one remaining state is brought to the register's qubit
        order by the in-place permutation — so a 34-qubit state is never duplicated."""
        if not self.split_untangled_states or len(self.qubits) <= _MAX_FACTOR_BITS:
            return super().create_merged_state()
        merged = self.join_for(tuple(self.qubits))
        phase = self._sim_states[None]
        scalar = complex(phase._state.to_numpy_tensor().reshape(-1)[0]) if phase is not None else 1.0
        if scalar != 1.0:  # zero-qubit operations (global phases) live in the [None] state
            merged._state.scale(scalar)
            phase._state.scale(1.0 / scalar)
        return merged.transpose_to_qubit_order(self.qubits, inplace=True)

    def join_for(self, qubits):
        """The sub-state holding all of `qubits`, joining sub-states by Kronecker
        products when they live apart (sim/simulation_product_state.py:110-123).
        A join that would pass _DENSE_JOIN_BITS merges the whole register at once,
        largest sub-state x everything else (see the constants' comment)."""
        states = self._sim_states
        target = states[qubits[0]]
        parts = [target]
        for q in qubits[1:]:
            if not any(states[q] is p for p in parts):
                parts.append(states[q])
        if len(parts) == 1:
            return target
        total = sum(len(p.qubits) for p in parts)
        n_all = len(self.qubits)
        if total > _DENSE_JOIN_BITS and n_all <= _MAX_DENSE_QUBITS:
            everything = []
            for q in self.qubits:
                if not any(states[q] is p for p in everything):
                    everything.append(states[q])
            everything.sort(key=lambda p: -len(p.qubits))
            target, rest = everything[0], everything[1:]
            small = rest[-1]
            for p in reversed(rest[:-1]):  # smallest first: the temporaries stay small
                small.kronecker_product(p, inplace=True)
            target.kronecker_product(small, inplace=True)
        else:
            for p in parts[1:]:
                target.kronecker_product(p, inplace=True)
        for q in target.qubits:
            states[q] = target
        return target

    def _act_on_fallback_(self, action, qubits, allow_decompose: bool = True):
        gate = action if isinstance(action, ops.Gate) else getattr(action, 'gate', None)
        swap = (isinstance(gate, ops.SwapPowGate) and gate.exponent % 2 == 1 and gate.global_shift == 0)
        if len(qubits) > 1 and not swap and not isinstance(gate, ops.IdentityGate):
            self.join_for(tuple(qubits))  # (the reference's joins, with the size policy above)
        return super()._act_on_fallback_(action, qubits, allow_decompose)

    def sample(self, qubits, repetitions: int = 1, seed=None) -> np.ndarray:
        q_set = set(qubits)
        owners = [v for v in dict.fromkeys(self.sim_states.values()) if any(q in q_set for q in v.qubits)]
        if len(owners) == 1 and len(q_set) == len(qubits) and hasattr(owners[0]._state, 'sample'):
            v = owners[0]
            qs = [q for q in v.qubits if q in q_set]
            if len(qs) == len(qubits):
                position = {q: i for i, q in enumerate(qs)}
                try:
                    return v._state.sample(
                        v.get_axes(qs), repetitions, seed,
                        out_columns=[position[q] for q in qubits],
                    )
                except TypeError:
                    pass  # a state representation without the out_columns extension
        return super().sample(qubits, repetitions, seed)


def create_product_state(simulator, initial_state, qubits):
    """``SimulatorBase._create_simulation_state`` (sim/simulator_base.py:322-352)
    building a ``B200ProductState`` when split_untangled_states is on."""
    from cirq.sim.simulation_state_base import SimulationStateBase

    if isinstance(initial_state, SimulationStateBase):
        return initial_state
    classical_data = value.ClassicalDataDictionaryStore()
    if not simulator._split_untangled_states:
        return simulator._create_partial_simulation_state(
            initial_state=initial_state, qubits=qubits, classical_data=classical_data
        )
    args_map = {}
    if isinstance(initial_state, int):
        for q in reversed(qubits):
            args_map[q] = simulator._create_partial_simulation_state(
                initial_state=initial_state % q.dimension, qubits=[q], classical_data=classical_data
            )
            initial_state = int(initial_state / q.dimension)
    else:
        args = simulator._create_partial_simulation_state(
            initial_state=initial_state, qubits=qubits, classical_data=classical_data
        )
        for q in qubits:
            args_map[q] = args
    args_map[None] = simulator._create_partial_simulation_state(0, (), classical_data)
    return B200ProductState(args_map, qubits, True, classical_data=classical_data)


class B200StateVectorSimulationState(SimulationState[B200StateVector]):
    """State and context for operations acting on a device state vector
    (replaces ``StateVectorSimulationState``)."""

    def __init__(
        self,
        *,
        prng: np.random.RandomState | None = None,
        qubits: Sequence['cirq.Qid'] | None = None,
        initial_state: Any = 0,
        dtype=np.complex64,
        classical_data: 'cirq.ClassicalDataStore' | None = None,
        max_fused_qubits: int | None = None,
    ):
        qubits = tuple(qubits) if qubits is not None else ()
        if isinstance(initial_state, B200StateVector):
            state = initial_state  # an existing device state, adopted as is
        else:
            state = B200StateVector.create(
                initial_state=initial_state,
                qid_shape=tuple(q.dimension for q in qubits),
                dtype=dtype,
                max_fused_qubits=max_fused_qubits,
            )
        super().__init__(state=state, prng=prng, qubits=qubits, classical_data=classical_data)
        self._dtype = np.dtype(dtype)
        self._max_fused_qubits = max_fused_qubits

    def transpose_to_qubit_order(self, qubits, *, inplace=False):
        """As SimulationState.transpose_to_qubit_order (sim/simulation_state.py:217-236);
        with `inplace` the device array is permuted in place (no second buffer), which is
        what the final merge of a product state asks for."""
        if not inplace:
            return super().transpose_to_qubit_order(qubits, inplace=False)
        if len(self.qubits) != len(qubits) or set(qubits) != set(self.qubits):
            raise ValueError(f'Qubits do not match. Existing: {self.qubits}, provided: {qubits}')
        self._state.reindex_inplace(self.get_axes(qubits))
        self._set_qubits(qubits)
        return self

    def add_qubits(self, qubits):
        """Ancilla / late-joining qubits in |0> (state_vector_simulation_state.py:357-363)."""
        ret = super().add_qubits(qubits)
        if ret is not NotImplemented:
            return ret
        fresh = type(self)(
            qubits=qubits, dtype=self._dtype, prng=self._prng,
            max_fused_qubits=self._max_fused_qubits,
        )
        return self.kronecker_product(fresh, inplace=True)

    def remove_qubits(self, qubits):
        """(state_vector_simulation_state.py:365-371)"""
        ret = super().remove_qubits(qubits)
        if ret is not NotImplemented:
            return ret
        extracted, remainder = self.factor(qubits, inplace=True)
        remainder._state.scale(complex(extracted._state.device_state.amplitudes([0])[0]))
        return remainder

    # ---- act_on entry point (protocols/act_on_protocol.py:90-170) ------------------------

    def _act_on_fallback_(
        self, action: Any, qubits: Sequence['cirq.Qid'], allow_decompose: bool = True
    ) -> bool:
        strats = [_strat_unitary, _strat_mixture, _strat_channel]
        if allow_decompose:
            if (
                len(qubits) > _MAX_DIRECT_UNITARY_QUBITS
                and not hasattr(action, '_unitary_')
                and _can_decompose(action, qubits)
            ):
                # Wide composite operation: apply its parts (each a cheap pass)
                # instead of materialising a 2^k x 2^k matrix on the host.
                strats.insert(0, strat_act_on_from_apply_decompose)
            else:
                strats.append(strat_act_on_from_apply_decompose)
        for strat in strats:
            result = strat(action, self, qubits)
            if result is True:
                return True
            assert result is NotImplemented, str(result)
        raise TypeError(
            "Can't simulate operations that don't implement "
            "SupportsUnitary, SupportsConsistentApplyUnitary, "
            f"SupportsMixture or is a measurement: {action!r}"
        )

    # ---- read-out ------------------------------------------------------------------------

    @property
    def target_tensor(self) -> np.ndarray:
        """Host copy of the state as a ``(2,)*n`` tensor (downloads 2^n amplitudes)."""
        return self._state.to_numpy_tensor()

    @property
    def device_state(self) -> DeviceState:
        return self._state.device_state

    def __repr__(self) -> str:
        return (
            'cirq_b200.B200StateVectorSimulationState('
            f'qubits={self.qubits!r}, classical_data={self.classical_data!r})'
        )


def _can_decompose(action: Any, qubits) -> bool:
    if isinstance(action, ops.Gate):
        return protocols.decompose_once_with_qubits(action, qubits, None) is not None
    return protocols.decompose_once(action, None) is not None


_UNITARY_CACHE: dict = {}
_UNITARY_CACHE_MAX = 4096


# Operation types that are nothing but "this gate on these qubits": cirq.X(q) and
# friends come as SingleQubitPauliStringGateOperation, a GateOperation subclass
# whose unitary is its gate's.
PLAIN_GATE_OPERATIONS = (ops.GateOperation, ops.SingleQubitPauliStringGateOperation)


def cached_unitary(action: Any):
    """``protocols.unitary`` with a small cache keyed by the (hashable,
    parameter-free) gate: circuits repeat a handful of gates thousands of
    times and building each matrix costs ~40 us of Python.  "Has no unitary" is
    cached too (False): for a channel the protocol tries ``_unitary_``,
    ``_apply_unitary_`` and a decomposition before giving up, ~120 us that a noise
    model would pay once per inserted operation."""
    gate = getattr(action, 'gate', None)
    key = None
    if gate is not None and type(action) in PLAIN_GATE_OPERATIONS:
        try:
            key = gate
            hit = _UNITARY_CACHE.get(key)
            if hit is not None:
                return None if hit is False else hit
        except TypeError:  # unhashable gate
            key = None
    u = protocols.unitary(action, None)
    if u is not None and gate is not None and u.shape[0] <= 8 and _only_applies_in_place(gate):
        u = _unitary_by_columns(action, u.shape[0].bit_length() - 1, u)
    if key is not None and not protocols.is_parameterized(gate):
        if len(_UNITARY_CACHE) >= _UNITARY_CACHE_MAX:
            _UNITARY_CACHE.clear()
        _UNITARY_CACHE[key] = False if u is None else u
    return u


def _only_applies_in_place(gate) -> bool:
    """A gate whose only description is an ``_apply_unitary_`` method."""
    cls = type(gate)
    return getattr(cls, '_apply_unitary_', None) is not None and getattr(cls, '_unitary_', None) is None


def _unitary_by_columns(action: Any, k: int, default: np.ndarray) -> np.ndarray:
    """Matrix of a gate that only has ``_apply_unitary_``, one basis state at a time.
    ``protocols.unitary`` hands such a gate the whole identity tensor
    (protocols/unitary_protocol.py:161-179), which an ad-hoc method written for a
    STATE (e.g. one that swaps ``target_tensor[0]`` and ``[1]`` and so needs them
    to be scalars — sparse_simulator_test.py:745-754) may not survive; the
    reference calls it on the state itself, so here it sees 2^k states."""
    dim = 1 << k
    cols = np.empty((dim, dim), dtype=np.complex128)
    for j in range(dim):
        basis = np.zeros((2,) * k, dtype=np.complex128)
        basis.flat[j] = 1
        out = protocols.apply_unitary(
            action, protocols.ApplyUnitaryArgs(basis, np.empty_like(basis), range(k)), default=None)
        if out is None:
            return default
        cols[:, j] = out.reshape(-1)
    return cols


_MIXTURE_CACHE: dict = {}


def cached_mixture(base_op) -> tuple | None:
    """(probabilities, unitaries, is_identity flags) of a plain gate operation that
    is a mixture of unitaries but not a unitary itself and records nothing (the
    Pauli channels a noise model inserts hundreds of times), cached per gate; None
    otherwise."""
    gate = base_op.gate
    try:
        hit = _MIXTURE_CACHE.get(gate)
    except TypeError:  # unhashable gate
        return None
    if hit is not None:
        return hit or None
    form: Any = False
    if (
        0 < len(base_op.qubits) <= 3
        and not protocols.is_parameterized(gate)
        and not protocols.is_measurement(base_op)
        and all(d == 2 for d in protocols.qid_shape(base_op))
        and not protocols.has_unitary(base_op)
    ):
        mixture = protocols.mixture(base_op, default=None)
        if mixture is not None:
            probabilities, unitaries = zip(*mixture)
            unitaries = [np.asarray(u, dtype=np.complex128) for u in unitaries]
            eye = np.eye(unitaries[0].shape[0])
            form = (probabilities, unitaries, [np.array_equal(u, eye) for u in unitaries])
    if len(_MIXTURE_CACHE) >= _UNITARY_CACHE_MAX:
        _MIXTURE_CACHE.clear()
    _MIXTURE_CACHE[gate] = form
    return form or None


def _strat_unitary(action: Any, args: B200StateVectorSimulationState, qubits) -> bool:
    """Unitary strategy: obtain the matrix from Cirq's query protocols and queue
    it.  Replaces ``_strat_act_on_state_vector_from_apply_unitary``
    (state_vector_simulation_state.py:402-407): the per-gate ``_apply_unitary_``
    numpy fast paths are never called."""
    u = cached_unitary(action)
    if u is None:
        return NotImplemented
    args._state.queue_unitary(u, args.get_axes(qubits))
    return True


def _strat_mixture(action: Any, args: B200StateVectorSimulationState, qubits) -> bool:
    """Samples one unitary of a mixture (state_vector_simulation_state.py:183-203)."""
    mixture = protocols.mixture(action, default=None)
    if mixture is None:
        return NotImplemented
    probabilities, unitaries = zip(*mixture)
    index = args.prng.choice(range(len(unitaries)), p=probabilities)
    args._state.queue_unitary(unitaries[index], args.get_axes(qubits))
    if protocols.is_measurement(action):
        key = protocols.measurement_key_obj(action)
        args._classical_data.record_channel_measurement(key, index)
    return True


def _strat_channel(action: Any, args: B200StateVectorSimulationState, qubits) -> bool:
    """Kraus-trajectory sampling, the algorithm of
    state_vector_simulation_state.py:205-257: try operators in order, weight =
    ||K_i psi||^2, stop when the uniform draw is used up, renormalise."""
    kraus_operators = protocols.kraus(action, default=None)
    if kraus_operators is None:
        return NotImplemented
    state = args._state
    axes = args.get_axes(qubits)
    p = args.prng.random()
    fallback_weight = 0.0
    fallback_index = 0
    base = state.copy()
    chosen = None
    index = 0
    weight = None
    for index, k in enumerate(kraus_operators):
        trial = base.copy()
        trial.apply_matrix_now(k, axes)
        weight = trial.norm2()
        if weight > fallback_weight:
            fallback_weight = weight
            fallback_index = index
        p -= weight
        if p < 0:
            chosen = trial
            break
    assert weight is not None, 'No Kraus operators'
    if chosen is None or weight == 0:
        chosen = base.copy()
        chosen.apply_matrix_now(kraus_operators[fallback_index], axes)
        weight = fallback_weight
        index = fallback_index
    chosen._dev.scale(1.0 / np.sqrt(weight))
    state.replace_device_state(chosen._dev)
    if protocols.is_measurement(action):
        key = protocols.measurement_key_obj(action)
        args._classical_data.record_channel_measurement(key, index)
    return True


class _FastConfuseMixin:
    """Vectorised ``StepResult.sample_measurement_ops`` (sim/simulator.py:733-820).

    Same validation, ordering, invert-mask, confusion-map and repeated-key
    semantics as the reference, but without its per-repetition Python loops:
    ``_confuse_results`` (:822-843) walks every repetition even when no
    measurement has a confusion map (2 s per million samples) and the per-qubit
    column copies are replaced by one gather."""

    def _confuse_results(self, bits, qubits, confusion_map, seed=None) -> None:
        if not confusion_map:
            return
        super()._confuse_results(bits, qubits, confusion_map, seed)

    def sample_measurement_ops(
        self, measurement_ops, repetitions: int = 1, seed=None, *, _allow_repeated=False
    ):
        import collections

        for op in measurement_ops:
            if not isinstance(op.gate, ops.MeasurementGate):
                raise ValueError(f'{op.gate} was not a MeasurementGate')
        result = collections.Counter(
            key for op in measurement_ops for key in protocols.measurement_key_names(op)
        )
        if result and not _allow_repeated:
            duplicates = [k for k, v in result.most_common() if v > 1]
            if duplicates:
                raise ValueError(f"Measurement key {','.join(duplicates)} repeated")

        measured_qubits = []
        seen_qubits = set()
        for op in measurement_ops:
            for q in op.qubits:
                if q not in seen_qubits:
                    seen_qubits.add(q)
                    measured_qubits.append(q)

        indexed_sample = self.sample(measured_qubits, repetitions, seed=seed)
        qubits_to_index = {q: i for i, q in enumerate(measured_qubits)}
        results = {}
        for op in measurement_ops:
            gate = op.gate
            key = gate.key
            cols = [qubits_to_index[q] for q in op.qubits]
            if cols == list(range(indexed_sample.shape[1])):
                arr = indexed_sample if len(measurement_ops) == 1 else indexed_sample.copy()
            else:
                arr = indexed_sample[:, cols]
            # the reference returns int8 on this path (sim/simulator.py:802)
            out = arr.view(np.int8) if arr.dtype == np.uint8 else arr.astype(np.int8, copy=False)
            inv = [i for i, flip in enumerate(gate.full_invert_mask()) if flip]
            if inv:
                out[:, inv] ^= 1
            self._confuse_results(out, op.qubits, gate.confusion_map, seed)
            if _allow_repeated:
                results.setdefault(key, []).append(out)
            else:
                results[key] = out
        if not _allow_repeated:
            return results
        return {
            k: (v[0][:, np.newaxis, :] if len(v) == 1 else np.array(v).swapaxes(0, 1))
            for k, v in results.items()
        }


class _DeviceReducedStateMixin:
    """``density_matrix_of`` / ``bloch_vector_of`` (sim/state_vector.py:109-167)
    evaluated by a reduction kernel on the device state instead of an einsum over
    a host copy (qis/states.py:676-693) — which at 34 qubits would first download
    137 GB, and which the reference refuses above 25 qubits.  Up to 5 kept
    qubits; wider requests (and ``qubits=None``, the full outer product) take
    the reference's host route."""

    _MAX_KEPT = 5

    def _merged_mapped_device_state(self):
        """(device array, where) of the merged state, see
        ``B200StateVector.mapped_device_state``."""
        raise NotImplementedError

    def density_matrix_of(self, qubits=None) -> np.ndarray:
        if qubits is None or not 1 <= len(qubits) <= self._MAX_KEPT:
            return super().density_matrix_of(qubits)
        axes = [self.qubit_map[q] for q in qubits]  # KeyError for foreign qubits, as the reference
        if len(set(axes)) != len(axes):
            return super().density_matrix_of(qubits)
        dev, where = self._merged_mapped_device_state()
        n = dev.n_bits
        rho = dev.reduced_density_matrix([where[n - 1 - a] for a in axes])
        return rho.astype(dev.dtype)

    def bloch_vector_of(self, qubit) -> np.ndarray:
        rho = self.density_matrix_of([qubit])
        v = np.zeros(3, dtype=np.float32)  # qis/states.py:614-620
        v[0] = 2 * np.real(rho[0][1])
        v[1] = 2 * np.imag(rho[1][0])
        v[2] = np.real(rho[0][0] - rho[1][1])
        return v


class B200SimulatorStep(
    _FastConfuseMixin, _DeviceReducedStateMixin, state_vector.StateVectorMixin,
    state_vector_simulator.StateVectorStepResult,
):
    """Step result of ``B200Simulator`` (replaces ``SparseSimulatorStep``,
    sim/sparse_simulator.py:221-290)."""

    def __init__(self, sim_state, dtype=np.complex64):
        qubit_map = {q: i for i, q in enumerate(sim_state.qubits)}
        super().__init__(sim_state=sim_state, qubit_map=qubit_map)
        self._dtype = dtype
        self._state_vector: np.ndarray | None = None

    def _merged_mapped_device_state(self):
        return self._merged_sim_state._state.mapped_device_state()

    def state_vector(self, copy: bool = False) -> np.ndarray:
        """Host copy of the state vector (big-endian), downloaded on first use."""
        if self._state_vector is None:
            self._state_vector = np.array([1])
            state = self._merged_sim_state
            if state is not None:
                vector = state.target_tensor
                self._state_vector = np.reshape(vector, -1)
        return self._state_vector.copy() if copy else self._state_vector

    def __repr__(self) -> str:
        return (
            f'cirq_b200.B200SimulatorStep(sim_state={self._sim_state!r},'
            f' dtype=np.{np.dtype(self._dtype)!r})'
        )


class B200StateVectorTrialResult(_DeviceReducedStateMixin, state_vector_simulator.StateVectorTrialResult):
    """Trial result whose final state stays on the device until asked for."""

    def _merged_mapped_device_state(self):
        return self._get_merged_sim_state()._state.mapped_device_state()

    @property
    def device_state(self) -> DeviceState:
        """The final state in HBM (complex[2^n], big-endian)."""
        return self._get_merged_sim_state().device_state


class B200Simulator(
    state_vector_simulator.SimulatesIntermediateStateVector['B200SimulatorStep'],
    simulator.SimulatesExpectationValues,
):
    """Drop-in for ``cirq.Simulator`` running on one B200.

    Implements SimulatesSamples / SimulatesFinalState /
    SimulatesIntermediateState / SimulatesAmplitudes /
    SimulatesExpectationValues: ``run``, ``run_sweep``, ``simulate``,
    ``simulate_sweep``, ``simulate_moment_steps``, ``compute_amplitudes`` and
    ``simulate_expectation_values`` work unchanged on ``cirq.Circuit`` objects.

    Args:
        dtype: ``np.complex64`` or ``np.complex128``.
        noise, seed: as ``cirq.Simulator``.
        split_untangled_states: as ``cirq.Simulator`` (default True): unentangled
            qubit sets are kept as separate device states and joined by a
            Kronecker-product kernel when a gate couples them, at any register
            size (above 31 qubits a join merges the whole register at once).
        max_fused_qubits: widest fused block (one GPU pass each).
        sweep_batch: False (default) runs ``run_sweep`` resolver by resolver like
            the reference; True lays all resolvers out as one device array and
            walks the circuit once (``cirq_b200.sweeps``) — same results.
        trajectory_batch: 0 (default) keeps the reference's one-simulation-per-
            repetition loop for noisy / mid-circuit-measured ``run`` calls, with
            its seeded results.  N > 1 advances up to N repetitions together as
            one device array (``cirq_b200.trajectories``): same distribution of
            results, random numbers consumed in a different order.
        plan_cache: True (default) keeps the device schedule of a circuit's unitary
            prefix from the second time the same circuit object is executed on
            (``cirq_b200.plan_cache``): later ``run`` / ``simulate`` calls skip the
            per-operation Python of the driver loop and the gate fuser.
    """

    def __init__(
        self,
        *,
        dtype=np.complex64,
        noise: 'cirq.NOISE_MODEL_LIKE' = None,
        seed: 'cirq.RANDOM_STATE_OR_SEED_LIKE' = None,
        split_untangled_states: bool = True,
        max_fused_qubits: int | None = None,
        trajectory_batch: int = 0,
        sweep_batch: bool = False,
        plan_cache: bool = True,
    ):
        if np.dtype(dtype).kind != 'c':
            raise ValueError(f'dtype must be a complex type but was {dtype}')
        if np.dtype(dtype) not in (np.dtype(np.complex64), np.dtype(np.complex128)):
            raise ValueError(f'dtype must be complex64 or complex128 but was {dtype}')
        super().__init__(
            dtype=dtype, noise=noise, seed=seed, split_untangled_states=split_untangled_states
        )
        # None = kernel-matched policy (cirq_b200.fusion.fuser_for)
        self._max_fused = None if max_fused_qubits is None else int(max_fused_qubits)
        self._trajectory_batch = int(trajectory_batch)
        self._sweep_batch = bool(sweep_batch)
        self._plan_cache = bool(plan_cache)
        self.last_run_info: dict = {}

    def _state_from_device(self, dev: DeviceState, qubits):
        """Simulation state adopting an existing device array (cirq_b200.sweeps)."""
        return B200StateVectorSimulationState(
            qubits=qubits, prng=self._prng, dtype=self._dtype, max_fused_qubits=self._max_fused,
            initial_state=B200StateVector(dev, len(qubits), self._max_fused),
        )

    # ---- per-circuit schedule cache (cirq_b200/plan_cache.py) ------------------------------

    def _lookup_plan(self, circuit, qubit_order=ops.QubitOrder.DEFAULT):
        """The cached schedule of `circuit`'s unitary prefix from |0...0>, built the
        second time the circuit is seen; None when there is none (first sighting,
        noise model, parameterized or qudit circuit, custom QubitOrder object)."""
        from cirq import devices
        from cirq_b200 import plan_cache

        if not self._plan_cache or self.noise is not devices.NO_NOISE or not plan_cache.enabled():
            return None
        if qubit_order is ops.QubitOrder.DEFAULT:
            order_key: Any = None
        elif isinstance(qubit_order, (list, tuple)):
            order_key = tuple(qubit_order)
        else:
            return None
        moments = circuit.moments
        if not moments:
            return None
        try:
            key = (tuple(map(id, moments)), order_key, np.dtype(self._dtype).str, self._max_fused,
                   self._split_untangled_states, DeviceState)
            found, plan = plan_cache.CACHE.get(key)
        except TypeError:  # unhashable entries in qubit_order
            return None
        if found:
            return plan
        if not plan_cache.CACHE.seen_before(key):
            return None
        plan = self._record_prefix(circuit, qubit_order)
        plan_cache.CACHE.put(key, plan)
        return plan

    def _record_prefix(self, circuit, qubit_order):
        """Dry run of the longest all-unitary prefix of `circuit` against a recording
        device state: the product-state logic and the fuser run as in a live call, the
        device operations are written down (no GPU work).  None if it cannot be cached."""
        from cirq_b200 import plan_cache

        tagged = ops.TaggedOperation
        first = 0
        for moment in circuit.moments:
            ok = True
            for op in moment.operations:
                base = op.untagged if type(op) is tagged else op
                # (wider operations end the prefix: asking a 40-qubit measurement for its
                # unitary would build a 2^40 x 2^40 identity)
                if (type(base) not in PLAIN_GATE_OPERATIONS or len(base.qubits) > 5
                        or cached_unitary(base) is None):
                    ok = False
                    break
            if not ok:
                break
            first += 1
        if first == 0 or protocols.is_parameterized(circuit):
            return None
        qubits = ops.QubitOrder.as_qubit_order(qubit_order).order_for(circuit.all_qubits())
        if not qubits or any(q.dimension != 2 for q in qubits):
            return None
        rec = plan_cache.recorder_for(DeviceState)
        classical_data = value.ClassicalDataDictionaryStore()
        prng = np.random.RandomState(0)  # (a unitary prefix draws nothing)

        def fresh(qs):
            return B200StateVectorSimulationState(
                qubits=qs, prng=prng, classical_data=classical_data, dtype=self._dtype,
                max_fused_qubits=self._max_fused,
                initial_state=B200StateVector(rec.basis(len(qs), self._dtype, 0), len(qs), self._max_fused),
            )

        if self._split_untangled_states:
            # as create_product_state for initial_state=0 (sim/simulator_base.py:322-352)
            args_map: dict = {q: fresh([q]) for q in reversed(qubits)}
            args_map[None] = fresh([])
            state: Any = B200ProductState(args_map, qubits, True, classical_data=classical_data)
        else:
            state = fresh(list(qubits))
        try:
            for _ in self._core_iterator(circuit=circuit[:first], sim_state=state):
                pass
            subs = list(dict.fromkeys(state.sim_states.values())) if self._split_untangled_states else [state]
            components = []
            for sub in subs:
                sv = sub._state
                blocks = sv.detach_pending()
                where = list(sv._where) if sv._where is not None else None
                components.append((sv._dev.ident, tuple(sub.qubits), where, blocks))
        except plan_cache.Untraceable:
            return None
        owner = None
        if self._split_untangled_states:
            index = {id(sub): i for i, sub in enumerate(subs)}
            owner = [(q, index[id(sub)]) for q, sub in state.sim_states.items()]
        rest = circuit.moments[first:]
        plan = plan_cache.PrefixPlan(tuple(circuit.moments), first, tuple(qubits), rec.ops, components, owner)
        # (for `run`: is what follows the prefix nothing but measurements?)
        plan.measure_tail = bool(rest) and all(
            isinstance(op.gate, ops.MeasurementGate) for m in rest for op in m.operations)
        return plan

    def _state_from_plan(self, plan):
        """Replays a cached prefix on the GPU and rebuilds the simulation state the live
        loop would have left behind."""
        from cirq_b200 import plan_cache

        live, passes = plan_cache.replay(plan, self._dtype, DeviceState)
        classical_data = value.ClassicalDataDictionaryStore()
        subs = []
        for ident, qs, where, blocks in plan.components:
            sv = B200StateVector(live[ident], len(qs), self._max_fused)
            sv._where = list(where) if where is not None else None
            sv._held = list(blocks)  # launched (paired) with whatever comes next, or by flush()
            sv.passes = passes[ident]
            subs.append(B200StateVectorSimulationState(
                qubits=qs, prng=self._prng, classical_data=classical_data, dtype=self._dtype,
                max_fused_qubits=self._max_fused, initial_state=sv,
            ))
        if plan.owner is None:
            return subs[0]
        return B200ProductState({q: subs[i] for q, i in plan.owner}, plan.qubits, True,
                                classical_data=classical_data)

    def simulate_sweep_iter(self, program, params, qubit_order=ops.QubitOrder.DEFAULT, initial_state=None):
        """``SimulatorBase.simulate_sweep_iter`` (sim/simulator_base.py:277-320); with
        ``sweep_batch=True`` a measurement-free sweep from |0...0> in the default
        qubit order advances all resolvers as one device array."""
        if self._sweep_batch and initial_state is None and qubit_order is ops.QubitOrder.DEFAULT:
            from cirq_b200 import sweeps

            batched = sweeps.simulate_sweep_batched(self, 'sv', program, params, DeviceState)
            if batched is not None:
                yield from batched
                return
        from cirq import devices

        resolvers = list(study.to_resolvers(params))
        if len(resolvers) == 1 and self.noise is devices.NO_NOISE:
            # a plain simulate(): nothing to share between resolvers, so the
            # reference's split into a resolver-independent prefix and the rest
            # (sim/simulator_base.py:304-320: ~40 us of Python per operation before
            # the first gate reaches the GPU) is skipped — same result (with a noise
            # model the split decides where noise lands, so it is kept)
            if (initial_state is None or (type(initial_state) is int and initial_state == 0)) and not resolvers[0]:
                plan = self._lookup_plan(program, qubit_order)
                if plan is not None:
                    # the unitary prefix replayed from the schedule cache, the rest of
                    # the circuit as in simulate_sweep_iter (sim/simulator.py:586-605)
                    sim_state = self._state_from_plan(plan)
                    measurements: dict = {}
                    if plan.first < len(program):
                        for step in self._core_iterator(circuit=program[plan.first:], sim_state=sim_state):
                            for k, v in step.measurements.items():
                                measurements[k] = np.array(v, dtype=np.uint8)
                    yield self._create_simulator_trial_result(
                        params=resolvers[0], measurements=measurements, final_simulator_state=sim_state)
                    return
            yield from simulator.SimulatesIntermediateState.simulate_sweep_iter(
                self, program, resolvers, qubit_order, initial_state)
            return
        yield from super().simulate_sweep_iter(program, resolvers, qubit_order, initial_state)

    def run_sweep_iter(self, program, params, repetitions: int = 1):
        """``SimulatesSamples.run_sweep_iter`` (sim/simulator.py:62-94); with
        ``sweep_batch=True`` all resolvers advance together as one device array
        when the circuit allows it (cirq_b200.sweeps)."""
        if self._sweep_batch:
            from cirq_b200 import sweeps

            batched = sweeps.run_sweep_batched(self, 'sv', program, params, repetitions, DeviceState)
            if batched is not None:
                yield from batched
                return
        yield from super().run_sweep_iter(program, params, repetitions)

    def _run(self, circuit, param_resolver, repetitions: int):
        """``SimulatorBase._run`` (sim/simulator_base.py:215-275) with the
        per-repetition loop (:249-264) replaced by batched trajectories when
        ``trajectory_batch`` asks for it and every suffix operation can be
        batched; anything else takes the reference's loop unchanged."""
        self.last_run_info = {'path': 'reference loop'}
        fast = self._run_unitary_then_measure(circuit, param_resolver, repetitions)
        if fast is not None:
            return fast
        if self._trajectory_batch <= 1 or repetitions <= 1:
            return super()._run(circuit, param_resolver, repetitions)
        from cirq.sim.simulator import check_all_resolved, split_into_matching_protocol_then_general
        from cirq_b200 import trajectories

        resolver = param_resolver or study.ParamResolver({})
        resolved = protocols.resolve_parameters(circuit, resolver)
        check_all_resolved(resolved)
        qubits = tuple(sorted(resolved.all_qubits()))
        prefix, suffix = (
            split_into_matching_protocol_then_general(resolved, self._can_be_in_run_prefix)
            if self._can_be_in_run_prefix(self.noise)
            else (resolved[0:0], resolved)
        )
        suffix_ops = list(suffix.all_operations())
        if not qubits or all(isinstance(op.gate, ops.MeasurementGate) for op in suffix_ops):
            return super()._run(circuit, param_resolver, repetitions)
        noisy = list(self.noise.noisy_moments(suffix, sorted(suffix.all_qubits())))
        plan = trajectories.plan_suffix(noisy, qubits)
        if plan is None or len(qubits) > trajectories.MAX_BATCH_STATE_BITS:
            return super()._run(circuit, param_resolver, repetitions)
        sim_state = self._create_simulation_state(0, qubits)
        step_result = None
        for step_result in self._core_iterator(circuit=prefix, sim_state=sim_state):
            pass
        merged = step_result._merged_sim_state
        # canonical order: axis i of the merged state is qubits[i]
        assert tuple(merged.qubits) == qubits
        info: dict = {}
        out = trajectories.run_plan(
            plan, merged.device_state, len(qubits), repetitions, self._dtype, self._prng,
            self._trajectory_batch, self._max_fused, info=info,
        )
        info['path'] = 'batched trajectories'
        self.last_run_info = info
        return out

    def _core_iterator(self, circuit, sim_state, all_measurements_are_terminal: bool = False):
        """``SimulatorBase._core_iterator`` (sim/simulator_base.py:169-214) with one
        shortcut: a plain gate operation on at most 5 qubits that has a unitary goes
        straight into the device state's queue (`B200ProductState.apply_unitary_op`)
        instead of through ``protocols.act_on`` -> ``_act_on_fallback_`` ->
        strategy list, which ends in exactly that call ~35 us of dispatch later.
        Everything else — measurements, channels, classical control, tagged or
        composite operations — takes ``protocols.act_on`` as in the reference."""
        import collections

        if len(circuit) == 0:
            yield self._create_step_result(sim_state)
            return
        noisy_moments = self.noise.noisy_moments(circuit, sorted(circuit.all_qubits()))
        measured: dict = collections.defaultdict(bool)
        product = isinstance(sim_state, B200ProductState)
        dense = isinstance(sim_state, B200StateVectorSimulationState)
        lean = product or dense
        plain, moment_type, tagged = PLAIN_GATE_OPERATIONS, circuits.Moment, ops.TaggedOperation
        for moment in noisy_moments:
            moment_ops = moment.operations if type(moment) is moment_type else ops.flatten_to_ops(moment)
            for op in moment_ops:
                try:
                    if all_measurements_are_terminal and measured[op.qubits]:
                        continue
                    if isinstance(op.gate, ops.MeasurementGate):
                        measured[op.qubits] = True
                        if all_measurements_are_terminal:
                            continue
                    base = op.untagged if type(op) is tagged else op
                    if lean and type(base) in plain and 0 < len(base.qubits) <= 5:
                        u = cached_unitary(base)
                        if u is not None:
                            if product:
                                sim_state.apply_unitary_op(base, u)
                            else:
                                sim_state._state.queue_unitary(u, sim_state.get_axes(base.qubits))
                            continue
                        mix = cached_mixture(base)
                        if mix is not None:
                            # _strat_mixture with the gate's mixture cached: the same
                            # single draw; an identity pick queues nothing
                            target = sim_state.join_for(base.qubits) if product else sim_state
                            index = target.prng.choice(range(len(mix[1])), p=mix[0])
                            if not mix[2][index]:
                                target._state.queue_unitary(mix[1][index], target.get_axes(base.qubits))
                            continue
                    protocols.act_on(op, sim_state)
                except TypeError:
                    raise TypeError(f"{self.__class__.__name__} doesn't support {op!r}")
            yield self._create_step_result(sim_state)

    def _run_unitary_then_measure(self, circuit, param_resolver, repetitions: int):
        """The most common shape — a noise-free circuit of unitary gates whose last
        moments are measurements — without the reference's generic prefix/suffix
        split (sim/simulator.py:952-985: ~15 ms of Python on a 900-operation
        circuit before the GPU gets its first gate).  For this shape the split is
        simply "moments before the first measurement | the rest", and the rest of
        `_run` (sim/simulator_base.py:229-244) follows unchanged.  Returns None for
        any other shape."""
        from cirq import devices

        if self.noise is not devices.NO_NOISE:
            return None
        if param_resolver and protocols.is_parameterized(circuit):
            return None  # sweeps over symbols: the generic path resolves them
        plan = self._lookup_plan(circuit)
        if plan is not None and plan.measure_tail:
            # seen before: the unitary prefix is replayed from the schedule cache
            sim_state = self._state_from_plan(plan)
            suffix = circuit[plan.first:]
            step_result = None
            for step_result in self._core_iterator(
                circuit=suffix, sim_state=sim_state, all_measurements_are_terminal=True
            ):
                pass
            self.last_run_info = {'path': 'unitary prefix + sampling', 'plan_cache': 'hit'}
            return step_result.sample_measurement_ops(
                list(suffix.all_operations()), repetitions, seed=self._prng, _allow_repeated=True
            )
        first = None
        for i, moment in enumerate(circuit):
            measuring = [isinstance(op.gate, ops.MeasurementGate) for op in moment]
            if first is None:
                if any(measuring):
                    if not all(measuring):
                        return None
                    first = i
                else:
                    for op in moment:
                        if type(op) not in PLAIN_GATE_OPERATIONS or cached_unitary(op) is None:
                            return None
            elif not all(measuring):
                return None
        if first is None:
            return None
        qubits = tuple(sorted(circuit.all_qubits()))
        sim_state = self._create_simulation_state(0, qubits)
        step_result = None
        for step_result in self._core_iterator(circuit=circuit[:first], sim_state=sim_state):
            pass
        suffix = circuit[first:]
        for step_result in self._core_iterator(
            circuit=suffix, sim_state=sim_state, all_measurements_are_terminal=True
        ):
            pass
        return step_result.sample_measurement_ops(
            list(suffix.all_operations()), repetitions, seed=self._prng, _allow_repeated=True
        )

    def _create_partial_simulation_state(self, initial_state, qubits, classical_data):
        if isinstance(initial_state, B200StateVectorSimulationState):
            return initial_state
        return B200StateVectorSimulationState(
            qubits=qubits,
            prng=self._prng,
            classical_data=classical_data,
            initial_state=initial_state,
            dtype=self._dtype,
            max_fused_qubits=self._max_fused,
        )

    def _create_simulation_state(self, initial_state, qubits):
        """As SimulatorBase (sim/simulator_base.py:322-352): product-state form at any
        size when split_untangled_states is on (B200ProductState.join_for bounds the
        memory of the joins)."""
        return create_product_state(self, initial_state, qubits)

    def _create_step_result(self, sim_state):
        return B200SimulatorStep(sim_state=sim_state, dtype=self._dtype)

    def _create_simulator_trial_result(self, params, measurements, final_simulator_state):
        return B200StateVectorTrialResult(
            params=params, measurements=measurements, final_simulator_state=final_simulator_state
        )

    # ---- device-side result consumers -------------------------------------

    def compute_amplitudes_sweep_iter(
        self, program, bitstrings, params, qubit_order=ops.QubitOrder.DEFAULT
    ) -> Iterator[Sequence[complex]]:
        """Amplitudes gathered on the device instead of indexing a downloaded
        state (sim/state_vector_simulator.py:74-98)."""
        if isinstance(bitstrings, np.ndarray) and len(bitstrings.shape) > 1:
            raise ValueError(
                'The list of bitstrings must be input as a '
                '1-dimensional array of ints. Got an array with '
                f'shape {bitstrings.shape}.'
            )
        idx = [int(b) for b in bitstrings]
        for trial_result in self.simulate_sweep_iter(program, params, qubit_order):
            # (no passes spent on putting relabelled SWAPs back: indices are translated)
            dev, where = trial_result._get_merged_sim_state()._state.mapped_device_state()
            total = 1 << dev.n_bits
            wrapped = [i % total if -total <= i < total else i for i in idx]
            for i in wrapped:
                if i < 0 or i >= total:
                    raise IndexError(f'index {i} is out of bounds for axis 0 with size {total}')
            if where != list(range(dev.n_bits)):
                wrapped = [sum(((i >> b) & 1) << w for b, w in enumerate(where)) for i in wrapped]
            amps = dev.amplitudes(wrapped).astype(self._dtype)
            yield amps.tolist()

    def simulate_expectation_values_sweep_iter(
        self,
        program,
        observables,
        params,
        qubit_order=ops.QubitOrder.DEFAULT,
        initial_state=None,
        permit_terminal_measurements: bool = False,
    ) -> Iterator[list[float]]:
        """<psi|O|psi> computed on the device, one reduction pass per Pauli
        string (replaces sim/sparse_simulator.py:193-218 ->
        ops/pauli_string.py:625-655)."""
        if not permit_terminal_measurements and program.are_any_measurements_terminal():
            raise ValueError(
                'Provided circuit has terminal measurements, which may '
                'skew expectation values. If this is intentional, set '
                'permit_terminal_measurements=True.'
            )
        if not isinstance(observables, list):
            observables = [observables]
        pslist = [ops.PauliSum.wrap(pslike) for pslike in observables]
        if self._sweep_batch and initial_state is None and qubit_order is ops.QubitOrder.DEFAULT:
            from cirq_b200 import sweeps

            batched = sweeps.expectation_sweep_batched(
                self, 'sv', program, pslist, params, DeviceState, pauli_sum_expectation)
            if batched is not None:
                yield from batched
                return
        order = ops.QubitOrder.as_qubit_order(qubit_order)
        qmap = {q: i for i, q in enumerate(order.order_for(program.all_qubits()))}
        for result in self.simulate_sweep_iter(
            program, params, qubit_order=qubit_order, initial_state=initial_state
        ):
            dev, where = result._get_merged_sim_state()._state.mapped_device_state()
            yield [pauli_sum_expectation(dev, obs, qmap, where) for obs in pslist]


def pauli_masks(pauli_string, qubit_map, n_qubits: int, where: Sequence[int] | None = None) -> tuple[int, int]:
    """x/z bit masks of a PauliString under axis -> bit p = n-1-axis (-> where[p]
    for a state whose bits were relabelled, ``B200StateVector.mapped_device_state``)."""
    x = z = 0
    for q, p in pauli_string.items():
        if q not in qubit_map:
            raise ValueError(f'Qubit {q} of the observable is not in the circuit')
        bit = n_qubits - 1 - qubit_map[q]
        if where is not None:
            bit = where[bit]
        if p == ops.X or p == ops.Y:
            x |= 1 << bit
        if p == ops.Z or p == ops.Y:
            z |= 1 << bit
    return x, z


def pauli_sum_expectation(dev: DeviceState, pauli_sum, qubit_map, where: Sequence[int] | None = None) -> complex:
    """sum_k c_k <psi|P_k|psi>.  Terms that flip the same bits (equal X mask) share
    ONE device reduction (b2q_sv_pauli_expectation_multi: the pair product
    conj(psi[i ^ x]) psi[i] is formed once, each term adds its own sign), so a sum of
    Z-type terms costs one read of the state, not one per term as in
    sim/sparse_simulator.py:193-218."""
    n = dev.n_bits
    groups: dict[int, list] = {}
    for ps in pauli_sum:
        if abs(complex(ps.coefficient).imag) > 0.0001:
            raise NotImplementedError(
                'Cannot compute expectation value of a non-Hermitian '
                f'PauliString <{ps}>. Coefficient must be real.'
            )
        x, z = pauli_masks(ps, qubit_map, n, where)
        groups.setdefault(x, []).append((z, ps.coefficient))
    total = 0.0 + 0.0j
    for x, terms in groups.items():
        if len(terms) == 1:
            total += terms[0][1] * dev.pauli_expectation(x, terms[0][0])
        else:
            values = dev.pauli_expectations(x, [z for z, _ in terms])
            total += sum(c * v for (_, c), v in zip(terms, values))
    return total
