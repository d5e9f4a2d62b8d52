"""B200 density-matrix simulator: a drop-in for ``cirq.DensityMatrixSimulator``.

rho over n qubits is stored as a 2n-"qubit" vector in HBM (flat index = row *
2^n + column; row bits above column bits — the layout of the reference's
``(2,)*2n`` tensor, sim/density_matrix_simulation_state.py:33-60).  Every
operation becomes a gate on that vector, applied by the same streaming
kernels as the state-vector path:

  * unitary U on qubits Q      ->  U on the row bits of Q, conj(U) on the column bits
                                   (``_apply_unitary``, protocols/apply_channel_protocol.py:274-294)
  * channel {K_i} on qubits Q  ->  the superoperator sum_i K_i (x) conj(K_i) on
                                   (row bits, column bits) of Q  (``_apply_kraus``, :297-356;
                                   the reference does copy + 2 multiplies + accumulate PER Kraus
                                   operator, ~16 passes for a 1-qubit depolarising channel; here
                                   it is one pass, and it fuses with neighbouring gates)

Operations are queued and fused exactly as in the state-vector simulator, so
e.g. a 2-qubit gate followed by depolarising noise on both qubits is a single
4-"qubit" pass over rho.  ``split_untangled_states`` is honoured through device
kron / partial-trace / axis-permutation kernels (``kron``, ``factor``, ``reindex``).
"""
from __future__ import annotations

from typing import Any, Sequence

import numpy as np

from cirq_b200._cirq_compat import import_cirq
from cirq_b200.device_state import DeviceState
from cirq_b200.fusion import GateFuser, fuser_for
from cirq_b200.sv_simulator import _FastConfuseMixin, create_product_state

cirq = import_cirq()

from cirq import ops, protocols, qis, study, value  # noqa: E402
from cirq.sim import density_matrix_simulator as _ref_dm  # noqa: E402
from cirq.sim import simulator, simulator_base  # noqa: E402
from cirq.sim.simulation_state import SimulationState, strat_act_on_from_apply_decompose  # noqa: E402

_MAX_DIRECT_QUBITS = 3  # widest op whose Kraus/unitary matrices are requested directly


class B200DensityMatrix(qis.QuantumStateRepresentation):
    """Device-resident replacement of ``_BufferedDensityMatrix`` (no scratch
    buffers: the reference keeps three)."""

    def __init__(self, dev: DeviceState, num_qubits: int, max_fused_qubits: int | None = None):
        self._dev = dev
        self._n = int(num_qubits)
        self._max_fused = max_fused_qubits
        # rho's natural block is 4 index bits (row+column bits of a qubit pair):
        # 5-bit blocks save no passes here
        self._fuser = fuser_for(dev.dtype, 4 if max_fused_qubits is None else max_fused_qubits, 2 * self._n)
        self._qid_shape = (2,) * self._n
        self.passes = 0
        self._since_drain = 0
        self._drain_every = max(8, 2 * self._n)
        self._held: list = []  # final blocks kept back by a drain to be paired later
        self._host = None  # cached host copy, dropped on every mutation

    @classmethod
    def create(cls, *, initial_state: Any = 0, qid_shape, dtype=np.complex64, max_fused_qubits=None):
        if any(d != 2 for d in qid_shape):
            raise ValueError(
                f'cirq_b200 simulates qubits only (dimension 2); got qid_shape={qid_shape}'
            )
        n = len(qid_shape)
        if isinstance(initial_state, (int, np.integer)):
            index = int(initial_state)
            if index < 0 or index >= (1 << n):
                raise ValueError(
                    f'initial_state={index} was out of range for {n} qubits (qid_shape={qid_shape})'
                )
            dev = DeviceState.basis(2 * n, dtype, index * ((1 << n) + 1))
        else:
            if isinstance(initial_state, np.ndarray) and dtype and initial_state.dtype != dtype:
                initial_state = initial_state.astype(dtype)
            rho = qis.to_valid_density_matrix(initial_state, n, qid_shape=qid_shape, dtype=dtype)
            dev = DeviceState.from_numpy(np.asarray(rho).reshape(-1), dtype)
        return cls(dev, n, max_fused_qubits)

    # ---- bit maps ---------------------------------------------------------------------------

    def _col_bits(self, axes: Sequence[int]) -> list[int]:
        return [self._n - 1 - int(a) for a in axes]

    def _row_bits(self, axes: Sequence[int]) -> list[int]:
        return [2 * self._n - 1 - int(a) for a in axes]

    # ---- queue ------------------------------------------------------------------------------

    def queue_unitary(self, u: np.ndarray, axes: Sequence[int]) -> None:
        self._host = None
        if len(axes) == 0:
            return  # |c|^2 = 1 for a unitary scalar: rho is unchanged
        u = np.asarray(u, dtype=np.complex128)
        self._fuser.add(u, self._row_bits(axes))
        self._fuser.add(np.conj(u), self._col_bits(axes))
        self._maybe_drain(2)

    def queue_kraus(self, kraus_ops: Sequence[np.ndarray], axes: Sequence[int]) -> None:
        self._host = None
        k = len(axes)
        if k == 0:
            scale = sum(abs(complex(np.asarray(op).reshape(-1)[0])) ** 2 for op in kraus_ops)
            self.flush()
            self._dev.scale(scale)
            return
        sup = None
        for op in kraus_ops:
            op = np.asarray(op, dtype=np.complex128).reshape(1 << k, 1 << k)
            term = np.kron(op, np.conj(op))
            sup = term if sup is None else sup + term
        self._fuser.add(sup, self._row_bits(axes) + self._col_bits(axes))
        self._maybe_drain(1)

    def queue_superoperator(self, sup: np.ndarray, axes: Sequence[int]) -> None:
        """A channel already in superoperator form (sum_i K_i (x) conj(K_i))."""
        self._host = None
        self._fuser.add(sup, self._row_bits(axes) + self._col_bits(axes))
        self._maybe_drain(1)

    def _maybe_drain(self, added: int) -> None:
        self._since_drain += added
        if self._since_drain < self._drain_every:
            return
        self._since_drain = 0
        ready = self._held + self._fuser.pop_final_blocks()
        # (a trailing block without a partner waits for the next batch: two blocks
        # share one pass over HBM, DeviceState.plan_passes)
        ready, self._held = self._dev.split_unpaired_tail(ready)
        if ready:
            self._dev.apply_batch(ready)
            self.passes += len(self._dev.plan_passes(ready))

    def flush(self) -> None:
        if len(self._fuser) == 0 and not self._held:
            return
        blocks = self._held + self._fuser.blocks()
        self._held = []
        self._fuser.clear()
        self._dev.apply_batch(blocks)
        self.passes += len(self._dev.plan_passes(blocks))

    @property
    def device_state(self) -> DeviceState:
        self.flush()
        return self._dev

    # ---- QuantumStateRepresentation ---------------------------------------------------------

    def copy(self, deep_copy_buffers: bool = True) -> 'B200DensityMatrix':
        self.flush()
        return B200DensityMatrix(self._dev.copy(), self._n, self._max_fused)

    def _marginal_device(self, axes: Sequence[int]):
        """Unnormalised float64 marginal of Re diag(rho) over `axes` (first axis
        = MSB), on the device: ``_probs`` of sim/density_matrix_utils.py:185-192."""
        diag = self._dev.dm_diagonal_device()
        return DeviceState.probs_marginal_device(diag, self._n, self._col_bits(axes))

    def measure(self, axes: Sequence[int], seed=None) -> list[int]:
        """``measure_density_matrix`` (sim/density_matrix_utils.py:97-182)."""
        axes = list(axes)
        if not axes:
            return []
        self.flush()
        self._host = None
        prng = value.parse_random_state(seed)
        m = len(axes)
        u = float(prng.random_sample())
        probs = self._marginal_device(axes)
        pick = int(DeviceState.cdf_sample_device(probs, np.array([u])).cpu()[0])
        values = [(pick >> (m - 1 - q)) & 1 for q in range(m)]
        p = np.clip(probs.cpu().numpy(), 0, None)
        self._dev.dm_collapse(self._col_bits(axes), values, float(p[pick] / p.sum()))
        return values

    def sample(self, axes: Sequence[int], repetitions: int = 1, seed=None,
               out_columns: Sequence[int] | None = None) -> np.ndarray:
        """``sample_density_matrix`` (sim/density_matrix_utils.py:31-94);
        `out_columns` reorders the result's columns on the device."""
        if repetitions < 0:
            raise ValueError(f'Number of repetitions cannot be negative. Was {repetitions}')
        axes = [int(a) for a in axes]
        for a in axes:
            if a < 0 or a >= self._n:
                raise IndexError(f'Out of range indices in {axes}, must be less than {self._n}')
        m = len(axes)
        if repetitions == 0 or m == 0:
            return np.zeros(shape=(repetitions, m), dtype=np.int8)
        self.flush()
        prng = value.parse_random_state(seed)
        uniforms = prng.random_sample(repetitions)
        probs = self._marginal_device(axes)
        idx = DeviceState.cdf_sample_device(probs, uniforms)
        cols = list(range(m)) if out_columns is None else [int(c) for c in out_columns]
        bits = DeviceState.unpack_bits_device(idx, [m - 1 - c for c in cols])
        return bits.cpu().numpy().astype(np.int8)

    @property
    def supports_factor(self) -> bool:
        return True

    @property
    def can_represent_mixed_states(self) -> bool:
        return True

    # ---- layout: Kronecker product, factoring, axis order (split_untangled_states) ----------

    def kron(self, other: 'B200DensityMatrix') -> 'B200DensityMatrix':
        """``density_matrix_kronecker_product`` (linalg/transformations.py:616-644):
        vec(rho) (x) vec(sigma) has bits [r1 c1 r2 c2]; regroup to [r1 r2 c1 c2]."""
        n1, n2 = self._n, other._n
        k = self.device_state.kron(other.device_state)
        src = [0] * (2 * (n1 + n2))
        for j in range(n2):  # c2
            src[j] = j
        for j in range(n1):  # c1
            src[n2 + j] = 2 * n2 + j
        for j in range(n2):  # r2
            src[n1 + n2 + j] = n2 + j
        for j in range(n1):  # r1
            src[n1 + 2 * n2 + j] = 2 * n2 + n1 + j
        dev = k if src == list(range(len(src))) else k.permute_bits(src)
        out = B200DensityMatrix(dev, n1 + n2, self._max_fused)
        out.passes = self.passes + other.passes
        return out

    def reindex(self, axes: Sequence[int]) -> 'B200DensityMatrix':
        """``transpose_density_matrix_to_axis_order``: new axis k = old axis axes[k],
        on rows and columns alike."""
        axes = [int(a) for a in axes]
        n = self._n
        if axes == list(range(n)):
            return B200DensityMatrix(self.device_state.copy(), n, self._max_fused)
        src = [0] * (2 * n)
        for k, a in enumerate(axes):
            src[n - 1 - k] = n - 1 - a  # column bits
            src[2 * n - 1 - k] = 2 * n - 1 - a  # row bits
        out = B200DensityMatrix(self.device_state.permute_bits(src), n, self._max_fused)
        out.passes = self.passes
        return out

    def factor(self, axes: Sequence[int], *, validate=True, atol=1e-07):
        """``factor_density_matrix`` (linalg/transformations.py:694-727): the two
        factors are the partial traces onto `axes` and onto the rest."""
        axes = [int(a) for a in axes]
        n = self._n
        rest = [a for a in range(n) if a not in axes]
        dev = self.device_state
        extracted = dev.dm_partial_trace([n - 1 - a for a in axes])
        remainder = dev.dm_partial_trace([n - 1 - a for a in rest])
        e_state = B200DensityMatrix(extracted, len(axes), self._max_fused)
        r_state = B200DensityMatrix(remainder, len(rest), self._max_fused)
        r_state.passes = self.passes
        if validate:
            product = e_state.kron(r_state)  # axes order: axes + rest
            order = axes + rest
            inverse = [order.index(a) for a in range(n)]
            back = product.reindex(inverse)
            if not back.device_state.allclose(dev, atol):
                raise ValueError('The tensor cannot be factored by the requested axes')
        return e_state, r_state

    def scale(self, factor: complex) -> None:
        self.flush()
        self._host = None
        self._dev.scale(factor)

    def to_numpy_tensor(self) -> np.ndarray:
        """Host copy as a ``(2,)*2n`` tensor, cached until the state changes."""
        if self._host is None:
            self.flush()
            self._host = self._dev.to_numpy().reshape(self._qid_shape * 2)
        return self._host


class B200DensityMatrixSimulationState(SimulationState[B200DensityMatrix]):
    """Replaces ``DensityMatrixSimulationState``
    (sim/density_matrix_simulation_state.py:252-354)."""

    def __init__(
        self,
        *,
        prng=None,
        qubits=None,
        initial_state: Any = 0,
        dtype=np.complex64,
        classical_data=None,
        max_fused_qubits: int | None = None,
    ):
        qubits = tuple(qubits) if qubits is not None else ()
        if isinstance(initial_state, B200DensityMatrix):
            state = initial_state  # an existing device state, adopted as is
        else:
            state = B200DensityMatrix.create(
                initial_state=initial_state,
                qid_shape=tuple(q.dimension for q in qubits),
                dtype=dtype,
                max_fused_qubits=max_fused_qubits,
            )
        super().__init__(state=state, prng=prng, qubits=qubits, classical_data=classical_data)
        self._dtype = dtype
        self._max_fused_qubits = max_fused_qubits

    def add_qubits(self, qubits):
        """(density_matrix_simulation_state.py:289-295)"""
        ret = super().add_qubits(qubits)
        if ret is not NotImplemented:
            return ret
        fresh = type(self)(
            qubits=qubits, dtype=self._dtype, prng=self._prng,
            max_fused_qubits=self._max_fused_qubits,
        )
        return self.kronecker_product(fresh, inplace=True)

    def remove_qubits(self, qubits):
        """(density_matrix_simulation_state.py:297-303)"""
        ret = super().remove_qubits(qubits)
        if ret is not NotImplemented:
            return ret
        extracted, remainder = self.factor(qubits, inplace=True)
        remainder._state.scale(complex(extracted._state.device_state.amplitudes([0])[0]))
        return remainder

    def _act_on_fallback_(self, action: Any, qubits, allow_decompose: bool = True) -> bool:
        strats = [_strat_apply_channel]
        if allow_decompose:
            strats.append(strat_act_on_from_apply_decompose)
        for strat in strats:
            result = strat(action, self, qubits)
            if result is True:
                return True
            assert result is NotImplemented, str(result)
        raise TypeError(
            "Can't simulate operations that don't implement "
            "SupportsUnitary, SupportsConsistentApplyUnitary, "
            f"SupportsMixture or SupportsKraus or is a measurement: {action!r}"
        )

    @property
    def target_tensor(self) -> np.ndarray:
        return self._state.to_numpy_tensor()

    @property
    def device_state(self) -> DeviceState:
        return self._state.device_state

    @property
    def qid_shape(self) -> tuple[int, ...]:
        return self._state._qid_shape

    def __repr__(self) -> str:
        return (
            'cirq_b200.B200DensityMatrixSimulationState('
            f'qubits={self.qubits!r}, classical_data={self.classical_data!r})'
        )


def can_decompose(action: Any, qubits) -> bool:
    if isinstance(action, ops.Gate):
        return protocols.decompose_once_with_qubits(action, qubits, None) is not None
    return protocols.decompose_once(action, None) is not None


def _strat_apply_channel(action: Any, args: B200DensityMatrixSimulationState, qubits) -> bool:
    """Strategy order of ``protocols.apply_channel``
    (protocols/apply_channel_protocol.py:239-262): unitary first, else Kraus
    operators (``kraus`` also covers mixtures, protocols/kraus_protocol.py:138-227)."""
    n = len(qubits)
    axes = args.get_axes(qubits)
    wide = n > _MAX_DIRECT_QUBITS
    if protocols.has_unitary(action) and not (wide and not hasattr(action, '_unitary_')):
        u = protocols.unitary(action, None)
        if u is not None:
            args._state.queue_unitary(u, axes)
            return True
    if wide and can_decompose(action, qubits):
        return NotImplemented
    ks = protocols.kraus(action, default=None)
    if ks is None:
        return NotImplemented
    args._state.queue_kraus(ks, axes)
    return True


_FORM_CACHE: dict = {}
_FORM_CACHE_MAX = 4096


def cached_channel_form(op) -> tuple | None:
    """('unitary', U) or ('super', sum_i K_i (x) conj(K_i)) for a plain gate operation
    (tags looked through) on at most 3 qubits, cached per gate — what
    ``_strat_apply_channel`` computes, once instead of per operation (a noise model
    inserts the same channel hundreds of times).  None for everything that must
    take ``protocols.act_on``: measurements, resets (the product state factors the
    qubit out afterwards), classical control, composite or symbolic operations."""
    from cirq_b200.sv_simulator import PLAIN_GATE_OPERATIONS

    base = op.untagged if type(op) is ops.TaggedOperation else op
    if type(base) not in PLAIN_GATE_OPERATIONS:
        return None
    gate = base.gate
    try:
        hit = _FORM_CACHE.get(gate)
    except TypeError:  # unhashable gate
        return None
    if hit is not None:
        return hit or None
    form: Any = False
    n = len(base.qubits)
    if (
        0 < n <= _MAX_DIRECT_QUBITS
        and not isinstance(gate, (ops.MeasurementGate, ops.ResetChannel, ops.IdentityGate, ops.SwapPowGate))
        and not protocols.is_parameterized(gate)
        and not protocols.is_measurement(base)
        and all(d == 2 for d in protocols.qid_shape(base))
    ):
        u = protocols.unitary(base, None) if protocols.has_unitary(base) else None
        if u is not None:
            form = ('unitary', np.asarray(u, dtype=np.complex128))
        else:
            ks = protocols.kraus(base, default=None)
            if ks is not None:
                form = ('super', sum(np.kron(k, np.conj(k)) for k in ks))
    if len(_FORM_CACHE) >= _FORM_CACHE_MAX:
        _FORM_CACHE.clear()
    _FORM_CACHE[gate] = form
    return form or None


class B200DensityMatrixStepResult(_FastConfuseMixin, _ref_dm.DensityMatrixStepResult):
    """Step result; ``density_matrix()`` downloads rho on first use."""


class B200DensityMatrixTrialResult(_ref_dm.DensityMatrixTrialResult):
    @property
    def device_state(self) -> DeviceState:
        """rho in HBM as complex[4^n] (row-major)."""
        return self._get_merged_sim_state().device_state


class B200DensityMatrixSimulator(
    simulator_base.SimulatorBase[
        'B200DensityMatrixStepResult',
        'B200DensityMatrixTrialResult',
        'B200DensityMatrixSimulationState',
    ],
    simulator.SimulatesExpectationValues,
):
    """Drop-in for ``cirq.DensityMatrixSimulator`` (sim/density_matrix_simulator.py:32-262)."""

    def __init__(
        self,
        *,
        dtype=np.complex64,
        noise: 'cirq.NOISE_MODEL_LIKE' = None,
        seed: 'cirq.RANDOM_STATE_OR_SEED_LIKE' = None,
        split_untangled_states: bool = True,
        max_fused_qubits: int | None = None,
        sweep_batch: bool = False,
    ):
        super().__init__(
            dtype=dtype, noise=noise, seed=seed, split_untangled_states=split_untangled_states
        )
        if dtype not in {np.complex64, np.complex128}:
            raise ValueError(f'dtype must be complex64 or complex128, was {dtype}')
        self._max_fused = None if max_fused_qubits is None else int(max_fused_qubits)
        # True: run_sweep advances all resolvers as one device array (cirq_b200.sweeps)
        self._sweep_batch = bool(sweep_batch)
        self.last_run_info: dict = {}

    def _state_from_device(self, dev: DeviceState, qubits):
        """Simulation state adopting an existing device array (cirq_b200.sweeps)."""
        return B200DensityMatrixSimulationState(
            qubits=qubits, prng=self._prng, dtype=self._dtype, max_fused_qubits=self._max_fused,
            initial_state=B200DensityMatrix(dev, len(qubits), self._max_fused),
        )

    def simulate_sweep_iter(self, program, params, qubit_order=ops.QubitOrder.DEFAULT, initial_state=None):
        """``SimulatorBase.simulate_sweep_iter`` (sim/simulator_base.py:277-320); with
        ``sweep_batch=True`` a measurement-free sweep from |0...0> in the default
        qubit order advances all resolvers as one device array."""
        if self._sweep_batch and initial_state is None and qubit_order is ops.QubitOrder.DEFAULT:
            from cirq_b200 import sweeps

            batched = sweeps.simulate_sweep_batched(self, 'dm', program, params, DeviceState)
            if batched is not None:
                yield from batched
                return
        yield from super().simulate_sweep_iter(program, params, qubit_order, initial_state)

    def run_sweep_iter(self, program, params, repetitions: int = 1):
        """``SimulatesSamples.run_sweep_iter`` (sim/simulator.py:62-94), optionally
        batched over the resolvers."""
        if self._sweep_batch:
            from cirq_b200 import sweeps

            batched = sweeps.run_sweep_batched(self, 'dm', program, params, repetitions, DeviceState)
            if batched is not None:
                yield from batched
                return
        yield from super().run_sweep_iter(program, params, repetitions)

    def _create_partial_simulation_state(self, initial_state, qubits, classical_data):
        if isinstance(initial_state, B200DensityMatrixSimulationState):
            return initial_state
        return B200DensityMatrixSimulationState(
            qubits=qubits,
            prng=self._prng,
            classical_data=classical_data,
            initial_state=initial_state,
            dtype=self._dtype,
            max_fused_qubits=self._max_fused,
        )

    def _create_simulation_state(self, initial_state, qubits):
        return create_product_state(self, initial_state, qubits)

    def _can_be_in_run_prefix(self, val: Any):
        return not protocols.measurement_keys_touched(val)

    def _create_step_result(self, sim_state):
        return B200DensityMatrixStepResult(sim_state=sim_state, dtype=self._dtype)

    def _create_simulator_trial_result(self, params, measurements, final_simulator_state):
        return B200DensityMatrixTrialResult(
            params=params, measurements=measurements, final_simulator_state=final_simulator_state
        )

    def _core_iterator(self, circuit, sim_state, all_measurements_are_terminal: bool = False):
        """``SimulatorBase._core_iterator`` (sim/simulator_base.py:169-214) with the
        shortcut of ``B200Simulator._core_iterator``: a plain unitary gate or channel
        on at most 3 qubits goes straight into the density matrix's queue as U (x)
        conj(U) / its cached superoperator, instead of through ``protocols.act_on``
        and a fresh ``protocols.kraus`` + Kronecker products per operation."""
        import collections

        from cirq_b200.sv_simulator import B200ProductState

        if len(circuit) == 0:
            yield self._create_step_result(sim_state)
            return
        noisy_moments = self.noise.noisy_moments(circuit, sorted(circuit.all_qubits()))
        measured: dict = collections.defaultdict(bool)
        product = isinstance(sim_state, B200ProductState)
        dense = isinstance(sim_state, B200DensityMatrixSimulationState)
        for moment in noisy_moments:
            for op in ops.flatten_to_ops(moment):
                try:
                    if all_measurements_are_terminal and measured[op.qubits]:
                        continue
                    if isinstance(op.gate, ops.MeasurementGate):
                        measured[op.qubits] = True
                        if all_measurements_are_terminal:
                            continue
                    form = cached_channel_form(op) if (product or dense) else None
                    if form is not None:
                        target = sim_state.join_for(op.qubits) if product else sim_state
                        axes = target.get_axes(op.qubits)
                        if form[0] == 'unitary':
                            target._state.queue_unitary(form[1], axes)
                        else:
                            target._state.queue_superoperator(form[1], axes)
                        continue
                    protocols.act_on(op, sim_state)
                except TypeError:
                    raise TypeError(f"{self.__class__.__name__} doesn't support {op!r}")
            yield self._create_step_result(sim_state)

    def simulate_expectation_values_sweep(
        self,
        program,
        observables,
        params,
        qubit_order=ops.QubitOrder.DEFAULT,
        initial_state=None,
        permit_terminal_measurements: bool = False,
    ) -> list[list[float]]:
        """tr(rho O) — sim/density_matrix_simulator.py:204-235."""
        if not permit_terminal_measurements and program.are_any_measurements_terminal():
            raise ValueError(
                'Provided circuit has terminal measurements, which may '
                'skew expectation values. If this is intentional, set '
                'permit_terminal_measurements=True.'
            )
        swept_evs = []
        if not isinstance(observables, list):
            observables = [observables]
        pslist = [ops.PauliSum.wrap(pslike) for pslike in observables]
        if self._sweep_batch and initial_state is None and qubit_order is ops.QubitOrder.DEFAULT:
            from cirq_b200 import sweeps

            batched = sweeps.expectation_sweep_batched(
                self, 'dm', program, pslist, params, DeviceState, dm_pauli_sum_expectation)
            if batched is not None:
                return batched
        qubit_order = ops.QubitOrder.as_qubit_order(qubit_order)
        qmap = {q: i for i, q in enumerate(qubit_order.order_for(program.all_qubits()))}
        for param_resolver in study.to_resolvers(params):
            result = self.simulate(
                program, param_resolver, qubit_order=qubit_order, initial_state=initial_state
            )
            # tr(rho P) per Pauli string on the device (2^n entries of rho each)
            # instead of the reference's einsum over a host copy of rho
            dev = result.device_state
            swept_evs.append([dm_pauli_sum_expectation(dev, obs, qmap) for obs in pslist])
        return swept_evs


def dm_pauli_sum_expectation(dev: DeviceState, pauli_sum, qubit_map) -> float:
    """sum_k c_k tr(rho P_k) (ops/linear_combinations.py expectation_from_density_matrix:
    the real part of each term, as ops/pauli_string.py:732 returns)."""
    from cirq_b200.sv_simulator import pauli_masks

    n = dev.n_bits // 2
    total = 0.0
    for ps in pauli_sum:
        if abs(complex(ps.coefficient).imag) > 0.0001:
            raise NotImplementedError(
                'Cannot compute expectation value of a non-Hermitian '
                f'PauliString <{ps}>. Coefficient must be real.'
            )
        x, z = pauli_masks(ps, qubit_map, n)
        total += complex(ps.coefficient).real * dev.dm_pauli_expectation(x, z).real
    return total
