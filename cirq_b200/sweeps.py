"""Sweep-level batching: all resolvers of a ``run_sweep`` advance together.

``SimulatesSamples.run_sweep_iter`` (cirq-core/cirq/sim/simulator.py:62-94)
re-simulates the circuit from scratch for every ParamResolver.  For the small
states typical of variational sweeps (QAOA / VQE parameter scans on 8-20 qubits)
each of those runs is a few hundred tiny launches plus the same host-side walk
over the circuit.  Here the P resolvers are laid out as ONE device array — state
of resolver i in index bits above the state's own bits, exactly like the batched
trajectories of ``cirq_b200.trajectories`` — and the circuit is walked once:

* an operation without symbols is the same matrix for every resolver: it goes
  through the ordinary fuser and gate kernels (one launch serves all resolvers);
* an operation with symbols becomes a table of matrices, one per distinct
  resolved value, applied by ``b2q_bsv_apply_select`` with choice[i] = row of
  resolver i (EigenGate families are resolved in one vectorised expression from
  their eigen-decomposition, ops/eigen_gate.py:295-310);
* density matrices use the same array with 2n state bits: U becomes
  U (x) conj(U) on (row bits, column bits), a channel its superoperator
  (the single-pass form of protocols/apply_channel_protocol.py:297-356).

Evolution is deterministic, so each resolver's final state equals the one the
per-resolver loop produces; sampling then runs per resolver, in resolver order,
through the unchanged ``sample_measurement_ops`` path — seeded results are those
of the sequential sweep.  Anything outside the supported shape (measurements
that are not terminal, classical control, stochastic operations on a state
vector, noise models other than a constant per-qubit channel, states too large
to batch) returns None and the caller falls back to the reference loop.
"""
from __future__ import annotations

import numpy as np

from cirq_b200._cirq_compat import import_cirq
from cirq_b200.fusion import fuser_for

cirq = import_cirq()
from cirq import devices, ops, protocols, study  # noqa: E402
from cirq.sim.simulator import split_into_matching_protocol_then_general  # noqa: E402

# largest batch array (state bits + resolver bits): 2^30 complex64 = 8.6 GB
MAX_BATCH_ARRAY_BITS = 30
# widest per-resolver operator the select kernel takes (index bits)
MAX_SELECT_BITS = 4


def _resolved_matrices(op, resolvers) -> np.ndarray | None:
    """complex128[P, d, d]: the unitary of `op` under each resolver, or None if
    some resolution has no unitary."""
    base = op.untagged
    gate = getattr(base, 'gate', None)
    if type(base) is ops.GateOperation and isinstance(gate, ops.EigenGate):
        shift = gate._global_shift
        if not protocols.is_parameterized(shift):
            try:
                exps = np.array([complex(r.value_of(gate._exponent, recursive=True)) for r in resolvers])
            except TypeError:
                exps = None
            if exps is not None and np.all(exps.imag == 0):
                e = exps.real
                total = None
                try:
                    # (eigen-components may themselves depend on another symbol, e.g.
                    # PhasedISwapPowGate(phase_exponent=a): then the generic loop below)
                    for half_turns, component in gate._eigen_components():
                        phase = np.exp(1j * np.pi * e * (float(half_turns) + float(shift)))
                        term = phase[:, None, None] * np.asarray(component, dtype=np.complex128)[None]
                        total = term if total is None else total + term
                except (TypeError, ValueError):
                    total = None
                if total is not None:
                    return total
    names = sorted(protocols.parameter_names(op))
    memo: dict = {}
    out = []
    for r in resolvers:
        key = tuple(r.value_of(name, recursive=True) for name in names)
        try:
            u = memo.get(key)
        except TypeError:
            key, u = None, None
        if u is None:
            u = protocols.unitary(protocols.resolve_parameters(op, r), None)
            if u is None:
                return None
            if key is not None:
                memo[key] = u
        out.append(u)
    return np.asarray(out, dtype=np.complex128)


def _resolved_superoperators(op, resolvers) -> np.ndarray | None:
    """complex128[P, d^2, d^2]: sum_k K (x) conj(K) of a symbolic channel."""
    out = []
    for r in resolvers:
        kraus = protocols.kraus(protocols.resolve_parameters(op, r), None)
        if kraus is None:
            return None
        out.append(sum(np.kron(k, np.conj(k)) for k in kraus))
    return np.asarray(out, dtype=np.complex128)


class SweepPlan:
    """Host-side walk of the circuit: [('shared', matrix, bits) |
    ('select', table[count, d, d], bits, choice[P])] over the index bits of one
    resolver's state, plus the terminal measurement operations."""

    def __init__(self, kind: str, qubits, resolvers):
        self.kind = kind  # 'sv' | 'dm'
        self.qubits = tuple(qubits)
        self.n = len(self.qubits)
        self.state_bits = self.n if kind == 'sv' else 2 * self.n
        self.resolvers = list(resolvers)
        self.items: list[tuple] = []
        self.measurement_ops: list = []
        self.scalar = 1.0 + 0.0j  # product of the zero-qubit (global phase) operations
        self._axis = {q: i for i, q in enumerate(self.qubits)}
        self._channel_cache: dict = {}

    def _bits(self, op):
        axes = [self._axis[q] for q in op.qubits]
        col = [self.n - 1 - a for a in axes]
        if self.kind == 'sv':
            return col, None
        return [2 * self.n - 1 - a for a in axes], col

    def add_op(self, op) -> bool:
        from cirq_b200.sv_simulator import cached_unitary

        if any(d != 2 for d in protocols.qid_shape(op)) or protocols.control_keys(op):
            return False
        if protocols.is_measurement(op):
            return False
        row, col = self._bits(op)
        if not row:
            # a global phase: no effect on samples or on rho; a final STATE VECTOR
            # carries it (simulate_sweep), so the scalar is kept for the end
            if protocols.is_parameterized(op):
                return False
            if self.kind == 'sv':
                u = protocols.unitary(op, None)
                if u is None:
                    return False
                self.scalar *= complex(np.asarray(u).reshape(-1)[0])
            return True
        if not protocols.is_parameterized(op):
            u = cached_unitary(op)
            if u is not None:
                self.items.append(('shared', np.asarray(u, dtype=np.complex128), row))
                if col is not None:
                    self.items.append(('shared', np.conj(u), col))
                return True
            if self.kind == 'sv' or len(row) > 3:
                return False  # stochastic on a pure state / too wide a channel
            gate = getattr(op.untagged, 'gate', None)
            try:
                sup = self._channel_cache.get(gate) if gate is not None else None
            except TypeError:
                gate, sup = None, None
            if sup is None:
                kraus = protocols.kraus(op, None)
                if kraus is None:
                    return False
                sup = sum(np.kron(k, np.conj(k)) for k in kraus)
                if gate is not None:
                    self._channel_cache[gate] = sup
            self.items.append(('shared', sup, row + col))
            return True
        # symbols: one matrix per resolver
        width = len(row) * (1 if self.kind == 'sv' else 2)
        if width > MAX_SELECT_BITS:
            return False
        mats = _resolved_matrices(op, self.resolvers)
        if mats is not None:
            if self.kind == 'dm':
                mats = np.einsum('pab,pcd->pacbd', mats, np.conj(mats)).reshape(
                    len(mats), 1 << width, 1 << width)
        elif self.kind == 'dm':
            mats = _resolved_superoperators(op, self.resolvers)
        if mats is None:
            return False
        bits = row if col is None else row + col
        table, choice = np.unique(mats.reshape(len(mats), -1), axis=0, return_inverse=True)
        d = 1 << width
        if len(table) == 1:
            self.items.append(('shared', table[0].reshape(d, d), bits))
        else:
            self.items.append(('select', table.reshape(-1, d, d), bits,
                               np.asarray(choice, dtype=np.int32).reshape(-1)))
        return True


def plan_sweep(simulator, kind: str, program, resolvers, sampled: bool = True) -> SweepPlan | None:
    """The batched schedule of `program` for `resolvers`, or None when the sweep
    has to take the per-resolver loop.  `sampled`: a run_sweep (the circuit must
    end in measurements, which are sampled); otherwise a final-state sweep (the
    circuit must not measure at all)."""
    noise = simulator.noise
    if noise is not devices.NO_NOISE and not isinstance(noise, devices.ConstantQubitNoiseModel):
        return None
    if kind == 'sv' and noise is not devices.NO_NOISE:
        return None
    if len(resolvers) < 2:
        return None
    qubits = tuple(sorted(program.all_qubits()))
    plan = SweepPlan(kind, qubits, resolvers)
    b = (len(resolvers) - 1).bit_length()
    if not qubits or plan.state_bits + b > MAX_BATCH_ARRAY_BITS:
        return None
    # every symbol must be resolved by every resolver (simulator.py:942-949)
    names = protocols.parameter_names(program)
    for r in resolvers:
        for name in names:
            if protocols.is_parameterized(r.value_of(name, recursive=True)):
                return None
    # The reference walks the circuit in two parts and hands the noise model each
    # part's own moments and qubits, so the split decides where noise lands:
    #  * run_sweep (simulator_base.py:224-244): a prefix without measurements, then
    #    a suffix that must consist of measurements only;
    #  * simulate_sweep (simulator_base.py:304-320): a prefix of operations without
    #    symbols (shared by the resolvers), then the rest.
    if sampled:
        prefix, suffix = split_into_matching_protocol_then_general(
            program, lambda op: not protocols.measurement_keys_touched(op))
        suffix_ops = list(suffix.all_operations())
        if not suffix_ops or not all(isinstance(op.gate, ops.MeasurementGate) for op in suffix_ops):
            return None
    else:
        if program.has_measurements():
            return None
        prefix, suffix = split_into_matching_protocol_then_general(
            program, lambda op: not protocols.is_parameterized(op))
        suffix_ops = []
    for part, skip_measurements in ((prefix, False), (suffix, True)):
        if len(part) == 0:
            continue
        # (the reference hands the noise model the qubits of the PART it is walking,
        # sim/simulator_base.py:196: a qubit that is only measured gets no noise
        # during the prefix)
        # (all_measurements_are_terminal walk of sim/simulator_base.py:196-209: once a
        # qubit tuple has been measured, every later operation on exactly that tuple —
        # the noise the model adds behind a terminal measurement — is skipped too)
        measured: dict = {}
        for moment in noise.noisy_moments(part, sorted(part.all_qubits())):
            for op in ops.flatten_to_ops(moment):
                if skip_measurements:
                    if measured.get(op.qubits):
                        continue
                    if isinstance(op.gate, ops.MeasurementGate):
                        measured[op.qubits] = True
                        continue
                if not plan.add_op(op):
                    return None
    plan.measurement_ops = suffix_ops
    return plan


def evolve_sweep(simulator, plan: SweepPlan, device_state_cls, info: dict | None = None):
    """Runs the gates of a SweepPlan; returns the batch array (state of resolver
    i at offset i << plan.state_bits)."""
    P = len(plan.resolvers)
    b = (P - 1).bit_length()
    count = 1 << b
    dtype = np.dtype(simulator._dtype)
    sb = plan.state_bits
    zero = device_state_cls.basis(sb, dtype, 0)
    dev = device_state_cls.from_numpy(np.ones(count, dtype=dtype), dtype).kron(zero) if b else zero
    del zero
    max_fused = simulator._max_fused
    if plan.kind == 'dm' and max_fused is None:
        max_fused = 4
    fuser = fuser_for(dtype, max_fused, sb + b, state_vector=plan.kind == 'sv')
    passes = 0

    def flush():
        nonlocal passes
        if fuser.pending:  # (a trailing relabelled SWAP counts: blocks() puts it back)
            blocks = fuser.blocks()
            fuser.clear()
            dev.apply_batch(blocks)
            passes += len(blocks)

    for item in plan.items:
        if item[0] == 'shared':
            fuser.add(item[1], item[2])
        else:
            flush()
            choice = np.zeros(count, dtype=np.int32)
            choice[:P] = item[3]
            dev.bsv_apply_select(sb, item[1], item[2], choice)
            passes += 1
    flush()
    if plan.scalar != 1.0:
        dev.scale(plan.scalar)
    if info is not None:
        info.update(path='batched sweep', resolvers=P, batch_bits=b, passes=passes,
                    select_passes=sum(1 for it in plan.items if it[0] == 'select'))
    return dev


def resolver_state(dev, plan: SweepPlan, i: int):
    """The state of resolver i inside the batch array, as a device state of its own."""
    if dev.n_bits == plan.state_bits:
        return dev
    return dev.slice_copy(i << plan.state_bits, plan.state_bits)


def execute_sweep(simulator, plan: SweepPlan, repetitions: int, device_state_cls, info: dict | None = None):
    """Runs a SweepPlan; yields one {key: records} dict per resolver, in order."""
    dev = evolve_sweep(simulator, plan, device_state_cls, info)
    for i in range(len(plan.resolvers)):
        sim_state = simulator._state_from_device(resolver_state(dev, plan, i), plan.qubits)
        step = simulator._create_step_result(sim_state)
        yield step.sample_measurement_ops(
            plan.measurement_ops, repetitions, seed=simulator._prng, _allow_repeated=True)


def run_sweep_batched(simulator, kind: str, program, params, repetitions: int, device_state_cls):
    """Iterator of cirq.ResultDict for the whole sweep, or None if it cannot be
    batched (the caller then runs the reference loop)."""
    resolvers = list(study.to_resolvers(params))
    if repetitions <= 0 or not program.has_measurements():
        return None
    plan = plan_sweep(simulator, kind, program, resolvers)
    if plan is None:
        return None

    def results():
        info: dict = {}
        for r, records in zip(resolvers, execute_sweep(simulator, plan, repetitions, device_state_cls, info)):
            simulator.last_run_info = info
            yield study.ResultDict(params=r, records=records)

    return results()


def expectation_sweep_batched(simulator, kind: str, program, pauli_sums, params, device_state_cls,
                              expectation):
    """[[<O_j> for j] for each resolver] with all resolvers evolved as one device
    array, or None if the sweep cannot be batched.  `expectation(dev, pauli_sum,
    qubit_map)` evaluates one observable on one resolver's device state (the
    simulators' own reduction kernels).  Replaces the per-resolver loops of
    sim/sparse_simulator.py:193-218 and sim/density_matrix_simulator.py:204-235."""
    resolvers = list(study.to_resolvers(params))
    plan = plan_sweep(simulator, kind, program, resolvers, sampled=False)
    if plan is None:
        return None
    qmap = {q: i for i, q in enumerate(plan.qubits)}
    if any(q not in qmap for obs in pauli_sums for q in obs.qubits):
        return None
    info: dict = {}
    dev = evolve_sweep(simulator, plan, device_state_cls, info)
    simulator.last_run_info = info
    out = []
    for i in range(len(resolvers)):
        piece = resolver_state(dev, plan, i)
        out.append([expectation(piece, obs, qmap) for obs in pauli_sums])
    return out


def simulate_sweep_batched(simulator, kind: str, program, params, device_state_cls):
    """Iterator of trial results (one per resolver, final state = that resolver's
    slice of the batch array), or None if the sweep cannot be batched.  Replaces
    the per-resolver loop of sim/simulator.py:585-606 for measurement-free
    circuits from |0...0> in the default qubit order."""
    resolvers = list(study.to_resolvers(params))
    plan = plan_sweep(simulator, kind, program, resolvers, sampled=False)
    if plan is None:
        return None

    def results():
        info: dict = {}
        dev = evolve_sweep(simulator, plan, device_state_cls, info)
        simulator.last_run_info = info
        for i, r in enumerate(resolvers):
            sim_state = simulator._state_from_device(resolver_state(dev, plan, i), plan.qubits)
            yield simulator._create_simulator_trial_result(
                params=r, measurements={}, final_simulator_state=sim_state)

    return results()
