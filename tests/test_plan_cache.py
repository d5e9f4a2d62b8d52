"""cirq_b200.plan_cache: the device schedule of a circuit's unitary prefix is kept
from the second execution of the same circuit on.  A cached call must give what
the uncached call gives (and what cirq.Simulator gives, seeded samples included),
consume random numbers identically, and never serve a stale schedule."""
import numpy as np
import pytest

from fake_device import OracleDeviceState


@pytest.fixture(params=['oracle', pytest.param('cuda', marks=pytest.mark.gpu)])
def SV(request, cirq, monkeypatch):
    import cirq_b200
    import cirq_b200.sv_simulator as svm
    from cirq_b200 import plan_cache
    from cirq_b200.device_state import DeviceState

    monkeypatch.setattr(svm, 'DeviceState', OracleDeviceState if request.param == 'oracle' else DeviceState)
    monkeypatch.delenv('CIRQ_B200_PLAN_CACHE', raising=False)
    monkeypatch.delenv('CIRQ_B200_PLAN_CACHE_EAGER', raising=False)
    plan_cache.CACHE.clear()
    yield cirq_b200.B200Simulator
    plan_cache.CACHE.clear()


def _circuit(cirq, qubits, depth, seed, extras=True):
    rng = np.random.RandomState(seed)
    c = cirq.Circuit()
    two = [cirq.CZ, cirq.ISWAP, cirq.SWAP, cirq.CNOT, cirq.SWAP ** 0.5, cirq.FSimGate(0.3, 0.7)]
    for d in range(depth):
        for q in qubits:
            if rng.rand() < 0.8:
                c.append(cirq.PhasedXPowGate(phase_exponent=rng.rand(), exponent=rng.rand())(q))
        for i in range(d % 2, len(qubits) - 1, 2):
            if rng.rand() < 0.7:
                c.append(two[rng.randint(len(two))](qubits[i], qubits[i + 1]))
        if extras and d == 2:
            c.append(cirq.global_phase_operation(np.exp(0.3j)))
            c.append(cirq.I(qubits[0]))
            c.append(cirq.CCZ(qubits[0], qubits[2], qubits[4]).with_tags('tagged'))
    return c


@pytest.mark.parametrize('split', [True, False], ids=['split', 'dense'])
@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_cached_simulate_equals_uncached_and_reference(cirq, SV, split, dtype):
    from cirq_b200 import plan_cache

    qubits = cirq.LineQubit.range(9)
    circuit = _circuit(cirq, qubits, 10, seed=1)
    want = cirq.Simulator(dtype=dtype, split_untangled_states=split).simulate(circuit).final_state_vector
    uncached = SV(dtype=dtype, split_untangled_states=split, plan_cache=False).simulate(circuit).final_state_vector
    atol = 1e-5 if dtype == np.complex64 else 1e-12
    np.testing.assert_allclose(uncached, want, atol=atol)
    assert plan_cache.CACHE.builds == 0
    for call in range(4):
        got = SV(dtype=dtype, split_untangled_states=split).simulate(circuit).final_state_vector
        np.testing.assert_allclose(got, want, atol=atol)
        # second sighting builds the schedule, later calls hit it
        assert plan_cache.CACHE.builds == (1 if call >= 1 else 0)
        assert plan_cache.CACHE.hits == max(0, call - 1)


def test_cached_run_gives_the_reference_seeded_samples(cirq, SV):
    from cirq_b200 import plan_cache

    qubits = cirq.LineQubit.range(8)
    circuit = _circuit(cirq, qubits, 8, seed=2) + cirq.Circuit(
        cirq.measure(*qubits[:5], key='a'), cirq.measure(*qubits[5:], key='b'))
    want = cirq.Simulator(seed=11, dtype=np.complex64).run(circuit, repetitions=200)
    for call in range(4):
        sim = SV(seed=11)
        got = sim.run(circuit, repetitions=200)
        for key in ('a', 'b'):
            np.testing.assert_array_equal(got.measurements[key], want.measurements[key])
        assert (sim.last_run_info.get('plan_cache') == 'hit') == (call >= 1)
    assert plan_cache.CACHE.builds == 1 and plan_cache.CACHE.hits == 2


def test_partly_entangled_register_keeps_its_factors(cirq, SV):
    """Qubits the prefix never couples stay separate sub-states in a replayed
    product state, in the reference's order (seeded samples are drawn per factor)."""
    qubits = cirq.LineQubit.range(10)
    circuit = cirq.Circuit(
        [cirq.H(q) for q in qubits[:7]], cirq.CNOT(qubits[0], qubits[1]), cirq.CNOT(qubits[1], qubits[2]),
        cirq.SWAP(qubits[2], qubits[8]), cirq.CZ(qubits[4], qubits[5]), cirq.X(qubits[9]) ** 0.3,
        cirq.measure(*qubits, key='m'),
    )
    want = cirq.Simulator(seed=4, dtype=np.complex64).run(circuit, repetitions=300).measurements['m']
    for _ in range(3):
        got = SV(seed=4).run(circuit, repetitions=300).measurements['m']
        np.testing.assert_array_equal(got, want)
    # the replayed state has the same factor structure as a live one
    body = circuit[:-1]
    live = SV(plan_cache=False).simulate(body)._final_simulator_state
    for _ in range(2):
        cached = SV().simulate(body)._final_simulator_state
    def shape(state):
        return [tuple(s.qubits) for s in dict.fromkeys(state.sim_states.values())]
    assert shape(cached) == shape(live)
    assert list(cached.sim_states.keys()) == list(live.sim_states.keys())


def test_rest_of_the_circuit_runs_after_a_replayed_prefix(cirq, SV):
    """Mid-circuit measurement, a channel-free tail and classical control after the
    unitary prefix: the tail goes through the normal loop on the replayed state."""
    qubits = cirq.LineQubit.range(6)
    circuit = _circuit(cirq, qubits, 6, seed=3) + cirq.Circuit(
        cirq.measure(qubits[0], key='a'), cirq.X(qubits[1]).with_classical_controls('a'),
        cirq.H(qubits[2]), cirq.CZ(qubits[2], qubits[3]), cirq.measure(qubits[3], key='b'),
    )
    for seed in (1, 2):
        want = cirq.Simulator(seed=seed, dtype=np.complex64).simulate(circuit)
        for _ in range(3):
            got = SV(seed=seed).simulate(circuit)
            assert {k: list(v) for k, v in got.measurements.items()} == {
                k: list(v) for k, v in want.measurements.items()}
            np.testing.assert_allclose(got.final_state_vector, want.final_state_vector, atol=1e-5)
    # run() with a non-terminal measurement does not take the sampling shortcut
    want = cirq.Simulator(seed=5, dtype=np.complex64).run(circuit, repetitions=20)
    for _ in range(3):
        got = SV(seed=5).run(circuit, repetitions=20)
        for key in ('a', 'b'):
            np.testing.assert_array_equal(got.measurements[key], want.measurements[key])


def test_qubit_order_and_initial_state_are_part_of_the_key(cirq, SV):
    from cirq_b200 import plan_cache

    qubits = cirq.LineQubit.range(5)
    circuit = _circuit(cirq, qubits, 5, seed=4, extras=False)
    orders = [qubits, qubits[::-1], qubits + [cirq.LineQubit(7)]]
    for order in orders:
        want = cirq.Simulator(dtype=np.complex64).simulate(circuit, qubit_order=order).final_state_vector
        for _ in range(3):
            got = SV().simulate(circuit, qubit_order=order).final_state_vector
            np.testing.assert_allclose(got, want, atol=1e-5)
    assert plan_cache.CACHE.builds == 3
    # a non-zero initial state is never served from the cache (schedules start at |0...0>)
    hits = plan_cache.CACHE.hits
    want = cirq.Simulator(dtype=np.complex64).simulate(circuit, initial_state=5).final_state_vector
    got = SV().simulate(circuit, initial_state=5).final_state_vector
    np.testing.assert_allclose(got, want, atol=1e-5)
    assert plan_cache.CACHE.hits == hits


def test_edited_circuit_is_a_different_circuit(cirq, SV):
    qubits = cirq.LineQubit.range(5)
    circuit = _circuit(cirq, qubits, 5, seed=5, extras=False)
    for _ in range(3):
        SV().simulate(circuit)
    circuit.append(cirq.CNOT(qubits[0], qubits[4]))
    circuit.insert(0, cirq.X(qubits[2]))
    circuit[3] = cirq.Moment(cirq.H(qubits[1]))
    want = cirq.Simulator(dtype=np.complex64).simulate(circuit).final_state_vector
    for _ in range(3):
        np.testing.assert_allclose(SV().simulate(circuit).final_state_vector, want, atol=1e-5)
    frozen = circuit.freeze()
    for _ in range(3):
        np.testing.assert_allclose(SV().simulate(frozen).final_state_vector, want, atol=1e-5)


def test_what_is_not_cached(cirq, SV, monkeypatch):
    import sympy

    from cirq_b200 import plan_cache

    qubits = cirq.LineQubit.range(4)
    circuit = _circuit(cirq, qubits, 4, seed=6, extras=False)
    # a noise model changes what is executed: never cached
    for _ in range(3):
        SV(noise=cirq.depolarize(0.01), seed=1).simulate(circuit)
    # parameterized circuits are resolved into new moments on every call
    sym = cirq.Circuit(cirq.X(qubits[0]) ** sympy.Symbol('t'), cirq.CZ(qubits[0], qubits[1]))
    for _ in range(3):
        SV().simulate(sym, param_resolver={'t': 0.5})
    with pytest.raises(ValueError, match='symbols'):
        SV().simulate(sym)
    # qudits are refused where they always were
    qutrit = cirq.Circuit(cirq.IdentityGate(qid_shape=(3,))(cirq.LineQid(0, 3)), cirq.X(qubits[0]))
    for _ in range(3):
        with pytest.raises(ValueError, match='qubits only'):
            SV().simulate(qutrit)
    # switched off by the constructor or the environment
    for _ in range(3):
        SV(plan_cache=False).simulate(circuit)
    monkeypatch.setenv('CIRQ_B200_PLAN_CACHE', '0')
    for _ in range(3):
        SV().simulate(circuit)
    assert plan_cache.CACHE.builds == 0 and plan_cache.CACHE.hits == 0


def test_cache_is_bounded(cirq, SV, monkeypatch):
    from cirq_b200 import plan_cache

    monkeypatch.setattr(plan_cache.PlanCache, 'MAX_ENTRIES', 3)
    qubits = cirq.LineQubit.range(3)
    circuits = [_circuit(cirq, qubits, 3, seed=s, extras=False) for s in range(6)]
    for c in circuits:
        for _ in range(2):
            SV().simulate(c)
    assert len(plan_cache.CACHE._plans) == 3
    assert len(plan_cache.CACHE._seen) <= 12
    want = cirq.Simulator(dtype=np.complex64).simulate(circuits[0]).final_state_vector
    np.testing.assert_allclose(SV().simulate(circuits[0]).final_state_vector, want, atol=1e-5)


def test_recorder_refuses_what_a_schedule_cannot_hold():
    from cirq_b200 import plan_cache

    rec = plan_cache.recorder_for(OracleDeviceState)
    a, b = rec.basis(2, np.complex64, 0), rec.basis(1, np.complex64, 1)
    c = a.kron(b)
    c.apply_batch([(np.eye(2), [1])])
    c.scale(1j)
    assert c.permute_bits_inplace([1, 0, 2]) == 0
    assert [op[0] for op in rec.ops] == ['basis', 'basis', 'kron', 'apply', 'scale', 'permute']
    assert c.n_bits == 3 and rec.ops[2][1:] == (c.ident, a.ident, b.ident)
    for reader in ('norm2', 'to_numpy', 'marginal_probs', 'copy', 'collapse'):
        with pytest.raises(plan_cache.Untraceable):
            getattr(c, reader)
    # a second recorder starts from an empty list
    assert plan_cache.recorder_for(OracleDeviceState).ops == []


def test_differential_fuzz_with_every_circuit_cached(cirq, SV, monkeypatch):
    """The simulator fuzz tests (random circuits, sizes, fusion widths, split settings,
    channels, resets, mid-circuit measurements; final states, amplitude gathers and
    seeded records against cirq.Simulator) with every circuit served from the schedule
    cache from its first call on."""
    import test_simulators as ts
    from cirq_b200 import plan_cache

    monkeypatch.setenv('CIRQ_B200_PLAN_CACHE_EAGER', '1')
    ts.test_differential_fuzz_state_vector(cirq, SV, seed=7)
    assert plan_cache.CACHE.builds >= 60 and plan_cache.CACHE.hits >= 60
    builds = plan_cache.CACHE.builds
    ts.test_differential_fuzz_noisy_state_vector_seeded(cirq, SV)
    assert plan_cache.CACHE.builds > builds  # (the noise-free cases with a unitary prefix)
