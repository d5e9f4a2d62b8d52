"""Drop-in parity of B200Simulator / B200DensityMatrixSimulator with the
reference cirq.Simulator / cirq.DensityMatrixSimulator on identical circuits.

Every test runs twice:
  * backend 'oracle' (CPU, no GPU needed): the simulators drive a test-only
    DeviceState backed by the numpy oracle, which checks the host logic;
  * backend 'cuda' (marked gpu): the real CUDA path through the C-ABI.

Modelled on the reference's cirq-core/cirq/sim/sparse_simulator_test.py and
density_matrix_simulator_test.py.  Tolerances: 1e-5 (complex64) / 1e-12
(complex128) max-abs on amplitudes (north star)."""
import numpy as np
import pytest
import sympy

from fake_device import OracleDeviceState

ATOL = {np.complex64: 1e-5, np.complex128: 1e-12}


@pytest.fixture(
    params=['oracle', pytest.param('cuda', marks=pytest.mark.gpu)], scope='module'
)
def backend(request, cirq):
    import cirq_b200.dm_simulator as dmm
    import cirq_b200.sv_simulator as svm
    from cirq_b200.device_state import DeviceState

    if request.param == 'oracle':
        svm.DeviceState = OracleDeviceState
        dmm.DeviceState = OracleDeviceState
    else:
        svm.DeviceState = DeviceState
        dmm.DeviceState = DeviceState
    yield request.param
    svm.DeviceState = DeviceState
    dmm.DeviceState = DeviceState


@pytest.fixture(scope='module')
def SV(backend):
    import cirq_b200

    return cirq_b200.B200Simulator


@pytest.fixture(params=[False, True], ids=['dense', 'split'])
def split(request):
    """split_untangled_states, applied to BOTH simulators in seeded comparisons:
    the random stream is consumed per unentangled factor, so seeded results
    agree between simulators configured alike."""
    return request.param


@pytest.fixture(scope='module')
def DM(backend):
    import cirq_b200

    return cirq_b200.B200DensityMatrixSimulator


# --------------------------------------------------------------------------- state vector


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
@pytest.mark.parametrize('max_fused', [1, 2, 3, 4, 5])
def test_random_circuits_match_reference(cirq, SV, dtype, max_fused):
    if dtype == np.complex128 and max_fused == 5:
        max_fused = 4
    for n, depth, seed in ((1, 5, 0), (2, 8, 1), (5, 10, 2), (9, 12, 3), (12, 16, 4)):
        qubits = cirq.LineQubit.range(n)
        circuit = cirq.testing.random_circuit(qubits, depth, 0.9, random_state=seed)
        want = cirq.Simulator(dtype=dtype).simulate(circuit, qubit_order=qubits)
        got = SV(dtype=dtype, max_fused_qubits=max_fused).simulate(circuit, qubit_order=qubits)
        np.testing.assert_allclose(
            got.final_state_vector, want.final_state_vector, atol=ATOL[dtype], rtol=0
        )
        assert got.qubit_map == want.qubit_map
        assert got.final_state_vector.dtype == dtype


def test_run_reports_config1_circuit_like_reference(cirq, SV):
    """BASELINE config 1 generator at a CPU-friendly size."""
    qubits = cirq.LineQubit.range(14)
    circuit = cirq.testing.random_circuit(qubits, 20, 0.9, random_state=1234)
    want = cirq.Simulator(dtype=np.complex64).simulate(circuit, qubit_order=qubits)
    got = SV(dtype=np.complex64).simulate(circuit, qubit_order=qubits)
    np.testing.assert_allclose(got.final_state_vector, want.final_state_vector, atol=1e-5, rtol=0)


def test_qubit_order_and_initial_states(cirq, SV):
    a, b, c = cirq.LineQubit.range(3)
    circuit = cirq.Circuit(cirq.X(a), cirq.H(b), cirq.CNOT(b, c), cirq.T(c))
    for order in ([a, b, c], [c, a, b], [b, c, a]):
        want = cirq.Simulator().simulate(circuit, qubit_order=order)
        got = SV().simulate(circuit, qubit_order=order)
        np.testing.assert_allclose(got.final_state_vector, want.final_state_vector, atol=1e-6)
    for init in (5, 0, 7):
        want = cirq.Simulator().simulate(circuit, initial_state=init)
        got = SV().simulate(circuit, initial_state=init)
        np.testing.assert_allclose(got.final_state_vector, want.final_state_vector, atol=1e-6)
    vec = cirq.testing.random_superposition(8, random_state=3).astype(np.complex64)
    keep = vec.copy()
    want = cirq.Simulator().simulate(circuit, initial_state=vec)
    got = SV().simulate(circuit, initial_state=vec)
    np.testing.assert_allclose(got.final_state_vector, want.final_state_vector, atol=1e-6)
    np.testing.assert_array_equal(vec, keep)  # test_does_not_modify_initial_state
    got = SV().simulate(circuit, initial_state=vec.reshape(2, 2, 2))
    np.testing.assert_allclose(got.final_state_vector, want.final_state_vector, atol=1e-6)
    want = cirq.Simulator().simulate(circuit, initial_state=(1, 0, 1))
    got = SV().simulate(circuit, initial_state=(1, 0, 1))
    np.testing.assert_allclose(got.final_state_vector, want.final_state_vector, atol=1e-6)
    with pytest.raises(ValueError):
        SV(split_untangled_states=False).simulate(circuit, initial_state=8)


def test_run_terminal_measurements_seeded_like_reference(cirq, SV, split):
    qubits = cirq.LineQubit.range(6)
    circuit = cirq.testing.random_circuit(qubits, 8, 0.9, random_state=7)
    circuit.append(cirq.measure(*qubits, key='m'))
    want = cirq.Simulator(seed=11, split_untangled_states=split).run(circuit, repetitions=400)
    got = SV(seed=11, split_untangled_states=split).run(circuit, repetitions=400)
    w, g = want.measurements['m'], got.measurements['m']
    assert g.shape == w.shape and g.dtype == w.dtype
    assert np.mean(np.any(w != g, axis=1)) <= 0.01
    # subset, permuted order, invert mask, two keys
    circuit2 = cirq.testing.random_circuit(qubits, 8, 0.9, random_state=8)
    circuit2.append(
        [
            cirq.measure(qubits[4], qubits[1], key='a', invert_mask=(True, False)),
            cirq.measure(qubits[0], key='b'),
        ]
    )
    want = cirq.Simulator(seed=3, split_untangled_states=split).run(circuit2, repetitions=300)
    got = SV(seed=3, split_untangled_states=split).run(circuit2, repetitions=300)
    for key in ('a', 'b'):
        assert got.measurements[key].shape == want.measurements[key].shape
        assert np.mean(np.any(want.measurements[key] != got.measurements[key], axis=1)) <= 0.01


def test_run_histogram_chi_squared(cirq, SV):
    qubits = cirq.LineQubit.range(5)
    circuit = cirq.testing.random_circuit(qubits, 10, 0.9, random_state=21)
    probs = np.abs(cirq.Simulator(dtype=np.complex128).simulate(circuit, qubit_order=qubits).final_state_vector) ** 2
    circuit.append(cirq.measure(*qubits, key='m'))
    reps = 20000
    res = SV(seed=1).run(circuit, repetitions=reps)
    ints = res.measurements['m'].astype(np.int64) @ (1 << np.arange(4, -1, -1))
    hist = np.bincount(ints, minlength=32)
    mask = probs * reps > 5
    chi2 = np.sum((hist[mask] - probs[mask] * reps) ** 2 / (probs[mask] * reps))
    dof = mask.sum() - 1
    assert chi2 < dof + 6 * np.sqrt(2 * dof) + 10
    assert hist[~mask].sum() <= 6 * max(1.0, (probs[~mask] * reps).sum()) + 5


def test_mid_circuit_measurement_classical_control_reset(cirq, SV, split):
    a, b = cirq.LineQubit.range(2)
    circuit = cirq.Circuit(
        cirq.H(a),
        cirq.measure(a, key='x'),
        cirq.X(b).with_classical_controls('x'),
        cirq.measure(b, key='y'),
        cirq.reset(a),
        cirq.measure(a, key='z'),
    )
    want = cirq.Simulator(seed=5, split_untangled_states=split).run(circuit, repetitions=60)
    got = SV(seed=5, split_untangled_states=split).run(circuit, repetitions=60)
    for k in ('x', 'y', 'z'):
        np.testing.assert_array_equal(got.measurements[k], want.measurements[k])
        assert got.measurements[k].dtype == want.measurements[k].dtype
    np.testing.assert_array_equal(got.measurements['x'], got.measurements['y'])
    assert not got.measurements['z'].any()
    # phase preserved through measure + control + reset (sparse_simulator_test.py:1453-1465)
    q = cirq.LineQubit.range(3)
    c = cirq.Circuit(cirq.Z(q[0]) ** 0.3, cirq.measure(q[1], key='m'), cirq.reset(q[2]))
    want = cirq.Simulator().simulate(c, initial_state=(1, 1, 1))
    got = SV().simulate(c, initial_state=(1, 1, 1))
    np.testing.assert_allclose(got.final_state_vector, want.final_state_vector, atol=1e-6)


def test_simulate_moment_steps_matches_reference(cirq, SV, split):
    qubits = cirq.LineQubit.range(4)
    circuit = cirq.testing.random_circuit(qubits, 6, 0.9, random_state=31)
    circuit.append(cirq.measure(qubits[0], qubits[2], key='m'))
    ref_steps = cirq.Simulator(seed=2, split_untangled_states=split).simulate_moment_steps(circuit, qubit_order=qubits)
    steps = SV(seed=2, split_untangled_states=split).simulate_moment_steps(
        circuit, qubit_order=qubits
    )
    count = 0
    for i, (step, ref) in enumerate(zip(steps, ref_steps)):
        np.testing.assert_allclose(step.state_vector(), ref.state_vector(), atol=1e-6)
        if i == 2:
            s1 = step.sample(qubits[:2], repetitions=5, seed=3)
            s2 = ref.sample(qubits[:2], repetitions=5, seed=3)
            np.testing.assert_array_equal(s1, s2)
        count += 1
    assert count == len(circuit)
    assert dict(step.measurements) == dict(ref.measurements)
    assert step.dirac_notation() == ref.dirac_notation()


def test_param_sweeps(cirq, SV, split):
    q = cirq.LineQubit.range(3)
    t, s = sympy.Symbol('t'), sympy.Symbol('s')
    circuit = cirq.Circuit(
        cirq.H.on_each(*q), cirq.CZ(q[0], q[1]) ** t, cirq.rx(s).on(q[2]), cirq.CNOT(q[2], q[0])
    )
    sweep = cirq.Product(cirq.Linspace('t', 0, 1, 3), cirq.Points('s', [0.1, 0.7]))
    want = cirq.Simulator().simulate_sweep(circuit, sweep)
    got = SV().simulate_sweep(circuit, sweep)
    assert len(got) == len(want) == 6
    for g, w in zip(got, want):
        assert g.params == w.params
        np.testing.assert_allclose(g.final_state_vector, w.final_state_vector, atol=1e-6)
    circuit.append(cirq.measure(*q, key='m'))
    want = cirq.Simulator(seed=9, split_untangled_states=split).run_sweep(
        circuit, sweep, repetitions=50
    )
    got = SV(seed=9, split_untangled_states=split).run_sweep(circuit, sweep, repetitions=50)
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g.measurements['m'], w.measurements['m'])
    with pytest.raises(ValueError, match='symbols'):
        SV().simulate(circuit)


def test_noisy_state_vector_trajectories_seeded(cirq, SV, split):
    q = cirq.LineQubit.range(3)
    circuit = cirq.Circuit(
        cirq.H(q[0]),
        cirq.CNOT(q[0], q[1]),
        cirq.bit_flip(0.3).on(q[1]),
        cirq.amplitude_damp(0.4).on(q[0]),
        cirq.depolarize(0.2).on(q[2]),
        cirq.measure(*q, key='m'),
    )
    want = cirq.Simulator(seed=17, split_untangled_states=split).run(circuit, repetitions=80)
    got = SV(seed=17, split_untangled_states=split).run(circuit, repetitions=80)
    np.testing.assert_array_equal(got.measurements['m'], want.measurements['m'])
    want = cirq.Simulator(seed=4, noise=cirq.depolarize(0.1), split_untangled_states=split).run(circuit, repetitions=40)
    got = SV(seed=4, noise=cirq.depolarize(0.1), split_untangled_states=split).run(
        circuit, repetitions=40
    )
    np.testing.assert_array_equal(got.measurements['m'], want.measurements['m'])


def _chi2_ok(counts, probs, reps):
    """Pearson chi-squared of a histogram against exact probabilities, bins of
    expected count < 5 pooled; passes below the 99.99 % quantile."""
    from scipy import stats

    expected = np.asarray(probs, dtype=np.float64) * reps
    small = expected < 5
    obs = np.append(counts[~small], counts[small].sum())
    exp = np.append(expected[~small], expected[small].sum())
    keep = exp > 0
    assert obs[~keep].sum() == 0, 'samples landed on zero-probability outcomes'
    obs, exp = obs[keep], exp[keep]
    if len(obs) < 2:
        return True
    chi2 = float(((obs - exp) ** 2 / exp).sum())
    return chi2 < stats.chi2.ppf(0.9999, len(obs) - 1)


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_batched_trajectories_distribution(cirq, SV, dtype):
    """trajectory_batch: same distribution as the reference's per-repetition loop
    (sim/simulator_base.py:249-264); checked against the exact outcome
    probabilities from cirq.DensityMatrixSimulator."""
    q = cirq.LineQubit.range(4)
    circuit = cirq.Circuit(
        cirq.H(q[0]),
        cirq.CNOT(q[0], q[1]),
        cirq.ry(0.7).on(q[2]),
        cirq.bit_flip(0.3).on(q[1]),
        cirq.amplitude_damp(0.4).on(q[0]),
        cirq.CZ(q[1], q[2]) ** 0.5,
        cirq.depolarize(0.2).on(q[2]),
        cirq.rx(0.4).on(q[3]),
        cirq.ISWAP(q[2], q[3]) ** 0.5,
        cirq.asymmetric_depolarize(0.05, 0.1, 0.15).on(q[3]),
        cirq.measure(q[2], q[0], q[3], q[1], key='m'),
    )
    reps = 6000
    rho = cirq.DensityMatrixSimulator(dtype=np.complex128).simulate(circuit[:-1]).final_density_matrix
    p = np.real(np.diag(rho))  # index = q0 q1 q2 q3 big-endian
    for batch in (64, 4096):
        sim = SV(seed=11, dtype=dtype, trajectory_batch=batch)
        res = sim.run(circuit, repetitions=reps)
        assert sim.last_run_info['path'] == 'batched trajectories'
        m = res.measurements['m']
        assert m.shape == (reps, 4) and m.dtype == np.uint8
        idx = (m[:, 1].astype(int) << 3) | (m[:, 3] << 2) | (m[:, 0] << 1) | m[:, 2]
        counts = np.bincount(idx, minlength=16)
        assert _chi2_ok(counts, p, reps)
    # noise model on the simulator + invert mask
    circuit2 = cirq.Circuit(
        cirq.X(q[0]), cirq.H(q[1]), cirq.CNOT(q[1], q[2]),
        cirq.measure(q[0], q[1], q[2], key='z', invert_mask=(True, False, True)),
    )
    noise = cirq.depolarize(0.1)
    noisy = cirq.Circuit(cirq.ConstantQubitNoiseModel(noise).noisy_moments(circuit2[:-1], q[:3]))
    rho = cirq.DensityMatrixSimulator(dtype=np.complex128).simulate(noisy, qubit_order=q[:3]).final_density_matrix
    # the measurement moment gets noise after it, which does not change the record
    p = np.real(np.diag(rho))
    sim = SV(seed=3, dtype=dtype, noise=noise, trajectory_batch=1024)
    m = sim.run(circuit2, repetitions=4000).measurements['z']
    assert sim.last_run_info['path'] == 'batched trajectories'
    idx = ((m[:, 0] ^ 1).astype(int) << 2) | (m[:, 1] << 1) | (m[:, 2] ^ 1)
    assert _chi2_ok(np.bincount(idx, minlength=8), p, 4000)


def test_batched_trajectories_mid_circuit_measurement(cirq, SV):
    """Mid-circuit measurements collapse per trajectory: correlations between a
    mid-circuit record, a later record of the same qubit and a terminal record."""
    q = cirq.LineQubit.range(3)
    circuit = cirq.Circuit(
        cirq.H(q[0]),
        cirq.CNOT(q[0], q[1]),
        cirq.measure(q[0], key='a'),
        cirq.H(q[2]),
        cirq.CNOT(q[1], q[2]),
        cirq.measure(q[1], key='b'),
        cirq.reset(q[1]),
        cirq.X(q[0]),
        cirq.measure(q[0], q[1], key='c'),
        cirq.measure(q[2], key='d'),
        cirq.measure(q[2], key='d'),
    )
    reps = 3000
    sim = SV(seed=5, trajectory_batch=512)
    res = sim.run(circuit, repetitions=reps)
    assert sim.last_run_info['path'] == 'batched trajectories'
    a, b, c = res.records['a'][:, 0], res.records['b'][:, 0], res.records['c'][:, 0]
    d = res.records['d']
    assert a.shape == (reps, 1) and c.shape == (reps, 2) and d.shape == (reps, 2, 1)
    np.testing.assert_array_equal(a, b)  # Bell pair
    np.testing.assert_array_equal(c[:, 0], 1 - a[:, 0])  # X after the collapse
    assert not c[:, 1].any()  # reset
    np.testing.assert_array_equal(d[:, 0], d[:, 1])
    # q2 = H-random XOR q1: uniform and independent of a
    joint = np.bincount(2 * a[:, 0].astype(int) + d[:, 0, 0], minlength=4)
    assert _chi2_ok(joint, np.full(4, 0.25), reps)
    # same seed, same records
    again = SV(seed=5, trajectory_batch=512).run(circuit, repetitions=reps)
    np.testing.assert_array_equal(again.records['a'][:, 0], a)
    np.testing.assert_array_equal(again.records['d'], d)


def test_batched_trajectories_fall_back_to_reference_loop(cirq, SV):
    """Operations that cannot be batched (classical control here) take the
    reference's per-repetition loop, with its seeded results."""
    q = cirq.LineQubit.range(2)
    circuit = cirq.Circuit(
        cirq.H(q[0]),
        cirq.measure(q[0], key='a'),
        cirq.X(q[1]).with_classical_controls('a'),
        cirq.measure(q[1], key='b'),
    )
    want = cirq.Simulator(seed=8).run(circuit, repetitions=30)
    sim = SV(seed=8, trajectory_batch=256)
    got = sim.run(circuit, repetitions=30)
    assert sim.last_run_info['path'] == 'reference loop'
    np.testing.assert_array_equal(got.measurements['a'], want.measurements['a'])
    np.testing.assert_array_equal(got.measurements['b'], want.measurements['b'])
    # terminal-measurement-only circuits keep the sample-once path
    c2 = cirq.Circuit(cirq.H(q[0]), cirq.CNOT(q[0], q[1]), cirq.measure(*q, key='m'))
    want = cirq.Simulator(seed=2).run(c2, repetitions=50)
    got = SV(seed=2, trajectory_batch=256).run(c2, repetitions=50)
    np.testing.assert_array_equal(got.measurements['m'], want.measurements['m'])


def test_expectation_values_and_amplitudes(cirq, SV):
    q = cirq.LineQubit.range(4)
    circuit = cirq.testing.random_circuit(q, 8, 0.9, random_state=5)
    circuit.append(cirq.I.on_each(*q))
    obs = [
        cirq.Z(q[0]) * cirq.X(q[2]),
        0.5 * cirq.Y(q[1]) + 2.0 * cirq.Z(q[3]) * cirq.Y(q[0]) - 1.5 * cirq.X(q[1]) * cirq.X(q[2]),
        cirq.PauliString(),
    ]
    want = cirq.Simulator(dtype=np.complex128).simulate_expectation_values(circuit, obs)
    got = SV(dtype=np.complex128).simulate_expectation_values(circuit, obs)
    np.testing.assert_allclose(got, want, atol=1e-10)
    bitstrings = [0, 3, 15, 9]
    want = cirq.Simulator().compute_amplitudes(circuit, bitstrings)
    got = SV().compute_amplitudes(circuit, bitstrings)
    np.testing.assert_allclose(got, want, atol=1e-6)
    with pytest.raises(ValueError, match='terminal measurements'):
        SV().simulate_expectation_values(
            cirq.Circuit(cirq.H(q[0]), cirq.measure(q[0])), [cirq.Z(q[0])]
        )


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_density_matrix_of_and_bloch_vector_on_device(cirq, SV, dtype):
    """sim/state_vector.py:109-167 through the device reduction kernel."""
    q = cirq.LineQubit.range(7)
    circuit = cirq.testing.random_circuit(q, 10, 0.9, random_state=21)
    circuit.append(cirq.I.on_each(*q))
    want = cirq.Simulator(dtype=dtype).simulate(circuit, qubit_order=q)
    got = SV(dtype=dtype).simulate(circuit, qubit_order=q)
    atol = 1e-6 if dtype == np.complex64 else 1e-12
    for keep in ([q[3]], [q[0], q[6]], [q[5], q[1], q[2]], [q[6], q[0], q[3], q[4]],
                 [q[2], q[4], q[1], q[6], q[0]], [q[1], q[2], q[3], q[4], q[5], q[6]]):
        np.testing.assert_allclose(got.density_matrix_of(keep), want.density_matrix_of(keep), atol=atol)
    for x in q:
        np.testing.assert_allclose(got.bloch_vector_of(x), want.bloch_vector_of(x), atol=atol * 4)
    np.testing.assert_allclose(got.density_matrix_of(), want.density_matrix_of(), atol=atol)
    with pytest.raises(KeyError):
        got.density_matrix_of([cirq.LineQubit(99)])
    steps_w = list(cirq.Simulator(dtype=dtype).simulate_moment_steps(circuit, qubit_order=q))
    steps_g = list(SV(dtype=dtype).simulate_moment_steps(circuit, qubit_order=q))
    for i in (0, 4, len(steps_w) - 1):
        np.testing.assert_allclose(
            steps_g[i].density_matrix_of([q[2], q[5]]), steps_w[i].density_matrix_of([q[2], q[5]]), atol=atol
        )
        np.testing.assert_allclose(steps_g[i].bloch_vector_of(q[4]), steps_w[i].bloch_vector_of(q[4]), atol=atol * 4)


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_relabelled_swaps_and_bit_addressed_readers(cirq, SV, dtype):
    """A dense state (split_untangled_states=False) with SWAP-heavy circuits: the
    scheduler relabels the SWAPs; amplitudes, expectation values and reduced
    density matrices read through the bit map, the raw state restores the order."""
    from cirq_b200 import workloads as W

    atol = 2e-6 if dtype == np.complex64 else 1e-12
    qft, q = W.qft_circuit(9)
    rng_circuit = cirq.testing.random_circuit(q, 12, 0.9, random_state=11)
    for circuit in (rng_circuit + qft, qft + rng_circuit + cirq.Circuit(cirq.SWAP(q[0], q[5]), cirq.SWAP(q[5], q[8]))):
        ref = cirq.Simulator(dtype=dtype, split_untangled_states=False)
        sim = SV(dtype=dtype, split_untangled_states=False)
        idx = [0, 5, 77, 300, 511]
        np.testing.assert_allclose(
            sim.compute_amplitudes(circuit, idx, qubit_order=q), ref.compute_amplitudes(circuit, idx, qubit_order=q),
            atol=atol)
        obs = [cirq.Z(q[0]) * cirq.X(q[7]), cirq.Y(q[3]) + 0.5 * cirq.Z(q[8]) * cirq.Z(q[1])]
        np.testing.assert_allclose(
            sim.simulate_expectation_values(circuit, obs, qubit_order=q),
            ref.simulate_expectation_values(circuit, obs, qubit_order=q), atol=atol * 4)
        want = ref.simulate(circuit, qubit_order=q)
        got = sim.simulate(circuit, qubit_order=q)
        state = got._get_merged_sim_state()._state
        np.testing.assert_allclose(got.density_matrix_of([q[6], q[0]]), want.density_matrix_of([q[6], q[0]]), atol=atol)
        np.testing.assert_allclose(got.bloch_vector_of(q[8]), want.bloch_vector_of(q[8]), atol=atol * 4)
        assert state._where is not None  # the readers above left the SWAPs relabelled
        # ... and the raw state comes back in canonical order
        np.testing.assert_allclose(got.final_state_vector, want.final_state_vector, atol=atol)
        assert state._where is None
        # sampling after a mapped read: canonical order again, seeded like the reference
        c2 = circuit + cirq.Circuit(cirq.measure(*q, key='m'))
        a = SV(dtype=dtype, seed=3, split_untangled_states=False).run(c2, repetitions=40)
        b = cirq.Simulator(dtype=dtype, seed=3, split_untangled_states=False).run(c2, repetitions=40)
        np.testing.assert_array_equal(a.measurements['m'], b.measurements['m'])


def test_wide_and_composite_operations(cirq, SV):
    q = cirq.LineQubit.range(7)
    circuit = cirq.Circuit(
        cirq.H.on_each(*q[:3]),
        cirq.qft(*q, without_reverse=True),
        cirq.MatrixGate(cirq.testing.random_unitary(8, random_state=1)).on(q[5], q[0], q[3]),
        cirq.CCX(q[1], q[2], q[6]),
        cirq.ControlledGate(cirq.ISWAP, num_controls=2).on(q[0], q[4], q[2], q[6]),
        cirq.MatrixGate(cirq.testing.random_unitary(64, random_state=2)).on(*q[:6]),
        cirq.CircuitOperation(cirq.FrozenCircuit(cirq.X(q[0]), cirq.CZ(q[0], q[1])), repetitions=3),
    )
    want = cirq.Simulator(dtype=np.complex128).simulate(circuit, qubit_order=q)
    got = SV(dtype=np.complex128).simulate(circuit, qubit_order=q)
    np.testing.assert_allclose(got.final_state_vector, want.final_state_vector, atol=1e-10)


def test_errors_match_reference(cirq, SV):
    q = cirq.LineQubit.range(2)
    with pytest.raises(ValueError, match='no measurements'):
        SV().run(cirq.Circuit(cirq.H(q[0])))
    with pytest.raises(ValueError, match='complex'):
        SV(dtype=np.float32)

    class Unsupported(cirq.Gate):
        def _num_qubits_(self):
            return 1

    with pytest.raises(TypeError, match="doesn't support"):
        SV().simulate(cirq.Circuit(Unsupported().on(q[0])))
    with pytest.raises(ValueError, match='dimension 2|qubits only'):
        SV().simulate(cirq.Circuit(cirq.IdentityGate(qid_shape=(3,)).on(cirq.LineQid(0, 3))))


def test_repeated_keys_and_empty_circuit(cirq, SV, split):
    q = cirq.LineQubit.range(2)
    circuit = cirq.Circuit(cirq.X(q[0]), cirq.measure(q[0], key='k'), cirq.measure(q[1], key='k'))
    want = cirq.Simulator(seed=1, split_untangled_states=split).run(circuit, repetitions=7)
    got = SV(seed=1, split_untangled_states=split).run(circuit, repetitions=7)
    np.testing.assert_array_equal(got.records['k'], want.records['k'])
    assert got.records['k'].shape == (7, 2, 1)
    res = SV().simulate(cirq.Circuit(), qubit_order=q)
    np.testing.assert_allclose(res.final_state_vector, [1, 0, 0, 0])
    assert str(res).startswith('measurements:')


def test_same_seed_same_result_and_global_state_untouched(cirq, SV):
    q = cirq.LineQubit.range(3)
    circuit = cirq.Circuit(cirq.H.on_each(*q), cirq.measure(*q, key='m'))
    a = SV(seed=123).run(circuit, repetitions=30).measurements['m']
    b = SV(seed=123).run(circuit, repetitions=30).measurements['m']
    np.testing.assert_array_equal(a, b)
    np.random.seed(10)
    before = np.random.get_state()[1].copy()
    SV(seed=1).run(circuit, repetitions=5)
    np.testing.assert_array_equal(before, np.random.get_state()[1])


# --------------------------------------------------------------------------- density matrix


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_density_matrix_matches_reference(cirq, DM, dtype):
    for n, depth, seed in ((1, 4, 0), (3, 6, 1), (5, 8, 2)):
        qubits = cirq.LineQubit.range(n)
        circuit = cirq.testing.random_circuit(qubits, depth, 0.9, random_state=seed)
        circuit.append(cirq.I.on_each(*qubits))
        circuit.append(cirq.amplitude_damp(0.2).on(qubits[0]))
        circuit.append(cirq.phase_damp(0.3).on(qubits[-1]))
        circuit.append(cirq.bit_flip(0.1).on(qubits[0]))
        if n >= 2:
            circuit.append(cirq.depolarize(0.1, n_qubits=2).on(qubits[0], qubits[1]))
        for noise in (None, cirq.depolarize(0.02)):
            want = cirq.DensityMatrixSimulator(dtype=dtype, noise=noise).simulate(
                circuit, qubit_order=qubits
            )
            for max_fused in (2, 4):
                got = DM(dtype=dtype, noise=noise, max_fused_qubits=max_fused).simulate(
                    circuit, qubit_order=qubits
                )
                np.testing.assert_allclose(
                    got.final_density_matrix, want.final_density_matrix, atol=ATOL[dtype], rtol=0
                )


def test_density_matrix_equals_state_vector_outer_product(cirq, DM):
    q = cirq.LineQubit.range(4)
    circuit = cirq.testing.random_circuit(q, 8, 0.9, random_state=12)
    circuit.append(cirq.I.on_each(*q))
    psi = cirq.Simulator(dtype=np.complex128).simulate(circuit, qubit_order=q).final_state_vector
    rho = DM(dtype=np.complex128).simulate(circuit, qubit_order=q).final_density_matrix
    np.testing.assert_allclose(rho, np.outer(psi, psi.conj()), atol=1e-10)


def test_density_matrix_run_and_measurement_seeded(cirq, DM, split):
    q = cirq.LineQubit.range(3)
    circuit = cirq.Circuit(
        cirq.H(q[0]),
        cirq.CNOT(q[0], q[1]),
        cirq.depolarize(0.2).on(q[1]),
        cirq.ry(0.7).on(q[2]),
        cirq.measure(q[2], q[0], key='a'),
        cirq.measure(q[1], key='b', invert_mask=(True,)),
    )
    want = cirq.DensityMatrixSimulator(
        seed=6, noise=cirq.depolarize(0.05), split_untangled_states=split
    ).run(circuit, repetitions=200)
    got = DM(seed=6, noise=cirq.depolarize(0.05), split_untangled_states=split).run(
        circuit, repetitions=200
    )
    for k in ('a', 'b'):
        assert got.measurements[k].dtype == want.measurements[k].dtype
        assert np.mean(np.any(got.measurements[k] != want.measurements[k], axis=1)) <= 0.01
    # mid-circuit measurement collapses rho (simulate path)
    c2 = cirq.Circuit(
        cirq.H(q[0]), cirq.CNOT(q[0], q[1]), cirq.measure(q[0], key='m'), cirq.H(q[2]),
        cirq.X(q[2]).with_classical_controls('m'),
    )
    want = cirq.DensityMatrixSimulator(seed=8, split_untangled_states=split).simulate(
        c2, qubit_order=q
    )
    got = DM(seed=8, split_untangled_states=split).simulate(c2, qubit_order=q)
    assert dict(got.measurements) == dict(want.measurements) or np.array_equal(
        got.measurements['m'], want.measurements['m']
    )
    np.testing.assert_allclose(got.final_density_matrix, want.final_density_matrix, atol=1e-6)


def test_density_matrix_sweep_qaoa_style(cirq, DM, split):
    """BASELINE config 5 shape at a CPU-friendly size: noisy QAOA, run_sweep."""
    q = cirq.LineQubit.range(4)
    beta, gamma = sympy.Symbol('beta0'), sympy.Symbol('gamma0')
    edges = [(0, 1), (1, 2), (2, 3), (0, 3)]
    circuit = cirq.Circuit(cirq.H.on_each(*q))
    circuit.append(cirq.ZZ(q[i], q[j]) ** gamma for i, j in edges)
    circuit.append(cirq.X.on_each(*q) ** beta if False else [cirq.X(x) ** beta for x in q])
    circuit.append(cirq.measure(*q, key='m'))
    sweep = cirq.Zip(cirq.Linspace('beta0', 0.1, 0.9, 4), cirq.Linspace('gamma0', 0.2, 0.8, 4))
    want = cirq.DensityMatrixSimulator(
        noise=cirq.depolarize(0.01), seed=0, split_untangled_states=split
    ).run_sweep(
        circuit, sweep, repetitions=100
    )
    got = DM(noise=cirq.depolarize(0.01), seed=0, split_untangled_states=split).run_sweep(
        circuit, sweep, repetitions=100
    )
    assert len(got) == 4
    for g, w in zip(got, want):
        assert g.params == w.params
        assert g.measurements['m'].shape == (100, 4)
        assert np.mean(np.any(g.measurements['m'] != w.measurements['m'], axis=1)) <= 0.02


def test_density_matrix_expectation_steps_and_initial_state(cirq, DM):
    q = cirq.LineQubit.range(3)
    circuit = cirq.Circuit(cirq.H(q[0]), cirq.CNOT(q[0], q[1]), cirq.rx(0.4).on(q[2]))
    obs = [cirq.Z(q[0]) * cirq.Z(q[1]), cirq.X(q[2]) + 0.5 * cirq.Y(q[2])]
    noise = cirq.depolarize(0.03)
    want = cirq.DensityMatrixSimulator(noise=noise).simulate_expectation_values(circuit, obs)
    got = DM(noise=noise).simulate_expectation_values(circuit, obs)
    np.testing.assert_allclose(got, want, atol=1e-6)
    ref_steps = cirq.DensityMatrixSimulator(noise=noise).simulate_moment_steps(circuit)
    for step, ref in zip(DM(noise=noise).simulate_moment_steps(circuit), ref_steps):
        np.testing.assert_allclose(step.density_matrix(), ref.density_matrix(), atol=1e-6)
    rho0 = cirq.testing.random_density_matrix(8, random_state=2).astype(np.complex64)
    want = cirq.DensityMatrixSimulator().simulate(circuit, initial_state=rho0)
    got = DM().simulate(circuit, initial_state=rho0)
    np.testing.assert_allclose(got.final_density_matrix, want.final_density_matrix, atol=1e-6)
    want = cirq.DensityMatrixSimulator().simulate(circuit, initial_state=5)
    got = DM().simulate(circuit, initial_state=5)
    np.testing.assert_allclose(got.final_density_matrix, want.final_density_matrix, atol=1e-6)


# --------------------------------------------------------------------------- mux entry points


def _qaoa_like(cirq, n, seed):
    import networkx as nx

    q = cirq.LineQubit.range(n)
    graph = nx.random_regular_graph(3, n, seed=seed)
    g0, g1, b0, b1 = sympy.symbols('g0 g1 b0 b1')
    c = cirq.Circuit(cirq.H.on_each(*q))
    for g, b in ((g0, b0), (g1, b1)):
        c.append(cirq.ZZ(q[i], q[j]) ** g for i, j in graph.edges)
        c.append(cirq.rx(2 * b).on_each(*q))
    c.append(cirq.measure(*q, key='m'))
    sweep = cirq.Zip(
        cirq.Linspace('g0', 0.1, 0.9, 7), cirq.Linspace('g1', 0.8, 0.2, 7),
        cirq.Linspace('b0', 0.3, 1.2, 7), cirq.Points('b1', [0.5] * 7),
    )
    return c, q, sweep


def test_batched_sweep_state_vector_matches_sequential(cirq, SV):
    """sweep_batch=True: every resolver's samples equal the reference's
    resolver-by-resolver run_sweep (sim/simulator.py:62-94) under the same seed."""
    c, q, sweep = _qaoa_like(cirq, 6, 3)
    c.insert(1, cirq.FSimGate(sympy.Symbol('g0'), 0.3).on(q[0], q[3]))  # non-EigenGate symbol
    c.insert(2, cirq.CZ(q[1], q[2]))
    want = cirq.Simulator(seed=7, split_untangled_states=False).run_sweep(c, sweep, repetitions=40)
    sim = SV(seed=7, sweep_batch=True)
    got = sim.run_sweep(c, sweep, repetitions=40)
    assert sim.last_run_info['path'] == 'batched sweep' and sim.last_run_info['resolvers'] == 7
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.params == w.params
        np.testing.assert_array_equal(g.measurements['m'], w.measurements['m'])
    # complex128, invert mask + two keys
    c2 = c[:-1] + cirq.Circuit(
        cirq.measure(q[0], q[2], key='a', invert_mask=(True, False)), cirq.measure(q[4], key='b'))
    want = cirq.Simulator(seed=1, dtype=np.complex128, split_untangled_states=False).run_sweep(c2, sweep, repetitions=25)
    got = SV(seed=1, dtype=np.complex128, sweep_batch=True).run_sweep(c2, sweep, repetitions=25)
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g.measurements['a'], w.measurements['a'])
        np.testing.assert_array_equal(g.measurements['b'], w.measurements['b'])


def test_batched_sweep_falls_back(cirq, SV):
    """Shapes the batch cannot take (mid-circuit measurement, noise on a state
    vector) use the reference loop and still match it."""
    q = cirq.LineQubit.range(3)
    t = sympy.Symbol('t')
    c = cirq.Circuit(cirq.rx(t).on(q[0]), cirq.measure(q[0], key='a'), cirq.CNOT(q[0], q[1]),
                     cirq.measure(q[1], q[2], key='b'))
    sweep = cirq.Linspace('t', 0, 2, 4)
    want = cirq.Simulator(seed=3).run_sweep(c, sweep, repetitions=20)
    sim = SV(seed=3, sweep_batch=True)
    got = sim.run_sweep(c, sweep, repetitions=20)
    assert sim.last_run_info.get('path') != 'batched sweep'
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g.measurements['b'], w.measurements['b'])
    with pytest.raises(ValueError, match='no measurements'):
        SV(sweep_batch=True).run_sweep(cirq.Circuit(cirq.X(q[0])), sweep)
    with pytest.raises(ValueError, match='symbols|resolved'):
        SV(sweep_batch=True).run_sweep(c, cirq.Linspace('other', 0, 1, 3), repetitions=2)


@pytest.mark.parametrize('noise_p', [0.0, 0.02])
def test_batched_sweep_density_matrix_matches_sequential(cirq, DM, noise_p):
    """Config-5 shape (noisy QAOA, run_sweep) on a small graph: batched over the
    resolvers = the reference resolver by resolver, seeded."""
    c, q, sweep = _qaoa_like(cirq, 4, 1)
    c.insert(3, cirq.X(q[2]).with_probability(sympy.Symbol('g1') / 2))  # symbolic channel
    noise = cirq.depolarize(noise_p) if noise_p else None
    want = cirq.DensityMatrixSimulator(seed=5, noise=noise, split_untangled_states=False).run_sweep(
        c, sweep, repetitions=30)
    sim = DM(seed=5, noise=noise, sweep_batch=True)
    got = sim.run_sweep(c, sweep, repetitions=30)
    assert sim.last_run_info['path'] == 'batched sweep'
    for g, w in zip(got, want):
        assert g.params == w.params
        np.testing.assert_array_equal(g.measurements['m'], w.measurements['m'])


def test_batched_sweep_expectation_values(cirq, SV, DM):
    """simulate_expectation_values_sweep with sweep_batch=True equals the
    reference's per-resolver loop (sim/sparse_simulator.py:193-218,
    sim/density_matrix_simulator.py:204-235)."""
    c, q, sweep = _qaoa_like(cirq, 6, 2)
    c = c[:-1]  # no measurement
    obs = [cirq.Z(q[0]) * cirq.Z(q[3]), 0.5 * cirq.X(q[1]) - 1.5 * cirq.Y(q[2]) * cirq.Z(q[5]),
           cirq.PauliString()]
    want = cirq.Simulator(dtype=np.complex128).simulate_expectation_values_sweep(c, obs, sweep)
    sim = SV(dtype=np.complex128, sweep_batch=True)
    got = sim.simulate_expectation_values_sweep(c, obs, sweep)
    assert sim.last_run_info['path'] == 'batched sweep'
    np.testing.assert_allclose(got, want, atol=1e-10)
    noise = cirq.depolarize(0.03)
    want = cirq.DensityMatrixSimulator(dtype=np.complex128, noise=noise).simulate_expectation_values_sweep(
        c, obs, sweep)
    dsim = DM(dtype=np.complex128, noise=noise, sweep_batch=True)
    got = dsim.simulate_expectation_values_sweep(c, obs, sweep)
    assert dsim.last_run_info['path'] == 'batched sweep'
    np.testing.assert_allclose(got, want, atol=1e-10)
    # a measured circuit takes the reference loop
    c2 = c + cirq.Circuit(cirq.measure(q[0], key='m'))
    with pytest.raises(ValueError, match='terminal measurements'):
        SV(sweep_batch=True).simulate_expectation_values_sweep(c2, obs, sweep)


def test_batched_simulate_sweep(cirq, SV, DM):
    """simulate_sweep with sweep_batch=True: every resolver's final state equals
    the reference's (sim/simulator_base.py:277-320)."""
    c, q, sweep = _qaoa_like(cirq, 6, 5)
    c = c[:-1]
    want = cirq.Simulator(dtype=np.complex128).simulate_sweep(c, sweep)
    sim = SV(dtype=np.complex128, sweep_batch=True)
    got = sim.simulate_sweep(c, sweep)
    assert sim.last_run_info['path'] == 'batched sweep' and len(got) == len(want)
    for g, w in zip(got, want):
        assert g.params == w.params and g.qubit_map == w.qubit_map
        np.testing.assert_allclose(g.final_state_vector, w.final_state_vector, atol=1e-12)
    noise = cirq.depolarize(0.02)
    want = cirq.DensityMatrixSimulator(dtype=np.complex128, noise=noise).simulate_sweep(c, sweep)
    dsim = DM(dtype=np.complex128, noise=noise, sweep_batch=True)
    got = dsim.simulate_sweep(c, sweep)
    assert dsim.last_run_info['path'] == 'batched sweep'
    for g, w in zip(got, want):
        np.testing.assert_allclose(g.final_density_matrix, w.final_density_matrix, atol=1e-12)
    # a given initial state or qubit order takes the reference loop
    want = cirq.Simulator(dtype=np.complex128).simulate_sweep(c, sweep, initial_state=3)
    got = SV(dtype=np.complex128, sweep_batch=True).simulate_sweep(c, sweep, initial_state=3)
    np.testing.assert_allclose(got[2].final_state_vector, want[2].final_state_vector, atol=1e-12)


@pytest.mark.parametrize('seed', [1, 2])
def test_differential_fuzz_state_vector(cirq, SV, seed):
    """Random circuits over a gate set rich in diagonal gates, SWAPs and 3-qubit
    gates, random sizes / fusion widths / split settings: final states, device-side
    amplitude gathers and seeded samples against cirq.Simulator."""
    rng = np.random.RandomState(seed)
    domain = {cirq.CNOT: 2, cirq.CZ: 2, cirq.H: 1, cirq.ISWAP: 2, cirq.CZPowGate(exponent=0.3): 2,
              cirq.S: 1, cirq.SWAP: 2, cirq.T: 1, cirq.X: 1, cirq.Y ** 0.5: 1, cirq.Z: 1, cirq.CCZ: 3,
              cirq.ZZPowGate(exponent=0.7): 2, cirq.FSimGate(0.4, 0.9): 2, cirq.rz(0.3): 1}
    for trial in range(60):
        n = int(rng.randint(2, 10))
        q = cirq.LineQubit.range(n)
        c = cirq.testing.random_circuit(
            q, int(rng.randint(3, 25)), float(rng.uniform(0.4, 1.0)),
            gate_domain={g: k for g, k in domain.items() if k <= n}, random_state=int(rng.randint(1 << 30)))
        c.append(cirq.I.on_each(*q))
        split = bool(rng.randint(2))
        dtype = [np.complex64, np.complex128][rng.randint(2)]
        mf = [None, 2, 3, 4][rng.randint(4)]
        atol = 2e-5 if dtype == np.complex64 else 1e-11
        want = cirq.Simulator(dtype=dtype, split_untangled_states=split).simulate(c, qubit_order=q).final_state_vector
        sim = SV(dtype=dtype, split_untangled_states=split, max_fused_qubits=mf)
        idx = [int(x) for x in rng.randint(0, 1 << n, size=4)]
        amps = sim.compute_amplitudes(c, idx, qubit_order=q)
        got = sim.simulate(c, qubit_order=q).final_state_vector
        assert np.max(np.abs(got - want)) <= atol, (trial, n, split, dtype, mf)
        assert np.max(np.abs(np.array(amps) - want[idx])) <= atol, (trial, n, split, dtype, mf)
        c2 = c + cirq.Circuit(cirq.measure(*q, key='m'))
        a = SV(dtype=dtype, seed=trial, split_untangled_states=split, max_fused_qubits=mf).run(c2, repetitions=20)
        b = cirq.Simulator(dtype=dtype, seed=trial, split_untangled_states=split).run(c2, repetitions=20)
        # (complex64 rounding can move a draw across a CDF bin edge: rare, never systematic)
        assert np.mean(a.measurements['m'] != b.measurements['m']) <= 0.05, (trial, n, split, dtype)


def test_differential_fuzz_noisy_state_vector_seeded(cirq, SV):
    """Noisy / mid-circuit-measured random circuits: the per-repetition loop
    consumes the simulator's random stream exactly like cirq.Simulator (mixtures,
    Kraus trajectories, resets, measurements, noise models), so seeded records agree."""
    rng = np.random.RandomState(1)
    domain = {cirq.CNOT: 2, cirq.CZ: 2, cirq.H: 1, cirq.ISWAP: 2, cirq.S: 1, cirq.SWAP: 2, cirq.T: 1,
              cirq.X: 1, cirq.Y ** 0.5: 1, cirq.FSimGate(0.4, 0.9): 2}
    chans = [cirq.depolarize(0.2), cirq.amplitude_damp(0.3), cirq.phase_damp(0.3), cirq.bit_flip(0.25),
             cirq.phase_flip(0.2), cirq.asymmetric_depolarize(0.1, 0.1, 0.2), cirq.depolarize(0.1, n_qubits=2),
             cirq.reset]
    for trial in range(25):
        n = int(rng.randint(1, 6))
        q = cirq.LineQubit.range(n)
        c = cirq.testing.random_circuit(
            q, int(rng.randint(2, 10)), 0.8, gate_domain={g: k for g, k in domain.items() if k <= n},
            random_state=int(rng.randint(1 << 30)))
        for _ in range(rng.randint(1, 6)):
            ch = chans[rng.randint(len(chans))]
            if ch is cirq.reset:
                c.insert(int(rng.randint(0, len(c) + 1)), cirq.reset(q[rng.randint(n)]))
                continue
            k = cirq.num_qubits(ch)
            if k <= n:
                c.insert(int(rng.randint(0, len(c) + 1)), ch.on(*[q[i] for i in rng.permutation(n)[:k]]))
        if rng.randint(2):
            c.insert(int(rng.randint(0, len(c) + 1)), cirq.measure(q[rng.randint(n)], key='mid'))
        c.append(cirq.measure(*q, key='m'))
        noise = [None, cirq.depolarize(0.05), cirq.bit_flip(0.1)][rng.randint(3)]
        split = bool(rng.randint(2))
        a = SV(noise=noise, seed=trial, split_untangled_states=split).run(c, repetitions=12)
        b = cirq.Simulator(noise=noise, seed=trial, split_untangled_states=split).run(c, repetitions=12)
        for key in b.records:
            # (a complex64 rounding difference at a probability bin edge may flip a draw)
            assert np.mean(a.records[key] != b.records[key]) <= 0.1, (trial, n, key, split, noise)


def test_differential_fuzz_density_matrix(cirq, DM):
    """Random circuits with channels sprinkled in (and optionally a noise model):
    final density matrices and seeded samples against cirq.DensityMatrixSimulator."""
    rng = np.random.RandomState(4)
    domain = {cirq.CNOT: 2, cirq.CZ: 2, cirq.H: 1, cirq.ISWAP: 2, cirq.S: 1, cirq.SWAP: 2, cirq.T: 1,
              cirq.X: 1, cirq.Y ** 0.5: 1, cirq.ZZPowGate(exponent=0.7): 2, cirq.FSimGate(0.4, 0.9): 2}
    chans = [cirq.depolarize(0.1), cirq.amplitude_damp(0.2), cirq.phase_damp(0.3), cirq.bit_flip(0.15),
             cirq.depolarize(0.05, n_qubits=2)]
    for trial in range(40):
        n = int(rng.randint(1, 6))
        q = cirq.LineQubit.range(n)
        c = cirq.testing.random_circuit(
            q, int(rng.randint(2, 12)), 0.8, gate_domain={g: k for g, k in domain.items() if k <= n},
            random_state=int(rng.randint(1 << 30)))
        for _ in range(rng.randint(1, 6)):
            ch = chans[rng.randint(len(chans))]
            k = cirq.num_qubits(ch)
            if k <= n:
                c.insert(int(rng.randint(0, len(c) + 1)), ch.on(*[q[i] for i in rng.permutation(n)[:k]]))
        c.append(cirq.I.on_each(*q))
        noise = [None, cirq.depolarize(0.02)][rng.randint(2)]
        split = bool(rng.randint(2))
        dtype = [np.complex64, np.complex128][rng.randint(2)]
        atol = 2e-5 if dtype == np.complex64 else 1e-11
        want = cirq.DensityMatrixSimulator(dtype=dtype, noise=noise, split_untangled_states=split).simulate(
            c, qubit_order=q).final_density_matrix
        got = DM(dtype=dtype, noise=noise, split_untangled_states=split).simulate(c, qubit_order=q).final_density_matrix
        assert np.max(np.abs(got - want)) <= atol, (trial, n, split, dtype, noise)
        c2 = c + cirq.Circuit(cirq.measure(*q, key='m'))
        a = DM(dtype=dtype, noise=noise, seed=trial, split_untangled_states=split).run(c2, repetitions=15)
        b = cirq.DensityMatrixSimulator(dtype=dtype, noise=noise, seed=trial, split_untangled_states=split).run(
            c2, repetitions=15)
        assert np.mean(a.measurements['m'] != b.measurements['m']) <= 0.05, (trial, n, split, dtype)


def test_differential_fuzz_batched_sweeps(cirq, SV, DM, backend):
    """sweep_batch=True against the reference resolver by resolver on random
    symbolic circuits (EigenGate and non-EigenGate symbols, SWAPs, idle qubits,
    noise models): seeded run_sweep records, simulate_sweep final states.  Caught
    two bugs when it was written: a trailing relabelled SWAP that was never put
    back, and noise applied to qubits the reference leaves alone (it hands the
    noise model the qubits of the circuit PART it is walking)."""
    rng = np.random.RandomState(5)
    a, b = sympy.symbols('a b')

    def rand_op(q, n):
        kind = rng.randint(9)
        i, j = rng.permutation(n)[:2] if n > 1 else (0, 0)
        if kind == 0:
            return cirq.rx(a * 2).on(q[i])
        if kind == 1:
            return cirq.rz(b + 0.3).on(q[i])
        if kind == 2 and n > 1:
            return cirq.ZZ(q[i], q[j]) ** a
        if kind == 3 and n > 1:
            return cirq.FSimGate(a, b).on(q[i], q[j])
        if kind == 4 and n > 1:
            return cirq.CZ(q[i], q[j]) ** (b * 0.5)
        if kind == 5 and n > 1:
            return cirq.CNOT(q[i], q[j])
        if kind == 6:
            return cirq.H(q[i])
        if kind == 7 and n > 1:
            return cirq.SWAP(q[i], q[j])
        return cirq.T(q[i])

    for trial in range(36):
        n = int(rng.randint(1, 6))
        q = cirq.LineQubit.range(n)
        body = cirq.Circuit(rand_op(q, n) for _ in range(rng.randint(3, 25)))
        c = body + cirq.Circuit(cirq.measure(*q, key='m'))
        count = int(rng.randint(2, 7))
        sweep = cirq.Zip(cirq.Points('a', list(rng.uniform(0, 1, count))),
                         cirq.Points('b', list(rng.uniform(0, 1, count))))
        dtype = [np.complex64, np.complex128][rng.randint(2)]
        atol = 2e-5 if dtype == np.complex64 else 1e-11
        if rng.randint(2):
            noise = [None, cirq.depolarize(0.03)][rng.randint(2)]
            ref = cirq.DensityMatrixSimulator(dtype=dtype, noise=noise, seed=trial, split_untangled_states=False)
            sim = DM(dtype=dtype, noise=noise, seed=trial, sweep_batch=True)
            final = lambda r: r.final_density_matrix
        else:
            ref = cirq.Simulator(dtype=dtype, seed=trial, split_untangled_states=False)
            sim = SV(dtype=dtype, seed=trial, sweep_batch=True)
            final = lambda r: r.final_state_vector
        want = ref.run_sweep(c, sweep, repetitions=15)
        got = sim.run_sweep(c, sweep, repetitions=15)
        assert sim.last_run_info.get('path') == 'batched sweep'
        for g, w in zip(got, want):
            assert np.mean(g.measurements['m'] != w.measurements['m']) <= 0.1, (trial, n, dtype)
        full = body + cirq.Circuit(cirq.I.on_each(*q))
        for g, w in zip(sim.simulate_sweep(full, sweep), ref.simulate_sweep(full, sweep)):
            assert np.max(np.abs(final(g) - final(w))) <= atol, (trial, n, dtype)
        assert sim.last_run_info.get('path') == 'batched sweep'
        paulis = [cirq.X, cirq.Y, cirq.Z]
        obs = [cirq.PauliString({x: paulis[rng.randint(3)] for x in q if rng.randint(2)},
                                coefficient=float(rng.uniform(-2, 2))) for _ in range(2)]
        obs.append(obs[0] + 0.5 * obs[1])
        got_ev = sim.simulate_expectation_values_sweep(full, obs, sweep)
        want_ev = ref.simulate_expectation_values_sweep(full, obs, sweep)
        assert np.max(np.abs(np.asarray(got_ev) - np.asarray(want_ev))) <= atol * 20, (trial, n, dtype)


def test_batched_sweep_noise_behind_per_qubit_measurements(cirq, DM):
    """Terminal measure-each pattern with a noise model: the reference's terminal
    walk (sim/simulator_base.py:203-209) skips every later operation on an already
    measured qubit tuple — the noise layer behind the measurements included.  A
    batched sweep has to do the same (it used to apply that noise: m0 = 0)."""
    a = sympy.Symbol('a')
    q0, q1 = cirq.LineQubit.range(2)
    c = cirq.Circuit(cirq.rx(a).on(q0), cirq.rx(a).on(q1), cirq.measure(q0, key='m0'),
                     cirq.measure(q1, key='m1'))
    sweep = cirq.Points('a', [0.0, np.pi, 0.0])
    noise = cirq.bit_flip(1.0)
    want = cirq.DensityMatrixSimulator(noise=noise, seed=1).run_sweep(c, sweep, repetitions=8)
    sim = DM(noise=noise, seed=1, sweep_batch=True)
    got = sim.run_sweep(c, sweep, repetitions=8)
    assert sim.last_run_info['path'] == 'batched sweep'
    plain = DM(noise=noise, seed=1).run_sweep(c, sweep, repetitions=8)
    for g, w, p in zip(got, want, plain):
        for key in ('m0', 'm1'):
            np.testing.assert_array_equal(g.measurements[key], w.measurements[key])
            np.testing.assert_array_equal(p.measurements[key], w.measurements[key])


def test_batched_simulate_sweep_keeps_global_phase(cirq, SV):
    """A zero-qubit operation (global phase) is part of the final state vector of a
    batched simulate_sweep, as in the per-resolver path and the reference."""
    a = sympy.Symbol('a')
    q0, q1 = cirq.LineQubit.range(2)
    c = cirq.Circuit(cirq.rx(a).on(q0), cirq.global_phase_operation(1j), cirq.H(q1),
                     cirq.global_phase_operation(np.exp(0.3j)))
    sweep = cirq.Points('a', [0.1, 0.7, 2.0])
    want = cirq.Simulator(dtype=np.complex128).simulate_sweep(c, sweep)
    sim = SV(dtype=np.complex128, sweep_batch=True)
    got = sim.simulate_sweep(c, sweep)
    assert sim.last_run_info['path'] == 'batched sweep'
    for g, w in zip(got, want):
        np.testing.assert_allclose(g.final_state_vector, w.final_state_vector, atol=1e-12)


def test_batched_sweep_symbolic_eigen_components(cirq, SV):
    """EigenGates whose eigen-components depend on a second symbol
    (PhasedISwapPowGate(phase_exponent=a, exponent=b)) take the generic
    per-resolver resolution instead of the vectorised eigen-decomposition."""
    a, b = sympy.symbols('a b')
    q0, q1 = cirq.LineQubit.range(2)
    c = cirq.Circuit(cirq.H(q0), cirq.PhasedISwapPowGate(phase_exponent=a, exponent=b).on(q0, q1),
                     cirq.rx(b).on(q1))
    sweep = cirq.Zip(cirq.Points('a', [0.1, 0.4, 0.9]), cirq.Points('b', [0.3, 0.5, 1.2]))
    want = cirq.Simulator(dtype=np.complex128).simulate_sweep(c, sweep)
    got = SV(dtype=np.complex128, sweep_batch=True).simulate_sweep(c, sweep)
    for g, w in zip(got, want):
        np.testing.assert_allclose(g.final_state_vector, w.final_state_vector, atol=1e-12)
    m = c + cirq.Circuit(cirq.measure(q0, q1, key='m'))
    want = cirq.Simulator(seed=3, split_untangled_states=False).run_sweep(m, sweep, repetitions=20)
    got = SV(seed=3, sweep_batch=True).run_sweep(m, sweep, repetitions=20)
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g.measurements['m'], w.measurements['m'])


def test_batched_trajectories_trailing_swap(cirq, SV):
    """A SWAP right before the measurement is relabelled by the scheduler and has
    to be put back before the records are read."""
    q = cirq.LineQubit.range(3)
    # (the bit flip puts q1 — and with it the SWAP — into the per-repetition part)
    circuit = cirq.Circuit(cirq.X(q[1]), cirq.bit_flip(0.0).on(q[1]), cirq.SWAP(q[1], q[2]),
                           cirq.measure(*q, key='m'))
    sim = SV(seed=1, trajectory_batch=64)
    m = sim.run(circuit, repetitions=50).measurements['m']
    assert sim.last_run_info['path'] == 'batched trajectories'
    assert np.all(m == [0, 0, 1])


def test_mux_entry_points_match_reference(cirq, SV, DM):
    """cirq_b200.sample / final_state_vector / final_density_matrix mirror
    cirq.sample / ... (sim/mux.py) with the same signatures."""
    import cirq_b200

    q = cirq.LineQubit.range(3)
    c = cirq.Circuit(cirq.H(q[0]), cirq.CNOT(q[0], q[1]), cirq.T(q[2]) ** 0.3, cirq.ISWAP(q[1], q[2]) ** 0.5)
    np.testing.assert_allclose(
        cirq_b200.final_state_vector(c, qubit_order=q), cirq.final_state_vector(c, qubit_order=q), atol=1e-6
    )
    np.testing.assert_allclose(
        cirq_b200.final_state_vector(c, initial_state=5, dtype=np.complex128),
        cirq.final_state_vector(c, initial_state=5, dtype=np.complex128), atol=1e-12,
    )
    noisy = c + cirq.Circuit(cirq.measure(q[0], key='m'))
    np.testing.assert_allclose(
        cirq_b200.final_density_matrix(noisy, noise=cirq.depolarize(0.05)),
        cirq.final_density_matrix(noisy, noise=cirq.depolarize(0.05)), atol=1e-6,
    )
    # non-Clifford circuit so the mux does not take its stabilizer shortcut
    m = c + cirq.Circuit(cirq.measure(*q, key='k'))
    got = cirq_b200.sample(m, repetitions=200, seed=4)
    assert got.measurements['k'].shape == (200, 3)
    got = cirq_b200.sample(m, noise=cirq.bit_flip(0.1), repetitions=50, seed=4)
    assert got.measurements['k'].shape == (50, 3)
    import sympy

    s = sympy.Symbol('s')
    sweep_c = cirq.Circuit(cirq.rx(s).on(q[0]), cirq.T(q[0]), cirq.measure(q[0], key='z'))
    res = cirq_b200.sample_sweep(sweep_c, cirq.Linspace('s', 0, 3, 4), repetitions=20, seed=1)
    assert len(res) == 4 and res[0].measurements['z'].shape == (20, 1)
    with pytest.raises(ValueError, match='measurement'):
        cirq_b200.final_state_vector(m)


def test_qvm_simulator_class_contract(cirq, SV, DM):
    """The Google QVM hook (cirq-google/cirq_google/engine/virtual_engine_factory.py:437-478)
    builds ``simulator_class(noise=noise_model, **kwargs)`` and uses the object as the
    processor's ``cirq.Sampler``.  Same construction and calls here with a
    NoiseModelFromNoiseProperties model (the base class of the QVM's
    NoiseModelFromGoogleNoiseProperties: per-gate depolarising + Kraus damping), and
    the real factory when cirq_google is importable."""
    from cirq.devices import noise_properties as nprop
    from cirq.devices.insertion_noise_model import InsertionNoiseModel
    from cirq.devices.noise_utils import OpIdentifier, PHYSICAL_GATE_TAG

    q = cirq.LineQubit.range(3)

    class Props(nprop.NoiseProperties):
        def build_noise_models(self):
            add = {OpIdentifier(cirq.HPowGate, x): cirq.depolarize(0.05).on(x) for x in q}
            add.update({OpIdentifier(cirq.CZPowGate, a, b): cirq.amplitude_damp(0.1).on(b)
                        for a, b in ((q[0], q[1]), (q[1], q[2]))})
            return [InsertionNoiseModel(ops_added=add, require_physical_tag=True)]

        def _value_equality_values_(self):
            return 'props'

    noise_model = nprop.NoiseModelFromNoiseProperties(Props())
    circuit = cirq.Circuit(cirq.H(q[0]), cirq.CZ(q[0], q[1]), cirq.H(q[1]), cirq.CZ(q[1], q[2]),
                           cirq.H(q[2]), cirq.measure(*q, key='m'))
    for simulator_class, ref_class, kwargs in (
            (SV, cirq.Simulator, dict(seed=11)),
            (DM, cirq.DensityMatrixSimulator, dict(seed=11))):
        sampler = simulator_class(noise=noise_model, **kwargs)  # the factory's exact call
        assert isinstance(sampler, cirq.Sampler) and isinstance(sampler, cirq.SimulatesSamples)
        want = ref_class(noise=noise_model, **kwargs).run(circuit, repetitions=40)
        got = sampler.run(circuit, repetitions=40)
        assert got.measurements['m'].shape == (40, 3)
        assert np.mean(got.measurements['m'] != want.measurements['m']) <= 0.05
        batch = sampler.run_batch([circuit, circuit], repetitions=5)  # SimulatedLocalProcessor's path
        assert len(batch) == 2 and batch[0][0].measurements['m'].shape == (5, 3)
    # with trajectory batching (the documented QVM recipe) the distribution is the same
    exact = cirq.DensityMatrixSimulator(noise=noise_model, dtype=np.complex128).simulate(
        circuit[:-1], qubit_order=q).final_density_matrix
    probs = np.real(np.diag(exact))
    m = SV(noise=noise_model, seed=5, trajectory_batch=512).run(circuit, repetitions=4000).measurements['m']
    hist = np.bincount(m @ (1 << np.arange(2, -1, -1)), minlength=8)
    expect = probs * 4000
    chi2 = float(np.sum((hist - expect) ** 2 / np.maximum(expect, 1e-9)))
    assert chi2 < 40, chi2  # 7 degrees of freedom
    try:
        import cirq_google
    except Exception:
        return  # not importable in this image (generated protobuf modules need `tunits`)
    engine = cirq_google.engine.create_default_noisy_quantum_virtual_machine(
        'rainbow', simulator_class=SV, seed=0)
    dev_q = sorted(engine.get_processor('rainbow').get_device().metadata.qubit_set)[:2]
    res = engine.get_sampler('rainbow').run(
        cirq.Circuit(cirq.X(dev_q[0]) ** 0.5, cirq.measure(*dev_q, key='z')), repetitions=20)
    assert res.measurements['z'].shape == (20, 2)


def test_use_b200_is_reentrant_and_thread_safe(cirq, SV):
    import threading

    import cirq_b200
    from cirq.sim import sparse_simulator

    original = sparse_simulator.Simulator
    q = cirq.LineQubit.range(2)
    c = cirq.Circuit(cirq.H(q[0]), cirq.T(q[0]), cirq.CNOT(q[0], q[1]))
    want = cirq.final_state_vector(c)
    with cirq_b200.use_b200():
        with cirq_b200.use_b200():
            assert sparse_simulator.Simulator is SV
        assert sparse_simulator.Simulator is SV  # the inner exit must not restore
        errors = []

        def work():
            try:
                for _ in range(5):
                    np.testing.assert_allclose(cirq_b200.final_state_vector(c), want, atol=1e-6)
            except Exception as exc:  # pragma: no cover
                errors.append(exc)

        threads = [threading.Thread(target=work) for _ in range(3)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors and sparse_simulator.Simulator is SV
    assert sparse_simulator.Simulator is original


def test_forty_qubit_mostly_unentangled_circuit_like_reference(cirq, SV):
    """split_untangled_states is honoured at any register size
    (sim/simulation_product_state.py:68-139; the state-vector analogue of
    density_matrix_simulator_test.py:1502-1520): 40 qubits whose entangled clusters
    stay small never need more than a few amplitudes.  Seeded records equal the
    reference's, mid-circuit measurement and reset included."""
    q = cirq.LineQubit.range(40)
    rng = np.random.RandomState(2)
    ops_list = [cirq.H(x) for x in q[::3]] + [cirq.X(q[7]) ** 0.3, cirq.Y(q[22]) ** 0.7]
    for start in (0, 9, 18, 30):  # four 4-qubit clusters
        cluster = q[start:start + 4]
        for _ in range(6):
            a, b = rng.choice(4, 2, replace=False)
            ops_list.append(cirq.FSimGate(rng.uniform(0, 2), rng.uniform(0, 2)).on(cluster[a], cluster[b]))
            ops_list.append(cirq.rx(rng.uniform(0, 3)).on(cluster[a]))
    ops_list += [cirq.SWAP(q[5], q[38]), cirq.CZ(q[13], q[14]) ** 0.4, cirq.measure(q[10], key='mid'),
                 cirq.ResetChannel().on(q[11]), cirq.CNOT(q[10], q[12])]
    # (per-qubit terminal measurements: with a mid-circuit measurement every repetition
    # walks the circuit, and a JOINT measurement of 40 qubits would join them all — in
    # the reference too)
    circuit = cirq.Circuit(ops_list) + cirq.Circuit(cirq.measure(x, key=f'm{i}') for i, x in enumerate(q))
    want = cirq.Simulator(seed=9).run(circuit, repetitions=6)
    got = SV(seed=9).run(circuit, repetitions=6)
    for key in ['mid'] + [f'm{i}' for i in range(40)]:
        np.testing.assert_array_equal(got.measurements[key], want.measurements[key])
    # all measurements terminal: one joint key, sampled sub-state by sub-state
    terminal = cirq.Circuit(op for op in ops_list if not cirq.is_measurement(op)
                            and not isinstance(op.gate, cirq.ResetChannel))
    terminal += cirq.Circuit(cirq.measure(*q, key='m'))
    want = cirq.Simulator(seed=3).run(terminal, repetitions=50)
    got = SV(seed=3).run(terminal, repetitions=50)
    np.testing.assert_array_equal(got.measurements['m'], want.measurements['m'])
    # the final state stays a product of small sub-states (never 2^40 amplitudes)
    sim = SV(seed=1)
    final = sim.simulate(cirq.Circuit(ops_list))._final_simulator_state
    comps = list({id(v): v for k, v in final.sim_states.items() if k is not None}.values())
    assert max(len(v.qubits) for v in comps) <= 6 and len(comps) >= 10
    ref = cirq.Simulator(seed=1).simulate(cirq.Circuit(ops_list))._final_simulator_state
    for qubit in (q[0], q[9], q[19], q[33]):
        a, b = final.sim_states[qubit], ref.sim_states[qubit]
        assert a.qubits == b.qubits
        np.testing.assert_allclose(a._state.to_numpy_tensor().reshape(-1), b.target_tensor.reshape(-1), atol=1e-6)


def test_product_state_densifies_before_joins_outgrow_memory(cirq, SV, monkeypatch):
    """Join policy of B200ProductState above _DENSE_JOIN_BITS (exercised here with the
    thresholds lowered): the first join that would pass the limit merges the whole
    register as (largest sub-state) x (everything else), later operations act on
    the one dense state, large states are not factored back after measurements, and
    the final merge transposes in place — results equal the reference's."""
    import cirq_b200.sv_simulator as svm

    monkeypatch.setattr(svm, '_DENSE_JOIN_BITS', 4)
    monkeypatch.setattr(svm, '_MAX_DENSE_QUBITS', 9)
    monkeypatch.setattr(svm, '_MAX_FACTOR_BITS', 4)
    q = cirq.LineQubit.range(9)
    circuit = cirq.testing.random_circuit(q, 14, 0.8, random_state=6)
    circuit.append(cirq.global_phase_operation(np.exp(0.7j)))  # lives in the empty [None] state
    for dtype, atol in ((np.complex64, 1e-5), (np.complex128, 1e-12)):
        want = cirq.Simulator(dtype=dtype).simulate(circuit, qubit_order=q)
        sim = SV(dtype=dtype)
        got = sim.simulate(circuit, qubit_order=q)
        np.testing.assert_allclose(got.final_state_vector, want.final_state_vector, atol=atol, rtol=0)
        comps = {id(v) for k, v in got._final_simulator_state.sim_states.items() if k is not None}
        assert len(comps) == 1  # densified
    noisy = circuit + cirq.Circuit(cirq.measure(q[2], key='a'), cirq.H(q[2]), cirq.CNOT(q[2], q[5]),
                                   cirq.measure(*q, key='m'))
    want = cirq.Simulator(seed=4, split_untangled_states=False).run(noisy, repetitions=10)
    got = SV(seed=4).run(noisy, repetitions=10)
    for key in ('a', 'm'):
        np.testing.assert_array_equal(got.measurements[key], want.measurements[key])


def test_relabelled_swaps_put_back_by_inplace_permutation(cirq, SV, monkeypatch):
    """With several SWAP gates relabelled by the scheduler, reading the raw state
    restores the bit order with the in-place permutation kernel (here from 3 qubits
    on) instead of real SWAP gates; same state as the reference."""
    import cirq_b200.sv_simulator as svm

    monkeypatch.setattr(svm, '_PERMUTE_RESTORE_MIN_BITS', 3)
    q = cirq.LineQubit.range(7)
    rng = np.random.RandomState(8)
    ops_list = [cirq.H.on_each(*q)]
    for i in range(12):
        a, b = rng.choice(7, 2, replace=False)
        ops_list.append(cirq.SWAP(q[a], q[b]))
        ops_list.append(cirq.CZ(q[a], q[(a + 1) % 7]) ** rng.uniform(0, 1) if a != (a + 1) % 7 else cirq.T(q[a]))
        ops_list.append(cirq.rx(rng.uniform(0, 3)).on(q[b]))
    circuit = cirq.Circuit(ops_list)
    for dtype, atol in ((np.complex64, 1e-5), (np.complex128, 1e-12)):
        want = cirq.Simulator(dtype=dtype).simulate(circuit, qubit_order=q)
        got = SV(dtype=dtype, split_untangled_states=False).simulate(circuit, qubit_order=q)
        np.testing.assert_allclose(got.final_state_vector, want.final_state_vector, atol=atol, rtol=0)
        steps = list(SV(dtype=dtype, split_untangled_states=False).simulate_moment_steps(circuit, qubit_order=q))
        np.testing.assert_allclose(steps[-1].state_vector(), want.final_state_vector, atol=atol, rtol=0)
