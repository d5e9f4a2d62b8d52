"""Pins the CPU oracle (oracle/sv_oracle.py) to the reference: every function
is checked against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import sv_oracle as orc


def tol(dtype):
    return 2e-6 if np.dtype(dtype) == np.complex64 else 1e-13


def test_apply_matrix_matches_targeted_left_multiply():
    g = load_golden('targeted_left_multiply.npz')
    for c in range(int(g['num_cases'])):
        n = int(g[f'c{c}_n'])
        state = g[f'c{c}_state']
        out = orc.apply_matrix(state, n, g[f'c{c}_matrix'], orc.axes_to_bits(n, g[f'c{c}_axes']))
        np.testing.assert_allclose(out, g[f'c{c}_out'], atol=tol(state.dtype) * 4, rtol=0)


def test_gate_lists_match_simulator_final_states():
    g = load_golden('simulator_final_states.npz')
    for c in range(int(g['num_cases'])):
        n = int(g[f'c{c}_n'])
        final = g[f'c{c}_final']
        gates = [
            (g[f'c{c}_g{i}_u'], orc.axes_to_bits(n, g[f'c{c}_g{i}_axes']))
            for i in range(int(g[f'c{c}_num_gates']))
        ]
        psi = orc.run_gate_list(n, gates, dtype=final.dtype)
        atol = 1e-5 if final.dtype == np.complex64 else 1e-12
        np.testing.assert_allclose(psi, final, atol=atol, rtol=0)


def test_sampling_matches_reference_choice_and_digit_order():
    g = load_golden('sampling.npz')
    for c in range(int(g['num_cases'])):
        n = int(g[f'c{c}_n'])
        state = g[f'c{c}_state']
        bits = orc.axes_to_bits(n, g[f'c{c}_indices'])
        probs = orc.marginal_probs(state, n, bits)
        np.testing.assert_allclose(probs / probs.sum(), g[f'c{c}_probs'], atol=1e-6, rtol=0)
        got = orc.sample(state, n, bits, g[f'c{c}_uniforms'])
        want = g[f'c{c}_bits']
        # identical up to float rounding of the cdf at bin edges
        assert np.mean(np.any(got != want, axis=1)) <= 1 / 64 + 1e-9
        assert got.dtype == np.uint8 and got.shape == want.shape


def test_measure_collapse_matches_reference():
    g = load_golden('sampling.npz')
    for c in range(int(g['num_cases'])):
        n = int(g[f'c{c}_n'])
        state = g[f'c{c}_state']
        bits = orc.axes_to_bits(n, g[f'c{c}_indices'])
        m = len(bits)
        probs = orc.marginal_probs(state, n, bits)
        pick = int(orc.choice_indices(probs, np.array([g[f'c{c}_measure_uniform']]))[0])
        values = [(pick >> (m - 1 - q)) & 1 for q in range(m)]
        assert values == g[f'c{c}_measure_bits'].tolist()
        out = orc.collapse(state, n, bits, values, probs[pick] / probs.sum())
        np.testing.assert_allclose(out, g[f'c{c}_measure_state'], atol=tol(state.dtype) * 8, rtol=0)


def test_density_matrix_channels_match_reference():
    g = load_golden('density_matrix_final_states.npz')
    for c in range(int(g['num_cases'])):
        n = int(g[f'c{c}_n'])
        final = g[f'c{c}_final']
        rho = np.zeros(1 << (2 * n), dtype=final.dtype)
        rho[0] = 1
        for i in range(int(g[f'c{c}_num_ops'])):
            rho = orc.dm_apply_channel(
                rho, n, list(g[f'c{c}_g{i}_kraus']), orc.axes_to_bits(n, g[f'c{c}_g{i}_axes'])
            )
        atol = 1e-5 if final.dtype == np.complex64 else 1e-12
        np.testing.assert_allclose(rho.reshape(final.shape), final, atol=atol, rtol=0)
        np.testing.assert_allclose(orc.dm_diagonal(rho, n).sum(), 1.0, atol=1e-5)


def test_pauli_expectation_matches_reference():
    g = load_golden('pauli_expectation.npz')
    for c in range(int(g['num_cases'])):
        n = int(g[f'c{c}_n'])
        codes = g[f'c{c}_codes']
        x = z = 0
        for axis, code in enumerate(codes):
            b = n - 1 - axis
            if code in (1, 2):
                x |= 1 << b
            if code in (2, 3):
                z |= 1 << b
        val = orc.pauli_expectation(g[f'c{c}_state'], n, x, z)
        np.testing.assert_allclose(val, complex(g[f'c{c}_value']), atol=1e-12)


def test_reference_known_answer_vectors():
    g = load_golden('reference_test_vectors.npz')
    # sim/state_vector_test.py:62-86: |x> sampled on [2,1,0] gives reversed bits
    for x in range(8):
        state = np.zeros(8, dtype=np.complex64)
        state[x] = 1
        got = orc.sample(state, 3, orc.axes_to_bits(3, [2, 1, 0]), np.array([0.5]))
        np.testing.assert_array_equal(got, g['big_endian_samples'][x])
    state = np.zeros(8, dtype=np.complex64)
    state[6] = 1
    for perm, want in zip(g['perms'], g['perm_samples']):
        got = orc.sample(state, 3, orc.axes_to_bits(3, perm), np.array([0.3]))
        np.testing.assert_array_equal(got, want)


def test_collapse_and_diagonal_helpers():
    rng = np.random.RandomState(0)
    n = 4
    psi = (rng.standard_normal(16) + 1j * rng.standard_normal(16)).astype(np.complex128)
    psi /= np.linalg.norm(psi)
    d = np.exp(1j * rng.standard_normal(4))
    out = orc.apply_diagonal(psi, n, d, [3, 0])
    np.testing.assert_allclose(out, orc.apply_matrix(psi, n, np.diag(d), [3, 0]), atol=1e-14)
    rho = np.outer(psi, psi.conj()).reshape(-1)
    p = orc.marginal_probs(psi, n, [2])
    got = orc.dm_collapse(rho, n, [2], [1], p[1])
    c = orc.collapse(psi, n, [2], [1], p[1])
    np.testing.assert_allclose(got.reshape(16, 16), np.outer(c, c.conj()), atol=1e-13)


def test_dist_pack_roundtrip():
    rng = np.random.RandomState(1)
    s = rng.standard_normal(64) + 1j * rng.standard_normal(64)
    for bits in ([5], [0, 3], [4, 1, 2]):
        p = orc.dist_pack(s, 6, bits)
        np.testing.assert_array_equal(orc.dist_unpack(p, 6, bits), s)
        seg = len(s) >> len(bits)
        # segment 0 has all the chosen bits clear
        i = np.arange(64)
        mask = sum(1 << b for b in bits)
        np.testing.assert_array_equal(p[:seg], s[(i & mask) == 0])


def test_reduced_density_matrix_matches_reference():
    g = load_golden('reduced_density_matrix.npz')
    for c in range(int(g['num_cases'])):
        n = int(g[f'c{c}_n'])
        bits = [n - 1 - int(a) for a in g[f'c{c}_indices']]
        rho = orc.reduced_density_matrix(g[f'c{c}_state'], n, bits)
        np.testing.assert_allclose(rho, g[f'c{c}_rho'], atol=1e-13)
        r1 = orc.reduced_density_matrix(g[f'c{c}_state'], n, bits[:1])
        bloch = [2 * r1[0, 1].real, 2 * r1[1, 0].imag, (r1[0, 0] - r1[1, 1]).real]
        np.testing.assert_allclose(bloch, g[f'c{c}_bloch'], atol=1e-6)  # reference returns float32


def test_trajectory_ops_match_reference():
    """bsv_* oracle functions against the reference's per-repetition pieces."""
    g = load_golden('trajectory_ops.npz')
    for c in range(int(g['num_cases'])):
        n = int(g[f'c{c}_n'])
        states = g[f'c{c}_states']
        B = states.shape[0]
        bits = [n - 1 - int(a) for a in g[f'c{c}_axes']]
        kraus = g[f'c{c}_kraus']
        w = orc.bsv_kraus_weights(states.reshape(-1), n, kraus, bits)
        np.testing.assert_allclose(w, g[f'c{c}_weights'], atol=1e-13)
        for i in range(len(kraus)):
            got = orc.bsv_apply_select(states.reshape(-1), n, kraus, bits, np.full(B, i))
            np.testing.assert_allclose(got.reshape(B, -1), g[f'c{c}_applied'][:, i], atol=1e-13)
        # skip index leaves trajectories untouched; scale multiplies
        choice = np.arange(B) % len(kraus)
        scale = 1.0 / np.sqrt(w[np.arange(B), choice])
        got = orc.bsv_apply_select(states.reshape(-1), n, kraus, bits, choice, scale, skip=0).reshape(B, -1)
        for t in range(B):
            want = states[t] if choice[t] == 0 else g[f'c{c}_applied'][t, choice[t]] * scale[t]
            np.testing.assert_allclose(got[t], want, atol=1e-13)
        results = g[f'c{c}_results']
        probs = np.array([
            orc.marginal_probs(states[t], n, bits)[int(''.join(str(int(b)) for b in results[t]), 2)]
            for t in range(B)
        ])
        got = orc.bsv_collapse(states.reshape(-1), n, bits, results, 1.0 / np.sqrt(probs)).reshape(B, -1)
        np.testing.assert_allclose(got, g[f'c{c}_collapsed'].reshape(B, -1), atol=1e-12)


def _masks(codes, n):
    x = z = 0
    for axis, code in enumerate(codes):
        b = n - 1 - axis
        if code in (1, 2):
            x |= 1 << b
        if code in (2, 3):
            z |= 1 << b
    return x, z


def test_dm_pauli_expectation_matches_reference():
    g = load_golden('dm_pauli_expectation.npz')
    for c in range(int(g['num_cases'])):
        n = int(g[f'c{c}_n'])
        x, z = _masks(g[f'c{c}_codes'], n)
        val = orc.dm_pauli_expectation(g[f'c{c}_rho'].reshape(-1), n, x, z)
        np.testing.assert_allclose(val.real, float(g[f'c{c}_value']), atol=1e-12)
        assert abs(val.imag) < 1e-12
