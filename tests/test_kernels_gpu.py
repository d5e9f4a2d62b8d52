"""Parity of the CUDA path (through the C-ABI, via cirq_b200.DeviceState)
against the CPU oracle and the committed reference goldens.  GPU only.

Tolerances (north star): max-abs amplitude error 1e-5 for complex64 and
1e-12 for complex128 on normalised states; index/bit work is bit-exact."""
import itertools

import numpy as np
import pytest

from conftest import load_golden
from oracle import sv_oracle as orc

pytestmark = pytest.mark.gpu

ATOL = {np.dtype(np.complex64): 1e-5, np.dtype(np.complex128): 1e-12}


@pytest.fixture(scope='module')
def DS():
    from cirq_b200.device_state import DeviceState

    return DeviceState


def rand_state(rng, n, dtype):
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    v /= np.linalg.norm(v)
    return v.astype(dtype)


def rand_unitary(rng, k):
    d = 1 << k
    q, r = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
    return q * (np.diag(r) / np.abs(np.diag(r)))


def rand_matrix(rng, k):
    d = 1 << k
    return (rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))) / np.sqrt(d)


def target_sets(n, k, rng, count):
    """Mix of hand-picked position classes (vector bit, zone bits, high bits)
    and random draws, as ordered tuples (gate order matters)."""
    sets = set()
    pools = [list(range(min(n, 7))), list(range(max(0, n - 6), n)), list(range(n))]
    for pool in pools:
        if len(pool) >= k:
            for _ in range(count):
                sets.add(tuple(rng.permutation(pool)[:k].tolist()))
    low = tuple(range(k))
    if k <= n:
        sets.add(low)
        sets.add(low[::-1])
        sets.add(tuple(range(n - k, n)))
    return sorted(sets)


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
@pytest.mark.parametrize('n', [1, 2, 3, 5, 8, 11, 12, 13, 16, 21])
def test_apply_matrix_matches_oracle(DS, dtype, n):
    rng = np.random.RandomState(1000 + n)
    max_k = 5 if dtype == np.complex64 else 4
    for k in range(1, min(n, max_k) + 1):
        for targets in target_sets(n, k, rng, 3 if n < 20 else 1):
            state = rand_state(rng, n, dtype)
            m = rand_matrix(rng, k)
            dev = DS.from_numpy(state)
            dev.apply_matrix(m, targets)
            got = dev.to_numpy()
            want = orc.apply_matrix(state, n, m, targets)
            err = np.max(np.abs(got - want))
            assert err <= ATOL[np.dtype(dtype)], (n, k, targets, err)


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_every_single_and_pair_position(DS, dtype):
    """All 1-qubit positions and all ordered 2-qubit pairs at n=13: covers the
    vector bit, every lane bit, the zone boundary and high bits."""
    n = 13
    rng = np.random.RandomState(7)
    state = rand_state(rng, n, dtype)
    for t in range(n):
        m = rand_matrix(rng, 1)
        dev = DS.from_numpy(state)
        dev.apply_matrix(m, [t])
        assert np.max(np.abs(dev.to_numpy() - orc.apply_matrix(state, n, m, [t]))) <= ATOL[np.dtype(dtype)]
    for a, b in itertools.permutations(range(n), 2):
        m = rand_matrix(rng, 2)
        dev = DS.from_numpy(state)
        dev.apply_matrix(m, [a, b])
        err = np.max(np.abs(dev.to_numpy() - orc.apply_matrix(state, n, m, [a, b])))
        assert err <= ATOL[np.dtype(dtype)], (a, b, err)


@pytest.mark.parametrize('k', [4, 5, 6])
def test_tensor_core_kernels_match_oracle(DS, k):
    """tcgen05 path (complex64, k = 4, 5 and 6): every position class, 3xTF32 split
    must stay within the complex64 tolerance."""
    from cirq_b200 import _lib

    rng = np.random.RandomState(60 + k)
    if k == 4:  # opt-in instantiation (the default keeps 4-qubit blocks on the CUDA cores)
        _lib.load().b2q_set_tc_mode(2)
    try:
        _tensor_core_cases(DS, k, rng)
    finally:
        _lib.load().b2q_set_tc_mode(1)


def _tensor_core_cases(DS, k, rng):
    for n in (k + 7, 16, 20):
        sets = [list(range(n - k, n)), list(range(k)), list(range(2, 2 + k))]
        # low index bits pick the 16-byte / pair-exchange access patterns
        sets += [list(range(1, 1 + k)), [0] + list(range(n - k + 1, n)), [1] + list(range(n - k + 1, n)),
                 [0, 1] + list(range(n - k + 2, n)), [n - 1, 0, 3, 5, 2, 7][:k], [4, 1, n - 2, 6, 9, 3][:k]]
        sets += [rng.permutation(n)[:k].tolist() for _ in range(4)]
        for targets in sets:
            state = rand_state(rng, n, np.complex64)
            m = rand_unitary(rng, k)
            dev = DS.from_numpy(state)
            dev.apply_matrix(m, targets)
            # against the float64 result: the error is the kernel's own (3xTF32 split
            # + fp32 accumulation), measured 3e-7 of the norm per pass (DESIGN.md
            # 3.1b); the bound is 4x that, so a broken split (1e-4) cannot hide
            want = orc.apply_matrix(state.astype(np.complex128), n, m, targets)
            diff = dev.to_numpy().astype(np.complex128) - want
            rel_l2 = np.linalg.norm(diff) / np.linalg.norm(want)
            assert rel_l2 <= 1.2e-6, (n, targets, rel_l2)
            # max-abs: amplitudes are ~2^(-n/2); 1.5e-5 of that (5 sigma of the above
            # over 2^20 amplitudes), 40x tighter than the north-star 1e-5 at n = 12
            assert np.max(np.abs(diff)) <= 1.5e-5 * 2.0 ** (-n / 2), (n, targets, np.max(np.abs(diff)))
    # accumulated error over a long run of unitary blocks (the depth of a 34-qubit
    # circuit's schedule): norm drift and distance from the float64 evolution
    n = 18
    dev = DS.basis(n, np.complex64, 1)
    ref = np.zeros(1 << n, dtype=np.complex128)
    ref[1] = 1
    for _ in range(40):
        u, t = rand_unitary(rng, k), rng.permutation(n)[:k].tolist()
        dev.apply_matrix(u, t)
        ref = orc.apply_matrix(ref, n, u, t)
    assert abs(dev.norm2() - 1.0) < 2e-5
    diff = dev.to_numpy().astype(np.complex128) - ref
    assert np.linalg.norm(diff) <= 40 * 3e-7, np.linalg.norm(diff)
    assert np.max(np.abs(diff)) <= 1e-5 * 2.0 ** (-n / 2) * 4, np.max(np.abs(diff))


def test_tile_kernel_two_blocks_per_pass_matches_oracle(DS):
    """b2q_sv_apply_tile_blocks: two blocks applied in one pass over HBM (tile
    staged in shared memory, both blocks on tcgen05) against the float64 oracle
    applying them one after the other.  Same error bound as the one-block
    tensor-core kernels, per block."""
    rng = np.random.RandomState(77)
    for n in (13, 14, 16, 21):
        cases = [([0, 1, 2, 3, 4], [5, 6, 7, 8, 9]), (list(range(n - 5, n)), list(range(n - 10, n - 5))),
                 ([n - 1, 0, 5, 2, 9], [9, 2, n - 2, 3, 7]), ([2, 3, 4, 5, 6], [6, 5, 4, 3, 2]),
                 ([1, 3, 5, 7, 9], [0, 2, 4, 6, 8]), ([4, n - 1, 8, 6, 11], [n - 2, 5, 4, 10, 3]),
                 ([3, 7], [n - 1, 7, 2]), ([0], [1, 0, n - 1, 5]), ([n - 1, n - 2, n - 3, n - 4], [0, 1])]
        cases += [(rng.permutation(n)[:rng.randint(1, 6)].tolist(), rng.permutation(n)[:rng.randint(1, 6)].tolist())
                  for _ in range(8)]
        for ta, tb in cases:
            if len(set(ta)) != len(ta) or len(set(tb)) != len(tb):
                continue  # (hand-picked sets collide at the smallest n)
            state = rand_state(rng, n, np.complex64)
            ma, mb = rand_unitary(rng, len(ta)), rand_unitary(rng, len(tb))
            dev = DS.from_numpy(state)
            dev.apply_tile_blocks([(ma, ta), (mb, tb)])
            want = orc.apply_matrix(orc.apply_matrix(state.astype(np.complex128), n, ma, ta), n, mb, tb)
            diff = dev.to_numpy().astype(np.complex128) - want
            rel = np.linalg.norm(diff) / np.linalg.norm(want)
            assert rel <= 2 * 1.2e-6, (n, ta, tb, rel)
            assert np.max(np.abs(diff)) <= 2 * 1.5e-5 * 2.0 ** (-n / 2), (n, ta, tb)
        # one block alone is a valid tile pass
        state = rand_state(rng, n, np.complex64)
        m, t = rand_unitary(rng, 5), rng.permutation(n)[:5].tolist()
        dev = DS.from_numpy(state)
        dev.apply_tile_blocks([(m, t)])
        want = orc.apply_matrix(state.astype(np.complex128), n, m, t)
        assert np.linalg.norm(dev.to_numpy() - want) / np.linalg.norm(want) <= 1.2e-6
    # a long run through apply_batch at a size where pairing is on: pairs, singles and
    # diagonal blocks mixed, against the oracle gate by gate
    n = 22
    assert DS.basis(n, np.complex64, 0).tile_pairing()
    gates = []
    for i in range(14):
        k = int(rng.randint(2, 6))
        gates.append((rand_unitary(rng, k), rng.permutation(n)[:k].tolist()))
        if i % 5 == 4:
            gates.append((np.exp(1j * rng.standard_normal(1 << 6)), rng.permutation(n)[:6].tolist()))
    dev = DS.basis(n, np.complex64, 3)
    passes = dev.plan_passes(gates)
    assert any(len(g) == 2 for g in passes) and sum(len(g) for g in passes) == len(gates)
    dev.apply_batch(gates)
    want = np.zeros(1 << n, dtype=np.complex128)
    want[3] = 1
    for m, b in gates:
        want = orc.apply_diagonal(want, n, m, b) if np.ndim(m) == 1 else orc.apply_matrix(want, n, m, b)
    diff = dev.to_numpy().astype(np.complex128) - want
    assert np.linalg.norm(diff) <= len(gates) * 6e-7, np.linalg.norm(diff)
    assert abs(dev.norm2() - 1.0) < 2e-5


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_inplace_bit_permutation_matches_out_of_place(DS, dtype):
    """b2q_sv_permute_bits_inplace (tile passes in shared memory, no second buffer)
    is bit-exact with the index definition out[o] = in[i], bit k of o = bit
    src_bit[k] of i (np.moveaxis of linalg/transformations.py:743-754)."""
    rng = np.random.RandomState(21)
    for n in (1, 2, 5, 12, 13, 14, 18, 21):
        perms = [list(range(n))[::-1], list(range(1, n)) + [0], list(range(n))]
        perms += [rng.permutation(n).tolist() for _ in range(3)]
        for src in perms:
            state = rand_state(rng, n, dtype)
            dev = DS.from_numpy(state)
            passes = dev.permute_bits_inplace(src)
            o = np.arange(1 << n, dtype=np.int64)
            i = np.zeros_like(o)
            for k, sb in enumerate(src):
                i |= ((o >> k) & 1) << sb
            np.testing.assert_array_equal(dev.to_numpy(), state[i])
            assert passes <= 4 and (passes == 0) == (src == list(range(n)))
            np.testing.assert_array_equal(DS.from_numpy(state).permute_bits(src).to_numpy(), state[i])


def test_generic_kernel_large_k(DS):
    rng = np.random.RandomState(5)
    for dtype, ks in ((np.complex64, (6, 7)), (np.complex128, (5, 6))):
        for k in ks:
            n = 12
            targets = rng.permutation(n)[:k].tolist()
            state = rand_state(rng, n, dtype)
            m = rand_matrix(rng, k)
            dev = DS.from_numpy(state)
            dev.apply_matrix(m, targets)
            err = np.max(np.abs(dev.to_numpy() - orc.apply_matrix(state, n, m, targets)))
            assert err <= ATOL[np.dtype(dtype)] * 4


def test_golden_targeted_left_multiply(DS):
    g = load_golden('targeted_left_multiply.npz')
    for c in range(int(g['num_cases'])):
        n = int(g[f'c{c}_n'])
        state = g[f'c{c}_state']
        dev = DS.from_numpy(state)
        dev.apply_matrix(g[f'c{c}_matrix'], orc.axes_to_bits(n, g[f'c{c}_axes']))
        # golden matrices are unnormalised gaussians: scale the tolerance by ||M||
        scale = max(1.0, np.linalg.norm(g[f'c{c}_matrix'], 2))
        err = np.max(np.abs(dev.to_numpy() - g[f'c{c}_out']))
        assert err <= ATOL[state.dtype] * scale, (c, err)


def test_golden_simulator_final_states(DS):
    g = load_golden('simulator_final_states.npz')
    for c in range(int(g['num_cases'])):
        n = int(g[f'c{c}_n'])
        final = g[f'c{c}_final']
        gates = [
            (g[f'c{c}_g{i}_u'], orc.axes_to_bits(n, g[f'c{c}_g{i}_axes']))
            for i in range(int(g[f'c{c}_num_gates']))
        ]
        for batch in (False, True):
            dev = DS.basis(n, final.dtype, 0)
            if batch:
                dev.apply_batch(gates)
            else:
                for m, b in gates:
                    dev.apply_matrix(m, b)
            err = np.max(np.abs(dev.to_numpy() - final))
            assert err <= ATOL[final.dtype], (c, batch, err)
            assert abs(dev.norm2() - 1.0) < 1e-4


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_diagonal_scale_init(DS, dtype):
    rng = np.random.RandomState(3)
    for n in (2, 9, 14):
        state = rand_state(rng, n, dtype)
        for k in (1, 2, min(n, 5)):
            targets = rng.permutation(n)[:k].tolist()
            d = np.exp(1j * rng.standard_normal(1 << k))
            dev = DS.from_numpy(state)
            dev.apply_diagonal(d, targets)
            err = np.max(np.abs(dev.to_numpy() - orc.apply_diagonal(state, n, d, targets)))
            assert err <= ATOL[np.dtype(dtype)]
        dev = DS.from_numpy(state)
        dev.scale(0.3 - 0.4j)
        assert np.max(np.abs(dev.to_numpy() - state * (0.3 - 0.4j))) <= ATOL[np.dtype(dtype)]
        idx = int(rng.randint(1 << n))
        b = DS.basis(n, dtype, idx).to_numpy()
        assert b[idx] == 1 and np.count_nonzero(b) == 1


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_wide_diagonal_blocks(DS, dtype):
    """Diagonal blocks of up to 16 wires: shared-memory table kernel (<= 13 wires
    complex64 / 12 complex128) and the global-table kernel above, every position
    class (chunk-local bits, chunk-selecting bits, mixed), bit-exact table indexing."""
    rng = np.random.RandomState(31)
    for n in (1, 3, 12, 13, 14, 18, 22):
        state = rand_state(rng, n, dtype)
        widths = sorted({1, min(n, 2), min(n, 7), min(n, 12), min(n, 13), min(n, 14), min(n, 16)})
        for k in widths:
            picks = [rng.permutation(n)[:k].tolist(), list(range(k)), list(range(n - k, n))[::-1]]
            for targets in picks:
                d = np.exp(1j * rng.standard_normal(1 << k))
                dev = DS.from_numpy(state)
                dev.apply_diagonal(d, targets)
                err = np.max(np.abs(dev.to_numpy() - orc.apply_diagonal(state, n, d, targets)))
                assert err <= ATOL[np.dtype(dtype)], (n, k, targets, err)
    # index mapping is exact: a table of distinct integers reproduces the index bits
    n, targets = 15, [14, 0, 7, 3, 9, 1, 12]
    d = np.arange(1 << len(targets)).astype(np.complex128)
    dev = DS.from_numpy(np.ones(1 << n, dtype=dtype))
    dev.apply_diagonal(d, targets)
    idx = np.arange(1 << n)
    want = np.zeros(1 << n, dtype=np.int64)
    for t in targets:
        want = (want << 1) | ((idx >> t) & 1)
    np.testing.assert_array_equal(dev.to_numpy().real.astype(np.int64), want)
    # mixed batches: dense blocks and diagonal blocks in one apply_batch call
    n = 16
    state = rand_state(rng, n, dtype)
    gates = []
    for i in range(12):
        if i % 3 == 1:
            k = int(rng.randint(2, 13))
            gates.append((np.exp(1j * rng.standard_normal(1 << k)), rng.permutation(n)[:k].tolist()))
        else:
            k = int(rng.randint(1, 4))
            gates.append((rand_unitary(rng, k), rng.permutation(n)[:k].tolist()))
    dev = DS.from_numpy(state)
    dev.apply_batch(gates)
    want = state.astype(np.complex128)
    for m, w in gates:
        want = orc.apply_diagonal(want, n, m, w) if np.ndim(m) == 1 else orc.apply_matrix(want, n, m, w)
    assert np.max(np.abs(dev.to_numpy() - want)) <= 4 * ATOL[np.dtype(dtype)]


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_norm_gather_pauli(DS, dtype):
    rng = np.random.RandomState(11)
    for n in (1, 4, 10, 17):
        state = rand_state(rng, n, dtype) * 1.7
        dev = DS.from_numpy(state)
        assert abs(dev.norm2() - orc.norm2(state)) <= 1e-5 * orc.norm2(state)
        idx = rng.randint(0, 1 << n, size=33)
        np.testing.assert_array_equal(dev.amplitudes(idx), state[idx].astype(np.complex128))
        for _ in range(4):
            x = int(rng.randint(1 << n))
            z = int(rng.randint(1 << n))
            got = dev.pauli_expectation(x, z)
            want = orc.pauli_expectation(state, n, x, z)
            assert abs(got - want) <= (1e-5 if dtype == np.complex64 else 1e-11) * 3


def test_golden_pauli_expectation(DS):
    g = load_golden('pauli_expectation.npz')
    for c in range(int(g['num_cases'])):
        n = int(g[f'c{c}_n'])
        x = z = 0
        for axis, code in enumerate(g[f'c{c}_codes']):
            b = n - 1 - axis
            if code in (1, 2):
                x |= 1 << b
            if code in (2, 3):
                z |= 1 << b
        dev = DS.from_numpy(g[f'c{c}_state'])
        assert abs(dev.pauli_expectation(x, z) - complex(g[f'c{c}_value'])) < 1e-11


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_pauli_expectations_sharing_an_x_mask(DS, dtype):
    """b2q_sv_pauli_expectation_multi: strings with one X mask in one pass (16 per
    launch) equal the one-string kernel, the oracle, and the golden cases."""
    rng = np.random.RandomState(31)
    tol = 2e-6 if dtype == np.complex64 else 1e-12
    for n in (1, 4, 9, 14, 19):
        state = rand_state(rng, n, dtype)
        dev = DS.from_numpy(state)
        for x in (0, int(rng.randint(0, 1 << n)), (1 << n) - 1):
            for count in (1, 5, 16, 21):
                zs = [int(v) for v in rng.randint(0, 1 << n, size=count)]
                got = dev.pauli_expectations(x, zs)
                assert got.shape == (count,)
                for z, v in zip(zs, got):
                    assert abs(v - orc.pauli_expectation(state, n, x, z)) < tol
                    assert abs(v - dev.pauli_expectation(x, z)) < tol
    g = load_golden('pauli_expectation.npz')
    for c in range(int(g['num_cases'])):
        n = int(g[f'c{c}_n'])
        x = z = 0
        for axis, code in enumerate(g[f'c{c}_codes']):
            b = n - 1 - axis
            if code in (1, 2):
                x |= 1 << b
            if code in (2, 3):
                z |= 1 << b
        if g[f'c{c}_state'].dtype == dtype:
            vals = DS.from_numpy(g[f'c{c}_state']).pauli_expectations(x, [z, z])
            assert abs(vals[0] - complex(g[f'c{c}_value'])) < 1e-6 and vals[0] == vals[1]


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_marginal_probs(DS, dtype):
    rng = np.random.RandomState(21)
    for n in (1, 3, 6, 7, 10, 15):
        state = rand_state(rng, n, dtype)
        dev = DS.from_numpy(state)
        for m in sorted({1, min(2, n), min(5, n), n}):
            for _ in range(3):
                bits = rng.permutation(n)[:m].tolist()
                got = dev.marginal_probs(bits)
                want = orc.marginal_probs(state, n, bits)
                np.testing.assert_allclose(got, want, atol=1e-6 if dtype == np.complex64 else 1e-13)


def test_golden_sampling_and_measurement(DS):
    g = load_golden('sampling.npz')
    for c in range(int(g['num_cases'])):
        n = int(g[f'c{c}_n'])
        state = g[f'c{c}_state']
        bits = orc.axes_to_bits(n, g[f'c{c}_indices'])
        m = len(bits)
        dev = DS.from_numpy(state)
        got = dev.sample_bits(bits, g[f'c{c}_uniforms'])
        want = g[f'c{c}_bits']
        assert got.dtype == np.uint8 and got.shape == want.shape
        assert np.mean(np.any(got != want, axis=1)) <= 1 / 64 + 1e-9, c
        # measurement: same outcome as the reference for the same uniform draw
        probs = dev.marginal_probs_device(bits)
        pick = int(DS.cdf_sample_device(probs, np.array([g[f'c{c}_measure_uniform']])).cpu()[0])
        values = [(pick >> (m - 1 - q)) & 1 for q in range(m)]
        assert values == g[f'c{c}_measure_bits'].tolist()
        p = probs.cpu().numpy()
        dev.collapse(bits, values, p[pick] / p.sum())
        err = np.max(np.abs(dev.to_numpy() - g[f'c{c}_measure_state']))
        assert err <= ATOL[state.dtype] * 4


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_full_state_sampler_matches_oracle_and_chi_squared(DS, dtype):
    rng = np.random.RandomState(31)
    for n in (1, 5, 7, 8, 13, 16, 18):
        state = rand_state(rng, n, dtype)
        dev = DS.from_numpy(state)
        reps = 20000
        u = rng.random_sample(reps)
        got = dev.sample_indices(u)
        probs = orc.marginal_probs(state, n, list(range(n - 1, -1, -1)))
        want = orc.choice_indices(probs, u)
        mismatch = np.mean(got != want.astype(np.uint64))
        assert mismatch < 2e-3, (n, mismatch)
        # chi-squared on a coarse marginal (top min(n,6) bits)
        mbits = min(n, 6)
        hist = np.bincount((got >> np.uint64(n - mbits)).astype(np.int64), minlength=1 << mbits)
        expect = probs.reshape(1 << mbits, -1).sum(axis=1) * reps
        chi2 = np.sum((hist - expect) ** 2 / np.maximum(expect, 1e-12))
        dof = (1 << mbits) - 1
        assert chi2 < dof + 6 * np.sqrt(2 * dof) + 10, (n, chi2)


def test_sampler_never_returns_zero_probability_states(DS):
    for n in (3, 9, 15):
        for idx in (0, (1 << n) - 1, 5):
            dev = DS.basis(n, np.complex64, idx)
            u = np.concatenate([[0.0, 1.0 - 2**-53], np.random.RandomState(0).random_sample(100)])
            assert np.all(dev.sample_indices(u) == idx)
            got = dev.sample_bits(list(range(n - 1, -1, -1)), u[:5])
            want = [(idx >> (n - 1 - a)) & 1 for a in range(n)]
            assert np.all(got == np.array(want, dtype=np.uint8))


def test_reference_known_answer_vectors(DS):
    g = load_golden('reference_test_vectors.npz')
    for x in range(8):
        dev = DS.basis(3, np.complex64, x)
        got = dev.sample_bits(orc.axes_to_bits(3, [2, 1, 0]), np.array([0.5]))
        np.testing.assert_array_equal(got, g['big_endian_samples'][x])
    dev = DS.basis(3, np.complex64, 6)
    for perm, want in zip(g['perms'], g['perm_samples']):
        got = dev.sample_bits(orc.axes_to_bits(3, perm), np.array([0.3]))
        np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_density_matrix_golden_and_helpers(DS, dtype):
    g = load_golden('density_matrix_final_states.npz')
    for c in range(int(g['num_cases'])):
        final = g[f'c{c}_final']
        if final.dtype != np.dtype(dtype):
            continue
        n = int(g[f'c{c}_n'])
        dev = DS.basis(2 * n, dtype, 0)
        for i in range(int(g[f'c{c}_num_ops'])):
            dev.dm_apply_channel(
                list(g[f'c{c}_g{i}_kraus']), orc.axes_to_bits(n, g[f'c{c}_g{i}_axes'])
            )
        rho = dev.to_numpy()
        err = np.max(np.abs(rho.reshape(final.shape) - final))
        assert err <= ATOL[np.dtype(dtype)], (c, err)
        assert abs(dev.dm_trace() - 1.0) < 1e-5
        diag = dev.dm_diagonal_device().cpu().numpy()
        np.testing.assert_allclose(diag, orc.dm_diagonal(rho, n), atol=1e-7)
        p = orc.marginal_probs(np.sqrt(np.maximum(diag, 0)).astype(np.complex128), n, [0])
        if p[1] > 1e-3:
            dev.dm_collapse([0], [1], p[1] / p.sum())
            want = orc.dm_collapse(rho, n, [0], [1], p[1] / p.sum())
            assert np.max(np.abs(dev.to_numpy() - want)) <= ATOL[np.dtype(dtype)] * 10


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_layout_ops_kron_permute_argmax(DS, dtype):
    rng = np.random.RandomState(71)
    a = rand_state(rng, 3, dtype)
    b = rand_state(rng, 5, dtype)
    k = DS.from_numpy(a).kron(DS.from_numpy(b))
    np.testing.assert_allclose(k.to_numpy(), np.kron(a, b), atol=ATOL[np.dtype(dtype)])
    n = 9
    s = rand_state(rng, n, dtype)
    src = rng.permutation(n).tolist()
    got = DS.from_numpy(s).permute_bits(src).to_numpy()
    o = np.arange(1 << n)
    i = np.zeros_like(o)
    for kk, sb in enumerate(src):
        i |= ((o >> kk) & 1) << sb
    np.testing.assert_array_equal(got, s[i])
    assert DS.from_numpy(s).argmax_abs() == int(np.argmax(np.abs(s.astype(np.complex128)) ** 2))
    t = DS.from_numpy(np.kron(a, b))
    assert t.kron_allclose(DS.from_numpy(a), DS.from_numpy(b), 1e-6)
    bad = b.copy()
    bad[3] += 0.01
    assert not t.kron_allclose(DS.from_numpy(a), DS.from_numpy(bad), 1e-6)
    sl = DS.from_numpy(s).slice_copy(64, 5).to_numpy()
    np.testing.assert_array_equal(sl, s[64:96])


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_allclose_and_kron_sizes_and_positions(DS, dtype):
    """b2q_sv_allclose / b2q_sv_kron_allclose / b2q_sv_kron with 16-byte accesses and
    four accesses in flight: every size class (scalar, one vector, unroll tails), a
    violation at the first, an odd and the last index, and np.allclose's threshold
    (|a - b| <= atol + rtol |b|; reference use: sim/simulation_product_state.py via
    qis/states.py validate / linalg.allclose_up_to_global_phase call sites)."""
    rng = np.random.RandomState(72)
    for n in (1, 2, 3, 5, 10, 13, 17):
        s = rand_state(rng, n, dtype)
        dev = DS.from_numpy(s)
        assert dev.allclose(DS.from_numpy(s.copy()), 1e-7)
        for pos in sorted({0, (1 << n) // 2 | 1, (1 << n) - 1}):
            bad = s.copy()
            bad[pos] += 3e-3
            assert not dev.allclose(DS.from_numpy(bad), 1e-3, rtol=0.0)
            assert dev.allclose(DS.from_numpy(bad), 1e-2, rtol=0.0)
            # the relative term: |bad[pos]| <= 1.1, so rtol = 1 admits nothing below 3e-3 only if
            # |bad[pos]| < 3e-3; a huge rtol admits everything non-zero
            assert dev.allclose(DS.from_numpy(bad), 1e-9, rtol=1e3) == bool(
                np.all(np.abs(s.astype(np.complex128) - bad) <= 1e-9 + 1e3 * np.abs(bad.astype(np.complex128))))
    for na, nb in ((1, 1), (3, 1), (1, 4), (4, 7), (9, 2), (2, 12), (8, 8)):
        a = rand_state(rng, na, dtype)
        b = rand_state(rng, nb, dtype)
        k = DS.from_numpy(a).kron(DS.from_numpy(b))
        want = np.kron(a, b)
        np.testing.assert_allclose(k.to_numpy(), want, atol=ATOL[np.dtype(dtype)])
        t = DS.from_numpy(want.astype(dtype))
        assert t.kron_allclose(DS.from_numpy(a), DS.from_numpy(b), 1e-6)
        for pos in sorted({0, (1 << (na + nb)) - 1, (1 << (na + nb)) // 3 | 1}):
            wrong = want.astype(dtype).copy()
            wrong[pos] += 5e-3
            assert not DS.from_numpy(wrong).kron_allclose(DS.from_numpy(a), DS.from_numpy(b), 1e-3, rtol=0.0)
            assert DS.from_numpy(wrong).kron_allclose(DS.from_numpy(a), DS.from_numpy(b), 1e-2, rtol=0.0)


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_reduced_density_matrix_gram_kernel_positions(DS, dtype):
    """The one-read Gram kernel (3-5 kept bits, >= 11 qubits) for kept bits at the
    bottom, at the top, straddling the tile, and in the caller's order; against
    qis/states.py:676-693 restated, and against the row-tile kernel's size class."""
    rng = np.random.default_rng(15)
    tol = 2e-6 if dtype == np.complex64 else 1e-13
    for n in (11, 12, 16, 22):
        psi = rand_state(rng, n, dtype)
        dev = DS.from_numpy(psi, dtype)
        cases = [[0, 1, 2], [2, 1, 0, 3], [4, 3, 2, 1, 0], [n - 1, n - 2, n - 3], [n - 3, n - 1, n - 2, n - 5, n - 4],
                 [0, n - 1, 5], [n - 1, 0, 6, 3], [9, 0, n - 1, 7, 3], [8, 9, 10, 7]]
        for bits in cases:
            got = dev.reduced_density_matrix(bits)
            want = orc.reduced_density_matrix(psi, n, bits)
            np.testing.assert_allclose(got, want, atol=tol, rtol=0)
            np.testing.assert_allclose(got, got.conj().T, atol=tol, rtol=0)  # (a,b) and (b,a) are summed by different threads
        again = dev.reduced_density_matrix(cases[-2])
        np.testing.assert_array_equal(again, dev.reduced_density_matrix(cases[-2]))  # fixed summation order


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_pauli_expectations_on_states_that_walk_the_sign_table(DS, dtype):
    """b2q_sv_pauli_expectation_multi above 2^21 amplitudes: a thread walks several runs
    and the signs of the index bits above the virtual-thread bits come from the
    per-CTA table (n = 22: two steps per thread).  The oracle (seconds per string at this
    size) checks three strings per X mask, the one-string kernel all of them."""
    rng = np.random.RandomState(33)
    tol = 2e-6 if dtype == np.complex64 else 1e-12
    for n in (12, 22):
        state = rand_state(rng, n, dtype)
        dev = DS.from_numpy(state)
        top = (1 << (n - 1)) | (1 << (n - 2))
        for x in (0, (1 << (n - 1)) | int(rng.randint(0, 1 << (n - 1))) | 5):
            zs = [int(v) for v in rng.randint(0, 1 << n, size=15)] + [top, (1 << n) - 1, 0, 7, top | 7]
            got = dev.pauli_expectations(x, zs)
            for i in (0, 15, 19):
                assert abs(got[i] - orc.pauli_expectation(state, n, x, zs[i])) < tol, (n, x, zs[i])
            for z, v in zip(zs, got):
                assert abs(v - dev.pauli_expectation(x, z)) < tol, (n, x, z)


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_dist_pack_unpack(DS, dtype):
    import torch

    rng = np.random.RandomState(41)
    n = 12
    state = rand_state(rng, n, dtype)
    dev = DS.from_numpy(state)
    for bits in ([11], [10, 11], [0, 5], [7, 3, 9]):
        packed = torch.empty_like(dev.tensor)
        dev.dist_pack(bits, packed)
        got = packed.cpu().numpy().view(np.dtype(dtype)).reshape(-1)
        np.testing.assert_array_equal(got, orc.dist_pack(state, n, bits))
        dev2 = DS(n, dtype)
        dev2.dist_unpack(bits, packed)
        np.testing.assert_array_equal(dev2.to_numpy(), state)


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_dm_pauli_expectation(DS, dtype):
    """b2q_dm_pauli_expectation: reference goldens (ops/pauli_string.py:657-770)
    and random non-Hermitian arrays against the oracle."""
    g = load_golden('dm_pauli_expectation.npz')
    tol = 1e-5 if dtype == np.complex64 else 1e-12
    for c in range(int(g['num_cases'])):
        n = int(g[f'c{c}_n'])
        x = z = 0
        for axis, code in enumerate(g[f'c{c}_codes']):
            b = n - 1 - axis
            if code in (1, 2):
                x |= 1 << b
            if code in (2, 3):
                z |= 1 << b
        dev = DS.from_numpy(g[f'c{c}_rho'].reshape(-1), dtype)
        assert abs(dev.dm_pauli_expectation(x, z) - float(g[f'c{c}_value'])) < tol
    rng = np.random.default_rng(8)
    for n in (3, 7, 10):
        rho = (rng.standard_normal(1 << 2 * n) + 1j * rng.standard_normal(1 << 2 * n)) / (1 << n)
        dev = DS.from_numpy(rho, dtype)
        for _ in range(5):
            x, z = int(rng.integers(0, 1 << n)), int(rng.integers(0, 1 << n))
            want = orc.dm_pauli_expectation(rho.astype(dtype), n, x, z)
            assert abs(dev.dm_pauli_expectation(x, z) - want) < tol * 10


def test_golden_reduced_density_matrix_and_trajectory_ops(DS):
    """CUDA path against the committed reference outputs (tests/golden/make_golden.py)."""
    g = load_golden('reduced_density_matrix.npz')
    for c in range(int(g['num_cases'])):
        n = int(g[f'c{c}_n'])
        bits = [n - 1 - int(a) for a in g[f'c{c}_indices']]
        dev = DS.from_numpy(g[f'c{c}_state'], np.complex128)
        np.testing.assert_allclose(dev.reduced_density_matrix(bits), g[f'c{c}_rho'], atol=1e-12)
    g = load_golden('trajectory_ops.npz')
    for c in range(int(g['num_cases'])):
        n = int(g[f'c{c}_n'])
        states = g[f'c{c}_states']
        B = states.shape[0]
        bits = [n - 1 - int(a) for a in g[f'c{c}_axes']]
        kraus = g[f'c{c}_kraus']
        dev = DS.from_numpy(states.reshape(-1), np.complex128)
        w = dev.bsv_kraus_weights(n, kraus, bits)
        np.testing.assert_allclose(w, g[f'c{c}_weights'], atol=1e-12)
        for i in range(len(kraus)):
            dev = DS.from_numpy(states.reshape(-1), np.complex128)
            dev.bsv_apply_select(n, kraus, bits, np.full(B, i))
            np.testing.assert_allclose(dev.to_numpy().reshape(B, -1), g[f'c{c}_applied'][:, i], atol=1e-12)
        results = g[f'c{c}_results']
        probs = np.array([
            orc.marginal_probs(states[t], n, bits)[int(''.join(str(int(b)) for b in results[t]), 2)]
            for t in range(B)
        ])
        dev = DS.from_numpy(states.reshape(-1), np.complex128)
        dev.bsv_collapse(n, bits, results, 1.0 / np.sqrt(probs))
        np.testing.assert_allclose(dev.to_numpy().reshape(B, -1), g[f'c{c}_collapsed'].reshape(B, -1), atol=1e-12)


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_reduced_density_matrix_matches_oracle(DS, dtype):
    """b2q_sv_reduced_density_matrix vs qis/states.py:676-693 restated."""
    rng = np.random.default_rng(5)
    for n in (1, 2, 5, 9, 14, 21):
        psi = rand_state(rng, n, dtype)
        dev = DS.from_numpy(psi, dtype)
        for m in range(1, min(n, 5) + 1):
            for _ in range(3):
                bits = rng.permutation(n)[:m].tolist()
                got = dev.reduced_density_matrix(bits)
                want = orc.reduced_density_matrix(psi, n, bits)
                np.testing.assert_allclose(got, want, atol=2e-6 if dtype == np.complex64 else 1e-13, rtol=0)
                assert abs(np.trace(got) - 1) < (1e-5 if dtype == np.complex64 else 1e-12)


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_batched_trajectory_kernels_match_oracle(DS, dtype):
    """b2q_bsv_apply_select / b2q_bsv_kraus_weights / b2q_bsv_collapse against the
    per-trajectory oracle (reference: sim/state_vector_simulation_state.py:183-257,
    sim/state_vector.py:300-318)."""
    rng = np.random.default_rng(77)
    atol = ATOL[np.dtype(dtype)]
    for n, b in ((1, 3), (3, 0), (4, 5), (7, 6), (11, 4), (2, 13)):
        B = 1 << b
        states = np.concatenate([rand_state(rng, n, dtype) for _ in range(B)])
        for k in (1, 2, 3):
            if k > n:
                continue
            for count in (1, 2, 4, 5):
                targets = rng.permutation(n)[:k].tolist()
                mats = np.stack([rand_matrix(rng, k) for _ in range(count)])
                mats[0] = np.eye(1 << k)
                choice = rng.integers(0, count, size=B)
                scale = rng.uniform(0.5, 2.0, size=B)
                for skip, sc in ((-1, None), (0, None), (-1, scale), (0, scale)):
                    dev = DS.from_numpy(states, dtype)
                    dev.bsv_apply_select(n, mats, targets, choice, sc, skip)
                    want = orc.bsv_apply_select(states, n, mats, targets, choice, sc, skip)
                    np.testing.assert_allclose(dev.to_numpy(), want, atol=4 * atol, rtol=0)
                dev = DS.from_numpy(states, dtype)
                got = dev.bsv_kraus_weights(n, mats, targets)
                want = orc.bsv_kraus_weights(states, n, mats, targets)
                np.testing.assert_allclose(got, want, rtol=2e-5 if dtype == np.complex64 else 1e-12, atol=atol)
                np.testing.assert_array_equal(dev.to_numpy(), states)
        # a layer of 1-qubit selections in one launch (targets may repeat)
        for count, layer in ((4, min(n, 5)), (3, n + 2), (2, 32)):
            mats = np.stack([rand_matrix(rng, 1) for _ in range(count)])
            mats[0] = np.eye(2)
            tg = rng.integers(0, n, size=layer).tolist()
            choices = rng.integers(0, count, size=(layer, B))
            choices[rng.random((layer, B)) < 0.6] = 0
            for skip in (0, -1):
                dev = DS.from_numpy(states, dtype)
                dev.bsv_apply_select_multi(n, mats, tg, choices, skip)
                want = states
                for j, bit in enumerate(tg):
                    want = orc.bsv_apply_select(want, n, mats, [bit], choices[j], None, skip)
                scale_up = max(1.0, float(np.max(np.abs(want))) * 4)
                np.testing.assert_allclose(dev.to_numpy(), want, atol=scale_up * 4 * atol, rtol=0)
        m = int(rng.integers(1, n + 1))
        bits = rng.permutation(n)[:m].tolist()
        values = rng.integers(0, 2, size=(B, m))
        scale = rng.uniform(0.5, 2.0, size=B)
        dev = DS.from_numpy(states, dtype)
        dev.bsv_collapse(n, bits, values, scale)
        want = orc.bsv_collapse(states, n, bits, values, scale)
        np.testing.assert_allclose(dev.to_numpy(), want, atol=4 * atol, rtol=0)
        # exact zeros outside the selected slice (bit-exact index work)
        assert np.array_equal(dev.to_numpy() == 0, want == 0)


def test_full_size_properties_30q(DS):
    """BASELINE config sizes: size-independent properties at 30 qubits c64
    (8.6 GB state): unitarity round trip, norm, basis-state sampling."""
    n = 30
    rng = np.random.RandomState(2024)
    dev = DS.basis(n, np.complex64, 0)
    gates = []
    for layer in range(3):
        for k in (1, 2, 3, 4):
            targets = rng.permutation(n)[:k].tolist()
            gates.append((rand_unitary(rng, k), targets))
        gates.append((rand_unitary(rng, 2), [0, 29]))
        gates.append((rand_unitary(rng, 2), [3, 1]))
    dev.apply_batch(gates)
    assert abs(dev.norm2() - 1.0) < 1e-4
    inv = [(m.conj().T, t) for m, t in reversed(gates)]
    dev.apply_batch(inv)
    amp0 = dev.amplitudes([0, 1, 12345])
    assert abs(amp0[0] - 1.0) < 1e-4 and abs(amp0[1]) < 1e-5 and abs(amp0[2]) < 1e-5
    assert abs(dev.norm2() - 1.0) < 1e-4
    # Hadamard on every qubit -> uniform superposition: analytic amplitude 2^-15
    h = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    dev.apply_batch([(h, [q]) for q in range(n)])
    amps = dev.amplitudes(rng.randint(0, 1 << n, size=64))
    np.testing.assert_allclose(amps, 2.0**-15, atol=1e-9)
    idx = dev.sample_indices(rng.random_sample(4096))
    # uniform distribution: top 4 bits chi-squared
    hist = np.bincount((idx >> np.uint64(26)).astype(np.int64), minlength=16)
    chi2 = np.sum((hist - 256.0) ** 2 / 256.0)
    assert chi2 < 15 + 6 * np.sqrt(30) + 10
    # a non-uniform 30-qubit state: 1M samples against the state's own exact
    # 12-qubit marginal (device reduction), the check BASELINE asks for at 30 q
    dev.apply_batch(gates)
    bits = [29, 3, 17, 0, 22, 9, 11, 28, 5, 14, 1, 20]
    probs = dev.marginal_probs(bits)
    probs = probs / probs.sum()
    reps = 1_000_000
    samples = dev.sample_bits(bits, rng.random_sample(reps))
    keys = samples.astype(np.int64) @ (1 << np.arange(len(bits) - 1, -1, -1))
    hist = np.bincount(keys, minlength=1 << len(bits))
    mask = probs * reps > 5
    chi2 = np.sum((hist[mask] - probs[mask] * reps) ** 2 / (probs[mask] * reps))
    dof = int(mask.sum()) - 1
    assert chi2 < dof + 6 * np.sqrt(2 * dof) + 10, (chi2, dof)
    full = dev.sample_indices(rng.random_sample(reps))
    keys = np.zeros(reps, dtype=np.int64)
    for b in bits:
        keys = (keys << 1) | ((full >> np.uint64(b)) & np.uint64(1)).astype(np.int64)
    hist = np.bincount(keys, minlength=1 << len(bits))
    chi2 = np.sum((hist[mask] - probs[mask] * reps) ** 2 / (probs[mask] * reps))
    assert chi2 < dof + 6 * np.sqrt(2 * dof) + 10, (chi2, dof)
