"""Host-side checks of the gate kernel's planning logic (no GPU needed).

`b2q_debug_plan` exposes the plan the C++ host code derives for the
register-tiled kernel (which register bit holds which index bit, which lane
bits are exchanged by shuffles).  A lane-by-lane numpy emulation of the kernel
driven by that plan must reproduce the oracle, which checks the plan, the
shuffle-exchange scheme and the matrix permutation before any GPU time is
spent."""
import ctypes
import itertools

import numpy as np
import pytest

from cirq_b200 import _lib
from oracle import sv_oracle as orc


def get_plan(dtype_code, n, targets):
    lib = _lib.load()
    out = (ctypes.c_int * 24)()
    lib.b2q_debug_plan(dtype_code, n, _lib.int_array(targets), len(targets), out)
    o = list(out)
    return dict(
        feasible=bool(o[0]), S=o[1], GT=o[2], swaps=bool(o[3] & 1), vec=bool(o[3] & 2), n_ins=o[4],
        ins_pos=o[5:5 + o[4]], reg_off_log2=o[11:17], swap_lane=o[17:23], log2_items=o[23],
    )


def permuted_matrix(matrix, targets):
    lib = _lib.load()
    k = len(targets)
    m = np.ascontiguousarray(matrix, dtype=np.complex128)
    out = np.empty_like(m)
    lib.b2q_debug_permute_matrix(m.ctypes.data, _lib.int_array(targets), k, out.ctypes.data)
    return out


def insert_zero_bits(x, positions):
    for p in positions:
        low = x & ((1 << p) - 1)
        x = ((x >> p) << (p + 1)) | low
    return x


def emulate_fast_kernel(state, n, matrix, targets, dtype_code):
    """Mirrors sv_apply_fast_kernel (cirq_b200/csrc/b2q_apply.cu) lane by lane."""
    k = len(targets)
    pl = get_plan(dtype_code, n, targets)
    assert pl['feasible']
    vec = pl['vec']
    assert not (vec and dtype_code != 0)
    zb = 6 if vec else 5
    S, GT = pl['S'], pl['GT']
    rb = S + k + GT
    nr = 1 << rb
    reg_off = [0 if l < 0 else (1 << l) for l in pl['reg_off_log2']]
    mat = permuted_matrix(matrix, targets)
    out = state.copy()
    vb = 1 if vec else 0
    for item in range(1 << pl['log2_items']):
        x = np.zeros((32, nr), dtype=np.complex128)
        addr = np.zeros((32, nr), dtype=np.int64)
        for lane in range(32):
            b = insert_zero_bits(((item << 5) | lane) << vb, pl['ins_pos'])
            for r in range(nr):
                off = 0
                if vec:
                    for bit in range(1, rb):
                        if (r >> bit) & 1:
                            off += reg_off[bit]
                    a = b + off + (r & 1)
                else:
                    for bit in range(rb):
                        if (r >> bit) & 1:
                            off += reg_off[bit]
                    a = b + off
                addr[lane, r] = a
                x[lane, r] = state[a]

        def swap_all(v):
            for slot in range(k):
                lb = pl['swap_lane'][slot]
                if lb < 0:
                    continue
                new = v.copy()
                for lane in range(32):
                    hi = (lane >> lb) & 1
                    partner = lane ^ (1 << lb)
                    for r in range(nr):
                        if (r >> slot) & 1:
                            continue
                        r0, r1 = r, r | (1 << slot)
                        # partner sends (hi_p ? x[r0] : x[r1])
                        hi_p = (partner >> lb) & 1
                        recv = v[partner, r0] if hi_p else v[partner, r1]
                        if hi:
                            new[lane, r0] = recv
                        else:
                            new[lane, r1] = recv
                v = new
            return v

        if pl['swaps']:
            assert S == 0
            x = swap_all(x)
        y = np.zeros_like(x)
        dim = 1 << k
        for lane in range(32):
            for g in range(1 << GT):
                for s in range(1 << S):
                    vin = np.array([x[lane, (((g << k) | c) << S) | s] for c in range(dim)])
                    vout = mat @ vin
                    for r in range(dim):
                        y[lane, (((g << k) | r) << S) | s] = vout[r]
        if pl['swaps']:
            y = swap_all(y)
        for lane in range(32):
            for r in range(nr):
                out[addr[lane, r]] = y[lane, r]
    return out


CASES_C64 = [
    [9], [0], [3], [5], [6],
    [7, 9], [0, 9], [0, 1], [2, 4], [1, 8], [5, 6], [9, 0], [8, 7], [4, 0],
    [7, 8, 9], [0, 1, 2], [0, 5, 9], [3, 9, 1], [2, 3, 4],
    [6, 7, 8, 9], [0, 1, 8, 9], [1, 2, 3, 4], [5, 0, 9, 3],
]


@pytest.mark.parametrize('targets', CASES_C64)
def test_emulated_kernel_matches_oracle_c64(targets):
    n = 11 if len(targets) <= 3 else 12
    rng = np.random.RandomState(hash(tuple(targets)) % (1 << 31))
    k = len(targets)
    state = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    matrix = rng.standard_normal((1 << k, 1 << k)) + 1j * rng.standard_normal((1 << k, 1 << k))
    got = emulate_fast_kernel(state, n, matrix, targets, 0)
    want = orc.apply_matrix(state, n, matrix, targets)
    np.testing.assert_allclose(got, want, atol=1e-12)


@pytest.mark.parametrize(
    'targets', [[9], [0], [4], [5], [0, 9], [0, 1], [3, 4], [4, 5], [0, 2, 4], [1, 6, 9], [0, 1, 2, 3], [4, 5, 6, 0]]
)
def test_emulated_kernel_matches_oracle_c128(targets):
    n = 11
    rng = np.random.RandomState(hash(tuple(targets)) % (1 << 31))
    k = len(targets)
    state = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    matrix = rng.standard_normal((1 << k, 1 << k)) + 1j * rng.standard_normal((1 << k, 1 << k))
    got = emulate_fast_kernel(state, n, matrix, targets, 1)
    want = orc.apply_matrix(state, n, matrix, targets)
    np.testing.assert_allclose(got, want, atol=1e-12)


def test_five_qubit_gate_all_in_zone_c64():
    targets = [1, 2, 3, 4, 5]
    n = 13
    rng = np.random.RandomState(3)
    state = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    matrix = rng.standard_normal((32, 32)) + 1j * rng.standard_normal((32, 32))
    got = emulate_fast_kernel(state, n, matrix, targets, 0)
    np.testing.assert_allclose(got, orc.apply_matrix(state, n, matrix, targets), atol=1e-11)


@pytest.fixture
def remap_mode():
    lib = _lib.load()
    lib.b2q_set_lane_mode(1)
    yield
    lib.b2q_set_lane_mode(2)


@pytest.fixture(params=[1, 2])
def vec_mode(request):
    lib = _lib.load()
    lib.b2q_set_vec_mode(request.param)
    yield request.param
    lib.b2q_set_vec_mode(0)


@pytest.mark.parametrize('targets', [[9], [0], [3], [0, 9], [1, 4], [0, 1, 2], [3, 9, 1], [6, 7, 8, 9], [0, 1, 8, 9], [1, 2, 3, 4], [5, 0, 9, 3], [1, 2, 3, 4, 5], [0, 6, 7, 8, 12]])
def test_emulated_kernel_both_access_widths_c64(targets, vec_mode):
    n = 13
    rng = np.random.RandomState(hash(tuple(targets)) % (1 << 31))
    k = len(targets)
    assert get_plan(0, n, targets)['vec'] == (vec_mode == 1)
    state = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    matrix = rng.standard_normal((1 << k, 1 << k)) + 1j * rng.standard_normal((1 << k, 1 << k))
    got = emulate_fast_kernel(state, n, matrix, targets, 0)
    np.testing.assert_allclose(got, orc.apply_matrix(state, n, matrix, targets), atol=1e-11)


@pytest.fixture
def shuffle_mode():
    lib = _lib.load()
    lib.b2q_set_lane_mode(0)
    yield
    lib.b2q_set_lane_mode(2)


@pytest.mark.parametrize('targets', CASES_C64 + [[1, 2, 3, 4, 5]])
def test_emulated_kernel_shuffle_mode_c64(targets, shuffle_mode):
    n = 11 if len(targets) <= 3 else 13
    rng = np.random.RandomState(hash(tuple(targets)) % (1 << 31))
    k = len(targets)
    state = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    matrix = rng.standard_normal((1 << k, 1 << k)) + 1j * rng.standard_normal((1 << k, 1 << k))
    got = emulate_fast_kernel(state, n, matrix, targets, 0)
    np.testing.assert_allclose(got, orc.apply_matrix(state, n, matrix, targets), atol=1e-11)


def test_auto_policy_exhaustive_small():
    """Default (per-target) policy: every target set of size <= 4 drawn from the
    low 8 bits plus one high bit, both dtypes."""
    import itertools

    rng = np.random.RandomState(123)
    n = 13
    pool = [0, 1, 2, 3, 4, 5, 6, 12]
    state = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    for code in (0, 1):
        for k in (1, 2, 3, 4):
            combos = list(itertools.combinations(pool, k))
            rng.shuffle(combos)
            for targets in combos[: 12 if k > 1 else 8]:
                targets = list(targets)
                rng.shuffle(targets)
                matrix = rng.standard_normal((1 << k, 1 << k)) + 1j * rng.standard_normal((1 << k, 1 << k))
                got = emulate_fast_kernel(state, n, matrix, targets, code)
                want = orc.apply_matrix(state, n, matrix, targets)
                np.testing.assert_allclose(got, want, atol=1e-11, err_msg=f'{code} {targets}')


@pytest.mark.parametrize('targets', CASES_C64 + [[1, 2, 3, 4, 5]])
def test_emulated_kernel_remap_mode_c64(targets, remap_mode):
    n = 11 if len(targets) <= 3 else 13
    rng = np.random.RandomState(hash(tuple(targets)) % (1 << 31))
    k = len(targets)
    pl = get_plan(0, n, targets)
    assert pl['feasible'] and not pl['swaps']
    state = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    matrix = rng.standard_normal((1 << k, 1 << k)) + 1j * rng.standard_normal((1 << k, 1 << k))
    got = emulate_fast_kernel(state, n, matrix, targets, 0)
    np.testing.assert_allclose(got, orc.apply_matrix(state, n, matrix, targets), atol=1e-11)


@pytest.mark.parametrize('targets', [[0], [4], [0, 1], [3, 4], [0, 2, 4], [0, 1, 2, 3], [4, 5, 6, 0]])
def test_emulated_kernel_remap_mode_c128(targets, remap_mode):
    n = 11
    rng = np.random.RandomState(hash(tuple(targets)) % (1 << 31))
    k = len(targets)
    assert not get_plan(1, n, targets)['swaps']
    state = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    matrix = rng.standard_normal((1 << k, 1 << k)) + 1j * rng.standard_normal((1 << k, 1 << k))
    got = emulate_fast_kernel(state, n, matrix, targets, 1)
    np.testing.assert_allclose(got, orc.apply_matrix(state, n, matrix, targets), atol=1e-12)


def test_plan_covers_every_amplitude_exactly_once():
    for dtype_code in (0, 1):
        for targets in ([3, 20], [0, 1, 2], [25, 26, 27, 28], [7]):
            pl = get_plan(dtype_code, 30, targets)
            assert pl['feasible']
            zb = 6 if pl['vec'] else 5
            rb = pl['S'] + len(targets) + pl['GT']
            assert pl['log2_items'] + zb + pl['n_ins'] == 30
            # register-resident bits are distinct and never the vector bit
            assert len(set(pl['ins_pos'])) == pl['n_ins'] and min(pl['ins_pos']) >= (1 if pl['vec'] else 0)
            assert pl['n_ins'] == sum(1 for l in pl['reg_off_log2'][:rb] if l >= 0)


def test_plan_infeasible_falls_back_for_tiny_states():
    assert not get_plan(0, 5, [0, 1])['feasible']
    assert not get_plan(1, 6, [0, 5])['feasible']
    assert not get_plan(0, 30, [0, 1, 2, 3, 4, 5])['feasible']  # k=6


def test_matrix_permutation_is_consistent_with_target_order():
    rng = np.random.RandomState(9)
    for targets in ([4, 1], [1, 4], [2, 0, 5], [5, 2, 0]):
        k = len(targets)
        m = rng.standard_normal((1 << k, 1 << k)) + 1j * rng.standard_normal((1 << k, 1 << k))
        pm = permuted_matrix(m, targets)
        # permuted matrix with ascending targets listed LSB-first == original in gate order
        srt = sorted(targets)
        n = 6
        state = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
        a = orc.apply_matrix(state, n, m, targets)
        b = orc.apply_matrix(state, n, pm, srt[::-1])
        np.testing.assert_allclose(a, b, atol=1e-12)
