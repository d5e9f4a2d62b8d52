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


# ---- staged tensor-core kernel: shared-memory layout ---------------------------
# b2q_debug_tc_stage_plan layout (int64[86]): rpos[10] | p5 | gpos[5] | swz_src[3] |
# swz_dst[3] | goff[16] | sreq[16] | smem_j[32].  The emulation replays both access
# phases of sv_apply_tc_staged_kernel for one warp region: every element must land
# where the group/member gather expects it, and no phase may have bank conflicts.

def _stage_plan(lib, n, targets):
    k = len(targets)
    arr = (ctypes.c_int * k)(*targets)
    out = (ctypes.c_int64 * 86)()
    _lib.check(lib.b2q_debug_tc_stage_plan(n, arr, k, out))
    o = list(out)
    return dict(rpos=o[0:k + 5], p5=o[10], gpos=o[11:16], src=o[16:19], dst=o[19:22],
                goff=o[22:22 + (1 << (k - 1))], sreq=o[38:38 + (1 << (k - 1))], smem_j=o[54:54 + (1 << k)])

def _stage_swz(x, p):
    for s, d in zip(p['src'], p['dst']):
        x ^= ((x >> s) & 1) << d
    return x

def _stage_insert_zero_bits(x, pos):
    for q in pos:
        low = x & ((1 << q) - 1)
        x = ((x >> q) << (q + 1)) | low
    return x

def _stage_check(lib, n, targets):
    k = len(targets)
    p = _stage_plan(lib, n, targets)
    region = 1 << (k + 5)
    free = [b for b in range(n) if b not in targets]
    for wt in (0, 1, 5, (1 << (n - k - 5)) - 1):
        base = _stage_insert_zero_bits(wt, p['rpos'])
        smem = {}
        # write phase
        for r in range(1 << (k - 1)):
            banks_q = {}
            for lane in range(32):
                goff = ((lane & 15) << 1) + ((lane >> 4) << p['p5'])
                s = _stage_swz(lane << 1, p) ^ p['sreq'][r]
                assert s % 2 == 0
                for e in (0, 1):
                    assert (s + e) not in smem
                    smem[s + e] = base + goff + p['goff'][r] + e
                q = lane // 8
                banks_q.setdefault(q, set()).add((s >> 1) & 7)
            for q, bs in banks_q.items():
                assert len(bs) == 8, ('write conflict', targets, r, q)
        assert sorted(smem) == list(range(region))
        # all region elements distinct and the right set
        want = set()
        for g in range(32):
            for j in range(1 << k):
                idx = base
                for i in range(5):
                    if (g >> i) & 1: idx += 1 << free[i]
                for b in range(k):
                    if (j >> b) & 1: idx += 1 << targets[b]
                want.add(idx)
        assert set(smem.values()) == want
        # read phase
        vec = targets[0] == 0
        for j in range(0, 1 << k, 2 if vec else 1):
            groups = {}
            for lane in range(32):
                gl = 0
                for i in range(5):
                    if (lane >> i) & 1: gl |= 1 << p['gpos'][i]
                s = _stage_swz(gl, p) ^ p['smem_j'][j]
                idx = base
                for i in range(5):
                    if (lane >> i) & 1: idx += 1 << free[i]
                for b in range(k):
                    if (j >> b) & 1: idx += 1 << targets[b]
                assert smem[s] == idx, (targets, lane, j)
                if vec:
                    assert s % 2 == 0 and smem[s + 1] == idx + 1
                    groups.setdefault(lane // 8, set()).add((s >> 1) & 7)
                else:
                    groups.setdefault(lane // 16, set()).add(s & 15)
            for q, bs in groups.items():
                assert len(bs) == (8 if vec else 16), ('read conflict', targets, j, q, bs)


def test_tc_staged_layout_is_a_conflict_free_bijection():
    import ctypes  # noqa: F401

    lib = _lib.load()
    rng = np.random.RandomState(0)
    for k in (4, 5):
        for n in (k + 7, 14, 20, 30):
            sets = [list(range(k)), list(range(1, k + 1)), list(range(n - k, n)),
                    [0] + list(range(n - k + 1, n)), [1] + list(range(n - k + 1, n)),
                    [0, 1] + list(range(n - k + 2, n))]
            sets += [sorted(rng.permutation(n)[:k].tolist()) for _ in range(20)]
            sets += [sorted(rng.permutation(min(n, 9))[:k].tolist()) for _ in range(20)]
            for t in sets:
                _stage_check(lib, n, t)


# ---- C-ABI surface --------------------------------------------------------------

def test_library_exports_every_symbol_the_header_declares():
    """include/cirq_b200.h is the drop-in boundary: each declared entry point
    must be exported by libcirq_b200.so and bound (with a signature) by the
    ctypes loader.  No compute call is made, so this runs without a GPU."""
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, 'include', 'cirq_b200.h')) as f:
        text = re.sub(r'/\*.*?\*/', '', f.read(), flags=re.S)
    declared = set(re.findall(r'\b(b2q_[a-z0-9_]+)\s*\(', text))
    assert len(declared) >= 40
    lib = _lib.load()
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, f'declared in the header but not exported: {missing}'
    bound = set(_lib.SIGNATURES) | set(_lib.DEBUG_SIGNATURES)
    assert declared <= bound, f'no ctypes signature for: {sorted(declared - bound)}'
    assert bound <= declared, f'bound but not declared in the header: {sorted(bound - declared)}'
    assert lib.b2q_version() > 0
    assert isinstance(lib.b2q_last_error(), bytes)


def test_read_out_entry_points_plan_without_crashing_at_every_size():
    """The host half of the read-out entry points (grid and table planning: virtual
    threads of the multi-string Pauli kernel, the reduced-density-matrix tile bits,
    vector widths of kron / allclose) runs before any launch.  On a machine without a
    device the calls must come back with an error code at every register size — a
    division by the zero runs of a 1-qubit state once took the process down."""
    import ctypes

    import torch

    if torch.cuda.is_available():
        pytest.skip('the device tests cover these calls where a GPU exists')
    lib = _lib.load()
    buf = np.zeros(1 << 12, dtype=np.complex128)
    ptr = ctypes.c_void_p(buf.ctypes.data)
    out = (ctypes.c_double * 4096)()
    zs = (ctypes.c_uint64 * 3)(1, 0, 1)
    ok = ctypes.c_int()
    for code in (_lib.C64, _lib.C128):
        for n in list(range(1, 24)) + [28, 34]:
            for x in (0, 1):
                assert lib.b2q_sv_pauli_expectation_multi(ptr, code, n, x, zs, 3, out, None) != 0
            for m in range(1, min(n, 5) + 1):
                bits = (ctypes.c_int * 5)(*([n - 1, 0, 2, 1, 3][:m] + [0] * (5 - m)))
                if len(set(bits[:m])) == m and max(bits[:m]) < n:
                    assert lib.b2q_sv_reduced_density_matrix(ptr, code, n, bits, m, out, None) != 0
            assert lib.b2q_sv_allclose(ptr, ptr, code, n, 1e-6, 1e-5, ctypes.byref(ok), None) != 0
        for na, nb in ((0, 0), (0, 1), (1, 0), (3, 2), (12, 1)):
            assert lib.b2q_sv_kron_allclose(ptr, na, ptr, nb, ptr, code, 1e-6, 1e-5, ctypes.byref(ok), None) != 0
            assert lib.b2q_sv_kron(ptr, na, ptr, nb, code, ptr, None) != 0
    assert lib.b2q_last_error()


# ---- read-out kernels: replay of their host plans ----------------------------------------

def _insert_zero_bits(x, positions):
    for pos in positions:
        x = ((x >> pos) << (pos + 1)) | (x & ((1 << pos) - 1))
    return x


@pytest.mark.parametrize('m', [3, 4, 5])
def test_reduced_density_matrix_tile_plan_replays_to_the_oracle(m):
    """b2q_debug_rdm_plan = the tile plan sv_reduced_dm_gram_kernel runs on (the same
    functions compute the offsets and slots on the device).  Replayed here: the tiles
    cover every amplitude once, a tile's slots are a bijection onto X[rest][a], the
    per-thread / per-load split of an element index is an OR of its parts, and the Gram
    sums of the staged tiles are the reference's reduced density matrix
    (qis/states.py:676-693 via the oracle)."""
    import ctypes

    lib = _lib.load()
    rng = np.random.default_rng(40 + m)
    for n in (11, 12, 14):
        psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
        psi /= np.linalg.norm(psi)
        cases = [rng.permutation(n)[:m].tolist() for _ in range(4)]
        cases += [list(range(m)), list(range(n - 1, n - 1 - m, -1)), [n - 1, 0, 5, 9, 2][:m]]
        for bits in cases:
            out = (ctypes.c_int64 * (16 + 2 * 2048))()
            assert lib.b2q_debug_rdm_plan(n, (ctypes.c_int * m)(*bits), m, out) == 0, lib.b2q_last_error()
            plan = np.array(out[:], dtype=np.int64)
            tile_pos, kept_rank = plan[:11].tolist(), plan[11:11 + m].tolist()
            offs, slots = plan[16::2], plan[17::2]
            assert tile_pos == sorted(set(tile_pos)) and max(tile_pos) < n
            assert set(sorted(bits)) <= set(tile_pos)
            free = [b for b in tile_pos if b not in bits]
            assert free == [b for b in range(n) if b not in bits][:11 - m]  # the LOWEST free bits
            assert [tile_pos[r] for r in kept_rank] == sorted(bits)
            assert sorted(slots.tolist()) == list(range(2048))
            for t in (0, 77, 255):
                for k in range(8):  # thread part | load part
                    assert offs[t | (k << 8)] == offs[t] | offs[k << 8]
                    assert slots[t | (k << 8)] == slots[t] | slots[k << 8]
            d = 1 << m
            rho = np.zeros((d, d), dtype=np.complex128)
            seen = np.zeros(1 << n, dtype=np.int64)
            for tile in range(1 << (n - 11)):
                base = _insert_zero_bits(tile, tile_pos)
                staged = np.zeros(2048, dtype=np.complex128)
                staged[slots] = psi[base + offs]
                seen[base + offs] += 1
                x = staged.reshape(2048 // d, d)
                rho += x.T @ x.conj()
            assert np.all(seen == 1)
            srt = sorted(bits)
            to_kernel = [sum(((idx >> (m - 1 - q)) & 1) << srt.index(bits[q]) for q in range(m)) for idx in range(d)]
            got = rho[np.ix_(to_kernel, to_kernel)]
            np.testing.assert_allclose(got, orc.reduced_density_matrix(psi, n, bits), atol=1e-13, rtol=0)
    assert lib.b2q_debug_rdm_plan(10, (ctypes.c_int * 3)(0, 1, 2), 3, out) != 0
    assert lib.b2q_debug_rdm_plan(12, (ctypes.c_int * 3)(0, 1, 1), 3, out) != 0


def test_pauli_run_plan_sign_decomposition_matches_the_oracle():
    """b2q_debug_pauli_plan = the launch shape of sv_pauli_multi_run_kernel.  Replayed with
    numpy: run r = k * vt + g, Walsh-Hadamard transform of the run's 8 pair products,
    W[z & 7], the sign of g's bits applied per walk and the sign of k's bits from the
    table — equal to ops/pauli_string.py:625-655 restated (the oracle), at sizes with
    one and with several steps per virtual thread."""
    import ctypes

    lib = _lib.load()
    out = (ctypes.c_int64 * 4)()
    shapes = {}
    for n in range(1, 35):
        assert lib.b2q_debug_pauli_plan(n, out) == 0
        by_runs, vt, k_count, blocks = out[:]
        shapes[n] = (by_runs, vt, k_count, blocks)
        if n < 12:
            assert not by_runs
            continue
        runs = (1 << n) >> 3
        assert by_runs and vt * k_count == runs and vt & (vt - 1) == 0 and vt >= 256
        assert 1 <= k_count <= 128 and 1 <= blocks <= min(vt >> 8, 148 * 4)
    assert shapes[21][2] == 1 and shapes[22][2] == 2 and shapes[28][2] == 128 and shapes[34][2] == 128
    had = np.array([[(-1) ** bin(j & w).count('1') for j in range(8)] for w in range(8)], dtype=np.float64)
    parity = lambda v: np.array([bin(int(x)).count('1') & 1 for x in v])  # noqa: E731
    rng = np.random.default_rng(9)
    for n, (vt, k_count) in ((12, (shapes[12][1], shapes[12][2])), (13, (256, 4)), (22, (shapes[22][1], shapes[22][2]))):
        psi = (rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)).astype(np.complex128)
        psi /= np.linalg.norm(psi)
        g = np.arange(vt, dtype=np.int64)
        for x in (0, (1 << (n - 1)) | 6):
            i = np.arange(1 << n, dtype=np.int64)
            prod = (np.conj(psi[i ^ x]) * psi).reshape(-1, 8)  # pair products by run
            w = prod @ had.T  # w[r, m] = sum_j (-1)^popcount(j & m) prod[r, j]
            zs = [int(v) for v in rng.integers(0, 1 << n, size=3)] + [(1 << n) - 1, 7, 1 << (n - 1)]
            for z in zs if n <= 13 else zs[1:4:2]:  # (the 2^22 case costs a second per string)
                col = w[:, z & 7].reshape(k_count, vt)  # [k][g]
                if n <= 13:
                    flip = parity((np.arange(k_count, dtype=np.int64) * vt << 3) & z)
                    fixed = parity((g << 3) & z)
                else:  # (vectorised parity for the big case)
                    kk = (np.arange(k_count, dtype=np.int64) * vt << 3) & z
                    gg = (g << 3) & z
                    flip = np.array([bin(int(v)).count('1') & 1 for v in kk])
                    gb = gg.copy()
                    fixed = np.zeros_like(gb)
                    while gb.any():
                        fixed ^= gb & 1
                        gb >>= 1
                walk = (np.where(flip[:, None] == 1, -col, col)).sum(axis=0)  # one virtual thread's walk
                value = np.where(fixed == 1, -walk, walk).sum()
                value *= 1j ** (bin(x & z).count('1') & 3)
                assert abs(value - orc.pauli_expectation(psi, n, x, z)) < 1e-11, (n, x, z)


# ---- tile kernel (two blocks per HBM pass): replay of its address tables ----------------

def _tile_slot(local, xmask):
    s = local & ~0xE
    for d in range(3):
        s |= (bin(local & xmask[d]).count('1') & 1) << (d + 1)
    return s


def _tile_plan(lib, n, blocks_sorted):
    import ctypes

    nb = len(blocks_sorted)
    flat = _lib.int_array([t for b in blocks_sorted for t in b])
    out = (ctypes.c_int64 * (13 + 3 + 8 + 8 + nb * 41))()
    rc = lib.b2q_debug_tile_plan(n, nb, flat, out)
    if rc != 0:
        return None
    v = list(out)
    plan = dict(tbits=v[0:13], xmask=v[13:16], rgoff=v[16:24], rslot=v[24:32], blocks=[])
    o = 32
    for _ in range(nb):
        plan['blocks'].append(dict(vec=v[o], gbit=v[o + 1:o + 9], mslot=v[o + 9:o + 41]))
        o += 41
    return plan


def _tile_emulate(lib, n, state, blocks):
    """Replays sv_apply_tc_tile_kernel's data movement on the CPU (512 threads: warp
    = q | h << 2 | g << 3; thread (g, q, lane) = amplitude group, h = half of its
    members): blocks = [(matrix 32x32 in the caller's target order, targets[5])].
    Returns the new state; asserts bijections and bank-conflict freedom on the way."""
    sorted_blocks = [sorted(t) for _, t in blocks]
    plan = _tile_plan(lib, n, sorted_blocks)
    assert plan is not None
    tbits, xmask = plan['tbits'], plan['xmask']
    assert tbits[:3] == [0, 1, 2] and tbits == sorted(tbits)
    mats = []
    for m, t in blocks:
        out = np.empty((32, 32), dtype=np.complex128)
        src = np.ascontiguousarray(m, dtype=np.complex128)
        assert lib.b2q_debug_permute_matrix(src.ctypes.data, _lib.int_array(t), 5, out.ctypes.data) == 0
        mats.append(out)
    state = state.copy()
    size = 1 << 13
    threads = np.arange(512)
    lane, warp = threads & 31, threads >> 5
    thr_local = (lane << 1) | (warp << 6)
    thr_goff = np.zeros(512, dtype=np.int64)
    for i in range(1, 10):
        thr_goff += ((thr_local >> i) & 1).astype(np.int64) << tbits[i]
    thr_slot = np.array([_tile_slot(int(x), xmask) for x in thr_local])
    # copy pattern: a bijection onto the slots, quarter-warps conflict free
    slots = (thr_slot[:, None] ^ np.array(plan['rslot'])[None, :])
    assert sorted(np.concatenate([slots.reshape(-1), slots.reshape(-1) + 1]).tolist()) == list(range(size))
    for r in range(8):
        for q0 in range(0, 512, 8):
            assert len({(int(s) >> 1) & 7 for s in slots[q0:q0 + 8, r]}) == 8
    goffs = thr_goff[:, None] + np.array(plan['rgoff'], dtype=np.int64)[None, :]
    q, h, g = warp & 3, (warp >> 2) & 1, warp >> 3
    gidx = lane | (q << 5) | (g << 7)
    for tile in range(1 << (n - 13)):
        base = tile
        for pos in tbits:
            base = ((base >> pos) << (pos + 1)) | (base & ((1 << pos) - 1))
        smem = np.zeros(size, dtype=state.dtype)
        smem[slots] = state[base + goffs]
        smem[slots + 1] = state[base + goffs + 1]
        for blk, m in zip(plan['blocks'], mats):
            gl = np.zeros(512, dtype=np.int64)
            for i in range(8):
                gl |= ((gidx >> i) & 1) << blk['gbit'][i]
            sg = np.array([_tile_slot(int(x), xmask) for x in gl])
            mslot = np.array(blk['mslot'])
            member = h[:, None] * 16 + np.arange(16)[None, :]  # [thread, 16]
            addr = sg[:, None] ^ mslot[member]
            if tile == 0:
                assert sorted(addr.reshape(-1).tolist()) == list(range(size))
                if blk['vec']:
                    assert np.all(addr[:, 1::2] == addr[:, 0::2] + 1) and np.all(addr[:, 0::2] % 2 == 0)
                    for j in range(0, 16, 2):
                        for q0 in range(0, 512, 8):
                            assert len({(int(s) >> 1) & 7 for s in addr[q0:q0 + 8, j]}) == 8
                else:
                    for j in range(16):
                        for h0 in range(0, 512, 16):
                            assert len({int(s) & 15 for s in addr[h0:h0 + 16, j]}) == 16
            # the two half-threads of a group together hold its 32 members
            order = np.argsort(gidx * 2 + h, kind='stable')
            full_addr = addr[order].reshape(256, 32)
            x = smem[full_addr]
            smem[full_addr] = x @ m.T
        state[base + goffs] = smem[slots]
        state[base + goffs + 1] = smem[slots + 1]
    return state


def test_tile_kernel_plan_replays_to_the_oracle():
    """Host emulation of the two-blocks-per-pass kernel against the oracle, over
    target sets that exercise every layout case (low bits, overlaps, widest union)."""
    lib = _lib.load()
    rng = np.random.RandomState(11)

    def unitary(k):
        d = 1 << k
        q, r = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
        return q * (np.diag(r) / np.abs(np.diag(r)))

    n = 15
    cases = [([0, 1, 2, 3, 4], [5, 6, 7, 8, 9]), ([9, 10, 11, 12, 13], [4, 5, 6, 7, 8]),
             ([13, 0, 5, 2, 9], [9, 2, 11, 3, 7]), ([2, 3, 4, 5, 6], [2, 3, 4, 5, 6]),
             ([1, 3, 5, 7, 9], [0, 2, 4, 6, 8]), ([13, 12, 11, 10, 9], [8, 7, 6, 5, 4]),
             ([4, 13, 8, 6, 11], [12, 5, 4, 10, 3])]
    for _ in range(25):
        a = rng.permutation(n)[:5].tolist()
        pool = [b for b in range(n)]
        b = rng.permutation(pool)[:5].tolist()
        cases.append((a, b))
    checked = 0
    for ta, tb in cases:
        assert len(set(ta) | set(tb) | {0, 1, 2}) <= 13  # always: 10 targets + 3
        state = (rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)) / 2 ** (n / 2)
        ma, mb = unitary(5), unitary(5)
        got = _tile_emulate(lib, n, state, [(ma, ta), (mb, tb)])
        want = orc.apply_matrix(orc.apply_matrix(state, n, ma, ta), n, mb, tb)
        np.testing.assert_allclose(got, want, atol=1e-12)
        checked += 1
    assert checked >= 20
    # a single block is a valid tile pass too
    state = (rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)) / 2 ** (n / 2)
    m = unitary(5)
    got = _tile_emulate(lib, n, state, [(m, [3, 12, 0, 7, 9])])
    np.testing.assert_allclose(got, orc.apply_matrix(state, n, m, [3, 12, 0, 7, 9]), atol=1e-12)


def test_tile_blocks_feasibility():
    lib = _lib.load()
    ten_high = [29, 28, 27, 26, 25, 24, 23, 22, 21, 20]
    # two 5-bit blocks always fit: 10 targets + index bits 0-2 = 13 tile bits
    assert lib.b2q_tile_blocks_feasible(_lib.C64, 30, 2, _lib.int_array([5, 5]), _lib.int_array(ten_high)) == 1
    assert lib.b2q_tile_blocks_feasible(_lib.C64, 30, 2, _lib.int_array([5, 5]),
                                        _lib.int_array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9])) == 1
    assert lib.b2q_tile_blocks_feasible(_lib.C128, 30, 2, _lib.int_array([5, 5]),
                                        _lib.int_array(list(range(2, 12)))) == 0
    assert lib.b2q_tile_blocks_feasible(_lib.C64, 12, 2, _lib.int_array([5, 5]),
                                        _lib.int_array(list(range(0, 10)))) == 0
    assert lib.b2q_tile_blocks_feasible(_lib.C64, 30, 2, _lib.int_array([3, 2]),
                                        _lib.int_array([4, 9, 1, 0, 17])) == 1
    assert lib.b2q_tile_blocks_feasible(_lib.C64, 30, 2, _lib.int_array([6, 2]),
                                        _lib.int_array([4, 9, 1, 0, 17, 3, 5, 6])) == 0


# ---- in-place bit permutation: the host planner's passes, replayed -----------------------

def _permute_reference(state, src_bit):
    n = len(src_bit)
    o = np.arange(1 << n, dtype=np.int64)
    i = np.zeros_like(o)
    for k, sb in enumerate(src_bit):
        i |= ((o >> k) & 1) << sb
    return state[i]


def test_inplace_permutation_planner_passes_compose_to_the_permutation():
    """b2q_sv_permute_bits_inplace = a product of tile passes, each moving at most 13
    (complex64) / 12 (complex128) index bits that include bits 0-3; replaying the
    planned passes on the CPU must give the requested permutation."""
    import ctypes

    lib = _lib.load()
    rng = np.random.RandomState(4)
    cases = []
    for n in (3, 9, 14, 16):
        cases.append((n, list(range(n))[::-1]))                    # bit reversal (QFT order)
        cases.append((n, list(range(1, n)) + [0]))                  # one long cycle
        cases += [(n, rng.permutation(n).tolist()) for _ in range(3)]
    worst = 0
    for dtype_code, cap in ((_lib.C64, 13), (_lib.C128, 12)):
        for n, src in cases:
            out = (ctypes.c_int * (27 * 16))()
            count = ctypes.c_int(0)
            assert lib.b2q_debug_permute_plan(dtype_code, n, _lib.int_array(src), 16, out, ctypes.byref(count)) == 0
            state = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
            want = _permute_reference(state, src)
            got = state
            assert count.value <= 16
            if src == list(range(n)):
                assert count.value == 0
            for p in range(count.value):
                row = list(out[27 * p: 27 * p + 27])
                T = row[0]
                tbits, src_local = row[1:1 + T], row[14:14 + T]
                assert T <= min(cap, n) and tbits == sorted(tbits) and tbits[0] == 0
                assert sorted(src_local) == list(range(T))
                if n >= 4:
                    assert tbits[:4] == [0, 1, 2, 3]  # runs of >= 128 bytes
                full = list(range(n))
                for k in range(T):
                    full[tbits[k]] = tbits[src_local[k]]
                got = _permute_reference(got, full)
            np.testing.assert_array_equal(got, want)
            worst = max(worst, count.value)
    assert worst <= 4
    # the 34-qubit bit reversal (the QFT's relabelled SWAPs): 17 transpositions, <= 5 passes
    out = (ctypes.c_int * (27 * 16))()
    count = ctypes.c_int(0)
    assert lib.b2q_debug_permute_plan(_lib.C64, 34, _lib.int_array(list(range(34))[::-1]), 16, out,
                                      ctypes.byref(count)) == 0
    assert 1 <= count.value <= 5
