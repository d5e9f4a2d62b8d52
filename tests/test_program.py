"""cirq_b200.program: recorded schedules compiled for b2q_run_schedule (one library
call per replay).  CPU: the packed arguments decode to exactly the launches a live
replay makes.  GPU: a native replay leaves bit-identical states."""
import ctypes

import numpy as np
import pytest

from cirq_b200 import _lib, program


def _lib_or_skip():
    try:
        return _lib.load()
    except Exception:  # pragma: no cover
        pytest.skip('C-ABI library not built')


def _unitary(rng, k):
    d = 1 << k
    q, r = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
    return q * (np.diag(r) / np.abs(np.diag(r)))


def _recorded_ops(rng, n, dtype, wide=False):
    """A schedule with every operation kind: lazy growth (basis + kron), dense runs,
    pairs of blocks, diagonal blocks, a scaling and an in-place permutation."""
    ops, ident = [], 0
    sizes = {}

    def new(kind, *rest, bits):
        nonlocal ident
        ident += 1
        sizes[ident] = bits
        ops.append((kind, ident) + rest)
        return ident

    cur = new('basis', 3, 5, bits=3)
    ops.append(('apply', cur, [(_unitary(rng, 2), (2, 0)), (_unitary(rng, 1), (1,))]))
    while sizes[cur] < n:
        add = min(int(rng.randint(1, 4)), n - sizes[cur])
        other = new('basis', add, int(rng.randint(0, 1 << add)), bits=add)
        ops.append(('apply', other, [(_unitary(rng, 1), (0,))], 1))
        cur = new('kron', cur, other, bits=sizes[cur] + add) if rng.randint(2) else new(
            'kron', other, cur, bits=sizes[cur] + add)
        k = sizes[cur]
        blocks = []
        for _ in range(int(rng.randint(1, 5))):
            w = min(k, int(rng.randint(1, 7 if wide else 6)))
            if _lib.dtype_code(np.dtype(dtype)) != _lib.C64:
                w = min(w, 5 if wide else 4)
            wires = tuple(int(x) for x in rng.permutation(k)[:w])
            if rng.rand() < 0.3:
                blocks.append((np.exp(1j * rng.standard_normal(1 << w)), wires))
            else:
                blocks.append((_unitary(rng, w), wires))
        ops.append(('apply', cur, blocks))
    ops.append(('scale', cur, np.exp(0.4j)))
    ops.append(('permute', cur, [int(x) for x in rng.permutation(n)]))
    spare = new('basis', 0, 0, bits=0)
    ops.append(('scale', spare, -1j))
    return ops, cur, spare


def _decode(ns):
    calls = []
    for op in ns.ops[:ns.num_ops]:
        iv, rv = ns.ints[op.ints_offset:], ns.reals[op.reals_offset:]
        if op.kind == program.OP_BASIS:
            calls.append(('basis', op.slot, op.n_bits, int(op.basis_index)))
        elif op.kind == program.OP_KRON:
            calls.append(('kron', op.slot, op.a, op.b, int(iv[0]), int(iv[1]), op.n_bits))
        elif op.kind in (program.OP_DENSE, program.OP_TILE):
            ks = [int(x) for x in iv[:op.count]]
            targets, mats, t, r = [], [], op.count, 0
            for k in ks:
                targets.append(tuple(int(x) for x in iv[t:t + k]))
                mats.append(np.array(rv[r:r + (2 << (2 * k))]).view(np.complex128).reshape(1 << k, 1 << k))
                t += k
                r += 2 << (2 * k)
            calls.append(('tile' if op.kind == program.OP_TILE else 'dense', op.slot, op.n_bits, targets, mats))
        elif op.kind == program.OP_DIAGONAL:
            calls.append(('diag', op.slot, op.n_bits, tuple(int(x) for x in iv[:op.count]),
                          np.array(rv[:2 << op.count]).view(np.complex128)))
        elif op.kind == program.OP_SCALE:
            calls.append(('scale', op.slot, op.n_bits, complex(rv[0], rv[1])))
        elif op.kind == program.OP_PERMUTE:
            assert op.count == op.n_bits
            calls.append(('permute', op.slot, op.n_bits, [int(x) for x in iv[:op.count]]))
        else:
            raise AssertionError(op.kind)
    return calls


def test_schedule_op_layout_matches_the_library():
    lib = _lib_or_skip()
    assert lib.b2q_schedule_op_bytes() == ctypes.sizeof(program.ScheduleOp) == 48
    assert [f[0] for f in program.ScheduleOp._fields_] == [
        'kind', 'slot', 'a', 'b', 'n_bits', 'count', 'ints_offset', 'reals_offset', 'basis_index']


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
@pytest.mark.parametrize('n', [7, 24])
def test_compiled_schedule_decodes_to_the_live_launches(dtype, n):
    _lib_or_skip()
    from cirq_b200.device_state import DeviceState

    rng = np.random.RandomState(n)
    ops, final, spare = _recorded_ops(rng, n, dtype)
    ns = program.compile_schedule(ops, dtype, DeviceState)
    assert ns is not None
    # what a live replay would launch, operation by operation
    want, slot, bits = [], {}, {}
    for op in ops:
        if op[0] == 'basis':
            slot[op[1]], bits[op[1]] = len(slot), op[2]
            want.append(('basis', slot[op[1]], op[2], op[3]))
        elif op[0] == 'kron':
            slot[op[1]], bits[op[1]] = len(slot), bits[op[2]] + bits[op[3]]
            want.append(('kron', slot[op[1]], slot[op[2]], slot[op[3]], bits[op[2]], bits[op[3]], bits[op[1]]))
        elif op[0] == 'apply':
            shape = program._Shape(bits[op[1]], dtype, DeviceState)
            for what, payload in DeviceState.lower_batch(shape, op[2]):
                if what == 'diag':
                    want.append(('diag', slot[op[1]], bits[op[1]], tuple(payload[1]), np.asarray(payload[0])))
                else:
                    want.append((what, slot[op[1]], bits[op[1]], [tuple(w) for _, w in payload],
                                 [np.asarray(m) for m, _ in payload]))
        elif op[0] == 'scale':
            want.append(('scale', slot[op[1]], bits[op[1]], complex(op[2])))
        else:
            want.append(('permute', slot[op[1]], bits[op[1]], list(op[2])))
    got = _decode(ns)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g[:3] == w[:3]
        if g[0] in ('dense', 'tile'):
            assert g[3] == w[3]
            for a, b in zip(g[4], w[4]):
                np.testing.assert_array_equal(a, np.asarray(b, dtype=np.complex128))
        elif g[0] == 'diag':
            assert g[3] == w[3]
            np.testing.assert_array_equal(g[4], np.asarray(w[4], dtype=np.complex128))
        else:
            assert g[3:] == w[3:]
    if n == 24 and dtype == np.complex64:
        assert any(g[0] == 'tile' for g in got)  # pairs of blocks share a tile pass from 22 bits on
    # bookkeeping: which states exist at the end, where they live, how many passes they took
    assert set(ns.alive) == {final, spare}
    assert ns.slot_bits[ns.alive[final]] == n and ns.slot_bits[ns.alive[spare]] == 0
    assert all(off % 256 == 0 for off in ns.offsets)
    amp = 8 if dtype == np.complex64 else 16
    assert ns.total_bytes >= sum(amp << b for b in ns.slot_bits)
    shape_passes = 0
    for op in ops:
        if op[0] == 'apply':
            shape_passes += op[3] if len(op) > 3 else len(
                DeviceState.plan_passes(program._Shape(bits[op[1]], dtype, DeviceState), op[2]))
    assert ns.passes[final] == shape_passes and ns.passes[spare] == 0


def test_schedules_that_are_not_compiled():
    _lib_or_skip()
    from cirq_b200.device_state import DeviceState

    rng = np.random.RandomState(0)
    big = [('basis', 1, program.MAX_BITS + 1, 0)]
    assert program.compile_schedule(big, np.complex64, DeviceState) is None
    grown = [('basis', 1, program.MAX_BITS, 0), ('basis', 2, 1, 0), ('kron', 3, 1, 2)]
    assert program.compile_schedule(grown, np.complex64, DeviceState) is None
    wide = [('basis', 1, 8, 0), ('apply', 1, [(_unitary(rng, 6), (0, 1, 2, 3, 4, 5))])]
    assert program.compile_schedule(wide, np.complex64, DeviceState) is None
    wide128 = [('basis', 1, 8, 0), ('apply', 1, [(_unitary(rng, 5), (0, 1, 2, 3, 4))])]
    assert program.compile_schedule(wide128, np.complex128, DeviceState) is None
    assert program.compile_schedule(wide128, np.complex64, DeviceState) is not None


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
@pytest.mark.parametrize('n', [5, 13, 23])
def test_native_replay_is_bit_identical_to_the_python_replay(dtype, n):
    from cirq_b200 import plan_cache
    from cirq_b200.device_state import DeviceState

    rng = np.random.RandomState(100 + n)
    ops, final, spare = _recorded_ops(rng, n, dtype)
    plan = plan_cache.PrefixPlan((), 1, (), ops, [], None)
    native = program.compile_schedule(ops, dtype, DeviceState)
    assert native is not None
    live_n, passes_n = native.run(DeviceState)

    import os

    os.environ['CIRQ_B200_NATIVE_REPLAY'] = '0'
    try:
        live_p, passes_p = plan_cache.replay(plan, dtype, DeviceState)
    finally:
        del os.environ['CIRQ_B200_NATIVE_REPLAY']
    assert set(live_n) == set(live_p) == {final, spare}
    assert passes_n == passes_p
    for ident in live_p:
        a, b = live_n[ident].to_numpy(), live_p[ident].to_numpy()
        assert a.dtype == b.dtype and a.shape == b.shape
        np.testing.assert_array_equal(a, b)
    assert abs(live_n[final].norm2() - 1) < 1e-4
    # and through plan_cache.replay with the native path on (the default)
    live_d, _ = plan_cache.replay(plan, dtype, DeviceState)
    np.testing.assert_array_equal(live_d[final].to_numpy(), live_p[final].to_numpy())
    # the results are ordinary device states: further passes work on them
    live_d[final].apply_batch([(_unitary(rng, 2), (0, n - 1))])
    assert abs(live_d[final].norm2() - 1) < 1e-4
