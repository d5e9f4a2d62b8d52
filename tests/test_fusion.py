"""The host scheduler must preserve circuit semantics: applying the fused
blocks equals applying the original gates one by one (checked with the CPU
oracle), for every max block width."""
import numpy as np
import pytest

from cirq_b200.fusion import GateFuser, expand_matrix, fuse_gates
from oracle import sv_oracle as orc


def rand_unitary(rng, k):
    d = 1 << k
    q, r = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
    return q * (np.diag(r) / np.abs(np.diag(r)))


def random_gates(rng, n, count, max_k=3):
    gates = []
    for _ in range(count):
        k = int(rng.randint(1, min(max_k, n) + 1))
        wires = rng.permutation(n)[:k].tolist()
        gates.append((rand_unitary(rng, k), wires))
    return gates


def run(n, gates, psi=None):
    if psi is None:
        psi = np.zeros(1 << n, dtype=np.complex128)
        psi[0] = 1
    for m, w in gates:
        m = np.asarray(m)
        if m.ndim == 1:  # a diagonal block: its 2^k diagonal entries
            m = np.diag(m)
        psi = orc.apply_matrix(psi, n, m, list(w))
    return psi


SWAP = np.eye(4)[[0, 2, 1, 3]]


def diag_heavy_gates(rng, n, count):
    """Mix of dense gates, 1/2/3-qubit diagonal gates and SWAPs."""
    gates = []
    for _ in range(count):
        kind = rng.randint(0, 6)
        if kind == 0:
            gates.append((SWAP.astype(np.complex128), rng.permutation(n)[:2].tolist()))
        elif kind in (1, 2, 3):
            k = int(rng.randint(1, min(3, n) + 1))
            d = np.exp(1j * rng.uniform(0, 2 * np.pi, size=1 << k))
            gates.append((np.diag(d), rng.permutation(n)[:k].tolist()))
        else:
            k = int(rng.randint(1, min(2, n) + 1))
            gates.append((rand_unitary(rng, k), rng.permutation(n)[:k].tolist()))
    return gates


@pytest.mark.parametrize('diag_max', [0, 6, 9])
@pytest.mark.parametrize('relabel', [False, True])
@pytest.mark.parametrize('max_q', [2, 4, 5])
def test_diagonal_blocks_and_swap_relabelling_preserve_semantics(max_q, relabel, diag_max):
    for n, seed in ((3, 0), (6, 1), (9, 2), (10, 3)):
        rng = np.random.RandomState(1000 * seed + 10 * max_q + diag_max)
        gates = diag_heavy_gates(rng, n, 150)
        psi0 = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
        psi0 /= np.linalg.norm(psi0)
        want = run(n, gates, psi0.copy())
        f = GateFuser(max_q, diag_max=diag_max, relabel_swaps=relabel)
        for m, w in gates:
            f.add(m, w)
        blocks = f.blocks()
        assert all(len(w) <= (diag_max if np.ndim(m) == 1 else max(max_q, 3)) for m, w in blocks)
        np.testing.assert_allclose(run(n, blocks, psi0.copy()), want, atol=1e-10)
        # streaming: early releases + a final flush, the permutation taken instead of restored
        f = GateFuser(max_q, diag_max=diag_max, relabel_swaps=relabel)
        emitted = []
        for i, (m, w) in enumerate(gates):
            f.add(m, w)
            if i % 11 == 10:
                emitted += f.pop_final_blocks()
        emitted += f.blocks(restore=False)
        perm = f.take_permutation()
        got = run(n, emitted, psi0.copy())
        if perm:
            assert relabel
            # caller's wire w now lives on wire perm[w]: undo by reading bit perm[w] for w
            idx = np.arange(1 << n)
            src = np.zeros_like(idx)
            for w in range(n):
                src |= ((idx >> w) & 1) << perm.get(w, w)
            got = got[src]
        np.testing.assert_allclose(got, want, atol=1e-10)
        assert f.take_permutation() == {}


def test_qft_schedule_uses_diagonal_blocks():
    """The example QFT (H, CZ**t, SWAP chains): relabelled swaps + diagonal blocks
    cut the pass count several times, result unchanged."""
    from cirq_b200 import workloads as W

    pytest.importorskip('sympy')
    n = 12
    try:
        circuit, qubits = W.qft_circuit(n)
    except Exception:
        pytest.skip('cirq not importable')
    gates = W.circuit_to_gates(circuit, qubits)
    dense = fuse_gates(gates, 5)
    f = GateFuser(5, diag_max=13, relabel_swaps=True)
    for m, w in gates:
        f.add(m, w)
    smart = f.blocks()
    assert len(smart) < len(dense)
    assert any(np.ndim(m) == 1 for m, _ in smart)
    rng = np.random.RandomState(4)
    psi0 = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    psi0 /= np.linalg.norm(psi0)
    np.testing.assert_allclose(run(n, smart, psi0.copy()), run(n, gates, psi0.copy()), atol=1e-10)


@pytest.mark.parametrize('max_q', [1, 2, 3, 4, 5])
@pytest.mark.parametrize('n', [2, 5, 8])
def test_fused_blocks_equal_sequential(n, max_q):
    rng = np.random.RandomState(100 * n + max_q)
    gates = random_gates(rng, n, 60)
    fused = fuse_gates(gates, max_q)
    assert all(len(w) <= max(max_q, 3) for _, w in fused)
    np.testing.assert_allclose(run(n, fused), run(n, gates), atol=1e-10)
    if max_q >= 2:
        assert len(fused) < len(gates)


def test_expand_matrix_matches_kron():
    rng = np.random.RandomState(0)
    a = rand_unitary(rng, 1)
    b = rand_unitary(rng, 2)
    np.testing.assert_allclose(expand_matrix(a, [5], [5, 3]), np.kron(a, np.eye(2)), atol=1e-14)
    np.testing.assert_allclose(expand_matrix(a, [3], [5, 3]), np.kron(np.eye(2), a), atol=1e-14)
    np.testing.assert_allclose(expand_matrix(b, [7, 2], [7, 4, 2]).reshape(2, 2, 2, 2, 2, 2)[:, 0, :, :, 0, :].reshape(4, 4), b, atol=1e-14)
    # wire order swap
    sw = np.eye(4)[[0, 2, 1, 3]]
    np.testing.assert_allclose(expand_matrix(b, [2, 7], [7, 2]), sw @ b @ sw, atol=1e-14)


def test_layered_circuit_fuses_to_few_blocks():
    """Sycamore-like layer structure: 1-qubit layer + disjoint 2-qubit layer
    must collapse into the 2-qubit blocks (no leftover 1-qubit passes)."""
    rng = np.random.RandomState(3)
    n = 8
    gates = []
    for layer in range(4):
        for q in range(n):
            gates.append((rand_unitary(rng, 1), [q]))
        start = layer % 2
        for q in range(start, n - 1, 2):
            gates.append((rand_unitary(rng, 2), [q, q + 1]))
    fused2 = fuse_gates(gates, 2)
    assert all(len(w) == 2 for _, w in fused2[:-1]) or len(fused2) <= 16
    np.testing.assert_allclose(run(n, fused2), run(n, gates), atol=1e-10)
    fused4 = fuse_gates(gates, 4)
    assert len(fused4) < len(fused2)
    np.testing.assert_allclose(run(n, fused4), run(n, gates), atol=1e-10)


def test_fuser_incremental_interface():
    f = GateFuser(3)
    rng = np.random.RandomState(1)
    gates = random_gates(rng, 4, 10, max_k=2)
    for m, w in gates:
        f.add(m, w)
    assert len(f) == 10
    np.testing.assert_allclose(run(4, f.blocks()), run(4, gates), atol=1e-10)
    f.clear()
    assert f.blocks() == [] and len(f) == 0


@pytest.mark.parametrize('max_q', [2, 3, 4])
def test_streaming_drain_preserves_semantics(max_q):
    """Blocks released early by pop_final_blocks followed by the rest equal the
    sequential circuit, and released blocks never reappear."""
    rng = np.random.RandomState(50 + max_q)
    n = 7
    gates = random_gates(rng, n, 120, max_k=2)
    f = GateFuser(max_q)
    emitted = []
    for i, (m, w) in enumerate(gates):
        f.add(m, w)
        if i % 9 == 8:
            emitted += f.pop_final_blocks()
    early = len(emitted)
    emitted += f.blocks()
    assert early > 0
    np.testing.assert_allclose(run(n, emitted), run(n, gates), atol=1e-10)
    assert len(emitted) <= len(fuse_gates(gates, max_q)) + 2


def test_split_executor_matches_sequential():
    """Lazily growing product of states (cirq_b200.plan) == dense application."""
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from fake_device import OracleDeviceState

    from cirq_b200.plan import run_gate_list

    rng = np.random.RandomState(77)
    n = 7
    gates = []
    for q in range(n):
        gates.append((rand_unitary(rng, 1), [q]))
    for q in range(0, n - 1, 2):
        gates.append((rand_unitary(rng, 2), [q + 1, q]))
    gates += random_gates(rng, n, 40, max_k=2)
    dev, bit_of, passes = run_gate_list(n, gates, np.complex128, 3, OracleDeviceState)
    got = dev.to_numpy()
    want = run(n, gates)
    idx = np.arange(1 << n)
    src = np.zeros_like(idx)
    for b in range(n):
        src |= ((idx >> b) & 1) << bit_of[b]
    np.testing.assert_allclose(got[src], want, atol=1e-10)
    assert passes > 0


def test_plan_replay_equals_direct_execution():
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from fake_device import OracleDeviceState

    from cirq_b200.plan import build_plan, replay_plan

    rng = np.random.RandomState(78)
    n = 6
    gates = [(rand_unitary(rng, 1), [q]) for q in range(n)] + random_gates(rng, n, 50, max_k=2)
    plan = build_plan(n, gates, np.complex128, 3)
    dev = replay_plan(plan, OracleDeviceState)
    got = dev.to_numpy()
    idx = np.arange(1 << n)
    src = np.zeros_like(idx)
    for b in range(n):
        src |= ((idx >> b) & 1) << plan['bit_of'][b]
    np.testing.assert_allclose(got[src], run(n, gates), atol=1e-10)
    # replaying twice gives the same state (plans are reusable)
    np.testing.assert_array_equal(replay_plan(plan, OracleDeviceState).to_numpy(), got)


def test_native_block_composition_matches_numpy(monkeypatch):
    """b2q_host_compose / b2q_host_compose_diag (one native call per emitted block)
    against the numpy composition the fuser falls back to without the library."""
    from cirq_b200 import fusion

    if not fusion._native():
        pytest.skip('C-ABI library not built')
    rng = np.random.RandomState(11)
    union = (9, 7, 4, 3, 1, 0)
    dense = []
    for wires in ((7, 0), (3,), (9, 4, 1), (0, 7), (4, 9)):
        d = 1 << len(wires)
        dense.append((rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)), wires))
    diags = [(np.exp(1j * rng.standard_normal(1 << len(w))), w) for w in ((1, 9), (0,), (7, 3, 4), (9, 1))]
    native_dense = fusion._materialize(dense, union)
    native_diag = fusion._materialize_diag(diags, union)
    monkeypatch.setattr(fusion, '_NATIVE', False)
    np.testing.assert_allclose(native_dense, fusion._materialize(dense, union), atol=1e-12)
    np.testing.assert_allclose(native_diag, fusion._materialize_diag(diags, union), atol=1e-13)
