"""Host logic of the sharded state vector with world_size 2 and 4 over gloo
(CPU): the sharded run must reproduce the single-state oracle exactly, whatever
swaps the scheduler chooses."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _diag_worker(rank, world, port, n, out_dir):
    """Only diagonal / controlled operations touch the global qubit: the whole
    circuit must run without a single exchange."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist

    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from cirq_b200.dist import ShardedStateVector
        from fake_dist import GlooShardBackend
        from oracle import sv_oracle as orc

        rng = np.random.RandomState(9)
        h = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
        top = n - 1  # the global qubit for world=2
        gates = [(h, [q]) for q in range(n - 1)]
        for q in range(n - 1):
            cp = np.diag([1, 1, 1, np.exp(1j * rng.standard_normal())])
            gates.append((cp, [top, q]))
            cu = np.eye(4, dtype=complex)
            cu[2:, 2:] = np.linalg.qr(rng.standard_normal((2, 2)) + 1j * rng.standard_normal((2, 2)))[0]
            gates.append((cu, [top, q]))
        gates.append((np.diag([1, 1j]), [top]))
        sv = ShardedStateVector(n, np.complex128, backend=GlooShardBackend(n - 1, np.complex128),
                                initial_index=1 << top)
        sv.apply_blocks(gates)
        got = sv.gather_state()
        want = orc.run_gate_list(n, gates, dtype=np.complex128, initial=1 << top)
        if rank == 0:
            np.savez(os.path.join(out_dir, 'diag.npz'), err=float(np.max(np.abs(got - want))),
                     swaps=sv.swaps, diag=sv.diag_global_blocks)
    finally:
        dist.destroy_process_group()


def test_diagonal_on_global_qubit_needs_no_exchange(tmp_path):
    port = 29500 + (os.getpid() * 3 + 11) % 2000
    mp.spawn(_diag_worker, args=(2, port, 6, str(tmp_path)), nprocs=2, join=True)
    res = np.load(os.path.join(str(tmp_path), 'diag.npz'))
    assert float(res['err']) < 1e-12
    assert int(res['swaps']) == 0 and int(res['diag']) >= 10


def _worker(rank, world, port, n, seed, max_fused, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist

    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from cirq_b200.dist import ShardedStateVector
        from cirq_b200.fusion import fuse_gates
        from fake_dist import GlooShardBackend
        from oracle import sv_oracle as orc

        rng = np.random.RandomState(seed)

        def unitary(k):
            d = 1 << k
            q, r = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
            return q * (np.diag(r) / np.abs(np.diag(r)))

        gates = []
        for _ in range(60):
            k = int(rng.randint(1, 3))
            wires = rng.permutation(n)[:k].tolist()
            kind = rng.randint(4)
            if kind == 0:
                m = np.diag(np.exp(1j * rng.standard_normal(1 << k)))  # diagonal: no swap needed
            elif kind == 1 and k == 2:
                m = np.eye(4, dtype=complex)
                m[2:, 2:] = unitary(1)  # controlled-U: diagonal in its control
            elif kind == 2 and k == 2 and rng.randint(2):
                m = np.eye(4, dtype=complex)[[0, 2, 1, 3]]  # SWAP: relabelled, never exchanged
            else:
                m = unitary(k)
            gates.append((m, wires))
        g = world.bit_length() - 1
        sv = ShardedStateVector(n, np.complex128, backend=GlooShardBackend(n - g, np.complex128),
                                initial_index=3)
        # the production schedule (dist.plan_sharded): diagonal blocks + relabelled SWAPs
        perm = {}
        blocks = fuse_gates(gates, max_fused, np.complex128, n - g, diagonal_blocks=True, permutation=perm)
        sv.apply_blocks(blocks)
        sv.rename_bits(perm)
        got = sv.gather_state()
        want = orc.run_gate_list(n, gates, dtype=np.complex128, initial=3)
        err = float(np.max(np.abs(got - want)))
        nrm = sv.norm2()
        samples = sv.sample(4000, seed=5)
        if rank == 0:
            np.savez(os.path.join(out_dir, f'res_{world}_{n}_{seed}.npz'), err=err, norm=nrm,
                     swaps=sv.swaps, passes=sv.passes, diag=sv.diag_global_blocks, fused=sv.fused_exchanges,
                     samples=samples, probs=np.abs(want) ** 2, phys=np.array(sv.phys))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n,seed,max_fused', [(2, 7, 1, 2), (2, 8, 2, 4), (4, 8, 3, 3), (4, 9, 4, 4)])
def test_sharded_matches_oracle(tmp_path, world, n, seed, max_fused):
    port = 29500 + (os.getpid() + seed * 7) % 2000
    mp.spawn(_worker, args=(world, port, n, seed, max_fused, str(tmp_path)), nprocs=world, join=True)
    res = np.load(os.path.join(str(tmp_path), f'res_{world}_{n}_{seed}.npz'))
    assert float(res['err']) < 1e-12
    assert abs(float(res['norm']) - 1.0) < 1e-12
    assert int(res['swaps']) >= 1  # dense gates on global qubits forced exchanges
    assert 1 <= int(res['fused']) <= int(res['swaps'])  # some rode along with a local pass
    # sampled bitstrings follow |psi|^2 in LOGICAL order: chi-squared on 5 top qubits
    samples, probs = res['samples'], res['probs']
    top = 5
    ints = samples[:, :top].astype(np.int64) @ (1 << np.arange(top - 1, -1, -1))
    hist = np.bincount(ints, minlength=1 << top)
    expect = probs.reshape(1 << top, -1).sum(axis=1) * len(samples)
    chi2 = np.sum((hist - expect) ** 2 / np.maximum(expect, 1e-9))
    dof = (1 << top) - 1
    assert chi2 < dof + 6 * np.sqrt(2 * dof) + 10


def _lazy_worker(rank, world, port, n, seed, pattern, out_dir):
    """Lazy state growth on the sharded path: the circuit prefix runs on small
    replicated sub-states, which are then joined straight into the shards."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist

    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from cirq_b200.dist import ShardedStateVector, execute_sharded_plan, plan_sharded
        from fake_dist import GlooShardBackend
        from oracle import sv_oracle as orc

        rng = np.random.RandomState(seed)

        def unitary(k):
            d = 1 << k
            q, r = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
            return q * (np.diag(r) / np.abs(np.diag(r)))

        gates = [(unitary(1), [q]) for q in range(n)]
        if pattern == 'chain':  # neighbours first: components grow one qubit at a time
            for _ in range(3):
                for q in range(n - 1):
                    gates.append((unitary(2), [q + 1, q]))
                gates += [(unitary(1), [q]) for q in range(n)]
        elif pattern == 'islands':  # never connects everything: the join happens at the end
            for q in range(0, n - 1, 2):
                gates.append((unitary(2), [q, q + 1]))
        else:
            for i in range(50):
                k = int(rng.randint(1, 3))
                wires = rng.permutation(n)[:k].tolist()
                if k == 2 and i % 5 == 0:
                    gates.append((np.eye(4, dtype=complex)[[0, 2, 1, 3]], wires))  # SWAP
                elif k == 2 and i % 5 == 1:
                    gates.append((np.diag(np.exp(1j * rng.standard_normal(4))), wires))
                else:
                    gates.append((unitary(k), wires))
        g = world.bit_length() - 1
        sv = ShardedStateVector(n, np.complex128, backend=GlooShardBackend(n - g, np.complex128),
                                initial_index=None)
        plan = plan_sharded(n, gates, np.complex128, 3, n - g)
        execute_sharded_plan(plan, sv)
        got = sv.gather_state()
        want = orc.run_gate_list(n, gates, dtype=np.complex128, initial=0)
        if rank == 0:
            np.savez(os.path.join(out_dir, f'lazy_{world}_{pattern}.npz'),
                     err=float(np.max(np.abs(got - want))), norm=sv.norm2() if False else 1.0,
                     prefix=plan['prefix_gates'], total=len(gates), passes=sv.passes)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n,seed,pattern', [(2, 8, 1, 'chain'), (4, 9, 2, 'random'), (2, 7, 3, 'islands'),
                                                  (4, 8, 4, 'chain')])
def test_sharded_lazy_growth_matches_oracle(tmp_path, world, n, seed, pattern):
    port = 29500 + (os.getpid() + seed * 13 + 5) % 2000
    mp.spawn(_lazy_worker, args=(world, port, n, seed, pattern, str(tmp_path)), nprocs=world, join=True)
    res = np.load(os.path.join(str(tmp_path), f'lazy_{world}_{pattern}.npz'))
    assert float(res['err']) < 1e-12
    assert int(res['prefix']) >= n  # at least the first layer ran on the small sub-states
    if pattern == 'islands':
        assert int(res['prefix']) == int(res['total']) and int(res['passes']) == 0


def test_block_diagonal_detection():
    sys.path.insert(0, ROOT)
    from cirq_b200.dist import block_diagonal_in

    rng = np.random.RandomState(0)
    u = np.linalg.qr(rng.standard_normal((2, 2)) + 1j * rng.standard_normal((2, 2)))[0]
    cu = np.eye(4, dtype=complex)
    cu[2:, 2:] = u
    subs = block_diagonal_in(cu, [7, 3], [7])
    assert subs is not None
    np.testing.assert_allclose(subs[(0,)], np.eye(2))
    np.testing.assert_allclose(subs[(1,)], u)
    assert block_diagonal_in(cu, [7, 3], [3]) is None
    # control listed second
    sw = np.eye(4)[[0, 2, 1, 3]]
    cu2 = sw @ cu @ sw
    subs = block_diagonal_in(cu2, [3, 7], [7])
    np.testing.assert_allclose(subs[(1,)], u)
    d = np.diag(np.exp(1j * rng.standard_normal(8)))
    subs = block_diagonal_in(d, [5, 4, 2], [5, 2])
    assert subs is not None and subs[(1, 0)].shape == (2, 2)
    np.testing.assert_allclose(subs[(1, 0)], np.diag([d[4, 4], d[6, 6]]))


def _sweep_worker(rank, world, port, out_dir):
    """run_sweep_sharded: resolvers dealt out over the ranks (replicas), results
    gathered in resolver order."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist

    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import sympy

        import cirq_b200
        import cirq_b200.dm_simulator as dmm
        from cirq_b200._cirq_compat import import_cirq
        from cirq_b200.dist import run_sweep_sharded
        from fake_device import OracleDeviceState

        dmm.DeviceState = OracleDeviceState
        cirq = import_cirq()
        q = cirq.LineQubit.range(3)
        t = sympy.Symbol('t')
        # X**t with t in {0, 1, 2, 3}: deterministic outcomes t % 2 per resolver; plus a fair coin
        circuit = cirq.Circuit(cirq.X(q[0]) ** t, cirq.CNOT(q[0], q[1]), cirq.H(q[2]),
                               cirq.measure(*q, key='m'))
        sweep = cirq.Points('t', [0, 1, 2, 3, 1, 0, 1])
        results = run_sweep_sharded(
            lambda s: cirq_b200.B200DensityMatrixSimulator(seed=s), circuit, sweep, repetitions=200, seed=5)
        ok = len(results) == 7
        for r, tv in zip(results, [0, 1, 2, 3, 1, 0, 1]):
            m = r.measurements['m']
            ok &= r.params.value_of('t') == tv and m.shape == (200, 3)
            ok &= bool(np.all(m[:, 0] == tv % 2) and np.all(m[:, 1] == tv % 2))
            ok &= 60 < int(m[:, 2].sum()) < 140
        np.savez(os.path.join(out_dir, f'sweep_{rank}.npz'), ok=ok,
                 first=results[1].measurements['m'])
    finally:
        dist.destroy_process_group()


def test_run_sweep_sharded_replicas(tmp_path):
    world = 2
    port = 29500 + (os.getpid() * 3 + 11) % 2000
    mp.spawn(_sweep_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    a = np.load(os.path.join(str(tmp_path), 'sweep_0.npz'))
    b = np.load(os.path.join(str(tmp_path), 'sweep_1.npz'))
    assert bool(a['ok']) and bool(b['ok'])
    np.testing.assert_array_equal(a['first'], b['first'])  # every rank holds the same gathered results


def _sim_worker(rank, world, port, out_dir):
    """B200ShardedSimulator.run end to end (Cirq circuit -> gates -> live prefix on
    replicated sub-states -> shards -> samples) on the gloo/oracle backend."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist

    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from cirq_b200._cirq_compat import import_cirq
        from cirq_b200.dist import B200ShardedSimulator
        from fake_dist import GlooShardBackend

        cirq = import_cirq()
        n = 9
        g = world.bit_length() - 1
        q = cirq.LineQubit.range(n)
        circuit = cirq.testing.random_circuit(q, 14, 0.9, random_state=3)
        circuit.append(cirq.measure(*q, key='m'))
        want = cirq.Simulator(dtype=np.complex128).simulate(circuit[:-1], qubit_order=q).final_state_vector
        sim = B200ShardedSimulator(dtype=np.complex128, seed=7)
        sim._backend = GlooShardBackend(n - g, np.complex128)  # instead of CUDA IPC shards
        res = sim.run(circuit, repetitions=6000)
        m = res['m']
        top = 5
        ints = m[:, :top].astype(np.int64) @ (1 << np.arange(top - 1, -1, -1))
        hist = np.bincount(ints, minlength=1 << top)
        expect = (np.abs(want) ** 2).reshape(1 << top, -1).sum(axis=1) * len(m)
        chi2 = float(np.sum((hist - expect) ** 2 / np.maximum(expect, 1e-9)))
        np.savez(os.path.join(out_dir, f'sim_{rank}.npz'), chi2=chi2, shape=np.array(m.shape), first=m[:50])
    finally:
        dist.destroy_process_group()


def test_sharded_simulator_run_end_to_end(tmp_path):
    world = 2
    port = 29500 + (os.getpid() * 5 + 17) % 2000
    mp.spawn(_sim_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    a = np.load(os.path.join(str(tmp_path), 'sim_0.npz'))
    b = np.load(os.path.join(str(tmp_path), 'sim_1.npz'))
    assert tuple(a['shape']) == (6000, 9)
    dof = 31
    assert float(a['chi2']) < dof + 6 * np.sqrt(2 * dof) + 10
    np.testing.assert_array_equal(a['first'], b['first'])  # identical on every rank
