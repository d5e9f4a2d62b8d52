"""Multi-GPU correctness + swap bandwidth check (run under torchrun on >= 2 GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/dist_check.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank = int(os.environ['RANK'])
    local_rank = int(os.environ['LOCAL_RANK'])
    world = int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(local_rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from cirq_b200.device_state import DeviceState
    from cirq_b200.dist import ShardedStateVector, execute_sharded_plan, plan_sharded
    from cirq_b200.fusion import fuse_gates

    rng = np.random.RandomState(7)

    def unitary(k):
        d = 1 << k
        q, r = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
        return q * (np.diag(r) / np.abs(np.diag(r)))

    ok = True
    for dtype, atol in ((np.complex64, 1e-5), (np.complex128, 1e-12)):
        for n in (12, 16, 20):
            gates = []
            for _ in range(80):
                k = int(rng.randint(1, 4))
                gates.append((unitary(k), rng.permutation(n)[:k].tolist()))
            blocks = fuse_gates(gates, 4)
            sv = ShardedStateVector(n, dtype, initial_index=5)
            sv.apply_blocks(blocks)
            got = sv.gather_state()
            ref = DeviceState.basis(n, dtype, 5)
            ref.apply_batch(blocks)
            want = ref.to_numpy()
            err = float(np.max(np.abs(got - want)))
            if n <= 16:
                # ... and against the CPU oracle (the checker; gate by gate, unfused)
                from oracle import sv_oracle as orc

                want_oracle = orc.run_gate_list(n, gates, dtype=dtype, initial=5)
                err = max(err, float(np.max(np.abs(got - want_oracle))))
            nrm = sv.norm2()
            samples = sv.sample(20000, seed=3)
            # chi-squared of the top 4 logical qubits against the single-GPU probabilities
            probs = np.abs(want.astype(np.complex128)) ** 2
            ints = samples[:, :4].astype(np.int64) @ (1 << np.arange(3, -1, -1))
            hist = np.bincount(ints, minlength=16)
            expect = probs.reshape(16, -1).sum(axis=1) * len(samples)
            chi2 = float(np.sum((hist - expect) ** 2 / np.maximum(expect, 1e-9)))
            good = err <= atol and abs(nrm - 1) < 1e-4 and chi2 < 15 + 6 * np.sqrt(30) + 10
            ok &= good
            if rank == 0:
                print(f'n={n} {np.dtype(dtype)} world={world}: max|diff| vs 1 GPU{" and oracle" if n <= 16 else ""}'
                      f'={err:.2e} norm={nrm:.6f} swaps={sv.swaps} exchanges={sv.exchanges} '
                      f'(volume {sv.exchange_volume:.3f} shards) passes={sv.passes} chi2={chi2:.1f} '
                      f'{"OK" if good else "FAIL"}',
                      flush=True)
            sv.close()
    # lazy state growth from |0...0>: prefix on replicated sub-states, join into the shards
    g = world.bit_length() - 1
    for dtype, atol in ((np.complex64, 1e-5), (np.complex128, 1e-12)):
        n = 18
        gates = [(unitary(1), [q]) for q in range(n)]
        for _ in range(3):
            gates += [(unitary(2), [q + 1, q]) for q in range(n - 1)]
            gates += [(unitary(3), rng.permutation(n)[:3].tolist()) for _ in range(6)]
        sv = ShardedStateVector(n, dtype, initial_index=None)
        plan = plan_sharded(n, gates, dtype, None, n - g)
        execute_sharded_plan(plan, sv)
        got = sv.gather_state()
        ref = DeviceState.basis(n, dtype, 0)
        ref.apply_batch(fuse_gates(gates, 4))
        err = float(np.max(np.abs(got - ref.to_numpy())))
        nrm = sv.norm2()
        good = err <= atol and abs(nrm - 1) < 1e-4 and plan['prefix_gates'] >= n
        ok &= good
        if rank == 0:
            print(f'lazy growth n={n} {np.dtype(dtype)} world={world}: prefix {plan["prefix_gates"]} of '
                  f'{len(gates)} gates, max|diff|={err:.2e} norm={nrm:.6f} swaps={sv.swaps} '
                  f'passes={sv.passes} {"OK" if good else "FAIL"}', flush=True)
        sv.close()
    # the production schedule on a diagonal / SWAP heavy circuit: diagonal blocks are
    # applied without communication, SWAP gates only rename bits
    swap = np.eye(4, dtype=complex)[[0, 2, 1, 3]]
    for dtype, atol in ((np.complex64, 1e-5), (np.complex128, 1e-12)):
        n = 18
        gates = [(unitary(1), [q]) for q in range(n)]
        for i in range(90):
            a, b = rng.permutation(n)[:2].tolist()
            kind = i % 4
            if kind == 0:
                gates.append((swap, [a, b]))
            elif kind == 1:
                gates.append((np.diag(np.exp(1j * rng.standard_normal(4))), [a, b]))
            elif kind == 2:
                gates.append((unitary(2), [a, b]))
            else:
                gates.append((unitary(1), [a]))
        sv = ShardedStateVector(n, dtype, initial_index=9)
        perm = {}
        blocks = fuse_gates(gates, None, dtype, n - g, diagonal_blocks=True, permutation=perm)
        sv.apply_blocks(blocks)
        sv.rename_bits(perm)
        got = sv.gather_state()
        ref = DeviceState.basis(n, dtype, 9)
        ref.apply_batch(fuse_gates(gates, 4))
        err = float(np.max(np.abs(got - ref.to_numpy())))
        good = err <= atol and any(np.ndim(m) == 1 for m, _ in blocks) and bool(perm)
        ok &= good
        if rank == 0:
            print(f'diagonal blocks + relabelled SWAPs n={n} {np.dtype(dtype)} world={world}: '
                  f'{sum(1 for m, _ in blocks if np.ndim(m) == 1)} diagonal of {len(blocks)} blocks, '
                  f'max|diff|={err:.2e} swaps={sv.swaps} diag-global={sv.diag_global_blocks} '
                  f'{"OK" if good else "FAIL"}', flush=True)
        sv.close()
    # swap bandwidth at a large shard
    n_local = int(os.environ.get('B2Q_SWAP_NLOCAL', '30'))
    n = n_local + world.bit_length() - 1
    sv = ShardedStateVector(n, np.complex64)
    for lbit in (n_local - 1, n_local - 5, 8):
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sv.swap_global_local(n_local, lbit)
        s.record()
        reps = 4
        for _ in range(reps):
            sv.swap_global_local(n_local, lbit)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / reps
        half = sv.local.nbytes / 2
        if rank == 0:
            print(f'swap global<->local bit {lbit}: {ms:.2f} ms, {half / ms / 1e6:.0f} GB/s per direction '
                  f'(half shard {half / 1e9:.2f} GB)', flush=True)
    if world >= 4:
        g = world.bit_length() - 1
        for m in range(2, g + 1):
            pairs = [(n_local + i, n_local - 1 - 2 * i) for i in range(m)]
            torch.cuda.synchronize()
            dist.barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sv.exchange_bits(pairs)
            s.record()
            reps = 4
            for _ in range(reps):
                sv.exchange_bits([(n_local + i, n_local - 1 - 2 * i) for i in range(m)])
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / reps
            out = sv.local.nbytes * (1 - 0.5 ** m)
            if rank == 0:
                print(f'{m}-bit exchange: {ms:.2f} ms, {out / ms / 1e6:.0f} GB/s per direction '
                      f'({out / 1e9:.2f} GB out per GPU = {1 - 0.5 ** m:.3f} shard; {m} single swaps move '
                      f'{m * 0.5:.1f} shards)', flush=True)
    sv.close()
    if rank == 0:
        print('DIST CHECK', 'PASSED' if ok else 'FAILED', flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
