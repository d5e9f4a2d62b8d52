"""bench.py body for N > 1 GPUs: the sharded state vector (config 4 family).

Weak scaling: N GPUs simulate a Sycamore-style circuit on 30 + log2(N) qubits
(30 local qubits = 8.6 GB per GPU, the N = 1 workload's state size), or, with
``--workload rc_hbm``, cirq.testing.random_circuit on 34 + log2(N) qubits — the
BASELINE config (37 qubits on 8 GPUs, 137 GB per GPU).  ``value`` counts
30-qubit-equivalent fused gates: gates * 2^(n-30) per second, so that perfect
weak scaling multiplies it by N.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np


def grid_qubits(cirq, n):
    """n GridQubits: the smallest near-square grid with >= n sites, row-major,
    truncated (Sycamore itself is a 54-site grid minus one)."""
    cols = int(np.ceil(np.sqrt(n)))
    rows = int(np.ceil(n / cols))
    return [cirq.GridQubit(r, c) for r in range(rows) for c in range(cols)][:n]


def build(args, world):
    from cirq_b200._cirq_compat import import_cirq
    from cirq_b200 import workloads as W

    cirq = import_cirq()
    g = world.bit_length() - 1
    if args.workload == 'rc_hbm':
        n = 34 + g
        circuit, qubits = W.random_circuit(n, 20, 1234)
        qubits = list(qubits)
        name = f'cirq.testing.random_circuit {n}q depth 20'
        reps = 0
    else:
        n = 30 + g
        qubits = sorted(grid_qubits(cirq, n))
        circuit = cirq.experiments.random_rotations_between_grid_interaction_layers_circuit(
            qubits, depth=20,
            two_qubit_op_factory=lambda a, b, _: cirq.FSimGate(np.pi / 2, np.pi / 6)(a, b), seed=1)
        name = f'Sycamore-style RQC {n}q depth 20'
        reps = 1_000_000
    gates = W.circuit_to_gates(circuit, qubits)
    return circuit, qubits, gates, n, reps, name


def run(args, world, rank, local_rank):
    import torch
    import torch.distributed as dist

    from cirq_b200 import _lib
    from cirq_b200.dist import ShardedStateVector, execute_sharded_plan, plan_sharded
    from cirq_b200.fusion import fuse_gates
    import bench as B

    lib = _lib.load()
    peak_gbs, peak_src = B.load_peaks()
    circuit, qubits, gates, n, reps, name = build(args, world)
    unit_gates = len(fuse_gates(gates, 2))
    sv = ShardedStateVector(n, np.complex64, initial_index=None)
    shard_bytes = sv.local.nbytes
    # Host scheduling happens once, outside the timed region (as in the 1-GPU
    # bench): the circuit prefix that runs on small replicated sub-states, the
    # join into the shards, and the fused blocks for the sharded state.
    plan = plan_sharded(n, gates, np.complex64, args.max_fused, sv.n_local)
    blocks = plan['blocks']

    def step():
        execute_sharded_plan(plan, sv)
        if reps:
            sv.sample(reps, seed=0)

    def timed(fn, count):
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(count):
            fn()
        e.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / count

    for _ in range(args.warmup):
        step()
    launches0 = int(lib.b2q_launch_count())
    sv.swaps = sv.passes = sv.fused_exchanges = 0
    with B.ClockSampler(local_rank) as clocks:
        ms_per_step = timed(step, args.steps)
    launches = (int(lib.b2q_launch_count()) - launches0) // max(args.steps, 1)
    swaps = sv.swaps // max(args.steps, 1)
    passes = sv.passes // max(args.steps, 1)
    fused_per_step = sv.fused_exchanges // max(args.steps, 1)

    # instrumented pieces: one local pass and one qubit swap, timed alone
    h = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    u2 = np.kron(h, h)
    pass_ms = timed(lambda: sv.local.apply_matrix(u2, [sv.n_local - 1, sv.n_local - 2]), 5)
    swap_ms = timed(lambda: sv.swap_global_local(sv.n_local, sv.n_local - 1), 4)
    fused_ms = None
    if getattr(sv.backend, 'can_fuse_exchange', False):
        rs = np.random.RandomState(5)
        q5, _ = np.linalg.qr(rs.standard_normal((32, 32)) + 1j * rs.standard_normal((32, 32)))
        bits5 = [sv.n_local - 3, 12, 9, 7, 3]
        fused_ms = timed(lambda: sv.swap_global_local(sv.n_local, sv.n_local - 1, fused_block=(q5, bits5)), 4)
    swap_bytes = shard_bytes // 2
    equiv = 2.0 ** (n - 30)
    value = unit_gates * equiv / (ms_per_step * 1e-3)

    # e2e through the Cirq-facing sharded API (host scheduling + result gather inside)
    from cirq_b200.dist import B200ShardedSimulator
    from cirq_b200._cirq_compat import import_cirq

    cirq = import_cirq()
    sv.close()
    del sv
    torch.cuda.empty_cache()
    sim = B200ShardedSimulator(dtype=np.complex64, seed=0, max_fused_qubits=args.max_fused)
    if reps:
        full = circuit + cirq.Circuit(cirq.measure(*qubits, key='m'))
        e2e_fn = lambda: sim.run(full, repetitions=reps)
    else:
        def e2e_fn():
            s = sim.simulate_sharded(circuit, qubit_order=qubits)
            s.norm2()
            s.close()
    e2e_fn()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    e2e_fn()
    torch.cuda.synchronize()
    dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device='cuda')
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_s = float(dt.item())

    sim.close()
    if rank == 0:
        achieved = 2 * shard_bytes / (pass_ms * 1e-3) / 1e9
        line = {
            'metric': 'fused_gates_per_s', 'value': value, 'unit': 'gates/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'c64',
            'data': 'synthetic',
            'config': {'workload': name, 'n_qubits': n, 'local_qubits': n - (world.bit_length() - 1),
                       'raw_ops': len(gates), 'unit_gates': unit_gates,
                       'gate_unit': 'k<=2 fused blocks, counted as 30-qubit equivalents (x 2^(n-30))',
                       'max_fused_qubits': max(len(w) for _, w in blocks), 'passes_per_step': passes,
                       'schedule': ('fusion + lazy state growth: %d of %d raw gates run on replicated '
                                    'sub-states before the join into the shards; planned once outside '
                                    'the timed region' % (plan.get('prefix_gates', 0), len(gates))),
                       'qubit_swaps_per_step': swaps,
                       'swaps_fused_with_a_gate_pass_per_step': fused_per_step, 'repetitions': reps,
                       'shard_bytes': shard_bytes,
                       'swap': {'bytes_out_per_gpu': swap_bytes, 'ms': swap_ms,
                                'GBps_per_direction': swap_bytes / (swap_ms * 1e-3) / 1e9,
                                'nvlink_ref_GBps': 770.0,
                                'how': 'one b2q_dist_swap_bit kernel per rank over peer memory',
                                'fused_gate_plus_swap_ms': fused_ms,
                                'fused_how': ('b2q_dist_apply_exchange: a 5-qubit tensor-core pass whose '
                                              'results are written straight to the partner over NVLink '
                                              '(vs pass + swap as two kernels: ms + local pass ms)')},
                       'l2': 'inputs larger than L2 (shard %.1f GB)' % (shard_bytes / 1e9)},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak_gbs, 'unit': 'GB/s',
                         'frac': achieved / peak_gbs, 'traffic': None,
                         'kernel': 'sv_apply_fast_kernel (per rank, local pass)',
                         'peak_source': peak_src, 'bytes_per_launch': 2 * shard_bytes,
                         'ms_per_launch': pass_ms},
            'cpu_baseline': None,
            'e2e': {'value': unit_gates * equiv / e2e_s, 'unit': 'gates/s', 'ms_per_step': e2e_s * 1e3,
                    'h2d_bytes_per_step': int(8 * reps), 'd2h_bytes_per_step': int(8 * reps),
                    'api': 'cirq_b200.dist.B200ShardedSimulator.run(circuit, repetitions)'},
            'gpu_launches': launches, 'clocks': clocks.summary(),
        }
        B.emit(line)
    dist.barrier()
    dist.destroy_process_group()
