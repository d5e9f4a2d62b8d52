"""bench.py body for N > 1 GPUs: the sharded state vector (BASELINE config 4).

Headline (default ``--workload rc_hbm``): cirq.testing.random_circuit on 34 +
log2(N) qubits — 34 LOCAL qubits = 137 GB per GPU, i.e. 35 / 36 / 37 qubits on
2 / 4 / 8 GPUs (37 qubits do not fit fewer than 8 GPUs) — weak scaling.
``value`` counts 30-qubit-equivalent fused gates (gates * 2^(n-30) per second),
the unit of the N = 1 line, so perfect weak scaling multiplies it by N.

The same line carries
  parity   the SAME sharded code path on a 24-qubit circuit of the same generator,
           compared amplitude by amplitude with the 1-GPU kernels and with the
           CPU oracle (rank 0), plus norm^2 of the big state after the timed
           steps; a mismatch makes the process exit non-zero;
  configs  ``rqc_weak30``: the round-1 proxy (Sycamore-style circuit on 30 +
           log2(N) qubits, 8.6 GB per GPU, 1M samples) with its own value /
           roofline / e2e;
  roofline the dominant IN-STEP kernel (CUDA events around every launch of two
           instrumented steps), exchange time and NVLink rate next to it;
  cpu_baseline  cirq.Simulator on rank 0's host cores on a 20-qubit circuit of
           the same generator.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np


def grid_qubits(cirq, n):
    """n GridQubits: the smallest near-square grid with >= n sites, row-major,
    truncated (Sycamore itself is a 54-site grid minus one)."""
    cols = int(np.ceil(np.sqrt(n)))
    rows = int(np.ceil(n / cols))
    return [cirq.GridQubit(r, c) for r in range(rows) for c in range(cols)][:n]


def build(workload, world, n_local=None):
    from cirq_b200._cirq_compat import import_cirq
    from cirq_b200 import workloads as W

    cirq = import_cirq()
    g = world.bit_length() - 1
    if workload == 'rc_hbm':
        n = (34 if n_local is None else n_local) + g
        circuit, qubits = W.random_circuit(n, 20, 1234)
        qubits = list(qubits)
        name = f'cirq.testing.random_circuit {n}q depth 20'
        reps = 0
    else:
        n = (30 if n_local is None else n_local) + g
        qubits = sorted(grid_qubits(cirq, n))
        circuit = cirq.experiments.random_rotations_between_grid_interaction_layers_circuit(
            qubits, depth=20,
            two_qubit_op_factory=lambda a, b, _: cirq.FSimGate(np.pi / 2, np.pi / 6)(a, b), seed=1)
        name = f'Sycamore-style RQC {n}q depth 20'
        reps = 1_000_000
    gates = W.circuit_to_gates(circuit, qubits)
    return circuit, qubits, gates, n, reps, name


def _timed(torch, dist, fn, count):
    """ms per call: CUDA events on this rank's stream between barriers, max over ranks."""
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


def parity_check(world, rank, max_fused, n=24):
    """The sharded path on an n-qubit random circuit vs the 1-GPU kernels and the
    CPU oracle.  Collective; returns the parity dict (identical on every rank)."""
    import torch
    import torch.distributed as dist

    from cirq_b200 import workloads as W
    from cirq_b200.dist import ShardedStateVector, execute_sharded_plan, plan_sharded
    from cirq_b200.fusion import fuse_gates
    from cirq_b200.plan import run_gate_list

    g = world.bit_length() - 1
    circuit, qubits = W.random_circuit(n, 20, 1234)
    gates = W.circuit_to_gates(circuit, list(qubits))
    sv = ShardedStateVector(n, np.complex64, initial_index=None)
    plan = plan_sharded(n, gates, np.complex64, max_fused, sv.n_local)
    execute_sharded_plan(plan, sv)
    got = sv.gather_state()
    norm2 = sv.norm2()
    swaps, passes = sv.swaps, sv.passes
    sv.close()
    out = {'circuit': f'cirq.testing.random_circuit {n}q depth 20 (seed 1234), {len(gates)} gates',
           'n_qubits': n, 'local_qubits': n - g, 'qubit_swaps': swaps, 'passes': passes,
           'tolerance_max_abs': 1e-5, 'norm2': norm2}
    verdict = torch.zeros(3, dtype=torch.float64, device='cuda')
    if rank == 0:
        dev, bit_of, _ = run_gate_list(n, gates, np.complex64, max_fused)
        raw = dev.to_numpy()
        idx = np.arange(1 << n, dtype=np.int64)
        src = np.zeros_like(idx)
        for logical, phys in bit_of.items():
            src |= ((idx >> logical) & 1) << phys
        one_gpu = raw[src]
        del dev, raw, idx, src
        from oracle import sv_oracle as orc  # the checker (bench.py may use it as such)

        want = orc.run_gate_list(n, fuse_gates(gates, 4), dtype=np.complex64)
        verdict[0] = float(np.max(np.abs(got - one_gpu)))
        verdict[1] = float(np.max(np.abs(got - want)))
        verdict[2] = float(np.max(np.abs(one_gpu - want)))
    dist.all_reduce(verdict)  # (zeros elsewhere: a broadcast)
    d1, d2, d3 = (float(v) for v in verdict.cpu())
    out.update(max_abs_diff_vs_1gpu=d1, max_abs_diff_vs_oracle=d2, max_abs_diff_1gpu_vs_oracle=d3)
    ok = d1 <= 1e-5 and d2 <= 1e-5 and abs(norm2 - 1.0) < 1e-4
    out['status'] = 'ok' if ok else 'FAILED'
    return out


def measure(workload, args, world, rank, local_rank, steps, warmup, n_local=None, with_cpu=True):
    import torch
    import torch.distributed as dist

    from cirq_b200 import _lib
    from cirq_b200.dist import B200ShardedSimulator, ShardedStateVector, execute_sharded_plan, plan_sharded
    from cirq_b200._cirq_compat import import_cirq
    import bench as B

    cirq = import_cirq()
    lib = _lib.load()
    peak_gbs, peak_src = B.load_peaks()
    circuit, qubits, gates, n, reps, name = build(workload, world, n_local)
    unit_gates = B.ref_unit_gates(cirq, circuit)
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

    for _ in range(warmup):
        step()
    launches0 = int(lib.b2q_launch_count())
    sv.swaps = sv.passes = sv.fused_exchanges = sv.exchanges = 0
    sv.exchange_volume = 0.0
    with B.ClockSampler(local_rank) as clocks:
        ms_per_step = _timed(torch, dist, step, steps)
        launches = (int(lib.b2q_launch_count()) - launches0) // max(steps, 1)
        swaps = sv.swaps // max(steps, 1)
        passes = sv.passes // max(steps, 1)
        fused_per_step = sv.fused_exchanges // max(steps, 1)
        exchanges_per_step = sv.exchanges // max(steps, 1)
        volume_per_step = sv.exchange_volume / max(steps, 1)
        norm2_big = sv.norm2()

        # two instrumented steps: CUDA events around every launch on the shard
        events: list = []

        def hook(kind, blk, run):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            run()
            b.record()
            if kind == 'exchange':
                kname = 'dist_swap_bit_kernel (+ 2 stream barriers)'
            elif kind.startswith('exchange'):
                kname = f'dist_swap_multi_kernel, {kind[8:]} bits (+ 2 stream barriers)'
            elif kind == 'pass+exchange':
                kname = 'sv_apply_tc_staged_kernel + exchange (b2q_dist_apply_exchange)'
            else:
                kname = B.pass_kernel_name(blk, sv.n_local)
            events.append((kname, a, b))

        record_steps = 2
        sv.on_kernel = hook
        for _ in range(record_steps):
            step()
        torch.cuda.synchronize()
        sv.on_kernel = None
    per_kernel: dict = {}
    for kname, a, b in events:
        per_kernel.setdefault(kname, []).append(a.elapsed_time(b))
    # per-kernel means -> max over ranks would need a fixed key order: keys are
    # identical on every rank (SPMD schedule), so reduce the vector of means
    keys = sorted(per_kernel)
    means = torch.tensor([float(np.mean(per_kernel[k])) for k in keys], dtype=torch.float64, device='cuda')
    dist.all_reduce(means, op=dist.ReduceOp.MAX)
    means = means.cpu().numpy()
    total = sum(means[i] * len(per_kernel[k]) for i, k in enumerate(keys))
    breakdown = {k: {'launches_per_step': len(per_kernel[k]) // record_steps, 'ms_per_launch': float(means[i]),
                     'share_of_step_kernel_time': float(means[i] * len(per_kernel[k]) / total)}
                 for i, k in enumerate(keys)}
    pass_keys = [k for k in keys if 'exchange' not in k and 'swap' not in k] or keys
    dominant = max(pass_keys, key=lambda k: breakdown[k]['share_of_step_kernel_time'])
    pass_ms = breakdown[dominant]['ms_per_launch']
    exch = [k for k in keys if k not in pass_keys]
    exchange_ms_per_step = sum(breakdown[k]['ms_per_launch'] * breakdown[k]['launches_per_step'] for k in exch)
    swap_bytes = shard_bytes // 2
    bare = breakdown.get('dist_swap_bit_kernel (+ 2 stream barriers)')
    multi = {k: {'ms': breakdown[k]['ms_per_launch'],
                 'bytes_out_per_gpu': int(shard_bytes * (1 - 0.5 ** int(k.split(',')[1].split()[0]))),
                 'GBps_per_direction': shard_bytes * (1 - 0.5 ** int(k.split(',')[1].split()[0]))
                 / (breakdown[k]['ms_per_launch'] * 1e-3) / 1e9}
             for k in keys if k.startswith('dist_swap_multi_kernel')}
    equiv = 2.0 ** (n - 30)
    value = unit_gates * equiv / (ms_per_step * 1e-3)
    units = 2 if '2 blocks per pass' in dominant else 1  # fused blocks per launch of the tile kernel
    achieved = units * 2 * shard_bytes / (pass_ms * 1e-3) / 1e9

    # e2e through the Cirq-facing sharded API (host scheduling + result gather inside)
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

    def wall(fn):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        dist.barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device='cuda')
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt.item())

    first_s = wall(e2e_fn)
    e2e_s = wall(e2e_fn)
    sim.close()
    torch.cuda.empty_cache()

    cpu = None
    if with_cpu and not args.no_cpu_baseline:
        if rank == 0:
            r = B.time_reference('rc20', 3, 0)
            cpu = {'value': r['value'] * 2.0 ** (r['bits'] - 30), 'unit': 'gates/s', 'cores': 1,
                   'kind': 'reference', 'raw_value_on_sample': r['value'], 'same_config': False,
                   'sample': f"cirq.Simulator(complex64) on cirq.testing.random_circuit 20q depth 20 (same "
                             f"generator and seed, fewer qubits): {r['raw_ops']} ops = {r['unit_gates']} k<=2 "
                             f"blocks, 3 x {r['seconds_per_step']:.2f} s: {r['value']:.3g} gates/s there, "
                             f"counted as 30-qubit-equivalent gates (x 2^(20-30)); single-threaded numpy, "
                             f"host has {os.cpu_count()} cores"}
        dist.barrier()
    out = {
        'metric': 'fused_gates_per_s', 'value': value, 'unit': 'gates/s', 'n_gpus': world,
        'steps': steps, 'warmup': warmup, 'ms_per_step': ms_per_step, 'dtype': 'c64',
        'config': {'workload': workload, 'circuit': name, 'n_qubits': n,
                   'local_qubits': n - (world.bit_length() - 1),
                   'raw_ops': len(gates), 'unit_gates': unit_gates,
                   'gate_unit': 'k<=2 fused blocks (cirq.merge_k_qubit_unitaries(k=2) count), counted as '
                                '30-qubit equivalents (x 2^(n-30))',
                   'max_fused_qubits': max(len(w) for _, w in B._flat_blocks(blocks)),
                   'passes_per_step': passes,
                   'schedule': ('fusion + lazy state growth: %d of %d raw gates run on replicated '
                                'sub-states before the join into the shards; planned once outside '
                                'the timed region' % (plan.get('prefix_gates', 0), len(gates))),
                   'qubit_swaps_per_step': swaps,
                   'swaps_fused_with_a_gate_pass_per_step': fused_per_step, 'repetitions': reps,
                   'shard_bytes': shard_bytes, 'norm2_after_timed_steps': norm2_big,
                   'exchange': {'ms_per_step': exchange_ms_per_step,
                                'share_of_step': exchange_ms_per_step / ms_per_step,
                                'bytes_out_per_gpu_per_swap': swap_bytes,
                                'bare_swap_ms': bare['ms_per_launch'] if bare else None,
                                'bare_swap_GBps_per_direction':
                                    swap_bytes / (bare['ms_per_launch'] * 1e-3) / 1e9 if bare else None,
                                'multi_bit_exchanges': multi, 'exchange_kernels_per_step': exchanges_per_step,
                                'shard_volumes_sent_per_step': volume_per_step,
                                'nvlink5_GBps_per_direction': 900.0,
                                'how': 'one peer-memory kernel per rank (b2q_dist_swap_bit / '
                                       'b2q_dist_apply_exchange) between two stream-ordered barriers'},
                   'l2': 'inputs larger than L2 (shard %.1f GB)' % (shard_bytes / 1e9)},
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak_gbs, 'unit': 'GB/s',
                     'frac': achieved / peak_gbs, 'traffic': B.measured_traffic(n - (world.bit_length() - 1), dominant),
                     'kernel': dominant + ' (per rank, in-step, max over ranks)',
                     'peak_source': peak_src, 'bytes_per_launch': units * 2 * shard_bytes,
                     'fused_blocks_per_launch': units, 'hbm_bytes_moved_per_launch': 2 * shard_bytes,
                     'frac_of_peak_by_bytes_moved': achieved / units / peak_gbs,
                     'ms_per_launch': pass_ms, 'kernels': breakdown},
        'cpu_baseline': cpu,
        'e2e': {'value': unit_gates * equiv / e2e_s, 'unit': 'gates/s', 'ms_per_step': e2e_s * 1e3,
                'first_call_ms': first_s * 1e3,
                'h2d_bytes_per_step': int(8 * reps), 'd2h_bytes_per_step': int(8 * reps) if reps else 8,
                'api': 'cirq_b200.dist.B200ShardedSimulator.run(circuit, repetitions)' if reps else
                       'cirq_b200.dist.B200ShardedSimulator.simulate_sharded(circuit).norm2()'},
        'gpu_launches': launches, 'clocks': clocks.summary(),
    }
    return out


def run(args, world, rank, local_rank):
    import torch.distributed as dist

    import bench as B

    workload = args.workload or 'rc_hbm'
    if workload not in ('rc_hbm', 'rqc_weak'):
        workload = 'rqc_weak'  # (round-1 spelling: any 1-GPU workload name meant the weak proxy)
    n_local = int(os.environ['B2Q_BENCH_NLOCAL']) if os.environ.get('B2Q_BENCH_NLOCAL') else None
    parity = parity_check(world, rank, args.max_fused)
    head = measure(workload, args, world, rank, local_rank, args.steps, args.warmup, n_local)
    line = {
        'metric': head['metric'], 'value': head['value'], 'unit': head['unit'], 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': head['ms_per_step'],
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'c64',
        'data': 'synthetic', 'config': head['config'], 'roofline': head['roofline'],
        'cpu_baseline': head['cpu_baseline'], 'e2e': head['e2e'], 'gpu_launches': head['gpu_launches'],
        'clocks': head['clocks'], 'parity': parity,
    }
    if not args.no_configs and args.workload is None:
        try:
            sub = measure('rqc_weak', args, world, rank, local_rank, max(3, min(args.steps, 10)),
                          max(3, min(args.warmup, 3)), None, with_cpu=False)
            line['configs'] = {'rqc_weak30': sub}
        except Exception as exc:
            line['configs'] = {'rqc_weak30': {'error': repr(exc)}}
    if rank == 0:
        B.emit(line)
    dist.barrier()
    dist.destroy_process_group()
    if parity['status'] != 'ok':
        sys.exit(3)
