"""Noisy state-vector `run`: batched trajectories against the per-repetition loop.

    python tools/traj_bench.py [--qubits 16] [--depth 8] [--reps 4096] [--out file.json]

The circuit is a brickwork of random 1-qubit rotations + CZ with
``cirq.depolarize(p)`` as the simulator's noise model and a terminal measurement
— the shape on which ``cirq.Simulator.run`` falls back to one simulation per
repetition (cirq-core/cirq/sim/simulator_base.py:249-264).  Three timings:

  batched   B200Simulator(trajectory_batch=B).run(circuit, reps)
  loop      B200Simulator().run(circuit, loop_reps)          (same kernels, one state per repetition)
  reference cirq.Simulator().run(circuit, ref_reps)          (host numpy, bounded sample)

and a chi-squared check of the batched histogram on a 6-qubit marginal against
the loop's histogram is left to tests/; here only throughput is reported.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def noisy_brickwork(cirq, n, depth, seed):
    rng = np.random.RandomState(seed)
    q = cirq.LineQubit.range(n)
    c = cirq.Circuit()
    for d in range(depth):
        c.append(cirq.Moment(cirq.PhasedXZGate(x_exponent=rng.rand(), z_exponent=rng.rand(),
                                               axis_phase_exponent=rng.rand()).on(x) for x in q))
        c.append(cirq.Moment(cirq.CZ(q[i], q[i + 1]) for i in range(d % 2, n - 1, 2)))
    c.append(cirq.measure(*q, key='m'))
    return c, q


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--qubits', type=int, default=16)
    ap.add_argument('--depth', type=int, default=8)
    ap.add_argument('--reps', type=int, default=4096)
    ap.add_argument('--batch', type=int, default=4096)
    ap.add_argument('--loop-reps', type=int, default=64)
    ap.add_argument('--ref-reps', type=int, default=8)
    ap.add_argument('--p', type=float, default=0.01)
    ap.add_argument('--out', default='')
    args = ap.parse_args()
    import torch

    import cirq_b200
    from cirq_b200._cirq_compat import import_cirq

    cirq = import_cirq()
    circuit, q = noisy_brickwork(cirq, args.qubits, args.depth, 0)
    noise = cirq.depolarize(args.p)
    rep = dict(n_qubits=args.qubits, depth=args.depth, noise=f'depolarize({args.p})',
               ops=len(list(circuit.all_operations())))

    def timed(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        return out, time.perf_counter() - t0

    sim = cirq_b200.B200Simulator(noise=noise, seed=1, trajectory_batch=args.batch)
    sim.run(circuit, repetitions=min(args.reps, 256))  # warm-up (allocator, unitary cache)
    res, dt = timed(lambda: sim.run(circuit, repetitions=args.reps))
    assert sim.last_run_info['path'] == 'batched trajectories', sim.last_run_info
    rep['batched'] = dict(reps=args.reps, seconds=dt, reps_per_s=args.reps / dt, **sim.last_run_info)
    ones_b = res.measurements['m'].mean(axis=0)

    loop = cirq_b200.B200Simulator(noise=noise, seed=1)
    loop.run(circuit, repetitions=2)
    res, dt = timed(lambda: loop.run(circuit, repetitions=args.loop_reps))
    rep['loop'] = dict(reps=args.loop_reps, seconds=dt, reps_per_s=args.loop_reps / dt)

    if args.ref_reps:
        ref = cirq.Simulator(noise=noise, seed=1)
        t0 = time.perf_counter()
        ref.run(circuit, repetitions=args.ref_reps)
        dt = time.perf_counter() - t0
        rep['reference_cpu'] = dict(reps=args.ref_reps, seconds=dt, reps_per_s=args.ref_reps / dt,
                                    cores=1, host_cores=os.cpu_count())
    rep['speedup_vs_loop'] = rep['batched']['reps_per_s'] / rep['loop']['reps_per_s']
    if 'reference_cpu' in rep:
        rep['speedup_vs_reference_cpu'] = rep['batched']['reps_per_s'] / rep['reference_cpu']['reps_per_s']
    rep['mean_ones_per_qubit_batched'] = [round(float(x), 4) for x in ones_b]
    line = json.dumps(rep)
    print(line)
    if args.out:
        with open(args.out, 'w') as f:
            f.write(line + '\n')


if __name__ == '__main__':
    main()
