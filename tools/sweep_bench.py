"""run_sweep over many resolvers: batched (sweep_batch=True) against the
resolver-by-resolver loop, on config 5's circuit family at sizes where the state
is small (noisy max-cut QAOA, examples/qaoa.py:128-158; 256 resolvers).

    python tools/sweep_bench.py [--kind dm|sv] [--qubits 10] [--resolvers 256] [--reps 1000]

Reports resolvers/s for: batched, the sequential B200 path (bounded sample of the
resolvers) and the reference simulator on the host (bounded sample).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--kind', default='dm', choices=['dm', 'sv'])
    ap.add_argument('--qubits', type=int, default=10)
    ap.add_argument('--resolvers', type=int, default=256)
    ap.add_argument('--reps', type=int, default=1000)
    ap.add_argument('--seq-resolvers', type=int, default=16)
    ap.add_argument('--ref-resolvers', type=int, default=2)
    ap.add_argument('--expect', action='store_true',
                    help='simulate_expectation_values_sweep of the max-cut cost instead of run_sweep')
    ap.add_argument('--out', default='')
    args = ap.parse_args()
    import torch

    import cirq_b200
    from cirq_b200 import workloads as W
    from cirq_b200._cirq_compat import import_cirq

    cirq = import_cirq()
    circuit, qubits, names = W.qaoa_circuit(args.qubits)
    resolvers = list(cirq.to_resolvers(W.qaoa_sweep(names, args.resolvers)))
    if args.kind == 'dm':
        noise = cirq.depolarize(0.01)
        make = lambda **kw: cirq_b200.B200DensityMatrixSimulator(noise=noise, seed=0, **kw)
        ref = cirq.DensityMatrixSimulator(noise=noise, seed=0)
    else:
        make = lambda **kw: cirq_b200.B200Simulator(seed=0, **kw)
        ref = cirq.Simulator(seed=0)

    def timed(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        return out, time.perf_counter() - t0

    rep = dict(kind=args.kind, n_qubits=args.qubits, resolvers=args.resolvers, repetitions=args.reps,
               ops=len(list(circuit.all_operations())))
    if args.expect:
        # the QAOA cost: sum over the graph's edges of Z_i Z_j, on the circuit without its measurement
        body = cirq.Circuit(op for op in circuit.all_operations() if not cirq.is_measurement(op))
        edges = {tuple(sorted(op.qubits)) for op in body.all_operations() if len(op.qubits) == 2}
        cost = sum(cirq.Z(a) * cirq.Z(b) for a, b in sorted(edges))
        rep['mode'] = f'simulate_expectation_values_sweep, {len(edges)} ZZ terms'
        sim = make(sweep_batch=True)
        sim.simulate_expectation_values_sweep(body, [cost], resolvers[:4])
        got, dt = timed(lambda: sim.simulate_expectation_values_sweep(body, [cost], resolvers))
        assert sim.last_run_info.get('path') == 'batched sweep', sim.last_run_info
        rep['batched'] = dict(seconds=dt, resolvers_per_s=len(resolvers) / dt, **sim.last_run_info)
        seq = make()
        seq.simulate_expectation_values_sweep(body, [cost], resolvers[:2])
        want, dt = timed(lambda: seq.simulate_expectation_values_sweep(body, [cost], resolvers[: args.seq_resolvers]))
        rep['sequential'] = dict(resolvers=args.seq_resolvers, seconds=dt, resolvers_per_s=args.seq_resolvers / dt)
        rep['max_abs_diff_batched_vs_sequential'] = float(
            np.max(np.abs(np.asarray(got[: args.seq_resolvers]) - np.asarray(want))))
        if args.ref_resolvers:
            t0 = time.perf_counter()
            ref_vals = ref.simulate_expectation_values_sweep(body, [cost], resolvers[: args.ref_resolvers])
            dt = time.perf_counter() - t0
            rep['reference_cpu'] = dict(resolvers=args.ref_resolvers, seconds=dt,
                                        resolvers_per_s=args.ref_resolvers / dt, cores=1, host_cores=os.cpu_count())
            rep['max_abs_diff_vs_reference'] = float(
                np.max(np.abs(np.asarray(got[: args.ref_resolvers]) - np.asarray(ref_vals))))
            rep['speedup_vs_reference_cpu'] = rep['batched']['resolvers_per_s'] / rep['reference_cpu']['resolvers_per_s']
        rep['speedup_vs_sequential'] = rep['batched']['resolvers_per_s'] / rep['sequential']['resolvers_per_s']
        line = json.dumps(rep)
        print(line)
        if args.out:
            with open(args.out, 'w') as f:
                f.write(line + '\n')
        return
    sim = make(sweep_batch=True)
    sim.run_sweep(circuit, resolvers[:4], repetitions=10)  # warm-up
    res, dt = timed(lambda: sim.run_sweep(circuit, resolvers, repetitions=args.reps))
    assert sim.last_run_info.get('path') == 'batched sweep', sim.last_run_info
    rep['batched'] = dict(seconds=dt, resolvers_per_s=len(resolvers) / dt, **sim.last_run_info)
    mean_b = np.mean([r.measurements['m'].mean() for r in res[: args.seq_resolvers]])

    seq = make()
    seq.run_sweep(circuit, resolvers[:2], repetitions=10)
    res, dt = timed(lambda: seq.run_sweep(circuit, resolvers[: args.seq_resolvers], repetitions=args.reps))
    rep['sequential'] = dict(resolvers=args.seq_resolvers, seconds=dt, resolvers_per_s=args.seq_resolvers / dt)
    mean_s = np.mean([r.measurements['m'].mean() for r in res])
    rep['mean_bit_batched_vs_sequential'] = [float(mean_b), float(mean_s)]
    if args.ref_resolvers:
        t0 = time.perf_counter()
        ref.run_sweep(circuit, resolvers[: args.ref_resolvers], repetitions=args.reps)
        dt = time.perf_counter() - t0
        rep['reference_cpu'] = dict(resolvers=args.ref_resolvers, seconds=dt,
                                    resolvers_per_s=args.ref_resolvers / dt, cores=1,
                                    host_cores=os.cpu_count())
        rep['speedup_vs_reference_cpu'] = rep['batched']['resolvers_per_s'] / rep['reference_cpu']['resolvers_per_s']
    rep['speedup_vs_sequential'] = rep['batched']['resolvers_per_s'] / rep['sequential']['resolvers_per_s']
    line = json.dumps(rep)
    print(line)
    if args.out:
        with open(args.out, 'w') as f:
            f.write(line + '\n')


if __name__ == '__main__':
    main()
