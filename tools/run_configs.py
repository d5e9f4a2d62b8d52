"""Runs the BASELINE configs that are parity-test cases rather than the bench
headline, at their full sizes, on one B200, with the size-independent checks:

  config 3: 34-qubit QFT (137 GB complex64 state, one GPU): analytic amplitudes
  config 5: 16-qubit noisy QAOA density matrix (34 GB rho), run_sweep

    python tools/run_configs.py [--qft 34] [--qaoa 16] [--resolvers 4] [--out file.json]
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
    ap.add_argument('--qft', type=int, default=34)
    ap.add_argument('--qaoa', type=int, default=16)
    ap.add_argument('--resolvers', type=int, default=4)
    ap.add_argument('--reps', type=int, default=1000)
    ap.add_argument('--out', default='')
    args = ap.parse_args()
    import torch

    import cirq_b200
    from cirq_b200 import workloads as W
    from cirq_b200._cirq_compat import import_cirq

    cirq = import_cirq()
    report = {}
    if args.qft:
        n = args.qft
        circuit, qubits = W.qft_circuit(n)
        sim = cirq_b200.B200Simulator(dtype=np.complex64)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = sim.simulate(circuit, qubit_order=qubits)
        dev = res.device_state
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        rng = np.random.RandomState(1)
        idx = rng.randint(0, 1 << n, size=256, dtype=np.int64)
        amps = dev.amplitudes(idx)
        err = float(np.max(np.abs(amps - 2.0 ** (-n / 2))))
        nrm = dev.norm2()
        passes = res._final_simulator_state._state.passes
        # second input: basis state x -> all amplitudes have modulus 2^-n/2
        del res, dev
        torch.cuda.empty_cache()
        x = int(rng.randint(1, 1 << n))
        res = sim.simulate(circuit, qubit_order=qubits, initial_state=x)
        amps = res.device_state.amplitudes(idx)
        mod_err = float(np.max(np.abs(np.abs(amps) - 2.0 ** (-n / 2))))
        nrm2 = res.device_state.norm2()
        report['qft'] = dict(n=n, ops=len(list(circuit.all_operations())), passes=passes, seconds=dt,
                             state_gb=(8 << n) / 1e9, max_abs_err_uniform=err, norm=nrm,
                             max_modulus_err_basis_input=mod_err, norm_basis_input=nrm2,
                             ms_per_pass=dt * 1e3 / passes,
                             GBps=2 * (8 << n) * passes / dt / 1e9)
        print('QFT', json.dumps(report['qft']), flush=True)
        assert err < 1e-7 and abs(nrm - 1) < 1e-3 and mod_err < 1e-7
        del res
        torch.cuda.empty_cache()
    if args.qaoa:
        n = args.qaoa
        circuit, qubits, names = W.qaoa_circuit(n)
        sweep = W.qaoa_sweep(names, 256)
        resolvers = list(cirq.to_resolvers(sweep))[: args.resolvers]
        sim = cirq_b200.B200DensityMatrixSimulator(noise=cirq.depolarize(0.01), seed=0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        results = sim.run_sweep(circuit, resolvers, repetitions=args.reps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        shapes = [r.measurements['m'].shape for r in results]
        # trace / hermiticity spot check on one resolver through simulate
        nomeas = cirq.Circuit(op for op in circuit.all_operations() if not cirq.is_measurement(op))
        t1 = time.perf_counter()
        # one dense rho from the start, so that `passes` counts 34 GB passes only
        dense = cirq_b200.B200DensityMatrixSimulator(noise=cirq.depolarize(0.01), seed=0,
                                                     split_untangled_states=False)
        fin = dense.simulate(nomeas, resolvers[0], qubit_order=qubits)
        tr = fin.device_state.dm_trace()
        torch.cuda.synchronize()
        dt1 = time.perf_counter() - t1
        passes = fin._final_simulator_state._state.passes
        report['qaoa'] = dict(n=n, rho_gb=(8 << (2 * n)) / 1e9, resolvers=len(resolvers),
                              repetitions=args.reps, seconds=dt, seconds_per_resolver=dt / len(resolvers),
                              shapes=[list(s) for s in shapes], trace=tr, passes_per_resolver=passes,
                              ms_per_pass=dt1 * 1e3 / passes,
                              GBps=2 * (8 << (2 * n)) * passes / dt1 / 1e9)
        print('QAOA', json.dumps(report['qaoa']), flush=True)
        assert abs(tr - 1) < 1e-3
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, 'w') as f:
            json.dump(report, f, indent=1)


if __name__ == '__main__':
    main()
