"""cProfile of the end-to-end call bench.py times as `e2e`
(B200Simulator(seed=0).run on the 30-qubit Sycamore-style circuit, 1M samples):
where the host time goes next to the ~130 ms of device work.

    python tools/e2e_profile.py [--workload rqc30] [--top 30]
"""
import argparse
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='rqc30')
    ap.add_argument('--top', type=int, default=30)
    args = ap.parse_args()
    import torch

    import bench as B
    import cirq_b200
    from cirq_b200._cirq_compat import import_cirq

    cirq = import_cirq()
    wl = B.build_workload(args.workload)
    circuit, reps = wl['circuit'], wl['reps']
    if reps:
        circuit = circuit + cirq.Circuit(cirq.measure(*wl['qubits'], key='m'))

    def step():
        sim = cirq_b200.B200Simulator(dtype=np.complex64, seed=0)
        if reps:
            return sim.run(circuit, repetitions=reps).measurements['m'].shape
        return sim.compute_amplitudes(circuit, [0, 1], qubit_order=wl['qubits'])

    step()
    torch.cuda.synchronize()
    for _ in range(2):
        t0 = time.perf_counter()
        step()
        torch.cuda.synchronize()
        print('e2e step %.1f ms' % ((time.perf_counter() - t0) * 1e3), flush=True)
    pr = cProfile.Profile()
    pr.enable()
    step()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(args.top)


if __name__ == '__main__':
    main()
