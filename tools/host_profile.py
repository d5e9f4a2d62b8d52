"""Host-side cost of one end-to-end call with the device taken out: the device state
is replaced by one whose gate passes do nothing (results are meaningless), so what is
timed is Cirq's driver loop + the scheduler + the launch bookkeeping.  Runs without a GPU.

    python tools/host_profile.py [--workload rc20] [--top 25] [--call simulate|run]
"""
import argparse
import cProfile
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='rc20')
    ap.add_argument('--top', type=int, default=25)
    ap.add_argument('--call', default='simulate', choices=['simulate', 'run'])
    ap.add_argument('--sort', default='tottime')
    args = ap.parse_args()

    import bench as B
    import cirq_b200
    import cirq_b200.sv_simulator as svm
    from cirq_b200._cirq_compat import import_cirq
    from fake_device import OracleDeviceState

    # (patched in place: kron / copy of the emulation build plain OracleDeviceStates)
    for name in ('apply_matrix', 'apply_batch', 'apply_tile_blocks', 'apply_diagonal', 'permute_bits_inplace'):
        setattr(OracleDeviceState, name, lambda self, *a, **k: None)
    OracleDeviceState.tile_pairing = lambda self: self.n_bits >= 22
    svm.DeviceState = OracleDeviceState
    cirq = import_cirq()
    wl = B.build_workload(args.workload)
    circuit, qubits = wl['circuit'], wl['qubits']
    if args.call == 'run':
        circuit = circuit + cirq.Circuit(cirq.measure(*qubits, key='m'))

    def step():
        sim = cirq_b200.B200Simulator(dtype=np.complex64, seed=0)
        if args.call == 'run':
            return sim.run(circuit, repetitions=10)
        return sim.simulate(circuit, qubit_order=qubits)

    step()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        step()
        ts.append((time.perf_counter() - t0) * 1e3)
    print('host-only %s of %s: %s ms' % (args.call, args.workload, ' '.join('%.1f' % t for t in ts)), flush=True)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        step()
    pr.disable()
    pstats.Stats(pr).sort_stats(args.sort).print_stats(args.top)


if __name__ == '__main__':
    main()
