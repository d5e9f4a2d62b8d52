"""Runs the REFERENCE's own simulator test files against the B200 drop-in
classes: `cirq.Simulator` / `cirq.DensityMatrixSimulator` are replaced by
`B200Simulator` / `B200DensityMatrixSimulator` before the reference test module
is imported (SURVEY.md §8c "reuse plan").

    python tests/ref_suite_runner.py {oracle|cuda} {sparse|density|mux} out.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import pytest  # noqa: E402


class Collector:
    def __init__(self):
        self.outcomes = {}

    def pytest_runtest_logreport(self, report):
        if report.when == 'call' or (report.when == 'setup' and report.outcome != 'passed'):
            self.outcomes[report.nodeid.split('::', 1)[1]] = report.outcome


def main():
    backend, which, out = sys.argv[1:4]
    from cirq_b200._cirq_compat import import_cirq

    cirq = import_cirq()
    import cirq_b200.dm_simulator as dmm
    import cirq_b200.sv_simulator as svm

    if backend == 'oracle':
        from fake_device import OracleDeviceState

        svm.DeviceState = OracleDeviceState
        dmm.DeviceState = OracleDeviceState
    cirq.Simulator = svm.B200Simulator
    cirq.DensityMatrixSimulator = dmm.B200DensityMatrixSimulator
    import cirq.sim as cs

    cs.Simulator = svm.B200Simulator
    cs.DensityMatrixSimulator = dmm.B200DensityMatrixSimulator
    ref_dir = os.path.join(os.path.dirname(cirq.__file__), 'sim')
    if which == 'mux' or which.endswith('.py'):
        # cirq.sample / final_state_vector / final_density_matrix look their simulator
        # classes up in these modules at call time (sim/mux.py:53-333)
        from cirq.sim import density_matrix_simulator, sparse_simulator

        sparse_simulator.Simulator = svm.B200Simulator
        density_matrix_simulator.DensityMatrixSimulator = dmm.B200DensityMatrixSimulator
    files = {'sparse': 'sparse_simulator_test.py', 'density': 'density_matrix_simulator_test.py',
             'mux': 'mux_test.py'}
    if which.endswith('.py'):
        # other reference test modules that build cirq.Simulator() /
        # cirq.DensityMatrixSimulator() themselves: comma-separated paths relative
        # to the cirq package
        targets = [os.path.join(os.path.dirname(cirq.__file__), w) for w in which.split(',')]
        ref_dir = os.path.dirname(cirq.__file__)
    else:
        targets = [os.path.join(ref_dir, files[which])]
    col = Collector()
    pytest.main([*targets, '-q', '-q', '-p', 'no:cacheprovider', '-c', os.devnull,
                 '--rootdir', ref_dir, '-W', 'ignore'], plugins=[col])
    with open(out, 'w') as f:
        json.dump(col.outcomes, f, indent=0)


if __name__ == '__main__':
    main()
