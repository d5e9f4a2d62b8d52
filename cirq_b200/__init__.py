"""cirq_b200 — a B200-native (sm_100a) simulator backend for Cirq.

``B200Simulator`` and ``B200DensityMatrixSimulator`` are drop-in replacements
for ``cirq.Simulator`` and ``cirq.DensityMatrixSimulator``; everything below
Cirq's simulator driver classes runs in hand-written CUDA kernels reached
through the C-ABI of ``libcirq_b200.so`` (``include/cirq_b200.h``).

Importing this package is cheap; the simulator classes (which import Cirq)
are loaded on first attribute access.
"""
from cirq_b200._lib import B200Error  # noqa: F401

__version__ = '0.1.0'

_LAZY = {
    'B200Simulator': 'cirq_b200.sv_simulator',
    'B200StateVectorSimulationState': 'cirq_b200.sv_simulator',
    'B200StateVectorTrialResult': 'cirq_b200.sv_simulator',
    'B200SimulatorStep': 'cirq_b200.sv_simulator',
    'B200DensityMatrixSimulator': 'cirq_b200.dm_simulator',
    'B200DensityMatrixSimulationState': 'cirq_b200.dm_simulator',
    'B200DensityMatrixTrialResult': 'cirq_b200.dm_simulator',
    'DeviceState': 'cirq_b200.device_state',
    'GateFuser': 'cirq_b200.fusion',
    'sample': 'cirq_b200.mux',
    'sample_sweep': 'cirq_b200.mux',
    'final_state_vector': 'cirq_b200.mux',
    'final_density_matrix': 'cirq_b200.mux',
    'use_b200': 'cirq_b200.mux',
    'B200ShardedSimulator': 'cirq_b200.dist',
    'ShardedStateVector': 'cirq_b200.dist',
    'run_sweep_sharded': 'cirq_b200.dist',
}


def __getattr__(name):
    if name in _LAZY:
        import importlib

        return getattr(importlib.import_module(_LAZY[name]), name)
    raise AttributeError(f'module {__name__!r} has no attribute {name!r}')
