"""B200 versions of Cirq's convenience entry points.

``cirq.sample``, ``cirq.sample_sweep``, ``cirq.final_state_vector`` and
``cirq.final_density_matrix`` (cirq-core/cirq/sim/mux.py:53-333) pick a simulator
class by module lookup at call time.  The functions here run the reference's own
dispatch logic (Clifford shortcut, unitary check, measurement dephasing, partial
trace) with ``Simulator`` / ``DensityMatrixSimulator`` bound to the B200 classes,
so behaviour and signatures are the reference's by construction.
"""
from __future__ import annotations

import contextlib

from cirq_b200._cirq_compat import import_cirq

cirq = import_cirq()

from cirq.sim import density_matrix_simulator, mux, sparse_simulator  # noqa: E402

from cirq_b200.dm_simulator import B200DensityMatrixSimulator  # noqa: E402
from cirq_b200.sv_simulator import B200Simulator  # noqa: E402


@contextlib.contextmanager
def use_b200():
    """Within the block, Cirq's mux functions simulate on the B200."""
    saved = (sparse_simulator.Simulator, density_matrix_simulator.DensityMatrixSimulator)
    sparse_simulator.Simulator = B200Simulator
    density_matrix_simulator.DensityMatrixSimulator = B200DensityMatrixSimulator
    try:
        yield
    finally:
        sparse_simulator.Simulator, density_matrix_simulator.DensityMatrixSimulator = saved


def sample(program, **kwargs):
    """``cirq.sample`` (sim/mux.py:53-93) on the B200."""
    with use_b200():
        return mux.sample(program, **kwargs)


def sample_sweep(program, params, **kwargs):
    """``cirq.sample_sweep`` (sim/mux.py:173-213) on the B200."""
    with use_b200():
        return mux.sample_sweep(program, params, **kwargs)


def final_state_vector(program, **kwargs):
    """``cirq.final_state_vector`` (sim/mux.py:106-170) on the B200."""
    with use_b200():
        return mux.final_state_vector(program, **kwargs)


def final_density_matrix(program, **kwargs):
    """``cirq.final_density_matrix`` (sim/mux.py:216-333) on the B200."""
    with use_b200():
        return mux.final_density_matrix(program, **kwargs)
