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
import threading

from cirq_b200._cirq_compat import import_cirq

cirq = import_cirq()

from cirq.sim import density_matrix_simulator, mux, sparse_simulator  # noqa: E402

from cirq_b200.dm_simulator import B200DensityMatrixSimulator  # noqa: E402
from cirq_b200.sv_simulator import B200Simulator  # noqa: E402


_LOCK = threading.RLock()
_DEPTH = 0
_SAVED = None


@contextlib.contextmanager
def use_b200():
    """Within the block, Cirq's mux functions simulate on the B200.

    The reference looks its simulator classes up as module globals at call time
    (sim/mux.py:90, 163, 322), so the switch is a swap of those two globals.  It is
    re-entrant (nested blocks, and the functions below calling each other) and
    safe against concurrent use: a process-wide lock guards the swap and a depth
    counter restores the reference classes only when the LAST block exits, so a
    thread leaving its block never un-binds the classes under another thread that
    is still inside one.  (While any block is open, every thread's
    ``cirq.sample`` runs on the B200 — the globals are process-wide.)"""
    global _DEPTH, _SAVED
    with _LOCK:
        if _DEPTH == 0:
            _SAVED = (sparse_simulator.Simulator, density_matrix_simulator.DensityMatrixSimulator)
            sparse_simulator.Simulator = B200Simulator
            density_matrix_simulator.DensityMatrixSimulator = B200DensityMatrixSimulator
        _DEPTH += 1
    try:
        yield
    finally:
        with _LOCK:
            _DEPTH -= 1
            if _DEPTH == 0:
                sparse_simulator.Simulator, density_matrix_simulator.DensityMatrixSimulator = _SAVED
                _SAVED = None


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
