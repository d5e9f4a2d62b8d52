"""Gate-list execution with lazily growing states (no Cirq objects involved).

``run_gate_list`` applies [(matrix, bits), ...] to |0...0> the way the
reference's ``SimulationProductState`` does for circuits
(cirq-core/cirq/sim/simulation_product_state.py:83-139): every qubit starts as
its own 1-qubit device state, a gate that couples two unentangled sets joins
them with the Kronecker-product kernel first, and gates are fused and applied
on whatever (small) state currently holds their qubits.  The 2^n-amplitude
state only comes into existence once the circuit has connected all qubits, so
the first cycles of a circuit cost almost no HBM traffic.  Used by ``bench.py``
for the device-resident measurement and usable as a Cirq-free entry point to
the library.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from cirq_b200.fusion import fuser_for


class _Component:
    __slots__ = ('bits', 'dev', 'fuser', 'passes', 'held')

    def __init__(self, bits, dev, fuser):
        self.bits = list(bits)  # logical bits, most significant first
        self.dev = dev
        self.fuser = fuser
        self.passes = 0
        self.held = []  # final blocks kept back by drain() to be paired later

    def wire(self, bit: int) -> int:
        return len(self.bits) - 1 - self.bits.index(bit)

    def flush(self) -> None:
        if len(self.fuser) or self.held:
            # relabelled SWAPs are not undone: the component renames its bits
            blocks = self.held + self.fuser.blocks(restore=False)
            self.held = []
            self.fuser.clear()
            self.dev.apply_batch(blocks)
            self.passes += len(blocks)
        perm = self.fuser.take_permutation()
        if perm:
            top = len(self.bits) - 1
            moved = list(self.bits)
            for w, now in perm.items():
                moved[top - now] = self.bits[top - w]
            self.bits = moved

    def drain(self) -> None:
        ready = self.held + self.fuser.pop_final_blocks()
        split = getattr(self.dev, 'split_unpaired_tail', None)
        # (a trailing block without a partner waits for the next batch: two blocks
        # share one pass over HBM, DeviceState.plan_passes)
        ready, self.held = split(ready) if split else (ready, [])
        if ready:
            self.dev.apply_batch(ready)
            self.passes += len(ready)


class SplitExecutor:
    """State of n logical bits kept as a product of device states."""

    # Joining sub-states is out of place (inputs + output live together), so above
    # this size the state is one dense tensor from the start.
    MAX_SPLIT_BITS = 30

    def __init__(self, n_bits: int, dtype=np.complex64, max_fused_qubits: int | None = None,
                 device_state_cls=None, max_component_bits: int | None = None):
        """`max_component_bits`: sub-states never grow beyond this many bits;
        `apply` then returns False for a gate that would need a larger join and
        leaves the caller to continue differently (the sharded path,
        cirq_b200/dist.py).  Implies product form from the start whatever n."""
        if device_state_cls is None:
            from cirq_b200.device_state import DeviceState as device_state_cls
        self._DS = device_state_cls
        self.n = int(n_bits)
        self.dtype = np.dtype(dtype)
        self.max_fused = max_fused_qubits
        self._comp = {}
        self.max_component_bits = max_component_bits
        if self.n > self.MAX_SPLIT_BITS and max_component_bits is None:
            bits = list(range(self.n - 1, -1, -1))
            c = _Component(bits, self._DS.basis(self.n, self.dtype, 0),
                           fuser_for(self.dtype, max_fused_qubits, self.n, state_vector=True))
            for b in bits:
                self._comp[b] = c
        else:
            for b in range(self.n):
                c = _Component([b], self._DS.basis(1, self.dtype, 0),
                               fuser_for(self.dtype, max_fused_qubits, 1, state_vector=True))
                self._comp[b] = c
        self.kron_count = 0
        self._since_drain = 0

    def _join(self, comps: Sequence[_Component]) -> _Component:
        first = comps[0]
        first.flush()
        dev, bits, passes = first.dev, list(first.bits), first.passes
        for other in comps[1:]:
            other.flush()
            dev = dev.kron(other.dev)
            bits += other.bits
            passes += other.passes
            self.kron_count += 1
        merged = _Component(bits, dev, fuser_for(self.dtype, self.max_fused, len(bits), state_vector=True))
        merged.passes = passes
        for b in bits:
            self._comp[b] = merged
        return merged

    def apply(self, matrix, bits: Sequence[int]) -> bool:
        comps = []
        for b in bits:
            c = self._comp[int(b)]
            if not any(c is x for x in comps):
                comps.append(c)
        if (self.max_component_bits is not None and len(comps) > 1
                and sum(len(c.bits) for c in comps) > self.max_component_bits):
            return False
        comp = comps[0] if len(comps) == 1 else self._join(comps)
        comp.fuser.add(matrix, [comp.wire(int(b)) for b in bits])
        self._since_drain += 1
        if self._since_drain >= max(8, self.n):
            self._since_drain = 0
            comp.drain()
        return True

    def components(self):
        """All sub-states, pending gates applied: [(DeviceState, logical bits
        most significant first)], largest first (ties: the one holding the
        highest bit)."""
        comps = []
        for b in range(self.n - 1, -1, -1):
            c = self._comp[b]
            if not any(c is x for x in comps):
                comps.append(c)
        for c in comps:
            c.flush()
        comps.sort(key=lambda c: -len(c.bits))
        return [(c.dev, list(c.bits)) for c in comps]

    def finalize(self):
        """Joins everything; returns (DeviceState, bit_of) with bit_of[logical bit]
        = index bit of that qubit in the merged state."""
        comps = []
        for b in range(self.n - 1, -1, -1):
            c = self._comp[b]
            if not any(c is x for x in comps):
                comps.append(c)
        merged = comps[0] if len(comps) == 1 else self._join(comps)
        merged.flush()
        bit_of = {b: merged.wire(b) for b in merged.bits}
        return merged.dev, bit_of, merged.passes


def run_gate_list(n_bits: int, gates, dtype=np.complex64, max_fused_qubits: int | None = None,
                  device_state_cls=None):
    """Applies `gates` to |0...0>; returns (DeviceState, bit_of, passes)."""
    ex = SplitExecutor(n_bits, dtype, max_fused_qubits, device_state_cls)
    for m, b in gates:
        ex.apply(m, b)
    return ex.finalize()


# ---- pre-built plans: schedule once on the host, replay on the device --------------------


class _RecordingState:
    """Stands in for DeviceState while a SplitExecutor schedules: records the
    device operations instead of performing them."""

    ops: list = []
    counter = 0

    def __init__(self, n_bits, ident, dtype=np.complex64):
        from cirq_b200 import _lib

        self.n_bits = n_bits
        self.ident = ident
        self.dtype = np.dtype(dtype)
        self.code = _lib.dtype_code(self.dtype)

    @classmethod
    def basis(cls, n_bits, dtype, index=0):
        cls.counter += 1
        cls.ops.append(('basis', cls.counter, n_bits, index))
        return cls(n_bits, cls.counter, dtype)

    def kron(self, other):
        cls = type(self)
        cls.counter += 1
        cls.ops.append(('kron', cls.counter, self.ident, other.ident))
        return cls(self.n_bits + other.n_bits, cls.counter, self.dtype)

    # the pass grouping of the device state the plan will be replayed on, so that a
    # recorded schedule holds back unpaired blocks exactly like a live one
    def tile_pairing(self):
        from cirq_b200.device_state import DeviceState

        return DeviceState.tile_pairing(self)

    TILE_MIN_BITS = 22

    def _pairable(self, m, b):
        return np.ndim(m) == 2 and len(b) <= 5

    def plan_passes(self, gates):
        from cirq_b200.device_state import DeviceState

        return DeviceState.plan_passes(self, gates)

    def split_unpaired_tail(self, gates):
        from cirq_b200.device_state import DeviceState

        return DeviceState.split_unpaired_tail(self, gates)

    def apply_batch(self, blocks):
        type(self).ops.append(('apply', self.ident, [(np.asarray(m), tuple(w)) for m, w in blocks]))


def build_plan(n_bits: int, gates, dtype=np.complex64, max_fused_qubits: int | None = None):
    """Schedules `gates` (fusion + lazy state growth) without touching the GPU.
    Returns a plan dict for `replay_plan`."""

    class Rec(_RecordingState):
        ops = []
        counter = 0

    ex = SplitExecutor(n_bits, dtype, max_fused_qubits, Rec)
    for m, b in gates:
        ex.apply(m, b)
    dev, bit_of, _ = ex.finalize()
    passes = sum(len(op[2]) for op in Rec.ops if op[0] == 'apply')
    full_passes = sum(len(op[2]) for op in Rec.ops if op[0] == 'apply' and True)
    return {'n': n_bits, 'dtype': np.dtype(dtype), 'ops': Rec.ops, 'final': dev.ident,
            'bit_of': bit_of, 'passes': passes, 'full_passes': full_passes}


def replay_ops(ops, dtype, device_state_cls=None, on_apply=None):
    """Executes recorded device operations; returns {ident: DeviceState} of the
    states still alive at the end."""
    if device_state_cls is None:
        from cirq_b200.device_state import DeviceState as device_state_cls
    live = {}
    for op in ops:
        if op[0] == 'basis':
            live[op[1]] = device_state_cls.basis(op[2], dtype, op[3])
        elif op[0] == 'kron':
            live[op[1]] = live.pop(op[2]).kron(live.pop(op[3]))
        else:
            if on_apply is None:
                live[op[1]].apply_batch(op[2])
            else:
                on_apply(live[op[1]], op[2])
    return live


def replay_plan(plan, device_state_cls=None, on_apply=None):
    """Executes a plan from `build_plan`; returns the final DeviceState.
    `on_apply(state, blocks)` replaces ``state.apply_batch(blocks)`` when given
    (bench.py times individual launches through it).  Small registers on the real
    device run as ONE library call (cirq_b200/program.py)."""
    if on_apply is None:
        from cirq_b200 import program
        from cirq_b200.device_state import DeviceState

        if (device_state_cls is None or device_state_cls is DeviceState) and program.enabled():
            if '_native' not in plan:
                plan['_native'] = program.compile_schedule(plan['ops'], plan['dtype'], DeviceState)
            if plan['_native'] is not None:
                return plan['_native'].run(DeviceState)[0][plan['final']]
    return replay_ops(plan['ops'], plan['dtype'], device_state_cls, on_apply)[plan['final']]
