"""A recorded schedule compiled for `b2q_run_schedule`: ONE library call executes
the device operations of a circuit's unitary prefix (basis states, Kronecker joins,
gate passes, in-place permutations, scalings — what the reference's
``SimulationProductState`` does to its sub-states,
cirq-core/cirq/sim/simulation_product_state.py:83-139, recorded by
cirq_b200/plan_cache.py or cirq_b200/plan.py).

Replaying the list from Python costs ~20 us of interpreter + ctypes per operation:
for a 20-qubit circuit (70 operations, 1.5 ms of kernels) that is most of the call.
Here the operations are lowered ONCE to the launches `DeviceState.apply_batch` would
make (`DeviceState.lower_batch`: the same grouping code), their arguments packed into
two flat arrays, and every state of the schedule carved out of one arena allocation.

Only for registers of at most MAX_BITS qubits: above that a replay is GPU time, and
states that die during the schedule should give their memory back as they do
(the arena keeps everything alive until the results are dropped).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from cirq_b200 import _lib

MAX_BITS = 26

OP_BASIS, OP_KRON, OP_DENSE, OP_TILE, OP_DIAGONAL, OP_SCALE, OP_PERMUTE = range(7)


class ScheduleOp(ctypes.Structure):
    """`b2q_schedule_op` of include/cirq_b200.h."""

    _fields_ = [
        ('kind', ctypes.c_int32), ('slot', ctypes.c_int32), ('a', ctypes.c_int32), ('b', ctypes.c_int32),
        ('n_bits', ctypes.c_int32), ('count', ctypes.c_int32), ('ints_offset', ctypes.c_int64),
        ('reals_offset', ctypes.c_int64), ('basis_index', ctypes.c_uint64),
    ]


def enabled() -> bool:
    return os.environ.get('CIRQ_B200_NATIVE_REPLAY', '1') != '0'


class _Shape:
    """What `DeviceState.lower_batch` needs to know about a state."""

    def __init__(self, n_bits, dtype, device_cls):
        self.n_bits = n_bits
        self.dtype = np.dtype(dtype)
        self.code = _lib.dtype_code(self.dtype)
        self._cls = device_cls
        self.TILE_MIN_BITS = device_cls.TILE_MIN_BITS

    def tile_pairing(self):
        return self._cls.tile_pairing(self)

    def _pairable(self, m, b):
        return self._cls._pairable(self, m, b)

    def plan_passes(self, gates):
        return self._cls.plan_passes(self, gates)

    def lower_batch(self, gates):
        return self._cls.lower_batch(self, gates)


class NativeSchedule:
    """Compiled form of a list of recorded device operations (see `compile_schedule`)."""

    def __init__(self, dtype, records, ints, reals, slot_bits, alive, passes, permutes=()):
        self.dtype = np.dtype(dtype)
        self.code = _lib.dtype_code(self.dtype)
        self.num_ops = len(records)
        self.ops = (ScheduleOp * max(1, len(records)))(*records)
        self.ints = np.ascontiguousarray(np.asarray(ints + [0], dtype=np.int32))
        self.reals = np.ascontiguousarray(np.asarray(reals + [0.0], dtype=np.float64))
        self.slot_bits = list(slot_bits)
        amp = 8 if self.code == _lib.C64 else 16
        self.offsets, total = [], 0
        for bits in self.slot_bits:
            self.offsets.append(total)
            total += (((amp << bits) + 255) // 256) * 256
        self.total_bytes = max(total, 256)
        self.amp_bytes = amp
        self.alive = dict(alive)  # ident -> slot of the states that exist at the end
        self.passes = dict(passes)  # ident -> gate passes issued on it (and its ancestors)
        self.permutes = list(permutes)  # (surviving ident, slot) of the in-place permutations

    def run(self, device_cls):
        """Executes the schedule on the current stream; ({ident: device state}, {ident:
        passes}) of the states that exist at the end."""
        import torch

        lib = _lib.load()
        arena = torch.empty(self.total_bytes, dtype=torch.uint8, device='cuda')
        base = arena.data_ptr()
        slots = (ctypes.c_void_p * max(1, len(self.offsets)))(*[base + off for off in self.offsets])
        permute_passes = (ctypes.c_int * max(1, len(self.offsets)))() if self.permutes else None
        _lib.check(lib.b2q_run_schedule(
            self.code, self.num_ops, self.ops, self.ints.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
            self.reals.ctypes.data, len(self.offsets), slots, permute_passes,
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        real = torch.float32 if self.code == _lib.C64 else torch.float64
        live = {}
        for ident, slot in self.alive.items():
            off, nbytes = self.offsets[slot], self.amp_bytes << self.slot_bits[slot]
            live[ident] = device_cls(self.slot_bits[slot], self.dtype,
                                     tensor=arena[off:off + nbytes].view(real).view(-1, 2))
        passes = dict(self.passes)
        for ident, slot in self.permutes:  # (how many tile passes a permutation took is the library's answer)
            passes[ident] += int(permute_passes[slot])
        return live, passes


def compile_schedule(ops, dtype, device_cls):
    """Lowers recorded operations — ('basis', id, bits, index), ('kron', id, a, b),
    ('apply', id, blocks[, passes]), ('scale', id, factor), ('permute', id, src_bit) —
    to a NativeSchedule; None when some operation has no place in one (a state above
    MAX_BITS, a dense block wider than the register-tiled kernels take)."""
    dtype = np.dtype(dtype)
    max_fast = 5 if _lib.dtype_code(dtype) == _lib.C64 else 4
    records, ints, reals = [], [], []
    slot_of, bits_of, alive, passes = {}, [], {}, {}
    permuted: list = []  # idents permuted in place (their pass counts come back from the library)
    heir: dict = {}  # ident -> the state it was joined into

    def new_slot(ident, bits):
        slot_of[ident] = len(bits_of)
        bits_of.append(bits)
        alive[ident] = slot_of[ident]
        passes[ident] = 0
        return slot_of[ident]

    def emit(kind, slot, n_bits, count=0, a=-1, b=-1, ivals=(), rvals=(), index=0):
        records.append(ScheduleOp(kind, slot, a, b, n_bits, count, len(ints), len(reals), index))
        ints.extend(int(v) for v in ivals)
        reals.extend(rvals)

    def matrix_reals(m):
        return _lib.as_c128_buffer(m).reshape(-1).view(np.float64).tolist()

    for op in ops:
        kind = op[0]
        if kind == 'basis':
            if op[2] > MAX_BITS:
                return None
            emit(OP_BASIS, new_slot(op[1], op[2]), op[2], index=int(op[3]))
        elif kind == 'kron':
            a, b = slot_of[op[2]], slot_of[op[3]]
            bits = bits_of[a] + bits_of[b]
            if bits > MAX_BITS:
                return None
            slot = new_slot(op[1], bits)
            passes[op[1]] = passes.pop(op[2]) + passes.pop(op[3])
            del alive[op[2]], alive[op[3]]
            heir[op[2]] = heir[op[3]] = op[1]
            emit(OP_KRON, slot, bits, a=a, b=b, ivals=(bits_of[a], bits_of[b]))
        elif kind == 'apply':
            slot = slot_of[op[1]]
            n = bits_of[slot]
            shape = _Shape(n, dtype, device_cls)
            blocks = list(op[2])
            passes[op[1]] += op[3] if len(op) > 3 else len(shape.plan_passes(blocks))
            for what, payload in shape.lower_batch(blocks):
                if what == 'diag':
                    m, wires = payload
                    emit(OP_DIAGONAL, slot, n, count=len(wires), ivals=wires, rvals=matrix_reals(m))
                    continue
                if any(len(w) > max_fast for _, w in payload):
                    return None  # (goes through the out-of-place kernel with a scratch buffer)
                ivals = [len(w) for _, w in payload] + [t for _, w in payload for t in w]
                rvals = [x for m, _ in payload for x in matrix_reals(m)]
                emit(OP_TILE if what == 'tile' else OP_DENSE, slot, n, count=len(payload),
                     ivals=ivals, rvals=rvals)
        elif kind == 'scale':
            slot = slot_of[op[1]]
            f = complex(op[2])
            emit(OP_SCALE, slot, bits_of[slot], rvals=(f.real, f.imag))
        elif kind == 'permute':
            slot = slot_of[op[1]]
            emit(OP_PERMUTE, slot, bits_of[slot], count=len(op[2]), ivals=op[2])
            permuted.append(op[1])
        else:
            return None
    permutes = []
    for ident in permuted:
        last = ident
        while last in heir:
            last = heir[last]
        permutes.append((last, slot_of[ident]))
    # (the library adds up the permutations of a slot itself: one entry per slot)
    return NativeSchedule(dtype, records, ints, reals, bits_of, alive, passes, list(dict.fromkeys(permutes)))
