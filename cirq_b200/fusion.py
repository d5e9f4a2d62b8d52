"""Host-side gate scheduler: fuses a stream of small matrices into k-qubit blocks.

One GPU pass over the state costs the same HBM traffic whether it applies a
1-qubit or a 4-qubit matrix, so the scheduler's job is to minimise the number
of passes.  The reference's fuser (``cirq.merge_k_qubit_unitaries``,
cirq-core/cirq/transformers/merge_k_qubit_gates.py:70-114, driven by
``_merge_operations_impl`` transformer_primitives.py:442-546) only merges two
operations when the qubits of one are a subset of the other's (:452-455), so
it never grows a block beyond the widest gate.  This scheduler works on
(matrix, wires) pairs — `wires` are bit positions of the flat index — and
grows blocks up to ``max_qubits`` wires by a frontier rule:

* a block is *movable* when it is the last block on every one of its wires
  (nothing scheduled after it touches them), so it can be executed later;
* a new gate is merged with all the last-blocks on its wires when they are
  all movable and the union fits; otherwise into the latest of those blocks
  (always order-safe, see ``add``) plus whatever movable neighbours still
  fit; otherwise it opens a new block.

Matrix convention everywhere: wires[0] is the most significant bit of the
row/column index (the ``reshape((2,)*2k)`` convention of
protocols/apply_unitary_protocol.py:440-466).
"""
from __future__ import annotations

from typing import Iterable, Sequence

import ctypes
import os

import numpy as np


_NATIVE = None  # libcirq_b200.so's host helper, False when the library is not built


def _native():
    global _NATIVE
    if _NATIVE is None:
        _NATIVE = False
        if os.environ.get('CIRQ_B200_NATIVE_FUSER', '1') != '0':
            try:
                from cirq_b200 import _lib

                lib = _lib.load()
                if hasattr(lib, 'b2q_host_left_apply'):
                    _NATIVE = lib
            except Exception:  # library not built (host-only use of the scheduler): numpy path
                _NATIVE = False
    return _NATIVE


def _left_apply_native(lib, block: np.ndarray, union, matrix: np.ndarray, wires) -> np.ndarray | None:
    """(matrix on `wires`) @ block through b2q_host_left_apply; `block` is consumed."""
    u, k = len(union), len(wires)
    m = np.ascontiguousarray(matrix, dtype=np.complex128)
    top = u - 1
    pos = (ctypes.c_int * k)(*[top - union.index(w) for w in wires])
    if lib.b2q_host_left_apply(block.ctypes.data, u, m.ctypes.data, pos, k) != 0:
        return None
    return block


def expand_matrix(matrix: np.ndarray, wires: Sequence[int], out_wires: Sequence[int]) -> np.ndarray:
    """Embeds `matrix` on `wires` into the space of `out_wires` (a superset),
    acting as identity elsewhere."""
    k = len(wires)
    u = len(out_wires)
    m = np.asarray(matrix, dtype=np.complex128).reshape((2,) * (2 * k))
    if tuple(wires) == tuple(out_wires):
        return m.reshape(1 << k, 1 << k)
    lib = _native()
    if lib and u <= 6:
        out = _left_apply_native(lib, np.eye(1 << u, dtype=np.complex128), tuple(out_wires),
                                 m.reshape(1 << k, 1 << k), wires)
        if out is not None:
            return out
    pos = [out_wires.index(w) for w in wires]
    full = np.eye(1 << u, dtype=np.complex128).reshape((2,) * (2 * u))
    # contract the input legs of m with the row legs `pos` of the identity
    res = np.tensordot(m, full, axes=(list(range(k, 2 * k)), pos))
    # res legs: m's k output legs, then the identity's legs minus `pos`, in order
    rest = [i for i in range(2 * u) if i not in pos]
    order = [0] * (2 * u)
    for j, p in enumerate(pos):
        order[p] = j
    for j, r in enumerate(rest):
        order[r] = k + j
    return np.transpose(res, order).reshape(1 << u, 1 << u)


def apply_to_block(block: np.ndarray, union: Sequence[int], matrix: np.ndarray,
                   wires: Sequence[int]) -> np.ndarray:
    """(matrix on `wires`) @ block, with `block` a 2^u x 2^u matrix on `union`
    (a superset of `wires`): one tensor contraction on the row legs, no
    expansion of `matrix` to the full space."""
    u = len(union)
    k = len(wires)
    if k == u and tuple(wires) == tuple(union):
        return matrix @ block
    lib = _native()
    if lib and u <= 6:
        out = _left_apply_native(lib, np.array(block, dtype=np.complex128, order='C'), tuple(union),
                                 matrix, wires)
        if out is not None:
            return out
    pos = [union.index(w) for w in wires]
    t = block.reshape((2,) * u + (1 << u,))
    res = np.tensordot(matrix.reshape((2,) * (2 * k)), t, axes=(list(range(k, 2 * k)), pos))
    res = np.moveaxis(res, list(range(k)), pos)
    return res.reshape(1 << u, 1 << u)


def _materialize(members, union) -> np.ndarray:
    """Product of `members` = [(matrix, wires), ...] (applied in order) on the wires
    `union`: ONE native call (b2q_host_compose) for blocks of up to 6 wires."""
    u = len(union)
    if len(members) == 1 and tuple(members[0][1]) == tuple(union):
        return np.asarray(members[0][0], dtype=np.complex128)
    lib = _native()
    if lib and u <= 6 and hasattr(lib, 'b2q_host_compose'):
        top = u - 1
        ks = (ctypes.c_int * len(members))(*[len(w) for _, w in members])
        pos = [top - union.index(x) for _, w in members for x in w]
        bitpos = (ctypes.c_int * len(pos))(*pos)
        flat = np.concatenate([np.ascontiguousarray(m, dtype=np.complex128).reshape(-1) for m, _ in members])
        out = np.empty((1 << u, 1 << u), dtype=np.complex128)
        if lib.b2q_host_compose(out.ctypes.data, u, len(members), ks, bitpos, flat.ctypes.data) == 0:
            return out
    total = None
    for m, w in members:
        total = expand_matrix(m, w, union) if total is None else apply_to_block(total, union, m, w)
    return total


def _materialize_diag(members, union) -> np.ndarray:
    """Table of a diagonal block from its member gates [(diagonal entries, wires), ...]."""
    u = len(union)
    lib = _native()
    if lib and u <= 16 and hasattr(lib, 'b2q_host_compose_diag'):
        top = u - 1
        ks = (ctypes.c_int * len(members))(*[len(w) for _, w in members])
        pos = [top - union.index(x) for _, w in members for x in w]
        bitpos = (ctypes.c_int * len(pos))(*pos)
        flat = np.concatenate([np.ascontiguousarray(d, dtype=np.complex128).reshape(-1) for d, _ in members])
        out = np.empty(1 << u, dtype=np.complex128)
        if lib.b2q_host_compose_diag(out.ctypes.data, u, len(members), ks, bitpos, flat.ctypes.data) == 0:
            return out
    total = None
    for d, w in members:
        e = expand_diagonal(d, w, union)
        total = e if total is None else total * e
    return total


class _Block:
    """A fused block.  While the scheduler is still growing it, a dense block is
    only the LIST of its member gates (`members`); the 2^k x 2^k product is formed
    once, when somebody reads `matrix` (emission), instead of after every gate."""

    __slots__ = ('wires', '_matrix', 'members', 'seq', 'alive', 'count', 'diag')

    def __init__(self, wires, matrix, seq, count=1, diag=False, members=None):
        self.wires = tuple(wires)
        self._matrix = matrix  # 2^k x 2^k, or the 2^k diagonal entries when `diag`; None = lazy
        self.members = members  # [(matrix, wires), ...] in application order, when lazy
        self.seq = seq
        self.alive = True
        self.count = count
        self.diag = diag

    @property
    def matrix(self):
        if self._matrix is None:
            self._matrix = (_materialize_diag if self.diag else _materialize)(self.members, self.wires)
            self.members = None
        return self._matrix

    @matrix.setter
    def matrix(self, value):
        self._matrix = value
        self.members = None

    def dense(self) -> np.ndarray:
        return np.diag(self.matrix) if self.diag else self.matrix

    def parts(self) -> list:
        """[(matrix, wires), ...] whose ordered product is this block."""
        if self._matrix is None and not self.diag:
            return self.members
        return [(self.dense(), self.wires)]


_SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)


def is_diagonal(matrix: np.ndarray) -> bool:
    m = np.asarray(matrix)
    # (every non-zero entry sits on the diagonal)
    return m.ndim == 2 and np.count_nonzero(m) == np.count_nonzero(np.diagonal(m))


def _is_swap(m: np.ndarray) -> bool:
    """m == SWAP exactly (4 x 4); most gates fail the first comparison."""
    return bool(m[1, 2] == 1 and m[2, 1] == 1 and m[0, 0] == 1 and m[3, 3] == 1 and np.count_nonzero(m) == 4)


def expand_diagonal(diag: np.ndarray, wires: Sequence[int], out_wires: Sequence[int]) -> np.ndarray:
    """The 2^u diagonal of (diag on `wires`) embedded in `out_wires` (a superset)."""
    k, u = len(wires), len(out_wires)
    idx = np.arange(1 << u)
    sub = np.zeros(1 << u, dtype=np.int64)
    for q, w in enumerate(wires):
        pos = u - 1 - out_wires.index(w)  # bit of the big index holding wire w
        sub |= ((idx >> pos) & 1) << (k - 1 - q)
    return np.asarray(diag)[sub]


class GateFuser:
    """Accumulates gates and emits fused blocks in a valid execution order."""

    def __init__(self, max_qubits: int = 4, narrow_wires: Sequence[int] = (), narrow_max: int | None = None,
                 diag_max: int = 0, relabel_swaps: bool = False):
        """`narrow_wires`: wires whose presence caps a block at `narrow_max`
        qubits (index bits on which the widest kernel is inefficient).

        `diag_max` > max_qubits turns on diagonal blocks: diagonal gates that do
        not fit into a dense block with their predecessors are collected in
        diagonal blocks of up to `diag_max` wires (one table-lookup pass each);
        diagonal gates commute, so such a gate may join any diagonal block that
        has no dense block after it on the gate's wires.

        `relabel_swaps`: a SWAP gate moves no data — the two wires trade names
        (as SimulationProductState does for whole qubits,
        sim/simulation_product_state.py:95-108); `blocks()` restores the order
        with real swaps at the end unless the caller takes the permutation
        (`blocks(restore=False)` + `take_permutation()`)."""
        self.max_qubits = int(max_qubits)
        self.diag_max = int(diag_max) if diag_max and diag_max > max_qubits else 0
        self.relabel_swaps = bool(relabel_swaps)
        self._map: dict[int, int] = {}  # caller's wire -> wire currently holding it
        self._last_dense: dict[int, int] = {}  # wire -> seq of the last non-diagonal block on it
        self._narrow = frozenset(int(w) for w in narrow_wires)
        self._narrow_max = int(narrow_max) if narrow_max is not None else self.max_qubits
        self._blocks: list[_Block] = []
        self._last: dict[int, _Block] = {}
        self._seq = 0
        self.num_gates = 0

    def __len__(self) -> int:
        return self.num_gates

    @property
    def pending(self) -> bool:
        """Whether a flush has anything to do: queued gates, or SWAP gates that
        were relabelled (they are not counted as gates) and still have to be
        put back or handed over."""
        return bool(self.num_gates or self._map)

    def _fits(self, union: Sequence[int]) -> bool:
        if len(union) <= self._narrow_max:
            return True
        if len(union) > self.max_qubits:
            return False
        return self._narrow.isdisjoint(union)

    def _movable(self, b: _Block) -> bool:
        return all(self._last.get(w) is b for w in b.wires)

    def _new_seq(self) -> int:
        self._seq += 1
        return self._seq

    @staticmethod
    def _union(wire_lists: Iterable[Sequence[int]]) -> tuple[int, ...]:
        seen: dict[int, None] = {}
        for ws in wire_lists:
            for w in ws:
                seen.setdefault(w, None)
        return tuple(sorted(seen, reverse=True))

    def add(self, matrix: np.ndarray, wires: Sequence[int]) -> None:
        wires = tuple(int(w) for w in wires)
        k = len(wires)
        if not (type(matrix) is np.ndarray and matrix.dtype == np.complex128 and matrix.shape == (1 << k, 1 << k)):
            matrix = np.asarray(matrix, dtype=np.complex128).reshape(1 << k, 1 << k)
        if self.relabel_swaps:
            if k == 2 and _is_swap(matrix):
                a, b = wires
                self._map[a], self._map[b] = self._map.get(b, b), self._map.get(a, a)
                return
            if self._map:
                wires = tuple(self._map.get(w, w) for w in wires)
        self._add(matrix, wires)

    def _add(self, matrix: np.ndarray, wires: tuple[int, ...]) -> None:
        k = len(wires)
        self.num_gates += 1
        if k > self.max_qubits:
            # Too wide to fuse with anything: its own block at the end.
            self._append(_Block(wires, matrix, self._new_seq()))
            return
        diag_gate = bool(self.diag_max) and k >= 2 and is_diagonal(matrix)
        cands: list[_Block] = []
        for w in wires:
            b = self._last.get(w)
            if b is not None and b not in cands:
                cands.append(b)
        movable = [b for b in cands if self._movable(b)]

        # (a diagonal gate never drags a diagonal block into a dense one)
        diag_pred = diag_gate and any(b.diag for b in cands)

        # 1. everything on our wires can be pulled together at the end
        if cands and len(movable) == len(cands) and not diag_pred:
            union = self._union([wires] + [b.wires for b in cands])
            if self._fits(union):
                self._merge_at_end(cands, matrix, wires, union)
                return
        # 2. merge into the latest predecessor (order-safe: every other
        #    predecessor on our wires is earlier, and nothing after the latest
        #    touches any of our wires), plus movable neighbours that fit
        if cands:
            latest = max(cands, key=lambda b: b.seq)
            if len(latest.wires) <= self.max_qubits and not (diag_gate and latest.diag):
                union = self._union([latest.wires, wires])
                if self._fits(union):
                    extra = []
                    for b in sorted(movable, key=lambda b: len(b.wires)):
                        if b is latest:
                            continue
                        u2 = self._union([union, b.wires])
                        if self._fits(u2):
                            union = u2
                            extra.append(b)
                    self._merge_into(latest, extra, matrix, wires, union)
                    return
        # 2b. a diagonal gate that found no room in a dense block: into a diagonal
        #     block that no dense block follows on these wires (diagonals commute)
        if diag_gate and cands:
            self._add_diagonal(np.diagonal(matrix).copy(), wires)
            return
        # 3. new block, pulling in movable predecessors that fit
        union = tuple(sorted(wires, reverse=True))
        extra = []
        for b in sorted(movable, key=lambda b: len(b.wires)):
            u2 = self._union([union, b.wires])
            if self._fits(u2):
                union = u2
                extra.append(b)
        self._merge_at_end(extra, matrix, wires, union)

    def _append(self, block: _Block) -> None:
        self._blocks.append(block)
        for w in block.wires:
            self._last[w] = block
            if not block.diag:
                self._last_dense[w] = block.seq

    def _add_diagonal(self, diag: np.ndarray, wires: tuple[int, ...]) -> None:
        floor = max(self._last_dense.get(w, 0) for w in wires)
        best, best_overlap = None, 0
        for b in self._blocks:
            if not (b.alive and b.diag and b.seq > floor):
                continue
            overlap = len(set(b.wires) & set(wires))
            if overlap > best_overlap and len(set(b.wires) | set(wires)) <= self.diag_max:
                best, best_overlap = b, overlap
        if best is None:
            self._append(_Block(tuple(sorted(wires, reverse=True)), None, self._new_seq(), diag=True,
                                members=[(diag, tuple(wires))]))
            return
        # (the table over the union is formed once, when the block is emitted)
        members = best.members if best._matrix is None else [(best._matrix, best.wires)]
        best.members = members + [(diag, tuple(wires))]
        best._matrix = None
        best.wires = self._union([best.wires, wires])
        best.count += 1
        for w in wires:
            last = self._last.get(w)
            if last is None or last.seq < best.seq:
                self._last[w] = best

    def _compose(self, blocks: Sequence[_Block], matrix, wires, union) -> tuple[list, int]:
        """Members of G . (product of the given blocks, seq order): the product itself
        is formed when the block is emitted (`_Block.matrix`)."""
        members: list = []
        count = 1
        for b in sorted(blocks, key=lambda b: b.seq):
            members.extend(b.parts())
            count += b.count
        members.append((matrix, tuple(wires)))
        return members, count

    def _merge_at_end(self, blocks, matrix, wires, union) -> None:
        members, count = self._compose(blocks, matrix, wires, union)
        for b in blocks:
            b.alive = False
        self._append(_Block(union, None, self._new_seq(), count, members=members))

    def _merge_into(self, target: _Block, extra, matrix, wires, union) -> None:
        members, count = self._compose([target] + list(extra), matrix, wires, union)
        for b in extra:
            b.alive = False
        was_diag = target.diag
        target.wires = union
        target._matrix = None
        target.members = members
        target.count = count
        target.diag = False
        # Only the wires that gained an operation move their frontier here; the
        # target's other wires may already have later blocks.
        for w in wires:
            self._last[w] = target
        for b in extra:
            for w in b.wires:
                self._last[w] = target
        for w in (union if was_diag else [x for b in extra for x in b.wires] + list(wires)):
            self._last_dense[w] = max(self._last_dense.get(w, 0), target.seq)

    def blocks(self, restore: bool = True) -> list[tuple[np.ndarray, tuple[int, ...]]]:
        """Fused (matrix, wires) in execution order, independent narrow blocks
        packed side by side (`pack_disjoint`).  A 1-D `matrix` is a diagonal
        block (its 2^k diagonal entries).  With relabelled SWAPs pending,
        `restore` appends the real swaps that put every wire back in place."""
        if restore and self._map:
            self._restore_order()
        return pack_disjoint([(b.matrix, b.wires) for b in self._blocks if b.alive], self._fits)

    def _restore_order(self) -> None:
        """Real SWAP gates undoing the pending relabelling (cycle by cycle)."""
        holder = {w: self._map.get(w, w) for w in self._map}  # caller's wire -> current wire
        self._map = {}
        at = {cur: w for w, cur in holder.items()}  # current wire -> caller's wire living there
        for w in sorted(holder):
            cur = holder[w]
            if cur == w:
                continue
            # bring caller's wire w home: exchange the contents of wires `cur` and `w`
            self._add(_SWAP, (cur, w))
            other = at[w]  # the caller's wire that was living at position w
            at[cur], holder[other] = other, cur
            at[w], holder[w] = w, w

    def fold_permutation(self, where: Sequence[int]) -> None:
        """Merges a permutation the caller kept from earlier `take_permutation`
        calls (its wire b lives on wire where[b]) back into the pending relabelling,
        so that the next `blocks()` restores the caller's original order."""
        pending = self._map
        self._map = {b: pending.get(w, w) for b, w in enumerate(where) if pending.get(w, w) != b}

    def take_permutation(self) -> dict[int, int]:
        """{caller's wire: wire that holds it now} of the pending relabelling, which
        is then forgotten (the caller renames its wires instead of moving data)."""
        perm = {w: c for w, c in self._map.items() if w != c}
        self._map = {}
        return perm

    def num_blocks(self) -> int:
        return sum(1 for b in self._blocks if b.alive)

    def pop_final_blocks(self) -> list[tuple[np.ndarray, tuple[int, ...]]]:
        """Removes and returns the blocks that can no longer change.

        A block is final when it is not the last block on ANY of its wires:
        merging only ever targets last-blocks, so nothing will be added to it.
        Final blocks are released in sequence order as long as no block that
        stays behind precedes them on a shared wire; this lets the caller launch
        them on the GPU while the host keeps scheduling the rest of the circuit.
        """
        out = []
        kept = []
        blocked: set[int] = set()
        for b in self._blocks:
            if not b.alive:
                continue
            if b.diag:
                # may still grow while no dense block has closed one of its wires
                final = len(b.wires) >= self.diag_max or all(
                    self._last_dense.get(w, 0) > b.seq for w in b.wires)
            else:
                final = all(self._last.get(w) is not b for w in b.wires)
            if final and not any(w in blocked for w in b.wires):
                out.append((b.matrix, b.wires))
                self.num_gates -= b.count
            else:
                kept.append(b)
                blocked.update(b.wires)
        self._blocks = kept
        return pack_disjoint(out, self._fits)

    def clear(self) -> None:
        """Forgets the scheduled blocks (a pending SWAP relabelling stays)."""
        self._blocks = []
        self._last = {}
        self._last_dense = {}
        self.num_gates = 0


def pack_disjoint(blocks, fits) -> list[tuple[np.ndarray, tuple[int, ...]]]:
    """Packs blocks that act on disjoint wires into one wider block (Kronecker
    product), as long as `fits(wires)` allows the union.

    The frontier rule of `GateFuser.add` only joins gates that share a wire, so a
    layer of sixteen 1-qubit gates with nothing after it (the last layer of a
    circuit, or every layer of a noisy circuit, where each channel forces a
    flush) leaves sixteen blocks = sixteen passes.  Since a pass costs the same
    HBM traffic for any width, side-by-side blocks are merged: a block may move
    up to any position after the last earlier block it shares a wire with (all
    blocks in between act on other wires and commute with it), and joins the
    first group there that still has room."""
    groups: list[list] = []  # [wire set, [(matrix, wires), ...]]
    for m, ws in blocks:
        wset = set(ws)
        first = 0
        for gi in range(len(groups) - 1, -1, -1):
            if not groups[gi][0].isdisjoint(wset):
                first = gi + 1
                break
        diagonal = np.ndim(m) == 1  # diagonal blocks keep their place and their shape
        for gi in range(first, len(groups)):
            g = groups[gi]
            if diagonal or np.ndim(g[1][0][0]) == 1:
                continue
            if fits(tuple(g[0] | wset)) and fits(tuple(g[0])) and fits(tuple(ws)):
                g[0] |= wset
                g[1].append((m, ws))
                break
        else:
            groups.append([wset, [(m, ws)]])
    out = []
    for _, members in groups:
        if len(members) == 1:
            out.append(members[0])
            continue
        matrix = np.asarray(members[0][0])
        wires = tuple(members[0][1])
        for m, ws in members[1:]:
            k = len(ws)
            matrix = np.kron(matrix, np.asarray(m).reshape(1 << k, 1 << k))
            wires += tuple(ws)
        out.append((matrix, wires))
    return out


# widest diagonal block whose table the kernel keeps in shared memory (64 KB)
DIAG_MAX_WIRES = {np.dtype(np.complex64): 13, np.dtype(np.complex128): 12}


def fuser_for(dtype, max_qubits: int | None = None, n_bits: int | None = None,
              state_vector: bool = False) -> 'GateFuser':
    """The fusion policy matched to the kernels: complex64 fuses
    up to 5 wires (tensor-core kernels; blocks touching index bits 0-1 take the
    shared-memory-staged variant, so no wire needs a narrower cap any more —
    ``CIRQ_B200_NARROW_WIRES=0,1`` restores the old cap of 4 on those wires for
    experiments), except for states too small for that kernel; complex128 up to
    4.  `state_vector=True` adds diagonal blocks and SWAP relabelling (GateFuser).  (A 6-qubit tensor-core kernel exists and is selected with max_qubits=6,
    but at 5.2 ms per 30-qubit pass against 2.7 ms for 5 qubits it does not pay
    for the passes it saves.)"""
    is_c64 = np.dtype(dtype) == np.dtype(np.complex64)
    if max_qubits is None:
        max_qubits = 5 if is_c64 and (n_bits is None or n_bits >= 12) else 4
    extra = {}
    if state_vector and os.environ.get('CIRQ_B200_SMART_FUSION', '1') != '0':
        # pure-state schedules also get diagonal blocks and relabelled SWAPs
        # (density matrices fuse a gate with the noise that follows it into one
        # dense 4-bit block, which diagonal blocks would only break up)
        extra = dict(diag_max=DIAG_MAX_WIRES[np.dtype(dtype)], relabel_swaps=True)
    if is_c64 and max_qubits >= 5:
        narrow = os.environ.get('CIRQ_B200_NARROW_WIRES', '')
        wires = tuple(int(w) for w in narrow.split(',') if w.strip())
        return GateFuser(max_qubits, narrow_wires=wires, narrow_max=4, **extra)
    return GateFuser(max_qubits, **extra)


def fuse_gates(gates, max_qubits: int = 4, dtype=None, n_bits: int | None = None,
               diagonal_blocks: bool = False, permutation: dict | None = None):
    """Convenience wrapper: [(matrix, wires)] -> fused [(matrix, wires)].  With
    `dtype` the kernel-matched policy of `fuser_for` is used; `diagonal_blocks`
    adds diagonal blocks (1-D matrices).  With a dict as `permutation`, SWAP gates
    are relabelled instead of executed and the dict receives {wire: wire that
    holds its content afterwards} for the caller to rename its wires."""
    f = GateFuser(max_qubits) if dtype is None else fuser_for(dtype, max_qubits, n_bits)
    if diagonal_blocks and dtype is not None:
        f.diag_max = DIAG_MAX_WIRES[np.dtype(dtype)]
    if permutation is not None:
        f.relabel_swaps = True
        for m, w in gates:
            f.add(m, w)
        blocks = f.blocks(restore=False)
        permutation.update(f.take_permutation())
        return blocks
    for m, w in gates:
        f.add(m, w)
    return f.blocks()
