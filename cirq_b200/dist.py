"""Sharded state vector over P = 2^g GPUs of one node (one process per GPU).

Layout: the flat big-endian index of the n-qubit state is split
on its top g bits — physical bits n_local .. n-1 are *global* (their value is
the rank id), bits 0 .. n_local-1 are *local*; rank r holds the contiguous slice
``[r * 2^n_local, (r+1) * 2^n_local)`` as an ordinary ``DeviceState``.

* A fused block whose wires are all local runs the unchanged 1-GPU kernel on
  every rank, no communication.
* A block that is diagonal in its global wires needs no communication either:
  each rank applies the sub-matrix selected by its rank bits.
* Otherwise the needed global bits are exchanged with local bits
  (``swap_global_local``): one peer-memory kernel per rank over NVLink
  (``b2q_dist_swap_bit``), half a shard out and in per GPU.  Nothing is moved
  back: the logical->physical bit map is updated instead, and the victim
  local bit is the one whose next use lies furthest in the future.

torch.distributed (NCCL, or gloo in the CPU tests) is used for rendezvous,
barriers, IPC-handle exchange and scalar reductions only; the state exchange
itself is the library's own kernel.

The reference has no counterpart; the closest concept is the
index-only SWAP relabel of cirq-core/cirq/sim/simulation_product_state.py:95-108.
"""
from __future__ import annotations

import ctypes
from typing import Sequence

import numpy as np

from cirq_b200 import _lib
from cirq_b200._lib import check


# Shards up to this many bits (17 GB complex64) get a second buffer for the fused
# gate + exchange kernel.
FUSED_EXCHANGE_MAX_BITS = 31


def block_diagonal_in(matrix: np.ndarray, wires: Sequence[int], diag_wires: Sequence[int], atol=0.0):
    """If `matrix` (on `wires`, wires[0] = MSB) never changes the basis value of
    `diag_wires`, returns {values tuple -> sub-matrix on the remaining wires};
    else None."""
    k = len(wires)
    t = np.asarray(matrix).reshape((2,) * (2 * k))
    pos = [wires.index(w) for w in diag_wires]
    rest = [i for i in range(k) if i not in pos]
    subs = {}
    for vals in np.ndindex(*(2,) * len(pos)):
        # the off-diagonal (out != in) blocks must vanish
        idx = [slice(None)] * (2 * k)
        for p, v in zip(pos, vals):
            idx[p] = v
        row_fixed = t[tuple(idx)]  # legs: rest outs..., all ins (k)
        # now the in legs: position of in-leg for wire index i is (len(rest) + i)
        sel = [slice(None)] * row_fixed.ndim
        for p, v in zip(pos, vals):
            sel[len(rest) + p] = v
        diag_block = row_fixed[tuple(sel)]
        off = np.array(row_fixed, copy=True)
        off[tuple(sel)] = 0
        if np.sum(np.abs(off) ** 2) > atol:
            return None
        subs[tuple(int(v) for v in vals)] = diag_block.reshape(1 << len(rest), 1 << len(rest))
    return subs


class _RawShard:
    """Device memory from b2q_dist_alloc exposed to torch (zero copy)."""

    def __init__(self, ptr: int, n_elems: int, real_typestr: str):
        self.ptr = ptr
        self.__cuda_array_interface__ = {
            'shape': (n_elems, 2),
            'typestr': real_typestr,
            'data': (ptr, False),
            'version': 2,
        }


class ShardBackend:
    """CUDA backend: shard in IPC-shared HBM, exchange by peer-memory kernel."""

    def __init__(self, n_local: int, dtype, group=None):
        import torch
        import torch.distributed as dist

        from cirq_b200.device_state import DeviceState

        self.torch = torch
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.n_local = n_local
        self.dtype = np.dtype(dtype)
        self.lib = _lib.load()
        code = _lib.dtype_code(self.dtype)
        self._nbytes = (1 << n_local) * (8 if code == _lib.C64 else 16)
        # Buffer 0 is the shard.  A second, equally shared buffer makes the fused
        # gate + exchange kernel possible (it is out of place); it is only worth
        # its memory while shards are small next to the 180 GB of HBM.
        self.can_fuse_exchange = bool(code == _lib.C64 and 12 <= n_local <= FUSED_EXCHANGE_MAX_BITS)
        self._bufs, self._states, self._peers = [], [], []
        for _ in range(2 if self.can_fuse_exchange else 1):
            self._add_buffer()
        self._cur = 0
        self._token = torch.zeros(1, dtype=torch.int32, device='cuda')
        self.barrier()

    def _add_buffer(self) -> None:
        """Allocates one shard-sized buffer and maps every peer's (collective)."""
        from cirq_b200.device_state import DeviceState

        torch, dist = self.torch, self.dist
        code = _lib.dtype_code(self.dtype)
        ptr = ctypes.c_void_p()
        check(self.lib.b2q_dist_alloc(ctypes.c_uint64(self._nbytes), ctypes.byref(ptr)))
        raw = _RawShard(ptr.value, 1 << self.n_local, '<f4' if code == _lib.C64 else '<f8')
        tensor = torch.as_tensor(raw, device='cuda')
        state = DeviceState(self.n_local, self.dtype, tensor=tensor)
        state._raw_owner = raw
        handle = (ctypes.c_ubyte * 64)()
        check(self.lib.b2q_dist_ipc_get(ctypes.c_void_p(ptr.value), handle))
        mine = torch.tensor(list(handle), dtype=torch.uint8, device='cuda')
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine, group=self.group)
        peers = {}
        for r, h in enumerate(gathered):
            if r == self.rank:
                continue
            buf = (ctypes.c_ubyte * 64)(*h.cpu().tolist())
            out = ctypes.c_void_p()
            check(self.lib.b2q_dist_ipc_open(buf, ctypes.byref(out)))
            peers[r] = out.value
        self._bufs.append(ptr.value)
        self._states.append(state)
        self._peers.append(peers)

    @property
    def local(self):
        """The buffer that currently holds this rank's shard."""
        return self._states[self._cur]

    @property
    def _ptr(self):
        return self._bufs[self._cur]

    @property
    def peer_ptrs(self):
        return self._peers[self._cur]

    def barrier(self):
        """Host barrier: this rank's stream is drained, then all ranks meet."""
        self.torch.cuda.current_stream().synchronize()
        self.dist.barrier(group=self.group)

    def device_barrier(self):
        """Stream-ordered barrier: kernels enqueued after it on ANY rank start only
        once every rank's kernels enqueued before it have finished.  A 4-byte NCCL
        all-reduce; the host does not wait, so the scheduler keeps running ahead."""
        self.dist.all_reduce(self._token, group=self.group)

    def swap_bit(self, partner: int, local_bit: int, my_gbit: int) -> None:
        """Both ranks of every pair call this; the exchange kernel touches the
        partner's shard, so it sits between two stream-ordered barriers."""
        self.device_barrier()
        stream = ctypes.c_void_p(self.torch.cuda.current_stream().cuda_stream)
        check(
            self.lib.b2q_dist_swap_bit(
                ctypes.c_void_p(self._ptr), ctypes.c_void_p(self.peer_ptrs[partner]),
                self.local.code, self.n_local, local_bit, my_gbit, stream,
            )
        )
        self.device_barrier()

    def swap_bits(self, pairs: Sequence[tuple[int, int]]) -> None:
        """Multi-bit exchange, one kernel per rank: `pairs` = [(rank bit index,
        local bit)], ascending in the local bit.  All ranks call it."""
        m = len(pairs)
        rho = 0
        for i, (gi, _) in enumerate(pairs):
            rho |= ((self.rank >> gi) & 1) << i
        base = self.rank
        for gi, _ in pairs:
            base &= ~(1 << gi)
        peers = (ctypes.c_void_p * (1 << m))()
        for c in range(1 << m):
            r = base
            for i, (gi, _) in enumerate(pairs):
                r |= ((c >> i) & 1) << gi
            peers[c] = None if r == self.rank else self.peer_ptrs[r]
        self.device_barrier()
        stream = ctypes.c_void_p(self.torch.cuda.current_stream().cuda_stream)
        check(
            self.lib.b2q_dist_swap_bits(
                ctypes.c_void_p(self._ptr), peers, self.local.code, self.n_local,
                _lib.int_array([l for _, l in pairs]), m, rho, stream,
            )
        )
        self.device_barrier()

    def apply_exchange(self, matrix, bits: Sequence[int], partner: int, local_bit: int,
                       my_gbit: int) -> None:
        """One 4/5-qubit block and the exchange of `local_bit` with the pair's
        global bit in a single kernel: reads the current buffer, writes the spare
        ones (its own and the partner's), which then become current."""
        other = 1 - self._cur
        m = np.ascontiguousarray(matrix, dtype=np.complex128)
        self.device_barrier()
        stream = ctypes.c_void_p(self.torch.cuda.current_stream().cuda_stream)
        check(
            self.lib.b2q_dist_apply_exchange(
                ctypes.c_void_p(self._bufs[self._cur]), ctypes.c_void_p(self._bufs[other]),
                ctypes.c_void_p(self._peers[other][partner]), self.local.code, self.n_local,
                m.ctypes.data, _lib.int_array(list(bits)), len(bits), local_bit, my_gbit, stream,
            )
        )
        self.device_barrier()
        self._cur = other

    def all_reduce_sum(self, value: float) -> float:
        t = self.torch.tensor([value], dtype=self.torch.float64, device='cuda')
        self.dist.all_reduce(t, group=self.group)
        return float(t.item())

    def all_gather_floats(self, value: float) -> list[float]:
        t = self.torch.tensor([value], dtype=self.torch.float64, device='cuda')
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t, group=self.group)
        return [float(x.item()) for x in out]

    def gather_objects(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def merge_samples(self, reps: int, positions: np.ndarray, local_uniforms: np.ndarray,
                      bits: Sequence[int]) -> np.ndarray:
        """Resolves this rank's samples on the device, merges the physical
        indices of all ranks with one all-reduce (disjoint positions) and
        extracts the requested physical bits: uint8[reps, len(bits)]."""
        from cirq_b200.device_state import DeviceState

        torch = self.torch
        full = torch.zeros(max(reps, 1), dtype=torch.int64, device='cuda')
        if positions.size:
            idx = self.local.sample_indices_device(local_uniforms)
            idx = idx | (self.rank << self.n_local)
            full[torch.from_numpy(positions).to('cuda', non_blocking=True)] = idx
        self.dist.all_reduce(full, group=self.group)
        dev = DeviceState.unpack_bits_device(full[:reps], bits)
        host = torch.empty(dev.shape, dtype=dev.dtype, pin_memory=True)
        host.copy_(dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host.numpy()

    def close(self):
        self.barrier()
        for peers in self._peers:
            for p in peers.values():
                self.lib.b2q_dist_ipc_close(ctypes.c_void_p(p))
        self._peers = [{} for _ in self._peers]
        self.barrier()
        self._states = []
        for ptr in self._bufs:
            self.lib.b2q_dist_free(ctypes.c_void_p(ptr))
        self._bufs = []


class ExchangePlanner:
    """Plans the global<->local exchanges of a block list on the wires alone.

    `metas[i]` = (wires, diagonal wires, fusable) of block i.  A block runs without
    communication when every wire it is NOT diagonal in is local.  When the first
    block in line needs global wires, they are exchanged with the local bits whose
    next use lies furthest away; further global wires that later blocks need before
    those evicted bits are used again may join the same exchange (m bits at once
    move 1 - 2^-m of a shard, m single swaps m/2).  How many join is decided per
    exchange by rolling the rest of the schedule out under both simple policies
    (none / all that qualify) and keeping the cheapest total.

    Cost unit: one shard volume over NVLink.  A single-bit swap that rides along
    with the preceding local pass (fused kernel) hides that pass: FUSE_GAIN.
    """

    FUSE_GAIN = 0.13  # measured: a fused pass + swap costs 4.7 ms more than the pass, a bare swap 6.3 ms (30-bit shards)

    def __init__(self, metas, n_local: int, lowest_victim: int, can_fuse: bool, max_bits: int):
        self.metas = metas
        self.n_local = n_local
        self.lowest_victim = lowest_victim
        self.preferred_lowest = lowest_victim + 3  # 128-byte runs (complex64: bit 4, complex128: bit 3)
        self.can_fuse = can_fuse
        self.max_bits = max_bits

    def _needs(self, i: int, phys) -> list[int]:
        ws, diag, _ = self.metas[i]
        return [w for w in ws if phys[w] >= self.n_local and w not in diag]

    def _drain_local(self, remaining, phys):
        """(blocks runnable now, in order; blocks left): commuting blocks overtake."""
        pending = []
        progressed = True
        while progressed and remaining:
            progressed = False
            blocked: set[int] = set()
            keep = []
            for i in remaining:
                ws = self.metas[i][0]
                if blocked.isdisjoint(ws) and not self._needs(i, phys):
                    pending.append(i)
                    progressed = True
                    continue
                keep.append(i)
                blocked.update(ws)
            remaining = keep
        return pending, remaining

    def _victim(self, remaining, phys, protected):
        next_use = {}
        for pos, i in enumerate(remaining):
            for w in self.metas[i][0]:
                p = phys[w]
                if p < self.n_local and p not in next_use:
                    next_use[p] = pos
        # Bits below `preferred_lowest` are evicted only when nothing else is free: with
        # the exchanged bit at position b the shard crosses NVLink in runs of 2^b
        # amplitudes, and runs under 128 bytes waste the link (bits 1-2 as victims made a
        # 3-bit exchange of an 8.6 GB shard take 17 ms instead of 11, profiles r2i).
        for floor in (max(self.lowest_victim, self.preferred_lowest), self.lowest_victim):
            best, best_pos = None, -1
            for p in range(self.n_local - 1, floor - 1, -1):
                if p in protected:
                    continue
                pos = next_use.get(p, 1 << 60)
                if pos > best_pos:
                    best, best_pos = p, pos
            if best is not None:
                return best, best_pos
        raise RuntimeError('no local bit available to swap with')

    def _options(self, remaining, phys):
        """(forced pairs, optional pairs in order of first use)."""
        first = remaining[0]
        needed = self._needs(first, phys)
        protected = {phys[w] for w in self.metas[first][0]}
        forced = []
        for w in needed:
            victim, _ = self._victim(remaining, phys, protected)
            protected.add(victim)
            forced.append((phys[w], victim))
        optional = []
        if self.max_bits > len(forced):
            first_use = {}
            for pos, i in enumerate(remaining):
                for w in self._needs(i, phys):
                    first_use.setdefault(w, pos)
            cands = sorted((pos, w) for w, pos in first_use.items() if w not in needed)
            for pos, w in cands:
                if len(forced) + len(optional) >= self.max_bits:
                    break
                try:
                    victim, victim_pos = self._victim(remaining, phys, protected)
                except RuntimeError:
                    break
                if victim_pos <= pos:
                    break  # the evicted bit would be needed first: no gain
                protected.add(victim)
                optional.append((phys[w], victim))
        return forced, optional

    def _cost(self, pairs, pending) -> tuple[float, bool]:
        fuse = bool(self.can_fuse and pending and self.metas[pending[-1]][2] and pairs[0][1] >= 1)
        if len(pairs) == 1 or self.max_bits == 1:
            # single swaps (the first may ride along with the last local pass)
            return 0.5 * len(pairs) - (self.FUSE_GAIN if fuse else 0.0), fuse
        return 1.0 - 0.5 ** len(pairs), False

    @staticmethod
    def _swapped(phys, pairs):
        phys = list(phys)
        for g, l in pairs:
            for w in range(len(phys)):
                if phys[w] == g:
                    phys[w] = l
                elif phys[w] == l:
                    phys[w] = g
        return phys

    def plan(self, phys, mode: str = 'search', remaining=None):
        """(actions, cost).  actions: ('run', [block indices]) |
        ('exchange', [(global phys bit, local phys bit), ...], fused block index | None).
        mode: 'none' = only the wires a block needs, 'greedy' = every qualifying
        wire joins, 'search' = per exchange, the count with the cheapest rollout."""
        phys = list(phys)
        remaining = list(range(len(self.metas))) if remaining is None else list(remaining)
        actions = []
        total = 0.0
        while remaining:
            pending, remaining = self._drain_local(remaining, phys)
            if not remaining:
                if pending:
                    actions.append(('run', pending))
                break
            forced, optional = self._options(remaining, phys)
            if mode == 'none' or not optional:
                take = 0
            elif mode == 'greedy':
                take = len(optional)
            else:
                best = None
                for j in range(len(optional) + 1):
                    pairs = forced + optional[:j]
                    now, _ = self._cost(pairs, pending)
                    after = self._swapped(phys, pairs)
                    rest = min(self.plan(after, 'none', remaining)[1], self.plan(after, 'greedy', remaining)[1])
                    if best is None or now + rest < best[0] - 1e-9:
                        best = (now + rest, j)
                take = best[1]
            pairs = forced + optional[:take]
            cost, fuse = self._cost(pairs, pending)
            total += cost
            if fuse:
                if pending[:-1]:
                    actions.append(('run', pending[:-1]))
                actions.append(('exchange', pairs, pending[-1]))
            else:
                if pending:
                    actions.append(('run', pending))
                actions.append(('exchange', pairs, None))
            phys = self._swapped(phys, pairs)
        return actions, total


class ShardedStateVector:
    """n-qubit state sharded over the ranks of a process group.

    All ranks must call every method collectively with identical arguments
    (SPMD), exactly like a torch.distributed collective.
    """

    def __init__(self, n_qubits: int, dtype=np.complex64, *, backend=None, group=None,
                 initial_index: int | None = 0):
        self._owns_backend = backend is None  # (a caller's backend outlives this state)
        if backend is None:
            import torch.distributed as dist

            world = dist.get_world_size(group)
            g = world.bit_length() - 1
            if 1 << g != world:
                raise ValueError(f'world size {world} is not a power of two')
            backend = ShardBackend(n_qubits - g, dtype, group)
        self.backend = backend
        self.rank = backend.rank
        self.world = backend.world
        self.g = self.world.bit_length() - 1
        if 1 << self.g != self.world:
            raise ValueError(f'world size {self.world} is not a power of two')
        self.n = int(n_qubits)
        self.n_local = self.n - self.g
        if self.n_local < 2:
            raise ValueError('need at least 2 local qubits per rank')
        assert backend.n_local == self.n_local
        self.dtype = np.dtype(dtype)
        # logical bit -> physical bit (physical bits >= n_local are global)
        self.phys = list(range(self.n))
        self.swaps = 0
        self.exchanges = 0        # exchange kernels (a multi-bit exchange counts once)
        self.exchange_volume = 0.0  # shard fractions sent per rank, summed
        # bench.py's instrumentation: on_kernel(kind, block, run) must call run()
        self.on_kernel = None
        self.fused_exchanges = 0
        self.passes = 0
        self.local_only_blocks = 0
        self.diag_global_blocks = 0
        import os

        self.multi_bit_exchange = os.environ.get('CIRQ_B200_MULTI_BIT_EXCHANGE', '1') != '0'
        if initial_index is not None:
            self._init_basis(initial_index)

    @property
    def local(self):
        """This rank's shard (the backend may move it between two buffers)."""
        return self.backend.local

    # ------------------------------------------------------------------ helpers

    def load_product(self, components) -> None:
        """Sets the state to the Kronecker product of sub-states that every rank
        holds in full: `components` = [(device state, logical bits most
        significant first)], together covering all n bits.  The first component
        supplies the global (rank) bits: each rank takes its slice of it and
        joins the rest locally, writing straight into its shard — the
        distributed form of SimulationProductState.create_merged_state
        (cirq-core/cirq/sim/simulation_product_state.py:68-81), without any
        exchange."""
        comps = [(dev, list(bits)) for dev, bits in components]
        if sorted(b for _, bits in comps for b in bits) != list(range(self.n)):
            raise ValueError('components must cover every bit exactly once')
        # the leading component must be wider than the rank bits
        while len(comps[0][1]) <= self.g and len(comps) > 1:
            (d0, b0), (d1, b1) = comps[0], comps[1]
            comps[:2] = [(d0.kron(d1), b0 + b1)]
        dev0, bits0 = comps[0]
        if self.g:
            keep = len(bits0) - self.g
            acc = dev0.slice_copy(self.rank << keep, keep)
        else:
            acc = dev0
        rest = comps[1:]
        if not rest:
            acc.copy_into(self.local)
        else:
            for dev, _ in rest[:-1]:
                acc = acc.kron(dev)
            acc.kron_into(rest[-1][0], self.local)
        order = [b for _, bits in comps for b in bits]  # most significant first
        for i, b in enumerate(order):
            self.phys[b] = self.n - 1 - i
        getattr(self.backend, 'device_barrier', self.backend.barrier)()

    def _init_basis(self, index: int) -> None:
        owner = index >> self.n_local
        lib = _lib.load()
        if hasattr(self.local, 'ptr'):
            import torch

            stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            if owner == self.rank:
                check(lib.b2q_sv_init_basis(self.local.ptr, self.local.code, self.n_local,
                                            ctypes.c_uint64(index & ((1 << self.n_local) - 1)), stream))
            else:
                self.local.tensor.zero_()
        else:  # test backend
            self.local.array[:] = 0
            if owner == self.rank:
                self.local.array[index & ((1 << self.n_local) - 1)] = 1
        getattr(self.backend, 'device_barrier', self.backend.barrier)()

    def _rank_bit(self, phys_bit: int) -> int:
        return (self.rank >> (phys_bit - self.n_local)) & 1

    def rename_bits(self, permutation: dict) -> None:
        """Applies a SWAP relabelling from the scheduler: the content of logical
        bit w now lives where logical bit permutation[w] used to be."""
        if permutation:
            old = list(self.phys)
            for w, now in permutation.items():
                self.phys[w] = old[now]

    def swap_global_local(self, global_phys: int, local_phys: int, fused_block=None) -> None:
        """Exchanges physical bits (global, local) of the index: data moves, the
        logical->physical map follows.  `fused_block` = (matrix, local bits) is
        applied first, inside the exchange kernel."""
        gi = global_phys - self.n_local
        partner = self.rank ^ (1 << gi)
        if fused_block is not None:
            run = lambda: self.backend.apply_exchange(fused_block[0], fused_block[1], partner, local_phys,
                                                      self._rank_bit(global_phys))
            if self.on_kernel is None:
                run()
            else:
                self.on_kernel('pass+exchange', fused_block, run)
            self.passes += 1
            self.fused_exchanges += 1
        else:
            run = lambda: self.backend.swap_bit(partner, local_phys, self._rank_bit(global_phys))
            if self.on_kernel is None:
                run()
            else:
                self.on_kernel('exchange', None, run)
        for l in range(self.n):
            if self.phys[l] == global_phys:
                self.phys[l] = local_phys
            elif self.phys[l] == local_phys:
                self.phys[l] = global_phys
        self.swaps += 1
        self.exchanges += 1
        self.exchange_volume += 0.5

    def exchange_bits(self, pairs: Sequence[tuple[int, int]]) -> None:
        """Exchanges several (global physical bit, local physical bit) pairs with ONE
        kernel per rank: (1 - 2^-m) of the shard travels instead of m halves."""
        pairs = sorted(pairs, key=lambda gl: gl[1])
        if len(pairs) == 1 or not hasattr(self.backend, 'swap_bits'):
            for g, l in pairs:
                self.swap_global_local(g, l)
            return
        run = lambda: self.backend.swap_bits([(g - self.n_local, l) for g, l in pairs])
        if self.on_kernel is None:
            run()
        else:
            self.on_kernel('exchange%d' % len(pairs), None, run)
        for g, l in pairs:
            for w in range(self.n):
                if self.phys[w] == g:
                    self.phys[w] = l
                elif self.phys[w] == l:
                    self.phys[w] = g
        self.swaps += len(pairs)
        self.exchanges += 1
        self.exchange_volume += 1.0 - 0.5 ** len(pairs)

    # ------------------------------------------------------------------ gates

    def apply_blocks(self, blocks: Sequence[tuple[np.ndarray, Sequence[int]]]) -> None:
        """Applies fused blocks [(matrix, logical bits)], reordering commuting
        blocks so that everything executable without communication runs before
        the next exchange.  The exchanges themselves are planned first, on the
        wires alone (`ExchangePlanner`): which global bits to bring in together and
        which local bits to evict."""
        blocks = [(np.asarray(m), tuple(int(w) for w in ws)) for m, ws in blocks]
        self._diag_cache = {}
        metas = [(ws, self._diagonal_wires(m, ws), np.ndim(m) == 2 and len(ws) <= 5) for m, ws in blocks]
        can_multi = self.multi_bit_exchange and hasattr(self.backend, 'swap_bits') and self.g > 1
        planner = ExchangePlanner(
            metas, self.n_local, lowest_victim=1 if self.dtype == np.dtype(np.complex64) else 0,
            can_fuse=bool(getattr(self.backend, 'can_fuse_exchange', False)),
            max_bits=min(3, self.g) if can_multi else 1)
        actions, _ = planner.plan(self.phys, 'search' if can_multi else 'none')
        for action in actions:
            if action[0] == 'run':
                forms = []
                for i in action[1]:
                    form = self._local_form(*blocks[i])
                    assert form is not None, 'planner scheduled a block that needs an exchange'
                    if form[1]:
                        forms.append(form)
                self._run_local(forms)
                continue
            _, pairs, fused_index = action
            if fused_index is not None:
                form = self._local_form(*blocks[fused_index])
                fused = self._fusable([form], pairs[0][1]) if form is not None and form[1] else None
                if fused is None:  # (became a pure phase on this rank: nothing to fuse)
                    self.swap_global_local(*pairs[0])
                else:
                    self.swap_global_local(pairs[0][0], pairs[0][1], fused_block=fused)
                for gbit, victim in pairs[1:]:
                    self.swap_global_local(gbit, victim)
            elif len(pairs) == 1 or not can_multi:
                for gbit, victim in pairs:
                    self.swap_global_local(gbit, victim)
            else:
                self.exchange_bits(pairs)

    def _diagonal_wires(self, m: np.ndarray, ws: tuple[int, ...]) -> frozenset:
        """Wires whose basis value the block never changes: with those global the
        block still runs without communication (each rank applies the sub-matrix
        selected by its rank bits)."""
        if np.ndim(m) == 1:
            return frozenset(ws)
        if all(self.phys[w] < self.n_local for w in ws) and self.g == 0:
            return frozenset()
        return frozenset(w for w in ws if block_diagonal_in(m, list(ws), [w], atol=1e-24) is not None)

    def _run_local(self, batch) -> None:
        if batch:
            if self.on_kernel is None:
                self.local.apply_batch(batch)
            else:
                # instrumented (bench.py): one call per pass, bracketed by the hook
                for group in self.local.plan_passes(batch):
                    self.on_kernel('pass', group, lambda group=group: self.local.apply_batch(group))
            self.passes += len(self.local.plan_passes(batch)) if hasattr(self.local, 'plan_passes') else len(batch)

    def _fusable(self, pending, victim: int):
        """(matrix, bits) of the last pending pass widened to the 4/5 qubits the
        fused gate + exchange kernel takes, or None if it cannot be used."""
        if not pending or not getattr(self.backend, 'can_fuse_exchange', False) or victim < 1:
            return None
        m, bits = pending[-1]
        bits = list(bits)
        k = len(bits)
        if k > 5 or np.ndim(m) == 1:
            return None
        if k < 4:
            # identity on spare wires (the highest free local bits)
            pad = [p for p in range(self.n_local - 1, -1, -1) if p not in bits][: 4 - k]
            m = np.kron(np.eye(1 << len(pad)), np.asarray(m).reshape(1 << k, 1 << k))
            bits = pad + bits
        return m, bits

    def _local_form(self, m: np.ndarray, ws: tuple[int, ...]):
        """(matrix, local physical bits) if the block can run without
        communication on this rank, else None."""
        pw = [self.phys[w] for w in ws]
        glob = [p for p in pw if p >= self.n_local]
        if not glob:
            self.local_only_blocks += 1
            return m, pw
        if np.ndim(m) == 1:
            # a diagonal block never needs communication: this rank applies the
            # entries selected by its rank bits
            k = len(pw)
            t = np.asarray(m).reshape((2,) * k)
            sel = tuple(self._rank_bit(p) if p >= self.n_local else slice(None) for p in pw)
            rest = [p for p in pw if p < self.n_local]
            self.diag_global_blocks += 1
            sub = np.ascontiguousarray(t[sel]).reshape(-1)
            if not rest:
                self.local.scale(complex(sub[0]))
                return sub, []
            return sub, rest
        key = (id(m), tuple(pw))
        if key not in self._diag_cache:
            self._diag_cache[key] = block_diagonal_in(m, pw, glob, atol=1e-24)
        subs = self._diag_cache[key]
        if subs is None:
            return None
        self.diag_global_blocks += 1
        vals = tuple(self._rank_bit(p) for p in glob)
        rest = [p for p in pw if p < self.n_local]
        sub = subs[vals]
        if not rest:
            # pure phase on this rank
            self.local.scale(complex(sub.reshape(-1)[0]))
            return sub, []
        return sub, rest

    def _choose_victim(self, remaining, protected: set[int], with_pos: bool = False):
        """Local physical bit (not protected, >= 1 for 16-byte vectors) whose
        next use is furthest in the future (`with_pos`: also that position)."""
        next_use = {}
        for pos, (_, ws) in enumerate(remaining):
            for w in ws:
                p = self.phys[w]
                if p < self.n_local and p not in next_use:
                    next_use[p] = pos
        best, best_pos = None, -1
        lowest = 1 if self.dtype == np.dtype(np.complex64) else 0
        for p in range(self.n_local - 1, lowest - 1, -1):
            if p in protected:
                continue
            pos = next_use.get(p, 1 << 60)
            if pos > best_pos:
                best, best_pos = p, pos
        if best is None:
            raise RuntimeError('no local bit available to swap with')
        return (best, best_pos) if with_pos else best

    # ------------------------------------------------------------------ read-out

    def norm2(self) -> float:
        return self.backend.all_reduce_sum(self.local.norm2())

    def sample(self, repetitions: int, seed=None, axes: Sequence[int] | None = None) -> np.ndarray:
        """uint8[reps, len(axes)] bitstrings (column i = logical qubit axis axes[i],
        i.e. logical bit n-1-axes[i]; all n axes in order by default) drawn from
        |psi|^2; identical on every rank.  Every rank draws
        the same uniforms; a sample belongs to the rank whose cumulative
        probability interval contains it and is resolved there by the 1-GPU
        sampler."""
        rng = np.random.RandomState(seed)
        u = rng.random_sample(repetitions)
        totals = np.array(self.backend.all_gather_floats(self.local.norm2()))
        cum = np.cumsum(totals)
        target = u * cum[-1]
        owner = np.minimum(np.searchsorted(cum, target, side='right'), self.world - 1)
        mine = np.nonzero(owner == self.rank)[0]
        before = cum[self.rank] - totals[self.rank]
        local_u = np.clip((target[mine] - before) / max(totals[self.rank], 1e-300), 0.0, 1.0 - 2**-53)
        # column `axis` of the result is logical bit n-1-axis = physical bit phys[...]
        bits = [self.phys[self.n - 1 - axis] for axis in (range(self.n) if axes is None else axes)]
        return self.backend.merge_samples(repetitions, mine, local_u, bits)

    def gather_state(self) -> np.ndarray:
        """Full state in LOGICAL order on every rank (tests / small n only)."""
        parts = self.backend.gather_objects(self.local.to_numpy())
        phys_state = np.concatenate(parts)
        # un-permute: logical index bit l sits at physical bit phys[l]
        idx = np.arange(1 << self.n, dtype=np.int64)
        src = np.zeros_like(idx)
        for l in range(self.n):
            src |= ((idx >> l) & 1) << self.phys[l]
        return phys_state[src]

    def close(self):
        """Releases the shard memory — unless it belongs to the caller (a
        B200ShardedSimulator keeps one set of shards and IPC mappings across calls)."""
        if self._owns_backend and hasattr(self.backend, 'close'):
            self.backend.close()


# Lazy state growth: sub-states stay replicated and small until a
# gate needs one wider than this many bits; 2^31 amplitudes = 17 GB (complex64)
# bounds the temporaries of the join next to the shard itself.
LAZY_MAX_SHARD_BITS = 31


def plan_sharded(n_qubits: int, gates, dtype, max_fused_qubits, n_local: int, live_cls=None):
    """Host-side schedule of a gate list for the sharded path: the prefix that
    runs on small replicated sub-states (a `cirq_b200.plan` op list), the
    sub-states to join, and the fused blocks left for the sharded state.

    `live_cls` (a DeviceState class): the prefix is EXECUTED on device states of
    that class while it is being scheduled instead of recorded for a later replay,
    so the GPU works on the sub-states while the host fuses the rest of the
    circuit (the end-to-end path; the recorded form is for plan-once / replay)."""
    from cirq_b200.fusion import fuse_gates
    from cirq_b200.plan import SplitExecutor, _RecordingState

    class Rec(_RecordingState):
        ops = []
        counter = 0

    gates = list(gates)
    perm: dict = {}  # SWAP gates of the sharded part: wires renamed, no data moved
    if n_local > LAZY_MAX_SHARD_BITS:
        return {'n': n_qubits, 'dtype': np.dtype(dtype), 'ops': [], 'components': None,
                'blocks': fuse_gates(gates, max_fused_qubits, dtype, n_local, diagonal_blocks=True,
                                     permutation=perm),
                'permutation': perm}
    ex = SplitExecutor(n_qubits, dtype, max_fused_qubits, live_cls or Rec,
                       max_component_bits=min(n_local, 30))
    done = len(gates)
    for i, (m, b) in enumerate(gates):
        if not ex.apply(m, b):
            done = i
            break
    if live_cls is not None:
        comps = ex.components()  # live (device state, bits): launched already
        return {'n': n_qubits, 'dtype': np.dtype(dtype), 'ops': None, 'components': comps,
                'blocks': fuse_gates(gates[done:], max_fused_qubits, dtype, n_local, diagonal_blocks=True,
                                     permutation=perm),
                'permutation': perm, 'prefix_gates': done}
    comps = [(dev.ident, bits) for dev, bits in ex.components()]
    return {'n': n_qubits, 'dtype': np.dtype(dtype), 'ops': Rec.ops, 'components': comps,
            'blocks': fuse_gates(gates[done:], max_fused_qubits, dtype, n_local, diagonal_blocks=True,
                                 permutation=perm),
            'permutation': perm, 'prefix_gates': done}


def execute_sharded_plan(plan, sv: 'ShardedStateVector', device_state_cls=None) -> None:
    """Runs a `plan_sharded` schedule on `sv` (collective)."""
    if plan['components'] is None:
        sv.phys = list(range(sv.n))
        sv._init_basis(0)
    elif plan['ops'] is None:  # scheduled live: the sub-states exist already
        sv.load_product(plan['components'])
        plan['components'] = []  # (release the sub-states)
    else:
        from cirq_b200.plan import replay_ops

        live = replay_ops(plan['ops'], plan['dtype'], device_state_cls or type(sv.local))
        sv.load_product([(live[ident], bits) for ident, bits in plan['components']])
        del live
    sv.apply_blocks(plan['blocks'])
    sv.rename_bits(plan.get('permutation') or {})


class B200ShardedSimulator:
    """Cirq-facing entry point of the sharded path (SPMD: every rank of the
    process group calls it with the same circuit).

    Supports what config 4 needs: unitary circuits, optionally followed by
    terminal measurements (``run``).  Mid-circuit measurement, noise and
    classical control stay on the single-GPU simulators.
    """

    def __init__(self, *, dtype=np.complex64, seed=None, max_fused_qubits: int | None = None, group=None):
        self.dtype = np.dtype(dtype)
        self.seed = seed
        self.max_fused = max_fused_qubits
        self.group = group
        self._backend = None  # shard memory + IPC mappings, reused by successive run() calls

    def _backend_for(self, n_qubits: int):
        import torch.distributed as dist

        world = dist.get_world_size(self.group)
        n_local = n_qubits - (world.bit_length() - 1)
        if self._backend is not None and self._backend.n_local != n_local:
            self.close()
        if self._backend is None:
            self._backend = ShardBackend(n_local, self.dtype, self.group)
        return self._backend

    def close(self) -> None:
        """Releases the cached shard (collective: call on every rank)."""
        if self._backend is not None:
            self._backend.close()
            self._backend = None

    def _gates(self, circuit, qubits):
        from cirq_b200._cirq_compat import import_cirq
        from cirq_b200.sv_simulator import cached_unitary

        cirq = import_cirq()
        n = len(qubits)
        axis = {q: i for i, q in enumerate(qubits)}
        gates, measured = [], []
        for moment in circuit:
            for op in moment:
                if cirq.is_measurement(op):
                    measured.append(op)
                    continue
                if measured and any(q in {x for m in measured for x in m.qubits} for q in op.qubits):
                    raise ValueError('B200ShardedSimulator only supports terminal measurements')
                u = cached_unitary(op)
                if u is None:
                    raise TypeError(f"B200ShardedSimulator doesn't support {op!r}")
                gates.append((u, [n - 1 - axis[q] for q in op.qubits]))
        return gates, measured

    def simulate_sharded(self, circuit, qubit_order=None, initial_state: int = 0) -> ShardedStateVector:
        """Evolves |initial_state> and returns the sharded final state.  The state lives in
        the simulator's cached shards: it is valid until the next call on this simulator
        (or `close()`)."""
        from cirq_b200._cirq_compat import import_cirq
        from cirq_b200.fusion import fuse_gates

        cirq = import_cirq()
        qubits = cirq.QubitOrder.as_qubit_order(
            qubit_order if qubit_order is not None else cirq.QubitOrder.DEFAULT
        ).order_for(circuit.all_qubits())
        gates, _ = self._gates(circuit, qubits)
        if initial_state != 0:
            sv = ShardedStateVector(len(qubits), self.dtype, group=self.group,
                                    backend=self._backend_for(len(qubits)), initial_index=initial_state)
            perm: dict = {}
            sv.apply_blocks(fuse_gates(gates, self.max_fused, self.dtype, sv.n_local, diagonal_blocks=True,
                                       permutation=perm))
            sv.rename_bits(perm)
            return sv
        sv = ShardedStateVector(len(qubits), self.dtype, group=self.group,
                                backend=self._backend_for(len(qubits)), initial_index=None)
        execute_sharded_plan(plan_sharded(sv.n, gates, self.dtype, self.max_fused, sv.n_local,
                                          live_cls=type(sv.local)), sv)
        return sv

    def run(self, circuit, repetitions: int = 1) -> dict:
        """{key: int8[reps, n_measured]} like ``cirq.Result.measurements``,
        identical on every rank."""
        from cirq_b200._cirq_compat import import_cirq
        from cirq_b200.fusion import fuse_gates

        cirq = import_cirq()
        qubits = cirq.QubitOrder.DEFAULT.order_for(circuit.all_qubits())
        gates, measured = self._gates(circuit, qubits)
        if not measured:
            raise ValueError('Circuit has no measurements to sample.')
        sv = ShardedStateVector(len(qubits), self.dtype, group=self.group,
                                backend=self._backend_for(len(qubits)), initial_index=None)
        execute_sharded_plan(plan_sharded(sv.n, gates, self.dtype, self.max_fused, sv.n_local,
                                          live_cls=type(sv.local)), sv)
        axis = {q: i for i, q in enumerate(qubits)}
        cols = [axis[q] for op in measured for q in op.qubits]
        # only the measured columns leave the device, already in result order
        bits = sv.sample(repetitions, seed=self.seed, axes=cols).view(np.int8)
        out = {}
        start = 0
        for op in measured:
            arr = bits[:, start:start + len(op.qubits)]
            start += len(op.qubits)
            inv = [i for i, f in enumerate(op.gate.full_invert_mask()) if f]
            if inv:
                arr = arr.copy()
                arr[:, inv] ^= 1
            out[op.gate.key] = arr
        return out


# ---- sweeps over independent resolvers: replicas, no data-path collective ----------------


def run_sweep_sharded(make_simulator, program, params, repetitions: int = 1, seed: int | None = None,
                      group=None):
    """``run_sweep`` with the resolvers dealt out over the ranks of a process group
    (SPMD: every rank calls it with the same arguments and gets the full result
    list back, in resolver order).

    The reference's ``run_sweep_iter`` simulates resolver after resolver
    (cirq-core/cirq/sim/simulator.py:85-94); the resolvers are independent, so
    rank r takes resolvers r, r + world, r + 2 world, ... on its own GPU — whole
    replicas, nothing exchanged while simulating (config 5: a 16-qubit density
    matrix is 34 GB, one per GPU) — and only the sampled records are gathered at
    the end.  `make_simulator(seed)` builds the rank's simulator (for instance
    ``lambda s: B200DensityMatrixSimulator(noise=..., seed=s)``); rank r is seeded
    with ``seed + r``, so results are reproducible for a given world size but are
    not the single-process stream's.
    """
    import torch.distributed as dist

    from cirq_b200._cirq_compat import import_cirq

    cirq = import_cirq()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    resolvers = list(cirq.to_resolvers(params))
    mine = list(range(rank, len(resolvers), world))
    sim = make_simulator(None if seed is None else int(seed) + rank)
    local = sim.run_sweep(program, [resolvers[i] for i in mine], repetitions) if mine else []
    payload = [(i, {k: np.asarray(v) for k, v in r.records.items()}) for i, r in zip(mine, local)]
    gathered = [None] * world
    dist.all_gather_object(gathered, payload, group=group)
    records = {i: rec for part in gathered for i, rec in part}
    return [cirq.ResultDict(params=resolvers[i], records=records[i]) for i in range(len(resolvers))]
