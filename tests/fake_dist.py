"""TEST-ONLY backend for cirq_b200.dist.ShardedStateVector: numpy shards (the
oracle-backed fake device) exchanged over torch.distributed/gloo.  Exercises
the host-side sharding logic (bit maps, scheduler, victim choice, sampling
routing) with world_size > 1 on a machine without GPUs."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from fake_device import OracleDeviceState


class GlooShardBackend:
    def __init__(self, n_local, dtype):
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.n_local = n_local
        self.dtype = np.dtype(dtype)
        self.local = OracleDeviceState(n_local, dtype)

    def barrier(self):
        dist.barrier()

    def swap_bit(self, partner, local_bit, my_gbit):
        # amplitudes with local bit != my global bit value go to the partner
        arr = self.local.array
        idx = np.arange(arr.size, dtype=np.int64)
        sel = ((idx >> local_bit) & 1) == (1 - my_gbit)
        send = np.ascontiguousarray(arr[sel]).view(np.float64 if self.dtype == np.complex128 else np.float32)
        send_t = torch.from_numpy(send.copy())
        recv_t = torch.empty_like(send_t)
        if self.rank < partner:
            dist.send(send_t, partner)
            dist.recv(recv_t, partner)
        else:
            dist.recv(recv_t, partner)
            dist.send(send_t, partner)
        arr[sel] = recv_t.numpy().view(self.dtype)
        dist.barrier()

    def swap_bits(self, pairs):
        """Multi-bit exchange (same result as the CUDA kernel): sub-block c of this
        rank trades places with sub-block rho of the rank whose sub-rank is c."""
        m = len(pairs)
        arr = self.local.array
        idx = np.arange(arr.size, dtype=np.int64)
        sub = np.zeros_like(idx)
        rho = 0
        base = self.rank
        for i, (gi, l) in enumerate(pairs):
            sub |= ((idx >> l) & 1) << i
            rho |= ((self.rank >> gi) & 1) << i
            base &= ~(1 << gi)
        real = np.float64 if self.dtype == np.complex128 else np.float32
        incoming = {}
        for c in range(1 << m):
            if c == rho:
                continue
            partner = base
            for i, (gi, _) in enumerate(pairs):
                partner |= ((c >> i) & 1) << gi
            send_t = torch.from_numpy(np.ascontiguousarray(arr[sub == c]).view(real).copy())
            recv_t = torch.empty_like(send_t)
            if self.rank < partner:
                dist.send(send_t, partner)
                dist.recv(recv_t, partner)
            else:
                dist.recv(recv_t, partner)
                dist.send(send_t, partner)
            incoming[c] = recv_t.numpy().view(self.dtype)
        for c, data in incoming.items():
            arr[sub == c] = data
        dist.barrier()
        self.multi_calls = getattr(self, 'multi_calls', 0) + 1

    can_fuse_exchange = True

    def apply_exchange(self, matrix, bits, partner, local_bit, my_gbit):
        """Same result as the fused CUDA kernel: the block, then the exchange."""
        assert len(bits) in (4, 5) and local_bit >= 1
        self.local.apply_matrix(np.asarray(matrix), list(bits))
        self.swap_bit(partner, local_bit, my_gbit)
        self.fused_calls = getattr(self, 'fused_calls', 0) + 1

    def all_reduce_sum(self, value):
        t = torch.tensor([value], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())

    def all_gather_floats(self, value):
        t = torch.tensor([value], dtype=torch.float64)
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t)
        return [float(x.item()) for x in out]

    def gather_objects(self, obj):
        out = [None] * self.world
        dist.all_gather_object(out, obj)
        return out

    def merge_samples(self, reps, positions, local_uniforms, bits):
        from oracle import sv_oracle as orc

        if positions.size:
            idx = np.asarray(self.local.sample_indices(local_uniforms), dtype=np.int64)
            idx = idx | (self.rank << self.n_local)
        else:
            idx = np.zeros(0, dtype=np.int64)
        full = torch.zeros(max(reps, 1), dtype=torch.int64)
        full[torch.from_numpy(positions)] = torch.from_numpy(idx)
        dist.all_reduce(full)
        return orc.unpack_bits(full[:reps].numpy().astype(np.uint64), list(bits))
