"""Single-process probe of the peer-memory exchange kernels (for ncu: a profiler
cannot attach to a multi-rank job).  Two shards live on cuda:0 and cuda:1 of ONE
process; cuda:0 runs its half of the exchange against cuda:1's memory over NVLink.

    ncu --clock-control none --metrics gpu__time_duration.sum,nvltx__bytes.sum,nvlrx__bytes.sum,\
dram__bytes_read.sum,dram__bytes_write.sum -k regex:dist_swap -o gpurun_out/nvlink \
        python tools/nvlink_probe.py --n-local 28

Only rank 0's half runs (the partner's kernel would run on the other GPU in the
real job), so the link carries HALF of a swap's traffic: 1/4 shard written to the
peer and 1/4 shard read from it.
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from cirq_b200 import _lib  # noqa: E402
from cirq_b200._lib import check  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n-local', type=int, default=28)
    ap.add_argument('--out', default='')
    args = ap.parse_args()
    assert torch.cuda.device_count() >= 2, 'needs 2 GPUs'
    lib = _lib.load()
    n = args.n_local
    a = torch.empty((1 << n, 2), dtype=torch.float32, device='cuda:0').normal_()
    b = torch.empty((1 << n, 2), dtype=torch.float32, device='cuda:1').normal_()
    # one peer copy makes torch enable peer access in both directions
    tmp = b[:16].to('cuda:0')
    tmp2 = a[:16].to('cuda:1')
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    del tmp, tmp2
    a0, b0 = a[:4096].cpu().clone(), b[:4096].cpu().clone()
    torch.cuda.set_device(0)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rows = []

    def timed(label, fn, out_bytes):
        fn()
        torch.cuda.synchronize(0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.nvtx.range_push('b2q')
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize(0)
        torch.cuda.nvtx.range_pop()
        ms = s.elapsed_time(e)
        rows.append({'kernel': label, 'ms': ms, 'bytes_written_to_peer': out_bytes, 'bytes_read_from_peer': out_bytes,
                     'GBps_per_direction': out_bytes / ms / 1e6})
        print(f'{label}: {ms:.3f} ms, {out_bytes / 1e9:.3f} GB each way, {out_bytes / ms / 1e6:.0f} GB/s per direction',
              flush=True)

    shard = 8 << n
    for lbit in (n - 1, 5):
        timed(f'dist_swap_bit_kernel (rank 0 half, local bit {lbit})',
              lambda lbit=lbit: check(lib.b2q_dist_swap_bit(ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()),
                                                           _lib.C64, n, lbit, 0, stream)), shard // 4)
    peers = (ctypes.c_void_p * 2)(None, b.data_ptr())
    timed('dist_swap_multi_kernel (m = 1, rank 0 half)',
          lambda: check(lib.b2q_dist_swap_bits(ctypes.c_void_p(a.data_ptr()), peers, _lib.C64, n,
                                               _lib.int_array([n - 2]), 1, 0, stream)), shard // 4)
    # fused gate + exchange: out of place into spare buffers (local one on cuda:0, peer one on cuda:1)
    if n <= 30:
        out_local = torch.empty_like(a)
        out_peer = torch.empty_like(b)
        rs = np.random.RandomState(5)
        q5, _ = np.linalg.qr(rs.standard_normal((32, 32)) + 1j * rs.standard_normal((32, 32)))
        m = np.ascontiguousarray(q5, dtype=np.complex128)
        bits5 = _lib.int_array([n - 3, 12, 9, 7, 3])
        timed('sv_apply_tc_staged_kernel + exchange (b2q_dist_apply_exchange, rank 0 side)',
              lambda: check(lib.b2q_dist_apply_exchange(ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(out_local.data_ptr()),
                                                        ctypes.c_void_p(out_peer.data_ptr()), _lib.C64, n,
                                                        m.ctypes.data, bits5, 5, n - 1, 0, stream)), shard // 2)
        rows[-1]['bytes_read_from_peer'] = 0
    # an even number of swaps of each bit: the shards are back where they started
    ok = bool(torch.equal(a[:4096].cpu(), a0) and torch.equal(b[:4096].cpu(), b0))
    print('shards restored after paired swaps:', ok, flush=True)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
        with open(args.out, 'w') as f:
            json.dump({'n_local': n, 'shard_bytes': shard, 'rows': rows, 'restored': ok}, f, indent=1)


if __name__ == '__main__':
    main()
