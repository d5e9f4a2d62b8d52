"""Phase timeline of the tile kernel (clock64 stamps of CTA 0's first tiles).

    python tools/tile_trace.py [--n 30] [--blocks 2]

Needs an INSTRUMENTED build of the kernel (the stamps cost registers, so they are not
in the product): `git apply tools/tile_kernel_experiments.patch`, rebuild with
`-DB2Q_TILE_TRACE`.  That patch is the last experiment of round 2 — the stamps plus a
dedicated 17th warp that issues the MMAs — which measured SLOWER (7.4-7.9 ms per
two-block pass: the 17th warp caps registers at 96 per thread) and was not kept;
profiles/r2l_tile_trace_*.log are timelines of the committed kernel structure.
"""
import argparse
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from cirq_b200 import _lib  # noqa: E402
from cirq_b200.device_state import DeviceState  # noqa: E402

NAMES = ['loop top', 'cp.async landed', 'CTA barrier', 'A: gather+split+st issued', 'A: operand ready (group barrier)',
         'A: MMAs issued + HBM traffic issued', 'A: MMAs done', 'A: epilogue done', 'A: CTA barrier',
         'B: gather+split+st issued', 'B: operand ready', 'B: MMAs issued', 'B: MMAs done', 'B: epilogue done',
         'B: CTA barrier']


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=30)
    ap.add_argument('--blocks', type=int, default=2)
    args = ap.parse_args()
    lib = _lib.load()
    n = args.n
    rng = np.random.RandomState(3)
    dev = DeviceState(n, np.complex64)
    dev.tensor.normal_()
    dev.tensor.mul_(2.0 ** (-(n + 1) / 2))
    q, _ = np.linalg.qr(rng.standard_normal((32, 32)) + 1j * rng.standard_normal((32, 32)))
    group = [(q, [n - 1, 17, 9, 5, 12]), (q, [22, 9, n - 4, 14, 3])][: args.blocks]
    for _ in range(3):
        dev.apply_tile_blocks(group)
    buf = torch.zeros(2 * 24 * 16, dtype=torch.int64, device='cuda')
    lib.b2q_debug_tile_trace(ctypes.c_void_p(buf.data_ptr()))
    dev.apply_tile_blocks(group)
    torch.cuda.synchronize()
    lib.b2q_debug_tile_trace(ctypes.c_void_p(0))
    t = buf.cpu().numpy().reshape(2, 24, 16)
    last = 3 + 6 * args.blocks
    for g in range(2):
        print(f'--- M-tile group {g}: cycles per phase, median over tiles 4..23')
        d = np.diff(t[g, 4:, :last], axis=1)
        for i in range(last - 1):
            print(f'  {NAMES[i]:<40s} -> {NAMES[i + 1]:<40s} {int(np.median(d[:, i])):6d}')
        period = np.diff(t[g, 4:, 0])
        print(f'  tile period {int(np.median(period))} cycles; sum of phases {int(np.median(d.sum(axis=1)))}')


if __name__ == '__main__':
    main()
