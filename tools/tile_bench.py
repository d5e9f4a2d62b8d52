"""Two blocks as two passes vs one tile pass (b2q_sv_apply_tile_blocks), by
target-position class, on a dense state (CUDA events, sustained).

    python tools/tile_bench.py [--n 30] [--reps 20] [--out gpurun_out/tile_bench.json]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from cirq_b200.device_state import DeviceState  # noqa: E402


def rand_unitary(rng, k):
    d = 1 << k
    q, r = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
    return q * (np.diag(r) / np.abs(np.diag(r)))


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=30)
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--out', default='')
    args = ap.parse_args()
    n = args.n
    rng = np.random.RandomState(3)
    dev = DeviceState(n, np.complex64)
    dev.tensor.normal_()
    dev.tensor.mul_(2.0 ** (-(n + 1) / 2))
    hi = list(range(n - 1, n - 11, -1))
    cases = {
        'high bits, disjoint': (hi[:5], hi[5:10]),
        'high bits, 3 shared': ([n - 1, n - 2, n - 3, n - 4, n - 5], [n - 3, n - 4, n - 5, n - 6, n - 7]),
        'mixed positions': ([n - 1, 17, 9, 5, 12], [22, 9, n - 4, 14, 3]),
        'low bits 0-4 + 5-9': ([0, 1, 2, 3, 4], [5, 6, 7, 8, 9]),
        'bit 0 in one block': ([0, 13, 21, 25, 7], [n - 1, 11, 16, 19, 23]),
        'k=4 + k=3': ([n - 1, 9, 14, 20], [6, 17, 22]),
        'random a': (rng.permutation(n)[:5].tolist(), rng.permutation(n)[:5].tolist()),
        'random b': (rng.permutation(n)[:5].tolist(), rng.permutation(n)[:5].tolist()),
        'random c': (rng.permutation(n)[:5].tolist(), rng.permutation(n)[:5].tolist()),
    }
    state_gb = (8 << n) / 1e9
    rows = []
    for name, (ta, tb) in cases.items():
        ma, mb = rand_unitary(rng, len(ta)), rand_unitary(rng, len(tb))
        two = timed(lambda: (dev.apply_matrix(ma, ta), dev.apply_matrix(mb, tb)), args.reps)
        one = timed(lambda: dev.apply_tile_blocks([(ma, ta), (mb, tb)]), args.reps)
        single = timed(lambda: dev.apply_tile_blocks([(ma, ta)]), args.reps)
        rows.append({'case': name, 'targets': [ta, tb], 'two_passes_ms': two, 'tile_pass_ms': one,
                     'tile_pass_one_block_ms': single, 'speedup': two / one,
                     'tile_GBps_algorithmic': 2 * state_gb / (one * 1e-3)})
        print(f'{name}: two passes {two:.3f} ms, one tile pass {one:.3f} ms (x{two / one:.2f}), '
              f'tile pass with one block {single:.3f} ms, {2 * state_gb / (one * 1e-3):.0f} GB/s', flush=True)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
        with open(args.out, 'w') as f:
            json.dump({'n': n, 'reps': args.reps, 'rows': rows}, f, indent=1)


if __name__ == '__main__':
    main()
