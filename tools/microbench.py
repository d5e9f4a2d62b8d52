"""Times single gate passes by target-position class (CUDA events, on a B200).

    python tools/microbench.py [--n 30] [--dtype c64] [--reps 10] [--out gpurun_out/microbench.json]
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


WARM_MS = 0.0
CLOCKS = {}


def _nvml_clock():
    """(SM MHz, power W) right now, via NVML (None when unavailable)."""
    try:
        import pynvml

        if not CLOCKS:
            pynvml.nvmlInit()
            CLOCKS['h'] = pynvml.nvmlDeviceGetHandleByIndex(0)
        h = CLOCKS['h']
        return (pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
    except Exception:
        return None


class _Applier:
    """dev.apply_matrix, or dev.apply_diagonal for a 1-D `m` (a diagonal block)."""

    def __init__(self, dev):
        self.dev = dev

    def apply_matrix(self, m, bits):
        if np.ndim(m) == 1:
            self.dev.apply_diagonal(m, bits)
        else:
            self.dev.apply_matrix(m, bits)


def time_pass(dev, m, bits, reps):
    dev = _Applier(dev)
    for _ in range(3):
        dev.apply_matrix(m, bits)
    if WARM_MS > 0:
        # sustained mode: keep the GPU under this kernel's load until clocks settle
        torch.cuda.synchronize()
        import time

        t0 = time.perf_counter()
        while (time.perf_counter() - t0) * 1e3 < WARM_MS:
            for _ in range(20):
                dev.apply_matrix(m, bits)
            torch.cuda.synchronize()
        start = torch.cuda.Event(enable_timing=True)
        end = torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(reps):
            dev.apply_matrix(m, bits)
        end.record()
        time.sleep(reps * 0.0027 * 0.6)
        time_pass.last_clock = _nvml_clock()
        torch.cuda.synchronize()
        return start.elapsed_time(end) / reps
    torch.cuda.synchronize()
    start = torch.cuda.Event(enable_timing=True)
    end = torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(reps):
        dev.apply_matrix(m, bits)
    end.record()
    torch.cuda.synchronize()
    return start.elapsed_time(end) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=30)
    ap.add_argument('--dtype', default='c64')
    ap.add_argument('--reps', type=int, default=10)
    ap.add_argument('--out', default='')
    ap.add_argument('--lane-mode', type=int, default=None)
    ap.add_argument('--only', default='')
    ap.add_argument('--vec-mode', type=int, default=None)
    ap.add_argument('--tc-mode', type=int, default=None)
    ap.add_argument('--tc-stage-mode', type=int, default=None)
    ap.add_argument('--dense', action='store_true', help='start from a dense random state')
    ap.add_argument('--stage-opts', default=None, help='early,l2_ahead for the staged TC kernel')
    ap.add_argument('--warm-ms', type=float, default=0.0,
                    help='sustained mode: run each class this long before timing it')
    args = ap.parse_args()
    global WARM_MS
    WARM_MS = args.warm_ms
    dtype = np.complex64 if args.dtype == 'c64' else np.complex128
    if args.tc_mode is not None:
        from cirq_b200 import _lib
        _lib.load().b2q_set_tc_mode(args.tc_mode)
    if args.stage_opts is not None:
        from cirq_b200 import _lib
        e, a = (int(v) for v in args.stage_opts.split(','))
        _lib.check(_lib.load().b2q_set_tc_stage_opts(e, a))
    if args.tc_stage_mode is not None:
        from cirq_b200 import _lib
        _lib.load().b2q_set_tc_stage_mode(args.tc_stage_mode)
    if args.vec_mode is not None:
        from cirq_b200 import _lib
        _lib.load().b2q_set_vec_mode(args.vec_mode)
    if args.lane_mode is not None:
        from cirq_b200 import _lib
        _lib.load().b2q_set_lane_mode(args.lane_mode)
    n = args.n
    rng = np.random.RandomState(0)
    dev = DeviceState.basis(n, dtype, 0)
    if args.dense:
        # a dense random state: data-dependent power draw like a real circuit's
        g = torch.Generator(device='cuda').manual_seed(1)
        dev.tensor.normal_(generator=g)
        dev.tensor.mul_(1.0 / float(np.sqrt(2.0 * (1 << n))))
    bytes_per_pass = 2 * dev.nbytes
    zb = 6 if dtype == np.complex64 else 5
    hi = n - 1
    classes = {
        'k1_high': [hi], 'k1_mid': [zb + 2], 'k1_bit0': [0], 'k1_lane': [3],
        'k2_high': [hi, hi - 1], 'k2_mid': [zb, zb + 1], 'k2_spread': [hi, zb + 3],
        'k2_bit0_high': [0, hi], 'k2_lane_high': [2, hi], 'k2_lane_lane': [1, 4],
        'k2_bit0_lane': [0, 3], 'k2_low': [1, 2], 'k2_bit01': [0, 1],
        'k3_high': [hi, hi - 1, hi - 2], 'k3_mid': [zb, zb + 1, zb + 2],
        'k3_lane_high': [2, hi, hi - 5], 'k3_lanes': [1, 2, 3], 'k3_bit012': [0, 1, 2], 'k3_lanes345': [3, 4, 5],
        'k4_high': [hi, hi - 1, hi - 2, hi - 3], 'k4_mid': [zb, zb + 2, zb + 4, zb + 6],
        'k4_lane_high': [1, 3, hi, hi - 1], 'k4_lanes': [1, 2, 3, 4], 'k4_lane5_high': [5, hi, hi - 1, hi - 2], 'k4_bit0_high': [0, hi, hi - 1, hi - 2], 'k4_lane1_high': [1, hi, hi - 1, hi - 2],
    }
    if dtype == np.complex64:
        classes.update({
            'k5_high': [hi, hi - 1, hi - 2, hi - 3, hi - 4],
            'k5_lane_high': [2, 4, hi, hi - 1, hi - 2],
            'k5_lanes': [1, 2, 3, 4, 5],
            'k5_bit0_high': [0, hi, hi - 1, hi - 2, hi - 3],
            'k5_bit1_high': [1, hi, hi - 1, hi - 2, hi - 3],
            'k5_bit01_high': [0, 1, hi, hi - 1, hi - 2],
            'k5_low': [0, 1, 2, 3, 4],
            'k5_bit0_mid': [0, 3, 7, 12, 20],
            'k5_bit1_mid': [1, 3, 7, 12, 20],
        })
    # diagonal blocks (table in shared memory up to 13 / 12 wires)
    kd = 13 if dtype == np.complex64 else 12
    classes.update({
        'diag2_high': [hi, hi - 1],
        f'diag{kd}_low': list(range(kd)),
        f'diag{kd}_high': list(range(hi, hi - kd, -1)),
        f'diag{kd}_mixed': [0, 2, 5, 9, 11, 12, 14, 17, 20, 23, 25, hi - 1, hi][:kd],
        'diag16_mixed': [0, 2, 5, 9, 11, 12, 13, 14, 17, 19, 20, 23, 25, 27, hi - 1, hi],
    })
    results = {}
    if args.only:
        classes = {k: v for k, v in classes.items() if any(k.startswith(p) for p in args.only.split(','))}
    for name, bits in classes.items():
        if name.startswith('diag'):
            m = np.exp(1j * rng.standard_normal(1 << len(bits)))
        else:
            m = rand_unitary(rng, len(bits))
        ms = time_pass(dev, m, bits, args.reps)
        gbs = bytes_per_pass / ms / 1e6
        results[name] = {'bits': bits, 'ms': ms, 'GBps': gbs}
        clock = getattr(time_pass, 'last_clock', None)
        if clock:
            results[name]['sm_mhz'], results[name]['power_w'] = clock
        print(f'{name:16s} bits={bits!s:28s} {ms:8.3f} ms  {gbs:8.1f} GB/s  {clock or ""}', flush=True)
    # reference points: torch copy (read+write) and the diagonal/scale kernels
    other = torch.empty_like(dev.tensor)
    for _ in range(3):
        other.copy_(dev.tensor)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.reps):
        other.copy_(dev.tensor)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / args.reps
    results['torch_copy'] = {'ms': ms, 'GBps': bytes_per_pass / ms / 1e6}
    print(f'torch_copy {ms:8.3f} ms {bytes_per_pass / ms / 1e6:8.1f} GB/s')
    del other
    s.record()
    for _ in range(args.reps):
        dev.scale(1.0)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / args.reps
    results['scale_inplace'] = {'ms': ms, 'GBps': bytes_per_pass / ms / 1e6}
    print(f'scale_inplace {ms:8.3f} ms {bytes_per_pass / ms / 1e6:8.1f} GB/s')
    dev.norm2()  # first call loads the kernel (CUDA lazy module loading)
    s.record()
    nrm = dev.norm2()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    results['norm2'] = {'ms': ms, 'GBps': dev.nbytes / ms / 1e6, 'value': nrm}
    print(f'norm2 {ms:8.3f} ms {dev.nbytes / ms / 1e6:8.1f} GB/s (read only) value={nrm}')
    u = rng.random_sample(1_000_000)
    s.record()
    idx = dev.sample_indices_device(u)
    e.record()
    torch.cuda.synchronize()
    results['sample_1M'] = {'ms': s.elapsed_time(e)}
    print(f'sample 1M: {s.elapsed_time(e):8.3f} ms')
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, 'w') as f:
            json.dump({'n': n, 'dtype': args.dtype, 'results': results}, f, indent=1)


if __name__ == '__main__':
    main()
