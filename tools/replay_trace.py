"""Per-operation device time of a cached schedule (cirq_b200/plan_cache.py): replays the
recorded operations of one workload with a synchronize after each and prints where the
time of an end-to-end call goes (joins, gate passes by state size, the final flush).

    python tools/replay_trace.py --workload qft34
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='qft34')
    ap.add_argument('--min-ms', type=float, default=1.0)
    ap.add_argument('--repeat-alone', type=int, default=1)
    args = ap.parse_args()
    import torch

    import bench as B
    import cirq_b200
    from cirq_b200.device_state import DeviceState

    wl = B.build_workload(args.workload)
    sim = cirq_b200.B200Simulator(dtype=B.WORKLOAD_DTYPE.get(args.workload, np.complex64))
    plan = sim._record_prefix(wl['circuit'], wl['qubits'])
    assert plan is not None

    def describe(blocks):
        return '[' + ','.join(('d%d' % len(w)) if np.ndim(m) == 1 else str(len(w)) for m, w in blocks) + ']'

    for rep in range(2):
        live, sizes, total = {}, {}, {}
        torch.cuda.synchronize()
        t_all = time.perf_counter()
        for op in plan.ops:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            kind = op[0]
            if kind == 'basis':
                live[op[1]] = DeviceState.basis(op[2], sim._dtype, op[3])
                sizes[op[1]] = op[2]
                what = 'basis %d' % op[2]
            elif kind == 'kron':
                a, b = live.pop(op[2]), live.pop(op[3])
                live[op[1]] = a.kron(b)
                del a, b
                sizes[op[1]] = sizes[op[2]] + sizes[op[3]]
                what = 'kron %d x %d' % (sizes[op[2]], sizes[op[3]])
            elif kind == 'apply':
                live[op[1]].apply_batch(op[2])
                what = 'apply on %d bits %s (%d passes)' % (sizes[op[1]], describe(op[2]), op[3])
            elif kind == 'permute':
                live[op[1]].permute_bits_inplace(op[2])
                what = 'permute %d bits' % sizes[op[1]]
            else:
                live[op[1]].scale(op[2])
                what = 'scale'
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3
            key = what.split(' [')[0] if kind == 'apply' else kind
            total[key] = total.get(key, 0.0) + ms
            if rep == 1 and ms >= args.min_ms:
                free, _ = torch.cuda.mem_get_info()
                print('%9.2f ms  %s   (free %.1f GiB, torch reserved %.1f GiB)' % (
                    ms, what, free / 2**30, torch.cuda.memory_reserved() / 2**30))
        for ident, qs, where, blocks in plan.components:
            if blocks:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                live[ident].apply_batch(blocks)
                torch.cuda.synchronize()
                ms = (time.perf_counter() - t0) * 1e3
                total['held blocks'] = total.get('held blocks', 0.0) + ms
                if rep == 1:
                    print('%9.2f ms  held blocks on %d bits %s = %d passes' % (
                        ms, len(qs), describe(blocks), len(live[ident].plan_passes(blocks))))
                    for again in range(args.repeat_alone):
                        for blk in blocks:
                            t0 = time.perf_counter()
                            live[ident].apply_batch([blk])
                            torch.cuda.synchronize()
                            ms = (time.perf_counter() - t0) * 1e3
                            if again == 0 or ms > 70:
                                print('            %8.2f ms  alone (round %d): %s on %s' % (
                                    ms, again, describe([blk]), list(blk[1])))
        torch.cuda.synchronize()
        if rep == 1:
            print('total (serialised) %.1f ms' % ((time.perf_counter() - t_all) * 1e3))
            for k, v in sorted(total.items(), key=lambda kv: -kv[1])[:12]:
                print('   %9.2f ms  %s' % (v, k))
        del live
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
