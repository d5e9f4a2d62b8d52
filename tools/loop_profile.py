"""cProfile of the per-repetition noisy loop (B200Simulator without trajectory_batch)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    import torch

    import cirq_b200
    from cirq_b200._cirq_compat import import_cirq
    from traj_bench import noisy_brickwork

    cirq = import_cirq()
    c, q = noisy_brickwork(cirq, 16, 8, 0)
    sim = cirq_b200.B200Simulator(noise=cirq.depolarize(0.01), seed=1)
    sim.run(c, repetitions=4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sim.run(c, repetitions=16)
    torch.cuda.synchronize()
    print('per repetition %.1f ms' % ((time.perf_counter() - t0) / 16 * 1e3), flush=True)
    pr = cProfile.Profile()
    pr.enable()
    sim.run(c, repetitions=8)
    pr.disable()
    pstats.Stats(pr).sort_stats('tottime').print_stats(18)


if __name__ == '__main__':
    main()
