"""BASELINE config 5 at full size across the GPUs of one node: 16-qubit noisy
QAOA, DensityMatrixSimulator with depolarizing noise, run_sweep over 256
resolvers — the resolvers dealt out as replicas (cirq_b200.dist.run_sweep_sharded).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29590 tools/qaoa_sweep_multi.py [--qubits 16] [--resolvers 256] [--reps 1000]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--qubits', type=int, default=16)
    ap.add_argument('--resolvers', type=int, default=256)
    ap.add_argument('--reps', type=int, default=1000)
    ap.add_argument('--out', default='')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local_rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    import cirq_b200
    from cirq_b200 import workloads as W
    from cirq_b200._cirq_compat import import_cirq
    from cirq_b200.dist import run_sweep_sharded

    cirq = import_cirq()
    circuit, qubits, names = W.qaoa_circuit(args.qubits)
    sweep = W.qaoa_sweep(names, args.resolvers)
    make = lambda s: cirq_b200.B200DensityMatrixSimulator(noise=cirq.depolarize(0.01), seed=s)
    # warm-up: one resolver per rank (allocator, kernels, unitary caches)
    run_sweep_sharded(make, circuit, list(cirq.to_resolvers(sweep))[:world], repetitions=10, seed=0)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    results = run_sweep_sharded(make, circuit, sweep, repetitions=args.reps, seed=0)
    torch.cuda.synchronize()
    dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        shapes = {tuple(r.measurements['m'].shape) for r in results}
        mean_bit = float(sum(r.measurements['m'].mean() for r in results) / len(results))
        line = dict(workload='config 5: noisy QAOA density matrix run_sweep', n_qubits=args.qubits,
                    rho_gb=(8 << (2 * args.qubits)) / 1e9, resolvers=len(results), repetitions=args.reps,
                    n_gpus=world, seconds=dt, resolvers_per_s=len(results) / dt,
                    seconds_per_resolver_per_gpu=dt * world / len(results),
                    result_shapes=[list(s) for s in shapes], mean_measured_bit=mean_bit,
                    how='resolvers dealt out over the ranks as replicas, records gathered at the end')
        print(json.dumps(line), flush=True)
        if args.out:
            with open(args.out, 'w') as f:
                f.write(json.dumps(line) + '\n')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
