"""Benchmark of the hot path: fused gates/s on BASELINE.json's configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload rqc30]

One "step" = one pass of the hot path over one synthetic circuit: |0..0> on the
device, every fused gate block applied (one streaming HBM pass each), and (for
the Sycamore-style config) 1M bitstrings sampled.  Prints ONE JSON line.

value     fused gates/s with all inputs resident in HBM (plan pre-built, uniforms
          uploaded before the timed region).  The gate unit is the k<=2 fused
          block — what cirq.merge_k_qubit_unitaries(k=2) yields (245 for the
          30-qubit config) — so numbers are comparable with BASELINE.md even
          though this backend fuses wider and needs fewer passes.
e2e       the same metric through the public API a Cirq user calls:
          B200Simulator(seed=0).run(circuit, repetitions=...) on the cirq.Circuit,
          host scheduling, uploads and the device->host copy of the samples
          inside the timed region.
roofline  dominant kernel (sv_apply_fast_kernel): algorithmic bytes per launch
          (2 * sizeof(complex) * 2^n) / mean launch duration measured with CUDA
          events around the gate passes, vs the measured copy peak.
cpu_baseline  the reference cirq.Simulator on the box's host cores on a bounded
          sample of the same generator (fewer qubits), timed in the same run.
          Its gates are counted in the workload's own unit — gates on the
          workload's state size: a gate pass over 2^m amplitudes counts as
          2^(m-n) gates of the n-qubit workload (the same convention the N > 1
          arm uses for its larger states); the raw number measured on the sample
          is reported next to it.  The reference's time per gate grows at least
          linearly in 2^n (20 -> 24 qubits: x24 for x16 amplitudes), so this
          favours the reference.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, params, repetitions)
    'rqc30': ('rqc', dict(rows=5, cols=6, depth=20, seed=1), 1_000_000),
    'rqc24': ('rqc', dict(rows=4, cols=6, depth=20, seed=1), 100_000),
    'rqc20': ('rqc', dict(rows=4, cols=5, depth=20, seed=1), 100_000),
    'qft34': ('qft', dict(n=34), 0),
    'qft30': ('qft', dict(n=30), 0),
    'rc20': ('rc', dict(n=20, depth=20, seed=1234), 0),
}
# bounded CPU samples (same generator, fewer qubits): ~5-20 s of reference time per step
CPU_SAMPLE = {'rqc30': 'rqc20', 'rqc24': 'rqc20', 'rqc20': 'rqc20', 'qft34': 'qft22', 'qft30': 'qft22',
              'qft22': 'qft22', 'rc20': 'rc20'}
# the reference arm (`--impl reference`) has minutes, not seconds: a larger sample
REFERENCE_ARM_SAMPLE = dict(CPU_SAMPLE, rqc30='rqc24', qft34='qft24', qft30='qft24')
WORKLOADS['qft22'] = ('qft', dict(n=22), 0)
WORKLOADS['qft24'] = ('qft', dict(n=24), 0)


def quiet_stdout():
    """Sends everything libraries write to stdout (NCCL prints a version banner
    there) to stderr and keeps the real stdout for the ONE JSON line.  The saved
    descriptor travels in the environment: bench.py runs as __main__, while
    cirq_b200.dist_bench imports it a second time as `bench`."""
    if 'B2Q_BENCH_RESULT_FD' not in os.environ:
        sys.stdout.flush()
        os.environ['B2Q_BENCH_RESULT_FD'] = str(os.dup(1))
        os.dup2(2, 1)


def emit(line: dict) -> None:
    """The bench result: one JSON line on the process's original stdout."""
    text = (json.dumps(line) + '\n').encode()
    sys.stdout.flush()
    fd = os.environ.get('B2Q_BENCH_RESULT_FD')
    os.write(int(fd) if fd is not None else 1, text)


def workload_equivalent(value, n_sample, workload):
    """gates/s measured on an n_sample-qubit sample -> gates/s in units of the
    workload's state size (a pass over 2^m amplitudes = 2^(m-n) workload gates)."""
    kind, params, _ = WORKLOADS[workload]
    n = params['rows'] * params['cols'] if kind == 'rqc' else params['n']
    return value * 2.0 ** (n_sample - n), n


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def measured_traffic(n_qubits, kernel):
    """dram read+write bytes per launch of `kernel` from the committed ncu --set
    full captures (profiles/ncu_traffic.json), valid for 30-qubit states only."""
    path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if n_qubits != 30 or not os.path.exists(path):
        return None
    with open(path) as f:
        table = json.load(f)['bytes_per_launch_30q']
    return table.get(kernel)


def kernel_class(wires):
    """Name of the kernel a fused block on these index bits runs on (complex64,
    default knobs: b2q_apply_tc.cu launch_tc_k / b2q_apply.cu apply_matrix_t)."""
    k = len(wires)
    tc4 = os.environ.get('CIRQ_B200_TC_MODE') == '2'  # opt-in: 4-qubit blocks on tcgen05 too
    if k == 5 or (k == 4 and tc4):
        return f'sv_apply_tc_staged_kernel<{k}>' if min(wires) < 2 else f'sv_apply_tc_kernel<{k}>'
    if k == 6:
        return 'sv_apply_tc_kernel<6>'
    return f'sv_apply_fast_kernel<float,{k}>'


def build_workload(name):
    """Returns dict(circuit, qubits, gates, n, reps, generator)."""
    from cirq_b200 import workloads as W
    from cirq_b200._cirq_compat import cirq_available

    kind, params, reps = WORKLOADS[name]
    if not cirq_available():
        if kind != 'rqc':
            raise RuntimeError('cirq is not importable and only the rqc workload has a builtin generator')
        gates = W.builtin_rqc_gates(**params)
        return dict(circuit=None, qubits=None, gates=gates, n=params['rows'] * params['cols'],
                    reps=reps, generator='builtin (cirq not importable)')
    if kind == 'rqc':
        circuit, qubits = W.rqc_circuit(**params)
    elif kind == 'qft':
        circuit, qubits = W.qft_circuit(**params)
    else:
        circuit, qubits = W.random_circuit(**params)
    gates = W.circuit_to_gates(circuit, qubits)
    return dict(circuit=circuit, qubits=qubits, gates=gates, n=len(qubits), reps=reps,
                generator='cirq ' + kind)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._thread = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
                     '-i', str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(',')]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith('active') for s in self.samples)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.samples[0][1]),
                'power_w_max': max(float(s[2]) for s in self.samples), 'reasons': reasons,
                'samples': len(self.samples)}


def time_reference(name, steps, warmup):
    """Times the unmodified reference (cirq.Simulator, numpy) on the host."""
    from cirq_b200._cirq_compat import import_cirq
    from cirq_b200.fusion import fuse_gates

    cirq = import_cirq()
    wl = build_workload(name)
    circuit, qubits, reps = wl['circuit'], wl['qubits'], wl['reps']
    unit_gates = len(fuse_gates(wl['gates'], 2))
    run_circuit = circuit
    reps_ref = min(reps, 10_000)
    if reps:
        run_circuit = circuit + cirq.Circuit(cirq.measure(*qubits, key='m'))

    def step():
        sim = cirq.Simulator(dtype=np.complex64, seed=0)
        if reps:
            sim.run(run_circuit, repetitions=reps_ref)
        else:
            sim.simulate(run_circuit, qubit_order=qubits)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dict(value=unit_gates / dt, seconds_per_step=dt, unit_gates=unit_gates, n=wl['n'],
                raw_ops=len(wl['gates']), reps=reps_ref)


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sample = REFERENCE_ARM_SAMPLE[args.workload]
    big = sample != CPU_SAMPLE[args.workload]  # ~90 s per step: keep the run within minutes
    steps = max(1, min(args.steps, 2 if big else 3))
    warmup = 0 if big else min(args.warmup, 1)
    r = time_reference(sample, steps, warmup)
    raw = r['value']
    r['value'], n_workload = workload_equivalent(raw, r['n'], args.workload)
    line = {
        'impl': 'reference', 'metric': 'fused_gates_per_s', 'value': r['value'], 'unit': 'gates/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup,
        'ms_per_step': r['seconds_per_step'] * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'c64', 'data': 'synthetic',
        'config': {'workload': args.workload, 'gate_unit': 'k<=2 fused blocks',
                   'reference_sample': sample, 'n_qubits': r['n'], 'raw_ops': r['raw_ops'],
                   'unit_gates': r['unit_gates'], 'repetitions': r['reps'],
                   'gates_per_s_on_sample': raw,
                   'value_unit': f"{n_workload}-qubit-equivalent gates/s = gates/s on the "
                                 f"{r['n']}-qubit sample x 2^({r['n']}-{n_workload})"},
        'cpu_baseline': {'value': r['value'], 'unit': 'gates/s', 'cores': 1, 'kind': 'reference',
                         'raw_value_on_sample': raw,
                         'sample': f"cirq.Simulator(complex64) on {sample} ({r['n']} qubits, same generator, "
                                   f"{r['seconds_per_step']:.1f} s per step): {raw:.3g} gates/s there, counted as "
                                   f"{n_workload}-qubit-equivalent gates (x 2^({r['n']}-{n_workload})); "
                                   f"numpy path is single-threaded; host has {os.cpu_count()} cores"},
        'e2e': {'value': r['value'], 'unit': 'gates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    from cirq_b200 import _lib
    from cirq_b200.device_state import DeviceState
    from cirq_b200.fusion import fuse_gates

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    if world > 1:
        from cirq_b200 import dist_bench

        dist_bench.run(args, world, rank, local_rank)
        return

    lib = _lib.load()
    peak_gbs, peak_src = load_peaks()
    wl = build_workload(args.workload)
    n, reps, gates = wl['n'], wl['reps'], wl['gates']
    unit_gates = len(fuse_gates(gates, 2))
    dtype = np.complex64
    from cirq_b200.plan import build_plan, replay_plan

    # Host scheduling happens ONCE, outside the timed region: fusion + lazy state
    # growth (sub-states joined by the kron kernel, as with split_untangled_states).
    plan = build_plan(n, gates, dtype, args.max_fused)
    full_blocks = [blk for op in plan['ops'] if op[0] == 'apply' for blk in op[2]]
    state_bytes = (8 << n)
    rng = np.random.RandomState(0)
    uniforms = rng.random_sample(max(reps, 1))
    u_dev = torch.from_numpy(uniforms).to('cuda')

    import ctypes

    ws_bytes = int(lib.b2q_sv_sample_workspace_bytes(n, reps)) if reps else 16
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device='cuda')
    out_idx = torch.empty(max(reps, 1), dtype=torch.int64, device='cuda')
    out_bits = torch.empty((max(reps, 1), n if reps else 1), dtype=torch.uint8, device='cuda')
    # column a of the samples = logical qubit axis a = logical bit n-1-a
    bits_order = _lib.int_array([plan['bit_of'][n - 1 - a] for a in range(n)])
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    per_kernel = {}

    def timed_apply(state, blocks):
        for m, w in blocks:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if np.ndim(m) == 1:  # a diagonal block (its table in shared memory)
                state.apply_diagonal(m, w)
                name = 'sv_apply_diag_smem_kernel<float>'
            else:
                state.apply_matrix(m, w)
                name = kernel_class(w)
            b.record()
            timed_apply.pairs.append((state.n_bits, name, a, b))

    timed_apply.pairs = []

    def step(record=False):
        dev = replay_plan(plan, on_apply=timed_apply if record else None)
        if reps:
            _lib.check(lib.b2q_sv_sample(dev.ptr, dev.code, n, ctypes.c_void_p(u_dev.data_ptr()), reps,
                                         ctypes.c_void_p(out_idx.data_ptr()),
                                         ctypes.c_void_p(ws.data_ptr()), ws_bytes, stream))
            _lib.check(lib.b2q_unpack_bits(ctypes.c_void_p(out_idx.data_ptr()), reps, bits_order, n,
                                           ctypes.c_void_p(out_bits.data_ptr()), stream))
        if record:
            torch.cuda.synchronize()
            for nb, name, a, b in timed_apply.pairs:
                if nb == n:  # launches on the full-size state: the HBM-bound ones
                    per_kernel.setdefault(name, []).append(a.elapsed_time(b))
            timed_apply.pairs = []
        del dev

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launches0 = int(lib.b2q_launch_count())
    with ClockSampler(local_rank) as clocks:
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        start.record()
        for _ in range(args.steps):
            step()
        end.record()
        torch.cuda.synchronize()
        ms_per_step = start.elapsed_time(end) / args.steps
        launches = (int(lib.b2q_launch_count()) - launches0) // max(args.steps, 1)
        # per-kernel duration of the gate passes (separate, event-bracketed steps)
        record_steps = min(3, args.steps)
        for _ in range(record_steps):
            step(record=True)
    total_gate_ms = sum(np.sum(v) for v in per_kernel.values())
    full_passes = sum(len(v) for v in per_kernel.values()) // record_steps
    mean_pass_ms = float(total_gate_ms / max(1, sum(len(v) for v in per_kernel.values())))
    breakdown = {k: {'launches_per_step': len(v) // record_steps,
                     'ms_per_launch': float(np.mean(v)),
                     'share_of_gate_time': float(np.sum(v) / total_gate_ms)}
                 for k, v in per_kernel.items()}
    dominant = max(breakdown, key=lambda k: breakdown[k]['share_of_gate_time'])
    pass_ms = breakdown[dominant]['ms_per_launch']
    achieved = 2 * state_bytes / (pass_ms * 1e-3) / 1e9
    value = unit_gates / (ms_per_step * 1e-3)
    blocks = full_blocks

    # ---- e2e through the public Cirq-facing API -------------------------------------------
    e2e = None
    if wl['circuit'] is not None:
        from cirq_b200._cirq_compat import import_cirq

        cirq = import_cirq()
        import cirq_b200

        circuit = wl['circuit']
        if reps:
            circuit = circuit + cirq.Circuit(cirq.measure(*wl['qubits'], key='m'))
        e2e_steps = max(1, min(args.steps, 3))

        def e2e_step():
            sim = cirq_b200.B200Simulator(dtype=dtype, seed=0, max_fused_qubits=args.max_fused)
            if reps:
                res = sim.run(circuit, repetitions=reps)
                return res.measurements['m'].shape
            # the reference's own API for reading a few amplitudes of a state too
            # large to download (SimulatesAmplitudes, sim/simulator.py:120-182)
            return sim.compute_amplitudes(circuit, [0, 1], qubit_order=wl['qubits'])

        torch.cuda.empty_cache()
        e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e2e_steps
        mat_bytes = int(sum(16 * m.size for m, _ in blocks))
        e2e = {'value': unit_gates / dt, 'unit': 'gates/s', 'ms_per_step': dt * 1e3,
               'h2d_bytes_per_step': int(8 * reps + mat_bytes),
               'd2h_bytes_per_step': int(reps * n if reps else 32),
               'api': 'cirq_b200.B200Simulator(seed=0).run(circuit, repetitions)' if reps
                      else 'cirq_b200.B200Simulator().compute_amplitudes(circuit, [0, 1])'}

    # ---- CPU baseline (reference on host cores, bounded sample) ---------------------------
    cpu = None
    if not args.no_cpu_baseline:
        try:
            sample = CPU_SAMPLE[args.workload]
            r = time_reference(sample, 3, 0)
            eq, _ = workload_equivalent(r['value'], r['n'], args.workload)
            cpu = {'value': eq, 'unit': 'gates/s', 'cores': 1, 'kind': 'reference',
                   'raw_value_on_sample': r['value'],
                   'sample': f"cirq.Simulator(complex64) on {sample}: {r['n']} qubits, {r['raw_ops']} ops = "
                             f"{r['unit_gates']} k<=2 blocks, {r['reps']} repetitions, 3 x {r['seconds_per_step']:.2f} s: "
                             f"{r['value']:.3g} gates/s there, counted as {n}-qubit-equivalent gates "
                             f"(x 2^({r['n']}-{n})); single-threaded numpy, host has {os.cpu_count()} cores"}
        except Exception as exc:  # reference not importable on this box
            cpu = {'value': None, 'unit': 'gates/s', 'cores': 1, 'kind': 'reference',
                   'sample': f'unavailable: {exc}'}

    line = {
        'metric': 'fused_gates_per_s', 'value': value, 'unit': 'gates/s', 'n_gpus': 1,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'c64',
        'data': 'synthetic',
        'config': {'workload': args.workload, 'generator': wl['generator'], 'n_qubits': n,
                   'raw_ops': len(gates), 'gate_unit': 'k<=2 fused blocks (reference merge_k_qubit_unitaries(k=2) count)',
                   'unit_gates': unit_gates, 'max_fused_qubits': max(len(w) for m, w in blocks if np.ndim(m) == 2),
                   'diagonal_passes_per_step': sum(1 for m, _ in blocks if np.ndim(m) == 1),
                   'passes_per_step': len(blocks), 'full_size_passes_per_step': full_passes,
                   'schedule': 'fusion + lazy state growth (kron-joined sub-states), planned once outside the timed region',
                   'repetitions': reps,
                   'state_bytes': state_bytes,
                   'l2': 'inputs larger than L2 (state %.1f GB vs 126 MB)' % (state_bytes / 1e9)
                         if state_bytes > 252e6 else 'state fits L2; not an HBM measurement'},
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak_gbs, 'unit': 'GB/s',
                     'frac': achieved / peak_gbs, 'traffic': measured_traffic(n, dominant), 'kernel': dominant,
                     'peak_source': peak_src, 'bytes_per_launch': 2 * state_bytes,
                     'ms_per_launch': pass_ms, 'mean_ms_over_all_passes': mean_pass_ms,
                     'kernels': breakdown},
        'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': launches, 'clocks': clocks.summary(),
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='rqc30', choices=sorted(WORKLOADS) + ['rc_hbm'])
    ap.add_argument('--max-fused', dest='max_fused', type=int, default=None,
                    help='widest fused block; default = kernel-matched policy (5 for complex64)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == '__main__':
    main()
