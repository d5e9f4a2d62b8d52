"""Benchmark of the hot path: fused gates/s on BASELINE.json's configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload rqc30]

One "step" = one pass of the hot path over one synthetic circuit: |0..0> on the
device, every fused gate block applied (one streaming HBM pass each), and (for
the sampled configs) the bitstrings drawn.  Prints ONE JSON line.

N = 1: the headline (`value`, `roofline`, `e2e`, `cpu_baseline`) is BASELINE
config 2 (`rqc30`: 30-qubit Sycamore-style circuit, 1M samples); the same line
carries `configs`: one entry per other single-GPU BASELINE config, each with its
own `value`, `ms_per_step`, `roofline`, `e2e` and `cpu_baseline`:
  rc20    config 1  cirq.testing.random_circuit, 20 qubits (lives in L2: latency-bound)
  qft34   config 3  34-qubit QFT, 137 GB state
  qaoa16  config 5  16-qubit noisy QAOA density matrix (34 GB rho), one resolver per step
  rqc24   the sample the reference arm runs, so that ONE ratio is same-config
N > 1 (`cirq_b200/dist_bench.py`): BASELINE config 4, the sharded state vector
with 34 local qubits per GPU (35 / 36 / 37 qubits on 2 / 4 / 8 GPUs).

value     fused gates/s with all inputs resident in HBM (plan pre-built, uniforms
          uploaded before the timed region).  The gate unit is the k<=2 fused
          block — the operation count of cirq.merge_k_qubit_unitaries(circuit,
          k=2) (245 for the 30-qubit config) — so numbers are comparable with
          BASELINE.md even though this backend fuses wider and needs fewer passes.
e2e       the same metric through the public API a Cirq user calls
          (B200Simulator.run / compute_amplitudes, B200DensityMatrixSimulator
          .run_sweep) on the cirq.Circuit: host scheduling, uploads and the
          device->host copy of the result inside the timed region.  `first_call_ms`
          is the cold call (per-gate unitary cache empty), `second_call_ms` the call
          that records the circuit's schedule (cirq_b200/plan_cache.py), `ms_per_step`
          the calls after it (schedule replayed; every gate pass still runs on the
          GPU), `ms_per_step_without_schedule_cache` the same with plan_cache=False.
roofline  dominant in-step kernel: algorithmic bytes per launch
          (2 * sizeof(complex) * 2^bits) / mean launch duration measured with CUDA
          events around the gate passes, vs the measured copy peak.
cpu_baseline  the reference (cirq.Simulator / cirq.DensityMatrixSimulator) on the
          box's host cores on a bounded sample of the same generator (fewer
          qubits), timed in the same run.  Its gates are counted in the workload's
          own unit: a gate pass over 2^m amplitudes counts as 2^(m-n) gates of the
          n-qubit workload; the raw number measured on the sample is reported next
          to it.  The reference's time per gate grows at least linearly in 2^n, so
          this favours the reference.

`--impl reference` runs ONLY the unmodified reference (no library of this repo is
loaded): its top-level value is measured on a sample (`same_config: false`,
`extrapolated: true` in `config`), and `configs.rqc24` is the like-for-like entry
(`same_config: true`) that the repo arm also reports.  With `--gpus N` > 1 it mirrors
the sharded line instead: workload `rc_hbm` (cirq.testing.random_circuit, same seed) on a
24-qubit sample, counted in 30-qubit-equivalent gates like `cirq_b200/dist_bench.py`.
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
    'rqc24': ('rqc', dict(rows=4, cols=6, depth=20, seed=1), 10_000),
    'rqc20': ('rqc', dict(rows=4, cols=5, depth=20, seed=1), 10_000),
    'qft34': ('qft', dict(n=34), 0),
    'qft30': ('qft', dict(n=30), 0),
    'qft24': ('qft', dict(n=24), 0),
    'qft22': ('qft', dict(n=22), 0),
    'rc20': ('rc', dict(n=20, depth=20, seed=1234), 0),
    # sample of the sharded workload `rc_hbm` (same generator and seed, 24 qubits) for the reference arm
    'rc24': ('rc', dict(n=24, depth=20, seed=1234), 0),
    # the headline circuit in complex128 (17 GB state, register-tiled fp64 kernels: no
    # tensor-core path for double precision)
    'rqc30_c128': ('rqc', dict(rows=5, cols=6, depth=20, seed=1), 1_000_000),
    'qaoa16': ('qaoa', dict(n=16, p=2, graph_seed=0, noise=0.01, resolvers=256), 1000),
    'qaoa12': ('qaoa', dict(n=12, p=2, graph_seed=0, noise=0.01, resolvers=256), 1000),
    'qaoa10': ('qaoa', dict(n=10, p=2, graph_seed=0, noise=0.01, resolvers=256), 1000),
    'qaoa8': ('qaoa', dict(n=8, p=2, graph_seed=0, noise=0.01, resolvers=256), 1000),
}
# bounded CPU samples (same generator, fewer qubits): ~5-20 s of reference time in all
CPU_SAMPLE = {'rc24': 'rc20', 'rqc30': 'rqc20', 'rqc30_c128': 'rqc20', 'rqc24': 'rqc20', 'rqc20': 'rqc20', 'qft34': 'qft22', 'qft30': 'qft22',
              'qft24': 'qft22', 'qft22': 'qft22', 'rc20': 'rc20', 'qaoa16': 'qaoa10', 'qaoa12': 'qaoa10',
              'qaoa10': 'qaoa10', 'qaoa8': 'qaoa8'}
# the reference arm (`--impl reference`) has minutes, not seconds: a larger sample
REFERENCE_ARM_SAMPLE = dict(CPU_SAMPLE, rqc30='rqc24', qft34='qft24', qft30='qft24')
WORKLOAD_DTYPE = {'rqc30_c128': np.complex128}
# sub-entries of the N = 1 line
SUB_CONFIGS = ('rc20', 'rqc24', 'qft34', 'qaoa16', 'rqc30_c128')


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


def workload_bits(workload):
    """Index bits of the workload's state (2n for a density matrix)."""
    kind, params, _ = WORKLOADS[workload]
    if kind == 'rqc':
        return params['rows'] * params['cols']
    return 2 * params['n'] if kind == 'qaoa' else params['n']


SHARDED_EQUIVALENT_BITS = 30  # the N > 1 lines count 30-qubit-equivalent gates (cirq_b200/dist_bench.py)


def workload_equivalent(value, bits_sample, workload):
    """gates/s measured on a sample with 2^bits_sample amplitudes -> gates/s in
    units of the workload's state size (a pass over 2^m amplitudes = 2^(m-n)
    workload gates); the sharded workloads are counted in 30-qubit equivalents."""
    n = SHARDED_EQUIVALENT_BITS if workload in ('rc_hbm', 'rqc_weak') else workload_bits(workload)
    return value * 2.0 ** (bits_sample - n), n


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def measured_traffic(n_bits, kernel):
    """dram read+write bytes per launch of `kernel` from the committed ncu --set
    full captures (profiles/ncu_traffic.json: measured on 30-bit states; traffic
    scales with the state size, so other sizes are scaled and marked)."""
    path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if not os.path.exists(path):
        return None
    with open(path) as f:
        table = json.load(f)['bytes_per_launch_30q']
    base = kernel.split(' ')[0]
    got = table.get(kernel, table.get(base))
    if got is None:
        return None
    return int(got * 2.0 ** (n_bits - 30))


def kernel_class(wires, n_bits=30, real='float'):
    """Name of the kernel a fused block on these index bits runs on (default
    knobs: b2q_apply_tc.cu launch_tc_k / b2q_apply.cu apply_matrix_t)."""
    k = len(wires)
    if real == 'double':
        return f'sv_apply_fast_kernel<double,{k}>'
    tc4 = os.environ.get('CIRQ_B200_TC_MODE') == '2'  # opt-in: 4-qubit blocks on tcgen05 too
    if n_bits >= k + 7 and (k == 5 or (k == 4 and tc4)):
        return f'sv_apply_tc_staged_kernel<{k}>' if min(wires) < 2 else f'sv_apply_tc_kernel<{k}>'
    if k == 6 and n_bits >= 13:
        return 'sv_apply_tc_kernel<6>'
    return f'sv_apply_fast_kernel<float,{k}>'


def block_kernel_name(m, w, n_bits, real='float'):
    """(kernel name, blocks in the launch) of one scheduled block."""
    if np.ndim(m) == 1:
        return f'sv_apply_diag_smem_kernel<{real}>', 1
    return kernel_class(w, n_bits, real), 1


def pass_kernel_name(group, n_bits, real='float'):
    """Kernel of one pass from DeviceState.plan_passes: a single block, or two
    blocks in one tile pass."""
    if len(group) == 2:
        return 'sv_apply_tc_tile_kernel (2 blocks per pass)'
    return block_kernel_name(group[0][0], group[0][1], n_bits, real)[0]


def ref_unit_gates(cirq, circuit):
    """Gate unit of the metric: operations left by the reference's own fuser,
    cirq.merge_k_qubit_unitaries(k=2) (transformers/merge_k_qubit_gates.py:70-114),
    on the circuit without its measurements.  Needs nothing but cirq."""
    body = cirq.Circuit(op for op in circuit.all_operations() if not cirq.is_measurement(op))
    return sum(1 for _ in cirq.merge_k_qubit_unitaries(body, k=2).all_operations())


def build_workload(name):
    """Returns dict(kind, circuit, qubits, n, bits, reps, unit_gates, ...); cirq only."""
    from cirq_b200 import workloads as W
    from cirq_b200._cirq_compat import import_cirq

    cirq = import_cirq()
    kind, params, reps = WORKLOADS[name]
    out = dict(kind=kind, name=name, reps=reps, generator='cirq ' + kind)
    if kind == 'qaoa':
        circuit, qubits, names = W.qaoa_circuit(params['n'], params['p'], params['graph_seed'])
        resolvers = list(cirq.to_resolvers(W.qaoa_sweep(names, params['resolvers'])))
        out.update(circuit=circuit, qubits=list(qubits), resolvers=resolvers, noise_p=params['noise'],
                   n=len(qubits), bits=2 * len(qubits),
                   unit_gates=ref_unit_gates(cirq, cirq.resolve_parameters(circuit, resolvers[0])))
        return out
    if kind == 'rqc':
        circuit, qubits = W.rqc_circuit(**params)
    elif kind == 'qft':
        circuit, qubits = W.qft_circuit(**params)
    else:
        circuit, qubits = W.random_circuit(**params)
    out.update(circuit=circuit, qubits=list(qubits), n=len(qubits), bits=len(qubits),
               unit_gates=ref_unit_gates(cirq, circuit))
    return out


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


# ---- the reference on the host cores -------------------------------------------------------

_REF_CACHE: dict = {}


def time_reference(name, steps, warmup, dtype=np.complex64):
    """Times the unmodified reference (cirq.Simulator / cirq.DensityMatrixSimulator,
    numpy) on the host; nothing of this repo's library is involved."""
    key = (name, steps, warmup, np.dtype(dtype).name)
    if key in _REF_CACHE:
        return _REF_CACHE[key]
    from cirq_b200._cirq_compat import import_cirq

    cirq = import_cirq()
    wl = build_workload(name)
    circuit, qubits, reps = wl['circuit'], wl['qubits'], wl['reps']
    if wl['kind'] == 'qaoa':
        resolvers = wl['resolvers']
        state = {'i': 0}

        def step():
            sim = cirq.DensityMatrixSimulator(dtype=dtype, noise=cirq.depolarize(wl['noise_p']), seed=0)
            r = resolvers[state['i'] % len(resolvers)]
            state['i'] += 1
            sim.run_sweep(circuit, [r], repetitions=reps)
    else:
        run_circuit = circuit + cirq.Circuit(cirq.measure(*qubits, key='m')) if reps else circuit

        def step():
            sim = cirq.Simulator(dtype=dtype, seed=0)
            if reps:
                sim.run(run_circuit, repetitions=reps)
            else:
                sim.simulate(run_circuit, qubit_order=qubits)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    out = dict(value=wl['unit_gates'] / dt, seconds_per_step=dt, unit_gates=wl['unit_gates'], n=wl['n'],
               bits=wl['bits'], raw_ops=sum(1 for _ in circuit.all_operations()), reps=reps, steps=steps)
    _REF_CACHE[key] = out
    return out


def cpu_baseline_for(workload, steps=3):
    """cpu_baseline object of `workload`: the reference on CPU_SAMPLE[workload]."""
    try:
        sample = CPU_SAMPLE[workload]
        dtype = WORKLOAD_DTYPE.get(workload, np.complex64)
        r = time_reference(sample, steps, 0, dtype)
        eq, n = workload_equivalent(r['value'], r['bits'], workload)
        api = 'cirq.DensityMatrixSimulator(noise=depolarize).run_sweep, one resolver per step' \
            if WORKLOADS[sample][0] == 'qaoa' else f'cirq.Simulator({np.dtype(dtype).name})'
        return {'value': eq, 'unit': 'gates/s', 'cores': 1, 'kind': 'reference',
                'raw_value_on_sample': r['value'], 'same_config': sample == workload,
                'sample': f"{api} on {sample}: {r['n']} qubits, {r['raw_ops']} ops = {r['unit_gates']} k<=2 "
                          f"blocks, {r['reps']} repetitions, {r['steps']} x {r['seconds_per_step']:.2f} s: "
                          f"{r['value']:.3g} gates/s there, counted as {n}-bit-equivalent gates "
                          f"(x 2^({r['bits']}-{n})); single-threaded numpy, host has {os.cpu_count()} cores"}
    except Exception as exc:  # reference not importable on this box
        return {'value': None, 'unit': 'gates/s', 'cores': 1, 'kind': 'reference',
                'sample': f'unavailable: {exc!r}'}


def reference_entry(sample, workload, steps, warmup):
    """One reference measurement as a (sub-)line: value in `workload` units."""
    r = time_reference(sample, steps, warmup)
    value, n = workload_equivalent(r['value'], r['bits'], workload)
    same = sample == workload
    return {
        'impl': 'reference', 'metric': 'fused_gates_per_s', 'value': value, 'unit': 'gates/s',
        'steps': steps, 'warmup': warmup, 'ms_per_step': r['seconds_per_step'] * 1e3,
        'higher_is_better': True, 'dtype': 'c64', 'data': 'synthetic',
        'config': {'workload': workload, 'gate_unit': 'k<=2 fused blocks (cirq.merge_k_qubit_unitaries(k=2) count)',
                   'same_config': same, 'extrapolated': not same,
                   'reference_sample': sample, 'n_qubits': r['n'], 'raw_ops': r['raw_ops'],
                   'unit_gates': r['unit_gates'], 'repetitions': r['reps'],
                   'gates_per_s_on_sample': r['value'],
                   'value_unit': ('gates/s on the workload itself' if same else
                                  f"{n}-bit-equivalent gates/s = gates/s on the {r['bits']}-bit sample "
                                  f"x 2^({r['bits']}-{n})")},
        'cpu_baseline': {'value': value, 'unit': 'gates/s', 'cores': 1, 'kind': 'reference',
                         'raw_value_on_sample': r['value'], 'same_config': same,
                         'sample': f"the reference on {sample} ({r['n']} qubits, same generator, "
                                   f"{r['seconds_per_step']:.1f} s per step): {r['value']:.3g} gates/s there"
                                   + ('' if same else f", counted as {n}-bit-equivalent gates "
                                                      f"(x 2^({r['bits']}-{n}))")
                                   + f"; numpy path is single-threaded; host has {os.cpu_count()} cores"},
        'e2e': {'value': value, 'unit': 'gates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    if args.workload == 'rc_hbm' or (args.workload is None and args.gpus > 1):
        # the repo arm's N > 1 workload: cirq.testing.random_circuit at 34 + log2(N) qubits,
        # counted in 30-qubit-equivalent gates; here the same generator and seed at 24 qubits
        workload, sample = 'rc_hbm', 'rc24'
        steps, warmup = max(1, min(args.steps, 3)), 0
    else:
        workload = 'rqc30' if args.workload in (None, 'rqc_weak') else args.workload
        sample = REFERENCE_ARM_SAMPLE[workload]
        big = sample != CPU_SAMPLE[workload]  # ~55 s per step: keep the run within minutes
        steps = max(1, min(args.steps, 2 if big else 3))
        warmup = 0 if big else min(args.warmup, 1)
    line = reference_entry(sample, workload, steps, warmup)
    line.update({'n_gpus': args.gpus, 'scaling': 'weak', 'vs_baseline': None, 'gpu_launches': 0})
    if workload == 'rc_hbm':
        line['config']['n_qubits_of_the_repo_arm'] = 34 + max(args.gpus, 1).bit_length() - 1
    elif sample != workload:
        # the like-for-like entry: the repo arm reports configs[sample] on the SAME
        # circuit, repetitions and gate unit (no extrapolation)
        sub = reference_entry(sample, sample, steps, warmup)
        line['configs'] = {sample: sub}
    emit(line)


# ---- the B200 arm, one GPU -----------------------------------------------------------------


def _roofline(per_kernel, record_steps, bytes_per_pass, n_bits, peak_gbs, peak_src):
    """`roofline` object of one measurement.  The unit of the metric is the fused gate
    pass: 2 * sizeof(complex) * 2^bits algorithmic bytes each.  A launch
    of the tile kernel processes TWO units while moving the state once, so its
    `achieved` (algorithmic bytes per launch / launch time, the definition the
    contract gives) is reported next to the rate of the bytes it really moves."""
    total_ms = sum(np.sum(v) for v in per_kernel.values())
    count = sum(len(v) for v in per_kernel.values())
    units = {k: (2 if '2 blocks per pass' in k else 1) for k in per_kernel}
    breakdown = {k: {'launches_per_step': len(v) // record_steps, 'ms_per_launch': float(np.mean(v)),
                     'fused_blocks_per_launch': units[k],
                     'share_of_gate_time': float(np.sum(v) / total_ms)} for k, v in per_kernel.items()}
    dominant = max(breakdown, key=lambda k: breakdown[k]['share_of_gate_time'])
    pass_ms = breakdown[dominant]['ms_per_launch']
    u = units[dominant]
    achieved = u * bytes_per_pass / (pass_ms * 1e-3) / 1e9
    moved = bytes_per_pass / (pass_ms * 1e-3) / 1e9
    total_units = sum(units[k] * len(v) for k, v in per_kernel.items())
    return {'bound': 'hbm', 'achieved': achieved, 'peak': peak_gbs, 'unit': 'GB/s',
            'frac': achieved / peak_gbs, 'traffic': measured_traffic(n_bits, dominant), 'kernel': dominant,
            'peak_source': peak_src, 'bytes_per_launch': u * bytes_per_pass,
            'fused_blocks_per_launch': u, 'algorithmic_bytes_per_fused_block': bytes_per_pass,
            'hbm_bytes_moved_per_launch': bytes_per_pass,
            'frac_of_peak_by_bytes_moved': moved / peak_gbs,
            'note': ('this kernel applies 2 fused blocks per pass over HBM: it moves half the algorithmic '
                     'bytes, so `achieved` (algorithmic bytes / time) can exceed what the memory system '
                     'carries; `frac_of_peak_by_bytes_moved` is the rate of the real traffic — the kernel is '
                     'bound by its tensor-core / TMEM round trips, not by HBM') if u > 1 else None,
            'ms_per_launch': pass_ms, 'mean_ms_over_all_passes': float(total_ms / max(1, count)),
            'ms_per_fused_block_over_all_passes': float(total_ms / max(1, total_units)),
            'frac_over_all_passes': float(total_units * bytes_per_pass / (total_ms * 1e-3) / 1e9 / peak_gbs),
            'kernels': breakdown}, count // record_steps


def _time_e2e(fn, steps, settle=0):
    """(seconds per call over `steps` calls, seconds of the first call, seconds of each
    of the `settle` untimed calls in between)."""
    import torch

    def once():
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    first = once()
    settled = [once() for _ in range(settle)]
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps, first, settled


def measure_sv(name, steps, warmup, args, local_rank=0):
    """One state-vector workload on one GPU: device-resident value, per-kernel
    roofline, e2e through B200Simulator, cpu_baseline."""
    import ctypes

    import torch

    from cirq_b200 import _lib, workloads as W
    from cirq_b200.plan import build_plan, replay_plan

    lib = _lib.load()
    peak_gbs, peak_src = load_peaks()
    wl = build_workload(name)
    n, reps = wl['n'], wl['reps']
    gates = W.circuit_to_gates(wl['circuit'], wl['qubits'])
    unit_gates = wl['unit_gates']
    dtype = WORKLOAD_DTYPE.get(name, np.complex64)
    amp_bytes = np.dtype(dtype).itemsize
    # Host scheduling happens ONCE, outside the timed region: fusion + lazy state
    # growth (sub-states joined by the kron kernel, as with split_untangled_states).
    plan = build_plan(n, gates, dtype, args.max_fused)
    blocks = [blk for op in plan['ops'] if op[0] == 'apply' for blk in op[2]]
    state_bytes = (amp_bytes << n)
    rng = np.random.RandomState(0)
    uniforms = rng.random_sample(max(reps, 1))
    u_dev = torch.from_numpy(uniforms).to('cuda')
    ws_bytes = int(lib.b2q_sv_sample_workspace_bytes(n, reps)) if reps else 16
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device='cuda')
    out_idx = torch.empty(max(reps, 1), dtype=torch.int64, device='cuda')
    out_bits = torch.empty((max(reps, 1), n if reps else 1), dtype=torch.uint8, device='cuda')
    # column a of the samples = logical qubit axis a = logical bit n-1-a
    bits_order = _lib.int_array([plan['bit_of'][n - 1 - a] for a in range(n)])
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    per_kernel: dict = {}
    pairs: list = []

    def timed_apply(state, blks):
        for group in state.plan_passes(blks):  # one launch each: a block, or two in a tile pass
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            state.apply_batch(group)
            b.record()
            pairs.append((state.n_bits, pass_kernel_name(group, state.n_bits,
                                                         'double' if amp_bytes == 16 else 'float'), a, b))

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
            for nb, kname, a, b in pairs:
                if nb == n:  # launches on the full-size state: the HBM-bound ones
                    per_kernel.setdefault(kname, []).append(a.elapsed_time(b))
            pairs.clear()
        del dev

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    launches0 = int(lib.b2q_launch_count())
    with ClockSampler(local_rank) as clocks:
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        start.record()
        for _ in range(steps):
            step()
        end.record()
        torch.cuda.synchronize()
        ms_per_step = start.elapsed_time(end) / steps
        launches = (int(lib.b2q_launch_count()) - launches0) // max(steps, 1)
        # per-kernel duration of the gate passes (separate, event-bracketed steps)
        record_steps = min(3, steps)
        for _ in range(record_steps):
            step(record=True)
    roofline, full_passes = _roofline(per_kernel, record_steps, 2 * state_bytes, n, peak_gbs, peak_src)
    value = unit_gates / (ms_per_step * 1e-3)
    del u_dev, ws, out_idx, out_bits
    torch.cuda.empty_cache()

    # ---- e2e through the public Cirq-facing API -------------------------------------------
    from cirq_b200._cirq_compat import import_cirq

    cirq = import_cirq()
    import cirq_b200

    circuit = wl['circuit']
    if reps:
        circuit = circuit + cirq.Circuit(cirq.measure(*wl['qubits'], key='m'))
    e2e_steps = max(1, min(steps, 3))

    def e2e_step(plan_cache=True):
        sim = cirq_b200.B200Simulator(dtype=dtype, seed=0, max_fused_qubits=args.max_fused, plan_cache=plan_cache)
        if reps:
            res = sim.run(circuit, repetitions=reps)
            return res.measurements['m'].shape
        # the reference's own API for reading a few amplitudes of a state too
        # large to download (SimulatesAmplitudes, sim/simulator.py:120-182)
        return sim.compute_amplitudes(circuit, [0, 1], qubit_order=wl['qubits'])

    # call 1 is cold (per-gate unitary cache empty), call 2 records the circuit's schedule
    # (cirq_b200/plan_cache.py: the second sighting of a circuit), the timed calls replay
    # it; the same calls with the schedule cache off are timed beside them
    from cirq_b200 import plan_cache as _pc

    _pc.CACHE.clear()
    dt, first, settled = _time_e2e(e2e_step, e2e_steps, settle=1)
    cache_hits = _pc.CACHE.hits
    _pc.CACHE.clear()
    dt_nocache, _, _ = _time_e2e(lambda: e2e_step(plan_cache=False), max(1, min(e2e_steps, 2)))
    mat_bytes = int(sum(16 * np.size(m) for m, _ in _flat_blocks(blocks)))
    e2e = {'value': unit_gates / dt, 'unit': 'gates/s', 'ms_per_step': dt * 1e3, 'first_call_ms': first * 1e3,
           'second_call_ms': settled[0] * 1e3, 'schedule_cache_hits': cache_hits,
           'ms_per_step_without_schedule_cache': dt_nocache * 1e3,
           'h2d_bytes_per_step': int(8 * reps + mat_bytes),
           'd2h_bytes_per_step': int(reps * n if reps else 32),
           'api': 'cirq_b200.B200Simulator(seed=0).run(circuit, repetitions)' if reps
                  else 'cirq_b200.B200Simulator().compute_amplitudes(circuit, [0, 1])'}
    torch.cuda.empty_cache()

    flat = _flat_blocks(blocks)
    dense_widths = [len(w) for m, w in flat if np.ndim(m) == 2]
    return {
        'metric': 'fused_gates_per_s', 'value': value, 'unit': 'gates/s', 'steps': steps, 'warmup': warmup,
        'ms_per_step': ms_per_step, 'dtype': 'c128' if amp_bytes == 16 else 'c64',
        'config': {'workload': name, 'generator': wl['generator'], 'n_qubits': n,
                   'raw_ops': len(gates),
                   'gate_unit': 'k<=2 fused blocks (cirq.merge_k_qubit_unitaries(k=2) count)',
                   'unit_gates': unit_gates, 'max_fused_qubits': max(dense_widths) if dense_widths else 0,
                   'diagonal_passes_per_step': sum(1 for m, _ in flat if np.ndim(m) == 1),
                   'blocks_per_step': len(flat),
                   'full_size_passes_per_step': full_passes,
                   'schedule': 'fusion + lazy state growth (kron-joined sub-states) + tile groups, planned '
                               'once outside the timed region',
                   'repetitions': reps, 'state_bytes': state_bytes,
                   'l2': 'inputs larger than L2 (state %.1f GB vs 126 MB)' % (state_bytes / 1e9)
                         if state_bytes > 252e6 else 'state fits L2; not an HBM measurement'},
        'roofline': roofline, 'cpu_baseline': None if args.no_cpu_baseline else cpu_baseline_for(name),
        'e2e': e2e, 'gpu_launches': launches, 'clocks': clocks.summary(),
    }


def _flat_blocks(blocks):
    return list(blocks)


def dm_gate_list(cirq, wl, resolver):
    """[(matrix, bits)] over the 2n index bits of rho for one resolver, exactly what
    B200DensityMatrixSimulator queues (U on the row bits + conj(U) on the column
    bits, a channel as its superoperator), built by the sweep planner's walk."""
    from cirq_b200.sweeps import SweepPlan

    resolved = cirq.resolve_parameters(wl['circuit'], resolver)
    body = cirq.Circuit(op for op in resolved.all_operations() if not cirq.is_measurement(op))
    noise = cirq.ConstantQubitNoiseModel(cirq.depolarize(wl['noise_p']))
    plan = SweepPlan('dm', wl['qubits'], [resolver])
    for moment in noise.noisy_moments(body, sorted(body.all_qubits())):
        for op in cirq.flatten_to_ops(moment):
            if not plan.add_op(op):
                raise RuntimeError(f'cannot schedule {op!r}')
    return [(item[1], list(item[2])) for item in plan.items]


def measure_dm(name, steps, warmup, args, local_rank=0):
    """Config 5: noisy QAOA on a density matrix; one step = one resolver of the
    sweep (rho evolved from |0><0| + `repetitions` samples of the diagonal)."""
    import torch

    import cirq_b200
    from cirq_b200 import _lib
    from cirq_b200._cirq_compat import import_cirq
    from cirq_b200.device_state import DeviceState
    from cirq_b200.fusion import fuser_for

    cirq = import_cirq()
    lib = _lib.load()
    peak_gbs, peak_src = load_peaks()
    wl = build_workload(name)
    n, bits, reps = wl['n'], wl['bits'], wl['reps']
    dtype = np.complex64
    resolvers = wl['resolvers']
    unit_gates = wl['unit_gates']
    max_fused = 4 if args.max_fused is None else args.max_fused

    def schedule(resolver):
        f = fuser_for(dtype, max_fused, bits)
        for m, b in dm_gate_list(cirq, wl, resolver):
            f.add(m, b)
        return f.blocks()

    # one schedule per step's resolver, built outside the timed region
    plans = [schedule(resolvers[(i * 37) % len(resolvers)]) for i in range(min(4, max(steps, 1)))]
    state_bytes = 8 << bits
    rng = np.random.RandomState(0)
    uniforms = rng.random_sample(reps)
    meas_bits = list(range(n - 1, -1, -1))
    per_kernel: dict = {}
    pairs: list = []

    def step(i, record=False):
        dev = DeviceState.basis(bits, dtype, 0)
        for group in dev.plan_passes(plans[i % len(plans)]):
            if record:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
            dev.apply_batch(group)
            if record:
                b.record()
                pairs.append((pass_kernel_name(group, bits), a, b))
        probs = dev.dm_diagonal_device()
        idx = DeviceState.cdf_sample_device(DeviceState.probs_marginal_device(probs, n, meas_bits), uniforms)
        out = DeviceState.unpack_bits_device(idx, [n - 1 - c for c in range(n)])
        if record:
            torch.cuda.synchronize()
            for kname, a, b in pairs:
                per_kernel.setdefault(kname, []).append(a.elapsed_time(b))
            pairs.clear()
        del dev, out

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    launches0 = int(lib.b2q_launch_count())
    with ClockSampler(local_rank) as clocks:
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for i in range(steps):
            step(i)
        end.record()
        torch.cuda.synchronize()
        ms_per_step = start.elapsed_time(end) / steps
        launches = (int(lib.b2q_launch_count()) - launches0) // max(steps, 1)
        record_steps = min(2, steps)
        for i in range(record_steps):
            step(i, record=True)
    roofline, passes = _roofline(per_kernel, record_steps, 2 * state_bytes, bits, peak_gbs, peak_src)
    value = unit_gates / (ms_per_step * 1e-3)
    torch.cuda.empty_cache()

    # e2e: the public sweep API on the first resolvers (measurement records to the host)
    e2e_res = 2
    circuit = wl['circuit']

    def e2e_step():
        sim = cirq_b200.B200DensityMatrixSimulator(noise=cirq.depolarize(wl['noise_p']), seed=0, dtype=dtype)
        res = sim.run_sweep(circuit, resolvers[:e2e_res], repetitions=reps)
        return [r.measurements['m'].shape for r in res]

    dt, first, _ = _time_e2e(e2e_step, 1)
    dt /= e2e_res
    mat_bytes = int(sum(16 * np.size(m) for m, _ in _flat_blocks(plans[0])))
    e2e = {'value': unit_gates / dt, 'unit': 'gates/s', 'ms_per_step': dt * 1e3,
           'first_call_ms': first * 1e3 / e2e_res,
           'h2d_bytes_per_step': int(8 * reps + mat_bytes), 'd2h_bytes_per_step': int(reps * n),
           'api': f'cirq_b200.B200DensityMatrixSimulator(noise=depolarize({wl["noise_p"]}), seed=0)'
                  f'.run_sweep(circuit, resolvers[:{e2e_res}], repetitions={reps}), per resolver'}
    torch.cuda.empty_cache()
    flat = _flat_blocks(plans[0])
    return {
        'metric': 'fused_gates_per_s', 'value': value, 'unit': 'gates/s', 'steps': steps, 'warmup': warmup,
        'ms_per_step': ms_per_step, 'dtype': 'c64',
        'config': {'workload': name, 'generator': 'examples/qaoa.py qaoa_max_cut_circuit, random 3-regular graph',
                   'n_qubits': n, 'state_bits': bits, 'noise': f'depolarize({wl["noise_p"]}) after every moment',
                   'step': 'one resolver of the 256-point sweep: rho from |0><0| + sampling',
                   'resolvers_in_sweep': len(resolvers),
                   'resolvers_per_s': 1e3 / ms_per_step,
                   'full_sweep_seconds_estimate': len(resolvers) * ms_per_step / 1e3,
                   'gate_unit': 'k<=2 fused blocks of the noiseless circuit (cirq.merge_k_qubit_unitaries(k=2) count)',
                   'unit_gates': unit_gates, 'blocks_per_step': len(flat),
                   'full_size_passes_per_step': passes,
                   'max_fused_bits': max(len(w) for m, w in flat if np.ndim(m) == 2),
                   'repetitions': reps, 'state_bytes': state_bytes,
                   'l2': 'inputs larger than L2 (rho %.1f GB vs 126 MB)' % (state_bytes / 1e9)},
        'roofline': roofline, 'cpu_baseline': None if args.no_cpu_baseline else cpu_baseline_for(name, steps=1),
        'e2e': e2e, 'gpu_launches': launches, 'clocks': clocks.summary(),
    }


def measure(name, steps, warmup, args, local_rank=0):
    if WORKLOADS[name][0] == 'qaoa':
        return measure_dm(name, steps, warmup, args, local_rank)
    return measure_sv(name, steps, warmup, args, local_rank)


def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        from cirq_b200 import dist_bench

        dist_bench.run(args, world, rank, local_rank)
        return

    workload = args.workload or 'rqc30'
    if workload in ('rc_hbm', 'rqc_weak'):
        raise SystemExit(f'--workload {workload} is the sharded path: launch with --gpus N > 1 under torchrun')
    head = measure(workload, args.steps, args.warmup, args, local_rank)
    line = {
        'metric': head['metric'], 'value': head['value'], 'unit': head['unit'], 'n_gpus': 1,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': head['ms_per_step'],
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'c64',
        'data': 'synthetic', 'config': head['config'], 'roofline': head['roofline'],
        'cpu_baseline': head['cpu_baseline'], 'e2e': head['e2e'], 'gpu_launches': head['gpu_launches'],
        'clocks': head['clocks'],
    }
    if not args.no_configs and args.workload is None:
        # The other single-GPU BASELINE configs, each a full measurement of its own
        # (fewer steps: a 34-qubit step moves 7 TB).  A failure is recorded, not fatal.
        configs = {}
        sub_steps = max(3, min(args.steps, 5))
        sub_warm = max(3, min(args.warmup, 3))
        for sub in SUB_CONFIGS:
            t0 = time.perf_counter()
            try:
                configs[sub] = measure(sub, sub_steps, sub_warm, args, local_rank)
            except Exception as exc:  # keep the headline even if a sub-config cannot run here
                configs[sub] = {'error': repr(exc)}
                torch.cuda.empty_cache()
            configs[sub]['wall_s'] = time.perf_counter() - t0
        line['configs'] = configs
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default=None, choices=sorted(WORKLOADS) + ['rc_hbm', 'rqc_weak'],
                    help='default: rqc30 (+ the other configs as sub-entries) on 1 GPU, rc_hbm on N > 1')
    ap.add_argument('--max-fused', dest='max_fused', type=int, default=None,
                    help='widest fused block; default = kernel-matched policy (5 for complex64)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='headline workload only')
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == '__main__':
    main()
