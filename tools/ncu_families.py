"""One call of every kernel family that had no committed ncu capture yet
(VERDICT r1, g1): sampler chain, marginal, collapse / scale, norm, Pauli,
Kronecker join, bit permutation, reduced density matrix, density-matrix diagonal /
marginal / collapse / trace / Pauli, batched-trajectory kernels.

Run it under ncu on one GPU and summarise the report here (no GPU needed):

    ncu --nvtx --nvtx-include "b2q/" --clock-control none -o gpurun_out/families \
        --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... \
        python tools/ncu_families.py --n 28
    python tools/ncu_families.py --summarise gpurun_out/families.ncu-rep --n 28 \
        --out profiles/r2_ncu_families_summary.json

Without ncu the script prints CUDA-event timings of the same calls
(`--out gpurun_out/families_events.json`), which are the un-profiled numbers.
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def calls(n, DS, torch):
    """[(label, algorithmic bytes, fn)]; every fn launches the family's kernels once."""
    rng = np.random.RandomState(0)
    amp = 8 << n
    dev = DS.basis(n, np.complex64, 0)
    h = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    for b in range(n):  # a dense state: every amplitude non-zero
        dev.apply_matrix(h * np.exp(0.1j * b), [b])
    out = []
    out.append(('norm2 (sv_norm_partial_kernel + final_sum_kernel)', amp, lambda: dev.norm2()))
    out.append(('marginal over 3 bits (sv_marginal_kernel)', amp, lambda: dev.marginal_probs([n - 1, 7, 0])))
    out.append(('marginal over 12 bits (sv_marginal_kernel)', amp,
                lambda: dev.marginal_probs(list(range(n - 1, n - 13, -1)))))
    u = rng.random_sample(1_000_000)
    out.append(('sample 1M (sv_chunk_sums, block_sums, scan_inclusive, sv_sample_resolve)', amp,
                lambda: dev.sample_indices_device(u)))
    idx = dev.sample_indices_device(u)
    out.append(('unpack_bits 1M x n (unpack_bits_kernel)', 8 * u.size + n * u.size,
                lambda: DS.unpack_bits_device(idx, list(range(n - 1, -1, -1)))))
    out.append(('collapse 2 bits (sv_scale_mask_kernel)', 2 * amp, lambda: dev.collapse([5, n - 2], [0, 0], 0.25)))
    out.append(('scale (sv_scale_mask_kernel)', 2 * amp, lambda: dev.scale(2.0)))
    xm = (1 << 3) | (1 << (n - 2))
    zm = (1 << 3) | (1 << 9)
    out.append(('pauli expectation X..Y..Z (sv_pauli_partial_kernel)', 2 * amp, lambda: dev.pauli_expectation(xm, zm)))
    out.append(('pauli expectation Z-only (sv_pauli_partial_kernel)', amp, lambda: dev.pauli_expectation(0, zm)))
    zz = [(1 << int(a)) | (1 << int(b)) for a, b in rng.randint(0, n, size=(16, 2))]
    out.append(('16 Z-type Pauli strings in one pass (sv_pauli_multi_partial_kernel<DIAG>)', amp,
                lambda: dev.pauli_expectations(0, zz)))
    out.append(('16 Pauli strings sharing an X mask in one pass (sv_pauli_multi_partial_kernel)', 2 * amp,
                lambda: dev.pauli_expectations(xm, zz)))
    out.append(('reduced density matrix of 2 qubits (sv_reduced_dm_kernel)', amp,
                lambda: dev.reduced_density_matrix([n - 3, 4])))
    out.append(('reduced density matrix of 5 qubits (sv_reduced_dm_kernel)', amp,
                lambda: dev.reduced_density_matrix([n - 1, 11, 6, 3, 0])))
    out.append(('amplitude gather 4096 (sv_gather_kernel)', 8 * 4096,
                lambda: dev.amplitudes(rng.randint(0, 1 << n, 4096))))
    out.append(('argmax |amp| (sv_argmax kernels)', amp, lambda: dev.argmax_abs()))
    a = DS.basis(n - 10, np.complex64, 0)
    for b in range(n - 10):
        a.apply_matrix(h, [b])
    b10 = DS.basis(10, np.complex64, 3)
    for b in range(10):
        b10.apply_matrix(h, [b])
    out.append(('kron 2^(n-10) x 2^10 (sv_kron_kernel)', amp + (8 << (n - 10)), lambda: a.kron(b10)))
    perm = list(range(n))
    perm[0], perm[n - 1] = perm[n - 1], perm[0]
    perm[3], perm[12] = perm[12], perm[3]
    out.append(('permute bits out of place (sv_permute_bits_kernel)', 2 * amp, lambda: dev.permute_bits(perm)))
    rev = list(range(n))[::-1]
    out.append(('bit reversal IN PLACE, 3-4 tile passes (sv_permute_tile_kernel)', 2 * amp * 3,
                lambda: dev.permute_bits_inplace(rev)))
    u5a, u5b = np.linalg.qr(rng.standard_normal((32, 32)) + 1j * rng.standard_normal((32, 32)))[0], \
        np.linalg.qr(rng.standard_normal((32, 32)) + 1j * rng.standard_normal((32, 32)))[0]
    out.append(('two 5-qubit blocks in one pass (sv_apply_tc_tile_kernel)', 2 * amp,
                lambda: dev.apply_tile_blocks([(u5a, [n - 1, 17, 9, 5, 12]), (u5b, [22, 9, n - 4, 14, 3])])))
    other = dev.copy()
    out.append(('allclose (sv_allclose_kernel)', 2 * amp, lambda: dev.allclose(other, 1e-6)))
    joined = a.kron(b10)
    out.append(('kron allclose 2^(n-10) x 2^10 (sv_kron_mismatch_kernel)', amp + (8 << (n - 10)),
                lambda: joined.kron_allclose(a, b10, 1e-6)))
    # density matrix view of the same array: n/2 qubits
    nq = n // 2
    rho = DS.basis(2 * nq, np.complex64, 0)
    for b in range(2 * nq):
        rho.apply_matrix(h, [b])
    out.append(('dm diagonal (dm_diag_kernel)', 8 << nq, lambda: rho.dm_diagonal_device()))
    probs = rho.dm_diagonal_device()
    out.append(('probs marginal 4 of n/2 bits (probs_marginal_kernel)', 8 << nq,
                lambda: DS.probs_marginal_device(probs, nq, [nq - 1, 5, 2, 0])))
    out.append(('dm trace', 8 << nq, lambda: rho.dm_trace()))
    out.append(('dm collapse 1 qubit (dm collapse kernel)', 2 * (8 << (2 * nq)),
                lambda: rho.dm_collapse([3], [0], 0.5)))
    out.append(('dm pauli expectation (dm_pauli_partial_kernel)', 8 << nq, lambda: rho.dm_pauli_expectation(5, 6)))
    out.append(('dm partial trace keep 3 (dm partial trace kernel)', 8 << (2 * nq),
                lambda: rho.dm_partial_trace([nq - 1, 4, 0])))
    # batched trajectories: 2^(n-16) states of 16 qubits
    nb = 16
    batch = n - nb
    traj = DS.basis(n, np.complex64, 0)
    for b in range(n):
        traj.apply_matrix(h, [b])
    mats = np.stack([np.eye(2), np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.diag([1, -1])]).astype(complex)
    choice = rng.randint(0, 4, 1 << batch).astype(np.int32)
    out.append(('bsv apply select 1 qubit, 25% identity skipped (bsv_apply_select_kernel)', 2 * amp * 0.75,
                lambda: traj.bsv_apply_select(nb, mats, [7], choice, skip=0)))
    choices = rng.randint(0, 4, (nb, 1 << batch)).astype(np.int32)
    out.append(('bsv select multi: 16 one-qubit selections (bsv_select_multi_kernel)', 2 * amp,
                lambda: traj.bsv_apply_select_multi(nb, mats, list(range(nb)), choices, skip=0)))
    kr = np.stack([np.diag([1, np.sqrt(0.9)]), np.array([[0, np.sqrt(0.1)], [0, 0]])]).astype(complex)
    out.append(('bsv kraus weights, 2 operators (bsv_kraus_weights_kernel)', amp,
                lambda: traj.bsv_kraus_weights(nb, kr, [4])))
    vals = rng.randint(0, 2, (1 << batch, 2))
    out.append(('bsv collapse 2 bits (bsv_collapse_kernel)', 2 * amp,
                lambda: traj.bsv_collapse(nb, [3, 9], vals, np.full(1 << batch, 2.0))))
    # the gate kernels not covered by the bench's own captures
    d13 = np.exp(1j * rng.standard_normal(1 << 13))
    bits13 = [0, 2, 5, 9, 11, 12, 14, 17, 20, 23, 25, n - 2, n - 1]
    out.append(('13-wire diagonal block (sv_apply_diag_smem_kernel)', 2 * amp, lambda: dev.apply_diagonal(d13, bits13)))
    d15 = np.exp(1j * rng.standard_normal(1 << 15))
    out.append(('15-wire diagonal (sv_apply_diag_kernel)', 2 * amp,
                lambda: dev.apply_diagonal(d15, bits13 + [7, 16])))
    return out


def run(args):
    import torch

    from cirq_b200.device_state import DeviceState as DS

    rows = []
    only = [w for w in args.only.split(',') if w]
    for label, nbytes, fn in calls(args.n, DS, torch):
        if only and not any(w in label for w in only):
            continue
        fn()  # warm (allocations, attribute setup)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # (ncu --nvtx --nvtx-include "b2q/" profiles only the measured call, not the
        # state preparation or the warm call)
        torch.cuda.nvtx.range_push('b2q')
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        ms = a.elapsed_time(b)
        rows.append({'call': label, 'algorithmic_bytes': int(nbytes), 'ms_events': ms,
                     'GBps_events': nbytes / ms / 1e6})
        print(f'{label}: {ms:.3f} ms, {nbytes / ms / 1e6:.0f} GB/s', flush=True)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
        with open(args.out, 'w') as f:
            json.dump({'n_bits': args.n, 'calls': rows}, f, indent=1)


METRICS = {'gpu__time_duration.sum': 'duration', 'dram__bytes_read.sum': 'dram_read',
           'dram__bytes_write.sum': 'dram_write', 'launch__registers_per_thread': 'regs',
           'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_pct',
           'dram__throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct',
           'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
           'launch__grid_size': 'grid', 'launch__block_size': 'block'}
UNIT = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9,
        'Tbyte': 1e12}


def summarise(args):
    raw = subprocess.run(['ncu', '-i', args.summarise, '--page', 'raw', '--csv'], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    col = {name: i for i, name in enumerate(head)}
    per = {}
    for r in rows[2:]:
        name = r[col['Kernel Name']]
        rec = {}
        for metric, key in METRICS.items():
            if metric not in col:
                continue
            try:
                v = float(r[col[metric]].replace(',', ''))
            except ValueError:
                continue
            rec[key] = v * UNIT.get(units[col[metric]], 1.0)
        per.setdefault(name, []).append(rec)
    out = {}
    for name, recs in per.items():
        # skip the state-preparation launches: keep the LAST launch of each kernel
        # (the measured call runs after the warm one)
        r = recs[-1]
        traffic = r.get('dram_read', 0) + r.get('dram_write', 0)
        out[name] = {'launches_captured': len(recs), 'duration_ms': r.get('duration'),
                     'dram_read_bytes': r.get('dram_read'), 'dram_write_bytes': r.get('dram_write'),
                     'dram_GBps': traffic / r['duration'] / 1e6 if r.get('duration') else None,
                     'regs': r.get('regs'), 'sm_pct': r.get('sm_pct'), 'dram_pct': r.get('dram_pct'),
                     'warps_active_pct': r.get('warps_active_pct'), 'grid': r.get('grid'), 'block': r.get('block')}
    os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
    with open(args.out, 'w') as f:
        json.dump({'source': f'ncu --set full --clock-control none python tools/ncu_families.py --n {args.n}; '
                             'last captured launch of each kernel; durations under ncu are cold-cache and serialised',
                   'n_bits': args.n, 'state_bytes': 8 << args.n, 'kernels': out}, f, indent=1)
    print(f'{len(out)} kernels -> {args.out}')


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=28)
    ap.add_argument('--out', default='')
    ap.add_argument('--summarise', default='')
    ap.add_argument('--only', default='', help='comma-separated substrings of the call labels to run')
    a = ap.parse_args()
    summarise(a) if a.summarise else run(a)
