"""bench.py's output contract, checked without a GPU through the reference arm
(`--impl reference` times cirq.Simulator on the host): exactly one JSON line on
stdout carrying the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    proc = subprocess.run(
        [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'rc20',
         '--steps', '1', '--warmup', '0'],
        capture_output=True, text=True, timeout=600, cwd=ROOT,
    )
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, proc.stdout
    line = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step',
                'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in line, key
    assert line['impl'] == 'reference' and line['metric'] == 'fused_gates_per_s'
    assert line['value'] > 0 and line['higher_is_better'] is True
    assert line['config']['workload'] == 'rc20'
    assert line['cpu_baseline']['kind'] == 'reference' and line['cpu_baseline']['cores'] == 1
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0


def test_reference_arm_of_the_sharded_line_uses_the_sharded_workload():
    """`--impl reference --gpus N` (N > 1) mirrors the repo arm's N > 1 line: workload
    `rc_hbm` (cirq.testing.random_circuit, same seed) on a 24-qubit sample, counted in
    30-qubit-equivalent gates like cirq_b200/dist_bench.py; rank 0 alone prints."""
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    idle = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2',
                           '--steps', '1', '--warmup', '0'], capture_output=True, text=True, timeout=600,
                          cwd=ROOT, env=env)
    assert idle.returncode == 0 and idle.stdout.strip() == ''
    proc = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2',
                           '--steps', '1', '--warmup', '0'], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, proc.stdout
    line = json.loads(lines[0])
    assert line['impl'] == 'reference' and line['n_gpus'] == 2 and line['scaling'] == 'weak'
    cfg = line['config']
    assert cfg['workload'] == 'rc_hbm' and cfg['reference_sample'] == 'rc24' and cfg['n_qubits'] == 24
    assert cfg['extrapolated'] is True and cfg['n_qubits_of_the_repo_arm'] == 35
    assert abs(line['value'] - cfg['gates_per_s_on_sample'] * 2.0 ** (24 - 30)) < 1e-12
    assert line['e2e']['value'] == line['value'] and line['cpu_baseline']['value'] == line['value']


def test_workload_equivalent_units():
    sys.path.insert(0, ROOT)
    import bench

    # a gate pass over 2^24 amplitudes is 2^-6 of a 30-qubit gate
    value, n = bench.workload_equivalent(64.0, 24, 'rqc30')
    assert n == 30 and abs(value - 1.0) < 1e-12
    value, n = bench.workload_equivalent(5.0, 20, 'rc20')
    assert n == 20 and value == 5.0
    value, n = bench.workload_equivalent(64.0, 24, 'rc_hbm')  # the sharded lines count 30-qubit equivalents
    assert n == 30 and abs(value - 1.0) < 1e-12


def test_density_matrix_bench_schedule_matches_reference(cirq):
    """The device-resident leg of the qaoa workload replays a gate list built by
    bench.dm_gate_list; applied by the oracle it must give the reference's rho."""
    sys.path.insert(0, ROOT)
    import numpy as np

    import bench
    from cirq_b200.fusion import fuser_for
    from oracle import sv_oracle as orc

    bench.WORKLOADS['qaoa4'] = ('qaoa', dict(n=4, p=2, graph_seed=0, noise=0.01, resolvers=8), 10)
    wl = bench.build_workload('qaoa4')
    resolver = wl['resolvers'][3]
    f = fuser_for(np.complex128, 4, wl['bits'])
    for m, b in bench.dm_gate_list(cirq, wl, resolver):
        f.add(m, b)
    rho = np.zeros(1 << wl['bits'], dtype=np.complex128)
    rho[0] = 1
    for m, b in f.blocks():
        rho = orc.apply_matrix(rho, wl['bits'], m, list(b))
    body = cirq.Circuit(op for op in wl['circuit'].all_operations() if not cirq.is_measurement(op))
    want = cirq.DensityMatrixSimulator(dtype=np.complex128, noise=cirq.depolarize(0.01)).simulate(
        body, resolver, qubit_order=wl['qubits']).final_density_matrix
    np.testing.assert_allclose(rho.reshape(16, 16), want, atol=1e-12)
    assert wl['unit_gates'] == bench.ref_unit_gates(cirq, cirq.resolve_parameters(wl['circuit'], resolver))


def test_unit_gates_equal_reference_fuser_count(cirq):
    """The metric's gate unit, counted with cirq alone, equals this repo's own
    k <= 2 fusion of the same circuit (so both arms divide by the same number)."""
    sys.path.insert(0, ROOT)
    import bench
    from cirq_b200 import workloads as W
    from cirq_b200.fusion import fuse_gates

    for name in ('rqc20', 'rc20', 'qft22'):
        wl = bench.build_workload(name)
        assert wl['unit_gates'] == len(fuse_gates(W.circuit_to_gates(wl['circuit'], wl['qubits']), 2)), name
