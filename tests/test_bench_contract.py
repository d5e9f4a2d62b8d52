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


def test_workload_equivalent_units():
    sys.path.insert(0, ROOT)
    import bench

    # a gate pass over 2^24 amplitudes is 2^-6 of a 30-qubit gate
    value, n = bench.workload_equivalent(64.0, 24, 'rqc30')
    assert n == 30 and abs(value - 1.0) < 1e-12
    value, n = bench.workload_equivalent(5.0, 20, 'rc20')
    assert n == 20 and value == 5.0
