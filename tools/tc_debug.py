"""Diagnostic for the tensor-core 5-qubit kernel: applies marker matrices to
marker states and prints observed vs expected, so a layout/descriptor mistake
can be read off from one run."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cirq_b200.device_state import DeviceState  # noqa: E402
from cirq_b200 import _lib  # noqa: E402

lib = _lib.load()
n = 12
targets = [11, 10, 9, 8, 7]  # all high: group index = low 7 bits
np.set_printoptions(linewidth=250, precision=4, suppress=True)


def run(matrix, state, tc):
    lib.b2q_set_tc_mode(tc)
    dev = DeviceState.from_numpy(state.astype(np.complex64))
    dev.apply_matrix(matrix, targets)
    return dev.to_numpy()


rng = np.random.RandomState(0)
# 1. identity
state = (rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)).astype(np.complex64)
out = run(np.eye(32), state, 1)
print('identity max err', np.max(np.abs(out - state)))
# 2. marker matrix, marker input: group 0..127 all get x[c0] = 1
M = np.array([[(r + 1) + 0.01j * (c + 1) for c in range(32)] for r in range(32)])
for c0 in (0, 1, 2, 5, 31):
    st = np.zeros(1 << n, dtype=np.complex64)
    # matrix index bit for targets[q] is (4-q): column c0 <-> bits of targets
    idx = 0
    for q, t in enumerate(targets):
        if (c0 >> (4 - q)) & 1:
            idx |= 1 << t
    st[idx + np.arange(128)] = 1.0
    got = run(M, st, 1)
    ref = run(M, st, 0)
    # gather output column of group 3
    rows = []
    for r in range(32):
        i = 3
        for q, t in enumerate(targets):
            if (r >> (4 - q)) & 1:
                i |= 1 << t
        rows.append(i)
    print(f'c0={c0}: max err vs register kernel {np.max(np.abs(got - ref)):.3e}')
    if np.max(np.abs(got - ref)) > 1e-3:
        print(' got', got[rows][:8])
        print(' ref', ref[rows][:8])
# 3. random complex
M = (rng.standard_normal((32, 32)) + 1j * rng.standard_normal((32, 32))) / 6
for tg in ([11, 10, 9, 8, 7], [7, 8, 9, 10, 11], [0, 1, 2, 3, 4], [5, 0, 11, 3, 8]):
    targets = tg
    got = run(M, state, 1)
    ref = run(M, state, 0)
    print('random', tg, 'max err', np.max(np.abs(got - ref)), 'rel', np.max(np.abs(got - ref)) / np.max(np.abs(ref)))

# 4. norm drift over many unitary passes (bias check)
def rand_unitary(k):
    d = 1 << k
    q, r = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
    return q * (np.diag(r) / np.abs(np.diag(r)))

n = 20
for tc in (1, 0):
    lib.b2q_set_tc_mode(tc)
    dev = DeviceState.basis(n, np.complex64, 0)
    r2 = np.random.RandomState(5)
    norms = []
    for i in range(60):
        tg = r2.permutation(n)[:5].tolist()
        d = 32
        q, r = np.linalg.qr(r2.standard_normal((d, d)) + 1j * r2.standard_normal((d, d)))
        dev.apply_matrix(q * (np.diag(r) / np.abs(np.diag(r))), tg)
        if i % 10 == 9:
            norms.append(dev.norm2())
    print('tc' if tc else 'fp32', 'norm after 10..60 passes', ['%.7f' % v for v in norms])
