"""Batched Monte-Carlo trajectories for noisy / mid-circuit-measured runs.

``cirq.Simulator.run`` falls back to one full simulation per repetition as
soon as the circuit has noise or a non-terminal measurement
(cirq-core/cirq/sim/simulator_base.py:249-264): every repetition copies the
state and walks the circuit again, choosing one operator per stochastic
operation (sim/state_vector_simulation_state.py:183-257) — thousands of tiny
numpy calls (or, on a GPU, tiny launches).  Here B = 2^b repetitions advance
together as ONE (n+b)-bit device array, trajectory t in index bits [n, n+b):

* unitary gates are queued, fused and applied by the ordinary gate kernels —
  one launch serves all trajectories;
* a mixture draws all B choices on the host in one ``prng.choice`` call and
  applies them with ``b2q_bsv_apply_select`` (trajectories that drew the
  identity are skipped, not streamed);
* a Kraus channel gets all trial weights ||K_i psi_t||^2 from one read-only pass
  (``b2q_bsv_kraus_weights``), picks per trajectory exactly as the reference's
  loop does, and applies K_i / sqrt(weight) with the same select kernel;
* a measurement takes the joint marginal over (trajectory, measured bits), picks
  one outcome per trajectory and collapses with ``b2q_bsv_collapse``; terminal
  measurements draw one sample per trajectory with the full-state sampler.

The distribution of results is the reference's; the ORDER in which random
numbers are consumed is not (per operation across trajectories instead of per
repetition), so seeded bit-for-bit equality with ``cirq.Simulator`` is given up
— which is why batching is opt-in (``B200Simulator(trajectory_batch=...)``).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from cirq_b200._cirq_compat import import_cirq
from cirq_b200.fusion import fuser_for

cirq = import_cirq()
from cirq import ops, protocols  # noqa: E402

# largest (n + b)-bit array a batch may occupy (2^30 complex64 = 8.6 GB)
MAX_BATCH_STATE_BITS = 30
# the joint (trajectory, outcome) table of a measurement is reduced on the device
# by b2q_sv_marginal_probs, which takes at most this many bits
_MAX_MARGINAL_BITS = 24


def _classify(op, axis_bit, cache: dict | None = None) -> tuple | None:
    """One suffix operation -> ('measure'|'unitary'|'mixture'|'kraus', ...), or
    None if it cannot be batched (classical control, confusion maps, keyed
    channels, qudits, operators wider than 3 qubits).  `cache` maps a gate to
    its qubit-independent classification (a noise model repeats one channel
    hundreds of times)."""
    from cirq_b200.sv_simulator import cached_unitary

    if any(d != 2 for d in protocols.qid_shape(op)):
        return None
    bits = [axis_bit[q] for q in op.qubits]
    if isinstance(op.gate, ops.MeasurementGate):
        if op.gate.confusion_map:
            return None
        return ('measure', bits, str(protocols.measurement_key_obj(op)), list(op.gate.full_invert_mask()))
    key = None
    # (noise models tag what they insert: look through the tags)
    if cache is not None and isinstance(op.untagged, ops.GateOperation):
        try:
            hit = cache.get(op.gate)
            key = op.gate
        except TypeError:  # unhashable gate
            hit = None
        if hit is not None:
            return hit[:-1] + (bits,) if hit[0] != 'mixture' else hit[:3] + (bits, hit[4])
    form = _classify_uncached(op, bits, cached_unitary)
    if key is not None and form is not None and not protocols.is_parameterized(key):
        cache[key] = form
    return form


def _classify_uncached(op, bits, cached_unitary) -> tuple | None:
    if protocols.is_measurement(op) or protocols.control_keys(op):
        return None
    u = cached_unitary(op)
    if u is not None:
        return ('unitary', u, bits)
    if len(bits) > 3:
        return None
    mixture = protocols.mixture(op, default=None)
    if mixture is not None:
        probs, unitaries = zip(*mixture)
        mats = np.stack([np.asarray(u, dtype=np.complex128) for u in unitaries])
        eye = np.eye(mats.shape[1])
        skip = next((i for i, m in enumerate(mats) if np.array_equal(m, eye)), -1)
        return ('mixture', np.asarray(probs, dtype=np.float64), mats, bits, skip)
    kraus = protocols.kraus(op, default=None)
    if kraus is not None:
        return ('kraus', np.stack([np.asarray(k, dtype=np.complex128) for k in kraus]), bits)
    return None


def plan_suffix(noisy_moments, qubits) -> list | None:
    """Classified operation list of the per-repetition part of a run, or None if
    any operation rules batching out."""
    n = len(qubits)
    axis_bit = {q: n - 1 - i for i, q in enumerate(qubits)}
    plan = []
    cache: dict = {}
    for moment in noisy_moments:
        for op in ops.flatten_to_ops(moment):
            item = _classify(op, axis_bit, cache)
            if item is None:
                return None
            plan.append(item)
    return plan


class TrajectoryBatch:
    """2^batch_bits copies of an n-qubit state, advanced together."""

    def __init__(self, psi0, n_qubits: int, batch_bits: int, dtype, prng, max_fused_qubits=None):
        self.n = int(n_qubits)
        self.b = int(batch_bits)
        self.count = 1 << self.b
        self.dtype = np.dtype(dtype)
        self.prng = prng
        if self.b == 0:
            self.dev = psi0.copy()
        else:
            ones = type(psi0).from_numpy(np.ones(self.count, dtype=self.dtype))
            self.dev = ones.kron(psi0)
        self.fuser = fuser_for(self.dtype, max_fused_qubits, self.n + self.b, state_vector=True)
        self.passes = 0

    # -- unitary part ---------------------------------------------------------------------

    def queue_unitary(self, u: np.ndarray, bits: Sequence[int]) -> None:
        self.fuser.add(u, list(bits))

    def flush(self) -> None:
        if self.fuser.pending:  # (a trailing relabelled SWAP counts: blocks() puts it back)
            blocks = self.fuser.blocks()
            self.fuser.clear()
            self.dev.apply_batch(blocks)
            self.passes += len(blocks)

    # -- stochastic operations -------------------------------------------------------------

    def mixture(self, probs: np.ndarray, unitaries: np.ndarray, bits: Sequence[int], skip: int) -> None:
        """sim/state_vector_simulation_state.py:183-203, all trajectories at once."""
        self.flush()
        choice = self._draw(probs, (self.count,), skip)
        if skip >= 0 and np.all(choice == skip):
            return
        self.dev.bsv_apply_select(self.n, unitaries, bits, choice, None, skip)
        self.passes += 1

    def _draw(self, probs: np.ndarray, shape: tuple, skip: int) -> np.ndarray:
        """Independent draws from `probs` of the given shape.  Weak noise (the
        identity, `skip`, takes most of the weight) is drawn sparsely: how many
        draws are NOT the identity (binomial), where they are (a uniform subset),
        and which operator each one is — the same joint distribution as one
        categorical draw per entry, at a cost proportional to the number of hits."""
        total = int(np.prod(shape))
        if skip < 0 or probs[skip] < 0.75 or total < 1024:
            return self.prng.choice(len(probs), size=shape, p=probs)
        rest = np.delete(np.arange(len(probs)), skip)
        p_rest = probs[rest]
        weight = float(p_rest.sum())
        out = np.full(total, skip, dtype=np.int64)
        hits = int(self.prng.binomial(total, weight)) if weight > 0 else 0
        if hits:
            # a uniform `hits`-subset: the first `hits` distinct values of a stream of
            # uniform integers (choice(..., replace=False) would permute all `total`)
            where = np.zeros(0, dtype=np.int64)
            while where.size < hits:
                draws = np.concatenate([where, self.prng.randint(0, total, size=hits - where.size + 16)])
                _, first = np.unique(draws, return_index=True)
                where = draws[np.sort(first)]
            where = where[:hits]
            out[where] = rest[self.prng.choice(len(rest), size=hits, p=p_rest / weight)]
        return out.reshape(shape)

    def mixture_layer(self, probs: np.ndarray, unitaries: np.ndarray, bits: Sequence[int], skip: int) -> None:
        """The same 1-qubit mixture on each of `bits` (a noise model's layer after a
        moment): all draws in one ``prng.choice`` call, all applications in one
        launch (``b2q_bsv_apply_select_multi``)."""
        self.flush()
        choices = self._draw(probs, (len(bits), self.count), skip)
        if skip >= 0:
            hit = np.flatnonzero((choices != skip).any(axis=1))
            if hit.size == 0:
                return
            choices = choices[hit]
            bits = [bits[i] for i in hit]
        self.dev.bsv_apply_select_multi(self.n, unitaries, list(bits), choices, skip)
        self.passes += 1

    def channel(self, kraus: np.ndarray, bits: Sequence[int]) -> None:
        """sim/state_vector_simulation_state.py:205-257: the first operator whose
        cumulative weight exceeds the uniform draw; the most likely one when
        rounding leaves none (or picks a zero weight)."""
        self.flush()
        w = self.dev.bsv_kraus_weights(self.n, kraus, bits)
        u = self.prng.random_sample(self.count)
        cum = np.cumsum(w, axis=1)
        chosen = (u[:, None] >= cum).sum(axis=1)
        rows = np.arange(self.count)
        bad = chosen >= w.shape[1]
        chosen = np.where(bad, 0, chosen)
        bad |= w[rows, chosen] <= 0
        chosen = np.where(bad, np.argmax(w, axis=1), chosen)
        weight = w[rows, chosen]
        self.dev.bsv_apply_select(self.n, kraus, bits, chosen, 1.0 / np.sqrt(weight), -1)
        self.passes += 2

    def _batch_bits_desc(self) -> list[int]:
        return list(range(self.n + self.b - 1, self.n - 1, -1))

    def measure(self, bits: Sequence[int]) -> np.ndarray:
        """Mid-circuit measurement: uint8[trajectories, len(bits)], state collapsed
        and renormalised per trajectory (sim/state_vector.py:235-322)."""
        self.flush()
        bits = list(bits)
        out = np.zeros((self.count, len(bits)), dtype=np.uint8)
        group = max(1, _MAX_MARGINAL_BITS - self.b)
        for g0 in range(0, len(bits), group):
            gb = bits[g0:g0 + group]
            joint = np.asarray(self.dev.marginal_probs(self._batch_bits_desc() + gb), dtype=np.float64)
            joint = joint.reshape(self.count, 1 << len(gb))
            total = joint.sum(axis=1, keepdims=True)
            cum = np.cumsum(joint / total, axis=1)
            u = self.prng.random_sample(self.count)
            pick = (u[:, None] >= cum).sum(axis=1)
            # never land on an outcome of probability zero (rounding at the top end)
            last = (joint.shape[1] - 1) - np.argmax((joint > 0)[:, ::-1], axis=1)
            pick = np.minimum(pick, last)
            p_sel = joint[np.arange(self.count), pick]
            vals = ((pick[:, None] >> np.arange(len(gb) - 1, -1, -1)) & 1).astype(np.uint8)
            self.dev.bsv_collapse(self.n, gb, vals, 1.0 / np.sqrt(p_sel))
            out[:, g0:g0 + len(gb)] = vals
            self.passes += 2
        return out

    def sample_terminal(self, bits: Sequence[int]) -> np.ndarray:
        """One draw per trajectory from its full distribution (no collapse): the
        uniform of trajectory t is mapped into t's interval of the cumulative
        distribution of the whole array, so one run of the full-state sampler
        serves every trajectory."""
        self.flush()
        if self.b:
            w = np.asarray(self.dev.marginal_probs(self._batch_bits_desc()), dtype=np.float64)
        else:
            w = np.array([self.dev.norm2()], dtype=np.float64)
        u = np.clip(self.prng.random_sample(self.count), 2.0 ** -40, 1.0 - 2.0 ** -40)
        cum = np.cumsum(w)
        mapped = (cum - w + u * w) / cum[-1]
        idx = np.asarray(self.dev.sample_indices_device(mapped).cpu().numpy(), dtype=np.int64).reshape(-1)
        idx = idx[: self.count]
        lo = np.arange(self.count, dtype=np.int64) << self.n
        idx = np.clip(idx, lo, lo + (1 << self.n) - 1) - lo
        bits = np.asarray(list(bits), dtype=np.int64)
        self.passes += 1
        return ((idx[:, None] >> bits[None, :]) & 1).astype(np.uint8)


def choose_batch_bits(n_qubits: int, repetitions: int, max_trajectories: int) -> int:
    """log2 of the trajectories per batch: enough for the repetitions, within the
    caller's cap and the memory budget of one batch."""
    want = max(0, int(repetitions - 1).bit_length())
    cap = max(0, int(max_trajectories).bit_length() - 1)
    room = max(0, MAX_BATCH_STATE_BITS - n_qubits)
    return min(want, cap, room, _MAX_MARGINAL_BITS - 1)


def run_plan(plan: list, psi0, n_qubits: int, repetitions: int, dtype, prng, max_trajectories: int,
             max_fused_qubits=None, info: dict | None = None) -> dict[str, np.ndarray]:
    """Executes a `plan_suffix` plan for `repetitions` trajectories starting from
    the device state psi0; returns {key: uint8[repetitions, instances, qubits]} like
    SimulatorBase._run (sim/simulator_base.py:266-275)."""
    records: dict[str, list[list[np.ndarray]]] = {}
    done = 0
    batches = passes = 0
    # nothing observes the state after the last measurement of a `run`: operations
    # behind it (a noise model's layer after the measurement moment) are dropped
    plan = list(plan)
    while plan and plan[-1][0] != 'measure':
        plan.pop()
    # measurements after which nothing else touches the state can be sampled, not collapsed
    last_non_measure = max((i for i, it in enumerate(plan) if it[0] != 'measure'), default=-1)
    seen_bits: set[int] = set()
    terminal_ok = True
    for it in plan[last_non_measure + 1:]:
        if seen_bits & set(it[1]):
            terminal_ok = False
        seen_bits |= set(it[1])
    while done < repetitions:
        b = choose_batch_bits(n_qubits, repetitions - done, max_trajectories)
        tb = TrajectoryBatch(psi0, n_qubits, b, dtype, prng, max_fused_qubits)
        take = min(tb.count, repetitions - done)
        chunk: dict[str, list[np.ndarray]] = {}
        i = 0
        while i < len(plan):
            item = plan[i]
            kind = item[0]
            if kind == 'unitary':
                tb.queue_unitary(item[1], item[2])
            elif kind == 'mixture':
                # a run of the same 1-qubit mixture (same operator table) on
                # different qubits travels as one layer
                j = i + 1
                if len(item[3]) == 1:
                    while (j < len(plan) and plan[j][0] == 'mixture' and plan[j][2] is item[2]
                           and plan[j][1] is item[1] and j - i < 32):
                        j += 1
                if j - i > 1:
                    tb.mixture_layer(item[1], item[2], [plan[x][3][0] for x in range(i, j)], item[4])
                    i = j
                    continue
                tb.mixture(item[1], item[2], item[3], item[4])
            elif kind == 'kraus':
                tb.channel(item[1], item[2])
            elif i > last_non_measure and terminal_ok:
                tail = plan[i:]
                all_bits = [bit for it in tail for bit in it[1]]
                cols = tb.sample_terminal(all_bits)
                start = 0
                for it in tail:
                    m = len(it[1])
                    vals = cols[:, start:start + m] ^ np.asarray(it[3], dtype=np.uint8)[None, :]
                    chunk.setdefault(it[2], []).append(vals[:take])
                    start += m
                break
            else:
                vals = tb.measure(item[1]) ^ np.asarray(item[3], dtype=np.uint8)[None, :]
                chunk.setdefault(item[2], []).append(vals[:take])
            i += 1
        for key, instances in chunk.items():
            records.setdefault(key, []).append(instances)
        done += take
        batches += 1
        passes += tb.passes
        del tb
    if info is not None:
        info.update(batches=batches, passes=passes, batch_bits=choose_batch_bits(
            n_qubits, repetitions, max_trajectories))
    out = {}
    for key, chunks in records.items():
        # chunks[c][instance] = uint8[take_c, m]  ->  [reps, instances, m]
        per_chunk = [np.stack(instances, axis=1) for instances in chunks]
        out[key] = np.concatenate(per_chunk, axis=0)
    return out
