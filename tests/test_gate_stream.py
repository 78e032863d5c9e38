"""CPU: the gate-stream compiler (sdimb_gate_stream, csrc/planes_stream.cuh) keeps the program's meaning.

The compiled per-warp streams are decoded back into an op list the way gate_stream_kernel executes them (the merged
Pauli gates first; then per layer: the N1 events of the layer's noise-table range, then every warp's wide ops, lane
group by lane group) and that list goes through the C oracle next to the original program: records and final tableaus
must be identical."""
import numpy as np
import pytest

from make_cases import random_program
from oracle import c_oracle
from sdim_b200 import _native as N

MAGIC = 0x47533033
END, SYNC, H, P, CNOT, CZ, SWAP = range(7)
HEADER_ROWS, PAD_ROWS = 3, 96
INV, ON = 0x100, 0x200


def _front(prog):
    sched = N.schedule(prog.num_qudits, prog.ops)
    tail = N.tail_run(sched)
    return sched, tail


def _decode(gs, d):
    """Op rows (opcode, a, b, slot) in an order the kernel's execution is equivalent to."""
    magic, gpw, nw, n_tab = (int(v) for v in gs[0])
    tab_base, total, layers, n_pauli = (int(v) for v in gs[1])
    rs_bytes = int(gs[2][0])
    assert magic == MAGIC and total == gs.shape[0] and tab_base + n_tab + n_pauli == total
    table = gs[tab_base:]
    out = []
    for q, a, b, _ in gs[tab_base + n_tab:].tolist():          # the merged Pauli gates, in front of everything
        assert 0 <= a < d and 0 <= b < d and (a or b)
        out += [(1, q, -1, -1)] * a + [(3, q, -1, -1)] * b
    per_warp = []
    for w in range(nw):
        base, entries = int(gs[HEADER_ROWS + w][0]), int(gs[HEADER_ROWS + w][1])
        rows = gs[base: base + entries * gpw].reshape(entries, gpw, 4)
        assert all((int(r[0]) & 0xFF) == END for r in gs[base + (entries - 1) * gpw: base + (entries - 1) * gpw + PAD_ROWS])
        # split at SYNC entries
        segs, cur = [], None
        for e in rows:
            fam = int(e[0][0]) & 0xFF
            assert all((int(r[0]) & 0xFF) == fam for r in e)          # one family per wide op
            if fam == SYNC:
                assert all(tuple(r) == tuple(e[0]) for r in e)
                cur = {"noise": (int(e[0][1]), int(e[0][2])), "ops": []}
                segs.append(cur)
            elif fam == END:
                break
            else:
                cur["ops"].append(e)
        assert (int(rows[-1][0][0]) & 0xFF) == END                   # followed by 96 END rows of read-ahead padding
        per_warp.append(segs)
    assert all(len(s) == layers for s in per_warp)
    for layer in range(layers):
        lo, hi = per_warp[0][layer]["noise"]
        assert all(s[layer]["noise"] == (lo, hi) for s in per_warp)
        rows_written, rows_read = set(), set()
        for t in range(lo, hi):
            out.append((17, int(table[t][1]), -1, int(table[t][0])))
            rows_read.add(int(table[t][1]))
        for w in range(nw):
            for e in per_warp[w][layer]["ops"]:
                qs = []
                for r in e:
                    x, a, b = int(r[0]), int(r[1]), int(r[2])
                    if not x & ON:
                        continue
                    fam, inv = x & 0xFF, bool(x & INV)
                    assert a % rs_bytes == 0 and b % rs_bytes == 0
                    a, b = a // rs_bytes, b // rs_bytes
                    op = {H: 5, P: 7, CNOT: 9, CZ: 11, SWAP: 13}[fam] + (1 if inv else 0)
                    out.append((op, a, b if fam >= CNOT else -1, -1))
                    qs += [a] + ([b] if fam >= CNOT else [])
                assert len(qs) == len(set(qs))
                assert not (set(qs) & rows_written)
                rows_written |= set(qs)
        # (an N1 may share its layer with the next writer of its row: the kernel applies a layer's events first)
    return np.array(out, dtype=np.int32).reshape(-1, 4)


@pytest.mark.parametrize("d,n,depth", [(3, 24, 900), (2, 40, 1500), (3, 70, 1200), (2, 130, 800), (3, 5, 200), (3, 256, 600)])
def test_gate_stream_is_equivalent_to_the_program(d, n, depth):
    prog = random_program(seed=17 * d + n, n=n, d=d, depth=depth, p_meas=0.0)
    sched, tail = _front(prog)
    assert tail == n
    front = sched[: sched.shape[0] - tail]
    gs = N.gate_stream(n, d, front)
    assert gs is not None
    ops = _decode(gs, d)
    src = [tuple(r) for r in prog.ops.tolist() if r[0] not in (0, 1, 2, 3, 4, 14)]
    # the same multiset of non-Pauli gates and N1 events (b and slot fields normalised); the Pauli gates were merged
    norm = lambda r: (r[0], r[1], r[2] if 9 <= r[0] <= 13 else -1, r[3] if r[0] == 17 else -1)
    assert sorted(map(norm, (r for r in ops.tolist() if r[0] > 4))) == sorted(map(norm, src))
    assert sum(1 for r in ops.tolist() if r[0] <= 4) <= 2 * (d - 1) * n
    if not c_oracle.available():
        pytest.skip("liboracle.so not built")
    meas = np.array([[14, q, -1, q] for q in range(n)], dtype=np.int32)
    kw = dict(thresh24=prog.noise_thresh24, channel=prog.noise_channel, want_final=True)
    shots = 6
    r0, f0 = c_oracle.run(n, d, prog.ops, shots, 0, 5, **kw)
    r1, f1 = c_oracle.run(n, d, np.concatenate([ops, meas]), shots, 0, 5, **kw)
    assert np.array_equal(r0, r1)
    assert all(np.array_equal(f0[k], f1[k]) for k in f0)


def test_gate_stream_rejects_what_it_cannot_compile():
    n = 8
    front = N.schedule(n, np.array([[5, 0, -1, -1], [9, 0, 1, -1]], dtype=np.int32))
    assert N.gate_stream(n, 3, front) is not None
    assert N.gate_stream(n, 5, front) is None                                # bit planes only
    assert N.gate_stream(600, 3, front) is None                              # rows wider than a warp
    with_m = N.schedule(n, np.array([[5, 0, -1, -1], [14, 0, -1, 0]], dtype=np.int32))
    assert N.gate_stream(n, 3, with_m) is None                               # a measurement in the stretch
    unlayered = np.array([[5, 0, -1, -1], [9, 0, 1, -1]], dtype=np.int32)    # the compiler layers by itself
    gs = N.gate_stream(n, 3, unlayered)
    assert gs is not None and int(gs[1][2]) == 2 and _decode(gs, 3).tolist() == [[5, 0, -1, -1], [9, 0, 1, -1]]
    empty = N.gate_stream(n, 3, np.zeros((0, 4), dtype=np.int32))
    assert empty is not None and int(empty[1][2]) == 0


def test_gate_stream_gives_up_above_its_noise_table_limit():
    """More than 65 536 N1 ops in the stretch: the fired bits would not fit the kernel's shared-memory table; the
    compiler returns SDIMB_EINVAL and the caller runs the interpreter (TableauEngine.gate_stream is None then)."""
    n, k = 8, 65537
    ops = np.zeros((k + 1, 4), dtype=np.int32)
    ops[:k, 0], ops[:k, 1], ops[:k, 2], ops[:k, 3] = 17, np.arange(k) % n, -1, np.arange(k)
    ops[k] = (5, 0, -1, -1)
    assert N.gate_stream(n, 3, ops) is None
    assert N.gate_stream(n, 3, ops[1:]) is not None            # 65 536 events: still compiled
