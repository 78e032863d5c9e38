"""CPU: the host-side op-stream scheduler (sdimb_schedule) keeps the program's meaning.

Layers must only group ops that commute exactly on the tableau; the proof here is empirical — the reordered
stream through the C oracle gives identical records AND final tableaus — plus the structural property itself."""
import numpy as np
import pytest

from make_cases import random_program
from oracle import c_oracle
from sdim_b200 import _native as N

WRITERS = {5, 6, 7, 8, 9, 10, 11, 12, 13}
READERS = {1, 2, 3, 4, 17}


def _layers(sched):
    cur = []
    for row in sched:
        op = int(row[0]) & 0xFF
        if op == N.OP_BARRIER or op in (14, 15, 16):
            if cur:
                yield cur
            cur = []
        else:
            cur.append(row)
    if cur:
        yield cur


@pytest.mark.parametrize("d,n,depth", [(3, 24, 900), (2, 40, 1500), (5, 9, 400), (3, 3, 200)])
def test_schedule_structure_and_equivalence(d, n, depth):
    prog = random_program(seed=31 * d + n, n=n, d=d, depth=depth)
    sched = N.schedule(n, prog.ops)
    ops_only = sched[(sched[:, 0] & 0xFF) != N.OP_BARRIER]
    # same multiset of ops, event slots travel with their ops, collectives keep their order
    a = sorted(map(tuple, np.column_stack((ops_only[:, 0] & 0xFF, ops_only[:, 1:])).tolist()))
    b = sorted(map(tuple, prog.ops.tolist()))
    assert a == b
    coll = [(r[0] & 0xFF,) + tuple(r[1:]) for r in ops_only.tolist() if (r[0] & 0xFF) in (14, 15, 16)]
    assert coll == [tuple(r) for r in prog.ops.tolist() if r[0] in (14, 15, 16)]
    # marks of measurement runs (SDIMB_GM_*): only on M ops, FIRST ... LAST bracket at least min_run marked ops with
    # nothing but marked M ops between them, FOLLOW iff the stream goes on behind the run
    min_run = max(n // 8, 4)
    marks = [((int(r[0]) >> 8) & 0xFF, int(r[0]) & 0xFF) for r in ops_only.tolist()]
    inside, length = False, 0
    for k, (mk, op) in enumerate(marks):
        if op != 14 or not (mk & 1):
            assert not inside and (op != 14 or mk == 0)
            continue
        if mk & 2:
            assert not inside
            inside, length = True, 0
        assert inside
        length += 1
        if mk & 4:
            assert length >= min_run and bool(mk & 8) == (k + 1 < len(marks))
            inside = False
    assert not inside
    # inside a layer: a written row is touched by exactly one op; warps are balanced
    for layer in _layers(sched):
        written, touched = set(), {}
        for row in layer:
            op, qa, qb = int(row[0]) & 0xFF, int(row[1]), int(row[2])
            rows = [qa] + ([qb] if op in (9, 10, 11, 12, 13) else [])
            for q in rows:
                touched[q] = touched.get(q, 0) + 1
                if op in WRITERS:
                    written.add(q)
        assert all(touched[q] == 1 for q in written)
        warps = np.bincount([(int(r[0]) >> 8) & 0xFF for r in layer], minlength=4)
        assert warps.max() - warps.min() <= 1
    if not c_oracle.available():
        pytest.skip("liboracle.so not built")
    plain = ops_only.copy()
    plain[:, 0] &= 0xFF
    kw = dict(thresh24=prog.noise_thresh24, channel=prog.noise_channel, want_final=True)
    r0, f0 = c_oracle.run(n, d, prog.ops, 12, 0, 9, **kw)
    r1, f1 = c_oracle.run(n, d, plain, 12, 0, 9, **kw)
    assert np.array_equal(r0, r1)
    assert all(np.array_equal(f0[k], f1[k]) for k in f0)


def test_schedule_rejects_bad_streams_and_handles_empty():
    assert N.schedule(4, np.zeros((0, 4), dtype=np.int32)).shape == (0, 4)
    with pytest.raises(ValueError):
        N.schedule(4, np.array([[9, 0, 0, -1]], dtype=np.int32))
    with pytest.raises(ValueError):
        N.schedule(4, np.array([[5, 7, -1, -1]], dtype=np.int32))
    out = N.schedule(3, np.array([[0, 1, -1, -1], [14, 0, -1, 0]], dtype=np.int32))     # I dropped, M kept
    assert out.tolist() == [[14, 0, -1, 0]]
