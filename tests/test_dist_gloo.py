"""CPU: the N > 1 path (shot sharding + record gather) on 2 gloo ranks.  The GPU engine is replaced by an
injected runner that derives each shot's records from its GLOBAL shot id, exactly like the Philox counters."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sdim_b200.dist import shard_range


def test_shard_range_partitions_exactly():
    for shots in (0, 1, 7, 8, 1000, 10 ** 7 + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_range(shots, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == shots
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shots, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from make_cases import random_circuit
    from sdim_b200 import Program
    from sdim_b200.dist import simulate_sharded
    from sdim_b200.ir import compile_circuits
    from sdim_b200.rng import measurement_draws
    circ = random_circuit(5, 6, 3, 40)
    compiled = compile_circuits([circ])

    def runner(lo, hi):
        return torch.from_numpy(measurement_draws(11, 3, np.arange(lo, hi), compiled.n_meas))

    rec = simulate_sharded(Program(circ), compiled, shots, 11, runner=runner)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), rec)
    dist.destroy_process_group()


@pytest.mark.parametrize("shots", [10, 7])
def test_two_rank_gather_equals_single_process(tmp_path, shots):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, shots, str(tmp_path)), nprocs=world, join=True)
    from make_cases import random_program
    from sdim_b200.rng import measurement_draws
    n_meas = random_program(5, 6, 3, 40).n_meas
    want = measurement_draws(11, 3, np.arange(shots), n_meas)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npy"))
        assert got.shape == (shots, n_meas) and np.array_equal(got, want)


def _worker_program(rank, world, port, out_dir):
    """The product's own distributed path (Program.simulate_records(distributed=True)) on the stand-in engine: initial
    tableau, replayed draws, shot_offset and an unseeded call all have to behave as in a single process."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import random
    import sdim_b200.engine as engine_mod
    from fake_engine import FakeEngine
    from make_cases import random_circuit
    from sdim_b200 import Program
    engine_mod.TableauEngine = FakeEngine
    torch.cuda.mem_get_info = lambda device=None: (1 << 34, 1 << 34)
    shots = 7
    circ = random_circuit(5, 6, 3, 60)
    # (a) plain: global shot ids, explicit seed and shot_offset
    a = Program(circ).simulate_records(shots, seed=11, shot_offset=100, distributed=True)
    # (b) unseeded: ranks seed `random` differently, rank 0's draw must win everywhere
    random.seed(1000 + rank)
    prog_b = Program(circ)
    b = prog_b.simulate_records(shots, distributed=True)
    # (c) replayed measurement outcomes + an initial tableau taken from a previous run
    rm = (np.arange(shots * a.values.shape[1], dtype=np.uint8).reshape(shots, -1) % 3)
    start = Program(circ)
    start.simulate_records(1, seed=3)
    c = Program(circ, tableau=start.stabilizer_tableau).simulate_records(shots, seed=11, replay_meas=rm, distributed=True)
    np.savez(os.path.join(out_dir, f"prog_rank{rank}.npz"), a=a.values, a_det=a.deterministic, b=b.values,
             b_seed=np.uint64(b.seed), c=c.values, a_off=a.shot_offset)
    dist.destroy_process_group()


def test_two_rank_program_path_equals_single_process(tmp_path, monkeypatch):
    world, port = 2, _free_port()
    mp.spawn(_worker_program, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    import random
    import sdim_b200.engine as engine_mod
    from fake_engine import FakeEngine
    from make_cases import random_circuit
    from sdim_b200 import Program
    monkeypatch.setattr(engine_mod, "TableauEngine", FakeEngine)
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda device=None: (1 << 34, 1 << 34))
    shots = 7
    circ = random_circuit(5, 6, 3, 60)
    got = [np.load(os.path.join(str(tmp_path), f"prog_rank{r}.npz")) for r in range(world)]
    a = Program(circ).simulate_records(shots, seed=11, shot_offset=100)
    for g in got:
        assert np.array_equal(g["a"], a.values) and np.array_equal(g["a_det"], a.deterministic) and int(g["a_off"]) == 100
    random.seed(1000)                                                      # rank 0's stream
    b = Program(circ).simulate_records(shots)
    assert int(got[0]["b_seed"]) == int(got[1]["b_seed"]) == b.seed
    for g in got:
        assert np.array_equal(g["b"], b.values)
    rm = (np.arange(shots * a.values.shape[1], dtype=np.uint8).reshape(shots, -1) % 3)
    start = Program(circ)
    start.simulate_records(1, seed=3)
    c = Program(circ, tableau=start.stabilizer_tableau).simulate_records(shots, seed=11, replay_meas=rm)
    for g in got:
        assert np.array_equal(g["c"], c.values)
    assert not np.array_equal(c.values, Program(circ).simulate_records(shots, seed=11, replay_meas=rm).values)
