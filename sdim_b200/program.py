"""`Program`: the reference's driver API on top of the CUDA tableau engine.

Mirror of `sdim.program.Program` (reference: sdim/program.py:177-574) for
prime dimensions.  `simulate()` keeps the reference signature and return
shapes — a flat list ordered by (qudit, round) for one shot
(program.py:357-363), the nested list `[qudit][round][shot]` otherwise
(program.py:364-365) — but every shot is a full stabilizer tableau simulated
on the GPU (`force_tableau` semantics for any shot count), with N1 noise applied
per shot with the distribution of `_build_ir` (program.py:486-507) instead of
being ignored (program.py:31).

Extra keyword-only arguments (not in the reference): `seed`, `replay_meas`,
`replay_noise`, `mode`, plus `simulate_records()` which returns the packed
record table without materialising Python objects.
"""
from __future__ import annotations

import copy
import random
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import records as R
from .circuit import Circuit, CircuitInstruction
from .gatedata import MEASURE_OPS, OP_RESET
from .ir import MAX_DIMENSION, CompiledProgram, compile_circuits, is_prime
from .results import MEASUREMENT_DTYPE, MeasurementResult
from .tableau import ExtendedTableau


def undo_reset_correction(arrays: dict, qudit: int, outcome: int, d: int) -> dict:
    """Exported tableau arrays (keys x, z, p, dx, dz, dp; [qudit, generator]) AFTER a RESET op -> the tableau right
    after its measurement, i.e. before the driver's X^((-m) mod d) correction (sdim/program.py:335-339).  X^k only
    moves phases: phase -= phase_order * k * z[qudit] on both halves, so the inverse adds it."""
    k = (-outcome) % d
    po, order = (2, 2 * d) if d % 2 == 0 else (1, d)
    out = dict(arrays)
    out["p"] = (np.asarray(arrays["p"]) + po * k * np.asarray(arrays["z"])[qudit]) % order
    out["dp"] = (np.asarray(arrays["dp"]) + po * k * np.asarray(arrays["dz"])[qudit]) % order
    return out


@dataclass
class SimulationOptions:
    """Same fields as the reference (sdim/program.py:167-175)."""

    shots: int = 1
    show_measurement: bool = False
    record_tableau: bool = False
    force_tableau: bool = False
    verbose: bool = False
    show_gate: bool = False
    exact: bool = False


@dataclass
class RecordTable:
    """Packed result of a run: one row per shot, one column per chronological measurement."""

    values: np.ndarray          # uint8 [shots, n_meas] (uint16 for d > 127)
    deterministic: np.ndarray   # bool  [shots, n_meas]
    meas_qudit: np.ndarray      # int32 [n_meas]
    meas_round: np.ndarray      # int32 [n_meas]
    seed: int
    shot_offset: int = 0

    @property
    def shots(self) -> int:
        return int(self.values.shape[0])

    def column(self, qudit: int, meas_round: int = 0) -> int:
        hit = np.nonzero((self.meas_qudit == qudit) & (self.meas_round == meas_round))[0]
        if hit.size == 0:
            raise ValueError(f"qudit {qudit} has no measurement round {meas_round}")
        return int(hit[0])

    def to_structured(self, num_qudits: Optional[int] = None) -> np.ndarray:
        """The reference's record array: structured `[n_qudits, rounds, shots]` of MEASUREMENT_DTYPE
        (sdim/program.py:34-40, the layout `simulate_frame` returns, :57-61).  Cells a qudit never reached (fewer
        rounds than the maximum) stay zero, as in the reference's `np.zeros` allocation."""
        n = int(num_qudits if num_qudits is not None else (self.meas_qudit.max() + 1 if self.meas_qudit.size else 0))
        rounds = int(self.meas_round.max() + 1) if self.meas_round.size else 0
        out = np.zeros((n, rounds, self.shots), dtype=MEASUREMENT_DTYPE)
        shot_ids = np.arange(self.shot_offset, self.shot_offset + self.shots, dtype=np.int64)
        for k in range(self.meas_qudit.size):
            cell = out[int(self.meas_qudit[k]), int(self.meas_round[k])]
            cell["qudit_index"] = int(self.meas_qudit[k])
            cell["meas_round"] = int(self.meas_round[k])
            cell["shot"] = shot_ids
            cell["deterministic"] = self.deterministic[:, k]
            cell["measurement_value"] = self.values[:, k]
        return out

    def save(self, path: str) -> str:
        """Write the table as one compressed `.npz` (the reference's TODO.md:12 asks for numpy-array results):
        values uint8 [shots, n_meas], the deterministic flags bit-packed along the measurement axis, the
        (qudit, round) of every column, seed and shot offset.  `RecordTable.load` reads it back."""
        if not path.endswith(".npz"):
            path += ".npz"
        np.savez_compressed(path, values=self.values, deterministic_bits=np.packbits(self.deterministic, axis=1),
                            meas_qudit=self.meas_qudit, meas_round=self.meas_round,
                            seed=np.uint64(self.seed), shot_offset=np.int64(self.shot_offset))
        return path

    @classmethod
    def load(cls, path: str) -> "RecordTable":
        with np.load(path) as z:
            values = z["values"]
            det = np.unpackbits(z["deterministic_bits"], axis=1, count=values.shape[1]).astype(bool)
            return cls(values=values, deterministic=det, meas_qudit=z["meas_qudit"], meas_round=z["meas_round"],
                       seed=int(z["seed"]), shot_offset=int(z["shot_offset"]))


class Program:
    def __init__(self, circuit: Circuit, tableau: Optional[ExtendedTableau] = None, device=None,
                 fold_gates: bool = False):
        d = circuit.dimension
        if not is_prime(d):
            raise ValueError(f"dimension {d} is not prime: only the prime-dimension (ExtendedTableau) path of the "
                             "reference is implemented; composite dimensions (WeylTableau) are out of scope")
        if d > MAX_DIMENSION:
            raise ValueError(f"dimension {d} exceeds the uint16-lane limit of {MAX_DIMENSION}")
        self.circuits: List[Circuit] = [circuit]
        self.measurement_results: list = []
        self.device = device
        self.initial_tableau: Optional[ExtendedTableau] = tableau
        self._tableau_cache: Optional[ExtendedTableau] = (
            copy.deepcopy(tableau) if tableau is not None else ExtendedTableau(circuit.num_qudits, d))
        self._tableau_thunk = None
        self._engine = None
        self._engine_key = None
        self.last_records: Optional[RecordTable] = None
        # Not in the reference: fold runs of self-cancelling gates before upload (sdim_b200/peephole.py).  Records,
        # final tableaus and the shot*gates count are unchanged; host-stepped modes (verbose / show_gate /
        # record_tableau) always run the stream as written.
        self.fold_gates = fold_gates

    # ---- engine plumbing -------------------------------------------------------------------------
    def _compiled(self, fold: bool = False) -> CompiledProgram:
        compiled = compile_circuits(self.circuits)
        if fold:
            from .peephole import fold_program
            compiled = fold_program(compiled)
        return compiled

    def _get_engine(self, compiled: CompiledProgram):
        from .engine import TableauEngine          # imports torch; kept out of module import time
        key = (compiled.num_qudits, compiled.dimension, compiled.ops.tobytes(),
               compiled.noise_thresh24.tobytes(), compiled.noise_channel.tobytes())
        if self._engine is None or self._engine_key != key:
            self._engine = TableauEngine(compiled, self.device)
            self._engine_key = key
        return self._engine

    def _initial_store(self, engine, shots: int):
        """Device store pre-loaded with the user-supplied initial tableau (None when starting from |0...0>)."""
        if self.initial_tableau is None:
            return None
        import torch
        t = self.initial_tableau
        if t.num_qudits != engine.prog.num_qudits or t.dimension != engine.prog.dimension:
            raise ValueError("initial tableau does not match the circuit's qudit count / dimension")
        img = torch.from_numpy(t.pack(engine.layout.np)).to(engine.device)
        return img.unsqueeze(0).repeat(shots, 1).contiguous()

    @property
    def stabilizer_tableau(self) -> ExtendedTableau:
        """Tableau of the LAST shot of the last simulate() (reference keeps it in self.stabilizer_tableau).

        Materialised on first access by re-simulating that one shot with the same counter-based random
        draws and exporting it, so large runs never hold per-shot tableaus in HBM for this purpose.
        """
        if self._tableau_thunk is not None:
            self._tableau_cache = self._tableau_thunk()
            self._tableau_thunk = None
        return self._tableau_cache

    @stabilizer_tableau.setter
    def stabilizer_tableau(self, value):
        self._tableau_cache = value
        self._tableau_thunk = None

    # ---- simulate -----------------------------------------------------------------------------------
    def simulate_records(self, shots: int = 1, *, seed: Optional[int] = None, shot_offset: int = 0,
                         replay_meas=None, replay_noise=None, mode: Optional[str] = None,
                         distributed: bool = False, method: str = "tableau") -> RecordTable:
        """Run `shots` shots on the GPU and return the packed records (no Python objects).

        method="tableau" (default): one full stabilizer tableau per shot.
        method="frame": the reference's default multi-shot shortcut (sdim/program.py:244-265) — shot 0 is one
        noiseless reference tableau shot, shots 1.. are Pauli frames propagated on the GPU (sdimb_frames)."""
        compiled = self._compiled(fold=self.fold_gates)
        if method not in ("tableau", "frame"):
            raise ValueError("method must be 'tableau' or 'frame'")
        if distributed:
            from .dist import broadcast_seed
            seed = broadcast_seed(seed)         # one seed for all ranks (rank 0's draw when none was given)
        elif seed is None:
            seed = random.getrandbits(63)       # follows the user's random.seed(), like the reference's draws
        if method == "frame":
            if replay_meas is not None or replay_noise is not None or shot_offset != 0 or mode is not None or distributed:
                raise ValueError("method='frame' takes none of replay_meas, replay_noise, shot_offset, mode, distributed: "
                                 "frames are drawn per extra shot (ids 1..shots-1) around one reference tableau shot")
            rec = self._run_frames(compiled, shots, seed)
        elif distributed:
            from .dist import simulate_sharded
            rec = simulate_sharded(self, compiled, shots, seed, mode=mode, shot_offset=shot_offset,
                                   replay_meas=replay_meas, replay_noise=replay_noise)
        else:
            rec = self._run_local(compiled, shots, shot_offset, seed, replay_meas, replay_noise, mode, split=True)
        values, det = rec if isinstance(rec, tuple) else R.split(rec)
        table = RecordTable(values=values, deterministic=det,
                            meas_qudit=compiled.meas_qudit, meas_round=compiled.meas_round,
                            seed=seed, shot_offset=shot_offset)
        self.last_records = table
        return table

    def _run_frames(self, compiled, shots, seed) -> np.ndarray:
        import torch
        engine = self._get_engine(compiled)
        # reference shot: N1 is the identity there (sdim/program.py:31,245-247)
        if compiled.dimension >= R.WIDE_MIN_DIMENSION:
            raise ValueError("method='frame' supports dimensions up to 127; larger primes run the tableau path")
        quiet = torch.zeros((1, compiled.n_noise, 2), dtype=torch.uint8) if compiled.n_noise else None
        store = self._initial_store(engine, 1)
        if store is None:
            store = engine.alloc_tableau(1)
            engine.init_tableau(store)
        ref = engine.run(1, 0, seed, None, quiet, tableau=store, fresh=False, keep_tableau=True)
        out = np.empty((shots, compiled.n_meas), dtype=np.uint8)
        out[:1] = ref.cpu().numpy()
        if shots > 1:
            out[1:] = engine.run_frames(shots - 1, ref[0], 1, seed).cpu().numpy()
        # like the reference, which leaves the reference shot's final tableau in self.stabilizer_tableau
        # (sdim/program.py:245-247)
        self.stabilizer_tableau = ExtendedTableau.from_arrays(compiled.num_qudits, compiled.dimension,
                                                              engine.export(store, 0))
        return out

    WAVE_RECORD_BYTES = 256 << 20      # records per launch when the shots do not need an HBM tableau each
    WAVE_DEVICE_BYTES = 24 << 30       # per-shot device state (HBM tableaus, generator-major slabs) of one wave

    def _run_local(self, compiled, shots, shot_offset, seed, replay_meas, replay_noise, mode, split: bool = False,
                   on_device: bool = False):
        """Packed record bytes uint8[shots, n_meas]; with `split`, (values, deterministic) instead — unpacked wave
        by wave while the device is busy with the next wave.  With `on_device` the packed records stay in ONE device
        tensor (the sharded path gathers them over NCCL from there); waves then only bound the HBM tableau store."""
        import torch
        engine = self._get_engine(compiled)
        rdt, tdt = R.np_dtype(compiled.dimension), R.torch_dtype(compiled.dimension)
        rm = None if replay_meas is None else R.to_device(replay_meas, compiled.dimension)
        rn = None if replay_noise is None else R.to_device(replay_noise, compiled.dimension)
        # Shots run in waves sized to the device: a run that needs one HBM tableau per shot (uint8 lanes in global
        # mode, or a user-supplied initial tableau) must fit next to the records; resident kernels take any count.
        # ... and the two-kernel bit-plane path keeps one generator-major slab per shot of the wave in scratch.
        dev_per_shot = engine.device_bytes_per_shot(max(shots, 1), mode, fresh=self.initial_tableau is None) if shots > 0 else 0
        need_tab = dev_per_shot > 0
        wave = shots
        if need_tab and shots > 0:
            free_bytes, _total = torch.cuda.mem_get_info(engine.device)
            per_shot = dev_per_shot + (3 * compiled.n_meas + 2 * compiled.n_noise) * np.dtype(rdt).itemsize
            budget = min(int(0.6 * free_bytes), self.WAVE_DEVICE_BYTES)
            wave = max(1, min(shots, budget // max(per_shot, 1)))
            wave = min(wave, max(1, self.WAVE_RECORD_BYTES // max(compiled.n_meas * np.dtype(rdt).itemsize, 1)))
        elif shots > 0:
            wave = max(1, min(shots, self.WAVE_RECORD_BYTES // max(compiled.n_meas * np.dtype(rdt).itemsize, 1)))
        if on_device:
            full = torch.empty((shots, compiled.n_meas), dtype=tdt, device=engine.device)
            step = max(wave, 1) if need_tab else max(shots, 1)
            for lo in range(0, shots, step):
                hi = min(shots, lo + step)
                store = self._initial_store(engine, hi - lo)
                engine.run(hi - lo, shot_offset + lo, seed, None if rm is None else rm[lo:hi],
                           None if rn is None else rn[lo:hi], mode=mode, tableau=store, fresh=store is None,
                           records=full[lo:hi])
                del store
            self._set_last_shot_thunk(compiled, engine, shots, shot_offset, seed, rm, rn, mode)
            return full
        out = np.empty((shots, compiled.n_meas), dtype=rdt)
        det = np.empty((shots, compiled.n_meas), dtype=bool) if split else None
        det_bit, val_mask = R.masks(rdt)

        def deliver(lo, hi, packed):
            packed = R.unsigned(packed)
            if split:
                np.bitwise_and(packed, rdt(val_mask), out=out[lo:hi])
                np.not_equal(packed & rdt(det_bit), 0, out=det[lo:hi])
            else:
                out[lo:hi] = packed

        n_waves = (shots + max(wave, 1) - 1) // max(wave, 1) if shots > 0 else 0
        if n_waves <= 1:
            if shots > 0:
                store = self._initial_store(engine, shots)
                rec = engine.run(shots, shot_offset, seed, rm, rn, mode=mode, tableau=store, fresh=store is None)
                deliver(0, shots, rec.cpu().numpy())
                del store, rec
        else:
            # Several waves: records leave the device through two pinned staging buffers on a copy stream, so the
            # device-to-host transfer and the host-side copy of wave k overlap the simulation of wave k + 1.
            dev = engine.device
            with torch.cuda.device(dev):
                compute, copier = torch.cuda.current_stream(dev), torch.cuda.Stream(dev)
                recs = [torch.empty((wave, compiled.n_meas), dtype=tdt, device=dev) for _ in range(2)]
                pins = [torch.empty((wave, compiled.n_meas), dtype=tdt, pin_memory=True) for _ in range(2)]
                copied = [None, None]                            # (event, lo, hi) of the copy in flight per buffer

                def collect(slot):
                    if copied[slot] is not None:
                        ev, lo, hi = copied[slot]
                        ev.synchronize()
                        deliver(lo, hi, pins[slot][: hi - lo].numpy())
                        copied[slot] = None

                for k, lo in enumerate(range(0, shots, wave)):
                    hi, slot = min(shots, lo + wave), k & 1
                    collect(slot)                                # wave k - 2 has left both buffers of this slot
                    store = self._initial_store(engine, hi - lo)
                    engine.run(hi - lo, shot_offset + lo, seed,
                               None if rm is None else rm[lo:hi], None if rn is None else rn[lo:hi],
                               mode=mode, tableau=store, fresh=store is None, records=recs[slot][: hi - lo])
                    done = torch.cuda.Event()
                    done.record(compute)
                    with torch.cuda.stream(copier):
                        copier.wait_event(done)
                        pins[slot][: hi - lo].copy_(recs[slot][: hi - lo], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(copier)
                    copied[slot] = (ev, lo, hi)
                    collect(slot ^ 1)                            # host-side copy of wave k - 1 while wave k runs
                    del store
                collect(0)
                collect(1)

        self._set_last_shot_thunk(compiled, engine, shots, shot_offset, seed, rm, rn, mode)
        return (out, det) if split else out

    def _set_last_shot_thunk(self, compiled, engine, shots, shot_offset, seed, rm, rn, mode):
        """`stabilizer_tableau` of a batched run = the tableau of its last shot, re-simulated on demand."""
        def last_shot_tableau(last=shots - 1):
            one = engine.alloc_tableau(1)
            if self.initial_tableau is not None:
                one.copy_(self._initial_store(engine, 1))
            engine.run(1, shot_offset + last, seed,
                       None if rm is None else rm[last:last + 1], None if rn is None else rn[last:last + 1],
                       keep_tableau=True, mode=mode, tableau=one, fresh=self.initial_tableau is None)
            return ExtendedTableau.from_arrays(compiled.num_qudits, compiled.dimension, engine.export(one, 0))

        self._tableau_thunk = last_shot_tableau if shots > 0 else None

    def simulate(self, shots: int = 1, show_measurement: bool = False, record_tableau: bool = False,
                 force_tableau: bool = False, verbose: bool = False, show_gate: bool = False, exact: bool = False,
                 options: Optional[SimulationOptions] = None, *, seed: Optional[int] = None,
                 replay_meas=None, replay_noise=None, mode: Optional[str] = None, distributed: bool = False,
                 method: str = "tableau"):
        """Same call as the reference's Program.simulate (sdim/program.py:206-267).

        Returns a flat list of MeasurementResult for shots == 1, else `[qudit][round][shot]`.
        By default every shot runs on the tableau path (what the reference does under `force_tableau=True`);
        `method="frame"` selects the reference's default multi-shot mechanism instead (one reference tableau shot +
        Pauli frames, sdim/program.py:244-265) unless `force_tableau` / `record_tableau` is set.
        `exact` only concerns composite dimensions and is ignored.
        """
        if options is None:
            options = SimulationOptions(shots, show_measurement, record_tableau, force_tableau, verbose, show_gate,
                                        exact)
        compiled = self._compiled()
        n = compiled.num_qudits
        if options.verbose or options.show_gate or options.record_tableau:
            tables = self._simulate_stepped(compiled, options, seed, replay_meas, replay_noise, mode)
            values, det, snaps = tables
        else:
            use_frames = method == "frame" and options.shots > 1 and not options.force_tableau
            table = self.simulate_records(options.shots, seed=seed, replay_meas=replay_meas,
                                          replay_noise=replay_noise, mode=mode, distributed=distributed,
                                          method="frame" if use_frames else "tableau")
            values, det, snaps = table.values, table.deterministic, None

        # group as measurement_results[qudit][round][shot]                    (program.py:321-332)
        results: list = [[] for _ in range(n)]
        for k in range(compiled.n_meas):
            q = int(compiled.meas_qudit[k])
            vals, dets = values[:, k].tolist(), det[:, k].tolist()
            if snaps is None:
                column = [MeasurementResult(q, dt, v) for v, dt in zip(vals, dets)]
            else:
                column = [MeasurementResult(q, dt, v, snaps[s][k]) for s, (v, dt) in enumerate(zip(vals, dets))]
            results[q].append(column)
        self.measurement_results = results
        if options.show_measurement:
            self.print_measurements()
        if options.shots == 1:
            return [rounds[0] for per_qudit in results for rounds in per_qudit]
        return results

    def _simulate_stepped(self, compiled, options, seed, replay_meas, replay_noise, mode):
        """Host-stepped execution for verbose / show_gate / record_tableau: the op stream is issued one op
        per launch on a persistent device store so the tableau can be exported between ops."""
        import torch
        engine = self._get_engine(compiled)
        shots = options.shots
        if seed is None:
            seed = random.getrandbits(63)
        rm = None if replay_meas is None else R.to_device(replay_meas, compiled.dimension)
        rn = None if replay_noise is None else R.to_device(replay_noise, compiled.dimension)
        values = np.zeros((shots, compiled.n_meas), dtype=R.np_dtype(compiled.dimension))
        det = np.zeros((shots, compiled.n_meas), dtype=bool)
        snaps = [[None] * compiled.n_meas for _ in range(shots)] if options.record_tableau else None
        # the reference's step numbering (sdim/program.py:301,311-346): `time` restarts in every circuit and counts I
        # gates, "Final step" compares it with the total number of operations of all circuits
        length = sum(len(c.operations) for c in self.circuits)
        last = None
        for s in range(shots):
            store = self._initial_store(engine, 1)
            if store is None:
                store = engine.alloc_tableau(1)
                engine.init_tableau(store)
            rec = torch.zeros((1, compiled.n_meas), dtype=R.torch_dtype(compiled.dimension), device=engine.device)
            srm = None if rm is None else rm[s:s + 1]
            srn = None if rn is None else rn[s:s + 1]
            steps = [(t, g) for c in self.circuits for t, g in enumerate(c.operations)]
            i = -1                                       # index of the current instruction in the compiled stream (I gates are not in it)
            for time, g in steps:
                if time == 0 and options.verbose:
                    print("Initial state")
                    ExtendedTableau.from_arrays(compiled.num_qudits, compiled.dimension, engine.export(store, 0)).print_tableau()
                    print("\n")
                op = slot = -1
                if g.gate_id != 0:
                    i += 1
                    engine.run(1, s, seed, srm, srn, keep_tableau=True, mode=mode, tableau=store, fresh=False,
                               op_range=(i, i + 1), records=rec)
                    op, _, _, slot = (int(v) for v in compiled.ops[i])
                if snaps is not None and op in MEASURE_OPS:
                    arrays = engine.export(store, 0)
                    if op == OP_RESET:
                        # The reference snapshots right after measure(), before the RESET correction
                        # (program.py:323-324 precede :335-339); the device op does both, so the correction's
                        # phase update is taken back out of the exported copy.
                        outcome = int(R.split(rec[0, slot:slot + 1].cpu().numpy())[0][0])
                        arrays = undo_reset_correction(arrays, int(compiled.ops[i][1]), outcome, compiled.dimension)
                    snaps[s][slot] = ExtendedTableau.from_arrays(compiled.num_qudits, compiled.dimension, arrays)
                if options.show_gate:
                    info = g.target_index if g.target_index is not None else ""
                    print("Time step" if time < length - 1 else "Final step", time, "\t", g.name, g.qudit_index, info)
                if options.verbose:
                    ExtendedTableau.from_arrays(compiled.num_qudits, compiled.dimension,
                                                engine.export(store, 0)).print_tableau()
                    print("\n")
            values[s], det[s] = R.split(rec.cpu().numpy()[0])
            last = store
        if last is not None:
            self.stabilizer_tableau = ExtendedTableau.from_arrays(compiled.num_qudits, compiled.dimension,
                                                                  engine.export(last, 0))
        return values, det, snaps

    def apply_gate(self, instruc: CircuitInstruction) -> Optional[MeasurementResult]:
        """Apply one instruction to the current `stabilizer_tableau` (reference: program.py:367-385)."""
        import torch
        from .engine import TableauEngine
        from .ir import NUM_OPS
        if instruc.gate_id is None or not 0 <= instruc.gate_id < NUM_OPS:
            raise ValueError("Invalid gate value")
        current = self.stabilizer_tableau
        one = Circuit(current.num_qudits, current.dimension)
        one.operations.append(instruc)
        compiled = compile_circuits([one])
        engine = TableauEngine(compiled, self.device)
        store = torch.from_numpy(current.pack(engine.layout.np)).to(engine.device).unsqueeze(0).contiguous()
        rec = engine.run(1, 0, random.getrandbits(63), keep_tableau=True, tableau=store, fresh=False)
        out = None
        arrays = engine.export(store, 0)
        if compiled.n_meas:
            rv, rd = R.split(rec.cpu().numpy()[0, :1])
            out = MeasurementResult(int(compiled.meas_qudit[0]), bool(rd[0]), int(rv[0]))
            if instruc.gate_id == OP_RESET:
                # the reference's apply_reset only measures (tableau_gates.py:331-346); the X correction belongs to
                # its shot loop (program.py:335-339).  The device op does both, so the correction is taken back out.
                arrays = undo_reset_correction(arrays, int(compiled.meas_qudit[0]), int(rv[0]), compiled.dimension)
        self.stabilizer_tableau = ExtendedTableau.from_arrays(compiled.num_qudits, compiled.dimension, arrays)
        return out

    # ---- reference helpers kept for drop-in compatibility -------------------------------------------
    @staticmethod
    def _results_to_array(measurements: list) -> np.ndarray:
        """Nested MeasurementResult lists -> structured array [n_qudits, max_rounds] (program.py:387-421)."""
        if not measurements or not measurements[0]:
            raise ValueError("Empty or invalid measurement results format")
        first = measurements[0][0]
        if not isinstance(first, (list, MeasurementResult)):
            raise ValueError("Invalid measurement results format")
        max_rounds = max(len(m) for m in measurements)
        out = np.empty((len(measurements), max_rounds), dtype=MEASUREMENT_DTYPE)
        for q, per_qudit in enumerate(measurements):
            for r, item in enumerate(per_qudit):
                m = item[0] if isinstance(first, list) else item
                out[q, r] = (m.qudit_index, r, 0, m.deterministic, m.measurement_value)
        return out

    def _combine_results(self, frame_results) -> list:
        """Append extra shots held in a structured array [n_qudits, rounds, shots] (program.py:423-454)."""
        n_qudits, _, extra = frame_results.shape
        for q in range(n_qudits):
            for r in range(len(self.measurement_results[q])):
                cell = frame_results[q, r]
                self.measurement_results[q][r].extend(
                    MeasurementResult(int(cell[s]["qudit_index"]), bool(cell[s]["deterministic"]),
                                      int(cell[s]["measurement_value"])) for s in range(extra))
        return self.measurement_results

    def _build_ir(self, circuits: List[Circuit], extra_shots: int):
        """The reference's IR + pre-sampled noise (program.py:456-526), for callers that consume that format.
        The GPU path does not use it (noise is drawn on the device); semantics are identical."""
        d = self.circuits[0].dimension
        triples, noise = [], []
        for circuit in circuits:
            for ins in circuit.operations:
                if ins.gate_id == 0:
                    continue
                triples.append((ins.gate_id, ins.qudit_index, -1 if ins.target_index is None else ins.target_index))
                if ins.gate_id != 17:
                    continue
                params = ins.params or {}
                channel = params.get("noise_channel", params.get("channel", "d"))
                a = np.zeros(extra_shots, dtype=np.int64)
                b = np.zeros(extra_shots, dtype=np.int64)
                if channel == "d":
                    r = np.random.randint(1, d * d, size=extra_shots)
                    a, b = r % d, r // d
                elif channel == "f":
                    a = np.random.randint(1, d, size=extra_shots)
                elif channel == "p":
                    b = np.random.randint(1, d, size=extra_shots)
                keep_identity = np.random.uniform(0.0, 1.0, size=extra_shots) < 1.0 - float(params.get("prob", 0.01))
                a[keep_identity] = 0
                b[keep_identity] = 0
                noise.append(np.stack((a, b), axis=1))
        ir_dtype = np.dtype([("gate_id", np.int64), ("qudit_index", np.int64), ("target_index", np.int64)])
        ir = np.array(triples, dtype=ir_dtype)
        noise_array = np.array(noise, dtype=np.int64) if noise else np.empty((1, extra_shots, 2), dtype=np.int64)
        return ir, noise_array

    def append_circuit(self, circuit: Circuit) -> None:
        tail = self.circuits[-1]
        if tail.num_qudits < circuit.num_qudits:
            tail.num_qudits = circuit.num_qudits
        else:
            circuit.num_qudits = tail.num_qudits
        if tail.dimension != circuit.dimension:
            raise ValueError("Circuits must have the same dimension")
        self.circuits.append(circuit)

    def print_measurements(self) -> None:
        """Print stored results (program.py:547-571; unlike the reference this never re-simulates, B-10)."""
        shot_count = max((len(group) for per_qudit in self.measurement_results for group in per_qudit), default=0)
        if shot_count == 0:
            print("No measurements recorded.")
            return
        if shot_count == 1:
            for per_qudit in self.measurement_results:
                for group in per_qudit:
                    print(group[0])
            return
        for s in range(shot_count):
            print(f"Shot {s + 1}:")
            for per_qudit in self.measurement_results:
                for r, group in enumerate(per_qudit):
                    print(f"{group[s]} during measurement {r}")
            print()

    def __str__(self) -> str:
        return str(self.stabilizer_tableau)
