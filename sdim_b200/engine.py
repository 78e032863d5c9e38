"""Device-side driver: owns the PyTorch buffers and calls the C ABI (include/sdimb.h).

`TableauEngine` is the GPU counterpart of the body of `Program._simulate_tableau`
(reference: sdim/program.py:269-365): it uploads the compiled op stream once,
then `run()` simulates a contiguous range of shots — one tableau per shot — and
returns the packed record matrix uint8[shots, n_meas] (bit 7 = deterministic,
low bits = value).  PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _native as N
from . import records as R
from .ir import CompiledProgram


def _require_cuda(device) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("sdim_b200 runs the tableau path on a CUDA device only; no GPU is visible "
                           "and there is no CPU fallback")
    dev = torch.device(device if device is not None else "cuda")
    if dev.type != "cuda":
        raise RuntimeError(f"sdim_b200 needs a CUDA device, got {dev}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None or t.numel() == 0 else t.data_ptr()


class TableauEngine:
    # auto: bit-plane resident for d in {2,3} when it fits (bit planes on a global image when it does not), else
    # uint8 lanes resident when they fit, else global (HBM store).  On the HBM store the library runs one CTA per shot, or one thread-block cluster per shot for
    # large tableaus with few shots; "cluster" / "global-cta" pin that choice (tests, A/B timings).
    MODES = {None: 0, "auto": 0, "global": N.FORCE_GLOBAL, "resident": N.FORCE_RESIDENT, "lanes": N.FORCE_LANES,
             "planes": N.FORCE_PLANES, "planes-global": N.FORCE_PLANES | N.FORCE_GLOBAL,
             "planes-warp": N.FORCE_PLANES | N.NO_TILE,      # one shot per warp even where the tile interpreter fits
             "cluster": N.FORCE_GLOBAL | N.CLUSTER, "global-cta": N.FORCE_GLOBAL | N.NO_CLUSTER}

    def __init__(self, prog: CompiledProgram, device=None):
        self.prog = prog
        self.device = _require_cuda(device)
        self.layout = N.layout(prog.num_qudits, prog.dimension)
        self.lib = N.lib()
        dev = self.device
        self.ops = torch.from_numpy(np.ascontiguousarray(prog.ops)).to(dev)
        # layered stream for the multi-warp bit-plane interpreter (same ops, commuting reorder + barriers)
        sched = N.schedule(prog.num_qudits, prog.ops) if prog.n_ops else None
        self.ops_sched = torch.from_numpy(sched).to(dev) if prog.n_ops else None
        # the run of M ops that ends the stream ("measure every qudit"): the library may hand it to a second kernel
        self.tail_run_len = N.tail_run(sched) if prog.n_ops else 0
        # ... and the gates in front of it, when it holds every measurement, as pre-decoded per-warp streams
        self.gate_stream = None
        if self.tail_run_len and self.tail_run_len == prog.n_meas and prog.dimension in (2, 3):
            gs = N.gate_stream(prog.num_qudits, prog.dimension, sched[: sched.shape[0] - self.tail_run_len])
            self.gate_stream = torch.from_numpy(gs).to(dev) if gs is not None else None
        # uint8 lanes on the HBM store run the stream as written: its trailing run of plain M ops (lanes_gm.cuh),
        # if at least as long as the scheduler's run threshold
        t = 0
        while t < prog.n_ops and int(prog.ops[prog.n_ops - 1 - t][0]) == 14:
            t += 1
        self.tail_run_len_raw = t if t >= max(prog.num_qudits // 8, 4) else 0
        self.noise_thresh = torch.from_numpy(prog.noise_thresh24.astype(np.int64)).to(dev).to(torch.int32) \
            if prog.n_noise else None
        if prog.n_noise:
            # uint32 thresholds travel as int32 bit patterns (values <= 2^24 are unaffected)
            self.noise_channel = torch.from_numpy(prog.noise_channel.copy()).to(dev)
        else:
            self.noise_channel = None
        self._planes_fit: Optional[bool] = None          # do the bit planes of one shot fit in shared memory?
        self._scratch: Dict[tuple, torch.Tensor] = {}   # per (mode flags, tail-run shots): counter, images, slabs
        self.tableau: Optional[torch.Tensor] = None     # uint8 [shots, shot_bytes] of the last run that kept it
        self.tableau_shots = 0
        self.rec_dtype = R.torch_dtype(prog.dimension)   # packed records / replay arrays: uint8, int16 bits for d > 127

    # ------------------------------------------------------------------------------------------
    def plan(self, mode: Optional[str] = None, fresh: bool = True, keep_tableau: bool = False):
        """(kernel name, needs_tableau) for a run in `mode`."""
        if mode not in self.MODES:
            raise ValueError(f"mode must be one of {sorted(k for k in self.MODES if k)}")
        flags = self.MODES[mode] | (N.FRESH if fresh else 0) | (N.WRITEBACK if keep_tableau else 0)
        k, need = N.plan(self.prog.num_qudits, self.prog.dimension, flags)
        return N.KERNEL_NAMES[k], need

    def _auto_mode(self, mode: Optional[str], shots: int, keep_tableau: bool = False) -> Optional[str]:
        """`auto` refined by the shot count: a few shots of a d = 2, 3 tableau too large for shared memory run one
        uint8 tableau per thread-block cluster ("lanes") instead of one CTA per shot on bit planes (n = 2048, 8 shots:
        12.8 ms against 30 ms) — as long as every shot gets a cluster."""
        if mode in (None, "auto") and self.prog.dimension in (2, 3) and self.plan(None)[0] == "planes-global":
            if self._planes_fit is None:
                try:
                    self.plan("planes")
                    self._planes_fit = True
                except ValueError:
                    self._planes_fit = False
            if not self._planes_fit and self.cluster_size(shots, "lanes") > 0:
                return "lanes"
        # every measurement in the trailing run and the gates compiled into per-warp streams: the two-kernel path
        # (gate_stream_kernel + run_tail_kernel) also beats the shared-memory interpreter where that fits
        # (d = 3: n = 160 12.4 ms against 20.5 ms per 16 384 shots, n = 192 10.4 / 19.7; d = 2, n = 200: 7.5 / 13.5)
        # ... and the tile interpreter above 64 qudits (d = 3, n = 128: 10.7 / 24.8 ms; d = 2, n = 100: 5.1 / 5.6; at
        # n = 64 the tiles win, 3.8 / 4.9)
        if mode in (None, "auto") and self.gate_stream is not None and not keep_tableau:
            kernel = self.plan(None)[0]
            if kernel == "planes-resident" or (kernel == "planes-tile" and self.prog.num_qudits > 64):
                return "planes-global"
        # uint8 lanes that would run resident in shared memory: from 96 qudits on, the HBM store with the trailing run of
        # M ops in the generator-major kernel (lanes_gm.cuh) is faster (d = 5, 4096 shots: n = 96 18.0 vs 20.2 ms,
        # n = 128 22.4 / 44.5, n = 200 17.8 / 52.6; below that shared memory wins, n = 64: 7.6 / 8.4)
        if mode in (None, "auto") and self.tail_run_len_raw and not keep_tableau and self.prog.num_qudits >= 96 and \
                self.plan(None)[0] == "lanes-resident":
            return "global"
        return mode

    def cluster_size(self, shots: int, mode: Optional[str] = None) -> int:
        """Thread-block cluster size the library would use for `shots` shots in `mode` (0: one CTA per shot)."""
        with torch.cuda.device(self.device):
            return int(self.lib.sdimb_cluster_size(self.prog.num_qudits, self.prog.dimension, shots, self.MODES[mode]))

    def _scratch_for(self, mode_flags: int, shots: int = 0) -> Optional[torch.Tensor]:
        """Scratch of the plane kernels; `shots` > 0: room for one image per shot (tail run in a second kernel)."""
        have = self._scratch.get(mode_flags)
        nbytes = int(self.lib.sdimb_scratch_bytes_shots(self.prog.num_qudits, self.prog.dimension, mode_flags, shots))
        if have is None or have.numel() < nbytes:
            have = torch.empty(nbytes, dtype=torch.uint8, device=self.device) if nbytes else None
            self._scratch[mode_flags] = have
        return have

    def device_bytes_per_shot(self, shots: int, mode: Optional[str] = None, fresh: bool = True,
                              keep_tableau: bool = False) -> int:
        """Device memory one more shot of a `run` costs besides its records: the HBM tableau where the kernel needs
        one, and the per-shot slab of the two-kernel bit-plane path (B + QX, 80 KB at n = 256).  Callers size their
        waves with it (Program.simulate_records)."""
        mode = self._auto_mode(mode, shots, keep_tableau)
        kernel, need_tab = self.plan(mode, fresh, keep_tableau)
        per = self.layout.shot_bytes if need_tab else 0
        if kernel == "planes-global" and self.tail_run_len and not keep_tableau:
            n, d, flags = self.prog.num_qudits, self.prog.dimension, self.MODES[mode]
            s1 = int(self.lib.sdimb_scratch_bytes_shots(n, d, flags, 1024))
            s2 = int(self.lib.sdimb_scratch_bytes_shots(n, d, flags, 2048))
            per += max(0, (s2 - s1) // 1024)
        return per

    def fits_resident(self, mode: Optional[str] = None) -> bool:
        return not self.plan(mode)[1]

    def alloc_tableau(self, shots: int) -> torch.Tensor:
        return torch.empty((shots, self.layout.shot_bytes), dtype=torch.uint8, device=self.device)

    def init_tableau(self, tab: torch.Tensor) -> None:
        stream = torch.cuda.current_stream(self.device).cuda_stream
        N.check(self.lib.sdimb_init(tab.data_ptr(), self.prog.num_qudits, self.prog.dimension, tab.shape[0], stream))

    def run(self, shots: int, shot_offset: int = 0, seed: int = 0,
            replay_meas: Optional[torch.Tensor] = None, replay_noise: Optional[torch.Tensor] = None,
            keep_tableau: bool = False, mode: Optional[str] = None,
            tableau: Optional[torch.Tensor] = None, fresh: bool = True,
            op_range: Optional[tuple] = None, records: Optional[torch.Tensor] = None,
            time_kernels: bool = False) -> torch.Tensor:
        """Simulate local shots [0, shots) with global ids shot_offset + local; returns records on device.

        replay_meas  uint8[shots, n_meas]      outcome to use where measurement k is random (else Philox)
        replay_noise uint8[shots, n_noise, 2]  (a, b) of every N1 event (else Philox)
        tableau      existing store to continue from (fresh=False) or to fill (keep_tableau=True)
        op_range     (lo, hi) slice of the op stream, for host-stepped execution
        """
        prog, L, dev = self.prog, self.layout, self.device
        mode = self._auto_mode(mode, shots, keep_tableau or op_range is not None)
        kernel, need_tab = self.plan(mode, fresh, keep_tableau)
        flags = self.MODES[mode] | (N.FRESH if fresh else 0) | (N.WRITEBACK if keep_tableau else 0)
        # layered stream: the multi-warp bit-plane CTAs and the cluster interpreter's gate groups want it; the
        # one-CTA lane kernels ignore the layering (the reorder is exact, tests/test_schedule.py)
        wants_layers = kernel in ("planes-resident", "planes-global") or \
            (kernel == "lanes-global" and self.cluster_size(shots, mode) > 0)
        use_sched = wants_layers and op_range is None and self.ops_sched is not None
        if use_sched:
            flags |= N.SCHEDULED
        if time_kernels:
            flags |= N.TIME_KERNELS       # per-kernel events, read with _native.kernel_times()
        with torch.cuda.device(dev):
            if need_tab:
                if tableau is None:
                    tableau = self.alloc_tableau(shots)
                if tableau.shape != (shots, L.shot_bytes) or tableau.dtype != torch.uint8 or not tableau.is_contiguous():
                    raise ValueError("tableau buffer has the wrong shape/dtype for this program")
            if records is None:
                records = torch.empty((shots, prog.n_meas), dtype=self.rec_dtype, device=dev)
            elif records.dtype != self.rec_dtype:
                raise ValueError(f"records buffer must be {self.rec_dtype} for dimension {prog.dimension}")
            if replay_meas is not None:
                replay_meas = replay_meas.to(device=dev, dtype=self.rec_dtype).contiguous()
                if replay_meas.shape != (shots, prog.n_meas):
                    raise ValueError("replay_meas must be [shots, n_meas]")
            if replay_noise is not None:
                replay_noise = replay_noise.to(device=dev, dtype=self.rec_dtype).contiguous()
                if replay_noise.shape != (shots, prog.n_noise, 2):
                    raise ValueError("replay_noise must be [shots, n_noise, 2]")
            lo, hi = (0, prog.n_ops) if op_range is None else op_range
            a = N.SdimbRunArgs()
            a.struct_size = C.sizeof(N.SdimbRunArgs)
            a.flags = flags
            a.n, a.d = prog.num_qudits, prog.dimension
            a.shots, a.shot_offset = shots, shot_offset
            a.tableau = _ptr(tableau) if need_tab else None
            if use_sched:
                a.ops, a.n_ops = self.ops_sched.data_ptr(), self.ops_sched.shape[0]
            else:
                a.ops = (self.ops.data_ptr() + 16 * lo) if hi > lo else None
                a.n_ops = hi - lo
            a.records = _ptr(records)
            a.n_meas = prog.n_meas
            a.rec_stride = records.stride(0) if prog.n_meas else 0
            a.replay_meas = _ptr(replay_meas)
            a.replay_noise = _ptr(replay_noise)
            a.noise_thresh24 = _ptr(self.noise_thresh)
            a.noise_channel = _ptr(self.noise_channel)
            a.n_noise = prog.n_noise
            a.seed = seed & 0xFFFFFFFFFFFFFFFF
            a.stream = torch.cuda.current_stream(dev).cuda_stream
            tail = self.tail_run_len if (use_sched and not keep_tableau) else 0
            if kernel == "lanes-global" and not use_sched and fresh and not keep_tableau and op_range is None:
                tail = self.tail_run_len_raw
            scratch = self._scratch_for(self.MODES[mode], shots if tail else 0)
            a.scratch, a.scratch_bytes = _ptr(scratch), (scratch.numel() if scratch is not None else 0)
            a.tail_run_len = tail
            if tail and self.gate_stream is not None:
                a.gate_stream, a.gate_stream_rows = self.gate_stream.data_ptr(), self.gate_stream.shape[0]
            N.check(self.lib.sdimb_run(C.byref(a)))
        if keep_tableau:
            self.tableau, self.tableau_shots = tableau, shots
        self._keepalive = (replay_meas, replay_noise)   # until the stream has consumed them
        return records

    def run_frames(self, shots: int, reference: torch.Tensor, shot_offset: int = 1, seed: int = 0,
                   replay_z0: Optional[torch.Tensor] = None, replay_zm: Optional[torch.Tensor] = None,
                   replay_noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Pauli-frame sampler (sdimb_frames; reference: simulate_frame, sdim/program.py:45-165).

        `reference` = packed records uint8[n_meas] of one reference tableau shot.  Returns uint8[shots, n_meas]
        for the extra shots with global ids shot_offset .. shot_offset + shots - 1."""
        prog, dev = self.prog, self.device
        with torch.cuda.device(dev):
            records = torch.empty((shots, prog.n_meas), dtype=torch.uint8, device=dev)
            pitch = (shots + 127) // 128 * 128
            frames = torch.empty((2, prog.num_qudits, max(pitch, 1)), dtype=torch.uint8, device=dev)
            reference = reference.to(device=dev, dtype=torch.uint8).contiguous()

            def dev_u8(t, shape):
                if t is None:
                    return None
                t = t.to(device=dev, dtype=torch.uint8).contiguous()
                if tuple(t.shape) != shape:
                    raise ValueError(f"replay array must have shape {shape}")
                return t

            z0 = dev_u8(replay_z0, (shots, prog.num_qudits))
            zm = dev_u8(replay_zm, (shots, prog.n_meas))
            rn = dev_u8(replay_noise, (shots, prog.n_noise, 2))
            stream = torch.cuda.current_stream(dev).cuda_stream
            N.check(self.lib.sdimb_frames(prog.num_qudits, prog.dimension, shots, shot_offset, _ptr(self.ops),
                                          prog.n_ops, _ptr(reference), _ptr(records), prog.n_meas,
                                          records.stride(0) if prog.n_meas else 0, frames.data_ptr(),
                                          _ptr(z0), _ptr(zm), _ptr(rn), _ptr(self.noise_thresh),
                                          _ptr(self.noise_channel), prog.n_noise, seed & 0xFFFFFFFFFFFFFFFF, stream))
            torch.cuda.current_stream(dev).synchronize()      # replay buffers and frames die with this scope
        return records

    def export(self, tableau: torch.Tensor, shot: int) -> Dict[str, np.ndarray]:
        """Six int64 arrays of one shot in the reference orientation [qudit, generator]."""
        n, dev = self.prog.num_qudits, self.device
        with torch.cuda.device(dev):
            big = torch.empty((4, n, n), dtype=torch.int64, device=dev)
            small = torch.empty((2, n), dtype=torch.int64, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            N.check(self.lib.sdimb_export(tableau.data_ptr(), n, self.prog.dimension, shot,
                                          big[0].data_ptr(), big[1].data_ptr(), small[0].data_ptr(),
                                          big[2].data_ptr(), big[3].data_ptr(), small[1].data_ptr(), stream))
            big_h, small_h = big.cpu().numpy(), small.cpu().numpy()
        return {"x": big_h[0], "z": big_h[1], "p": small_h[0], "dx": big_h[2], "dz": big_h[3], "dp": small_h[1]}


def simulate_host(prog: CompiledProgram, shots: int, shot_offset: int = 0, seed: int = 0,
                  replay_meas: Optional[np.ndarray] = None, replay_noise: Optional[np.ndarray] = None,
                  mode: Optional[str] = None, out: Optional[np.ndarray] = None):
    """Host-buffer entry (sdimb_simulate_host): numpy in, numpy records out, plus device ms of the call.

    `out`: a C-contiguous [shots, n_meas] array of the record dtype to receive the records instead of a fresh one; if it
    lives in pinned memory (e.g. `torch.empty(..., pin_memory=True).numpy()`) the device writes it directly."""
    lib = N.lib()
    ops = np.ascontiguousarray(prog.ops, dtype=np.int32)
    rdt = R.np_dtype(prog.dimension)
    if out is None:
        rec = np.empty((shots, prog.n_meas), dtype=rdt)
    else:
        if out.shape != (shots, prog.n_meas) or out.dtype != rdt or not out.flags["C_CONTIGUOUS"] or not out.flags["WRITEABLE"]:
            raise ValueError(f"out must be a writeable C-contiguous {rdt.__name__}[{shots}, {prog.n_meas}] array")
        rec = out
    thr = np.ascontiguousarray(prog.noise_thresh24, dtype=np.uint32)
    ch = np.ascontiguousarray(prog.noise_channel, dtype=np.uint8)
    rm = None if replay_meas is None else np.ascontiguousarray(replay_meas, dtype=rdt)
    rn = None if replay_noise is None else np.ascontiguousarray(replay_noise, dtype=rdt)
    ms = C.c_float(0.0)

    def p(arr):
        return None if arr is None or arr.size == 0 else arr.ctypes.data

    N.check(lib.sdimb_simulate_host(prog.num_qudits, prog.dimension, shots, shot_offset, p(ops), prog.n_ops,
                                    p(rec), prog.n_meas, p(rm), p(rn), p(thr), p(ch), prog.n_noise,
                                    seed & 0xFFFFFFFFFFFFFFFF, TableauEngine.MODES[mode], C.byref(ms)))
    return rec, float(ms.value)
