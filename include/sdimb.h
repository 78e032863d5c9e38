/*
 * sdimb.h — C ABI of libsdimb, the B200 (sm_100a) stabilizer-tableau engine.
 *
 * Drop-in boundary for the prime-dimension tableau hot path of events555/sdim.
 * The reference has no native boundary (it is one Python package); its seam at
 * the right granularity is the batch call
 *     simulate_frame(ir_array, reference_results, n_qudits, dimension, extra_shots, noise_array)
 *         reference: sdim/program.py:45-47, called at sdim/program.py:253-262
 * i.e. "IR + noise in, records out, once per Program.simulate".  The entry points
 * below are what an FFI binding for that seam binds; INTEGRATION.md shows the
 * ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 (SDIMB_OK) or a negative SDIMB_E* code; nothing throws
 *   - device entry points are asynchronous on `stream` (a cudaStream_t passed as void*,
 *     NULL = legacy default stream) and never synchronise; pointers marked [device]
 *     are plain CUDA device addresses owned by the caller (PyTorch in this repo)
 *   - the device entry points keep no global mutable state: calls on distinct (stream, buffer) sets may run
 *     concurrently.  Only sdimb_simulate_host owns state (one reusable workspace per device) and serialises its
 *     callers per device: host threads driving different GPUs run concurrently
 *   - there is NO CPU fallback: without a CUDA device every launch returns SDIMB_ECUDA
 *
 * Tableau store (one tableau per shot; replaces the six int64 arrays of
 * sdim/tableau/dataclasses.py:24-39 + sdim/tableau/tableau_prime.py:24-26):
 *   lanes  np = n rounded up to 16;  W = 2*np generator lanes per row
 *          lane g in [0,n)      = stabilizer generator g   (reference column g of x_block/z_block)
 *          lane np+g            = destabilizer generator g (reference column g of destab_*_block)
 *          padding lanes hold 0 forever
 *   shot s at  tab + s*shot_bytes:
 *          row q (qudit q):  X[q][0..W) at q*2W,  Z[q][0..W) at q*2W + W      (uint8, reduced mod d)
 *          phases:           P[0..W)    at n*2W                                 (uint8, reduced mod order)
 *   order = 2d for d = 2, d for odd prime d (sdim/tableau/dataclasses.py:88-106)
 *
 * Primes 127 < d < 32768 ("uint16 lanes"): the same store with TWO bytes per entry (SdimbLayout.elem_bytes = 2; all
 * offsets double), and record / replay_meas / replay_noise elements are uint16 as well (SdimbLayout.rec_bytes = 2,
 * bit 15 of a record = deterministic flag).  The pointer types below stay uint8_t*: for those dimensions they are
 * read as uint16_t* (element strides rec_stride / n_meas / n_noise unchanged).  sdimb_frames stops at d = 127.
 */
#ifndef SDIMB_H
#define SDIMB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDIMB_VERSION 4

enum {
  SDIMB_OK = 0,
  SDIMB_EINVAL = -1,   /* bad argument (null pointer, n < 1, shots < 0, bad struct_size ...) */
  SDIMB_EDIM = -2,     /* dimension not a prime below 32768 (sdimb_frames: not a prime in [2, 127]) */
  SDIMB_EOP = -3,      /* invalid opcode / qudit index in the op stream ("Invalid gate value", sdim/program.py:381-382) */
  SDIMB_ECUDA = -4,    /* CUDA runtime error (no device, launch failure, out of memory) */
  SDIMB_ETOOBIG = -5   /* forced resident mode but one tableau does not fit in shared memory; d > 127 with n > 16384 */
};

/* Opcodes = gate ids of the reference gate table (sdim/gatedata.py:65-102). */
enum {
  SDIMB_OP_I = 0, SDIMB_OP_X = 1, SDIMB_OP_X_INV = 2, SDIMB_OP_Z = 3, SDIMB_OP_Z_INV = 4,
  SDIMB_OP_H = 5, SDIMB_OP_H_INV = 6, SDIMB_OP_P = 7, SDIMB_OP_P_INV = 8,
  SDIMB_OP_CNOT = 9, SDIMB_OP_CNOT_INV = 10, SDIMB_OP_CZ = 11, SDIMB_OP_CZ_INV = 12, SDIMB_OP_SWAP = 13,
  SDIMB_OP_M = 14, SDIMB_OP_M_X = 15, SDIMB_OP_RESET = 16, SDIMB_OP_N1 = 17,
  SDIMB_OP_BARRIER = 18   /* layer boundary emitted by sdimb_schedule; never written by users */
};
/* In a SCHEDULED stream (output of sdimb_schedule) bits 8..15 of the opcode field carry the warp that
 * executes the op inside its layer (index in the layer mod SDIMB_SCHED_WARPS) and bits 16..30 the index in the
 * layer itself (mod 2^15; the cluster interpreter deals layers over its own number of gate groups); bits 0..7
 * are the opcode.  An unscheduled stream has those bits 0. */
/* On an M op bits 8..15 carry the marks of a measurement RUN instead (sdimb_schedule: consecutive M ops with no
 * other op between): the bit-plane interpreter on a global image may execute such a run on a generator-major copy
 * of its image (sdim_b200/csrc/planes_gm.cuh); every other interpreter ignores them. */
#define SDIMB_GM_IN 1      /* this M belongs to a marked run */
#define SDIMB_GM_FIRST 2   /* first M of the run */
#define SDIMB_GM_LAST 4    /* last M of the run */
#define SDIMB_GM_FOLLOW 8  /* (with LAST) the stream has further ops behind the run */
#define SDIMB_OP_MASK 0xFF
#define SDIMB_OP_WARP_SHIFT 8
#define SDIMB_OP_INDEX_SHIFT 16
#ifndef SDIMB_SCHED_WARPS          /* compile-time knob of A/B builds; the shipped library uses 4 */
#define SDIMB_SCHED_WARPS 4
#endif

/* Record byte: low 7 bits = measured value, bit 7 = deterministic flag
 * (MeasurementResult.measurement_value / .deterministic, sdim/tableau/dataclasses.py:166-180).
 * d > 127: uint16 records, low 15 bits = value, bit 15 = deterministic flag. */
#define SDIMB_REC_DET 0x80u
#define SDIMB_REC_VALUE 0x7Fu
#define SDIMB_REC16_DET 0x8000u
#define SDIMB_REC16_VALUE 0x7FFFu

/* sdimb_run flags */
#define SDIMB_FRESH 0x1u           /* start every shot from |0...0> (ignore/skip the contents of `tableau`) */
#define SDIMB_WRITEBACK 0x2u       /* leave the final tableau of every shot in `tableau` */
#define SDIMB_FORCE_GLOBAL 0x4u    /* never stage the tableau in shared memory */
#define SDIMB_FORCE_RESIDENT 0x8u  /* require the shared-memory resident uint8-lane interpreter (else SDIMB_ETOOBIG) */
#define SDIMB_FORCE_LANES 0x10u    /* never use the bit-plane interpreter (d = 2, 3), keep uint8 lanes */
#define SDIMB_FORCE_PLANES 0x20u   /* require the bit-plane resident interpreter (d = 2, 3; else SDIMB_ETOOBIG); with
                                      SDIMB_FORCE_GLOBAL: the bit-plane interpreter on a global image (scratch) */
#define SDIMB_SCHEDULED 0x40u      /* `ops` is the output of sdimb_schedule: the bit-plane interpreter may run
                                      SDIMB_SCHED_WARPS warps per shot, one commuting layer at a time; the cluster
                                      interpreter deals each layer over its gate groups */
#define SDIMB_CLUSTER 0x80u        /* HBM store: run one shot per thread-block cluster whatever n and shots are
                                      (default: only for n > 512 with fewer shots than clusters fit on the GPU) */
#define SDIMB_NO_CLUSTER 0x100u    /* HBM store: never use the cluster interpreter */
#define SDIMB_NO_TILE 0x400u       /* d = 2, 3 with n <= 128: one shot per warp (interp_planes_kernel) instead of the tile
                                      interpreter with several shots per warp (tests, A/B timings) */
#define SDIMB_TIME_KERNELS 0x200u  /* measurement aid: a call that runs two kernels (interpreter + tail run) records CUDA
                                      events around each on `stream`; sdimb_kernel_times reads them (one set per
                                      process, not for concurrent callers) */

typedef struct SdimbLayout {
  int32_t n, d, np, lanes;   /* lanes = W = 2*np */
  int32_t order, phase_order;
  int64_t row_bytes;         /* 2*W: X row then Z row of one qudit */
  int64_t phase_offset;      /* n*row_bytes */
  int64_t shot_bytes;        /* phase_offset + W * elem_bytes */
  int32_t elem_bytes;        /* bytes per tableau entry: 1 (d <= 127), 2 (127 < d < 32768); the byte counts above include it */
  int32_t rec_bytes;         /* bytes per record / replay element, same rule */
} SdimbLayout;

typedef struct SdimbRunArgs {
  uint32_t struct_size;          /* = sizeof(SdimbRunArgs) */
  uint32_t flags;
  int32_t n, d;
  int64_t shots;                 /* shots simulated by this call */
  int64_t shot_offset;           /* global id of local shot 0 (Philox counter = shot_offset + local) */
  void* tableau;                 /* [device] shots*shot_bytes; may be NULL iff FRESH && !WRITEBACK && resident fits */
  const int32_t* ops;            /* [device] n_ops rows (opcode, a, b, slot) */
  int64_t n_ops;
  uint8_t* records;              /* [device] [shots][rec_stride], column k = chronological measurement k */
  int64_t n_meas;
  int64_t rec_stride;            /* >= n_meas */
  const uint8_t* replay_meas;    /* [device] nullable [shots][n_meas]: value used if measurement k is random */
  const uint8_t* replay_noise;   /* [device] nullable [shots][n_noise][2]: (a, b) of X^a Z^b for N1 event j */
  const uint32_t* noise_thresh24;/* [device] [n_noise] no-fire threshold, see sdim_b200/rng.py (Philox mode) */
  const uint8_t* noise_channel;  /* [device] [n_noise] 0 = 'd', 1 = 'f', 2 = 'p'  (sdim/program.py:486-497) */
  int64_t n_noise;
  uint64_t seed;
  void* stream;
  void* scratch;                 /* [device] nullable, sdimb_scratch_bytes() bytes: shot counter of the bit-plane */
  int64_t scratch_bytes;         /* interpreter (CTAs then claim shots dynamically instead of grid-striding)      */
  int64_t tail_run_len;          /* sdimb_tail_run(ops) of a SCHEDULED stream, or 0: the last tail_run_len ops are a marked
                                    run of M ops.  With scratch of sdimb_scratch_bytes_shots(.., shots) bytes and no
                                    WRITEBACK the global-image bit-plane interpreter hands that run to a second kernel
                                    (one warp per shot on a generator-major image, sdim_b200/csrc/planes_gm.cuh).
                                    Callers built against the struct without this field are accepted (struct_size). */
  const int32_t* gate_stream;    /* [device] nullable: output of sdimb_gate_stream for the ops in FRONT of the tail run */
  int64_t gate_stream_rows;      /* (ops [0, n_ops - tail_run_len) of the same scheduled stream) and its row count.  When
                                    the two-kernel path above applies and every measurement sits in the tail run, the
                                    front part then runs as pre-decoded per-warp streams (sdim_b200/csrc/planes_stream.cuh)
                                    instead of being interpreted.  Callers without these two fields are accepted. */
} SdimbRunArgs;

int sdimb_version(void);
const char* sdimb_strerror(int code);

/* Sizes and strides of the tableau store so the caller can allocate it. */
int sdimb_layout(int n, int d, SdimbLayout* out);

/* All shots <- |0...0>: stabilizers Z_q, destabilizers X_q
 * (sdim/tableau/dataclasses.py:34-39, sdim/tableau/tableau_prime.py:81-86). */
int sdimb_init(void* tableau, int n, int d, int64_t shots, void* stream);

/* Run the op stream over `shots` tableaus: the body of Program._simulate_tableau's shot loop
 * (sdim/program.py:308-351) with GATE_FUNCTIONS dispatch (sdim/program.py:13-32), the primitives of
 * sdim/tableau/tableau_optimized.py:5-118, measurement (sdim/tableau/tableau_prime.py:262-363),
 * the RESET correction (sdim/program.py:335-339) and N1 noise drawn with the distribution of
 * sdim/program.py:486-507. */
int sdimb_run(const SdimbRunArgs* args);

/* Widen one shot back to the reference's six int64 arrays, reference orientation [qudit][generator]
 * (x, z, dx, dz: n*n each; p, dp: n each).  All outputs [device]. */
int sdimb_export(const void* tableau, int n, int d, int64_t shot,
                 int64_t* x, int64_t* z, int64_t* p, int64_t* dx, int64_t* dz, int64_t* dp, void* stream);

/* Same job as sdimb_run with HOST buffers only: schedules the op stream (sdimb_schedule), copies it and the noise
 * tables in, simulates from |0...0>, copies the records out through pinned staging and synchronises.  Device
 * scratch, the staging buffer, the stream and the events live in a grow-only workspace of the CURRENT DEVICE that
 * later calls reuse (calls on one device are serialised internally; sdimb_release_workspace frees all of them).  If `records`
 * lies in pinned host memory (cudaHostAlloc / cudaHostRegister) the device writes it directly, without the staging copy.  This is the call a non-PyTorch host
 * (e.g. the reference itself through ctypes) makes; `elapsed_ms` (nullable) receives the device time of the
 * whole call measured with CUDA events. */
int sdimb_simulate_host(int n, int d, int64_t shots, int64_t shot_offset,
                        const int32_t* ops, int64_t n_ops,
                        uint8_t* records, int64_t n_meas,
                        const uint8_t* replay_meas, const uint8_t* replay_noise,
                        const uint32_t* noise_thresh24, const uint8_t* noise_channel, int64_t n_noise,
                        uint64_t seed, uint32_t flags, float* elapsed_ms);

int sdimb_release_workspace(void);

/* Pauli-frame sampler: the reference's default multi-shot mechanism, simulate_frame (sdim/program.py:45-165).
 * Given the records of ONE reference tableau shot (`reference[k]`, packed record bytes, N1 ignored in that shot as
 * in sdim/program.py:31,245-247), propagate one Pauli frame per extra shot through the op stream — x frame 0,
 * z frame uniform (:63-64); H: (x,z)<-(-z,x); H^-1: (z,-x); P: z+=x; CNOT: x[t]+=x[c], z[c]-=z[t]; CZ: z[t]+=x[c],
 * z[c]+=x[t]; SWAP; N1: (x,z)[q] += (a,b) (:82-120,159-162) — and record (reference + x[q]) mod d at M / M_X
 * (:121-144), then redraw z[q].  RESET records the PHYSICAL outcome (reference + x[q]) mod d, clears x[q] and
 * redraws z[q]; the reference records its own unchanged value there (:146-157), which is wrong under noise
 * (SURVEY Appendix B-5).  `records` rows are the extra shots (global ids shot_offset + local), same packed format,
 * deterministic flag copied from the reference shot as the reference does (:139).
 * `frames` is scratch, 2 * n * frames_pitch bytes with frames_pitch = shots rounded up to 128 [device].
 * Draws: Philox streams 1 (noise, as sdimb_run), 2 (initial z, slot = qudit), 3 (z redraw, slot = measurement),
 * or the nullable replay arrays replay_z0 [shots][n], replay_zm [shots][n_meas], replay_noise [shots][n_noise][2]. */
int sdimb_frames(int n, int d, int64_t shots, int64_t shot_offset, const int32_t* ops, int64_t n_ops,
                 const uint8_t* reference, uint8_t* records, int64_t n_meas, int64_t rec_stride, uint8_t* frames,
                 const uint8_t* replay_z0, const uint8_t* replay_zm, const uint8_t* replay_noise,
                 const uint32_t* noise_thresh24, const uint8_t* noise_channel, int64_t n_noise, uint64_t seed,
                 void* stream);

/* Host-side op-stream scheduler (no GPU work).  Reorders the gates between two collective ops (M, M_X, RESET)
 * into layers of ops on pairwise disjoint qudits (ASAP levels over row read/write dependencies), assigns the
 * ops of a layer round-robin to SDIMB_SCHED_WARPS warps and separates layers with SDIMB_OP_BARRIER.  Ops on
 * disjoint qudits commute exactly on the tableau (they touch disjoint rows and their phase increments add), so
 * records and final tableaus are unchanged; event slots travel with their ops.  `out` needs room for
 * 2*n_ops + 1 rows; *out_n receives the number written.  Replaces nothing in the reference (its loop is strictly
 * sequential, sdim/program.py:311-312); SURVEY 8f rank 3. */
int sdimb_schedule(int n, const int32_t* ops, int64_t n_ops, int32_t* out, int64_t out_cap, int64_t* out_n);

/* Scratch the bit-plane interpreter uses for (n, d, flags), 0 for the other interpreters.  Resident interpreter
 * (optional): a shot counter, CTAs then claim shots dynamically (shots differ in cost when noise fires) instead of
 * grid-striding.  Global-image interpreter (d = 2, 3 beyond the shared-memory limit; REQUIRED): the counter plus
 * one bit-plane image per CTA the current device keeps resident. */
int64_t sdimb_scratch_bytes(int n, int d, uint32_t flags);

/* Scratch for a call over `shots` shots whose stream ends in a marked measurement run (SdimbRunArgs.tail_run_len > 0):
 * at least sdimb_scratch_bytes(); for the global-image bit-plane interpreter with n <= 512 additionally one image per
 * shot plus the second kernel's slabs.  shots = 0 gives sdimb_scratch_bytes(). */
int64_t sdimb_scratch_bytes_shots(int n, int d, uint32_t flags, int64_t shots);

/* Length of the marked run of M ops (SDIMB_GM_*) that ends a scheduled HOST stream, 0 if it does not end in one. */
int64_t sdimb_tail_run(const int32_t* ops, int64_t n_ops);

/* Host-side compiler (no GPU work) of a gate-only stretch of a SCHEDULED stream — rows [0, n_ops) of sdimb_schedule's
 * output, holding unitary gates, N1 and SDIMB_OP_BARRIER only — into the pre-decoded per-warp streams of
 * gate_stream_kernel (sdim_b200/csrc/planes_stream.cuh): per layer, gates of one family on disjoint qudits are packed
 * 32 / Wb to a warp-wide op, N1 events move to a table evaluated once per shot.  d = 2, 3 and n <= 512 only
 * (SDIMB_EINVAL otherwise, or if the stretch holds a measurement, or more than 65 536 N1 ops: the caller then simply
 * passes no gate stream).  Call with out = NULL to learn the row count (*out_rows; rows are 4 x int32), then again
 * with room for it.  Replaces nothing in the reference, whose loop interprets one gate at a time
 * (sdim/program.py:311-312, GATE_FUNCTIONS sdim/program.py:13-32). */
int sdimb_gate_stream(int n, int d, const int32_t* sched_ops, int64_t n_ops, int32_t* out, int64_t out_cap_rows,
                      int64_t* out_rows);

/* Which interpreter sdimb_run would use for (n, d, flags): *kernel = 0 uint8 lanes on the HBM store (one CTA or
 * one thread-block cluster per shot, chosen per call from n and shots), 1 uint8 lanes resident in shared memory,
 * 2 bit-plane resident (d = 2, 3), 3 bit planes on a global image held in SdimbRunArgs.scratch (d = 2, 3 beyond
 * the shared-memory limit), 4 uint16 lanes on the HBM store (d > 127), 5 bit planes resident with several shots per
 * warp (d = 2, 3, n <= 128; sdim_b200/csrc/planes_tile.cuh); *needs_tableau = whether SdimbRunArgs.tableau must be a valid store for these flags. */
int sdimb_plan(int n, int d, uint32_t flags, int* kernel, int* needs_tableau);

/* Number of kernel launches issued by this library since load (for bench.py's gpu_launches). */
int64_t sdimb_launch_count(void);

/* Device time of the two kernels of the last sdimb_run that carried SDIMB_TIME_KERNELS and split its stream into
 * interpreter + tail run (waits for that call to finish); SDIMB_EINVAL if there was none.  bench.py's per-kernel
 * roofline uses it. */
int sdimb_kernel_times(float* front_ms, float* tail_ms);

/* Cluster size sdimb_run would give the cluster interpreter for (n, d, shots, flags) on the current device;
 * 0 = that call runs one CTA per shot (or another interpreter).  Needs a CUDA device (else 0). */
int sdimb_cluster_size(int n, int d, int64_t shots, uint32_t flags);

#ifdef __cplusplus
}
#endif
#endif /* SDIMB_H */
