// Shared device helpers: mod-d arithmetic, SWAR lanes, Philox4x32-10, kernel parameter block.
// Part of libsdimb (sdim_b200/csrc); included by sdimb.cu inside its anonymous namespace.
#pragma once

std::atomic<int64_t> g_launches{0};

constexpr int kMaxThreads = 256;
constexpr int kWideThreads = 1024;
constexpr uint32_t kNoPivot = 0xFFFFFFFFu;
constexpr int kSmemLimit = 227 * 1024;  // opt-in dynamic shared memory per CTA on sm_100

// ---------------------------------------------------------------------------------------------
// mod-d arithmetic with a precomputed reciprocal: x mod m == x - m * mulhi(x, ceil(2^32/m)) for x < 2^32/m
// ---------------------------------------------------------------------------------------------
struct Arith {
  uint32_t d, order, po, md, mo;
};

Arith make_arith(int d) {
  Arith a;
  a.d = (uint32_t)d;
  a.po = (d == 2) ? 2u : 1u;  // phase_order, sdim/tableau/dataclasses.py:98-106
  a.order = a.d * a.po;       // order,       sdim/tableau/dataclasses.py:88-96
  a.md = (uint32_t)((1ull << 32) / a.d) + 1u;
  a.mo = (uint32_t)((1ull << 32) / a.order) + 1u;
  return a;
}

__device__ __forceinline__ uint32_t mod_d(const Arith& A, uint32_t x) { return x - A.d * __umulhi(x, A.md); }
__device__ __forceinline__ uint32_t mod_o(const Arith& A, uint32_t x) { return x - A.order * __umulhi(x, A.mo); }
__device__ __forceinline__ uint32_t neg_d(const Arith& A, uint32_t x) { return x ? A.d - x : 0u; }
__device__ __forceinline__ uint32_t byte_of(uint32_t w, int k) { return (w >> (8 * k)) & 0xFFu; }

// SWAR arithmetic on four packed uint8 lanes, every lane reduced mod m (m <= 127, so a + b < 256 per lane).
// `rep` = m * 0x01010101, `bias` = (0x80 - m) * 0x01010101: adding the bias sets bit 7 of a lane iff lane >= m.
struct Swar {
  uint32_t m, rep, bias;
};
__device__ __forceinline__ Swar make_swar(uint32_t m) { return Swar{m, m * 0x01010101u, (0x80u - m) * 0x01010101u}; }
__device__ __forceinline__ uint32_t swar_reduce(const Swar& S, uint32_t s) {      // lanes in [0, 2m) -> [0, m)
  const uint32_t ge = ((s + S.bias) >> 7) & 0x01010101u;
  return s - ge * S.m;
}
__device__ __forceinline__ uint32_t swar_add(const Swar& S, uint32_t a, uint32_t b) { return swar_reduce(S, a + b); }
__device__ __forceinline__ uint32_t swar_neg(const Swar& S, uint32_t a) { return swar_reduce(S, S.rep - a); }
__device__ __forceinline__ uint32_t swar_sub(const Swar& S, uint32_t a, uint32_t b) { return swar_reduce(S, a + S.rep - b); }

// ---------------------------------------------------------------------------------------------
// Philox4x32-10, counter = (shot_lo, shot_hi, slot, stream), key = seed.  Host mirror: sdim_b200/rng.py
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                            uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

struct PlaneGeo {   // bit-plane interpreter: geometry and shared-memory offsets computed once on the host (planes::make_plane_geo)
  int np, Wb, RS, gpw, jstep, row_words, slab_words, acc_words;
  int off_ops, off_f, off_dotw, off_cnt, off_ar, off_br, off_xz, off_next;   // bytes from the phase accumulators' base
};

struct KParams {
  uint8_t* tab;
  const int4* ops;
  int64_t n_ops;
  uint8_t* records;
  int64_t rec_stride, n_meas;
  const uint8_t* replay_meas;
  const uint8_t* replay_noise;
  const uint32_t* thresh;
  const uint8_t* chan;
  int64_t n_noise;
  uint64_t seed;
  int64_t shots, shot_offset;
  int n, np, W;
  int64_t row_bytes, phase_off, shot_bytes;
  Arith A;
  uint32_t flags;
  int resident;
  int vec;                     // lane interpreter: lane words per thread (1, or 4 with 128-bit accesses)
  unsigned int* shot_counter;  // bit-plane kernel: next unclaimed shot (nullable: static grid-stride)
  int wpc;                     // cluster interpreter: lane words owned by each CTA of the cluster
  uint32_t* plane_slab;        // bit-plane interpreter on a global image: gridDim slabs of planes_row_bytes each
  PlaneGeo pg;                 // bit-plane interpreter: host-computed geometry (constant-bank operands instead of per-use arithmetic)
  // trailing measurement run on a generator-major image (planes_gm.cuh): the interpreter keeps the image of shot s at
  // plane_slab + s * img_stride_words (phase planes behind the rows) instead of one slab per CTA, run_tail_kernel
  // picks it up there
  int img_per_shot;
  int64_t img_stride_words;
  int64_t tail_start;          // run_tail_kernel: index of the first op of the run (ops [tail_start, n_ops) are M ops)
  uint32_t* gm_slab;           // run_tail_kernel: one B + QX slab per resident warp
  int64_t gm_slab_words;
  int gm_per_shot;             // gate_stream_kernel leaves B + QX + phase planes of shot s at gm_slab + s * gm_shot_stride_words
  int64_t gm_shot_stride_words;  // (it transposes its shared-memory image itself); run_tail_kernel works on them in place
  int64_t tile_stride_words;   // interp_tile_kernel: words between the shared-memory images of two tiles
  const int32_t* gate_stream;  // gate_stream_kernel: pre-decoded per-warp streams (sdimb_gate_stream, planes_stream.cuh)
};
