// libsdimb — B200 (sm_100a) stabilizer-tableau engine behind the C ABI of include/sdimb.h.
//
// One CTA owns one shot's tableau and interprets the whole op stream on it, so a shot never
// leaves the SM between gates.  The tableau lives in shared memory when one fits (resident mode)
// and in global memory (L2/HBM) otherwise; the device code is the same either way.
//
// Reference behaviour restated (file:line in events555/sdim):
//   primitives       sdim/tableau/tableau_optimized.py:5-118
//   composites       sdim/tableau/tableau_gates.py:27-137,229-261,298-329  (folded to closed forms)
//   measurement      sdim/tableau/tableau_prime.py:262-380
//   shot loop/RESET  sdim/program.py:308-351
//   noise            sdim/program.py:486-507
#include "sdimb.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace {

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
  unsigned int* shot_counter;  // bit-plane kernel: next unclaimed shot (nullable: static grid-stride)
};

// Shared scratch common to both modes (carved from dynamic shared memory after the resident tableau).
struct Scratch {
  uint32_t* dot;       // [W]    per-lane accumulator of Z[:,i] . x_p          (random branch)
  uint32_t* fw;        // [W/4]  packed factors f = -X[q,i] mod d, 4 lanes/word (random branch)
  uint32_t* red;       // [32]   cross-warp reduction scratch
  uint32_t* cnt;       // [4]    list lengths
  int4* ops;           // [32]   staged op batch (N1 rows carry the decoded event in .z)
  uint16_t* ar;        // [np]   active rows: qudits on which the pivot acts / active generators (det branch)
  uint16_t* br;        // [np]   rows outside the support whose destabilizer-p entry is stale
  uint16_t* aw;        // [W/4]  active words: lane quads holding a non-zero factor
  uint8_t* xs;         // [np]   pivot column X (random branch) / factors of the active generators (det branch)
  uint8_t* zs;         // [np]   pivot column Z
  uint8_t* inv;        // [128]  multiplicative inverses mod d
};

// Block reductions: warp reduce -> one word per warp in shared memory -> every warp reduces those words again with
// shuffles (one LDS + one redux instead of a 32-step loop per thread; it matters with 1024-thread CTAs).
__device__ __forceinline__ uint32_t block_sum(uint32_t v, uint32_t* red) {
  v = __reduce_add_sync(0xFFFFFFFFu, v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  return __reduce_add_sync(0xFFFFFFFFu, lane < (blockDim.x >> 5) ? red[lane] : 0u);
}

__device__ __forceinline__ uint32_t block_min(uint32_t v, uint32_t* red) {
  v = __reduce_min_sync(0xFFFFFFFFu, v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  return __reduce_min_sync(0xFFFFFFFFu, lane < (blockDim.x >> 5) ? red[lane] : kNoPivot);
}

// ---------------------------------------------------------------------------------------------
// |0...0>: stabilizers Z_q, destabilizers X_q (sdim/tableau/dataclasses.py:34-39, tableau_prime.py:81-86)
// ---------------------------------------------------------------------------------------------
__device__ void init_tableau(uint8_t* T, const KParams& p) {
  uint4* v = reinterpret_cast<uint4*>(T);
  const int64_t nvec = p.shot_bytes / 16;
  for (int64_t i = threadIdx.x; i < nvec; i += blockDim.x) v[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int q = threadIdx.x; q < p.n; q += blockDim.x) {
    uint8_t* row = T + (int64_t)q * p.row_bytes;
    row[p.W + q] = 1;      // Z[q][stab q]
    row[p.np + q] = 1;     // X[q][destab q]
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Unitary gates: every generator lane is independent, so a thread owns its 4-lane words for the whole gate
// sequence and consecutive gates need no barrier.  One gate on one word: `r` holds the X/Z words of row a (and
// row b for two-qudit gates), already loaded — possibly prefetched while the previous gate was computing; the
// new words are stored here, the new phase word is returned.  Words that cannot change are not rewritten.
// Closed forms: SURVEY Appendix A-1/A-2 (restating tableau_optimized.py:5-118, tableau_gates.py:27-261,298-329).
// ---------------------------------------------------------------------------------------------
struct Rows {
  uint32_t xa, za, xb, zb;
};

__device__ __forceinline__ bool is_two_qudit(int op) { return op >= SDIMB_OP_CNOT && op <= SDIMB_OP_SWAP; }
__device__ __forceinline__ bool is_unitary_like(int op) { return op < SDIMB_OP_M || op == SDIMB_OP_N1; }

__device__ __forceinline__ Rows load_rows(const uint8_t* T, const KParams& p, int op, int a, int b, int w) {
  const int wz = p.W / 4;
  const uint32_t* rowa = reinterpret_cast<const uint32_t*>(T + (int64_t)a * p.row_bytes);
  Rows r;
  r.xa = rowa[w];
  r.za = rowa[wz + w];
  r.xb = r.zb = 0u;
  if (is_two_qudit(op)) {
    const uint32_t* rowb = reinterpret_cast<const uint32_t*>(T + (int64_t)b * p.row_bytes);
    r.xb = rowb[w];
    r.zb = rowb[wz + w];
  }
  return r;
}

// pa, pb: Pauli exponents for X/X_INV/Z/Z_INV/N1 (phase += po*(pb*x - pa*z)), unused otherwise.
// Additions, subtractions and negations run on all four lanes of a word at once (Swar); only the lane-by-lane
// products of the phase terms are computed per byte.
__device__ __forceinline__ uint32_t gate_word(uint8_t* T, const KParams& p, int op, int a, int b, uint32_t pa,
                                              uint32_t pb, const Rows r, int w, uint32_t ph) {
  const Arith& A = p.A;
  const Swar Sd = make_swar(A.d), So = make_swar(A.order);
  const int wz = p.W / 4;
  uint32_t* rowa = reinterpret_cast<uint32_t*>(T + (int64_t)a * p.row_bytes);
  uint32_t* rowb = reinterpret_cast<uint32_t*>(T + (int64_t)b * p.row_bytes);
  switch (op) {
    case SDIMB_OP_H:
    case SDIMB_OP_H_INV: {
      if ((r.xa | r.za) == 0) return ph;
      // phase += po * new_x * new_z == -po * x * z        (tableau_optimized.py:17-30,45-58)
      uint32_t prod = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) prod |= (A.po * mod_d(A, byte_of(r.xa, k) * byte_of(r.za, k))) << (8 * k);
      if (op == SDIMB_OP_H) { rowa[w] = swar_neg(Sd, r.za); rowa[wz + w] = r.xa; }     // (x,z) <- (-z, x)
      else { rowa[w] = r.za; rowa[wz + w] = swar_neg(Sd, r.xa); }                      // (x,z) <- (z, -x)
      return swar_sub(So, ph, prod);
    }
    case SDIMB_OP_P:
    case SDIMB_OP_P_INV: {
      if (r.xa == 0) return ph;
      // even d: phase +-= x^2 (mod 2d); odd d: phase +-= x(x-1)/2 (mod d)   (tableau_optimized.py:62-96)
      uint32_t inc = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t xb = byte_of(r.xa, k);
        inc |= ((A.po == 2) ? mod_o(A, xb * xb) : mod_d(A, (xb * (xb - 1u)) >> 1)) << (8 * k);
      }
      if (op == SDIMB_OP_P) { rowa[wz + w] = swar_add(Sd, r.za, r.xa); return swar_add(So, ph, inc); }
      rowa[wz + w] = swar_sub(Sd, r.za, r.xa);
      return swar_sub(So, ph, inc);
    }
    case SDIMB_OP_X: case SDIMB_OP_X_INV: case SDIMB_OP_Z: case SDIMB_OP_Z_INV: case SDIMB_OP_N1: {
      // conjugation by X^pa Z^pb: phase += po * (pb*x - pa*z)  (tableau_gates.py:27-137, program.py:335-339)
      const uint32_t na = pa ? A.d - pa : 0u;
      const uint32_t x = pb ? r.xa : 0u, z = na ? r.za : 0u;
      if ((x | z) == 0) return ph;
      const bool unit_x = pb == 0u || pb == 1u || pb == A.d - 1u, unit_z = pa == 0u || pa == 1u || pa == A.d - 1u;
      if (unit_x && unit_z) {                                 // exponents +-1: no products, all four lanes at once
        const uint32_t sx = A.po == 2 ? x << 1 : x, sz = A.po == 2 ? z << 1 : z;
        uint32_t q = ph;
        if (pb) q = (pb == 1u) ? swar_add(So, q, sx) : swar_sub(So, q, sx);
        if (pa) q = (pa == 1u) ? swar_sub(So, q, sz) : swar_add(So, q, sz);
        return q;
      }
      uint32_t t = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) t |= mod_d(A, pb * byte_of(x, k) + na * byte_of(z, k)) << (8 * k);
      return swar_add(So, ph, A.po == 2 ? t << 1 : t);
    }
    case SDIMB_OP_CNOT: {                                     // x[t] += x[c];  z[c] -= z[t]   (tableau_optimized.py:99-107)
      if ((r.xa | r.zb) == 0) return ph;
      rowb[w] = swar_add(Sd, r.xb, r.xa);
      rowa[wz + w] = swar_sub(Sd, r.za, r.zb);
      return ph;
    }
    case SDIMB_OP_CNOT_INV: {                                 // x[t] -= x[c];  z[c] += z[t]   (tableau_optimized.py:110-118)
      if ((r.xa | r.zb) == 0) return ph;
      rowb[w] = swar_sub(Sd, r.xb, r.xa);
      rowa[wz + w] = swar_add(Sd, r.za, r.zb);
      return ph;
    }
    case SDIMB_OP_CZ:
    case SDIMB_OP_CZ_INV: {
      if ((r.xa | r.xb) == 0) return ph;
      // CZ = H^-1(t) CNOT(c,t) H(t) folded: z[a] +-= x[b]; z[b] +-= x[a]; phase +-= po*x[a]*x[b]
      uint32_t prod = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) prod |= (A.po * mod_d(A, byte_of(r.xa, k) * byte_of(r.xb, k))) << (8 * k);
      if (op == SDIMB_OP_CZ) {
        rowa[wz + w] = swar_add(Sd, r.za, r.xb); rowb[wz + w] = swar_add(Sd, r.zb, r.xa);
        return swar_add(So, ph, prod);
      }
      rowa[wz + w] = swar_sub(Sd, r.za, r.xb); rowb[wz + w] = swar_sub(Sd, r.zb, r.xa);
      return swar_sub(So, ph, prod);
    }
    case SDIMB_OP_SWAP:
      rowa[w] = r.xb; rowa[wz + w] = r.zb;
      rowb[w] = r.xa; rowb[wz + w] = r.za;
      return ph;
    default:
      return ph;
  }
}

// Pauli exponents (a, b) of an op of the phase-only family
__device__ __forceinline__ void pauli_exponents(const KParams& p, const int4& op, uint32_t& pa, uint32_t& pb) {
  pa = pb = 0u;
  switch (op.x) {
    case SDIMB_OP_X: pa = 1u; break;
    case SDIMB_OP_X_INV: pa = p.A.d - 1u; break;
    case SDIMB_OP_Z: pb = 1u; break;
    case SDIMB_OP_Z_INV: pb = p.A.d - 1u; break;
    case SDIMB_OP_N1: pa = (uint32_t)op.z & 0xFFu; pb = (uint32_t)op.z >> 8; break;
    default: break;
  }
}

// A whole gate when rows span several words per thread (n > 4 * blockDim): plain loop, phases in memory.
__device__ __forceinline__ void gate_rows(uint8_t* T, const KParams& p, const int4& op, uint32_t pa, uint32_t pb) {
  uint32_t* P = reinterpret_cast<uint32_t*>(T + p.phase_off);
  for (int w = threadIdx.x; w < p.W / 4; w += blockDim.x) {
    const Rows r = load_rows(T, p, op.x, op.y, op.z, w);
    const uint32_t ph = P[w];
    const uint32_t nph = gate_word(T, p, op.x, op.y, op.z, pa, pb, r, w, ph);
    if (nph != ph) P[w] = nph;
  }
}

// ---------------------------------------------------------------------------------------------
// Noise: N1 event j of this shot -> (a | b << 8), 0 if it does not fire.  Replayed, or Philox with the
// distribution of sdim/program.py:486-507.  Evaluated by the thread that fetched the op, so events that do
// not fire never reach the dispatch loop.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t noise_event(const KParams& p, int64_t j, int64_t shot_local) {
  const uint32_t d = p.A.d;
  uint32_t a = 0, b = 0;
  if (p.replay_noise) {
    const uint8_t* src = p.replay_noise + (shot_local * p.n_noise + j) * 2;
    a = src[0]; b = src[1];
  } else {
    const uint64_t gshot = (uint64_t)(p.shot_offset + shot_local);
    const uint4 r = philox4x32((uint32_t)gshot, (uint32_t)(gshot >> 32), (uint32_t)j, 1u, (uint32_t)p.seed,
                               (uint32_t)(p.seed >> 32));
    if ((r.x >> 8) >= __ldg(p.thresh + j)) {
      const uint32_t ch = __ldg(p.chan + j);
      if (ch == 0) {                                        // 'd': r ~ U{1..d^2-1}, a = r % d, b = r // d
        const uint32_t v = 1u + __umulhi(r.y, d * d - 1u);
        a = v % d; b = v / d;
      } else {                                              // 'f': X^e, 'p': Z^e, e ~ U{1..d-1}
        const uint32_t e = 1u + __umulhi(r.y, d - 1u);
        if (ch == 1) a = e; else b = e;
      }
    }
  }
  return a | (b << 8);
}

// ---------------------------------------------------------------------------------------------
// Measurement of qudit q in the Z basis (tableau_prime.py:262-363).  Returns the outcome to every thread.
//
// The reference skips generators whose factor is zero (tableau_prime.py:308,315,351) and its column
// updates are no-ops on qudits where the pivot is the identity.  Both sparsities are exploited here:
// the update runs over (active row) x (active lane-quad) pairs spread across the whole CTA, with kBatch
// independent loads in flight per thread, so a sparse measurement costs a handful of memory round trips
// and a dense one streams the tableau with full memory-level parallelism.
// ---------------------------------------------------------------------------------------------
constexpr int kBatch = 4;
constexpr int kWalk = 4;     // rows per thread whose column loads are issued together

__device__ uint32_t measure(uint8_t* T, const KParams& p, Scratch& S, int q, int64_t slot, int64_t shot_local,
                            uint32_t draw) {
  const Arith& A = p.A;
  const int n = p.n, W = p.W, npad = p.np, nt = blockDim.x, tid = threadIdx.x;
  const int wz = W / 4;
  uint8_t* rowq = T + (int64_t)q * p.row_bytes;
  uint8_t* P8 = T + p.phase_off;
  if (tid < 3) S.cnt[tid] = 0;   // cnt[3] holds the live-op mask of the current batch
  __syncthreads();   // gate writes of other threads' lanes become visible; counters reset

  // -- pivot: FIRST stabilizer with an X component on q (tableau_prime.py:273-283) ------------------
  uint32_t best = kNoPivot;
  {
    const uint32_t* xq = reinterpret_cast<const uint32_t*>(rowq);
    for (int w = tid; w < npad / 4; w += nt) {
      const uint32_t x = xq[w];
      if (x) { best = 4u * w + ((__ffs(x) - 1) >> 3); break; }
    }
  }
  const uint32_t piv = block_min(best, S.red);

  uint32_t outcome, rec;
  if (piv != kNoPivot) {
    // -- random branch (tableau_prime.py:294-334), with exponentiate (:365-380) folded into the gather ----
    const uint32_t v = rowq[piv];
    const uint32_t e = S.inv[v];
    const uint32_t ps_old = P8[piv];
    uint32_t sd_raw = 0;
    // One walk down the pivot column AND the destabilizer-p column (kWalk rows' loads in flight per thread):
    // xs/zs, the support list `ar`, and the list `br` of rows outside the support whose destabilizer-p entry is
    // non-zero and must be cleared when the destabilizer is overwritten with the pivot.
    for (int base = tid; base < n; base += nt * kWalk) {
      uint32_t xr[kWalk], zr[kWalk], od[kWalk];
#pragma unroll
      for (int u = 0; u < kWalk; ++u) {
        const int r = base + u * nt;
        xr[u] = zr[u] = od[u] = 0;
        if (r < n) {
          const uint8_t* row = T + (int64_t)r * p.row_bytes;
          xr[u] = row[piv]; zr[u] = row[W + piv];
          od[u] = (uint32_t)row[npad + piv] | row[W + npad + piv];
        }
      }
#pragma unroll
      for (int u = 0; u < kWalk; ++u) {
        const int r = base + u * nt;
        if (r >= n) break;
        S.xs[r] = (uint8_t)mod_d(A, xr[u] * e);
        S.zs[r] = (uint8_t)mod_d(A, zr[u] * e);
        sd_raw += mod_d(A, xr[u] * zr[u]);
        if (xr[u] | zr[u]) S.ar[atomicAdd(&S.cnt[0], 1u)] = (uint16_t)r;
        else if (od[u]) S.br[atomicAdd(&S.cnt[2], 1u)] = (uint16_t)r;
      }
    }
    for (int w = tid; w < wz; w += nt) {                      // factors f = -X[q,i] mod d, active-word list
      const uint32_t xq_w = reinterpret_cast<const uint32_t*>(rowq)[w];
      uint32_t fw = 0;
      if (xq_w) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (4u * w + k != piv) fw |= neg_d(A, byte_of(xq_w, k)) << (8 * k);   // the pivot itself is skipped
      }
      S.fw[w] = fw;
      if (fw) {
        S.aw[atomicAdd(&S.cnt[1], 1u)] = (uint16_t)w;
        *reinterpret_cast<uint4*>(S.dot + 4 * w) = make_uint4(0, 0, 0, 0);
      }
    }
    sd_raw = mod_d(A, block_sum(sd_raw, S.red));              // barrier: publishes xs/zs/ar/fw/aw/dot/cnt
    const uint32_t ps = mod_o(A, ps_old * e + A.po * mod_d(A, sd_raw * mod_d(A, (e * (e - 1u)) >> 1)));
    const uint32_t sd = mod_d(A, mod_d(A, sd_raw * e) * e);   // x_p . z_p after exponentiation
    const int nr_a = (int)S.cnt[0], nw_a = (int)S.cnt[1];

    // col_i += f_i * col_p over all (active row, active word) pairs; pair id = ri * nw_a + wi
    if (nw_a > 0) {
      const int npairs = nr_a * nw_a;
      const int dw = nt % nw_a, dr = nt / nw_a;
      int wi = tid % nw_a, ri = tid / nw_a;
      int cur_w = -1;
      uint32_t dot0 = 0, dot1 = 0, dot2 = 0, dot3 = 0;
      for (int base = tid; base < npairs; base += nt * kBatch) {
        uint32_t xw[kBatch], zw[kBatch];
        int rr[kBatch], ww[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          rr[u] = -1;
          if (base + u * nt < npairs) {
            rr[u] = S.ar[ri];
            ww[u] = S.aw[wi];
            const uint32_t* src = reinterpret_cast<const uint32_t*>(T + (int64_t)rr[u] * p.row_bytes) + ww[u];
            xw[u] = src[0];
            zw[u] = src[wz];
            wi += dw; ri += dr;
            if (wi >= nw_a) { wi -= nw_a; ++ri; }
          }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          if (rr[u] < 0) continue;
          if (ww[u] != cur_w) {
            if (cur_w >= 0) {
              atomicAdd(&S.dot[4 * cur_w + 0], dot0); atomicAdd(&S.dot[4 * cur_w + 1], dot1);
              atomicAdd(&S.dot[4 * cur_w + 2], dot2); atomicAdd(&S.dot[4 * cur_w + 3], dot3);
            }
            cur_w = ww[u]; dot0 = dot1 = dot2 = dot3 = 0;
          }
          const uint32_t s = S.xs[rr[u]], t = S.zs[rr[u]], fw = S.fw[ww[u]];
          const uint32_t z0 = byte_of(zw[u], 0), z1 = byte_of(zw[u], 1), z2 = byte_of(zw[u], 2), z3 = byte_of(zw[u], 3);
          dot0 = mod_d(A, dot0 + z0 * s); dot1 = mod_d(A, dot1 + z1 * s);    // Z[:,i] . x_p (old Z)
          dot2 = mod_d(A, dot2 + z2 * s); dot3 = mod_d(A, dot3 + z3 * s);
          uint32_t nx = 0, nz = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t fk = byte_of(fw, k);
            nx |= mod_d(A, byte_of(xw[u], k) + fk * s) << (8 * k);
            nz |= mod_d(A, byte_of(zw[u], k) + fk * t) << (8 * k);
          }
          uint32_t* dst = reinterpret_cast<uint32_t*>(T + (int64_t)rr[u] * p.row_bytes) + ww[u];
          dst[0] = nx;
          dst[wz] = nz;
        }
      }
      if (cur_w >= 0) {
        atomicAdd(&S.dot[4 * cur_w + 0], dot0); atomicAdd(&S.dot[4 * cur_w + 1], dot1);
        atomicAdd(&S.dot[4 * cur_w + 2], dot2); atomicAdd(&S.dot[4 * cur_w + 3], dot3);
      }
    }
    __syncthreads();
    // phase_i += f_i * phase_p + po * (f_i * (Z_i . x_p) + (x_p . z_p) * f_i(f_i-1)/2 * po)     (:310-312,317-319)
    for (int i = tid; i < nw_a; i += nt) {
      const int w = S.aw[i];
      const uint32_t fw = S.fw[w];
      uint32_t* Pw = reinterpret_cast<uint32_t*>(P8) + w;
      const uint32_t ph = *Pw;
      uint32_t nph = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t fk = byte_of(fw, k);
        const uint32_t g = mod_d(A, (fk * (fk - 1u)) >> 1);
        const uint32_t cp = mod_d(A, mod_d(A, S.dot[4 * w + k]) * fk + sd * g * A.po);
        nph |= mod_o(A, byte_of(ph, k) + fk * ps + A.po * cp) << (8 * k);
      }
      *Pw = nph;
    }
    __syncthreads();
    // destabilizer p <- old pivot, stabilizer p <- Z_q with phase -m*po (tableau_prime.py:323-333).
    // Column accesses cost one DRAM sector per byte, so only entries that change are written: both lanes are
    // rewritten on the pivot's support (list ar), stale destabilizer entries elsewhere (list br) are cleared.
    const int nr_b = (int)S.cnt[2];
    for (int i = tid; i < nr_a; i += nt) {
      const int r = S.ar[i];
      uint8_t* row = T + (int64_t)r * p.row_bytes;
      row[piv] = 0;
      row[W + piv] = (r == q) ? 1 : 0;
      row[npad + piv] = S.xs[r];
      row[W + npad + piv] = S.zs[r];
    }
    for (int i = tid; i < nr_b; i += nt) {
      uint8_t* row = T + (int64_t)S.br[i] * p.row_bytes;
      row[npad + piv] = 0;
      row[W + npad + piv] = 0;
    }
    outcome = draw;
    if (tid == 0) {
      P8[npad + piv] = (uint8_t)ps;
      P8[piv] = (uint8_t)mod_o(A, A.order - outcome * A.po);
    }
    rec = outcome;
  } else {
    // -- deterministic branch (tableau_prime.py:336-363): ordered accumulation over generators ----------
    // Ordered compaction of the generators with a non-zero factor f_i = destab X[q,i] (order matters: the
    // cross term uses the running ancilla, :354-357).
    uint32_t a1 = 0;
    int total = 0;
    __syncthreads();   // every warp has finished reading S.red in block_min before it is reused below
    for (int base = 0; base < n; base += nt) {
      const int i = base + tid;
      const uint32_t f = (i < n) ? rowq[npad + i] : 0u;
      const uint32_t mask = __ballot_sync(0xFFFFFFFFu, f != 0);
      if ((tid & 31) == 0) S.red[tid >> 5] = __popc(mask);
      __syncthreads();
      // exclusive scan of the per-warp counts, done by every warp with shuffles
      const int lane_id = tid & 31;
      const int mine_cnt = lane_id < (nt >> 5) ? (int)S.red[lane_id] : 0;
      int incl = mine_cnt;
#pragma unroll
      for (int d2 = 1; d2 < 32; d2 <<= 1) {
        const int o = __shfl_up_sync(0xFFFFFFFFu, incl, d2);
        if (lane_id >= d2) incl += o;
      }
      const int off = total + __shfl_sync(0xFFFFFFFFu, incl - mine_cnt, tid >> 5);
      const int all = total + __shfl_sync(0xFFFFFFFFu, incl, 31);
      if (f) {
        const int pos = off + __popc(mask & ((1u << (tid & 31)) - 1u));
        S.ar[pos] = (uint16_t)i;
        S.xs[pos] = (uint8_t)f;
        a1 += f * P8[i];
      }
      total = all;
      __syncthreads();
    }
    a1 = mod_o(A, block_sum(mod_o(A, a1), S.red));            // sum_i f_i * phase_i; publishes the lists
    uint32_t part = 0;
    for (int r = tid; r < n; r += nt) {
      const uint8_t* xr = T + (int64_t)r * p.row_bytes;
      const uint8_t* zr = xr + W;
      uint32_t az = 0, cross = 0, sdg = 0;
      for (int base = 0; base < total; base += kBatch) {
        uint32_t xi[kBatch], zi[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u)
          if (base + u < total) { const int g = S.ar[base + u]; xi[u] = xr[g]; zi[u] = zr[g]; }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          if (base + u >= total) break;
          const uint32_t f = S.xs[base + u];
          cross = mod_d(A, cross + mod_d(A, f * xi[u]) * az);  // ancilla_z . (f * x_i), running ancilla
          az = mod_d(A, az + f * zi[u]);
          sdg = mod_d(A, sdg + mod_d(A, xi[u] * zi[u]) * mod_d(A, (f * (f - 1u)) >> 1));
        }
      }
      part += mod_d(A, cross + A.po * sdg);
    }
    part = mod_d(A, block_sum(part, S.red));
    const uint32_t ap = mod_o(A, a1 + A.po * part);
    // (-ap // po) % d with Python floor semantics (tableau_prime.py:362)
    outcome = (A.po == 1) ? neg_d(A, ap) : (((ap + 1u) >> 1) & 1u);
    rec = outcome | SDIMB_REC_DET;
  }
  if (tid == 0) p.records[shot_local * p.rec_stride + slot] = (uint8_t)rec;
  __syncthreads();
  return outcome;
}

#include "planes.cuh"

// ---------------------------------------------------------------------------------------------
// The interpreter: one CTA per shot, grid-stride over shots.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void interp_body(const KParams& p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int64_t tab_smem = p.resident ? p.shot_bytes : 0;
  Scratch S;
  S.dot = reinterpret_cast<uint32_t*>(smem + tab_smem);
  S.fw = S.dot + p.W;
  S.red = S.fw + p.W / 4;
  S.cnt = S.red + 32;
  S.ops = reinterpret_cast<int4*>(S.cnt + 4);
  S.ar = reinterpret_cast<uint16_t*>(S.ops + 32);
  S.br = S.ar + p.np;
  S.aw = S.br + p.np;
  S.xs = reinterpret_cast<uint8_t*>(S.aw + p.W / 4);
  S.zs = S.xs + p.np;
  S.inv = S.zs + p.np;
  const Arith& A = p.A;

  for (uint32_t v = threadIdx.x; v < A.d; v += blockDim.x) {   // inverse table; inv[0] unused
    uint32_t e = 0;
    for (uint32_t c = 1; c < A.d; ++c)
      if (mod_d(A, v * c) == 1u) { e = c; break; }
    S.inv[v] = (uint8_t)e;
  }
  __syncthreads();

  for (int64_t shot = blockIdx.x; shot < p.shots; shot += gridDim.x) {
    uint8_t* G = p.tab ? p.tab + shot * p.shot_bytes : nullptr;
    uint8_t* T = p.resident ? smem : G;
    if (p.flags & SDIMB_FRESH) {
      init_tableau(T, p);
    } else if (p.resident) {
      const uint4* src = reinterpret_cast<const uint4*>(G);
      uint4* dst = reinterpret_cast<uint4*>(T);
      for (int64_t i = threadIdx.x; i < p.shot_bytes / 16; i += blockDim.x) dst[i] = src[i];
      __syncthreads();
    }

    // When a row is at most one word per thread, the thread's phase word stays in a register between
    // measurements (gates never read another lane's phase).
    const int w0 = threadIdx.x;
    const bool one_word = p.W / 4 <= (int)blockDim.x;
    const bool own_word = one_word && w0 < p.W / 4;
    uint32_t* Pw = reinterpret_cast<uint32_t*>(T + p.phase_off) + w0;
    uint32_t pw = own_word ? *Pw : 0u;

    for (int64_t i0 = 0; i0 < p.n_ops; i0 += 32) {
      // warp 0 fetches 32 ops (one per lane) and resolves their N1 events; only live ops are dispatched
      __syncthreads();
      if (threadIdx.x < 32) {
        int4 mine = make_int4(SDIMB_OP_I, 0, 0, 0);
        if (i0 + threadIdx.x < p.n_ops) mine = __ldg(p.ops + i0 + threadIdx.x);
        mine.x &= SDIMB_OP_MASK;             // a scheduled stream carries a warp id here; lanes need no schedule
        bool live = mine.x != SDIMB_OP_I && mine.x != SDIMB_OP_BARRIER;
        if (mine.x == SDIMB_OP_N1) {
          mine.z = (int)noise_event(p, mine.w, shot);
          live = mine.z != 0;
        } else if (mine.x >= SDIMB_OP_M && mine.x <= SDIMB_OP_RESET) {
          // outcome this measurement takes if it is random: replayed draw or Philox (reference: random.choice,
          // tableau_prime.py:332), resolved by the fetching lane so that it is off the measurement's critical path
          if (p.replay_meas) {
            mine.z = p.replay_meas[shot * p.n_meas + mine.w];
          } else {
            const uint64_t gshot = (uint64_t)(p.shot_offset + shot);
            const uint4 r = philox4x32((uint32_t)gshot, (uint32_t)(gshot >> 32), (uint32_t)mine.w, 0u,
                                       (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
            mine.z = (int)__umulhi(r.x, A.d);
          }
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, live);
        S.ops[threadIdx.x] = mine;
        if (threadIdx.x == 0) S.cnt[3] = m;
      }
      __syncthreads();
      uint32_t todo = S.cnt[3];
#pragma unroll 1
      while (todo) {
        const int k = __ffs(todo) - 1;
        todo &= todo - 1;
        const int4 op = S.ops[k];
        if (is_unitary_like(op.x)) {
          uint32_t pa, pb;
          pauli_exponents(p, op, pa, pb);
          if (!one_word) {
            gate_rows(T, p, op, pa, pb);
          } else if (own_word) {
            pw = gate_word(T, p, op.x, op.y, op.z, pa, pb, load_rows(T, p, op.x, op.y, op.z, w0), w0, pw);
          }
          continue;
        }
        // collective ops: M, M_X, RESET
        if (op.x == SDIMB_OP_M_X) {                          // tableau_gates.py:292-296: H^-1 then measure
          const int4 h = make_int4(SDIMB_OP_H_INV, op.y, -1, -1);
          if (!one_word) gate_rows(T, p, h, 0u, 0u);
          else if (own_word) pw = gate_word(T, p, h.x, h.y, h.z, 0u, 0u, load_rows(T, p, h.x, h.y, h.z, w0), w0, pw);
        }
        if (own_word) *Pw = pw;                              // measurement reads and writes phases in memory
        const uint32_t m = measure(T, p, S, op.y, op.w, shot, (uint32_t)op.z);
        if (own_word) pw = *Pw;
        if (op.x == SDIMB_OP_RESET && m) {                   // program.py:335-339: X applied (-m) mod d times
          const int4 x = make_int4(SDIMB_OP_N1, op.y, (int)(A.d - m), -1);
          if (!one_word) gate_rows(T, p, x, A.d - m, 0u);
          else if (own_word) pw = gate_word(T, p, x.x, x.y, -1, A.d - m, 0u, load_rows(T, p, x.x, x.y, -1, w0), w0, pw);
        }
      }
    }
    if (own_word) *Pw = pw;
    __syncthreads();
    if (p.resident && (p.flags & SDIMB_WRITEBACK)) {
      const uint4* src = reinterpret_cast<const uint4*>(T);
      uint4* dst = reinterpret_cast<uint4*>(G);
      for (int64_t i = threadIdx.x; i < p.shot_bytes / 16; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
  }
}

// Two launch shapes of the same body: up to 256 threads (many shots in flight, 12 CTAs of 128 threads per SM) and
// up to 1024 threads for tableaus whose rows span more than 256 lane words (n > 512: few, large shots — config 5).
__global__ void __launch_bounds__(kMaxThreads, 6) interp_kernel(const __grid_constant__ KParams p) { interp_body(p); }
__global__ void __launch_bounds__(kWideThreads, 1) interp_kernel_wide(const __grid_constant__ KParams p) { interp_body(p); }

__global__ void init_kernel(uint8_t* tab, int n, int np, int W, int64_t row_bytes, int64_t shot_bytes, int64_t shots) {
  for (int64_t shot = blockIdx.x; shot < shots; shot += gridDim.x) {
    uint8_t* T = tab + shot * shot_bytes;
    uint4* v = reinterpret_cast<uint4*>(T);
    for (int64_t i = threadIdx.x; i < shot_bytes / 16; i += blockDim.x) v[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
      T[(int64_t)q * row_bytes + W + q] = 1;
      T[(int64_t)q * row_bytes + np + q] = 1;
    }
    __syncthreads();
  }
}

__global__ void export_kernel(const uint8_t* T, int n, int np, int W, int64_t row_bytes, int64_t phase_off,
                              int64_t* x, int64_t* z, int64_t* ph, int64_t* dx, int64_t* dz, int64_t* dph) {
  const int64_t total = (int64_t)n * n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i / n), g = (int)(i % n);
    const uint8_t* row = T + (int64_t)q * row_bytes;
    x[i] = row[g];
    z[i] = row[W + g];
    dx[i] = row[np + g];
    dz[i] = row[W + np + g];
    if (q == 0) {
      ph[g] = T[phase_off + g];
      dph[g] = T[phase_off + np + g];
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Pauli-frame sampler (simulate_frame, sdim/program.py:45-165): one thread per extra shot, frames laid out
// [x|z][qudit][shot] so that every access of a warp is one contiguous 32-byte run.
// ---------------------------------------------------------------------------------------------
struct FParams {
  const int4* ops;
  int64_t n_ops;
  const uint8_t* reference;
  uint8_t* records;
  int64_t n_meas, rec_stride;
  uint8_t* frames;
  int64_t pitch;
  const uint8_t *replay_z0, *replay_zm, *replay_noise;
  const uint32_t* thresh;
  const uint8_t* chan;
  int64_t n_noise;
  uint64_t seed;
  int64_t shots, shot_offset;
  int n;
  Arith A;
};

__global__ void __launch_bounds__(128) frame_kernel(const __grid_constant__ FParams p) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= p.shots) return;
  const Arith& A = p.A;
  const uint32_t d = A.d;
  uint8_t* X = p.frames + s;
  uint8_t* Z = p.frames + (int64_t)p.n * p.pitch + s;
  const uint64_t gshot = (uint64_t)(p.shot_offset + s);
  const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
  for (int q = 0; q < p.n; ++q) {                       // x_frame = 0, z_frame uniform   (program.py:63-64)
    X[(int64_t)q * p.pitch] = 0;
    Z[(int64_t)q * p.pitch] = p.replay_z0 ? p.replay_z0[s * p.n + q]
        : (uint8_t)__umulhi(philox4x32((uint32_t)gshot, (uint32_t)(gshot >> 32), (uint32_t)q, 2u, k0, k1).x, d);
  }
  for (int64_t i = 0; i < p.n_ops; ++i) {
    const int4 op = __ldg(p.ops + i);
    const int code = op.x & SDIMB_OP_MASK;
    uint8_t* xa = X + (int64_t)op.y * p.pitch;
    uint8_t* za = Z + (int64_t)op.y * p.pitch;
    uint8_t* xb = X + (int64_t)op.z * p.pitch;
    uint8_t* zb = Z + (int64_t)op.z * p.pitch;
    switch (code) {
      case SDIMB_OP_H: { const uint32_t x = *xa, z = *za; *xa = (uint8_t)neg_d(A, z); *za = (uint8_t)x; break; }
      case SDIMB_OP_H_INV: { const uint32_t x = *xa, z = *za; *xa = (uint8_t)z; *za = (uint8_t)neg_d(A, x); break; }
      case SDIMB_OP_P: *za = (uint8_t)mod_d(A, (uint32_t)*za + *xa); break;
      case SDIMB_OP_P_INV: *za = (uint8_t)mod_d(A, (uint32_t)*za + d - *xa); break;
      case SDIMB_OP_CNOT: *xb = (uint8_t)mod_d(A, (uint32_t)*xb + *xa); *za = (uint8_t)mod_d(A, (uint32_t)*za + d - *zb); break;
      case SDIMB_OP_CNOT_INV: *xb = (uint8_t)mod_d(A, (uint32_t)*xb + d - *xa); *za = (uint8_t)mod_d(A, (uint32_t)*za + *zb); break;
      case SDIMB_OP_CZ: *zb = (uint8_t)mod_d(A, (uint32_t)*zb + *xa); *za = (uint8_t)mod_d(A, (uint32_t)*za + *xb); break;
      case SDIMB_OP_CZ_INV: *zb = (uint8_t)mod_d(A, (uint32_t)*zb + d - *xa); *za = (uint8_t)mod_d(A, (uint32_t)*za + d - *xb); break;
      case SDIMB_OP_SWAP: { const uint8_t x = *xa, z = *za; *xa = *xb; *za = *zb; *xb = x; *zb = z; break; }
      case SDIMB_OP_M:
      case SDIMB_OP_M_X:
      case SDIMB_OP_RESET: {
        if (code == SDIMB_OP_M_X) { const uint32_t x = *xa, z = *za; *xa = (uint8_t)z; *za = (uint8_t)neg_d(A, x); }
        const uint32_t ref = p.reference[op.w];
        const uint32_t val = mod_d(A, (ref & SDIMB_REC_VALUE) + *xa);       // program.py:135
        p.records[s * p.rec_stride + op.w] = (uint8_t)(val | (ref & SDIMB_REC_DET));
        if (code == SDIMB_OP_RESET) *xa = 0;                                 // program.py:155
        *za = p.replay_zm ? p.replay_zm[s * p.n_meas + op.w]                 // program.py:144,156
            : (uint8_t)__umulhi(philox4x32((uint32_t)gshot, (uint32_t)(gshot >> 32), (uint32_t)op.w, 3u, k0, k1).x, d);
        break;
      }
      case SDIMB_OP_N1: {
        uint32_t a = 0, b = 0;
        const int64_t j = op.w;
        if (p.replay_noise) {
          a = p.replay_noise[(s * p.n_noise + j) * 2]; b = p.replay_noise[(s * p.n_noise + j) * 2 + 1];
        } else {
          const uint4 r = philox4x32((uint32_t)gshot, (uint32_t)(gshot >> 32), (uint32_t)j, 1u, k0, k1);
          if ((r.x >> 8) >= __ldg(p.thresh + j)) {
            const uint32_t ch = __ldg(p.chan + j);
            if (ch == 0) { const uint32_t v = 1u + __umulhi(r.y, d * d - 1u); a = v % d; b = v / d; }
            else { const uint32_t e = 1u + __umulhi(r.y, d - 1u); if (ch == 1) a = e; else b = e; }
          }
        }
        if (a) *xa = (uint8_t)mod_d(A, (uint32_t)*xa + a);                   // program.py:160-161
        if (b) *za = (uint8_t)mod_d(A, (uint32_t)*za + b);
        break;
      }
      default: break;     // I, Paulis (frames commute with them up to phase, program.py:82-89), BARRIER
    }
  }
}

bool is_prime(int d) {
  if (d < 2) return false;
  for (int f = 2; f * f <= d; ++f)
    if (d % f == 0) return false;
  return true;
}

int check_dims(int n, int d) {
  if (n < 1) return SDIMB_EINVAL;
  if (d < 2 || d > 127 || !is_prime(d)) return SDIMB_EDIM;
  return SDIMB_OK;
}

int block_threads(int W) {
  int t = ((W / 4) + 31) / 32 * 32;
  if (t < 32) t = 32;
  if (t > kMaxThreads) t = t > kWideThreads ? kWideThreads : t;   // > 256 threads run interp_kernel_wide
  return t;
}

size_t scratch_bytes(int np) {
  const size_t W = 2 * (size_t)np;
  return 4 * W + W + 32 * 4 + 4 * 4 + 32 * 16 + 4 * (size_t)np + 2 * (W / 4) + 2 * (size_t)np + 128;
}

// Which interpreter a (n, d, flags) call runs: 0 = uint8 lanes in global memory, 1 = uint8 lanes resident in
// shared memory, 2 = bit-plane resident (d = 2, 3).  Negative = error code.
int plan_kernel(int n, int d, uint32_t flags, int np) {
  if ((flags & SDIMB_FORCE_GLOBAL) && (flags & (SDIMB_FORCE_RESIDENT | SDIMB_FORCE_PLANES))) return SDIMB_EINVAL;
  if ((flags & SDIMB_FORCE_LANES) && (flags & SDIMB_FORCE_PLANES)) return SDIMB_EINVAL;
  SdimbLayout L;
  const int rc = sdimb_layout(n, d, &L);
  if (rc) return rc;
  const bool fits = (size_t)L.shot_bytes + scratch_bytes(np) <= (size_t)kSmemLimit;
  const bool planes_fit = (d == 2 || d == 3) && planes::planes_smem_bytes(n, d) <= (size_t)kSmemLimit;
  if ((flags & SDIMB_FORCE_PLANES) && !planes_fit) return SDIMB_ETOOBIG;
  if ((flags & SDIMB_FORCE_RESIDENT) && !fits) return SDIMB_ETOOBIG;
  if (planes_fit && !(flags & (SDIMB_FORCE_GLOBAL | SDIMB_FORCE_RESIDENT | SDIMB_FORCE_LANES))) return 2;
  return (fits && !(flags & SDIMB_FORCE_GLOBAL)) ? 1 : 0;
}

}  // namespace

extern "C" {

int sdimb_version(void) { return SDIMB_VERSION; }

const char* sdimb_strerror(int code) {
  switch (code) {
    case SDIMB_OK: return "ok";
    case SDIMB_EINVAL: return "invalid argument";
    case SDIMB_EDIM: return "dimension must be a prime in [2, 127]";
    case SDIMB_EOP: return "Invalid gate value";
    case SDIMB_ECUDA: return "CUDA error (is a GPU present? there is no CPU fallback)";
    case SDIMB_ETOOBIG: return "tableau does not fit in shared memory for the resident interpreter";
    default: return "unknown error";
  }
}

int sdimb_layout(int n, int d, SdimbLayout* out) {
  if (!out) return SDIMB_EINVAL;
  const int rc = check_dims(n, d);
  if (rc) return rc;
  const Arith A = make_arith(d);
  out->n = n;
  out->d = d;
  out->np = (n + 15) / 16 * 16;
  out->lanes = 2 * out->np;
  out->order = (int32_t)A.order;
  out->phase_order = (int32_t)A.po;
  out->row_bytes = 2ll * out->lanes;
  out->phase_offset = (int64_t)n * out->row_bytes;
  out->shot_bytes = out->phase_offset + out->lanes;
  return SDIMB_OK;
}

int sdimb_init(void* tableau, int n, int d, int64_t shots, void* stream) {
  SdimbLayout L;
  const int rc = sdimb_layout(n, d, &L);
  if (rc) return rc;
  if (!tableau || shots < 0) return SDIMB_EINVAL;
  if (shots == 0) return SDIMB_OK;
  const int grid = (int)(shots < 148 * 16 ? shots : 148 * 16);
  init_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((uint8_t*)tableau, n, L.np, L.lanes, L.row_bytes, L.shot_bytes,
                                                      shots);
  g_launches++;
  return cudaGetLastError() == cudaSuccess ? SDIMB_OK : SDIMB_ECUDA;
}

int sdimb_run(const SdimbRunArgs* a) {
  if (!a || a->struct_size != sizeof(SdimbRunArgs)) return SDIMB_EINVAL;
  SdimbLayout L;
  const int rc = sdimb_layout(a->n, a->d, &L);
  if (rc) return rc;
  if (a->shots < 0 || a->n_ops < 0 || a->n_meas < 0 || a->n_noise < 0) return SDIMB_EINVAL;
  if (plan_kernel(a->n, a->d, a->flags, L.np) < 0) return plan_kernel(a->n, a->d, a->flags, L.np);
  if (a->shots == 0) return SDIMB_OK;
  if (a->n_ops > 0 && !a->ops) return SDIMB_EINVAL;
  if (a->n_meas > 0 && (!a->records || a->rec_stride < a->n_meas)) return SDIMB_EINVAL;
  if (a->n_noise > 0 && !a->replay_noise && (!a->noise_thresh24 || !a->noise_channel)) return SDIMB_EINVAL;

  const int kernel = plan_kernel(a->n, a->d, a->flags, L.np);
  if (kernel < 0) return kernel;
  const size_t scratch = scratch_bytes(L.np);
  const bool use_planes = kernel == 2, resident = kernel >= 1;
  const bool need_tab = !resident || !(a->flags & SDIMB_FRESH) || (a->flags & SDIMB_WRITEBACK);
  if (need_tab && !a->tableau) return SDIMB_EINVAL;

  KParams p;
  std::memset(&p, 0, sizeof(p));
  p.tab = (uint8_t*)a->tableau;
  p.ops = (const int4*)a->ops;
  p.n_ops = a->n_ops;
  p.records = a->records;
  p.rec_stride = a->rec_stride;
  p.n_meas = a->n_meas;
  p.replay_meas = a->replay_meas;
  p.replay_noise = a->replay_noise;
  p.thresh = a->noise_thresh24;
  p.chan = a->noise_channel;
  p.n_noise = a->n_noise;
  p.seed = a->seed;
  p.shots = a->shots;
  p.shot_offset = a->shot_offset;
  p.n = L.n; p.np = L.np; p.W = L.lanes;
  p.row_bytes = L.row_bytes; p.phase_off = L.phase_offset; p.shot_bytes = L.shot_bytes;
  p.A = make_arith(a->d);
  p.flags = a->flags;
  p.resident = kernel == 1 ? 1 : 0;

  int dev = 0, sms = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return SDIMB_ECUDA;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return SDIMB_ECUDA;
  if (use_planes) {
    auto kern = (a->d == 2) ? planes::interp_planes_kernel<2> : planes::interp_planes_kernel<3>;
    // one warp per shot unless shared memory leaves the SM short of warps and the stream is scheduled
    int nw = 1;
    size_t smem = planes::planes_smem_bytes(a->n, a->d, 1);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit) != cudaSuccess)
      return SDIMB_ECUDA;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem) != cudaSuccess || per_sm < 1)
      return SDIMB_ECUDA;
    if ((a->flags & SDIMB_SCHEDULED) && per_sm < 12) {
      nw = SDIMB_SCHED_WARPS;
      smem = planes::planes_smem_bytes(a->n, a->d, nw);
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * nw, smem) != cudaSuccess || per_sm < 1)
        return SDIMB_ECUDA;
    }
    int64_t grid = (int64_t)sms * per_sm;
    if (grid > a->shots) grid = a->shots;
    if (a->scratch && a->scratch_bytes >= (int64_t)sizeof(unsigned int)) {   // dynamic shot claiming
      p.shot_counter = (unsigned int*)a->scratch;
      if (cudaMemsetAsync(p.shot_counter, 0, sizeof(unsigned int), (cudaStream_t)a->stream) != cudaSuccess)
        return SDIMB_ECUDA;
    }
    kern<<<(unsigned)grid, 32 * nw, smem, (cudaStream_t)a->stream>>>(p);
    g_launches++;
    return cudaGetLastError() == cudaSuccess ? SDIMB_OK : SDIMB_ECUDA;
  }
  const int threads = block_threads(L.lanes);
  const size_t smem = scratch + (resident ? (size_t)L.shot_bytes : 0);
  auto lane_kernel = threads > kMaxThreads ? interp_kernel_wide : interp_kernel;
  if (cudaFuncSetAttribute(lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return SDIMB_ECUDA;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lane_kernel, threads, smem) != cudaSuccess || per_sm < 1)
    return SDIMB_ECUDA;
  int64_t grid = (int64_t)sms * per_sm;
  if (grid > a->shots) grid = a->shots;
  lane_kernel<<<(unsigned)grid, threads, smem, (cudaStream_t)a->stream>>>(p);
  g_launches++;
  return cudaGetLastError() == cudaSuccess ? SDIMB_OK : SDIMB_ECUDA;
}

int sdimb_export(const void* tableau, int n, int d, int64_t shot, int64_t* x, int64_t* z, int64_t* p, int64_t* dx,
                 int64_t* dz, int64_t* dp, void* stream) {
  SdimbLayout L;
  const int rc = sdimb_layout(n, d, &L);
  if (rc) return rc;
  if (!tableau || shot < 0 || !x || !z || !p || !dx || !dz || !dp) return SDIMB_EINVAL;
  const uint8_t* T = (const uint8_t*)tableau + shot * L.shot_bytes;
  const int64_t total = (int64_t)n * n;
  const int grid = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  export_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(T, n, L.np, L.lanes, L.row_bytes, L.phase_offset, x, z, p, dx,
                                                        dz, dp);
  g_launches++;
  return cudaGetLastError() == cudaSuccess ? SDIMB_OK : SDIMB_ECUDA;
}

// Workspace of the host-buffer entry: one grow-only device arena, one grow-only pinned staging buffer, a stream
// and two events, created on first use and reused by later calls (cudaMalloc/cudaFree per call cost more than
// the simulation of a small batch).  Calls are serialised by a mutex; sdimb_release_workspace() frees it.
namespace {
struct HostWorkspace {
  std::mutex mu;
  int device = -1;
  void* dev = nullptr;
  size_t dev_cap = 0;
  void* pin = nullptr;
  size_t pin_cap = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  void release() {
    if (dev) cudaFree(dev);
    if (pin) cudaFreeHost(pin);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (stream) cudaStreamDestroy(stream);
    dev = pin = nullptr; dev_cap = pin_cap = 0; stream = nullptr; e0 = e1 = nullptr; device = -1;
  }
  bool prepare(size_t dev_bytes, size_t pin_bytes) {
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) return false;
    if (cur != device) { release(); device = cur; }
    if (!stream && cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) return false;
    if (!e0 && cudaEventCreate(&e0) != cudaSuccess) return false;
    if (!e1 && cudaEventCreate(&e1) != cudaSuccess) return false;
    if (dev_bytes > dev_cap) {
      if (dev) cudaFree(dev);
      dev = nullptr; dev_cap = 0;
      if (cudaMalloc(&dev, dev_bytes) != cudaSuccess) return false;
      dev_cap = dev_bytes;
    }
    if (pin_bytes > pin_cap) {
      if (pin) cudaFreeHost(pin);
      pin = nullptr; pin_cap = 0;
      if (cudaHostAlloc(&pin, pin_bytes, cudaHostAllocDefault) != cudaSuccess) return false;
      pin_cap = pin_bytes;
    }
    return true;
  }
};
HostWorkspace g_ws;
inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
}  // namespace

int sdimb_release_workspace(void) {
  std::lock_guard<std::mutex> lock(g_ws.mu);
  g_ws.release();
  return SDIMB_OK;
}

int sdimb_simulate_host(int n, int d, int64_t shots, int64_t shot_offset, const int32_t* ops, int64_t n_ops,
                        uint8_t* records, int64_t n_meas, const uint8_t* replay_meas, const uint8_t* replay_noise,
                        const uint32_t* noise_thresh24, const uint8_t* noise_channel, int64_t n_noise, uint64_t seed,
                        uint32_t flags, float* elapsed_ms) {
  SdimbLayout L;
  int rc = sdimb_layout(n, d, &L);
  if (rc) return rc;
  if (shots < 0 || n_ops < 0 || n_meas < 0 || n_noise < 0) return SDIMB_EINVAL;
  if ((n_ops > 0 && !ops) || (n_meas > 0 && shots > 0 && !records)) return SDIMB_EINVAL;
  for (int64_t i = 0; i < n_ops; ++i) {   // host-side validation: "Invalid gate value" (sdim/program.py:381-382)
    const int32_t* o = ops + 4 * i;
    if (o[0] < 0 || o[0] > SDIMB_OP_N1 || o[1] < 0 || o[1] >= n) return SDIMB_EOP;
    const bool two = o[0] >= SDIMB_OP_CNOT && o[0] <= SDIMB_OP_SWAP;
    if (two && (o[2] < 0 || o[2] >= n || o[2] == o[1])) return SDIMB_EOP;
    const bool meas = o[0] >= SDIMB_OP_M && o[0] <= SDIMB_OP_RESET;
    if (meas && (o[3] < 0 || o[3] >= n_meas)) return SDIMB_EOP;
    if (o[0] == SDIMB_OP_N1 && (o[3] < 0 || o[3] >= n_noise)) return SDIMB_EOP;
  }
  if (n_noise > 0 && !replay_noise && (!noise_thresh24 || !noise_channel)) return SDIMB_EINVAL;
  if (shots == 0) return SDIMB_OK;

  const uint32_t mode_flags = flags & (SDIMB_FORCE_GLOBAL | SDIMB_FORCE_RESIDENT | SDIMB_FORCE_LANES | SDIMB_FORCE_PLANES);
  const int kernel = plan_kernel(n, d, mode_flags, L.np);
  if (kernel < 0) return kernel;
  std::vector<int32_t> sched;
  const int32_t* up_ops = ops;
  int64_t up_n = n_ops;
  uint32_t sched_flag = 0;
  if (kernel == 2 && n_ops > 0) {          // bit-plane interpreter: upload the layered stream
    sched.resize((size_t)(2 * n_ops + 1) * 4);
    rc = sdimb_schedule(n, ops, n_ops, sched.data(), 2 * n_ops + 1, &up_n);
    if (rc) return rc;
    up_ops = sched.data();
    sched_flag = SDIMB_SCHEDULED;
  }

  // carve the device arena
  const size_t b_ops = align256((size_t)up_n * 16), b_rec = align256((size_t)shots * n_meas);
  const size_t b_rm = replay_meas ? b_rec : 0;
  const size_t b_rn = replay_noise ? align256((size_t)shots * n_noise * 2) : 0;
  const size_t b_th = (n_noise && noise_thresh24) ? align256((size_t)n_noise * 4) : 0;
  const size_t b_ch = (n_noise && noise_channel) ? align256((size_t)n_noise) : 0;
  const size_t b_tab = kernel == 0 ? align256((size_t)shots * L.shot_bytes) : 0;
  const size_t b_scr = align256((size_t)sdimb_scratch_bytes(n, d, mode_flags));
  const size_t total = b_ops + b_rec + b_rm + b_rn + b_th + b_ch + b_tab + b_scr + 256;

  std::lock_guard<std::mutex> lock(g_ws.mu);
  if (!g_ws.prepare(total, b_rec + 256)) { cudaGetLastError(); return SDIMB_ECUDA; }
  cudaStream_t st = g_ws.stream;
  uint8_t* base = (uint8_t*)g_ws.dev;
  uint8_t* d_ops = base; base += b_ops;
  uint8_t* d_rec = base; base += b_rec;
  uint8_t* d_rm = b_rm ? base : nullptr; base += b_rm;
  uint8_t* d_rn = b_rn ? base : nullptr; base += b_rn;
  uint8_t* d_th = b_th ? base : nullptr; base += b_th;
  uint8_t* d_ch = b_ch ? base : nullptr; base += b_ch;
  uint8_t* d_tab = b_tab ? base : nullptr; base += b_tab;
  uint8_t* d_scr = b_scr ? base : nullptr;
  rc = SDIMB_ECUDA;
  do {
    if (cudaEventRecord(g_ws.e0, st) != cudaSuccess) break;
    if (up_n && cudaMemcpyAsync(d_ops, up_ops, (size_t)up_n * 16, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
    if (d_rm && cudaMemcpyAsync(d_rm, replay_meas, (size_t)shots * n_meas, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
    if (d_rn && cudaMemcpyAsync(d_rn, replay_noise, (size_t)shots * n_noise * 2, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
    if (d_th && cudaMemcpyAsync(d_th, noise_thresh24, (size_t)n_noise * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
    if (d_ch && cudaMemcpyAsync(d_ch, noise_channel, (size_t)n_noise, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
    SdimbRunArgs a;
    std::memset(&a, 0, sizeof(a));
    a.struct_size = sizeof(a);
    a.flags = mode_flags | SDIMB_FRESH | sched_flag;
    a.n = n; a.d = d; a.shots = shots; a.shot_offset = shot_offset;
    a.tableau = d_tab;
    a.ops = (const int32_t*)d_ops; a.n_ops = up_n;
    a.records = d_rec; a.n_meas = n_meas; a.rec_stride = n_meas;
    a.replay_meas = d_rm; a.replay_noise = d_rn;
    a.noise_thresh24 = (const uint32_t*)d_th; a.noise_channel = d_ch; a.n_noise = n_noise;
    a.seed = seed; a.stream = st;
    a.scratch = d_scr; a.scratch_bytes = (int64_t)b_scr;
    rc = sdimb_run(&a);
    if (rc) break;
    rc = SDIMB_ECUDA;
    const size_t rec_bytes = (size_t)shots * n_meas;
    if (rec_bytes && cudaMemcpyAsync(g_ws.pin, d_rec, rec_bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
    if (cudaEventRecord(g_ws.e1, st) != cudaSuccess) break;
    if (cudaStreamSynchronize(st) != cudaSuccess) break;
    if (rec_bytes) std::memcpy(records, g_ws.pin, rec_bytes);
    if (elapsed_ms && cudaEventElapsedTime(elapsed_ms, g_ws.e0, g_ws.e1) != cudaSuccess) break;
    rc = SDIMB_OK;
  } while (0);
  if (rc == SDIMB_ECUDA) cudaGetLastError();
  return rc;
}

int64_t sdimb_launch_count(void) { return g_launches.load(); }

int sdimb_schedule(int n, const int32_t* ops, int64_t n_ops, int32_t* out, int64_t out_cap, int64_t* out_n) {
  if (n < 1 || n_ops < 0 || (n_ops > 0 && !ops) || !out || !out_n || out_cap < 2 * n_ops + 1) return SDIMB_EINVAL;
  // lw[q]: first layer in which row q may be read again (last writer + 1); lr[q]: first layer in which row q
  // may be written again (last reader or writer + 1).  Layers restart after every collective op.
  std::vector<int> lw(n, 0), lr(n, 0);
  std::vector<std::vector<int64_t>> layers;
  int64_t w = 0;
  auto flush = [&]() {
    for (auto& layer : layers) {
      if (layer.empty()) continue;
      int k = 0;
      for (int64_t i : layer) {
        const int32_t* o = ops + 4 * i;
        int32_t* r = out + 4 * w++;
        r[0] = o[0] | ((k++ % SDIMB_SCHED_WARPS) << SDIMB_OP_WARP_SHIFT);
        r[1] = o[1]; r[2] = o[2]; r[3] = o[3];
      }
      int32_t* b = out + 4 * w++;
      b[0] = SDIMB_OP_BARRIER; b[1] = 0; b[2] = -1; b[3] = -1;
    }
    layers.clear();
    std::fill(lw.begin(), lw.end(), 0);
    std::fill(lr.begin(), lr.end(), 0);
  };
  for (int64_t i = 0; i < n_ops; ++i) {
    const int32_t* o = ops + 4 * i;
    const int op = o[0] & SDIMB_OP_MASK, a = o[1], b = o[2];
    if (op < 0 || op > SDIMB_OP_BARRIER || a < 0 || a >= n) return SDIMB_EOP;
    if (op == SDIMB_OP_I || op == SDIMB_OP_BARRIER) continue;
    if (op >= SDIMB_OP_M && op <= SDIMB_OP_RESET) {     // collective: everything before it completes first
      flush();
      int32_t* r = out + 4 * w++;
      r[0] = op; r[1] = a; r[2] = -1; r[3] = o[3];
      continue;
    }
    const bool two = op >= SDIMB_OP_CNOT && op <= SDIMB_OP_SWAP;
    if (two && (b < 0 || b >= n || b == a)) return SDIMB_EOP;
    const bool reads_only = (op >= SDIMB_OP_X && op <= SDIMB_OP_Z_INV) || op == SDIMB_OP_N1;
    int level;
    if (reads_only) {
      level = lw[a];
      lr[a] = std::max(lr[a], level + 1);
    } else {
      level = std::max(lw[a], lr[a]);
      if (two) level = std::max(level, std::max(lw[b], lr[b]));
      lw[a] = lr[a] = level + 1;
      if (two) lw[b] = lr[b] = level + 1;
    }
    if ((size_t)level >= layers.size()) layers.resize(level + 1);
    layers[level].push_back(i);
  }
  flush();
  *out_n = w;
  return SDIMB_OK;
}

int sdimb_frames(int n, int d, int64_t shots, int64_t shot_offset, const int32_t* ops, int64_t n_ops,
                 const uint8_t* reference, uint8_t* records, int64_t n_meas, int64_t rec_stride, uint8_t* frames,
                 const uint8_t* replay_z0, const uint8_t* replay_zm, const uint8_t* replay_noise,
                 const uint32_t* noise_thresh24, const uint8_t* noise_channel, int64_t n_noise, uint64_t seed,
                 void* stream) {
  const int rc = check_dims(n, d);
  if (rc) return rc;
  if (shots < 0 || n_ops < 0 || n_meas < 0 || n_noise < 0) return SDIMB_EINVAL;
  if (shots == 0) return SDIMB_OK;
  if ((n_ops > 0 && !ops) || !frames) return SDIMB_EINVAL;
  if (n_meas > 0 && (!reference || !records || rec_stride < n_meas)) return SDIMB_EINVAL;
  if (n_noise > 0 && !replay_noise && (!noise_thresh24 || !noise_channel)) return SDIMB_EINVAL;
  FParams p;
  std::memset(&p, 0, sizeof(p));
  p.ops = (const int4*)ops; p.n_ops = n_ops;
  p.reference = reference; p.records = records; p.n_meas = n_meas; p.rec_stride = rec_stride;
  p.frames = frames; p.pitch = (shots + 127) / 128 * 128;
  p.replay_z0 = replay_z0; p.replay_zm = replay_zm; p.replay_noise = replay_noise;
  p.thresh = noise_thresh24; p.chan = noise_channel; p.n_noise = n_noise;
  p.seed = seed; p.shots = shots; p.shot_offset = shot_offset; p.n = n;
  p.A = make_arith(d);
  const int64_t grid = (shots + 127) / 128;
  frame_kernel<<<(unsigned)grid, 128, 0, (cudaStream_t)stream>>>(p);
  g_launches++;
  return cudaGetLastError() == cudaSuccess ? SDIMB_OK : SDIMB_ECUDA;
}

int64_t sdimb_scratch_bytes(int n, int d, uint32_t flags) {
  SdimbLayout L;
  if (sdimb_layout(n, d, &L)) return 0;
  if (plan_kernel(n, d, flags, L.np) != 2) return 0;
  return 256;   // the shot counter
}

int sdimb_plan(int n, int d, uint32_t flags, int* kernel, int* needs_tableau) {
  SdimbLayout L;
  const int rc = sdimb_layout(n, d, &L);
  if (rc) return rc;
  const int k = plan_kernel(n, d, flags, L.np);
  if (k < 0) return k;
  if (kernel) *kernel = k;
  if (needs_tableau) *needs_tableau = (k == 0 || !(flags & SDIMB_FRESH) || (flags & SDIMB_WRITEBACK)) ? 1 : 0;
  return SDIMB_OK;
}

}  // extern "C"
