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
#include <cstring>
#include <vector>

namespace {

std::atomic<int64_t> g_launches{0};

constexpr int kMaxThreads = 256;
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
};

// Shared scratch common to both modes (carved from dynamic shared memory after the resident tableau).
struct Scratch {
  uint32_t* dot;       // [W]    per-lane accumulator of Z[:,i] . x_p          (random branch)
  uint32_t* fw;        // [W/4]  packed factors f = -X[q,i] mod d, 4 lanes/word (random branch)
  uint32_t* red;       // [32]   cross-warp reduction scratch
  uint32_t* cnt;       // [4]    list lengths
  int4* ops;           // [32]   staged op batch (N1 rows carry the decoded event in .z)
  uint16_t* ar;        // [np]   active rows: qudits on which the pivot acts / active generators (det branch)
  uint16_t* aw;        // [W/4]  active words: lane quads holding a non-zero factor
  uint8_t* xs;         // [np]   pivot column X (random branch) / factors of the active generators (det branch)
  uint8_t* zs;         // [np]   pivot column Z
  uint8_t* inv;        // [128]  multiplicative inverses mod d
};

__device__ __forceinline__ uint32_t block_sum(uint32_t v, uint32_t* red) {
  v = __reduce_add_sync(0xFFFFFFFFu, v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  uint32_t t = 0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}

__device__ __forceinline__ uint32_t block_min(uint32_t v, uint32_t* red) {
  v = __reduce_min_sync(0xFFFFFFFFu, v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  uint32_t t = kNoPivot;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t = min(t, red[i]);
  return t;
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
// Unitary gates: every generator lane is independent, so threads sweep 4-lane words with no barrier.
// Closed forms: SURVEY Appendix A-1/A-2 (restating tableau_optimized.py:5-118, tableau_gates.py:27-261,298-329).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void gate_h(uint8_t* T, const KParams& p, int a, bool inverse) {
  const Arith& A = p.A;
  uint32_t* xa = reinterpret_cast<uint32_t*>(T + (int64_t)a * p.row_bytes);
  uint32_t* za = xa + p.W / 4;
  uint32_t* P = reinterpret_cast<uint32_t*>(T + p.phase_off);
  for (int w = threadIdx.x; w < p.W / 4; w += blockDim.x) {
    const uint32_t x = xa[w], z = za[w];
    if ((x | z) == 0) continue;
    const uint32_t ph = P[w];
    uint32_t nx = 0, nz = 0, nph = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t xb = byte_of(x, k), zb = byte_of(z, k);
      // phase += po * new_x * new_z == -po * x * z        (tableau_optimized.py:17-30,45-58)
      uint32_t pb = byte_of(ph, k) + A.order - A.po * mod_d(A, xb * zb);
      pb = pb >= A.order ? pb - A.order : pb;
      const uint32_t nxb = inverse ? zb : neg_d(A, zb);   // H: (x,z) <- (-z,x);  H^-1: (x,z) <- (z,-x)
      const uint32_t nzb = inverse ? neg_d(A, xb) : xb;
      nx |= nxb << (8 * k); nz |= nzb << (8 * k); nph |= pb << (8 * k);
    }
    xa[w] = nx; za[w] = nz; P[w] = nph;
  }
}

__device__ __forceinline__ void gate_p(uint8_t* T, const KParams& p, int a, bool inverse) {
  const Arith& A = p.A;
  const uint32_t* xa = reinterpret_cast<const uint32_t*>(T + (int64_t)a * p.row_bytes);
  uint32_t* za = reinterpret_cast<uint32_t*>(T + (int64_t)a * p.row_bytes) + p.W / 4;
  uint32_t* P = reinterpret_cast<uint32_t*>(T + p.phase_off);
  for (int w = threadIdx.x; w < p.W / 4; w += blockDim.x) {
    const uint32_t x = xa[w];
    if (x == 0) continue;
    const uint32_t z = za[w], ph = P[w];
    uint32_t nz = 0, nph = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t xb = byte_of(x, k), zb = byte_of(z, k);
      // even d: phase +-= x^2 (mod 2d); odd d: phase +-= x(x-1)/2 (mod d)   (tableau_optimized.py:62-96)
      const uint32_t inc = (A.po == 2) ? mod_o(A, xb * xb) : mod_d(A, (xb * (xb - 1u)) >> 1);
      const uint32_t pb = mod_o(A, byte_of(ph, k) + (inverse ? A.order - inc : inc));
      const uint32_t nzb = mod_d(A, zb + (inverse ? A.d - xb : xb));
      nz |= nzb << (8 * k); nph |= pb << (8 * k);
    }
    za[w] = nz; P[w] = nph;
  }
}

// Conjugation by the Pauli X^a Z^b on qudit q: phase += po * (b*x - a*z).  Covers X, X_INV, Z, Z_INV,
// N1 noise and the RESET correction (tableau_gates.py:27-137, program.py:335-339).
__device__ __forceinline__ void gate_pauli(uint8_t* T, const KParams& p, int q, uint32_t a, uint32_t b) {
  const Arith& A = p.A;
  const uint32_t* xq = reinterpret_cast<const uint32_t*>(T + (int64_t)q * p.row_bytes);
  const uint32_t* zq = xq + p.W / 4;
  uint32_t* P = reinterpret_cast<uint32_t*>(T + p.phase_off);
  const uint32_t na = a ? A.d - a : 0u;
  for (int w = threadIdx.x; w < p.W / 4; w += blockDim.x) {
    const uint32_t x = b ? xq[w] : 0u, z = na ? zq[w] : 0u;
    if ((x | z) == 0) continue;
    const uint32_t ph = P[w];
    uint32_t nph = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t t = mod_d(A, b * byte_of(x, k) + na * byte_of(z, k));
      nph |= mod_o(A, byte_of(ph, k) + A.po * t) << (8 * k);
    }
    P[w] = nph;
  }
}

__device__ __forceinline__ void gate_cnot(uint8_t* T, const KParams& p, int a, int b, bool inverse) {
  const Arith& A = p.A;
  uint32_t* rowa = reinterpret_cast<uint32_t*>(T + (int64_t)a * p.row_bytes);
  uint32_t* rowb = reinterpret_cast<uint32_t*>(T + (int64_t)b * p.row_bytes);
  const int wz = p.W / 4;
  for (int w = threadIdx.x; w < wz; w += blockDim.x) {
    const uint32_t xa = rowa[w], zb = rowb[wz + w];
    if ((xa | zb) == 0) continue;
    const uint32_t xb = rowb[w], za = rowa[wz + w];
    uint32_t nxb = 0, nza = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // x[t] +-= x[c];  z[c] -+= z[t]        (tableau_optimized.py:99-118, (d-1)*z == -z)
      const uint32_t xak = byte_of(xa, k), zbk = byte_of(zb, k);
      nxb |= mod_d(A, byte_of(xb, k) + (inverse ? neg_d(A, xak) : xak)) << (8 * k);
      nza |= mod_d(A, byte_of(za, k) + (inverse ? zbk : neg_d(A, zbk))) << (8 * k);
    }
    rowb[w] = nxb; rowa[wz + w] = nza;
  }
}

__device__ __forceinline__ void gate_cz(uint8_t* T, const KParams& p, int a, int b, bool inverse) {
  const Arith& A = p.A;
  uint32_t* rowa = reinterpret_cast<uint32_t*>(T + (int64_t)a * p.row_bytes);
  uint32_t* rowb = reinterpret_cast<uint32_t*>(T + (int64_t)b * p.row_bytes);
  uint32_t* P = reinterpret_cast<uint32_t*>(T + p.phase_off);
  const int wz = p.W / 4;
  for (int w = threadIdx.x; w < wz; w += blockDim.x) {
    const uint32_t xa = rowa[w], xb = rowb[w];
    if ((xa | xb) == 0) continue;
    const uint32_t za = rowa[wz + w], zb = rowb[wz + w], ph = P[w];
    uint32_t nza = 0, nzb = 0, nph = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // CZ = H^-1(t) CNOT(c,t) H(t) folded: z[a] +-= x[b]; z[b] +-= x[a]; phase +-= po*x[a]*x[b]
      const uint32_t xak = byte_of(xa, k), xbk = byte_of(xb, k);
      const uint32_t prod = A.po * mod_d(A, xak * xbk);
      nph |= mod_o(A, byte_of(ph, k) + (inverse ? A.order - prod : prod)) << (8 * k);
      nza |= mod_d(A, byte_of(za, k) + (inverse ? neg_d(A, xbk) : xbk)) << (8 * k);
      nzb |= mod_d(A, byte_of(zb, k) + (inverse ? neg_d(A, xak) : xak)) << (8 * k);
    }
    rowa[wz + w] = nza; rowb[wz + w] = nzb; P[w] = nph;
  }
}

__device__ __forceinline__ void gate_swap(uint8_t* T, const KParams& p, int a, int b) {
  uint32_t* rowa = reinterpret_cast<uint32_t*>(T + (int64_t)a * p.row_bytes);
  uint32_t* rowb = reinterpret_cast<uint32_t*>(T + (int64_t)b * p.row_bytes);
  const int wz = p.W / 4;
  // each thread swaps the X and Z words of the lanes it owns (lane ownership must hold across gates: there is
  // no barrier between consecutive gates)
  for (int w = threadIdx.x; w < wz; w += blockDim.x) {
    const uint32_t tx = rowa[w], tz = rowa[wz + w];
    rowa[w] = rowb[w]; rowa[wz + w] = rowb[wz + w];
    rowb[w] = tx; rowb[wz + w] = tz;
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

__device__ uint32_t measure(uint8_t* T, const KParams& p, Scratch& S, int q, int64_t slot, int64_t shot_local) {
  const Arith& A = p.A;
  const int n = p.n, W = p.W, npad = p.np, nt = blockDim.x, tid = threadIdx.x;
  const int wz = W / 4;
  uint8_t* rowq = T + (int64_t)q * p.row_bytes;
  uint8_t* P8 = T + p.phase_off;
  if (tid < 2) S.cnt[tid] = 0;   // cnt[3] holds the live-op mask of the current batch
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

  // outcome used if the measurement is random: replayed draw or Philox (reference: random.choice, :332)
  uint32_t draw;
  if (p.replay_meas) {
    draw = p.replay_meas[shot_local * p.n_meas + slot];
  } else {
    const uint64_t gshot = (uint64_t)(p.shot_offset + shot_local);
    const uint4 r = philox4x32((uint32_t)gshot, (uint32_t)(gshot >> 32), (uint32_t)slot, 0u, (uint32_t)p.seed,
                               (uint32_t)(p.seed >> 32));
    draw = __umulhi(r.x, A.d);
  }

  uint32_t outcome, rec;
  if (piv != kNoPivot) {
    // -- random branch (tableau_prime.py:294-334), with exponentiate (:365-380) folded into the gather ----
    const uint32_t v = rowq[piv];
    const uint32_t e = S.inv[v];
    const uint32_t ps_old = P8[piv];
    uint32_t sd_raw = 0;
    for (int r = tid; r < n; r += nt) {                       // pivot column -> xs/zs, active-row list
      const uint8_t* row = T + (int64_t)r * p.row_bytes;
      const uint32_t xr = row[piv], zr = row[W + piv];
      S.xs[r] = (uint8_t)mod_d(A, xr * e);
      S.zs[r] = (uint8_t)mod_d(A, zr * e);
      sd_raw += mod_d(A, xr * zr);
      if (xr | zr) S.ar[atomicAdd(&S.cnt[0], 1u)] = (uint16_t)r;
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
    // Column accesses cost one DRAM sector per byte, so only entries that change are written: the
    // stabilizer lane is non-zero exactly on the pivot's support (the active rows), the destabilizer
    // lane is read back and rewritten where it differs.
    for (int i = tid; i < nr_a; i += nt) {
      uint8_t* row = T + (int64_t)S.ar[i] * p.row_bytes;
      row[piv] = 0;
      row[W + piv] = 0;
    }
    for (int r = tid; r < n; r += nt) {
      uint8_t* row = T + (int64_t)r * p.row_bytes;
      const uint8_t xs = S.xs[r], zs = S.zs[r];
      if (row[npad + piv] != xs) row[npad + piv] = xs;
      if (row[W + npad + piv] != zs) row[W + npad + piv] = zs;
    }
    __syncthreads();
    if (tid == 0) rowq[W + piv] = 1;
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
      int off = total, all = total;
      for (int wv = 0; wv < (nt >> 5); ++wv) {
        const int c = (int)S.red[wv];
        if (wv < (tid >> 5)) off += c;
        all += c;
      }
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
__global__ void __launch_bounds__(kMaxThreads) interp_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int64_t tab_smem = p.resident ? p.shot_bytes : 0;
  Scratch S;
  S.dot = reinterpret_cast<uint32_t*>(smem + tab_smem);
  S.fw = S.dot + p.W;
  S.red = S.fw + p.W / 4;
  S.cnt = S.red + 32;
  S.ops = reinterpret_cast<int4*>(S.cnt + 4);
  S.ar = reinterpret_cast<uint16_t*>(S.ops + 32);
  S.aw = S.ar + p.np;
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
        switch (op.x) {
          case SDIMB_OP_X: gate_pauli(T, p, op.y, 1u, 0u); break;
          case SDIMB_OP_X_INV: gate_pauli(T, p, op.y, A.d - 1u, 0u); break;
          case SDIMB_OP_Z: gate_pauli(T, p, op.y, 0u, 1u); break;
          case SDIMB_OP_Z_INV: gate_pauli(T, p, op.y, 0u, A.d - 1u); break;
          case SDIMB_OP_H: gate_h(T, p, op.y, false); break;
          case SDIMB_OP_H_INV: gate_h(T, p, op.y, true); break;
          case SDIMB_OP_P: gate_p(T, p, op.y, false); break;
          case SDIMB_OP_P_INV: gate_p(T, p, op.y, true); break;
          case SDIMB_OP_CNOT: gate_cnot(T, p, op.y, op.z, false); break;
          case SDIMB_OP_CNOT_INV: gate_cnot(T, p, op.y, op.z, true); break;
          case SDIMB_OP_CZ: gate_cz(T, p, op.y, op.z, false); break;
          case SDIMB_OP_CZ_INV: gate_cz(T, p, op.y, op.z, true); break;
          case SDIMB_OP_SWAP: gate_swap(T, p, op.y, op.z); break;
          case SDIMB_OP_M_X:
            gate_h(T, p, op.y, true);                          // tableau_gates.py:292-296: H^-1 then measure
            // fallthrough
          case SDIMB_OP_M:
          case SDIMB_OP_RESET: {
            const uint32_t m = measure(T, p, S, op.y, op.w, shot);
            if (op.x == SDIMB_OP_RESET && m) gate_pauli(T, p, op.y, A.d - m, 0u);   // program.py:335-339
            break;
          }
          case SDIMB_OP_N1: gate_pauli(T, p, op.y, (uint32_t)op.z & 0xFFu, (uint32_t)op.z >> 8); break;
          default: break;   // rejected on the host before launch
        }
      }
    }
    __syncthreads();
    if (p.resident && (p.flags & SDIMB_WRITEBACK)) {
      const uint4* src = reinterpret_cast<const uint4*>(T);
      uint4* dst = reinterpret_cast<uint4*>(G);
      for (int64_t i = threadIdx.x; i < p.shot_bytes / 16; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
  }
}

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
  if (t > kMaxThreads) t = kMaxThreads;
  return t;
}

size_t scratch_bytes(int np) {
  const size_t W = 2 * (size_t)np;
  return 4 * W + W + 32 * 4 + 4 * 4 + 32 * 16 + 2 * (size_t)np + 2 * (W / 4) + 2 * (size_t)np + 128;
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
    kern<<<(unsigned)grid, 32 * nw, smem, (cudaStream_t)a->stream>>>(p);
    g_launches++;
    return cudaGetLastError() == cudaSuccess ? SDIMB_OK : SDIMB_ECUDA;
  }
  const int threads = block_threads(L.lanes);
  const size_t smem = scratch + (resident ? (size_t)L.shot_bytes : 0);
  if (cudaFuncSetAttribute(interp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return SDIMB_ECUDA;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, interp_kernel, threads, smem) != cudaSuccess || per_sm < 1)
    return SDIMB_ECUDA;
  int64_t grid = (int64_t)sms * per_sm;
  if (grid > a->shots) grid = a->shots;
  interp_kernel<<<(unsigned)grid, threads, smem, (cudaStream_t)a->stream>>>(p);
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
  if (shots == 0) return SDIMB_OK;

  cudaStream_t st = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  void *d_ops = nullptr, *d_rec = nullptr, *d_rm = nullptr, *d_rn = nullptr, *d_th = nullptr, *d_ch = nullptr,
       *d_tab = nullptr;
  const uint32_t mode_flags = flags & (SDIMB_FORCE_GLOBAL | SDIMB_FORCE_RESIDENT | SDIMB_FORCE_LANES | SDIMB_FORCE_PLANES);
  const int kernel = plan_kernel(n, d, mode_flags, L.np);
  if (kernel < 0) return kernel;
  const bool resident = kernel >= 1;
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
  rc = SDIMB_ECUDA;
  do {
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) break;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) break;
    if (cudaEventRecord(e0, st) != cudaSuccess) break;
    if (up_n && cudaMalloc(&d_ops, (size_t)up_n * 16) != cudaSuccess) break;
    if (n_meas && cudaMalloc(&d_rec, (size_t)shots * n_meas) != cudaSuccess) break;
    if (replay_meas && n_meas && cudaMalloc(&d_rm, (size_t)shots * n_meas) != cudaSuccess) break;
    if (replay_noise && n_noise && cudaMalloc(&d_rn, (size_t)shots * n_noise * 2) != cudaSuccess) break;
    if (n_noise && noise_thresh24 && cudaMalloc(&d_th, (size_t)n_noise * 4) != cudaSuccess) break;
    if (n_noise && noise_channel && cudaMalloc(&d_ch, (size_t)n_noise) != cudaSuccess) break;
    if (!resident && cudaMalloc(&d_tab, (size_t)shots * L.shot_bytes) != cudaSuccess) break;
    if (d_ops && cudaMemcpyAsync(d_ops, up_ops, (size_t)up_n * 16, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
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
    a.records = (uint8_t*)d_rec; a.n_meas = n_meas; a.rec_stride = n_meas;
    a.replay_meas = (const uint8_t*)d_rm; a.replay_noise = (const uint8_t*)d_rn;
    a.noise_thresh24 = (const uint32_t*)d_th; a.noise_channel = (const uint8_t*)d_ch; a.n_noise = n_noise;
    a.seed = seed; a.stream = st;
    rc = sdimb_run(&a);
    if (rc) break;
    rc = SDIMB_ECUDA;
    if (d_rec && cudaMemcpyAsync(records, d_rec, (size_t)shots * n_meas, cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
    if (cudaEventRecord(e1, st) != cudaSuccess) break;
    if (cudaStreamSynchronize(st) != cudaSuccess) break;
    if (elapsed_ms && cudaEventElapsedTime(elapsed_ms, e0, e1) != cudaSuccess) break;
    rc = SDIMB_OK;
  } while (0);
  cudaFree(d_ops); cudaFree(d_rec); cudaFree(d_rm); cudaFree(d_rn); cudaFree(d_th); cudaFree(d_ch); cudaFree(d_tab);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (st) cudaStreamDestroy(st);
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
