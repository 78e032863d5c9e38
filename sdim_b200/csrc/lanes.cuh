// uint8-lane interpreter (any prime d <= 127): tableau in shared memory (resident) or in the HBM store (global).
// Part of libsdimb (sdim_b200/csrc); included by sdimb.cu inside its anonymous namespace.
#pragma once

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
// products of the phase terms are computed per byte.  Pure arithmetic: `r` is updated in place, the new phase
// word is written to `ph`, and the return value says which of the four row words changed (bit 0 xa, 1 za, 2 xb,
// 3 zb) so that callers store only those.
enum { CH_XA = 1, CH_ZA = 2, CH_XB = 4, CH_ZB = 8 };

__device__ __forceinline__ uint32_t gate_math(const KParams& p, int op, uint32_t pa, uint32_t pb, Rows& r,
                                              uint32_t& ph) {
  const Arith& A = p.A;
  const Swar Sd = make_swar(A.d), So = make_swar(A.order);
  switch (op) {
    case SDIMB_OP_H:
    case SDIMB_OP_H_INV: {
      if ((r.xa | r.za) == 0) return 0;
      // phase += po * new_x * new_z == -po * x * z        (tableau_optimized.py:17-30,45-58)
      uint32_t prod = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) prod |= (A.po * mod_d(A, byte_of(r.xa, k) * byte_of(r.za, k))) << (8 * k);
      const uint32_t x = r.xa, z = r.za;
      if (op == SDIMB_OP_H) { r.xa = swar_neg(Sd, z); r.za = x; }            // (x,z) <- (-z, x)
      else { r.xa = z; r.za = swar_neg(Sd, x); }                             // (x,z) <- (z, -x)
      ph = swar_sub(So, ph, prod);
      return CH_XA | CH_ZA;
    }
    case SDIMB_OP_P:
    case SDIMB_OP_P_INV: {
      if (r.xa == 0) return 0;
      // even d: phase +-= x^2 (mod 2d); odd d: phase +-= x(x-1)/2 (mod d)   (tableau_optimized.py:62-96)
      uint32_t inc = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t xb = byte_of(r.xa, k);
        inc |= ((A.po == 2) ? mod_o(A, xb * xb) : mod_d(A, (xb * (xb - 1u)) >> 1)) << (8 * k);
      }
      if (op == SDIMB_OP_P) { r.za = swar_add(Sd, r.za, r.xa); ph = swar_add(So, ph, inc); }
      else { r.za = swar_sub(Sd, r.za, r.xa); ph = swar_sub(So, ph, inc); }
      return CH_ZA;
    }
    case SDIMB_OP_X: case SDIMB_OP_X_INV: case SDIMB_OP_Z: case SDIMB_OP_Z_INV: case SDIMB_OP_N1: {
      // conjugation by X^pa Z^pb: phase += po * (pb*x - pa*z)  (tableau_gates.py:27-137, program.py:335-339)
      const uint32_t na = pa ? A.d - pa : 0u;
      const uint32_t x = pb ? r.xa : 0u, z = na ? r.za : 0u;
      if ((x | z) == 0) return 0;
      const bool unit_x = pb == 0u || pb == 1u || pb == A.d - 1u, unit_z = pa == 0u || pa == 1u || pa == A.d - 1u;
      if (unit_x && unit_z) {                                 // exponents +-1: no products, all four lanes at once
        const uint32_t sx = A.po == 2 ? x << 1 : x, sz = A.po == 2 ? z << 1 : z;
        if (pb) ph = (pb == 1u) ? swar_add(So, ph, sx) : swar_sub(So, ph, sx);
        if (pa) ph = (pa == 1u) ? swar_sub(So, ph, sz) : swar_add(So, ph, sz);
        return 0;
      }
      uint32_t t = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) t |= mod_d(A, pb * byte_of(x, k) + na * byte_of(z, k)) << (8 * k);
      ph = swar_add(So, ph, A.po == 2 ? t << 1 : t);
      return 0;
    }
    case SDIMB_OP_CNOT:                                       // x[t] += x[c];  z[c] -= z[t]   (tableau_optimized.py:99-107)
      if ((r.xa | r.zb) == 0) return 0;
      r.xb = swar_add(Sd, r.xb, r.xa);
      r.za = swar_sub(Sd, r.za, r.zb);
      return CH_XB | CH_ZA;
    case SDIMB_OP_CNOT_INV:                                   // x[t] -= x[c];  z[c] += z[t]   (tableau_optimized.py:110-118)
      if ((r.xa | r.zb) == 0) return 0;
      r.xb = swar_sub(Sd, r.xb, r.xa);
      r.za = swar_add(Sd, r.za, r.zb);
      return CH_XB | CH_ZA;
    case SDIMB_OP_CZ:
    case SDIMB_OP_CZ_INV: {
      if ((r.xa | r.xb) == 0) return 0;
      // CZ = H^-1(t) CNOT(c,t) H(t) folded: z[a] +-= x[b]; z[b] +-= x[a]; phase +-= po*x[a]*x[b]
      uint32_t prod = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) prod |= (A.po * mod_d(A, byte_of(r.xa, k) * byte_of(r.xb, k))) << (8 * k);
      if (op == SDIMB_OP_CZ) {
        r.za = swar_add(Sd, r.za, r.xb); r.zb = swar_add(Sd, r.zb, r.xa);
        ph = swar_add(So, ph, prod);
      } else {
        r.za = swar_sub(Sd, r.za, r.xb); r.zb = swar_sub(Sd, r.zb, r.xa);
        ph = swar_sub(So, ph, prod);
      }
      return CH_ZA | CH_ZB;
    }
    case SDIMB_OP_SWAP: {
      const uint32_t x = r.xa, z = r.za;
      r.xa = r.xb; r.za = r.zb; r.xb = x; r.zb = z;
      return CH_XA | CH_ZA | CH_XB | CH_ZB;
    }
    default:
      return 0;
  }
}

// One gate on one lane word: math + stores of the words that changed; returns the new phase word.
__device__ __forceinline__ uint32_t gate_word(uint8_t* T, const KParams& p, int op, int a, int b, uint32_t pa,
                                              uint32_t pb, Rows r, int w, uint32_t ph) {
  const int wz = p.W / 4;
  uint32_t* rowa = reinterpret_cast<uint32_t*>(T + (int64_t)a * p.row_bytes);
  uint32_t* rowb = reinterpret_cast<uint32_t*>(T + (int64_t)b * p.row_bytes);
  const uint32_t ch = gate_math(p, op, pa, pb, r, ph);
  if (ch & CH_XA) rowa[w] = r.xa;
  if (ch & CH_ZA) rowa[wz + w] = r.za;
  if (ch & CH_XB) rowb[w] = r.xb;
  if (ch & CH_ZB) rowb[wz + w] = r.zb;
  return ph;
}

// One gate on FOUR consecutive lane words owned by one thread (128-bit loads and stores, dispatch and address
// arithmetic paid once for 16 lanes): the streaming shape of the lane interpreter for wide rows on the HBM store.
__device__ __forceinline__ void gate_vec4(uint8_t* T, const KParams& p, int op, int a, int b, uint32_t pa, uint32_t pb,
                                          int w0, uint32_t (&ph)[4]) {
  const int wz = p.W / 4;
  uint4* rowa = reinterpret_cast<uint4*>(T + (int64_t)a * p.row_bytes);
  uint4* rowb = reinterpret_cast<uint4*>(T + (int64_t)b * p.row_bytes);
  const bool two = is_two_qudit(op);
  const int v = w0 >> 2, vz = (wz + w0) >> 2;
  const uint4 xa4 = rowa[v], za4 = rowa[vz];
  uint4 xb4 = make_uint4(0, 0, 0, 0), zb4 = xb4;
  if (two) { xb4 = rowb[v]; zb4 = rowb[vz]; }
  uint32_t xa[4] = {xa4.x, xa4.y, xa4.z, xa4.w}, za[4] = {za4.x, za4.y, za4.z, za4.w};
  uint32_t xb[4] = {xb4.x, xb4.y, xb4.z, xb4.w}, zb[4] = {zb4.x, zb4.y, zb4.z, zb4.w};
  uint32_t ch = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    Rows r{xa[c], za[c], xb[c], zb[c]};
    ch |= gate_math(p, op, pa, pb, r, ph[c]);
    xa[c] = r.xa; za[c] = r.za; xb[c] = r.xb; zb[c] = r.zb;
  }
  if (ch & CH_XA) rowa[v] = make_uint4(xa[0], xa[1], xa[2], xa[3]);
  if (ch & CH_ZA) rowa[vz] = make_uint4(za[0], za[1], za[2], za[3]);
  if (ch & CH_XB) rowb[v] = make_uint4(xb[0], xb[1], xb[2], xb[3]);
  if (ch & CH_ZB) rowb[vz] = make_uint4(zb[0], zb[1], zb[2], zb[3]);
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

// The pieces of a measurement.  They take explicit row / lane ranges so that the one-CTA interpreter below (whole
// tableau) and the cluster interpreter in clusters.cuh (one slice of the rows per CTA) share them.

// Pivot: FIRST stabilizer with an X component on q (tableau_prime.py:273-283).  Thread-local candidate; the caller
// reduces it with block_min.
__device__ __forceinline__ uint32_t pivot_candidate(const uint8_t* rowq, const KParams& p) {
  const uint32_t* xq = reinterpret_cast<const uint32_t*>(rowq);
  for (int w = threadIdx.x; w < p.np / 4; w += blockDim.x) {
    const uint32_t x = xq[w];
    if (x) return 4u * w + ((__ffs(x) - 1) >> 3);
  }
  return kNoPivot;
}

// One walk down the pivot column AND the destabilizer-p column over rows [r_lo, r_hi) (kWalk rows' loads in flight
// per thread): xs/zs (exponentiate, tableau_prime.py:365-380, folded in), the support list `ar` (count cnt[0]), and
// the list `br` (count cnt[2]) of rows outside the support whose destabilizer-p entry is non-zero and must be
// cleared when the destabilizer is overwritten with the pivot.  Returns this thread's share of x_p . z_p.
__device__ __forceinline__ uint32_t column_walk(const uint8_t* T, const KParams& p, Scratch& S, uint32_t piv,
                                                uint32_t e, int r_lo, int r_hi) {
  const Arith& A = p.A;
  const int W = p.W, npad = p.np, nt = blockDim.x;
  uint32_t sd_raw = 0;
  for (int base = r_lo + threadIdx.x; base < r_hi; base += nt * kWalk) {
    uint32_t xr[kWalk], zr[kWalk], od[kWalk];
#pragma unroll
    for (int u = 0; u < kWalk; ++u) {
      const int r = base + u * nt;
      xr[u] = zr[u] = od[u] = 0;
      if (r < r_hi) {
        const uint8_t* row = T + (int64_t)r * p.row_bytes;
        xr[u] = row[piv]; zr[u] = row[W + piv];
        od[u] = (uint32_t)row[npad + piv] | row[W + npad + piv];
      }
    }
#pragma unroll
    for (int u = 0; u < kWalk; ++u) {
      const int r = base + u * nt;
      if (r >= r_hi) break;
      S.xs[r] = (uint8_t)mod_d(A, xr[u] * e);
      S.zs[r] = (uint8_t)mod_d(A, zr[u] * e);
      sd_raw += mod_d(A, xr[u] * zr[u]);
      if (xr[u] | zr[u]) S.ar[atomicAdd(&S.cnt[0], 1u)] = (uint16_t)r;
      else if (od[u]) S.br[atomicAdd(&S.cnt[2], 1u)] = (uint16_t)r;
    }
  }
  return sd_raw;
}

// Factors f = -X[q,i] mod d of every lane but the pivot, packed per lane word (fw), the active-word list `aw`
// (count cnt[1]) and zeroed dot accumulators of the active words.
__device__ __forceinline__ void factor_words(const uint8_t* rowq, const KParams& p, Scratch& S, uint32_t piv) {
  const Arith& A = p.A;
  for (int w = threadIdx.x; w < p.W / 4; w += blockDim.x) {
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
}

// col_i += f_i * col_p over all (active row, active word) pairs; pair id = ri * nw_a + wi.  Partial dot products
// Z[:,i] . x_p (old Z) of the rows in `ar` are added to S.dot.
__device__ __forceinline__ void rank1_update(uint8_t* T, const KParams& p, Scratch& S, int nr_a, int nw_a) {
  const Arith& A = p.A;
  const int nt = blockDim.x, tid = threadIdx.x, wz = p.W / 4;
  if (nw_a <= 0) return;
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

// New phase word of four generators:
// phase_i += f_i * phase_p + po * (f_i * (Z_i . x_p) + (x_p . z_p) * f_i(f_i-1)/2 * po)     (:310-312,317-319)
__device__ __forceinline__ uint32_t phase_word_update(const Arith& A, uint32_t fw, uint32_t ph, const uint32_t* dot4,
                                                      uint32_t sd, uint32_t ps) {
  uint32_t nph = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t fk = byte_of(fw, k);
    const uint32_t g = mod_d(A, (fk * (fk - 1u)) >> 1);
    const uint32_t cp = mod_d(A, mod_d(A, dot4[k]) * fk + sd * g * A.po);
    nph |= mod_o(A, byte_of(ph, k) + fk * ps + A.po * cp) << (8 * k);
  }
  return nph;
}

// destabilizer p <- old pivot, stabilizer p <- Z_q (tableau_prime.py:323-333) on the rows of the lists.
// Column accesses cost one DRAM sector per byte, so only entries that change are written: both lanes are
// rewritten on the pivot's support (list ar), stale destabilizer entries elsewhere (list br) are cleared.
__device__ __forceinline__ void column_writes(uint8_t* T, const KParams& p, Scratch& S, int q, uint32_t piv, int nr_a,
                                              int nr_b) {
  const int W = p.W, npad = p.np;
  for (int i = threadIdx.x; i < nr_a; i += blockDim.x) {
    const int r = S.ar[i];
    uint8_t* row = T + (int64_t)r * p.row_bytes;
    row[piv] = 0;
    row[W + piv] = (r == q) ? 1 : 0;
    row[npad + piv] = S.xs[r];
    row[W + npad + piv] = S.zs[r];
  }
  for (int i = threadIdx.x; i < nr_b; i += blockDim.x) {
    uint8_t* row = T + (int64_t)S.br[i] * p.row_bytes;
    row[npad + piv] = 0;
    row[W + npad + piv] = 0;
  }
}

// Deterministic branch, part 1: ordered compaction of the generators with a non-zero factor f_i = destab X[q,i]
// into S.ar / S.xs (order matters: the cross term uses the running ancilla, :354-357).  Returns the list length;
// `a1` receives this thread's share of sum_i f_i * phase_i over the generators whose phase word lies in
// [pw_lo, pw_hi).  Contains block barriers: call it from uniform code.
__device__ __forceinline__ int det_list(const uint8_t* rowq, const uint8_t* P8, const KParams& p, Scratch& S,
                                        int pw_lo, int pw_hi, uint32_t& a1) {
  const int n = p.n, npad = p.np, nt = blockDim.x, tid = threadIdx.x;
  int total = 0;
  a1 = 0;
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
      if ((i >> 2) >= pw_lo && (i >> 2) < pw_hi) a1 += f * P8[i];
    }
    total = all;
    __syncthreads();
  }
  return total;
}

// Deterministic branch, part 2: this thread's share, over rows [r_lo, r_hi), of
// sum_r ( ancilla_z . (f * x_i) + po * (x_i . z_i) * f(f-1)/2 ) accumulated over the list in order.
__device__ __forceinline__ uint32_t det_rows(const uint8_t* T, const KParams& p, const Scratch& S, int total,
                                             int r_lo, int r_hi) {
  const Arith& A = p.A;
  uint32_t part = 0;
  for (int r = r_lo + threadIdx.x; r < r_hi; r += blockDim.x) {
    const uint8_t* xr = T + (int64_t)r * p.row_bytes;
    const uint8_t* zr = xr + p.W;
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
  return part;
}

__device__ uint32_t measure(uint8_t* T, const KParams& p, Scratch& S, int q, int64_t slot, int64_t shot_local,
                            uint32_t draw) {
  const Arith& A = p.A;
  const int n = p.n, npad = p.np, nt = blockDim.x, tid = threadIdx.x;
  const int wz = p.W / 4;
  uint8_t* rowq = T + (int64_t)q * p.row_bytes;
  uint8_t* P8 = T + p.phase_off;
  if (tid < 3) S.cnt[tid] = 0;   // cnt[3] holds the live-op mask of the current batch
  __syncthreads();   // gate writes of other threads' lanes become visible; counters reset

  const uint32_t piv = block_min(pivot_candidate(rowq, p), S.red);

  uint32_t outcome, rec;
  if (piv != kNoPivot) {
    // -- random branch (tableau_prime.py:294-334), with exponentiate (:365-380) folded into the gather ----
    const uint32_t v = rowq[piv];
    const uint32_t e = S.inv[v];
    const uint32_t ps_old = P8[piv];
    uint32_t sd_raw = column_walk(T, p, S, piv, e, 0, n);
    factor_words(rowq, p, S, piv);
    sd_raw = mod_d(A, block_sum(sd_raw, S.red));              // barrier: publishes xs/zs/ar/fw/aw/dot/cnt
    const uint32_t ps = mod_o(A, ps_old * e + A.po * mod_d(A, sd_raw * mod_d(A, (e * (e - 1u)) >> 1)));
    const uint32_t sd = mod_d(A, mod_d(A, sd_raw * e) * e);   // x_p . z_p after exponentiation
    const int nr_a = (int)S.cnt[0], nw_a = (int)S.cnt[1];
    rank1_update(T, p, S, nr_a, nw_a);
    __syncthreads();
    for (int i = tid; i < nw_a; i += nt) {
      const int w = S.aw[i];
      uint32_t* Pw = reinterpret_cast<uint32_t*>(P8) + w;
      *Pw = phase_word_update(A, S.fw[w], *Pw, S.dot + 4 * w, sd, ps);
    }
    __syncthreads();
    column_writes(T, p, S, q, piv, nr_a, (int)S.cnt[2]);
    outcome = draw;
    if (tid == 0) {
      P8[npad + piv] = (uint8_t)ps;                           // destabilizer p <- old pivot (phase)
      P8[piv] = (uint8_t)mod_o(A, A.order - outcome * A.po);  // stabilizer p <- Z_q with phase -m*po
    }
    rec = outcome;
  } else {
    // -- deterministic branch (tableau_prime.py:336-363): ordered accumulation over generators ----------
    uint32_t a1;
    const int total = det_list(rowq, P8, p, S, 0, wz, a1);
    a1 = mod_o(A, block_sum(mod_o(A, a1), S.red));            // sum_i f_i * phase_i; publishes the lists
    const uint32_t part = mod_d(A, block_sum(det_rows(T, p, S, total, 0, n), S.red));
    const uint32_t ap = mod_o(A, a1 + A.po * part);
    // (-ap // po) % d with Python floor semantics (tableau_prime.py:362)
    outcome = (A.po == 1) ? neg_d(A, ap) : (((ap + 1u) >> 1) & 1u);
    rec = outcome | SDIMB_REC_DET;
  }
  if (tid == 0) p.records[shot_local * p.rec_stride + slot] = (uint8_t)rec;
  __syncthreads();
  return outcome;
}


// ---------------------------------------------------------------------------------------------
// The interpreter: one CTA per shot, grid-stride over shots.
// ---------------------------------------------------------------------------------------------
template <bool VEC4>
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
    constexpr bool vec4 = VEC4;                               // four words (16 lanes) per thread, 128-bit accesses
    const int w0 = vec4 ? 4 * (int)threadIdx.x : (int)threadIdx.x;
    const bool one_word = vec4 || p.W / 4 <= (int)blockDim.x;  // the thread's phase word(s) can live in registers
    const bool own_word = one_word && w0 < p.W / 4;
    uint32_t* Pw = reinterpret_cast<uint32_t*>(T + p.phase_off) + w0;
    uint32_t pw = (own_word && !vec4) ? *Pw : 0u;
    uint32_t pw4[4] = {0u, 0u, 0u, 0u};
    if (own_word && vec4) { const uint4 v = *reinterpret_cast<const uint4*>(Pw); pw4[0] = v.x; pw4[1] = v.y; pw4[2] = v.z; pw4[3] = v.w; }
    auto phases_out = [&]() {                                 // measurement reads and writes phases in memory
      if (!own_word) return;
      if (vec4) *reinterpret_cast<uint4*>(Pw) = make_uint4(pw4[0], pw4[1], pw4[2], pw4[3]);
      else *Pw = pw;
    };
    auto phases_in = [&]() {
      if (!own_word) return;
      if (vec4) { const uint4 v = *reinterpret_cast<const uint4*>(Pw); pw4[0] = v.x; pw4[1] = v.y; pw4[2] = v.z; pw4[3] = v.w; }
      else pw = *Pw;
    };
    auto one_gate = [&](int code, int qa, int qb, uint32_t ea, uint32_t eb) {
      if (!one_word) {
        gate_rows(T, p, make_int4(code, qa, qb, -1), ea, eb);
      } else if (own_word) {
        if (vec4) gate_vec4(T, p, code, qa, qb, ea, eb, w0, pw4);
        else pw = gate_word(T, p, code, qa, qb, ea, eb, load_rows(T, p, code, qa, qb, w0), w0, pw);
      }
    };

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
          one_gate(op.x, op.y, op.z, pa, pb);
          continue;
        }
        // collective ops: M, M_X, RESET
        if (op.x == SDIMB_OP_M_X) one_gate(SDIMB_OP_H_INV, op.y, -1, 0u, 0u);   // tableau_gates.py:292-296
        phases_out();
        const uint32_t m = measure(T, p, S, op.y, op.w, shot, (uint32_t)op.z);
        phases_in();
        if (op.x == SDIMB_OP_RESET && m) one_gate(SDIMB_OP_N1, op.y, -1, A.d - m, 0u);   // program.py:335-339
      }
    }
    phases_out();
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
__global__ void __launch_bounds__(kMaxThreads, 6) interp_kernel(const __grid_constant__ KParams p) { interp_body<false>(p); }
__global__ void __launch_bounds__(kWideThreads, 1) interp_kernel_wide(const __grid_constant__ KParams p) { interp_body<false>(p); }
// streaming shape for gate-dominated streams on the HBM store: 16 lanes per thread, 64 registers (32 one-warp CTAs
// per SM at n = 256)
__global__ void __launch_bounds__(kMaxThreads, 4) interp_kernel_stream(const __grid_constant__ KParams p) { interp_body<true>(p); }
