// Bit-plane (bit-sliced) interpreter for d = 2 and d = 3.
//
// For the two smallest primes an exponent needs 1 (d = 2) or 2 (d = 3) bits, so a tableau row is kept as
// bit-planes over the generator lanes: one 32-bit word carries 32 generators, a Clifford gate on a qudit is a
// handful of LOP3s per word, and a whole n = 256 qutrit tableau is 69.6 KB.  One CTA owns one shot for the entire
// circuit; the image lives in shared memory (GLOBAL = false) or in a per-CTA slab of caller scratch that stays in
// L2 / L1 (GLOBAL = true, what sdimb_plan picks when fewer than four images fit in one SM's shared memory).  HBM sees
// the op stream, the record bytes and the write-back of slab lines.
//
//   d = 3  value v = 2*h + l, planes (l, h):  0 = (0,0), 1 = (1,0), 2 = (0,1);  negation swaps the planes
//   d = 2  value v = l;  phases are mod 4 = 2*h + l                         
//
// Image of one shot.  Lane word j (32 generators) of a qudit row is ONE vector entry holding all planes of X and Z,
// so a gate is one 128-bit load + one 128-bit store per lane (64-bit for d = 2):
//   d = 3: entry = uint4 (x_l, x_h, z_l, z_h)        d = 2: entry = uint2 (x, z)
//   row q:   entry[0..Wb) at q * (Wb + 1) entries    (+1 entry of padding: shared-memory column walks — one entry per
//                                                     row, 32 rows per instruction — then hit distinct banks)
//   phases:  uint2 (p_l, p_h) [Wb]                   per-warp accumulators, always in shared memory
// Lane numbering is that of the uint8 store (include/sdimb.h): stabilizer g -> lane g, destabilizer g -> lane
// np + g, with np a multiple of 32 here so the two halves never share a word.
//
// With an unscheduled op stream the CTA is a single warp.  With a SCHEDULED stream (sdimb_schedule: layers of gates
// on pairwise disjoint qudits, separated by barriers) the CTA has SDIMB_SCHED_WARPS warps: inside a layer each warp
// executes its share of the gates — every warp adds its phase increments to a PRIVATE phase accumulator, so gates of
// one layer never touch the same word — and measurements split their column walks and row updates over all warps.
// The accumulators are folded into accumulator 0 at the start of each measurement.
//
// Gate handlers live in planes_gates.inc, included twice (specialised for the shared-memory instantiation, compact
// for the global-image one: see COMPACT below).  Same reference behaviour as the uint8 interpreter in sdimb.cu
// (file:line citations there).
#pragma once

namespace planes {

struct E {   // 32 lanes of one exponent (or phase): l = low plane, h = high plane
  uint32_t l, h;
};
struct XZ {  // one lane word of a qudit row
  E x, z;
};

// ---- GF(3), bit-sliced --------------------------------------------------------------------------
__device__ __forceinline__ E add3(E a, E b) {
  E c;
  c.l = (a.l & ~(b.l | b.h)) | (b.l & ~(a.l | a.h)) | (a.h & b.h);
  c.h = (a.h & ~(b.l | b.h)) | (b.h & ~(a.l | a.h)) | (a.l & b.l);
  return c;
}
__device__ __forceinline__ E neg3(E a) { return E{a.h, a.l}; }
__device__ __forceinline__ E mul3(E a, E b) { return E{(a.l & b.l) | (a.h & b.h), (a.l & b.h) | (a.h & b.l)}; }
__device__ __forceinline__ E smul3(E a, uint32_t s) {   // s in {0,1,2}, per-thread scalar
  const uint32_t m1 = (s == 1u) ? 0xFFFFFFFFu : 0u, m2 = (s == 2u) ? 0xFFFFFFFFu : 0u;
  return E{(a.l & m1) | (a.h & m2), (a.h & m1) | (a.l & m2)};
}
// ---- Z4 phases for d = 2 --------------------------------------------------------------------------
__device__ __forceinline__ E add4(E p, E a) {   // lane-wise p + a mod 4
  const uint32_t carry = p.l & a.l;
  return E{p.l ^ a.l, p.h ^ a.h ^ carry};
}
__device__ __forceinline__ uint32_t bit2(E v, int b) { return ((v.l >> b) & 1u) | (((v.h >> b) & 1u) << 1); }
__device__ __forceinline__ E setbit2(E w, int b, uint32_t v) {
  const uint32_t m = 1u << b;
  return E{(w.l & ~m) | ((v & 1u) << b), (w.h & ~m) | (((v >> 1) & 1u) << b)};
}

// IL (global image of large tableaus): the entries of a row are stored INTERLEAVED — stabilizer lane word j at entry
// 2j, destabilizer lane word j (lane word np/32 + j) at entry 2j + 1 — and rows carry no padding entry, so that the two
// entries a measurement's column walk and column writes touch per row (pivot lane p and its destabilizer lane np + p)
// share one 32-byte sector (d = 3; half a sector for d = 2).  Gates loop over all lane words and do not care.
template <int D, bool IL = false>
struct Geo {
  static constexpr int EW = (D == 2) ? 2 : 4;   // words per entry
  uint32_t* tab;                                // image of the n rows: shared memory, or this CTA's slab
  uint2* ph_base;                               // [NW][Wb] phase accumulators, always in shared memory
  uint2* pacc;                                  // phase accumulator used by ldp/stp (this warp's, or #0 in measure)
  int n, np, Wb, RS;                            // RS = row stride in words = EW * (Wb + 1)
  int gpw, gsub, j0, jstep;                     // rank-1 update: row groups per warp (32 / Wb when Wb divides 32), this
                                                // lane's group inside the warp, its first lane word and its word stride
  __device__ __forceinline__ uint32_t* entry(int q, int j) const {
    if (IL) { const int h = np >> 5; j = (j < h) ? 2 * j : 2 * (j - h) + 1; }
    return tab + q * RS + j * EW;
  }
  __device__ __forceinline__ XZ ld_at(const uint32_t* e) const {      // entry address computed by the caller
    if (D == 3) {
      const uint4 v = *reinterpret_cast<const uint4*>(e);
      return XZ{E{v.x, v.y}, E{v.z, v.w}};
    }
    const uint2 v = *reinterpret_cast<const uint2*>(e);
    return XZ{E{v.x, 0u}, E{v.y, 0u}};
  }
  __device__ __forceinline__ uint2* phase() const { return pacc; }
  __device__ __forceinline__ uint2* phase_of(int w) const { return ph_base + w * Wb; }
  __device__ __forceinline__ XZ ld(int q, int j) const {
    if (D == 3) {
      const uint4 v = *reinterpret_cast<const uint4*>(entry(q, j));
      return XZ{E{v.x, v.y}, E{v.z, v.w}};
    }
    const uint2 v = *reinterpret_cast<const uint2*>(entry(q, j));
    return XZ{E{v.x, 0u}, E{v.y, 0u}};
  }
  __device__ __forceinline__ void st(int q, int j, XZ v) const {
    if (D == 3) *reinterpret_cast<uint4*>(entry(q, j)) = make_uint4(v.x.l, v.x.h, v.z.l, v.z.h);
    else *reinterpret_cast<uint2*>(entry(q, j)) = make_uint2(v.x.l, v.z.l);
  }
  __device__ __forceinline__ void stx(int q, int j, E x) const {
    if (D == 3) *reinterpret_cast<uint2*>(entry(q, j)) = make_uint2(x.l, x.h);
    else entry(q, j)[0] = x.l;
  }
  __device__ __forceinline__ void stz(int q, int j, E z) const {
    if (D == 3) *reinterpret_cast<uint2*>(entry(q, j) + 2) = make_uint2(z.l, z.h);
    else entry(q, j)[1] = z.l;
  }
  __device__ __forceinline__ E ldp(int j) const { const uint2 v = phase()[j]; return E{v.x, v.y}; }
  __device__ __forceinline__ void stp(int j, E v) const { phase()[j] = make_uint2(v.l, v.h); }
  // scalar accessors
  __device__ __forceinline__ uint32_t getx(int q, int lane) const { return bit2(ld(q, lane >> 5).x, lane & 31); }
  __device__ __forceinline__ uint32_t getz(int q, int lane) const { return bit2(ld(q, lane >> 5).z, lane & 31); }
  __device__ __forceinline__ uint32_t getp(int lane) const { return bit2(ldp(lane >> 5), lane & 31); }
  __device__ __forceinline__ void setx(int q, int lane, uint32_t v) const {
    stx(q, lane >> 5, setbit2(ld(q, lane >> 5).x, lane & 31, v));
  }
  __device__ __forceinline__ void setz(int q, int lane, uint32_t v) const {
    stz(q, lane >> 5, setbit2(ld(q, lane >> 5).z, lane & 31, v));
  }
  __device__ __forceinline__ void setp(int lane, uint32_t v) const {
    stp(lane >> 5, setbit2(ldp(lane >> 5), lane & 31, v));
  }
};

struct PScratch {
  int4* ops;         // [NW][32] staged op batch, private to each warp
  uint2* f;          // [Wb] factor planes f = -X[q,i]
  uint2* dotw;       // [NW][Wb] per-warp partial dot products
  uint32_t* cnt;     // [2][4] list lengths / block accumulators, double-buffered by measurement parity
  uint32_t parity;   // which half the current measurement uses (uniform across the CTA)
  uint16_t* ar;      // [np] active rows (pivot support) / active generators (det branch)
  uint16_t* br;      // [np] rows whose destabilizer-p entry must be cleared
  uint8_t* xz;       // [np] pivot column: xs | zs << 2   (det branch: factor of active generator k)
  int* next;         // [2] shot claimed from the global counter (double-buffered)
};

template <bool FOUR_WARPS = false>   // FOUR_WARPS: the CTA is known to be multi-warp (global image), no test
__device__ __forceinline__ void cta_sync() {
  if (FOUR_WARPS) { __syncthreads(); return; }
  if (blockDim.x == 32) __syncwarp(); else __syncthreads();
}

// COMPACT (global image): the interpreter's hot code (dispatch, 13 gate handlers, measurement: ~30 KB) sits at the
// capacity of the 32 KB instruction-cache level, and the 32 warps of an SM are in different handlers at any time.  The
// compact form keeps one handler per gate FAMILY (direction and Pauli exponents become run-time selects), does not
// unroll the lane-word loops and drops the one-warp tests of the CTA barriers: 26 % less code, +6.5 % on the headline.
// The shared-memory instantiation keeps the specialised handlers (its SASS is unchanged).
#ifndef SDIMB_PG_COMPACT
#define SDIMB_PG_COMPACT 1
#endif
#ifndef SDIMB_PR_COMPACT        // the same dispatch form for the shared-memory instantiation: not measured yet (round 2)
#define SDIMB_PR_COMPACT 0
#endif
#define SDIMB_GATES_NS gates_std
#define SDIMB_GATE_LOOP
#include "planes_gates.inc"
#undef SDIMB_GATES_NS
#undef SDIMB_GATE_LOOP
#define SDIMB_GATES_NS gates_compact
#define SDIMB_GATE_LOOP _Pragma("unroll 1")
#include "planes_gates.inc"
#undef SDIMB_GATES_NS
#undef SDIMB_GATE_LOOP

// N1 event -> (a | b << 8), 0 if it does not fire: replayed, or Philox with the distribution of
// sdim/program.py:486-507.  Evaluated by the lane that fetched the op, so events that do not fire never reach
// the dispatch loop.
template <int D>
__device__ __forceinline__ uint32_t p_noise_event(const KParams& p, int64_t j, int64_t shot_local) {
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
      if (ch == 0) { const uint32_t v = 1u + __umulhi(r.y, D * D - 1u); a = v % D; b = v / D; }
      else { const uint32_t e = 1u + __umulhi(r.y, D - 1u); if (ch == 1) a = e; else b = e; }
    }
  }
  return a | (b << 8);
}

// ---- measurement (whole CTA: 1 or SDIMB_SCHED_WARPS warps) -----------------------------------------------------
#ifndef SDIMB_P_NOUNROLL        // measurement and shot-init loops left rolled (smaller code, both instantiations): not measured yet
#define SDIMB_P_NOUNROLL 0
#endif
#if SDIMB_P_NOUNROLL
#define SDIMB_P_LOOP _Pragma("unroll 1")
#else
#define SDIMB_P_LOOP
#endif
#ifndef SDIMB_PG_SKIPLIST       // column walk: skip the list bookkeeping of a warp none of whose 32 rows is listed
#define SDIMB_PG_SKIPLIST 0
#endif
// FW: the CTA is known to have four warps (global image), see cta_sync.
// MERGE (used with the interleaved image): the column writes of the random branch happen inside the rank-1 pass and the
// two pivot phases inside the phase pass — one pass over the support and one block barrier less per measurement.
template <int D, bool FW, bool MERGE, class GEO>
__device__ uint32_t p_measure(GEO G, const KParams& p, PScratch& S, int q, int64_t slot, int64_t shot_local,
                              bool fold, uint32_t draw) {
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  constexpr uint32_t PO = (D == 2) ? 2u : 1u, ORDER = D * PO;
  const int n = G.n, np = G.np, Wb = G.Wb;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  // `fold` = some gate ran since the last measurement: wait for every warp's gates, then fold the per-warp phase
  // accumulators into accumulator 0.  Back-to-back measurements skip both barriers (the previous measurement
  // ended with one and already reset the list counters).
  G.pacc = G.phase_of(0);
  // counters: this measurement uses one half, and clears the other half for the next one (whose first barrier-free
  // use is ordered behind this measurement's barriers)
  uint32_t* const cnt = S.cnt + 4 * S.parity;
  if (tid < 4) S.cnt[4 * (S.parity ^ 1u) + tid] = 0;
  S.parity ^= 1u;
  if (fold) {
    cta_sync<FW>();
    SDIMB_P_LOOP
    for (int j = tid; j < Wb && nw > 1; j += nt) {
      E acc = G.ldp(j);
      SDIMB_P_LOOP
      for (int w = 1; w < nw; ++w) {
        uint2* pw = G.phase_of(w) + j;
        const E o{pw->x, pw->y};
        acc = (D == 3) ? add3(acc, o) : add4(acc, o);
        *pw = make_uint2(0u, 0u);
      }
      G.stp(j, acc);
    }
    cta_sync<FW>();
  }

  // pivot: first stabilizer lane with an X component on q (tableau_prime.py:273-283); every warp looks itself
  uint32_t best = kNoPivot;
  SDIMB_P_LOOP
  for (int j = lane; j < np / 32; j += 32) {
    const E x = G.ld(q, j).x;
    const uint32_t m = x.l | x.h;
    if (m) { best = 32u * j + (__ffs(m) - 1); break; }
  }
  const uint32_t piv = __reduce_min_sync(FULL, best);

  uint32_t outcome, rec;
  if (piv != kNoPivot) {
    // ---- random branch (tableau_prime.py:294-334, exponentiate :365-380 folded in) ----
    const int jp = piv >> 5, bp = piv & 31, jd = np / 32 + jp;          // stab word / bit, destab word of lane p
    const uint32_t e = (D == 3) ? G.getx(q, piv) : 1u;                  // inverse of v mod 3 is v itself
    const uint32_t ps_old = G.getp(piv);
    if (MERGE && tid == 0) cnt[1] = draw;                               // only warp 0 resolved the draw: publish it
    // one pass down the pivot column AND the destabilizer-p column: support list, values, stale destab entries
    uint32_t sd_part = 0;
    // one row of the walk, entries already loaded (all lanes of the warp call it together: it ballots)
    auto walk_row = [&](const int r, const bool in, const XZ& s, const XZ& dd) {
      uint32_t xr = 0, zr = 0, od = 0;
      if (in) {
        const uint32_t bm = 1u << bp;
        od = (dd.x.l | dd.x.h | dd.z.l | dd.z.h) & bm;
        if ((s.x.l | s.x.h | s.z.l | s.z.h) & bm) {            // the pivot acts on few qudits: most rows stop here
          xr = bit2(s.x, bp); zr = bit2(s.z, bp);
          sd_part += xr * zr;
          if (D == 3 && e == 2u) { xr = (xr >> 1) | ((xr & 1u) << 1); zr = (zr >> 1) | ((zr & 1u) << 1); }   // * 2 = negate
          S.xz[r] = (uint8_t)(xr | (zr << 2));
        }
      }
      const bool act = (xr | zr) != 0, stale = !act && od != 0;
      const uint32_t ma = __ballot_sync(FULL, act), mb = __ballot_sync(FULL, stale);
#if SDIMB_PG_SKIPLIST
      if ((ma | mb) == 0) return;                        // warp-uniform: no row of these 32 is on either list (the usual case)
#endif
      uint32_t both = 0;                                 // list lengths packed: support | stale << 16
      if (lane == 0 && (ma | mb)) both = atomicAdd(&cnt[0], (uint32_t)__popc(ma) | ((uint32_t)__popc(mb) << 16));
      both = __shfl_sync(FULL, both, 0);
      if (act) S.ar[(both & 0xFFFFu) + __popc(ma & lt)] = (uint16_t)r;
      if (stale) S.br[(both >> 16) + __popc(mb & lt)] = (uint16_t)r;
    };
    const uint32_t* const col_s = G.entry(0, jp);        // entries of lane p / of its destabilizer in row 0: the walk
    const uint32_t* const col_d = G.entry(0, jd);        // strides by rows (interleaved image: one 32-byte sector)
    SDIMB_P_LOOP
    for (int base = 0; base < n; base += nt) {
      const int r = base + tid;
      const XZ zero{E{0u, 0u}, E{0u, 0u}};
      XZ s = zero, dd = zero;
      if (MERGE) { if (r < n) { s = G.ld_at(col_s + r * G.RS); dd = G.ld_at(col_d + r * G.RS); } }
      else if (r < n) { s = G.ld(r, jp); dd = G.ld(r, jd); }
      walk_row(r, r < n, s, dd);
    }
    sd_part = __reduce_add_sync(FULL, sd_part);
    if (lane == 0 && sd_part) atomicAdd(&cnt[2], sd_part);
    // factors f = -X[q,i] for every lane but the pivot itself
    SDIMB_P_LOOP
    for (int j = tid; j < Wb; j += nt) {
      E x = G.ld(q, j).x;
      if (j == jp) { x.l &= ~(1u << bp); x.h &= ~(1u << bp); }
      S.f[j] = (D == 3) ? make_uint2(x.h, x.l) : make_uint2(x.l, 0u);
    }
    cta_sync<FW>();
    const int nr_a = (int)(cnt[0] & 0xFFFFu), nr_b = (int)(cnt[0] >> 16);
    const uint32_t sd_raw = cnt[2] % D;
    const uint32_t ps = (ps_old * e + PO * ((sd_raw * ((e * (e - 1u)) >> 1)) % D)) % ORDER;
    const uint32_t sd = (sd_raw * e * e) % D;
    if (MERGE) draw = cnt[1];
    // col_i += f_i * col_p on the pivot's support.  Threads form row groups: 32/Wb per warp when Wb divides 32
    // (geometry precomputed in Geo: the divisions cost more than the update of a sparse measurement).
    const int gpw = G.gpw, gtot = gpw * nw, gid = warp * gpw + G.gsub, jstep = G.jstep;
    SDIMB_P_LOOP
    for (int j = G.j0; j < Wb; j += jstep) {
      const uint2 fv = S.f[j];
      const E f{fv.x, fv.y};
      E dot{0u, 0u};
      // MERGE: the words of lane p and of its destabilizer are always visited, and the column writes
      // (destabilizer p <- old pivot, stabilizer p <- Z_q, tableau_prime.py:323-333) ride on the same store
      if ((f.l | f.h) || (MERGE && (j == jp || j == jd))) {
        auto update_row = [&](const int r, const uint32_t c, const XZ& v) {
          const uint32_t s = c & 3u, t = c >> 2;
          XZ nv;
          if (D == 3) {
            dot = add3(dot, smul3(v.z, s));                               // Z[:,i] . x_p  (old Z)
            nv = XZ{add3(v.x, smul3(f, s)), add3(v.z, smul3(f, t))};
          } else {
            if (s) dot.l ^= v.z.l;
            nv = XZ{E{v.x.l ^ (s ? f.l : 0u), 0u}, E{v.z.l ^ (t ? f.l : 0u), 0u}};
          }
          if (MERGE) {
            if (j == jp) { nv.x = setbit2(nv.x, bp, 0u); nv.z = setbit2(nv.z, bp, (r == q) ? 1u : 0u); }
            else if (j == jd) { nv.x = setbit2(nv.x, bp, s); nv.z = setbit2(nv.z, bp, t); }
          }
          G.st(r, j, nv);
        };
        SDIMB_P_LOOP
        for (int ri = gid; ri < nr_a; ri += gtot) {
          const int r = S.ar[ri];
          update_row(r, S.xz[r], G.ld(r, j));
        }
      }
      SDIMB_P_LOOP
      for (int off = Wb; off < 32 && gpw > 1; off <<= 1) {
        const E o{__shfl_xor_sync(FULL, dot.l, off), __shfl_xor_sync(FULL, dot.h, off)};
        dot = (D == 3) ? add3(dot, o) : E{dot.l ^ o.l, 0u};
      }
      if (gpw == 1 || lane < Wb) S.dotw[warp * Wb + j] = make_uint2(dot.l, dot.h);
    }
    if (MERGE) {            // rows outside the support (untouched above): clear their stale destabilizer-p entry
      SDIMB_P_LOOP
      for (int i = tid; i < nr_b; i += nt) {
        const int r = S.br[i];
        XZ dd = G.ld(r, jd);
        dd.x = setbit2(dd.x, bp, 0u);
        dd.z = setbit2(dd.z, bp, 0u);
        G.st(r, jd, dd);
      }
    }
    cta_sync<FW>();
    // phase_i += f_i*ps + po*(f_i*dot_i + sd*f_i(f_i-1)/2*po)      (tableau_prime.py:310-312,317-319)
    // MERGE: and, in the same store, destabilizer p <- old pivot phase, stabilizer p <- -m*po
    SDIMB_P_LOOP
    for (int j = tid; j < Wb; j += nt) {
      const uint2 fv = S.f[j];
      const E f{fv.x, fv.y};
      if ((f.l | f.h) == 0 && !(MERGE && (j == jp || j == jd))) continue;
      E dot{0u, 0u};
      SDIMB_P_LOOP
      for (int w = 0; w < nw; ++w) {
        const uint2 o = S.dotw[w * Wb + j];
        dot = (D == 3) ? add3(dot, E{o.x, o.y}) : E{dot.l ^ o.x, 0u};
      }
      E ph = G.ldp(j);
      if (D == 3) {
        E t = add3(smul3(f, ps), mul3(dot, f));
        t = add3(t, smul3(E{f.h, 0u}, sd));                               // f(f-1)/2 = [f == 2]
        ph = add3(ph, t);
      } else {
        ph = add4(ph, E{(ps & 1u) ? f.l : 0u, (ps & 2u) ? f.l : 0u});
        ph.h ^= dot.l & f.l;
      }
      if (MERGE) {
        if (j == jd) ph = setbit2(ph, bp, ps);
        if (j == jp) ph = setbit2(ph, bp, (ORDER - draw * PO) % ORDER);
      }
      G.stp(j, ph);
    }
    if (!MERGE) {
    cta_sync<FW>();
    // destabilizer p <- old pivot; stabilizer p <- Z_q with phase -m*po   (tableau_prime.py:323-333).
    // Only rows where something changes are touched: the support (list ar) and stale destabilizer rows (br).
    SDIMB_P_LOOP
    for (int i = tid; i < nr_a; i += nt) {
      const int r = S.ar[i];
      const uint32_t c = S.xz[r];
      XZ s = G.ld(r, jp);
      s.x = setbit2(s.x, bp, 0u);
      s.z = setbit2(s.z, bp, (r == q) ? 1u : 0u);
      G.st(r, jp, s);
      XZ dd = G.ld(r, jd);
      dd.x = setbit2(dd.x, bp, c & 3u);
      dd.z = setbit2(dd.z, bp, c >> 2);
      G.st(r, jd, dd);
    }
    SDIMB_P_LOOP
    for (int i = tid; i < nr_b; i += nt) {
      const int r = S.br[i];
      XZ dd = G.ld(r, jd);
      dd.x = setbit2(dd.x, bp, 0u);
      dd.z = setbit2(dd.z, bp, 0u);
      G.st(r, jd, dd);
    }
    }
    outcome = draw;        // replayed or Philox, resolved when the op was fetched (reference: random.choice, :332)
    if (!MERGE) {
      if (tid == 0) G.setp(np + piv, ps);
      if (tid == 1) G.setp(piv, (ORDER - outcome * PO) % ORDER);
    }
    rec = outcome;
  } else {
    // ---- deterministic branch (tableau_prime.py:336-363) ----
    // warp 0: ordered list of the generators with factor f_i = destab X[q,i] != 0, straight from the plane words
    if (warp == 0) {
      uint32_t a1 = 0;
      int total = 0;
      SDIMB_P_LOOP
      for (int base = 0; base < np / 32; base += 32) {
        const int j = base + lane;
        E f{0u, 0u}, ph{0u, 0u};
        if (j < np / 32) { f = G.ld(q, np / 32 + j).x; ph = G.ldp(j); }
        uint32_t m = f.l | f.h;
        const int cnt = __popc(m);
        int off = cnt;                                                    // inclusive warp scan of the counts
        for (int d2 = 1; d2 < 32; d2 <<= 1) {
          const int o = __shfl_up_sync(FULL, off, d2);
          if (lane >= d2) off += o;
        }
        int pos = total + off - cnt;
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          const uint32_t fv = bit2(f, b);
          S.ar[pos] = (uint16_t)(32 * j + b);
          S.xz[pos] = (uint8_t)fv;
          a1 += fv * bit2(ph, b);
          ++pos;
        }
        total += __shfl_sync(FULL, off, 31);
      }
      a1 = __reduce_add_sync(FULL, a1);
      if (lane == 0) { cnt[3] = (uint32_t)total; cnt[2] = a1 % ORDER; }
    }
    cta_sync<FW>();
    const int total = (int)cnt[3];
    uint32_t part = 0;
    SDIMB_P_LOOP
    for (int r = tid; r < n; r += nt) {
      uint32_t az = 0, cross = 0, sdg = 0;
      SDIMB_P_LOOP
      for (int k = 0; k < total; ++k) {
        const int g = S.ar[k];
        const uint32_t f = S.xz[k];
        const XZ v = G.ld(r, g >> 5);
        if (((v.x.l | v.x.h | v.z.l | v.z.h) >> (g & 31) & 1u) == 0) continue;   // generator g is the identity on row r
        const uint32_t xi = bit2(v.x, g & 31), zi = bit2(v.z, g & 31);
        cross += (f * xi) * az;                                           // ancilla_z . (f * x_i), running ancilla
        az = (az + f * zi) % D;
        sdg += xi * zi * ((f * (f - 1u)) >> 1);
        if ((k & 15) == 15) { cross %= D; sdg %= D; }
      }
      part += (cross + PO * sdg) % D;
    }
    part = __reduce_add_sync(FULL, part);
    if (lane == 0 && part) atomicAdd(&cnt[0], part);
    cta_sync<FW>();
    const uint32_t ap = (cnt[2] + PO * (cnt[0] % D)) % ORDER;
    outcome = (D == 3) ? (3u - ap) % 3u : (((ap + 1u) >> 1) & 1u);        // (-ap // po) % d  (tableau_prime.py:362)
    rec = outcome | SDIMB_REC_DET;
  }
  if (tid == 0) p.records[shot_local * p.rec_stride + slot] = (uint8_t)rec;
  cta_sync<FW>();
  return outcome;
}

// ---- the interpreter: one CTA per shot (1 warp, or SDIMB_SCHED_WARPS warps on a scheduled stream) --------------
// GLOBAL = false: the row image lives in shared memory (the resident interpreter).  GLOBAL = true: it lives in a
// per-CTA slab of caller-provided scratch (L2 / HBM) — the same bit planes for tableaus beyond the shared-memory
// limit (d = 3: n > ~440, d = 2: n > ~630), 4x / 8x fewer bytes per gate than the uint8 lanes; phase accumulators and
// measurement scratch stay in shared memory.  Two instantiations, so the resident one keeps LDS / STS.
// The global-image variant is bounded by latency, not by shared memory: 8 CTAs (32 warps) per SM at 64 registers
// measured best (7: -1 %, 9: -1 %, 10: -17 %, the default 80 registers / 6 CTAs: -12 %).  The resident variant keeps
// the compiler's own register choice (0 = no minimum: small tableaus run a dozen CTAs per SM).
#ifndef SDIMB_PLANES_GLOBAL_MIN_CTAS
#define SDIMB_PLANES_GLOBAL_MIN_CTAS 8
#endif
#ifndef SDIMB_PG_NOPAD          // global image: rows without the bank-conflict padding entry
#define SDIMB_PG_NOPAD 0
#endif
#ifndef SDIMB_PG_HOSTGEO        // geometry and shared-memory offsets as kernel parameters (both instantiations)
#define SDIMB_PG_HOSTGEO 0
#endif
#ifndef SDIMB_PG_SHFLOPS        // global image: staged ops broadcast by shuffle instead of a shared-memory round trip
#define SDIMB_PG_SHFLOPS 1
#endif
// IL: global image with interleaved entries and merged measurement passes (Geo, p_measure MERGE) — what sdimb_run
// launches for tableaus of SDIMB_PG_IL_MIN_NP padded qudits and more (round-2 A/B, profiles/r2_ab_measurement.json:
// d = 3 n = 500 18.0 -> 13.3 ms, d = 2 n = 400 19.6 -> 19.0 ms, but d = 3 n = 256 19.8 -> 20.8 ms, so the headline
// shape keeps the plain image).
#ifndef SDIMB_PG_IL_MIN_NP
#define SDIMB_PG_IL_MIN_NP 384
#endif
// GATES_ONLY: the stream holds no M / M_X / RESET (its measurements all sit in the tail run that run_tail_kernel
// executes): the measurement code is not instantiated at all — a third less code for the instruction caches.
#ifndef SDIMB_PG_GATES_ONLY_CTAS
#define SDIMB_PG_GATES_ONLY_CTAS 12   // 40 registers: 12 CTAs per SM measured best (8: +4 %, 10: +1 %, 6: +10 %)
#endif
template <int D, bool GLOBAL, bool IL = false, bool GATES_ONLY = false>
__global__ void __launch_bounds__(32 * SDIMB_SCHED_WARPS, GLOBAL ? (GATES_ONLY ? SDIMB_PG_GATES_ONLY_CTAS : SDIMB_PLANES_GLOBAL_MIN_CTAS) : 0)
interp_planes_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  static_assert(GLOBAL || !IL, "the interleaved image exists for the global placement only");
  constexpr bool FW = GLOBAL && SDIMB_PG_COMPACT != 0;     // the CTA is known to have four warps: barriers without the test
  constexpr bool CP = FW || (!GLOBAL && SDIMB_PR_COMPACT != 0);   // compact dispatch form (see COMPACT above)
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  Geo<D, IL> G;
  PScratch S;
  uint32_t* sm = reinterpret_cast<uint32_t*>(smem);
#if SDIMB_PG_HOSTGEO
  // geometry and scratch offsets come from the host as kernel parameters: constant-bank operands, nothing for the
  // compiler to rematerialise inside the dispatch loop (the arithmetic below was 11 % of the executed instructions)
  G.n = p.n;
  G.np = p.pg.np;
  G.Wb = p.pg.Wb;
  G.RS = p.pg.RS;
  G.gpw = p.pg.gpw;
  G.jstep = p.pg.jstep;
  G.gsub = G.gpw > 1 ? lane / G.Wb : 0;
  G.j0 = G.gpw > 1 ? lane % G.Wb : lane;
  const int row_words = p.pg.row_words, acc_words = p.pg.acc_words;
  if (GLOBAL) {
    G.tab = p.plane_slab + (int64_t)blockIdx.x * p.pg.slab_words;
  } else {
    G.tab = sm;
    sm += row_words;
  }
  G.ph_base = reinterpret_cast<uint2*>(sm);
  G.pacc = G.phase_of(warp);
  uint8_t* const sb = reinterpret_cast<uint8_t*>(sm);
  S.ops = reinterpret_cast<int4*>(sb + p.pg.off_ops);
  S.f = reinterpret_cast<uint2*>(sb + p.pg.off_f);
  S.dotw = reinterpret_cast<uint2*>(sb + p.pg.off_dotw);
  S.cnt = reinterpret_cast<uint32_t*>(sb + p.pg.off_cnt);
  S.ar = reinterpret_cast<uint16_t*>(sb + p.pg.off_ar);
  S.parity = 0;
  S.br = reinterpret_cast<uint16_t*>(sb + p.pg.off_br);
  S.xz = sb + p.pg.off_xz;
  S.next = reinterpret_cast<int*>(sb + p.pg.off_next);
#else
  G.n = p.n;
  G.np = (p.n + 31) / 32 * 32;
  G.Wb = 2 * G.np / 32;
  // the padding entry keeps shared-memory column walks off one bank; a global image may drop it so that rows start on
  // sector (n = 256, d = 3: cache-line) boundaries — its slab keeps the padded size either way
  G.RS = Geo<D>::EW * (G.Wb + ((IL || (GLOBAL && SDIMB_PG_NOPAD != 0)) ? 0 : 1));
  G.gpw = (G.Wb <= 32 && (32 % G.Wb) == 0) ? 32 / G.Wb : 1;
  G.gsub = G.gpw > 1 ? lane / G.Wb : 0;
  G.j0 = G.gpw > 1 ? lane % G.Wb : lane;
  G.jstep = G.gpw > 1 ? G.Wb : 32;
  const int row_words = (p.n * G.RS + 3) & ~3;
  if (GLOBAL) {
    const int slab_words = (p.n * Geo<D>::EW * (G.Wb + 1) + 3) & ~3;       // = planes_row_bytes / 4, the host's stride
    G.tab = p.plane_slab + (int64_t)blockIdx.x * slab_words;
  } else {
    G.tab = sm;
    sm += row_words;
  }
  G.ph_base = reinterpret_cast<uint2*>(sm);
  G.pacc = G.phase_of(warp);
  const int acc_words = (nw * 2 * G.Wb + 3) & ~3;
  S.ops = reinterpret_cast<int4*>(sm + acc_words);
  S.f = reinterpret_cast<uint2*>(S.ops + 32 * nw);
  S.dotw = S.f + G.Wb;
  S.cnt = reinterpret_cast<uint32_t*>(S.dotw + nw * G.Wb);
  S.ar = reinterpret_cast<uint16_t*>(S.cnt + 8);
  S.parity = 0;
  S.br = S.ar + G.np;
  S.xz = reinterpret_cast<uint8_t*>(S.br + G.np);
  S.next = reinterpret_cast<int*>(S.xz + G.np + ((4 - (G.np & 3)) & 3));
#endif
  int4* const my_ops = S.ops + 32 * warp;

  for (int round = 0;; ++round) {
    // claim the next shot: dynamic when the caller provided a counter (shots differ in cost, CTAs in speed),
    // static grid-stride otherwise
    if (tid == 0)
      S.next[round & 1] = p.shot_counter ? (int)atomicAdd(p.shot_counter, 1u) : (int)(blockIdx.x + round * gridDim.x);
    cta_sync<FW>();
    const int64_t shot = S.next[round & 1];
    if (shot >= p.shots) break;
    if (GLOBAL && p.img_per_shot) G.tab = p.plane_slab + shot * p.img_stride_words;   // the image outlives this kernel
    // ---- load: |0...0> or pack from the uint8 store ----
    SDIMB_P_LOOP
    for (int i = tid; i < row_words; i += nt) G.tab[i] = 0u;
    SDIMB_P_LOOP
    for (int i = tid; i < acc_words; i += nt) reinterpret_cast<uint32_t*>(G.ph_base)[i] = 0u;
    if (tid < 8) S.cnt[tid] = 0;
    cta_sync<FW>();
    uint8_t* T8 = p.tab ? p.tab + shot * p.shot_bytes : nullptr;
    G.pacc = G.phase_of(0);
    if (p.flags & SDIMB_FRESH) {
      for (int q = tid; q < p.n; q += nt) {
        G.setz(q, q, 1u);              // stabilizer q = Z_q
        G.setx(q, G.np + q, 1u);       // destabilizer q = X_q
      }
    } else {
      for (int q = warp; q < p.n; q += nw) {
        const uint8_t* row8 = T8 + (int64_t)q * p.row_bytes;
        for (int j = lane; j < G.Wb; j += 32) {
          XZ v{E{0u, 0u}, E{0u, 0u}};
          for (int b = 0; b < 32; ++b) {
            const int ln = 32 * j + b;
            const int half = ln >= G.np, g = half ? ln - G.np : ln;
            if (g >= p.n) continue;
            const uint32_t xv = row8[half * p.np + g], zv = row8[p.W + half * p.np + g];
            v.x.l |= (xv & 1u) << b; v.x.h |= ((xv >> 1) & 1u) << b;
            v.z.l |= (zv & 1u) << b; v.z.h |= ((zv >> 1) & 1u) << b;
          }
          G.st(q, j, v);
        }
      }
      for (int j = tid; j < G.Wb; j += nt) {
        E ph{0u, 0u};
        for (int b = 0; b < 32; ++b) {
          const int ln = 32 * j + b;
          const int half = ln >= G.np, g = half ? ln - G.np : ln;
          if (g >= p.n) continue;
          const uint32_t v = T8[p.phase_off + half * p.np + g];
          ph.l |= (v & 1u) << b; ph.h |= ((v >> 1) & 1u) << b;
        }
        G.stp(j, ph);
      }
    }
    G.pacc = G.phase_of(warp);
    cta_sync<FW>();
    bool dirty = false;                    // some gate may have added to a private phase accumulator since the last fold

    int4 ahead = make_int4(SDIMB_OP_I, 0, 0, 0);                // ops of the next batch, fetched one batch early
    if (lane < p.n_ops) ahead = __ldg(p.ops + lane);
    for (int64_t i0 = 0; i0 < p.n_ops; i0 += 32) {
      // each warp fetches the same 32 ops, one per lane, and keeps the ones it has to execute: collective ops
      // (measurements, barriers) and its own share of the gates; the fetching lane resolves N1 events, so events
      // that do not fire are never dispatched.  The load of the following batch is issued now and lands while
      // this batch executes.
      int4 mine = ahead;
      ahead = make_int4(SDIMB_OP_I, 0, 0, 0);
      if (i0 + 32 + lane < p.n_ops) ahead = __ldg(p.ops + i0 + 32 + lane);
      const int owner = (mine.x >> SDIMB_OP_WARP_SHIFT) & 0xFF;
      mine.x &= SDIMB_OP_MASK;
      const bool collective = mine.x >= SDIMB_OP_M && mine.x != SDIMB_OP_N1;      // M, M_X, RESET, BARRIER
      bool live = mine.x != SDIMB_OP_I && (collective || nw == 1 || owner == warp);
      if (live && mine.x == SDIMB_OP_N1) {
        mine.z = (int)p_noise_event<D>(p, mine.w, shot);
        live = mine.z != 0;
      }
      if (collective && mine.x != SDIMB_OP_BARRIER && warp == 0) {   // outcome this measurement takes if it is
        // random; only warp 0 consumes it (new stabilizer phase, record, RESET correction)
        if (p.replay_meas) {
          mine.z = p.replay_meas[shot * p.n_meas + mine.w];
        } else {
          const uint64_t gshot = (uint64_t)(p.shot_offset + shot);
          const uint4 r = philox4x32((uint32_t)gshot, (uint32_t)(gshot >> 32), (uint32_t)mine.w, 0u, (uint32_t)p.seed,
                                     (uint32_t)(p.seed >> 32));
          mine.z = (int)__umulhi(r.x, (uint32_t)D);
        }
      }
      if (mine.x == SDIMB_OP_BARRIER && nw == 1) live = false;
      uint32_t todo = __ballot_sync(0xFFFFFFFFu, live);
      // positions of ops (executed by ANY warp) that may touch a phase accumulator: identical in every warp
      const uint32_t gate_pos = __ballot_sync(0xFFFFFFFFu, mine.x != SDIMB_OP_I && !collective);
      if constexpr (!(GLOBAL && SDIMB_PG_SHFLOPS != 0)) {
        __syncwarp();
        my_ops[lane] = mine;
        __syncwarp();
      }
#pragma unroll 1
      while (todo) {
        const int k = __ffs(todo) - 1;
        todo &= todo - 1;
        int4 op;
        if constexpr (GLOBAL && SDIMB_PG_SHFLOPS != 0) {      // the warp is converged here: broadcast from the fetching lane
          op = make_int4(__shfl_sync(0xFFFFFFFFu, mine.x, k), __shfl_sync(0xFFFFFFFFu, mine.y, k),
                         __shfl_sync(0xFFFFFFFFu, mine.z, k), __shfl_sync(0xFFFFFFFFu, mine.w, k));
        } else {
          op = my_ops[k];
        }
        // the measurement cases, shared by the two dispatch forms below
#define SDIMB_COLLECTIVE_CASES(NS)                                                                                      \
          case SDIMB_OP_M_X:                                                                                            \
            cta_sync<FW>();                                                                                             \
            if (warp == 0) NS::g_h<D>(G, op.y, true);                                                                   \
            dirty = true;                       /* row q changed: the measurement must sync */                         \
            /* fallthrough */                                                                                           \
          case SDIMB_OP_M:                                                                                              \
          case SDIMB_OP_RESET: {                                                                                        \
            const bool fold = dirty || (gate_pos & ((1u << k) - 1u)) != 0;   /* gates since the last measurement? */   \
            const uint32_t m = p_measure<D, FW, IL>(G, p, S, op.y, op.w, shot, fold, (uint32_t)op.z);              \
            dirty = false;                                                                                              \
            if (op.x == SDIMB_OP_RESET) {                                                                               \
              if (m && warp == 0) NS::g_pauli<D>(G, op.y, D - m, 0u);   /* program.py:335-339 */                     \
              cta_sync<FW>();                                                                                           \
            }                                                                                                           \
            break;                                                                                                      \
          }
        if constexpr (CP) {
          switch (op.x) {
          case SDIMB_OP_X: case SDIMB_OP_X_INV: case SDIMB_OP_Z: case SDIMB_OP_Z_INV: case SDIMB_OP_N1: {
            // X^a Z^b: X (1,0), X^-1 (d-1,0), Z (0,1), Z^-1 (0,d-1), N1 the event resolved at fetch
            const uint32_t e = (op.x == SDIMB_OP_X || op.x == SDIMB_OP_Z) ? 1u : D - 1u;
            uint32_t pa = (op.x <= SDIMB_OP_X_INV) ? e : 0u, pb = (op.x <= SDIMB_OP_X_INV) ? 0u : e;
            if (op.x == SDIMB_OP_N1) { pa = (uint32_t)op.z & 0xFFu; pb = (uint32_t)op.z >> 8; }
            gates_compact::g_pauli<D>(G, op.y, pa, pb);
            break;
          }
          case SDIMB_OP_H: case SDIMB_OP_H_INV: gates_compact::g_h<D>(G, op.y, op.x == SDIMB_OP_H_INV); break;
          case SDIMB_OP_P: case SDIMB_OP_P_INV: gates_compact::g_p<D>(G, op.y, op.x == SDIMB_OP_P_INV); break;
          case SDIMB_OP_CNOT: case SDIMB_OP_CNOT_INV: gates_compact::g_cnot<D>(G, op.y, op.z, op.x == SDIMB_OP_CNOT_INV); break;
          case SDIMB_OP_CZ: case SDIMB_OP_CZ_INV: gates_compact::g_cz<D>(G, op.y, op.z, op.x == SDIMB_OP_CZ_INV); break;
          case SDIMB_OP_SWAP: gates_compact::g_swap<D>(G, op.y, op.z); break;
          case SDIMB_OP_M_X: case SDIMB_OP_M: case SDIMB_OP_RESET:
            if constexpr (!GATES_ONLY) {
              switch (op.x) {
              SDIMB_COLLECTIVE_CASES(gates_compact)
              default: break;
              }
            }
            break;
          case SDIMB_OP_BARRIER: cta_sync<FW>(); break;
          default: break;
          }
        } else {
          switch (op.x) {
          case SDIMB_OP_X: gates_std::g_pauli<D>(G, op.y, 1u, 0u); break;
          case SDIMB_OP_X_INV: gates_std::g_pauli<D>(G, op.y, D - 1u, 0u); break;
          case SDIMB_OP_Z: gates_std::g_pauli<D>(G, op.y, 0u, 1u); break;
          case SDIMB_OP_Z_INV: gates_std::g_pauli<D>(G, op.y, 0u, D - 1u); break;
          case SDIMB_OP_H: gates_std::g_h<D>(G, op.y, false); break;
          case SDIMB_OP_H_INV: gates_std::g_h<D>(G, op.y, true); break;
          case SDIMB_OP_P: gates_std::g_p<D>(G, op.y, false); break;
          case SDIMB_OP_P_INV: gates_std::g_p<D>(G, op.y, true); break;
          case SDIMB_OP_CNOT: gates_std::g_cnot<D>(G, op.y, op.z, false); break;
          case SDIMB_OP_CNOT_INV: gates_std::g_cnot<D>(G, op.y, op.z, true); break;
          case SDIMB_OP_CZ: gates_std::g_cz<D>(G, op.y, op.z, false); break;
          case SDIMB_OP_CZ_INV: gates_std::g_cz<D>(G, op.y, op.z, true); break;
          case SDIMB_OP_SWAP: gates_std::g_swap<D>(G, op.y, op.z); break;
          SDIMB_COLLECTIVE_CASES(gates_std)
          case SDIMB_OP_N1: gates_std::g_pauli<D>(G, op.y, (uint32_t)op.z & 0xFFu, (uint32_t)op.z >> 8); break;
          case SDIMB_OP_BARRIER: cta_sync<FW>(); break;
          default: break;
          }
        }
#undef SDIMB_COLLECTIVE_CASES
      }
      dirty = dirty || gate_pos != 0;      // conservative: gates of this batch behind its last measurement
    }
    cta_sync<FW>();
    if (GLOBAL && p.img_per_shot) {       // folded phase planes behind the rows, for run_tail_kernel
      uint2* const out = reinterpret_cast<uint2*>(G.tab + row_words);
      for (int j = tid; j < G.Wb; j += nt) {
        uint2 o = G.phase_of(0)[j];
        E acc{o.x, o.y};
        for (int w = 1; w < nw; ++w) {
          o = G.phase_of(w)[j];
          acc = (D == 3) ? add3(acc, E{o.x, o.y}) : add4(acc, E{o.x, o.y});
        }
        out[j] = make_uint2(acc.l, acc.h);
      }
    }
    if (p.flags & SDIMB_WRITEBACK) {      // fold the accumulators, then unpack into the uint8 store
      G.pacc = G.phase_of(0);
      for (int j = tid; j < G.Wb && nw > 1; j += nt) {
        E acc = G.ldp(j);
        for (int w = 1; w < nw; ++w) {
          const uint2 o = G.phase_of(w)[j];
          acc = (D == 3) ? add3(acc, E{o.x, o.y}) : add4(acc, E{o.x, o.y});
        }
        G.stp(j, acc);
      }
      cta_sync<FW>();
      for (int q = warp; q < p.n; q += nw) {
        uint8_t* row8 = T8 + (int64_t)q * p.row_bytes;
        for (int ln = lane; ln < p.W; ln += 32) {
          const int half = ln >= p.np, g = half ? ln - p.np : ln;
          const bool live = g < p.n;
          row8[ln] = live ? (uint8_t)G.getx(q, half * G.np + g) : 0;
          row8[p.W + ln] = live ? (uint8_t)G.getz(q, half * G.np + g) : 0;
        }
      }
      for (int ln = tid; ln < p.W; ln += nt) {
        const int half = ln >= p.np, g = half ? ln - p.np : ln;
        T8[p.phase_off + ln] = (g < p.n) ? (uint8_t)G.getp(half * G.np + g) : 0;
      }
      G.pacc = G.phase_of(warp);
    }
    cta_sync<FW>();
  }
}

// what the kernel prologue computes from (n, d, warps per CTA), for KParams::pg
inline PlaneGeo make_plane_geo(int n, int d, int nw, bool global) {
  PlaneGeo g;
  const int EW = (d == 2) ? 2 : 4;
  g.np = (n + 31) / 32 * 32;
  g.Wb = 2 * g.np / 32;
  g.RS = EW * (g.Wb + ((global && SDIMB_PG_NOPAD != 0) ? 0 : 1));
  g.gpw = (g.Wb <= 32 && (32 % g.Wb) == 0) ? 32 / g.Wb : 1;
  g.jstep = g.gpw > 1 ? g.Wb : 32;
  g.row_words = (n * g.RS + 3) & ~3;
  g.slab_words = (n * EW * (g.Wb + 1) + 3) & ~3;
  g.acc_words = (nw * 2 * g.Wb + 3) & ~3;
  g.off_ops = 4 * g.acc_words;
  g.off_f = g.off_ops + 16 * 32 * nw;
  g.off_dotw = g.off_f + 8 * g.Wb;
  g.off_cnt = g.off_dotw + 8 * nw * g.Wb;
  g.off_ar = g.off_cnt + 4 * 8;
  g.off_br = g.off_ar + 2 * g.np;
  g.off_xz = g.off_br + 2 * g.np;
  g.off_next = g.off_xz + g.np + ((4 - (g.np & 3)) & 3);
  return g;
}

// bytes of the row image of one shot (shared memory of a resident CTA, or one global slab of an overflow CTA)
inline size_t planes_row_bytes(int n, int d) {
  const size_t EW = (d == 2) ? 2 : 4;
  const size_t np = (size_t)(n + 31) / 32 * 32, Wb = 2 * np / 32, RS = EW * (Wb + 1);
  return 4 * (((size_t)n * RS + 3) & ~(size_t)3);
}
// shared memory besides the rows: phase accumulators + scratch
inline size_t planes_scratch_bytes(int n, int nw) {
  const size_t np = (size_t)(n + 31) / 32 * 32, Wb = 2 * np / 32;
  const size_t acc_words = ((size_t)nw * 2 * Wb + 3) & ~(size_t)3;
  return 4 * acc_words + (size_t)nw * 32 * 16 + 8 * Wb + (size_t)nw * 8 * Wb + 32 + 2 * np + 2 * np + np + 4 + 8 + 16;
}
inline size_t planes_smem_bytes(int n, int d, int nw = SDIMB_SCHED_WARPS) {
  return planes_row_bytes(n, d) + planes_scratch_bytes(n, nw);
}
// words between the per-shot images of a run that hands its tail to run_tail_kernel: rows, then Wb phase entries
inline size_t planes_img_stride_words(int n, int d) {
  const size_t np = (size_t)(n + 31) / 32 * 32, Wb = 2 * np / 32;
  return (planes_row_bytes(n, d) / 4 + 2 * Wb + 7) & ~(size_t)7;
}

#include "planes_gm.cuh"
#include "planes_stream.cuh"
#include "planes_tile.cuh"

}  // namespace planes
