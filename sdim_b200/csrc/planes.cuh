// Bit-plane (bit-sliced) resident interpreter for d = 2 and d = 3.
//
// For the two smallest primes an exponent needs 1 (d = 2) or 2 (d = 3) bits, so a tableau row is kept as
// bit-planes over the generator lanes: one 32-bit word carries 32 generators, a Clifford gate on a qudit is a
// handful of LOP3s per word, and a whole n = 256 qutrit tableau is 64 KiB — it fits in shared memory, where one
// warp owns one shot for the entire circuit and HBM sees nothing but the record bytes.
//
//   d = 3  value v = 2*h + l, planes (l, h):  0 = (0,0), 1 = (1,0), 2 = (0,1);  negation swaps the planes
//   d = 2  value v = l;  phases are mod 4 = 2*h + l                         
//
// Shared-memory image of one shot.  Lane word j (32 generators) of a qudit row is ONE vector entry holding
// all planes of X and Z, so a gate is one LDS.128 + one STS.128 per lane (LDS.64 for d = 2):
//   d = 3: entry = uint4 (x_l, x_h, z_l, z_h)        d = 2: entry = uint2 (x, z)
//   row q:   entry[0..Wb) at q * (Wb + 1) entries    (+1 entry of padding: column walks — one entry per row,
//                                                     32 rows per instruction — then hit distinct banks)
//   phases:  uint2 (p_l, p_h) [Wb]                   after the n rows
// Lane numbering is that of the uint8 store (include/sdimb.h): stabilizer g -> lane g, destabilizer g -> lane
// np + g, with np a multiple of 32 here so the two halves never share a word.
//
// One CTA owns one shot.  With an unscheduled op stream the CTA is a single warp.  With a SCHEDULED stream
// (sdimb_schedule: layers of gates on pairwise disjoint qudits, separated by barriers) the CTA has
// SDIMB_SCHED_WARPS warps: inside a layer each warp executes its share of the gates — every warp adds its phase
// increments to a PRIVATE phase accumulator, so gates of one layer never touch the same word — and
// measurements split their column walks and row updates over all warps.  The accumulators are folded into
// accumulator 0 at the start of each measurement.
//
// Same reference behaviour as the uint8 interpreter in sdimb.cu (file:line citations there).
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

template <int D>
struct Geo {
  static constexpr int EW = (D == 2) ? 2 : 4;   // words per entry
  uint32_t* tab;                                // shared-memory image of the n rows
  uint2* ph_base;                               // [NW][Wb] phase accumulators, always in shared memory
  uint2* pacc;                                  // phase accumulator used by ldp/stp (this warp's, or #0 in measure)
  int n, np, Wb, RS;                            // RS = row stride in words = EW * (Wb + 1)
  int gpw, gsub, j0, jstep;                     // rank-1 update: row groups per warp (32 / Wb when Wb divides 32), this
                                                // lane's group inside the warp, its first lane word and its word stride
  __device__ __forceinline__ uint32_t* entry(int q, int j) const { return tab + q * RS + j * EW; }
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

__device__ __forceinline__ void cta_sync() {
  if (blockDim.x == 32) __syncwarp(); else __syncthreads();
}

// ---- gates: lane j of the warp owns lane word j of every row ------------------------------------------
template <int D>
__device__ __forceinline__ void g_h(const Geo<D>& G, int a, bool inverse) {
  for (int j = threadIdx.x & 31; j < G.Wb; j += 32) {
    const XZ v = G.ld(a, j);
    E p = G.ldp(j);
    if (D == 3) {
      p = add3(p, neg3(mul3(v.x, v.z)));                                  // phase -= x*z
      G.st(a, j, inverse ? XZ{v.z, neg3(v.x)} : XZ{neg3(v.z), v.x});     // H: (x,z)<-(-z,x); H^-1: (x,z)<-(z,-x)
    } else {
      p.h ^= v.x.l & v.z.l;                                               // phase += 2*x*z (mod 4); H == H^-1
      G.st(a, j, XZ{v.z, v.x});
    }
    G.stp(j, p);
  }
}

template <int D>
__device__ __forceinline__ void g_p(const Geo<D>& G, int a, bool inverse) {
  for (int j = threadIdx.x & 31; j < G.Wb; j += 32) {
    const XZ v = G.ld(a, j);
    E p = G.ldp(j);
    if (D == 3) {
      p = add3(p, inverse ? E{0u, v.x.h} : E{v.x.h, 0u});                 // phase +-= x(x-1)/2 = [x == 2]
      G.stz(a, j, add3(v.z, inverse ? neg3(v.x) : v.x));                  // z +-= x
    } else {
      if (inverse) { const uint32_t borrow = ~p.l & v.x.l; p.l ^= v.x.l; p.h ^= borrow; }   // phase -= x (mod 4)
      else { const uint32_t carry = p.l & v.x.l; p.l ^= v.x.l; p.h ^= carry; }              // phase += x^2 = x
      G.stz(a, j, E{v.z.l ^ v.x.l, 0u});
    }
    G.stp(j, p);
  }
}

// Pauli X^a Z^b on qudit q: phase += po*(b*x - a*z)
template <int D>
__device__ __forceinline__ void g_pauli(const Geo<D>& G, int q, uint32_t a, uint32_t b) {
  for (int j = threadIdx.x & 31; j < G.Wb; j += 32) {
    const XZ v = G.ld(q, j);
    E p = G.ldp(j);
    if (D == 3) p = add3(p, add3(smul3(v.x, b), smul3(v.z, (3u - a) % 3u)));
    else p.h ^= ((b & 1u) ? v.x.l : 0u) ^ ((a & 1u) ? v.z.l : 0u);
    G.stp(j, p);
  }
}

template <int D>
__device__ __forceinline__ void g_cnot(const Geo<D>& G, int a, int b, bool inverse) {
  for (int j = threadIdx.x & 31; j < G.Wb; j += 32) {
    const XZ va = G.ld(a, j), vb = G.ld(b, j);
    if (D == 3) {
      G.stx(b, j, add3(vb.x, inverse ? neg3(va.x) : va.x));               // x[t] +-= x[c]
      G.stz(a, j, add3(va.z, inverse ? vb.z : neg3(vb.z)));               // z[c] -+= z[t]
    } else {
      G.stx(b, j, E{vb.x.l ^ va.x.l, 0u});
      G.stz(a, j, E{va.z.l ^ vb.z.l, 0u});
    }
  }
}

template <int D>
__device__ __forceinline__ void g_cz(const Geo<D>& G, int a, int b, bool inverse) {
  for (int j = threadIdx.x & 31; j < G.Wb; j += 32) {
    const XZ va = G.ld(a, j), vb = G.ld(b, j);
    E p = G.ldp(j);
    if (D == 3) {
      const E prod = mul3(va.x, vb.x);
      p = add3(p, inverse ? neg3(prod) : prod);                           // phase +-= x[a]*x[b]
      G.stz(a, j, add3(va.z, inverse ? neg3(vb.x) : vb.x));
      G.stz(b, j, add3(vb.z, inverse ? neg3(va.x) : va.x));
    } else {
      p.h ^= va.x.l & vb.x.l;
      G.stz(a, j, E{va.z.l ^ vb.x.l, 0u});
      G.stz(b, j, E{vb.z.l ^ va.x.l, 0u});
    }
    G.stp(j, p);
  }
}

template <int D>
__device__ __forceinline__ void g_swap(const Geo<D>& G, int a, int b) {
  for (int j = threadIdx.x & 31; j < G.Wb; j += 32) {   // lane j swaps lane word j (lane ownership holds across gates)
    const XZ va = G.ld(a, j), vb = G.ld(b, j);
    G.st(a, j, vb);
    G.st(b, j, va);
  }
}

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
template <int D>
__device__ uint32_t p_measure(Geo<D> G, const KParams& p, PScratch& S, int q, int64_t slot, int64_t shot_local,
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
    cta_sync();
    for (int j = tid; j < Wb && nw > 1; j += nt) {
      E acc = G.ldp(j);
      for (int w = 1; w < nw; ++w) {
        uint2* pw = G.phase_of(w) + j;
        const E o{pw->x, pw->y};
        acc = (D == 3) ? add3(acc, o) : add4(acc, o);
        *pw = make_uint2(0u, 0u);
      }
      G.stp(j, acc);
    }
    cta_sync();
  }

  // pivot: first stabilizer lane with an X component on q (tableau_prime.py:273-283); every warp looks itself
  uint32_t best = kNoPivot;
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
    // one pass down the pivot column AND the destabilizer-p column: support list, values, stale destab entries
    uint32_t sd_part = 0;
    for (int base = 0; base < n; base += nt) {
      const int r = base + tid;
      uint32_t xr = 0, zr = 0, od = 0;
      if (r < n) {
        const XZ s = G.ld(r, jp), dd = G.ld(r, jd);
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
      uint32_t both = 0;                                 // list lengths packed: support | stale << 16
      if (lane == 0 && (ma | mb)) both = atomicAdd(&cnt[0], (uint32_t)__popc(ma) | ((uint32_t)__popc(mb) << 16));
      both = __shfl_sync(FULL, both, 0);
      if (act) S.ar[(both & 0xFFFFu) + __popc(ma & lt)] = (uint16_t)r;
      if (stale) S.br[(both >> 16) + __popc(mb & lt)] = (uint16_t)r;
    }
    sd_part = __reduce_add_sync(FULL, sd_part);
    if (lane == 0 && sd_part) atomicAdd(&cnt[2], sd_part);
    // factors f = -X[q,i] for every lane but the pivot itself
    for (int j = tid; j < Wb; j += nt) {
      E x = G.ld(q, j).x;
      if (j == jp) { x.l &= ~(1u << bp); x.h &= ~(1u << bp); }
      S.f[j] = (D == 3) ? make_uint2(x.h, x.l) : make_uint2(x.l, 0u);
    }
    cta_sync();
    const int nr_a = (int)(cnt[0] & 0xFFFFu), nr_b = (int)(cnt[0] >> 16);
    const uint32_t sd_raw = cnt[2] % D;
    const uint32_t ps = (ps_old * e + PO * ((sd_raw * ((e * (e - 1u)) >> 1)) % D)) % ORDER;
    const uint32_t sd = (sd_raw * e * e) % D;
    // col_i += f_i * col_p on the pivot's support.  Threads form row groups: 32/Wb per warp when Wb divides 32
    // (geometry precomputed in Geo: the divisions cost more than the update of a sparse measurement).
    const int gpw = G.gpw, gtot = gpw * nw, gid = warp * gpw + G.gsub, jstep = G.jstep;
    for (int j = G.j0; j < Wb; j += jstep) {
      const uint2 fv = S.f[j];
      const E f{fv.x, fv.y};
      E dot{0u, 0u};
      if (f.l | f.h) {
        for (int ri = gid; ri < nr_a; ri += gtot) {
          const int r = S.ar[ri];
          const uint32_t c = S.xz[r], s = c & 3u, t = c >> 2;
          const XZ v = G.ld(r, j);
          if (D == 3) {
            dot = add3(dot, smul3(v.z, s));                               // Z[:,i] . x_p  (old Z)
            G.st(r, j, XZ{add3(v.x, smul3(f, s)), add3(v.z, smul3(f, t))});
          } else {
            if (s) dot.l ^= v.z.l;
            G.st(r, j, XZ{E{v.x.l ^ (s ? f.l : 0u), 0u}, E{v.z.l ^ (t ? f.l : 0u), 0u}});
          }
        }
      }
      for (int off = Wb; off < 32 && gpw > 1; off <<= 1) {
        const E o{__shfl_xor_sync(FULL, dot.l, off), __shfl_xor_sync(FULL, dot.h, off)};
        dot = (D == 3) ? add3(dot, o) : E{dot.l ^ o.l, 0u};
      }
      if (gpw == 1 || lane < Wb) S.dotw[warp * Wb + j] = make_uint2(dot.l, dot.h);
    }
    cta_sync();
    // phase_i += f_i*ps + po*(f_i*dot_i + sd*f_i(f_i-1)/2*po)      (tableau_prime.py:310-312,317-319)
    for (int j = tid; j < Wb; j += nt) {
      const uint2 fv = S.f[j];
      const E f{fv.x, fv.y};
      if ((f.l | f.h) == 0) continue;
      E dot{0u, 0u};
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
      G.stp(j, ph);
    }
    cta_sync();
    // destabilizer p <- old pivot; stabilizer p <- Z_q with phase -m*po   (tableau_prime.py:323-333).
    // Only rows where something changes are touched: the support (list ar) and stale destabilizer rows (br).
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
    for (int i = tid; i < nr_b; i += nt) {
      const int r = S.br[i];
      XZ dd = G.ld(r, jd);
      dd.x = setbit2(dd.x, bp, 0u);
      dd.z = setbit2(dd.z, bp, 0u);
      G.st(r, jd, dd);
    }
    outcome = draw;        // replayed or Philox, resolved when the op was fetched (reference: random.choice, :332)
    if (tid == 0) G.setp(np + piv, ps);
    if (tid == 1) G.setp(piv, (ORDER - outcome * PO) % ORDER);
    rec = outcome;
  } else {
    // ---- deterministic branch (tableau_prime.py:336-363) ----
    // warp 0: ordered list of the generators with factor f_i = destab X[q,i] != 0, straight from the plane words
    if (warp == 0) {
      uint32_t a1 = 0;
      int total = 0;
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
    cta_sync();
    const int total = (int)cnt[3];
    uint32_t part = 0;
    for (int r = tid; r < n; r += nt) {
      uint32_t az = 0, cross = 0, sdg = 0;
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
    cta_sync();
    const uint32_t ap = (cnt[2] + PO * (cnt[0] % D)) % ORDER;
    outcome = (D == 3) ? (3u - ap) % 3u : (((ap + 1u) >> 1) & 1u);        // (-ap // po) % d  (tableau_prime.py:362)
    rec = outcome | SDIMB_REC_DET;
  }
  if (tid == 0) p.records[shot_local * p.rec_stride + slot] = (uint8_t)rec;
  cta_sync();
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
template <int D, bool GLOBAL>
__global__ void __launch_bounds__(32 * SDIMB_SCHED_WARPS, GLOBAL ? SDIMB_PLANES_GLOBAL_MIN_CTAS : 0)
interp_planes_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  Geo<D> G;
  G.n = p.n;
  G.np = (p.n + 31) / 32 * 32;
  G.Wb = 2 * G.np / 32;
  G.RS = Geo<D>::EW * (G.Wb + 1);
  G.gpw = (G.Wb <= 32 && (32 % G.Wb) == 0) ? 32 / G.Wb : 1;
  G.gsub = G.gpw > 1 ? lane / G.Wb : 0;
  G.j0 = G.gpw > 1 ? lane % G.Wb : lane;
  G.jstep = G.gpw > 1 ? G.Wb : 32;
  const int row_words = (p.n * G.RS + 3) & ~3;
  uint32_t* sm = reinterpret_cast<uint32_t*>(smem);
  if (GLOBAL) {
    G.tab = p.plane_slab + (int64_t)blockIdx.x * row_words;
  } else {
    G.tab = sm;
    sm += row_words;
  }
  G.ph_base = reinterpret_cast<uint2*>(sm);
  G.pacc = G.phase_of(warp);
  const int acc_words = (nw * 2 * G.Wb + 3) & ~3;
  PScratch S;
  S.ops = reinterpret_cast<int4*>(sm + acc_words);
  S.f = reinterpret_cast<uint2*>(S.ops + 32 * nw);
  S.dotw = S.f + G.Wb;
  S.cnt = reinterpret_cast<uint32_t*>(S.dotw + nw * G.Wb);
  S.ar = reinterpret_cast<uint16_t*>(S.cnt + 8);
  S.parity = 0;
  S.br = S.ar + G.np;
  S.xz = reinterpret_cast<uint8_t*>(S.br + G.np);
  S.next = reinterpret_cast<int*>(S.xz + G.np + ((4 - (G.np & 3)) & 3));
  int4* my_ops = S.ops + 32 * warp;

  for (int round = 0;; ++round) {
    // claim the next shot: dynamic when the caller provided a counter (shots differ in cost, CTAs in speed),
    // static grid-stride otherwise
    if (tid == 0)
      S.next[round & 1] = p.shot_counter ? (int)atomicAdd(p.shot_counter, 1u) : (int)(blockIdx.x + round * gridDim.x);
    cta_sync();
    const int64_t shot = S.next[round & 1];
    if (shot >= p.shots) break;
    // ---- load: |0...0> or pack from the uint8 store ----
    for (int i = tid; i < row_words; i += nt) G.tab[i] = 0u;
    for (int i = tid; i < acc_words; i += nt) reinterpret_cast<uint32_t*>(G.ph_base)[i] = 0u;
    if (tid < 8) S.cnt[tid] = 0;
    cta_sync();
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
    cta_sync();
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
      __syncwarp();
      my_ops[lane] = mine;
      __syncwarp();
#pragma unroll 1
      while (todo) {
        const int k = __ffs(todo) - 1;
        todo &= todo - 1;
        const int4 op = my_ops[k];
        switch (op.x) {
          case SDIMB_OP_X: g_pauli<D>(G, op.y, 1u, 0u); break;
          case SDIMB_OP_X_INV: g_pauli<D>(G, op.y, D - 1u, 0u); break;
          case SDIMB_OP_Z: g_pauli<D>(G, op.y, 0u, 1u); break;
          case SDIMB_OP_Z_INV: g_pauli<D>(G, op.y, 0u, D - 1u); break;
          case SDIMB_OP_H: g_h<D>(G, op.y, false); break;
          case SDIMB_OP_H_INV: g_h<D>(G, op.y, true); break;
          case SDIMB_OP_P: g_p<D>(G, op.y, false); break;
          case SDIMB_OP_P_INV: g_p<D>(G, op.y, true); break;
          case SDIMB_OP_CNOT: g_cnot<D>(G, op.y, op.z, false); break;
          case SDIMB_OP_CNOT_INV: g_cnot<D>(G, op.y, op.z, true); break;
          case SDIMB_OP_CZ: g_cz<D>(G, op.y, op.z, false); break;
          case SDIMB_OP_CZ_INV: g_cz<D>(G, op.y, op.z, true); break;
          case SDIMB_OP_SWAP: g_swap<D>(G, op.y, op.z); break;
          case SDIMB_OP_M_X:
            cta_sync();
            if (warp == 0) g_h<D>(G, op.y, true);
            dirty = true;                                                  // row q changed: the measurement must sync
            // fallthrough
          case SDIMB_OP_M:
          case SDIMB_OP_RESET: {
            const bool fold = dirty || (gate_pos & ((1u << k) - 1u)) != 0;   // gates since the last measurement?
            const uint32_t m = p_measure<D>(G, p, S, op.y, op.w, shot, fold, (uint32_t)op.z);
            dirty = false;
            if (op.x == SDIMB_OP_RESET) {
              if (m && warp == 0) g_pauli<D>(G, op.y, D - m, 0u);      // program.py:335-339
              cta_sync();
            }
            break;
          }
          case SDIMB_OP_N1: g_pauli<D>(G, op.y, (uint32_t)op.z & 0xFFu, (uint32_t)op.z >> 8); break;
          case SDIMB_OP_BARRIER: cta_sync(); break;
          default: break;
        }
      }
      dirty = dirty || gate_pos != 0;      // conservative: gates of this batch behind its last measurement
    }
    cta_sync();
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
      cta_sync();
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
    cta_sync();
  }
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

}  // namespace planes
