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
// Shared-memory image of one shot (uint32 words, Wb = W/32 words per plane, B = bits per exponent):
//   row q:   X planes [B][Wb], Z planes [B][Wb]            at q * 2*B*Wb
//   phases:  P_l [Wb], P_h [Wb]                            at n * 2*B*Wb
// Lane numbering is that of the uint8 store (include/sdimb.h): stabilizer g -> lane g, destabilizer g -> lane
// np + g, with np a multiple of 32 here so the two halves never share a word.
//
// Same reference behaviour as the uint8 interpreter in sdimb.cu (file:line citations there).
#pragma once

namespace planes {

struct E {   // 32 lanes of one exponent (or phase): l = low plane, h = high plane
  uint32_t l, h;
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

template <int D>
struct Geo {
  static constexpr int B = (D == 2) ? 1 : 2;
  uint32_t* tab;   // shared-memory image
  int n, np, Wb, RW;
  __device__ __forceinline__ uint32_t* row(int q) const { return tab + q * RW; }
  __device__ __forceinline__ uint32_t* phase() const { return tab + n * RW; }
  __device__ __forceinline__ E ldx(int q, int j) const {
    const uint32_t* r = row(q);
    return E{r[j], D == 3 ? r[Wb + j] : 0u};
  }
  __device__ __forceinline__ E ldz(int q, int j) const {
    const uint32_t* r = row(q) + B * Wb;
    return E{r[j], D == 3 ? r[Wb + j] : 0u};
  }
  __device__ __forceinline__ void stx(int q, int j, E v) const {
    uint32_t* r = row(q);
    r[j] = v.l;
    if (D == 3) r[Wb + j] = v.h;
  }
  __device__ __forceinline__ void stz(int q, int j, E v) const {
    uint32_t* r = row(q) + B * Wb;
    r[j] = v.l;
    if (D == 3) r[Wb + j] = v.h;
  }
  __device__ __forceinline__ E ldp(int j) const { return E{phase()[j], phase()[Wb + j]}; }
  __device__ __forceinline__ void stp(int j, E v) const { phase()[j] = v.l; phase()[Wb + j] = v.h; }
  // scalar accessors (column operations)
  __device__ __forceinline__ uint32_t getx(int q, int lane) const {
    const E v = ldx(q, lane >> 5);
    return ((v.l >> (lane & 31)) & 1u) | (((v.h >> (lane & 31)) & 1u) << 1);
  }
  __device__ __forceinline__ uint32_t getz(int q, int lane) const {
    const E v = ldz(q, lane >> 5);
    return ((v.l >> (lane & 31)) & 1u) | (((v.h >> (lane & 31)) & 1u) << 1);
  }
  __device__ __forceinline__ uint32_t getp(int lane) const {
    const E v = ldp(lane >> 5);
    return ((v.l >> (lane & 31)) & 1u) | (((v.h >> (lane & 31)) & 1u) << 1);
  }
  __device__ __forceinline__ void setx(int q, int lane, uint32_t v) const {
    E w = ldx(q, lane >> 5);
    const uint32_t bit = 1u << (lane & 31);
    w.l = (w.l & ~bit) | ((v & 1u) ? bit : 0u);
    w.h = (w.h & ~bit) | ((v & 2u) ? bit : 0u);
    stx(q, lane >> 5, w);
  }
  __device__ __forceinline__ void setz(int q, int lane, uint32_t v) const {
    E w = ldz(q, lane >> 5);
    const uint32_t bit = 1u << (lane & 31);
    w.l = (w.l & ~bit) | ((v & 1u) ? bit : 0u);
    w.h = (w.h & ~bit) | ((v & 2u) ? bit : 0u);
    stz(q, lane >> 5, w);
  }
  __device__ __forceinline__ void setp(int lane, uint32_t v) const {
    E w = ldp(lane >> 5);
    const uint32_t bit = 1u << (lane & 31);
    w.l = (w.l & ~bit) | ((v & 1u) ? bit : 0u);
    w.h = (w.h & ~bit) | ((v & 2u) ? bit : 0u);
    stp(lane >> 5, w);
  }
};

struct PScratch {
  uint32_t* fl;      // [Wb] factor planes f = -X[q,i]
  uint32_t* fh;      // [Wb]
  int4* ops;         // [32] staged op batch
  uint16_t* ar;      // [np] active rows / active generators
  uint8_t* xs;       // [np]
  uint8_t* zs;       // [np]
};

// ---- gates: lane j of the warp owns word j of every plane -------------------------------------------
template <int D>
__device__ __forceinline__ void g_h(const Geo<D>& G, int a, bool inverse) {
  for (int j = threadIdx.x; j < G.Wb; j += 32) {
    const E x = G.ldx(a, j), z = G.ldz(a, j);
    if ((x.l | x.h | z.l | z.h) == 0) continue;
    E p = G.ldp(j);
    if (D == 3) {
      p = add3(p, neg3(mul3(x, z)));                       // phase -= x*z
      G.stx(a, j, inverse ? z : neg3(z));                  // H: (x,z) <- (-z,x); H^-1: (x,z) <- (z,-x)
      G.stz(a, j, inverse ? neg3(x) : x);
    } else {
      p.h ^= x.l & z.l;                                    // phase += 2*x*z (mod 4); H == H^-1 on qubits
      G.stx(a, j, z);
      G.stz(a, j, x);
    }
    G.stp(j, p);
  }
}

template <int D>
__device__ __forceinline__ void g_p(const Geo<D>& G, int a, bool inverse) {
  for (int j = threadIdx.x; j < G.Wb; j += 32) {
    const E x = G.ldx(a, j);
    if ((x.l | x.h) == 0) continue;
    const E z = G.ldz(a, j);
    E p = G.ldp(j);
    if (D == 3) {
      // phase +-= x(x-1)/2 = [x == 2];  z +-= x
      p = add3(p, inverse ? E{0u, x.h} : E{x.h, 0u});
      G.stz(a, j, add3(z, inverse ? neg3(x) : x));
    } else {
      // phase +-= x^2 = x (mod 4);  z ^= x
      if (inverse) { const uint32_t borrow = ~p.l & x.l; p.l ^= x.l; p.h ^= borrow; }
      else { const uint32_t carry = p.l & x.l; p.l ^= x.l; p.h ^= carry; }
      G.stz(a, j, E{z.l ^ x.l, 0u});
    }
    G.stp(j, p);
  }
}

// Pauli X^a Z^b on qudit q: phase += po*(b*x - a*z)
template <int D>
__device__ __forceinline__ void g_pauli(const Geo<D>& G, int q, uint32_t a, uint32_t b) {
  for (int j = threadIdx.x; j < G.Wb; j += 32) {
    E p = G.ldp(j);
    if (D == 3) {
      E t{0u, 0u};
      if (b) t = smul3(G.ldx(q, j), b);
      if (a) t = add3(t, smul3(G.ldz(q, j), 3u - a));
      if ((t.l | t.h) == 0) continue;
      p = add3(p, t);
    } else {
      uint32_t t = 0;
      if (b & 1u) t ^= G.ldx(q, j).l;
      if (a & 1u) t ^= G.ldz(q, j).l;
      if (!t) continue;
      p.h ^= t;
    }
    G.stp(j, p);
  }
}

template <int D>
__device__ __forceinline__ void g_cnot(const Geo<D>& G, int a, int b, bool inverse) {
  for (int j = threadIdx.x; j < G.Wb; j += 32) {
    const E xa = G.ldx(a, j), zb = G.ldz(b, j);
    if ((xa.l | xa.h | zb.l | zb.h) == 0) continue;
    const E xb = G.ldx(b, j), za = G.ldz(a, j);
    if (D == 3) {
      G.stx(b, j, add3(xb, inverse ? neg3(xa) : xa));      // x[t] +-= x[c]
      G.stz(a, j, add3(za, inverse ? zb : neg3(zb)));      // z[c] -+= z[t]
    } else {
      G.stx(b, j, E{xb.l ^ xa.l, 0u});
      G.stz(a, j, E{za.l ^ zb.l, 0u});
    }
  }
}

template <int D>
__device__ __forceinline__ void g_cz(const Geo<D>& G, int a, int b, bool inverse) {
  for (int j = threadIdx.x; j < G.Wb; j += 32) {
    const E xa = G.ldx(a, j), xb = G.ldx(b, j);
    if ((xa.l | xa.h | xb.l | xb.h) == 0) continue;
    const E za = G.ldz(a, j), zb = G.ldz(b, j);
    E p = G.ldp(j);
    if (D == 3) {
      const E prod = mul3(xa, xb);
      p = add3(p, inverse ? neg3(prod) : prod);            // phase +-= x[a]*x[b]
      G.stz(a, j, add3(za, inverse ? neg3(xb) : xb));
      G.stz(b, j, add3(zb, inverse ? neg3(xa) : xa));
    } else {
      p.h ^= xa.l & xb.l;
      G.stz(a, j, E{za.l ^ xb.l, 0u});
      G.stz(b, j, E{zb.l ^ xa.l, 0u});
    }
    G.stp(j, p);
  }
}

template <int D>
__device__ __forceinline__ void g_swap(const Geo<D>& G, int a, int b) {
  for (int j = threadIdx.x; j < G.Wb; j += 32) {   // lane j swaps word j of every plane (lane ownership)
    const E xa = G.ldx(a, j), za = G.ldz(a, j), xb = G.ldx(b, j), zb = G.ldz(b, j);
    G.stx(a, j, xb); G.stz(a, j, zb);
    G.stx(b, j, xa); G.stz(b, j, za);
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

// ---- measurement (one warp) ---------------------------------------------------------------------------
template <int D>
__device__ uint32_t p_measure(const Geo<D>& G, const KParams& p, PScratch& S, int q, int64_t slot,
                              int64_t shot_local) {
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  constexpr uint32_t PO = (D == 2) ? 2u : 1u, ORDER = D * PO;
  const int n = G.n, np = G.np, Wb = G.Wb, lane = threadIdx.x;
  __syncwarp();

  // pivot: first stabilizer lane with an X component on q (tableau_prime.py:273-283)
  uint32_t best = kNoPivot;
  for (int j = lane; j < np / 32; j += 32) {
    const E x = G.ldx(q, j);
    const uint32_t m = x.l | x.h;
    if (m) { best = 32u * j + (__ffs(m) - 1); break; }
  }
  const uint32_t piv = __reduce_min_sync(FULL, best);

  uint32_t draw;
  if (p.replay_meas) {
    draw = p.replay_meas[shot_local * p.n_meas + slot];
  } else {
    const uint64_t gshot = (uint64_t)(p.shot_offset + shot_local);
    const uint4 r = philox4x32((uint32_t)gshot, (uint32_t)(gshot >> 32), (uint32_t)slot, 0u, (uint32_t)p.seed,
                               (uint32_t)(p.seed >> 32));
    draw = __umulhi(r.x, (uint32_t)D);
  }

  uint32_t outcome, rec;
  if (piv != kNoPivot) {
    // ---- random branch (tableau_prime.py:294-334, exponentiate :365-380 folded in) ----
    const uint32_t e = (D == 3) ? G.getx(q, piv) : 1u;      // inverse of v mod 3 is v itself
    const uint32_t ps_old = G.getp(piv);
    uint32_t sd_raw = 0;
    int nr_a = 0;
    for (int base = 0; base < n; base += 32) {
      const int r = base + lane;
      uint32_t xr = 0, zr = 0;
      if (r < n) {
        xr = G.getx(r, piv); zr = G.getz(r, piv);
        S.xs[r] = (uint8_t)((xr * e) % D);
        S.zs[r] = (uint8_t)((zr * e) % D);
        sd_raw += xr * zr;
      }
      const uint32_t mask = __ballot_sync(FULL, (xr | zr) != 0);
      if (xr | zr) S.ar[nr_a + __popc(mask & ((1u << lane) - 1u))] = (uint16_t)r;
      nr_a += __popc(mask);
    }
    sd_raw = __reduce_add_sync(FULL, sd_raw) % D;
    const uint32_t ps = (ps_old * e + PO * ((sd_raw * ((e * (e - 1u)) >> 1)) % D)) % ORDER;
    const uint32_t sd = (sd_raw * e * e) % D;
    // factors f = -X[q,i] for every lane but the pivot itself
    for (int j = lane; j < Wb; j += 32) {
      E x = G.ldx(q, j);
      if (j == (int)(piv >> 5)) { x.l &= ~(1u << (piv & 31)); x.h &= ~(1u << (piv & 31)); }
      S.fl[j] = (D == 3) ? x.h : x.l;
      S.fh[j] = (D == 3) ? x.l : 0u;
    }
    __syncwarp();
    // col_i += f_i * col_p on the pivot's support; lanes split into 32/Wb row groups when Wb divides 32
    const int groups = (Wb <= 32 && (32 % Wb) == 0) ? 32 / Wb : 1;
    const int jstep = groups > 1 ? Wb : 32;
    const int grp = groups > 1 ? lane / Wb : 0;
    for (int j = groups > 1 ? lane % Wb : lane; j < Wb; j += jstep) {
      const E f{S.fl[j], S.fh[j]};
      E dot{0u, 0u};
      if (f.l | f.h) {
        for (int ri = grp; ri < nr_a; ri += groups) {
          const int r = S.ar[ri];
          const uint32_t s = S.xs[r], t = S.zs[r];
          const E x = G.ldx(r, j), z = G.ldz(r, j);
          if (D == 3) {
            dot = add3(dot, smul3(z, s));                   // Z[:,i] . x_p  (old Z)
            if (s) G.stx(r, j, add3(x, smul3(f, s)));
            if (t) G.stz(r, j, add3(z, smul3(f, t)));
          } else {
            if (s) { dot.l ^= z.l; G.stx(r, j, E{x.l ^ f.l, 0u}); }
            if (t) G.stz(r, j, E{z.l ^ f.l, 0u});
          }
        }
      }
      for (int off = Wb; off < 32 && groups > 1; off <<= 1) {
        const E o{__shfl_xor_sync(FULL, dot.l, off), __shfl_xor_sync(FULL, dot.h, off)};
        dot = (D == 3) ? add3(dot, o) : E{dot.l ^ o.l, 0u};
      }
      if (grp == 0 && (f.l | f.h)) {
        // phase_i += f_i*ps + po*(f_i*dot_i + sd*f_i(f_i-1)/2*po)      (tableau_prime.py:310-312,317-319)
        E ph = G.ldp(j);
        if (D == 3) {
          E t = add3(smul3(f, ps), mul3(dot, f));
          t = add3(t, smul3(E{f.h, 0u}, sd));               // f(f-1)/2 = [f == 2]
          ph = add3(ph, t);
        } else {
          ph = add4(ph, E{(ps & 1u) ? f.l : 0u, (ps & 2u) ? f.l : 0u});
          ph.h ^= dot.l & f.l;
        }
        G.stp(j, ph);
      }
    }
    __syncwarp();
    // destabilizer p <- old pivot; stabilizer p <- Z_q with phase -m*po   (tableau_prime.py:323-333)
    for (int i = lane; i < nr_a; i += 32) {
      const int r = S.ar[i];
      G.setx(r, piv, 0u);
      G.setz(r, piv, 0u);
    }
    for (int r = lane; r < n; r += 32) {
      G.setx(r, np + piv, S.xs[r]);
      G.setz(r, np + piv, S.zs[r]);
    }
    __syncwarp();
    outcome = draw;
    if (lane == 0) {
      G.setz(q, piv, 1u);
      G.setp(np + piv, ps);
      G.setp(piv, (ORDER - outcome * PO) % ORDER);
    }
    rec = outcome;
  } else {
    // ---- deterministic branch (tableau_prime.py:336-363) ----
    uint32_t a1 = 0;
    int total = 0;
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      const uint32_t f = (i < n) ? G.getx(q, np + i) : 0u;
      const uint32_t mask = __ballot_sync(FULL, f != 0);
      if (f) {
        const int pos = total + __popc(mask & ((1u << lane) - 1u));
        S.ar[pos] = (uint16_t)i;
        S.xs[pos] = (uint8_t)f;
        a1 += f * G.getp(i);
      }
      total += __popc(mask);
    }
    a1 = __reduce_add_sync(FULL, a1) % ORDER;
    __syncwarp();
    uint32_t part = 0;
    for (int r = lane; r < n; r += 32) {
      uint32_t az = 0, cross = 0, sdg = 0;
      for (int k = 0; k < total; ++k) {
        const int g = S.ar[k];
        const uint32_t f = S.xs[k];
        const uint32_t xi = G.getx(r, g), zi = G.getz(r, g);
        cross += (f * xi) * az;                              // ancilla_z . (f * x_i), running ancilla
        az = (az + f * zi) % D;
        sdg += xi * zi * ((f * (f - 1u)) >> 1);
        if ((k & 15) == 15) { cross %= D; sdg %= D; }
      }
      part += (cross + PO * sdg) % D;
    }
    part = __reduce_add_sync(FULL, part) % D;
    const uint32_t ap = (a1 + PO * part) % ORDER;
    outcome = (D == 3) ? (3u - ap) % 3u : (((ap + 1u) >> 1) & 1u);   // (-ap // po) % d  (tableau_prime.py:362)
    rec = outcome | SDIMB_REC_DET;
  }
  if (lane == 0) p.records[shot_local * p.rec_stride + slot] = (uint8_t)rec;
  __syncwarp();
  return outcome;
}

// ---- the interpreter: one warp (= one CTA) per shot ---------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(32) interp_planes_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  Geo<D> G;
  G.n = p.n;
  G.np = (p.n + 31) / 32 * 32;
  G.Wb = 2 * G.np / 32;
  G.RW = 2 * Geo<D>::B * G.Wb;
  G.tab = reinterpret_cast<uint32_t*>(smem);
  const int tab_words = p.n * G.RW + 2 * G.Wb;
  PScratch S;
  S.fl = G.tab + tab_words;
  S.fh = S.fl + G.Wb;
  S.ops = reinterpret_cast<int4*>(S.fh + G.Wb + ((4 - ((tab_words + 2 * G.Wb) & 3)) & 3));
  S.ar = reinterpret_cast<uint16_t*>(S.ops + 32);
  S.xs = reinterpret_cast<uint8_t*>(S.ar + G.np);
  S.zs = S.xs + G.np;
  const int lane = threadIdx.x;

  for (int64_t shot = blockIdx.x; shot < p.shots; shot += gridDim.x) {
    // ---- load: |0...0> or pack from the uint8 store ----
    for (int i = lane; i < tab_words; i += 32) G.tab[i] = 0u;
    __syncwarp();
    uint8_t* T8 = p.tab ? p.tab + shot * p.shot_bytes : nullptr;
    if (p.flags & SDIMB_FRESH) {
      for (int q = lane; q < p.n; q += 32) {
        G.setz(q, q, 1u);              // stabilizer q = Z_q
        G.setx(q, G.np + q, 1u);       // destabilizer q = X_q
      }
    } else {
      for (int q = 0; q < p.n; ++q) {
        const uint8_t* row8 = T8 + (int64_t)q * p.row_bytes;
        for (int j = lane; j < G.Wb; j += 32) {
          E x{0u, 0u}, z{0u, 0u};
          for (int b = 0; b < 32; ++b) {
            const int ln = 32 * j + b;
            const int half = ln >= G.np, g = half ? ln - G.np : ln;
            if (g >= p.n) continue;
            const uint32_t xv = row8[half * p.np + g], zv = row8[p.W + half * p.np + g];
            x.l |= (xv & 1u) << b; x.h |= ((xv >> 1) & 1u) << b;
            z.l |= (zv & 1u) << b; z.h |= ((zv >> 1) & 1u) << b;
          }
          G.stx(q, j, x); G.stz(q, j, z);
        }
      }
      for (int j = lane; j < G.Wb; j += 32) {
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
    __syncwarp();
    for (int64_t i0 = 0; i0 < p.n_ops; i0 += 32) {
      // fetch 32 ops, one per lane; the fetching lane resolves N1 events, so only live ops are dispatched
      int4 mine = make_int4(SDIMB_OP_I, 0, 0, 0);
      if (i0 + lane < p.n_ops) mine = __ldg(p.ops + i0 + lane);
      bool live = mine.x != SDIMB_OP_I;
      if (mine.x == SDIMB_OP_N1) {
        mine.z = (int)p_noise_event<D>(p, mine.w, shot);
        live = mine.z != 0;
      }
      uint32_t todo = __ballot_sync(0xFFFFFFFFu, live);
      __syncwarp();
      S.ops[lane] = mine;
      __syncwarp();
#pragma unroll 1
      while (todo) {
        const int k = __ffs(todo) - 1;
        todo &= todo - 1;
        const int4 op = S.ops[k];
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
            g_h<D>(G, op.y, true);
            // fallthrough
          case SDIMB_OP_M:
          case SDIMB_OP_RESET: {
            const uint32_t m = p_measure<D>(G, p, S, op.y, op.w, shot);
            if (op.x == SDIMB_OP_RESET && m) g_pauli<D>(G, op.y, D - m, 0u);
            break;
          }
          case SDIMB_OP_N1: g_pauli<D>(G, op.y, (uint32_t)op.z & 0xFFu, (uint32_t)op.z >> 8); break;
          default: break;
        }
      }
    }
    __syncwarp();
    if (p.flags & SDIMB_WRITEBACK) {      // unpack into the uint8 store
      for (int q = 0; q < p.n; ++q) {
        uint8_t* row8 = T8 + (int64_t)q * p.row_bytes;
        for (int ln = lane; ln < p.W; ln += 32) {
          const int half = ln >= p.np, g = half ? ln - p.np : ln;
          const bool live = g < p.n;
          row8[ln] = live ? (uint8_t)G.getx(q, half * G.np + g) : 0;
          row8[p.W + ln] = live ? (uint8_t)G.getz(q, half * G.np + g) : 0;
        }
      }
      for (int ln = lane; ln < p.W; ln += 32) {
        const int half = ln >= p.np, g = half ? ln - p.np : ln;
        T8[p.phase_off + ln] = (g < p.n) ? (uint8_t)G.getp(half * G.np + g) : 0;
      }
    }
    __syncwarp();
  }
}

inline size_t planes_smem_bytes(int n, int d) {
  const int B = (d == 2) ? 1 : 2;
  const size_t np = (size_t)(n + 31) / 32 * 32, Wb = 2 * np / 32, RW = 2 * B * Wb;
  const size_t tab_words = (size_t)n * RW + 2 * Wb;
  return 4 * (tab_words + 2 * Wb + 4) + 32 * 16 + 2 * np + 2 * np + 16;
}

}  // namespace planes
