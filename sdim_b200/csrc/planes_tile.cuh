// Tile interpreter of the bit planes: SEVERAL SHOTS PER WARP for small tableaus (included by planes.cuh inside
// namespace planes).
//
// A row of the bit-plane image has Wb = 2 * ceil(n / 32) lane words: 4 at n <= 64, 8 at n <= 128.  The one-shot-per-
// warp interpreter (interp_planes_kernel) then runs every gate with 4 or 8 of its 32 lanes — configs 2, 3 and 4 of
// BASELINE.json (n = 64, 97, 49) used 12-25 % of the lanes they issued for.  Here a shot belongs to a TILE of
// LPS = 4 or 8 lanes (the power of two >= Wb) and the 32 / LPS tiles of a warp run different shots in lock step:
//   - one op fetch and one dispatch for all tiles (the op stream is the same for every shot);
//   - lane j of a tile owns lane word j of every row of ITS shot's image (shared memory, one image per tile) and keeps
//     phase word j in registers for the whole circuit: a gate is one LDS + a few LOP3 + one STS with every lane busy;
//   - N1 events are drawn per (op, shot): the draws of a batch of 32 ops are spread over the warp's lanes and handed to
//     the executing tiles as one packed word (4 bits per tile), ops whose event fires in no tile are never dispatched;
//   - a measurement (tableau_prime.py:262-363, same closed forms as p_measure) runs inside the tile: pivot search by
//     a tile reduction, the pivot-column walk with LPS rows per step, the rank-1 update with each lane on its own lane
//     word (column writes merged into it), phases in registers.  Its control flow depends on the X / Z blocks only,
//     which are the same in every shot of a batch that started from one tableau, so the tiles agree on it; the code
//     does not rely on that: a branch is entered when ANY tile needs it, the others are predicated off, loop bounds
//     are the maximum over the warp — control flow stays warp-uniform and every collective uses the full mask.
// Warps are independent (one-warp CTAs, no block barrier anywhere); shots are claimed 32 / LPS at a time.
#pragma once

template <int D>
struct TImg {
  static constexpr int EW = (D == 2) ? 2 : 4;    // words per entry
  uint32_t* tab;                                 // this tile's image: n rows of Wb entries (no padding entry: shared
                                                 // memory bounds the shots in flight, column walks are the minority)
  uint16_t* ar;                                  // [np] support rows (row | xs << 12 | zs << 14) / listed generators (g | f << 12)
  uint16_t* br;                                  // [np] rows whose stale destabilizer-p entry must be cleared
  int n, np, Wb, Wq, RS;
  __device__ __forceinline__ uint32_t* entry(int q, int j) const { return tab + q * RS + j * EW; }
  __device__ __forceinline__ XZ ld(int q, int j) const {
    if (D == 3) { const uint4 v = *reinterpret_cast<const uint4*>(entry(q, j)); return XZ{E{v.x, v.y}, E{v.z, v.w}}; }
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
};

// shared memory of one tile / of one warp (host and device agree through these)
inline size_t tile_img_words(int n, int d) {
  const size_t EW = (d == 2) ? 2 : 4, np = (size_t)(n + 31) / 32 * 32, Wb = 2 * np / 32;
  return ((size_t)n * EW * Wb + 3) & ~(size_t)3;
}
inline int tile_lps(int n) {
  const int np = (n + 31) / 32 * 32, Wb = 2 * np / 32;
  int lps = 4;
  while (lps < Wb) lps *= 2;
  return lps;
}
// words between the images of two tiles: image + scratch, shifted so that the LPS-lane rows of the tiles of one
// 128-byte wavefront (8 lanes of 16 bytes for d = 3, 16 lanes of 8 bytes for d = 2) fall into different banks
inline size_t tile_stride_words(int n, int d) {
  const size_t np = (size_t)(n + 31) / 32 * 32, EW = (d == 2) ? 2 : 4;
  size_t w = tile_img_words(n, d) + np;            // ar + br: 2 * np uint16 = np words
  w = (w + 31) & ~(size_t)31;                      // multiple of 128 bytes
  const size_t row = (size_t)tile_lps(n) * EW;     // words one tile touches per gate
  return row < 32 ? w + row : w;
}
// Global-image form: the tiles' images live in caller scratch (L1 / L2), shared memory keeps only the lists and the
// noise events, so that registers, not shared memory, bound the warps in flight (config 3: 8 -> 20+ warps per SM; the
// shared-memory form ran at 12 % warps-active, 33 % issue-active, top stall "wait": too few warps to hide even the
// ALU latency of its own dependent instructions).
#ifndef SDIMB_TILE_GLB_CTAS
#define SDIMB_TILE_GLB_CTAS 20     // one-warp CTAs per SM the global-image form is compiled for (<= 102 registers)
#endif
inline size_t tile_glb_img_stride_words(int n, int d) { return (tile_img_words(n, d) + 31) & ~(size_t)31; }
inline size_t tile_glb_lists_words(int n) { return ((size_t)(n + 31) / 32 * 32 + 31) & ~(size_t)31; }   // ar + br per tile
inline size_t tile_glb_smem_bytes(int n) { return 4 * ((32 / (size_t)tile_lps(n)) * tile_glb_lists_words(n) + 32) + 16; }
inline size_t tile_smem_bytes(int n, int d) {      // one warp: its tiles + the packed noise events of a batch
  return 4 * ((32 / (size_t)tile_lps(n)) * tile_stride_words(n, d) + 32) + 16;
}

// one gate on the lane word this lane owns (code is warp-uniform; pa / pb are per tile)
template <int D>
__device__ __forceinline__ void t_gate(const TImg<D>& G, const int j, const int code, const int a, const int b,
                                       const uint32_t pa, const uint32_t pb, E& ph) {
  switch (code) {
  case SDIMB_OP_X: case SDIMB_OP_X_INV: case SDIMB_OP_Z: case SDIMB_OP_Z_INV: case SDIMB_OP_N1: {
    const XZ v = G.ld(a, j);                                            // phase += po*(pb*x - pa*z)
    if (D == 3) ph = add3(ph, add3(smul3(v.x, pb), smul3(v.z, (3u - pa) % 3u)));
    else ph.h ^= ((pb & 1u) ? v.x.l : 0u) ^ ((pa & 1u) ? v.z.l : 0u);
    break;
  }
  case SDIMB_OP_H: case SDIMB_OP_H_INV: {
    const XZ v = G.ld(a, j);
    if (D == 3) {
      ph = add3(ph, neg3(mul3(v.x, v.z)));                              // phase -= x*z
      G.st(a, j, code == SDIMB_OP_H_INV ? XZ{v.z, neg3(v.x)} : XZ{neg3(v.z), v.x});
    } else {
      ph.h ^= v.x.l & v.z.l;                                            // phase += 2*x*z (mod 4); H == H^-1
      G.st(a, j, XZ{v.z, v.x});
    }
    break;
  }
  case SDIMB_OP_P: case SDIMB_OP_P_INV: {
    const XZ v = G.ld(a, j);
    const bool inv = code == SDIMB_OP_P_INV;
    if (D == 3) {
      ph = add3(ph, inv ? E{0u, v.x.h} : E{v.x.h, 0u});                 // phase +-= x(x-1)/2 = [x == 2]
      G.stz(a, j, add3(v.z, inv ? neg3(v.x) : v.x));
    } else {
      if (inv) { const uint32_t borrow = ~ph.l & v.x.l; ph.l ^= v.x.l; ph.h ^= borrow; }
      else { const uint32_t carry = ph.l & v.x.l; ph.l ^= v.x.l; ph.h ^= carry; }
      G.stz(a, j, E{v.z.l ^ v.x.l, 0u});
    }
    break;
  }
  case SDIMB_OP_CNOT: case SDIMB_OP_CNOT_INV: {
    const XZ va = G.ld(a, j), vb = G.ld(b, j);
    const bool inv = code == SDIMB_OP_CNOT_INV;
    if (D == 3) {
      G.stx(b, j, add3(vb.x, inv ? neg3(va.x) : va.x));                 // x[t] +-= x[c]
      G.stz(a, j, add3(va.z, inv ? vb.z : neg3(vb.z)));                 // z[c] -+= z[t]
    } else {
      G.stx(b, j, E{vb.x.l ^ va.x.l, 0u});
      G.stz(a, j, E{va.z.l ^ vb.z.l, 0u});
    }
    break;
  }
  case SDIMB_OP_CZ: case SDIMB_OP_CZ_INV: {
    const XZ va = G.ld(a, j), vb = G.ld(b, j);
    const bool inv = code == SDIMB_OP_CZ_INV;
    if (D == 3) {
      const E prod = mul3(va.x, vb.x);
      ph = add3(ph, inv ? neg3(prod) : prod);                           // phase +-= x[a]*x[b]
      G.stz(a, j, add3(va.z, inv ? neg3(vb.x) : vb.x));
      G.stz(b, j, add3(vb.z, inv ? neg3(va.x) : va.x));
    } else {
      ph.h ^= va.x.l & vb.x.l;
      G.stz(a, j, E{va.z.l ^ vb.x.l, 0u});
      G.stz(b, j, E{vb.z.l ^ va.x.l, 0u});
    }
    break;
  }
  case SDIMB_OP_SWAP: {
    const XZ va = G.ld(a, j), vb = G.ld(b, j);
    G.st(a, j, vb);
    G.st(b, j, va);
    break;
  }
  default: break;
  }
}

// Tile collectives of a CONVERGED warp: every lane of the warp executes them together (full member mask, a
// compile-time constant: no run-time mask checks), shuffles stay inside the tile by their width argument.  The tile
// interpreter keeps its control flow warp-uniform — a branch is taken when ANY tile needs it and the tiles that do not
// are predicated off — so these are legal everywhere in it.
template <int LPS>
struct TW {
  static constexpr uint32_t FULL = 0xFFFFFFFFu;
  int lane, base;
  __device__ __forceinline__ TW() {
    const int wl = threadIdx.x & 31;
    base = wl & ~(LPS - 1);
    lane = wl & (LPS - 1);
  }
  template <class V> __device__ __forceinline__ V shfl(V v, int src) const { return __shfl_sync(FULL, v, src, LPS); }
  template <class V> __device__ __forceinline__ V shfl_up(V v, int d) const { return __shfl_up_sync(FULL, v, d, LPS); }
  template <class V> __device__ __forceinline__ V shfl_down(V v, int d) const { return __shfl_down_sync(FULL, v, d, LPS); }
  __device__ __forceinline__ uint32_t ballot(bool pr) const { return (__ballot_sync(FULL, pr) >> base) & ((1u << LPS) - 1u); }
  __device__ __forceinline__ uint32_t sum(uint32_t v) const {
#pragma unroll
    for (int off = 1; off < LPS; off <<= 1) v += __shfl_xor_sync(FULL, v, off);
    return v;
  }
  __device__ __forceinline__ uint32_t min(uint32_t v) const {
#pragma unroll
    for (int off = 1; off < LPS; off <<= 1) { const uint32_t o = __shfl_xor_sync(FULL, v, off); v = o < v ? o : v; }
    return v;
  }
  __device__ __forceinline__ static int wmax(int v) { return (int)__reduce_max_sync(FULL, (uint32_t)v); }   // over the warp
  __device__ __forceinline__ static void sync() { __syncwarp(); }
};

// Measurement of qudit q by every tile of the warp at once.  `on` lanes (j < Wb) own a lane word; ph is this lane's
// phase word.  All lanes of the warp must call it together.
//
// UNI: every shot of the batch started from |0...0> (SDIMB_FRESH).  The X / Z blocks of a tableau do not depend on
// measurement outcomes or Pauli noise — gates act on them linearly, a random measurement picks its pivot and its row
// operations from the X block alone, Paulis (N1, the RESET correction) touch phases only — so the tiles of the warp hold
// IDENTICAL X / Z images and differ in their phase words only.  Everything a measurement READS from the X / Z blocks
// without writing (the pivot-column walk and its lists, the row pass of a deterministic measurement) is then done
// ONCE per warp, by all 32 lanes on the image of tile 0 (G0), instead of once per tile by LPS lanes; every tile still
// applies the row operations to its own image and its own phases.
template <int D, int LPS, bool UNI, class DRAW>
__device__ __forceinline__ uint32_t t_measure(const TW<LPS>& T, const TImg<D>& G, const TImg<D>& G0, const int q,
                                              const DRAW& draw_fn, E& ph) {
  constexpr uint32_t PO = (D == 2) ? 2u : 1u, ORDER = D * PO, FULL = 0xFFFFFFFFu;
  constexpr int WQM = LPS / 2;                                           // most stabilizer lane words a tile can have
  constexpr int WL = UNI ? 32 : LPS;                                     // lanes that share one walk
  const int j = T.lane, Wb = G.Wb, Wq = G.Wq, n = G.n;
  const bool on = j < Wb;
  const TImg<D>& W = UNI ? G0 : G;                                       // image the walks read; its scratch holds their lists
  const int wj = UNI ? (int)(threadIdx.x & 31) : j;
  const uint32_t lt = (1u << wj) - 1u;
  auto wballot = [&](bool pr) -> uint32_t { return UNI ? __ballot_sync(FULL, pr) : T.ballot(pr); };
  auto wsum = [&](uint32_t v) -> uint32_t { return UNI ? __reduce_add_sync(FULL, v) : T.sum(v); };
  E xq{0u, 0u};
  if (on) xq = G.ld(q, j).x;
  const uint32_t nz = xq.l | xq.h;
  // pivot: first stabilizer lane with an X component on q (tableau_prime.py:273-283)
  const uint32_t piv = T.min((j < Wq && nz) ? 32u * j + (uint32_t)(__ffs(nz) - 1) : kNoPivot);
  const bool rnd = piv != kNoPivot;
  uint32_t rec = 0;
  if (__any_sync(FULL, rnd)) {
    // ---- random branch (tableau_prime.py:294-334, exponentiate :365-380 folded in); tiles with !rnd idle through it ----
    const uint32_t draw = draw_fn();        // replayed or Philox: only a random measurement consumes it
    const int jp = rnd ? (int)(piv >> 5) : 0, bp = piv & 31, jd = Wq + jp;
    const uint32_t e = (D == 3) ? bit2(E{T.shfl(xq.l, jp), T.shfl(xq.h, jp)}, bp) : 1u;   // inverse of v mod 3 is v
    const uint32_t ps_old = bit2(E{T.shfl(ph.l, jp), T.shfl(ph.h, jp)}, bp);
    // one pass down the pivot column and the destabilizer-p column, LPS rows per step
    uint32_t sd_part = 0;
    int na = 0, nb = 0;
    // (two steps' loads are issued together: the list stores between them would otherwise order the loads)
    for (int base = 0; base < n; base += 2 * WL) {
      uint32_t xr[2], zr[2], od[2];
      XZ s[2], dd[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int r = base + u * WL + wj;
        s[u] = dd[u] = XZ{E{0u, 0u}, E{0u, 0u}};
        if (r < n && rnd) { s[u] = W.ld(r, jp); dd[u] = W.ld(r, jd); }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int r = base + u * WL + wj;
        od[u] = ((dd[u].x.l | dd[u].x.h | dd[u].z.l | dd[u].z.h) >> bp) & 1u;
        xr[u] = bit2(s[u].x, bp); zr[u] = bit2(s[u].z, bp);
        sd_part += xr[u] * zr[u];
        if (D == 3 && e == 2u) { xr[u] = (xr[u] >> 1) | ((xr[u] & 1u) << 1); zr[u] = (zr[u] >> 1) | ((zr[u] & 1u) << 1); }   // * 2 = negate
        const bool act = (xr[u] | zr[u]) != 0, stale = !act && od[u] != 0;
        const uint32_t ma = wballot(act), mb = wballot(stale);
        if (act) W.ar[na + __popc(ma & lt)] = (uint16_t)((uint32_t)r | (xr[u] << 12) | (zr[u] << 14));
        if (stale) W.br[nb + __popc(mb & lt)] = (uint16_t)r;
        na += __popc(ma);
        nb += __popc(mb);
      }
    }
    const uint32_t sd_raw = wsum(sd_part) % D;
    const uint32_t ps = (ps_old * e + PO * ((sd_raw * ((e * (e - 1u)) >> 1)) % D)) % ORDER;
    const uint32_t sd = (sd_raw * e * e) % D;
    T.sync();
    // row_i += f_i * pivot on the pivot's support, every lane on its own lane word; the column writes
    // (destabilizer p <- old pivot, stabilizer p <- Z_q, tableau_prime.py:323-333) ride on the same store
    E f = (D == 3) ? E{xq.h, xq.l} : xq;                                 // f = -X[q,i]
    const bool fixp = rnd && on && j == jp, fixd = rnd && on && j == jd;
    if (fixp) { f.l &= ~(1u << bp); f.h &= ~(1u << bp); }                // the pivot itself
    if (!rnd) f = E{0u, 0u};
    const bool work = rnd && on && ((f.l | f.h) != 0u || fixp || fixd);
    E dot{0u, 0u};
    const int na_w = T.wmax(na);
    for (int k0 = 0; k0 < na_w; k0 += 2) {           // two support rows per step: both loads before either store
      XZ v[2];
      uint32_t ent[2];
      bool act[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        act[u] = work && k0 + u < na;
        ent[u] = act[u] ? (uint32_t)W.ar[k0 + u] : 0u;
        v[u] = XZ{E{0u, 0u}, E{0u, 0u}};
        if (act[u]) v[u] = G.ld(ent[u] & 0xFFFu, j);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (!act[u]) continue;
        const int r = ent[u] & 0xFFFu;
        const uint32_t s = (ent[u] >> 12) & 3u, t = (ent[u] >> 14) & 3u;
        XZ nv;
        if (D == 3) {
          dot = add3(dot, smul3(v[u].z, s));                             // Z[:,i] . x_p (old Z)
          nv = XZ{add3(v[u].x, smul3(f, s)), add3(v[u].z, smul3(f, t))};
        } else {
          if (s) dot.l ^= v[u].z.l;
          nv = XZ{E{v[u].x.l ^ (s ? f.l : 0u), 0u}, E{v[u].z.l ^ (t ? f.l : 0u), 0u}};
        }
        if (fixp) { nv.x = setbit2(nv.x, bp, 0u); nv.z = setbit2(nv.z, bp, (r == q) ? 1u : 0u); }
        else if (fixd) { nv.x = setbit2(nv.x, bp, s); nv.z = setbit2(nv.z, bp, t); }
        G.st(r, j, nv);
      }
    }
    if (fixd) {            // rows outside the support: clear their stale destabilizer-p entry
      for (int k = 0; k < nb; ++k) {
        const int r = W.br[k];
        XZ dd = G.ld(r, jd);
        dd.x = setbit2(dd.x, bp, 0u);
        dd.z = setbit2(dd.z, bp, 0u);
        G.st(r, jd, dd);
      }
    }
    // P_i += f*ps + po*((Z_i.xs)*f + sd*f(f-1)/2*po)      (tableau_prime.py:310-312,317-319)
    if (D == 3) {
      E t = add3(smul3(f, ps), mul3(dot, f));
      t = add3(t, smul3(E{f.h, 0u}, sd));                                // f(f-1)/2 = [f == 2]
      ph = add3(ph, t);
    } else {
      ph = add4(ph, E{(ps & 1u) ? f.l : 0u, (ps & 2u) ? f.l : 0u});
      ph.h ^= dot.l & f.l;
    }
    if (fixd) ph = setbit2(ph, bp, ps);
    if (fixp) ph = setbit2(ph, bp, (ORDER - draw * PO) % ORDER);
    if (rnd) rec = draw;        // replayed or Philox (reference: random.choice, :332)
    T.sync();
  }
  if (__any_sync(FULL, !rnd)) {
    // ---- deterministic branch (tableau_prime.py:336-363): stabilizers i with f_i = destab X[q,i] != 0, in order ----
    const bool det = !rnd;
    const E fd = (det && on && j >= Wq) ? xq : E{0u, 0u};
    const E fsh{T.shfl_down(fd.l, Wq), T.shfl_down(fd.h, Wq)};          // lane j < Wq: destabilizer word j
    const E fl = (j < Wq) ? fsh : E{0u, 0u};
    uint32_t a1 = (D == 3) ? popsum<3>(mul3(fl, ph)) : (uint32_t)__popc(fl.l & ph.l) + 2u * (uint32_t)__popc(fl.l & ph.h);
    a1 = T.sum(a1) % ORDER;
    // ordered list (generator | f << 12) and, per stabilizer lane word, the mask of listed generators
    uint32_t m = fl.l | fl.h;
    const int cnt = __popc(m);
    int incl = cnt;
#pragma unroll
    for (int d2 = 1; d2 < LPS; d2 <<= 1) {
      const int o = T.shfl_up(incl, d2);
      if (j >= d2) incl += o;
    }
    const int total = T.shfl(incl, LPS - 1);
    {
      int pos = incl - cnt;
      uint32_t mm = m;
      while (mm) {
        const int b = __ffs(mm) - 1;
        mm &= mm - 1;
        G.ar[pos++] = (uint16_t)((32 * j + b) | (bit2(fl, b) << 12));
      }
    }
    uint32_t mw[WQM];
#pragma unroll
    for (int w = 0; w < WQM; ++w) mw[w] = T.shfl(m, w);
    T.sync();
    // rows over the tile's lanes; a row on which every listed generator is the identity (the usual case: stabilizers
    // of a code act on a few qudits) costs its loads and one test
    uint32_t part = 0;
    constexpr int RU = 4;                                                // rows per lane whose loads are in flight together
    for (int r0 = wj; r0 < n && total != 0; r0 += RU * WL) {
      XZ v[RU][WQM];
      uint32_t hit[RU];
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        const int r = r0 + u * WL;
        hit[u] = 0;
#pragma unroll
        for (int w = 0; w < WQM; ++w) {
          v[u][w] = XZ{E{0u, 0u}, E{0u, 0u}};
          if (r < n && w < Wq && mw[w]) {
            v[u][w] = W.ld(r, w);
            hit[u] |= (v[u][w].x.l | v[u][w].x.h | v[u][w].z.l | v[u][w].z.h) & mw[w];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        if (!hit[u]) continue;
        uint32_t az = 0, cross = 0, sdg = 0;
        for (int k = 0; k < total; ++k) {
          const uint32_t ent = G.ar[k];
          const int g = ent & 0xFFFu, w = g >> 5, b = g & 31;
          const uint32_t fv = ent >> 12;
          XZ vv = v[u][0];
#pragma unroll
          for (int x = 1; x < WQM; ++x) if (w == x) vv = v[u][x];
          if ((((vv.x.l | vv.x.h | vv.z.l | vv.z.h) >> b) & 1u) == 0u) continue;    // generator g is the identity on this row
          const uint32_t xi = bit2(vv.x, b), zi = bit2(vv.z, b);
          cross += (fv * xi) * az;                                       // ancilla_z . (f * x_i), running ancilla
          az = (az + fv * zi) % D;
          sdg += xi * zi * ((fv * (fv - 1u)) >> 1);
          if ((k & 15) == 15) { cross %= D; sdg %= D; }
        }
        part += (cross + PO * sdg) % D;
      }
    }
    part = wsum(part);
    const uint32_t ap = (a1 + PO * (part % D)) % ORDER;
    const uint32_t outcome = (D == 3) ? (3u - ap) % 3u : (((ap + 1u) >> 1) & 1u);   // (-ap // po) % d  (:362)
    if (det) rec = outcome | SDIMB_REC_DET;
    T.sync();
  }
  return rec;
}

template <int D, int LPS, bool UNI, bool GLB = false>
__global__ void __launch_bounds__(32, GLB ? SDIMB_TILE_GLB_CTAS : 1) interp_tile_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int TPW = 32 / LPS, EW = TImg<D>::EW;
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  const TW<LPS> T;
  const int lane = threadIdx.x & 31, tw = lane / LPS, j = T.lane;
  TImg<D> G;
  G.n = p.n;
  G.np = (p.n + 31) / 32 * 32;
  G.Wq = G.np / 32;
  G.Wb = 2 * G.Wq;
  G.RS = EW * G.Wb;
  const int img_words = (p.n * G.RS + 3) & ~3;
  uint32_t* const sm = reinterpret_cast<uint32_t*>(smem);
  TImg<D> G0 = G;                                                        // tile 0 of the warp (UNI: the image the walks read)
  uint32_t* evs;                                                         // [32] packed N1 events of the current batch
  if (GLB) {                   // images in caller scratch, one per resident tile; lists and events in shared memory
    uint32_t* const img0 = p.plane_slab + (int64_t)blockIdx.x * TPW * p.img_stride_words;
    G.tab = img0 + (int64_t)tw * p.img_stride_words;
    G.ar = reinterpret_cast<uint16_t*>(sm + (size_t)tw * p.tile_stride_words);
    G.br = G.ar + G.np;
    G0.tab = img0;
    G0.ar = reinterpret_cast<uint16_t*>(sm);
    G0.br = G0.ar + G.np;
    evs = sm + (size_t)TPW * p.tile_stride_words;
  } else {
    G.tab = sm + (size_t)tw * p.tile_stride_words;
    G.ar = reinterpret_cast<uint16_t*>(G.tab + img_words);
    G.br = G.ar + G.np;
    G0.tab = sm;
    G0.ar = reinterpret_cast<uint16_t*>(G0.tab + img_words);
    G0.br = G0.ar + G.np;
    evs = sm + (size_t)TPW * p.tile_stride_words;
  }
  const bool on = j < G.Wb;
  const int jj = on ? j : 0;                                             // idle lanes of a tile (Wb < LPS) shadow word 0, never store

  for (int64_t round = 0;; ++round) {
    int64_t s0 = 0;
    if (lane == 0) s0 = p.shot_counter ? (int64_t)atomicAdd(p.shot_counter, (unsigned)TPW) : ((int64_t)blockIdx.x + round * gridDim.x) * TPW;
    s0 = __shfl_sync(FULL, s0, 0);
    if (s0 >= p.shots) break;
    const bool real = s0 + tw < p.shots;                                 // tiles past the end shadow the last shot, write nothing
    const int64_t shot = real ? s0 + tw : p.shots - 1;
    // ---- load: |0...0> or pack from the uint8 store ----
    for (int i = j; i < img_words; i += LPS) G.tab[i] = 0u;
    T.sync();
    E ph{0u, 0u};
    uint8_t* const T8 = p.tab ? p.tab + shot * p.shot_bytes : nullptr;
    if (p.flags & SDIMB_FRESH) {
      for (int q = j; q < p.n; q += LPS) {
        G.entry(q, q >> 5)[EW / 2] |= 1u << (q & 31);                    // stabilizer q = Z_q   (z_l plane)
        G.entry(q, G.Wq + (q >> 5))[0] |= 1u << (q & 31);                // destabilizer q = X_q (x_l plane)
      }
    } else {
      for (int q = 0; q < p.n; ++q) {
        const uint8_t* row8 = T8 + (int64_t)q * p.row_bytes;
        if (on) {
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
      if (on) {
        for (int b = 0; b < 32; ++b) {
          const int ln = 32 * j + b;
          const int half = ln >= G.np, g = half ? ln - G.np : ln;
          if (g >= p.n) continue;
          const uint32_t v = T8[p.phase_off + half * p.np + g];
          ph.l |= (v & 1u) << b; ph.h |= ((v >> 1) & 1u) << b;
        }
      }
    }
    __syncwarp();

    int4 ahead = make_int4(SDIMB_OP_I, 0, 0, 0);
    if (lane < p.n_ops) ahead = __ldg(p.ops + lane);
    for (int64_t i0 = 0; i0 < p.n_ops; i0 += 32) {
      int4 mine = ahead;
      ahead = make_int4(SDIMB_OP_I, 0, 0, 0);
      if (i0 + 32 + lane < p.n_ops) ahead = __ldg(p.ops + i0 + 32 + lane);
      mine.x &= SDIMB_OP_MASK;
      // N1 events of this batch for every tile of the warp: (noise op, tile) pairs dealt over the 32 lanes, one packed
      // word per op (4 bits per tile: a | b << 2); an op whose event fires in no tile is never dispatched
      const uint32_t nm = __ballot_sync(FULL, mine.x == SDIMB_OP_N1);
      uint32_t my_ev = 0;
      if (nm) {
        evs[lane] = 0u;
        __syncwarp();
        const int items = __popc(nm) * TPW;
        for (int base = 0; base < items; base += 32) {
          const int item = base + lane;
          const bool valid = item < items;
          const int k = valid ? (int)__fns(nm, 0u, item / TPW + 1) : 0, t = item % TPW;
          const int slot = __shfl_sync(FULL, mine.w, k);
          if (valid) {
            const int64_t sh = (s0 + t < p.shots) ? s0 + t : p.shots - 1;
            const uint32_t ev = p_noise_event<D>(p, slot, sh);
            if (ev) atomicOr(&evs[k], ((ev & 3u) | (((ev >> 8) & 3u) << 2)) << (4 * t));
          }
        }
        __syncwarp();
        my_ev = evs[lane];
        __syncwarp();
      }
      const bool live = mine.x != SDIMB_OP_I && mine.x != SDIMB_OP_BARRIER && (mine.x != SDIMB_OP_N1 || my_ev != 0u);
      uint32_t todo = __ballot_sync(FULL, live);
#pragma unroll 1
      while (todo) {
        const int k = __ffs(todo) - 1;
        todo &= todo - 1;
        const int code = __shfl_sync(FULL, mine.x, k), a = __shfl_sync(FULL, mine.y, k), b = __shfl_sync(FULL, mine.z, k);
        if (code < SDIMB_OP_M) {
          uint32_t pa = 0u, pb = 0u;
          if (code <= SDIMB_OP_Z_INV) {
            const uint32_t e = (code == SDIMB_OP_X || code == SDIMB_OP_Z) ? 1u : D - 1u;
            if (code <= SDIMB_OP_X_INV) pa = e; else pb = e;
          }
          t_gate<D>(G, jj, on ? code : SDIMB_OP_I, a, b, pa, pb, ph);
        } else if (code == SDIMB_OP_N1) {
          const uint32_t pack = __shfl_sync(FULL, my_ev, k), nib = (pack >> (4 * tw)) & 0xFu;
          if (nib && on) t_gate<D>(G, jj, SDIMB_OP_N1, a, 0, nib & 3u, nib >> 2, ph);
        } else {
          const int slot = __shfl_sync(FULL, mine.w, k);
          if (code == SDIMB_OP_M_X) {
            if (on) t_gate<D>(G, jj, SDIMB_OP_H_INV, a, 0, 0u, 0u, ph);   // tableau_gates.py:292-296
            T.sync();
          }
          auto draw_fn = [&]() -> uint32_t {      // outcome this measurement takes if it is random
            if (p.replay_meas) return p.replay_meas[shot * p.n_meas + slot];
            const uint64_t gshot = (uint64_t)(p.shot_offset + shot);
            const uint4 r = philox4x32((uint32_t)gshot, (uint32_t)(gshot >> 32), (uint32_t)slot, 0u, (uint32_t)p.seed,
                                       (uint32_t)(p.seed >> 32));
            return __umulhi(r.x, (uint32_t)D);
          };
          T.sync();                     // gates of this tile's lanes in front of the measurement
          const uint32_t rec = t_measure<D, LPS, UNI>(T, G, G0, a, draw_fn, ph);
          if (j == 0 && real) p.records[shot * p.rec_stride + slot] = (uint8_t)rec;
          const uint32_t m = rec & SDIMB_REC_VALUE;
          if (code == SDIMB_OP_RESET && m && on) t_gate<D>(G, jj, SDIMB_OP_N1, a, 0, D - m, 0u, ph);   // program.py:335-339
        }
      }
    }
    __syncwarp();
    if (p.flags & SDIMB_WRITEBACK) {                // unpack into the uint8 store; phase words travel through ar
      uint2* const phs = reinterpret_cast<uint2*>(G.ar);
      if (on) phs[j] = make_uint2(ph.l, ph.h);
      T.sync();                                     // (warp-wide: tiles past the end skip the stores only)
      for (int q = 0; q < p.n && real; ++q) {
        uint8_t* row8 = T8 + (int64_t)q * p.row_bytes;
        for (int ln = j; ln < p.W; ln += LPS) {
          const int half = ln >= p.np, g = half ? ln - p.np : ln;
          const bool lv = g < p.n;
          const int gl = half * G.np + g;
          const XZ v = G.ld(q, lv ? gl >> 5 : 0);
          row8[ln] = lv ? (uint8_t)bit2(v.x, gl & 31) : 0;
          row8[p.W + ln] = lv ? (uint8_t)bit2(v.z, gl & 31) : 0;
        }
      }
      for (int ln = j; ln < p.W && real; ln += LPS) {
        const int half = ln >= p.np, g = half ? ln - p.np : ln;
        const int gl = half * G.np + g;
        const uint2 w = phs[(g < p.n) ? gl >> 5 : 0];
        T8[p.phase_off + ln] = (g < p.n) ? (uint8_t)bit2(E{w.x, w.y}, gl & 31) : 0;
      }
    }
    __syncwarp();
  }
}
