// Generator-major measurement runs of the bit-plane interpreter (included by planes.cuh inside namespace planes).
//
// The interpreter's image is qudit-major — row q holds (x, z) of all 2np generator lanes — because a gate then
// touches one contiguous row.  A measurement wants the other orientation: the pivot's support and the rowsum
// (tableau_prime.py:294-334) are properties of GENERATORS, and in the qudit-major image every measurement walks a
// column over all n rows (one 32-byte sector per row: 16 KB through L1/L2 per measurement at n = 256, two thirds of
// the headline's time in round 1).  The run of M ops that ends a circuit (sdimb_schedule marks it, SDIMB_GM_*; the
// usual "measure every qudit") is therefore executed by a second kernel, run_tail_kernel: the interpreter leaves
// every shot's image at its own address (KParams::img_per_shot), and ONE WARP PER SHOT
//   - transposes the image in registers (32 x 32 bit blocks, five butterfly stages per block) into
//       B   [np stabilizers][Wq = np/32 qudit words]  entries (x_l, x_h, z_l, z_h) / (x, z) for d = 2: rows contiguous
//       Bd  [np destabilizers][Wq]                     entries (x_l, x_h) / (x): X ONLY — the run does not keep the tableau,
//                                                       and a destabilizer's Z part and phase never reach a record
//       QX  [n qudits][Wb = 2np/32 lane words]        the X half of the qudit-major image, compact (no z, no padding)
//       ph8 [2np] phases as bytes in shared memory
//   - and runs every measurement of the run on them:
//       row q of QX             -> pivot (first stabilizer lane with x != 0) and the factors f_i = -X[q,i] of all lanes
//       row p of B              -> the pivot's support over qudits; xs . zs by popcounts of bit-sliced products
//       rows i of B, f_i != 0   -> row_i += f_i * row_p, dot products Z_i . xs by popcount (32 / Wq generators per pass)
//       rows r of QX, xs_r != 0 -> X[r,:] += xs_r * f, plus the two column fixes (stabilizer p <- Z_q, destabilizer p <- pivot)
//       deterministic           -> product of the listed stabilizer rows of B, 32 qudits per lane and instruction
// Every access is a contiguous row; nothing walks a column, nothing needs a block barrier, and with 32 independent
// warps (shots) per SM the dependent round trips of one shot hide behind the others'.  Needs Wb <= 32 (n <= 512).
#pragma once

// ---- arithmetic generic in D (bit-sliced over 32 lanes) ------------------------------------------------------
template <int D> __device__ __forceinline__ E addD(E a, E b) { return (D == 3) ? add3(a, b) : E{a.l ^ b.l, 0u}; }
template <int D> __device__ __forceinline__ E negD(E a) { return (D == 3) ? neg3(a) : a; }
template <int D> __device__ __forceinline__ E mulD(E a, E b) { return (D == 3) ? mul3(a, b) : E{a.l & b.l, 0u}; }
template <int D> __device__ __forceinline__ E smulD(E a, uint32_t s) { return (D == 3) ? smul3(a, s) : E{s ? a.l : 0u, 0u}; }
template <int D> __device__ __forceinline__ uint32_t popsum(E a) {          // sum of the 32 lane values
  return (D == 3) ? (uint32_t)__popc(a.l) + 2u * (uint32_t)__popc(a.h) : (uint32_t)__popc(a.l);
}

// transpose of a 32 x 32 bit block held in registers: out[c] bit k = in[k] bit c
__device__ __forceinline__ void transpose32(uint32_t (&m)[32]) {
  uint32_t mask = 0x0000FFFFu;
#pragma unroll
  for (int j = 16; j; j >>= 1) {
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      if (k & j) continue;
      const uint32_t t = ((m[k] >> j) ^ m[k + j]) & mask;
      m[k] ^= t << j;
      m[k + j] ^= t;
    }
    mask ^= mask << (j >> 1);
  }
}

template <int D>
struct GMImg {
  static constexpr int EW = (D == 2) ? 2 : 4;    // words per entry of B
  static constexpr int EX = EW / 2;              // words per entry of QX
  uint32_t* B;                                   // stabilizer rows: [np][Wq] entries of EW words
  uint32_t* Bd;                                  // destabilizer rows: [np][Wq] entries of EX words — X ONLY: a run that
                                                 // does not keep the tableau never reads a destabilizer's Z part or phase
                                                 // (factors come from its X part, tableau_prime.py:350; rowsum writes it)
  uint32_t* QX;
  uint32_t ph8;                                  // shared-space ADDRESSES (32-bit) of the tile's phase bytes [2np] and
  uint32_t list;                                 // of its scratch list [2np] uint16: explicit ld/st.shared below — through
                                                 // generic pointers the compiler rebuilt the CTA's shared window (S2R
                                                 // SR_CgaCtaId + LEA) at many of the ~15 access sites of a measurement
  int n, np, Wq, Wb, gs_shift;                   // group of lanes per B row = 1 << gs_shift >= Wq
  __device__ __forceinline__ uint32_t ldl(int k) const {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(list + 2u * (uint32_t)k) : "memory");
    return v;
  }
  __device__ __forceinline__ void stl(int k, uint32_t v) const {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(list + 2u * (uint32_t)k), "r"(v) : "memory");
  }
  __device__ __forceinline__ uint32_t ldp(int i) const {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(ph8 + (uint32_t)i) : "memory");
    return v;
  }
  __device__ __forceinline__ void stp(int i, uint32_t v) const {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(ph8 + (uint32_t)i), "r"(v) : "memory");
  }
  __device__ __forceinline__ void place(uint32_t* slab) {        // B, Bd, QX of one shot, in this order
    B = slab;
    Bd = B + (size_t)np * Wq * EW;
    QX = Bd + (size_t)np * Wq * EX;
  }
  __device__ __forceinline__ uint32_t* bent(int g, int w) const { return B + (uint32_t)((g * Wq + w) * EW); }
  __device__ __forceinline__ uint32_t* dent(int g, int w) const { return Bd + (uint32_t)(((g - np) * Wq + w) * EX); }
  __device__ __forceinline__ uint32_t* qent(int r, int j) const { return QX + (uint32_t)((r * Wb + j) * EX); }
  __device__ __forceinline__ E ldBx(int g, int w) const {
    const uint32_t* e = (g >= np) ? dent(g, w) : bent(g, w);
    if (D == 3) { const uint2 v = *reinterpret_cast<const uint2*>(e); return E{v.x, v.y}; }
    return E{e[0], 0u};
  }
  __device__ __forceinline__ XZ ldB(int g, int w) const {
    XZ r{ldBx(g, w), E{0u, 0u}};
    if (g < np) {
      if (D == 3) { const uint2 v = *reinterpret_cast<const uint2*>(bent(g, w) + 2); r.z = E{v.x, v.y}; }
      else r.z = E{bent(g, w)[1], 0u};
    }
    return r;
  }
  __device__ __forceinline__ void stB(int g, int w, XZ v) const {
    if (g >= np) {
      if (D == 3) *reinterpret_cast<uint2*>(dent(g, w)) = make_uint2(v.x.l, v.x.h);
      else dent(g, w)[0] = v.x.l;
    } else {
      if (D == 3) *reinterpret_cast<uint4*>(bent(g, w)) = make_uint4(v.x.l, v.x.h, v.z.l, v.z.h);
      else *reinterpret_cast<uint2*>(bent(g, w)) = make_uint2(v.x.l, v.z.l);
    }
  }
  __device__ __forceinline__ E ldqx(int r, int j) const {
    if (D == 3) { const uint2 v = *reinterpret_cast<const uint2*>(qent(r, j)); return E{v.x, v.y}; }
    return E{qent(r, j)[0], 0u};
  }
  // lane words j (even) and j + 1 of row r in one access
  __device__ __forceinline__ void ldqx2(int r, int j, E& a, E& b) const {
    if (D == 3) { const uint4 v = *reinterpret_cast<const uint4*>(qent(r, j)); a = E{v.x, v.y}; b = E{v.z, v.w}; }
    else { const uint2 v = *reinterpret_cast<const uint2*>(qent(r, j)); a = E{v.x, 0u}; b = E{v.y, 0u}; }
  }
  __device__ __forceinline__ void stqx2(int r, int j, E a, E b) const {
    if (D == 3) *reinterpret_cast<uint4*>(qent(r, j)) = make_uint4(a.l, a.h, b.l, b.h);
    else *reinterpret_cast<uint2*>(qent(r, j)) = make_uint2(a.l, b.l);
  }
  __device__ __forceinline__ void stqx(int r, int j, E x) const {
    if (D == 3) *reinterpret_cast<uint2*>(qent(r, j)) = make_uint2(x.l, x.h);
    else qent(r, j)[0] = x.l;
  }
};

// Both directions of the transposition, by the nt threads (tid = 0..nt-1) that own the image.  One work item = one 32 x 32 bit block = one plane of
// (qudit block w, lane word j): to_gm reads 32 rows of the qudit-major image and writes 32 generator rows of B (and
// the X planes, untransposed, into QX); !to_gm reads B and rebuilds the qudit-major image.
template <int D, class GEO>
__device__ __noinline__ void gm_transpose(const GEO G, const GMImg<D> M, const bool to_gm, const int tid, const int nt) {
  constexpr int EW = GMImg<D>::EW, EX = GMImg<D>::EX;
  const int items = M.Wq * M.Wb * EW;
  for (int it = tid; it < items; it += nt) {
    const int pl = it % EW, jw = it / EW, j = jw % M.Wb, w = jw / M.Wb;
    const bool destab = j >= M.Wq;                                     // lane words of the destabilizer half
    if (destab && pl >= EX) continue;                                  // their Z planes are not kept (GMImg::Bd)
    uint32_t* const a = G.entry(32 * w, j) + pl;                       // + k * RS, rows k < n - 32 w
    uint32_t* const b = destab ? M.dent(32 * j, w) + pl : M.bent(32 * j, w) + pl;   // + k * Wq * EX / EW
    const int bs = destab ? M.Wq * EX : M.Wq * EW;
    const int alim = min(32, G.n - 32 * w);
    uint32_t* const src = to_gm ? a : b;
    uint32_t* const dst = to_gm ? b : a;
    const int ss = to_gm ? G.RS : bs, ds = to_gm ? bs : G.RS;
    const int sl = to_gm ? alim : 32, dl = to_gm ? 32 : alim;
    uint32_t m[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) m[k] = (k < sl) ? src[k * ss] : 0u;
    if (to_gm && pl < EX) {
      uint32_t* const qx = M.qent(32 * w, j) + pl;
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (k < alim) qx[k * M.Wb * EX] = m[k];
    }
    transpose32(m);
#pragma unroll
    for (int k = 0; k < 32; ++k)
      if (k < dl) dst[k * ds] = m[k];
  }
}

// The same transposition for the kernel that ran the gates on a SHARED-MEMORY image (gate_stream_kernel): to B / Bd / QX
// only, shared-space loads with 32-bit addresses, no per-element predicates on full 32-row blocks.  (The general
// routine above spent ~11 instructions per loaded word on generic addressing and predication: a quarter of the front
// kernel's instructions.)
template <int D, class GEO>
__device__ __noinline__ void gm_transpose_out_smem(const GEO G, const GMImg<D> M, const int tid, const int nt) {
  constexpr int EW = GMImg<D>::EW, EX = GMImg<D>::EX;
  const int items = M.Wq * M.Wb * EW;
  const uint32_t ssb = 4u * (uint32_t)G.RS;                               // bytes between two rows of the image
  const int64_t qs = (int64_t)M.Wb * EX;                                  // words between two rows of QX
  for (int it = tid; it < items; it += nt) {
    const int pl = it % EW, jw = it / EW, j = jw % M.Wb, w = jw / M.Wb;
    const bool destab = j >= M.Wq;
    if (destab && pl >= EX) continue;                                     // Z planes of destabilizers are not kept
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(G.entry(32 * w, j) + pl);
    uint32_t* b = destab ? M.dent(32 * j, w) + pl : M.bent(32 * j, w) + pl;
    const int64_t bs = destab ? (int64_t)M.Wq * EX : (int64_t)M.Wq * EW;
    const int alim = min(32, G.n - 32 * w);
    uint32_t m[32];
    if (alim == 32) {
#pragma unroll
      for (int k = 0; k < 32; ++k) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(m[k]) : "r"(sa + (uint32_t)k * ssb));
      if (pl < EX) {
        uint32_t* qx = M.qent(32 * w, j) + pl;
#pragma unroll
        for (int k = 0; k < 32; ++k) { *qx = m[k]; qx += qs; }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        m[k] = 0u;
        if (k < alim) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(m[k]) : "r"(sa + (uint32_t)k * ssb));
      }
      if (pl < EX) {
        uint32_t* const qx = M.qent(32 * w, j) + pl;
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (k < alim) qx[k * qs] = m[k];
      }
    }
    transpose32(m);
#pragma unroll
    for (int k = 0; k < 32; ++k) { *b = m[k]; b += bs; }
  }
}

// ---- one shot per TILE of LPS lanes ---------------------------------------------------------------------------
// At n = 256 a row of QX has 16 lane words and a row of B 8 entries: half a warp holds everything a measurement
// touches at once.  A shot therefore belongs to a tile of LPS = 4..32 lanes (the power of two >= Wb); the tiles of a
// warp run different shots in lock step (the control flow of a measurement depends on the X/Z blocks only, which are
// the same in every shot of a fresh batch; where shots do differ the hardware diverges the tiles, all collectives
// below are restricted to the tile's member mask).  Per shot this halves the instructions that do not scale with the
// data (pivot search, scans, loop control) and doubles the shots in flight per SM.
template <int LPS>
struct Tile {
  uint32_t mask;   // member mask of this tile inside its warp
  int base;        // first warp lane of the tile
  int lane;        // rank inside the tile
  __device__ __forceinline__ Tile() {
    const int wl = threadIdx.x & 31;
    base = wl & ~(LPS - 1);
    lane = wl & (LPS - 1);
    mask = (LPS == 32) ? 0xFFFFFFFFu : (((1u << LPS) - 1u) << base);
  }
  template <class T> __device__ __forceinline__ T shfl(T v, int src) const { return __shfl_sync(mask, v, src, LPS); }
  template <class T> __device__ __forceinline__ T shfl_xor(T v, int m) const { return __shfl_xor_sync(mask, v, m, LPS); }
  template <class T> __device__ __forceinline__ T shfl_up(T v, int d) const { return __shfl_up_sync(mask, v, d, LPS); }
  template <class T> __device__ __forceinline__ T shfl_down(T v, int d) const { return __shfl_down_sync(mask, v, d, LPS); }
  __device__ __forceinline__ uint32_t ballot(bool pr) const { return (__ballot_sync(mask, pr) & mask) >> base; }
  __device__ __forceinline__ uint32_t min(uint32_t v) const { return __reduce_min_sync(mask, v); }
  __device__ __forceinline__ uint32_t max(uint32_t v) const { return __reduce_max_sync(mask, v); }
  __device__ __forceinline__ uint32_t sum(uint32_t v) const { return __reduce_add_sync(mask, v); }
  __device__ __forceinline__ void sync() const { __syncwarp(mask); }
};

// Set bits of up to two mask words per lane (word indices w0 < w1 ascending with the lane) -> list entries
// (32 * w + bit) | value << 12, in ascending order; returns the total (tile-uniform).
template <int LPS, class IMG>
__device__ __forceinline__ int gm_compact(const Tile<LPS>& T, const IMG& M, uint32_t m0, E v0, int w0, uint32_t m1, E v1,
                                          int w1) {
  const int c0 = __popc(m0), cnt = c0 + __popc(m1);
  int incl = cnt;
#pragma unroll
  for (int d2 = 1; d2 < LPS; d2 <<= 1) {
    const int o = T.shfl_up(incl, d2);
    if (T.lane >= d2) incl += o;
  }
  int pos = incl - cnt;
  while (m0) {
    const int b = __ffs(m0) - 1;
    m0 &= m0 - 1;
    M.stl(pos++, (uint32_t)(32 * w0 + b) | (bit2(v0, b) << 12));
  }
  while (m1) {
    const int b = __ffs(m1) - 1;
    m1 &= m1 - 1;
    M.stl(pos++, (uint32_t)(32 * w1 + b) | (bit2(v1, b) << 12));
  }
  const int total = T.shfl(incl, LPS - 1);
  T.sync();
  return total;
}

// Measurement of qudit q on the generator-major image by one tile (tableau_prime.py:262-363; same closed forms as
// p_measure).  Needs Wb <= LPS.  Row q of QX is held one lane word per lane (pivot search, factors); rows of B and the
// rows of QX that get updated are held TWO entries per lane, so a B row occupies gb = pow2 >= Wq/2 lanes and LPS / gb
// rows are in flight per pass.  The shot's state lives in L2 / HBM, so the loads of a batch of passes are issued
// together, the generator rows together with the pivot row (they depend on row q only), and row q of the NEXT
// measurement (known from the op stream) is prefetched meanwhile.
#ifndef SDIMB_GM_UNROLL_B
#define SDIMB_GM_UNROLL_B 2        // passes over B rows whose loads are in flight together
#endif
#ifndef SDIMB_GM_UNROLL_Q
#define SDIMB_GM_UNROLL_Q 2        // passes over QX rows whose loads are in flight together
#endif
template <int D, int LPS>
__device__ __forceinline__ uint32_t gm_measure(const Tile<LPS>& T, const GMImg<D>& M, const int q, const int q_next,
                                               const uint32_t draw) {
  constexpr uint32_t PO = (D == 2) ? 2u : 1u, ORDER = D * PO;
  constexpr int UB = SDIMB_GM_UNROLL_B, UQ = SDIMB_GM_UNROLL_Q;
  const int lane = T.lane;
  const int Wq = M.Wq, Wb = M.Wb, np = M.np;
  const XZ zero{E{0u, 0u}, E{0u, 0u}};
  const E ezero{0u, 0u};
  E xq = ezero;
  if (lane < Wb) xq = M.ldqx(q, lane);
  if (q_next >= 0 && lane < Wb) asm volatile("prefetch.global.L2 [%0];" ::"l"(M.qent(q_next, lane)));
  const uint32_t nz = xq.l | xq.h;
  // pivot: first stabilizer lane with an X component on q (tableau_prime.py:273-283)
  const uint32_t piv = T.min((lane < Wq && nz) ? 32u * lane + (uint32_t)(__ffs(nz) - 1) : kNoPivot);
  // B rows: gb lanes per row, lane `sub` of a group holds entries 2 sub and 2 sub + 1
  const int gb = 1 << M.gs_shift, sub = lane & (gb - 1), grp = lane >> M.gs_shift, ng = LPS >> M.gs_shift;
  const int w0 = 2 * sub, w1 = 2 * sub + 1;
  const bool has0 = w0 < Wq, has1 = w1 < Wq;
  uint32_t rec;
  if (piv != kNoPivot) {
    // ---- random branch (tableau_prime.py:294-334, exponentiate :365-380 folded in) ----
    const int jp = piv >> 5, bp = piv & 31;
    XZ pv0 = zero, pv1 = zero;                           // the pivot row, replicated in every group
    E dx0 = ezero, dx1 = ezero;                          // X support of the destabilizer that is about to be overwritten
    if (has0) { pv0 = M.ldB(piv, w0); dx0 = M.ldBx(np + piv, w0); }
    if (has1) { pv1 = M.ldB(piv, w1); dx1 = M.ldBx(np + piv, w1); }
    const uint32_t e = (D == 3) ? bit2(E{T.shfl(xq.l, jp), T.shfl(xq.h, jp)}, bp) : 1u;
    E f = negD<D>(xq);                                   // f_i = -X[q,i]; the pivot and its destabilizer are replaced below
    if (lane == jp || lane == Wq + jp) { f.l &= ~(1u << bp); f.h &= ~(1u << bp); }
    const int total = gm_compact<LPS>(T, M, f.l | f.h, f, lane, 0u, ezero, 0);
    // first batch of generator rows: requested before anything waits for the pivot row
    XZ va[UB], vb[UB];
    uint32_t ent[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int k = u * ng + grp;
      ent[u] = (k < total) ? M.ldl(k) | 0x8000u : 0u;
      va[u] = vb[u] = zero;
      if ((ent[u] & 0x8000u) && has0) va[u] = M.ldB(ent[u] & 0xFFFu, w0);
      if ((ent[u] & 0x8000u) && has1) vb[u] = M.ldB(ent[u] & 0xFFFu, w1);
    }
    uint32_t sdp = popsum<D>(mulD<D>(pv0.x, pv0.z)) + popsum<D>(mulD<D>(pv1.x, pv1.z));
    for (int off = 1; off < gb; off <<= 1) sdp += T.shfl_xor(sdp, off);
    const uint32_t sd_raw = sdp % D;
    const uint32_t ps_old = M.ldp(piv);
    const uint32_t ps = (ps_old * e + PO * ((sd_raw * ((e * (e - 1u)) >> 1)) % D)) % ORDER;
    const uint32_t sd = (sd_raw * e * e) % D;
    if (D == 3 && e == 2u) { pv0.x = neg3(pv0.x); pv0.z = neg3(pv0.z); pv1.x = neg3(pv1.x); pv1.z = neg3(pv1.z); }
    // row_i += f_i * pivot for every generator with a factor; ng generators per pass   (pivot <- pivot^e = (xs, zs, ps))
    for (int k0 = 0;;) {
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        if (k0 + u * ng >= total) break;                                  // tile-uniform
        const bool act = (ent[u] & 0x8000u) != 0u;
        const int i = ent[u] & 0xFFFu;
        const uint32_t fi = (ent[u] >> 12) & 3u;
        uint32_t dot = popsum<D>(mulD<D>(va[u].z, pv0.x)) + popsum<D>(mulD<D>(vb[u].z, pv1.x));   // Z[:,i] . xs (old Z)
        if (act && has0) M.stB(i, w0, XZ{addD<D>(va[u].x, smulD<D>(pv0.x, fi)), addD<D>(va[u].z, smulD<D>(pv0.z, fi))});
        if (act && has1) M.stB(i, w1, XZ{addD<D>(vb[u].x, smulD<D>(pv1.x, fi)), addD<D>(vb[u].z, smulD<D>(pv1.z, fi))});
        for (int off = 1; off < gb; off <<= 1) dot += T.shfl_xor(dot, off);
        if (act && sub == 0 && i < np) {                                  // (a destabilizer's phase is never read)
          // P_i += f*ps + po*((Z_i.xs)*f + sd*f(f-1)/2*po)      (tableau_prime.py:310-312,317-319)
          const uint32_t ph = M.ldp(i);
          M.stp(i, (ph + fi * ps + PO * ((dot * fi + sd * ((fi * (fi - 1u)) >> 1) * PO) % D)) % ORDER);
        }
      }
      k0 += UB * ng;
      if (k0 >= total) break;
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int k = k0 + u * ng + grp;
        ent[u] = (k < total) ? M.ldl(k) | 0x8000u : 0u;
        va[u] = vb[u] = zero;
        if ((ent[u] & 0x8000u) && has0) va[u] = M.ldB(ent[u] & 0xFFFu, w0);
        if ((ent[u] & 0x8000u) && has1) vb[u] = M.ldB(ent[u] & 0xFFFu, w1);
      }
    }
    T.sync();
    // X[r,:] += xs_r * f on the pivot's X support; column p <- 0 (stabilizer p becomes Z_q), column np+p <- xs
    // (destabilizer p becomes the pivot), including rows where only the old destabilizer had an entry.
    // QX rows: gq = 2 gb lanes per row, lane `subq` holds lane words 2 subq and 2 subq + 1.
    {
      const bool lead = grp == 0;                                        // group 0 lists the rows
      const int totq = gm_compact<LPS>(T, M, (lead && has0) ? (pv0.x.l | pv0.x.h | dx0.l | dx0.h) : 0u, pv0.x, w0,
                                       (lead && has1) ? (pv1.x.l | pv1.x.h | dx1.l | dx1.h) : 0u, pv1.x, w1);
      const int gq = 2 << M.gs_shift, subq = lane & (gq - 1), grpq = lane >> (M.gs_shift + 1), ngq = LPS >> (M.gs_shift + 1);
      const int j0 = 2 * subq, j1 = 2 * subq + 1;
      const bool hq = j0 < Wb;                                            // Wb is even: j1 < Wb as well
      const E f0{T.shfl(f.l, j0 & (LPS - 1)), T.shfl(f.h, j0 & (LPS - 1))}, f1{T.shfl(f.l, j1 & (LPS - 1)), T.shfl(f.h, j1 & (LPS - 1))};
      const bool fixp0 = j0 == jp, fixp1 = j1 == jp, fixd0 = j0 == Wq + jp, fixd1 = j1 == Wq + jp;
      const bool fix = fixp0 || fixp1 || fixd0 || fixd1, any_f = (f0.l | f0.h | f1.l | f1.h) != 0u;
      for (int k0 = 0; k0 < totq; k0 += UQ * ngq) {
        E xa[UQ], xb[UQ];
        uint32_t en[UQ];
#pragma unroll
        for (int u = 0; u < UQ; ++u) {
          const int k = k0 + u * ngq + grpq;
          en[u] = 0u;
          xa[u] = xb[u] = ezero;
          if (k < totq && hq) {
            const uint32_t t = M.ldl(k);
            if (((t >> 12) && any_f) || fix) { en[u] = t | 0x8000u; M.ldqx2(t & 0xFFFu, j0, xa[u], xb[u]); }
          }
        }
#pragma unroll
        for (int u = 0; u < UQ; ++u) {
          if (en[u] & 0x8000u) {
            const uint32_t sv = (en[u] >> 12) & 3u;
            E na = addD<D>(xa[u], smulD<D>(f0, sv)), nb = addD<D>(xb[u], smulD<D>(f1, sv));
            if (fixp0) na = setbit2(na, bp, 0u);
            if (fixp1) nb = setbit2(nb, bp, 0u);
            if (fixd0) na = setbit2(na, bp, sv);
            if (fixd1) nb = setbit2(nb, bp, sv);
            M.stqx2(en[u] & 0xFFFu, j0, na, nb);
          }
        }
      }
    }
    // destabilizer p <- (xs, zs, ps); stabilizer p <- Z_q with phase -m*po   (tableau_prime.py:323-333)
    if (grp == 0) {
      XZ u0 = zero, u1 = zero;
      if (w0 == (q >> 5)) u0.z.l = 1u << (q & 31);
      if (w1 == (q >> 5)) u1.z.l = 1u << (q & 31);
      if (has0) { M.stB(np + piv, w0, pv0); M.stB(piv, w0, u0); }
      if (has1) { M.stB(np + piv, w1, pv1); M.stB(piv, w1, u1); }
    }
    if (lane == 0) {
      M.stp(np + piv, ps);
      M.stp(piv, (ORDER - draw * PO) % ORDER);
    }
    rec = draw;        // replayed or Philox, resolved when the op was fetched (reference: random.choice, :332)
  } else {
    // ---- deterministic branch (tableau_prime.py:336-363): product of the stabilizers i with f_i = destab X[q,i],
    // in increasing i, accumulated 64 qudits per lane (every group computes the same); rows requested UB at a time ----
    E fd = ezero;
    if (lane >= Wq && lane < Wb) fd = xq;
    const E fsh{T.shfl_down(fd.l, Wq & (LPS - 1)), T.shfl_down(fd.h, Wq & (LPS - 1))};   // lane j: destabilizer word j
    const E fl = (lane < Wq && Wq < LPS) ? fsh : ezero;
    const int total = gm_compact<LPS>(T, M, fl.l | fl.h, fl, lane, 0u, ezero, 0);   // ordered (stabilizer | f << 12)
    E az0 = ezero, az1 = ezero;
    uint32_t cross = 0, sdg = 0, a1 = 0;
    for (int k0 = 0; k0 < total; k0 += UB) {
      XZ va[UB], vb[UB];
      uint32_t ent[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        ent[u] = (k0 + u < total) ? M.ldl(k0 + u) : 0u;
        va[u] = vb[u] = zero;
        if (k0 + u < total && has0) va[u] = M.ldB(ent[u] & 0xFFFu, w0);
        if (k0 + u < total && has1) vb[u] = M.ldB(ent[u] & 0xFFFu, w1);
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        if (k0 + u >= total) break;                                        // tile-uniform
        const uint32_t fi = ent[u] >> 12;
        cross += popsum<D>(mulD<D>(az0, smulD<D>(va[u].x, fi))) + popsum<D>(mulD<D>(az1, smulD<D>(vb[u].x, fi)));
        az0 = addD<D>(az0, smulD<D>(va[u].z, fi));                         // running ancilla
        az1 = addD<D>(az1, smulD<D>(vb[u].z, fi));
        if (D == 3 && fi == 2u) sdg += popsum<D>(mulD<D>(va[u].x, va[u].z)) + popsum<D>(mulD<D>(vb[u].x, vb[u].z));
        a1 += fi * M.ldp(ent[u] & 0xFFFu);
      }
    }
    uint32_t part = cross + PO * sdg;
    for (int off = 1; off < gb; off <<= 1) part += T.shfl_xor(part, off);
    const uint32_t ap = (a1 % ORDER + PO * (part % D)) % ORDER;
    const uint32_t outcome = (D == 3) ? (3u - ap) % 3u : (((ap + 1u) >> 1) & 1u);   // (-ap // po) % d  (:362)
    rec = outcome | SDIMB_REC_DET;
  }
  T.sync();
  return rec;
}

// ---- the kernel of a trailing measurement run: one tile per shot --------------------------------------------------
#ifndef SDIMB_RUN_THREADS
#define SDIMB_RUN_THREADS 128
#endif
#ifndef SDIMB_RUN_CTAS
#define SDIMB_RUN_CTAS 8
#endif
constexpr int kRunThreads = SDIMB_RUN_THREADS;    // threads per CTA: kRunThreads / LPS shots in flight
constexpr int kRunCtasPerSm = SDIMB_RUN_CTAS;     // 32 warps per SM at 64 registers
inline int run_lps(int n) {                        // lanes per shot: the power of two >= Wb, at least 4
  const int np = (n + 31) / 32 * 32, Wb = 2 * np / 32;
  int lps = 4;
  while (lps < Wb) lps *= 2;
  if (const char* env = std::getenv("SDIMB_RUN_LPS")) {                  // developer knob (A/B timings)
    const int m = std::atoi(env);
    if (m >= lps && (m == 4 || m == 8 || m == 16 || m == 32)) lps = m;
  }
  return lps;
}
inline size_t run_smem_bytes(int n) {  // per CTA: list (2np uint16) + ph8 (2np bytes) per tile
  const size_t np = (size_t)(n + 31) / 32 * 32;
  return (size_t)(kRunThreads / run_lps(n)) * (4 * np + 2 * np);
}
inline size_t run_shot_stride_words(int n, int d);
inline size_t run_slab_words(int n, int d) {   // B + QX of one tile
  const size_t np = (size_t)(n + 31) / 32 * 32, Wq = np / 32, Wb = 2 * Wq, EW = (d == 2) ? 2 : 4;
  return (np * Wq * EW + np * Wq * (EW / 2) + (size_t)n * Wb * (EW / 2) + 7) & ~(size_t)7;   // B, Bd (X only), QX
}

inline size_t run_shot_stride_words(int n, int d) {   // B + QX + phase planes of one shot (gm_per_shot)
  const size_t np = (size_t)(n + 31) / 32 * 32, Wb = 2 * np / 32;
  return (run_slab_words(n, d) + 2 * Wb + 7) & ~(size_t)7;
}

// PRE: the kernel that ran the gates left B / Bd / QX in place (KParams::gm_per_shot): no transposition code in this
// instantiation.
template <int D, bool IL, int LPS, bool PRE = false>
__global__ void __launch_bounds__(kRunThreads, kRunCtasPerSm) run_tail_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int TILES = kRunThreads / LPS;
  const Tile<LPS> T;
  const int lane = T.lane, tile = threadIdx.x / LPS;
  Geo<D, IL> G;
  G.n = p.n;
  G.np = (p.n + 31) / 32 * 32;
  G.Wb = 2 * G.np / 32;
  G.RS = Geo<D>::EW * (G.Wb + (IL ? 0 : 1));
  const int row_words = (p.n * G.RS + 3) & ~3;
  GMImg<D> M;
  M.n = G.n; M.np = G.np; M.Wq = G.np / 32; M.Wb = G.Wb;
  M.gs_shift = 0;
  while ((2 << M.gs_shift) < M.Wq) ++M.gs_shift;                        // lanes per B row: two entries per lane
  M.list = (uint32_t)__cvta_generic_to_shared(smem) + (uint32_t)(tile * 4 * G.np);
  M.ph8 = (uint32_t)__cvta_generic_to_shared(smem) + (uint32_t)(TILES * 4 * G.np + tile * 2 * G.np);
  M.place(p.gm_slab + ((int64_t)blockIdx.x * TILES + tile) * p.gm_slab_words);
  for (;;) {
    int64_t shot = 0;
    if (lane == 0) shot = (int64_t)atomicAdd(p.shot_counter, 1u);
    shot = T.shfl(shot, 0);
    if (shot >= p.shots) break;
    const uint2* ph;
    if (PRE || p.gm_per_shot) {  // gate_stream_kernel already left B, QX and the phase planes of this shot in place
      M.place(p.gm_slab + shot * p.gm_shot_stride_words);
      ph = reinterpret_cast<const uint2*>(M.B + p.gm_slab_words);
    } else {
      G.tab = p.plane_slab + shot * p.img_stride_words;          // the image the interpreter left for this shot
      ph = reinterpret_cast<const uint2*>(G.tab + row_words);
    }
    for (int g = lane; g < 2 * G.np; g += LPS) {
      const uint2 w = ph[g >> 5];
      M.stp(g, ((w.x >> (g & 31)) & 1u) | (((w.y >> (g & 31)) & 1u) << 1));
    }
    if (!PRE && !p.gm_per_shot) gm_transpose<D>(G, M, true, lane, LPS);
    T.sync();
    for (int64_t i0 = p.tail_start; i0 < p.n_ops; i0 += LPS) {
      int4 mine = make_int4(SDIMB_OP_I, 0, 0, 0);
      if (i0 + lane < p.n_ops) mine = __ldg(p.ops + i0 + lane);
      mine.x &= SDIMB_OP_MASK;
      const bool is_m = mine.x == SDIMB_OP_M;
      if (is_m) {                       // outcome this measurement takes if it is random (same draws as the interpreter)
        if (p.replay_meas) {
          mine.z = p.replay_meas[shot * p.n_meas + mine.w];
        } else {
          const uint64_t gshot = (uint64_t)(p.shot_offset + shot);
          const uint4 r = philox4x32((uint32_t)gshot, (uint32_t)(gshot >> 32), (uint32_t)mine.w, 0u, (uint32_t)p.seed,
                                     (uint32_t)(p.seed >> 32));
          mine.z = (int)__umulhi(r.x, (uint32_t)D);
        }
      }
      uint32_t todo = T.ballot(is_m);
      uint32_t myrec = 0;
      while (todo) {
        const int k = __ffs(todo) - 1;
        todo &= todo - 1;
        const int qn = todo ? T.shfl(mine.y, __ffs(todo) - 1) : -1;      // next measurement of this batch, if any
        const uint32_t rec = gm_measure<D, LPS>(T, M, T.shfl(mine.y, k), qn, (uint32_t)T.shfl(mine.z, k));
        if (lane == k) myrec = rec;
      }
      if (is_m) p.records[shot * p.rec_stride + mine.w] = (uint8_t)myrec;
    }
  }
}
