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
//       B   [2np generators][Wq = np/32 qudit words]  entries (x_l, x_h, z_l, z_h) / (x, z) for d = 2: rows contiguous
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
  uint32_t* B;
  uint32_t* QX;
  uint8_t* ph8;
  uint16_t* list;                                // [2np] scratch: compacted generator / row lists
  int n, np, Wq, Wb, gs_shift;                   // group of lanes per B row = 1 << gs_shift >= Wq
  __device__ __forceinline__ uint32_t* bent(int g, int w) const { return B + ((size_t)g * Wq + w) * EW; }
  __device__ __forceinline__ uint32_t* qent(int r, int j) const { return QX + ((size_t)r * Wb + j) * EX; }
  __device__ __forceinline__ XZ ldB(int g, int w) const {
    if (D == 3) { const uint4 v = *reinterpret_cast<const uint4*>(bent(g, w)); return XZ{E{v.x, v.y}, E{v.z, v.w}}; }
    const uint2 v = *reinterpret_cast<const uint2*>(bent(g, w));
    return XZ{E{v.x, 0u}, E{v.y, 0u}};
  }
  __device__ __forceinline__ E ldBx(int g, int w) const {
    if (D == 3) { const uint2 v = *reinterpret_cast<const uint2*>(bent(g, w)); return E{v.x, v.y}; }
    return E{bent(g, w)[0], 0u};
  }
  __device__ __forceinline__ void stB(int g, int w, XZ v) const {
    if (D == 3) *reinterpret_cast<uint4*>(bent(g, w)) = make_uint4(v.x.l, v.x.h, v.z.l, v.z.h);
    else *reinterpret_cast<uint2*>(bent(g, w)) = make_uint2(v.x.l, v.z.l);
  }
  __device__ __forceinline__ E ldqx(int r, int j) const {
    if (D == 3) { const uint2 v = *reinterpret_cast<const uint2*>(qent(r, j)); return E{v.x, v.y}; }
    return E{qent(r, j)[0], 0u};
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
    uint32_t* const a = G.entry(32 * w, j) + pl;                       // + k * RS, rows k < n - 32 w
    uint32_t* const b = M.bent(32 * j, w) + pl;                        // + k * Wq * EW
    const int alim = min(32, G.n - 32 * w);
    uint32_t* const src = to_gm ? a : b;
    uint32_t* const dst = to_gm ? b : a;
    const int ss = to_gm ? G.RS : M.Wq * EW, ds = to_gm ? M.Wq * EW : G.RS;
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

// set bits of m (this lane's word of a 32-lane-wide mask) -> list entries (32 * lane + bit) | value << 12, in order;
// returns the total.  One warp.
__device__ __forceinline__ int gm_compact(uint16_t* list, int lane, uint32_t m, E val) {
  const int cnt = __popc(m);
  int incl = cnt;
#pragma unroll
  for (int d2 = 1; d2 < 32; d2 <<= 1) {
    const int o = __shfl_up_sync(0xFFFFFFFFu, incl, d2);
    if (lane >= d2) incl += o;
  }
  int pos = incl - cnt;
  while (m) {
    const int b = __ffs(m) - 1;
    m &= m - 1;
    list[pos++] = (uint16_t)((32 * lane + b) | (bit2(val, b) << 12));
  }
  const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
  __syncwarp();
  return total;
}

// Measurement of qudit q on the generator-major image, one warp (tableau_prime.py:262-363; same closed forms as
// p_measure).  Needs Wb <= 32: lane j holds lane word j of row q, lane w holds qudit word w of a generator row.
template <int D>
__device__ __noinline__ uint32_t gm_measure(const GMImg<D> M, const int q, const uint32_t draw) {
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  constexpr uint32_t PO = (D == 2) ? 2u : 1u, ORDER = D * PO;
  const int lane = threadIdx.x & 31;
  const int Wq = M.Wq, Wb = M.Wb, np = M.np;
  const XZ zero{E{0u, 0u}, E{0u, 0u}};
  E xq{0u, 0u};
  if (lane < Wb) xq = M.ldqx(q, lane);
  const uint32_t nz = xq.l | xq.h;
  // pivot: first stabilizer lane with an X component on q (tableau_prime.py:273-283)
  const uint32_t best = (lane < Wq && nz) ? 32u * lane + (uint32_t)(__ffs(nz) - 1) : kNoPivot;
  const uint32_t piv = __reduce_min_sync(FULL, best);
  uint32_t rec;
  if (piv != kNoPivot) {
    // ---- random branch (tableau_prime.py:294-334, exponentiate :365-380 folded in) ----
    const int jp = piv >> 5, bp = piv & 31;
    const uint32_t e = (D == 3) ? bit2(E{__shfl_sync(FULL, xq.l, jp), __shfl_sync(FULL, xq.h, jp)}, bp) : 1u;
    E f = negD<D>(xq);                                   // f_i = -X[q,i]; the pivot and its destabilizer are replaced below
    if (lane == jp || lane == Wq + jp) { f.l &= ~(1u << bp); f.h &= ~(1u << bp); }
    XZ pv = zero;
    E dxo{0u, 0u};                                       // X support of the destabilizer that is about to be overwritten
    if (lane < Wq) { pv = M.ldB(piv, lane); dxo = M.ldBx(np + piv, lane); }
    const uint32_t sd_raw = __reduce_add_sync(FULL, popsum<D>(mulD<D>(pv.x, pv.z))) % D;
    const uint32_t ps_old = M.ph8[piv];
    const uint32_t ps = (ps_old * e + PO * ((sd_raw * ((e * (e - 1u)) >> 1)) % D)) % ORDER;
    const uint32_t sd = (sd_raw * e * e) % D;
    if (D == 3 && e == 2u) { pv.x = neg3(pv.x); pv.z = neg3(pv.z); }    // pivot <- pivot^e: (xs, zs, ps)
    // row_i += f_i * pivot for every generator with a factor; 32 >> gs_shift generators per pass
    {
      const int total = gm_compact(M.list, lane, f.l | f.h, f);
      const int gs = 1 << M.gs_shift, sub = lane & (gs - 1), grp = lane >> M.gs_shift, ng = 32 >> M.gs_shift;
      const XZ pvs{E{__shfl_sync(FULL, pv.x.l, sub), __shfl_sync(FULL, pv.x.h, sub)},
                   E{__shfl_sync(FULL, pv.z.l, sub), __shfl_sync(FULL, pv.z.h, sub)}};
      for (int k0 = 0; k0 < total; k0 += ng) {
        const int k = k0 + grp;
        const bool act = k < total && sub < Wq;
        const uint32_t ent = act ? (uint32_t)M.list[k] : 0u;
        const int i = ent & 0xFFFu;
        const uint32_t fi = ent >> 12;
        XZ v = zero;
        if (act) v = M.ldB(i, sub);
        uint32_t dot = popsum<D>(mulD<D>(v.z, pvs.x));                   // Z[:,i] . xs (old Z)
        if (act) M.stB(i, sub, XZ{addD<D>(v.x, smulD<D>(pvs.x, fi)), addD<D>(v.z, smulD<D>(pvs.z, fi))});
        for (int off = 1; off < gs; off <<= 1) dot += __shfl_xor_sync(FULL, dot, off);
        if (act && sub == 0) {
          // P_i += f*ps + po*((Z_i.xs)*f + sd*f(f-1)/2*po)      (tableau_prime.py:310-312,317-319)
          const uint32_t ph = M.ph8[i];
          M.ph8[i] = (uint8_t)((ph + fi * ps + PO * (((dot % D) * fi + sd * ((fi * (fi - 1u)) >> 1) * PO) % D)) % ORDER);
        }
      }
      __syncwarp();
    }
    // X[r,:] += xs_r * f on the pivot's X support; column p <- 0 (stabilizer p becomes Z_q), column np+p <- xs
    // (destabilizer p becomes the pivot), including rows where only the old destabilizer had an entry
    {
      const E xs = pv.x;
      const uint32_t mr = (lane < Wq) ? (xs.l | xs.h | dxo.l | dxo.h) : 0u;
      const int total = gm_compact(M.list, lane, mr, xs);
      const int gsq = 2 << M.gs_shift, sub = lane & (gsq - 1), grp = lane >> (M.gs_shift + 1), ng = 16 >> M.gs_shift;
      const E fq{__shfl_sync(FULL, f.l, sub), __shfl_sync(FULL, f.h, sub)};
      const bool fix_p = sub == jp, fix_d = sub == Wq + jp;
      for (int k0 = 0; k0 < total; k0 += ng) {
        const int k = k0 + grp;
        if (k < total && sub < Wb) {
          const uint32_t ent = M.list[k];
          const int r = ent & 0xFFFu;
          const uint32_t s = ent >> 12;
          if ((s && (fq.l | fq.h)) || fix_p || fix_d) {
            E x = M.ldqx(r, sub);
            x = addD<D>(x, smulD<D>(fq, s));
            if (fix_p) x = setbit2(x, bp, 0u);
            if (fix_d) x = setbit2(x, bp, s);
            M.stqx(r, sub, x);
          }
        }
      }
    }
    // destabilizer p <- (xs, zs, ps); stabilizer p <- Z_q with phase -m*po   (tableau_prime.py:323-333)
    if (lane < Wq) {
      M.stB(np + piv, lane, pv);
      XZ unit = zero;
      if (lane == (q >> 5)) unit.z.l = 1u << (q & 31);
      M.stB(piv, lane, unit);
    }
    if (lane == 0) {
      M.ph8[np + piv] = (uint8_t)ps;
      M.ph8[piv] = (uint8_t)((ORDER - draw * PO) % ORDER);
    }
    rec = draw;        // replayed or Philox, resolved when the op was fetched (reference: random.choice, :332)
  } else {
    // ---- deterministic branch (tableau_prime.py:336-363): product of the stabilizers i with f_i = destab X[q,i],
    // in increasing i, accumulated 32 qudits per lane ----
    uint32_t wm = __ballot_sync(FULL, nz != 0u) >> Wq;                  // destabilizer lane words with a factor
    E az{0u, 0u};
    uint32_t cross = 0, sdg = 0, a1 = 0;
    while (wm) {
      const int j = __ffs(wm) - 1;
      wm &= wm - 1;
      const uint32_t fl = __shfl_sync(FULL, xq.l, Wq + j), fh = __shfl_sync(FULL, xq.h, Wq + j);
      uint32_t mm = fl | fh;
      while (mm) {
        const int b = __ffs(mm) - 1;
        mm &= mm - 1;
        const int i = 32 * j + b;
        const uint32_t fi = ((fl >> b) & 1u) | (((fh >> b) & 1u) << 1);
        XZ v = zero;
        if (lane < Wq) v = M.ldB(i, lane);
        cross += popsum<D>(mulD<D>(az, smulD<D>(v.x, fi)));            // ancilla_z . (f * x_i), running ancilla
        az = addD<D>(az, smulD<D>(v.z, fi));
        if (D == 3 && fi == 2u) sdg += popsum<D>(mulD<D>(v.x, v.z));    // x_i . z_i * f(f-1)/2
        a1 += fi * M.ph8[i];
      }
    }
    const uint32_t part = __reduce_add_sync(FULL, cross + PO * sdg) % D;
    const uint32_t ap = (a1 % ORDER + PO * part) % ORDER;
    const uint32_t outcome = (D == 3) ? (3u - ap) % 3u : (((ap + 1u) >> 1) & 1u);   // (-ap // po) % d  (:362)
    rec = outcome | SDIMB_REC_DET;
  }
  __syncwarp();
  return rec;
}

// ---- the kernel of a trailing measurement run: one warp per shot ------------------------------------------------
constexpr int kRunWarps = 4;          // warps (= shots in flight) per CTA
constexpr int kRunCtasPerSm = 8;      // 32 warps per SM at 64 registers
inline size_t run_smem_bytes(int n) {  // per CTA: list (2np uint16) + ph8 (2np bytes) per warp
  const size_t np = (size_t)(n + 31) / 32 * 32;
  return (size_t)kRunWarps * (4 * np + 2 * np);
}
inline size_t run_slab_words(int n, int d) {   // B + QX of one warp
  const size_t np = (size_t)(n + 31) / 32 * 32, Wq = np / 32, Wb = 2 * Wq, EW = (d == 2) ? 2 : 4;
  return (2 * np * Wq * EW + (size_t)n * Wb * (EW / 2) + 7) & ~(size_t)7;
}

template <int D, bool IL>
__global__ void __launch_bounds__(32 * kRunWarps, kRunCtasPerSm) run_tail_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Geo<D, IL> G;
  G.n = p.n;
  G.np = (p.n + 31) / 32 * 32;
  G.Wb = 2 * G.np / 32;
  G.RS = Geo<D>::EW * (G.Wb + (IL ? 0 : 1));
  const int row_words = (p.n * G.RS + 3) & ~3;
  GMImg<D> M;
  M.n = G.n; M.np = G.np; M.Wq = G.np / 32; M.Wb = G.Wb;
  M.gs_shift = 0;
  while ((1 << M.gs_shift) < M.Wq) ++M.gs_shift;
  M.list = reinterpret_cast<uint16_t*>(smem) + (size_t)warp * 2 * G.np;
  M.ph8 = smem + (size_t)kRunWarps * 4 * G.np + (size_t)warp * 2 * G.np;
  M.B = p.gm_slab + ((int64_t)blockIdx.x * kRunWarps + warp) * p.gm_slab_words;
  M.QX = M.B + (size_t)2 * G.np * M.Wq * GMImg<D>::EW;
  for (;;) {
    int64_t shot = 0;
    if (lane == 0) shot = (int64_t)atomicAdd(p.shot_counter, 1u);
    shot = __shfl_sync(FULL, shot, 0);
    if (shot >= p.shots) break;
    G.tab = p.plane_slab + shot * p.img_stride_words;          // the image the interpreter left for this shot
    {
      const uint2* ph = reinterpret_cast<const uint2*>(G.tab + row_words);
      for (int j = 0; j < G.Wb; ++j) {
        const uint2 w = ph[j];
        M.ph8[32 * j + lane] = (uint8_t)(((w.x >> lane) & 1u) | (((w.y >> lane) & 1u) << 1));
      }
    }
    gm_transpose<D>(G, M, true, lane, 32);
    __syncwarp();
    for (int64_t i0 = p.tail_start; i0 < p.n_ops; i0 += 32) {
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
      uint32_t todo = __ballot_sync(FULL, is_m);
      uint32_t myrec = 0;
      while (todo) {
        const int k = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t rec = gm_measure<D>(M, __shfl_sync(FULL, mine.y, k), (uint32_t)__shfl_sync(FULL, mine.z, k));
        if (lane == k) myrec = rec;
      }
      if (is_m) p.records[shot * p.rec_stride + mine.w] = (uint8_t)myrec;
    }
  }
}
