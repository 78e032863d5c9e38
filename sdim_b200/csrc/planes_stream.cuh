// Gate-stream kernel of the bit-plane interpreter (included by planes.cuh inside namespace planes).
//
// The general interpreter (interp_planes_kernel) pays ~150 warp instructions per executed gate: every warp fetches and
// decodes the whole op stream, filters its share, broadcasts each op by shuffles, resolves noise draws at fetch and
// runs one gate with half of its lanes (Wb = 16 lane words at n = 256).  For the part of a stream that holds no
// collective op — everything in front of the trailing measurement run that run_tail_kernel executes — the host
// compiles the layers of sdimb_schedule into one pre-decoded stream PER WARP (sdimb_gate_stream):
//   * a "wide op" = GPW = 32 / Wb gates of ONE family on pairwise disjoint rows (same layer), one row of the stream
//     per lane group: lane group g of the warp loads its own row (family | flags, qudit a, qudit b) and runs its gate
//     on its lane word — no filtering, no shuffles, all 32 lanes busy;
//   * layers (ASAP over the gates that write rows, computed by the compiler itself) are separated by SYNC rows (one
//     block barrier each, same count in every warp's stream);
//   * the Pauli gates X, X^-1, Z, Z^-1 (a third of a random Clifford circuit) are not in the stream either: a Pauli
//     commutes through a Clifford gate into another Pauli (Pauli-frame propagation, the mechanism of the reference's
//     simulate_frame, sdim/program.py:82-120, applied to gates instead of errors), so the host pushes every one of
//     them BACK to the start of the segment, where they merge into one X^a Z^b per qudit.  The tableau update of a
//     Pauli, phase += po (b x - a z), is the symplectic form, which the gate rules preserve: exact.  On |0...0> that
//     initial Pauli is a closed form (stabilizer q: -a_q po, destabilizer q: +b_q po), on a loaded store n phase passes;
//   * N1 noise ops are not in the stream at all: a pre-pass per shot evaluates every event of the segment (Philox or
//     replay, all threads of the CTA, fully packed) and sets one bit per FIRED event in shared memory; the SYNC row
//     that opens a layer carries that layer's range of the noise table and the index of its "some event fired" bit:
//     if that is set (rare), warp 0 applies the fired events and the CTA takes a second barrier before the layer's
//     gates.  An N1 belongs to the start of the layer of the NEXT writer of its row, so it sees the row as the program
//     order says and costs no layer of its own.
// SM = true keeps the shot's image in SHARED memory while the gates run (a gate is then one LDS.128 + LOP3s + one
// STS.128 per lane, ~30 cycles instead of an L1/L2 round trip, and neither L2 nor HBM sees the 2 000 row updates of
// a shot) and copies it out once, coalesced, to the per-shot image run_tail_kernel picks up; SM = false runs on that
// global image directly (tableaus whose image does not fit in shared memory).  Same image format and phase planes as
// the general interpreter (one accumulator per warp AND lane group, folded at the end of the segment).
//
// Stream layout (int4 rows, built by sdimb_gate_stream in sdimb.cu):
//   row 0            (magic, GPW, NW, n_tab)            n_tab = N1 events of the segment
//   row 1            (tab_base, total rows, layers, n_pauli)
//   row 2            (row stride of the image in bytes, interleaved image?, 0, 0)
//   rows 3 .. 3+NW   (first row of warp w's stream, entries, 0, 0)
//   warp streams     entries of GPW rows each: SYNC (noise range [y, z) of the layer it opens), wide ops (family | flags,
//                    BYTE offset of row a, of row b), END ... (read-ahead padding)
//   noise table      (event slot, qudit, 0, 0) per N1 op, in layer order
//   initial Paulis   (qudit, a, b, 0) for every qudit with (a, b) != (0, 0)
#pragma once

// families of the gate stream (row.x & 0xFF); flags above
enum { GS_END = 0, GS_SYNC = 1, GS_H = 2, GS_P = 3, GS_CNOT = 4, GS_CZ = 5, GS_SWAP = 6 };
#define GS_INV 0x100          // inverse gate (H_INV, P_INV, CNOT_INV, CZ_INV)
#define GS_ON 0x200           // this lane group has a gate (padding rows of a wide op do not)
#define GS_LAYER_SHIFT 12     // SYNC rows: ordinal of the layer among the layers that have N1 events
constexpr int kGateStreamMagic = 0x47533033;   // "GS03"
constexpr int kGateStreamMaxNoise = 1 << 16;   // fired bits live in shared memory: 8 KB at most
constexpr int kGateStreamMaxWarps = 8;
constexpr int kGateStreamHeaderRows = 3;
constexpr int kGateStreamPadRows = 96;        // END rows behind every warp stream: two 32-row batches of read-ahead

#ifndef SDIMB_GS_CTAS
#define SDIMB_GS_CTAS 10      // resident CTAs per SM the global-image form is compiled for (4 warps each)
#endif

inline int gate_stream_gpw(int n) {
  const int np = (n + 31) / 32 * 32, Wb = 2 * np / 32;
  return Wb <= 32 ? 32 / Wb : 0;                // 0: rows wider than a warp are not supported
}
// shared memory besides the image: phase accumulators, fired bits, the claimed shot
inline size_t gate_stream_smem_bytes(int n, int64_t n_noise, int nw) {
  const size_t np = (size_t)(n + 31) / 32 * 32, Wb = 2 * np / 32;
  const size_t gpw = Wb <= 32 ? 32 / Wb : 1;
  return 8 * (size_t)nw * gpw * Wb + 2 * 4 * (((size_t)n_noise + 31) / 32) + 16;   // fired bits per event and per layer
}

// This lane's window on the image: byte offsets of rows come from the stream.  SM: explicit shared-space accesses with
// a 32-bit address computed once per shot (through a generic pointer the compiler rebuilt the CTA's shared window
// address — S2R SR_CgaCtaId + LEA — in every iteration of the gate loop).
template <int D, bool SM>
struct RowAcc {
  uint32_t s;       // shared address of lane word j of row 0 (SM)
  uint8_t* g;       // generic pointer to the same (global image)
  __device__ __forceinline__ XZ ld(int off) const {
    if (SM) {
      if (D == 3) {
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(s + off) : "memory");
        return XZ{E{v.x, v.y}, E{v.z, v.w}};
      }
      uint2 v;
      asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(s + off) : "memory");
      return XZ{E{v.x, 0u}, E{v.y, 0u}};
    }
    if (D == 3) { const uint4 v = *reinterpret_cast<const uint4*>(g + off); return XZ{E{v.x, v.y}, E{v.z, v.w}}; }
    const uint2 v = *reinterpret_cast<const uint2*>(g + off);
    return XZ{E{v.x, 0u}, E{v.y, 0u}};
  }
  __device__ __forceinline__ void st(int off, XZ v) const {
    if (SM) {
      if (D == 3) asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(s + off), "r"(v.x.l), "r"(v.x.h), "r"(v.z.l), "r"(v.z.h) : "memory");
      else asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(s + off), "r"(v.x.l), "r"(v.z.l) : "memory");
    } else if (D == 3) *reinterpret_cast<uint4*>(g + off) = make_uint4(v.x.l, v.x.h, v.z.l, v.z.h);
    else *reinterpret_cast<uint2*>(g + off) = make_uint2(v.x.l, v.z.l);
  }
  __device__ __forceinline__ void stx(int off, E x) const {
    if (SM) {
      if (D == 3) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(s + off), "r"(x.l), "r"(x.h) : "memory");
      else asm volatile("st.shared.u32 [%0], %1;" ::"r"(s + off), "r"(x.l) : "memory");
    } else if (D == 3) *reinterpret_cast<uint2*>(g + off) = make_uint2(x.l, x.h);
    else *reinterpret_cast<uint32_t*>(g + off) = x.l;
  }
  __device__ __forceinline__ void stz(int off, E z) const {
    if (SM) {
      if (D == 3) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(s + off + 8), "r"(z.l), "r"(z.h) : "memory");
      else asm volatile("st.shared.u32 [%0], %1;" ::"r"(s + off + 4), "r"(z.l) : "memory");
    } else if (D == 3) *reinterpret_cast<uint2*>(g + off + 8) = make_uint2(z.l, z.h);
    else *reinterpret_cast<uint32_t*>(g + off + 4) = z.l;
  }
};

template <int D, bool IL, bool SM>
__global__ void __launch_bounds__(SM ? 32 * kGateStreamMaxWarps : 32 * SDIMB_SCHED_WARPS, SM ? 3 : SDIMB_GS_CTAS)
gate_stream_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int EW = Geo<D>::EW;
  constexpr uint32_t PO = (D == 2) ? 2u : 1u, ORDER = D * PO;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5, NT = blockDim.x;
  Geo<D, IL> G;
  G.n = p.n;
  G.np = (p.n + 31) / 32 * 32;
  G.Wb = 2 * G.np / 32;
  G.RS = EW * (G.Wb + (IL ? 0 : 1));
  const int Wb = G.Wb, row_words = (p.n * G.RS + 3) & ~3;
  const int4* const S = reinterpret_cast<const int4*>(p.gate_stream);
  const int4 hdr = __ldg(S), misc = __ldg(S + 1), geo = __ldg(S + 2);
  // not a stream compiled for this launch shape / image (a caller mixing library builds or developer knobs between
  // sdimb_gate_stream and sdimb_run): fail loudly — the launch ends in a CUDA error, never in silent garbage records
  if (hdr.x != kGateStreamMagic || hdr.z != NW || geo.x != 4 * G.RS || geo.y != (IL ? 1 : 0)) __trap();
  const int GPW = hdr.y, n_tab = hdr.w, tab_base = misc.x, n_pauli = misc.w;
  const int gsub = lane / Wb, j = lane - gsub * Wb;
  const bool lane_on = gsub < GPW;
  const int g = lane_on ? gsub : 0;
  uint32_t* const img = reinterpret_cast<uint32_t*>(smem);                  // SM: the shot's image (row_words words)
  uint2* const acc = reinterpret_cast<uint2*>(smem + (SM ? 4 * (size_t)row_words : 0));   // [NW * GPW][Wb]
  uint32_t* const fired = reinterpret_cast<uint32_t*>(acc + NW * GPW * Wb);  // [(n_tab + 31) / 32]
  uint32_t* const lfired = fired + (n_tab + 31) / 32;                        // [(n_tab + 31) / 32] one bit per noisy layer
  int* const next = reinterpret_cast<int*>(lfired + (n_tab + 31) / 32);      // [2]
  uint2* const pacc = acc + (warp * GPW + g) * Wb + j;
  int joff;                                                                  // word offset of lane word j inside a row
  { int e = j; if (IL) { const int h = G.np >> 5; e = (j < h) ? 2 * j : 2 * (j - h) + 1; } joff = e * EW; }
  const int4* const mystream = S + __ldg(S + kGateStreamHeaderRows + warp).x;
  const int EPB = 32 / GPW;

  for (int round = 0;; ++round) {
    if (tid == 0) next[round & 1] = (int)atomicAdd(p.shot_counter, 1u);
    __syncthreads();
    const int64_t shot = next[round & 1];
    if (shot >= p.shots) break;
    uint32_t* const gimg = p.plane_slab + shot * p.img_stride_words;          // the image run_tail_kernel picks up
    G.tab = SM ? img : gimg;
    uint8_t* const tabj = reinterpret_cast<uint8_t*>(G.tab + joff);            // this lane's word of row 0
    RowAcc<D, SM> R;
    R.g = tabj;
    R.s = SM ? (uint32_t)__cvta_generic_to_shared(tabj) : 0u;
    // ---- |0...0>, or pack from the uint8 store (same as interp_planes_kernel) ----
    for (int i = tid; i < row_words / 4; i += NT) reinterpret_cast<uint4*>(G.tab)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < NW * GPW * Wb; i += NT) acc[i] = make_uint2(0u, 0u);
    for (int i = tid; i < 2 * ((n_tab + 31) / 32); i += NT) fired[i] = 0u;     // fired and lfired
    __syncthreads();
    G.ph_base = acc;
    G.pacc = acc;
    if (p.flags & SDIMB_FRESH) {
      for (int q = tid; q < p.n; q += NT) {
        G.setz(q, q, 1u);              // stabilizer q = Z_q
        G.setx(q, G.np + q, 1u);       // destabilizer q = X_q
      }
      // the segment's Pauli gates, merged into one X^a Z^b per qudit in front of it: on |0...0> the phase of stabilizer
      // q (= Z_q) becomes -a po, that of destabilizer q (= X_q) +b po.  Distinct bits of accumulator 0.
      uint32_t* const acc0 = reinterpret_cast<uint32_t*>(acc);
      for (int t = tid; t < n_pauli; t += NT) {
        const int4 row = __ldg(S + tab_base + n_tab + t);                // (qudit, a, b, -)
        const uint32_t vs = (ORDER - (uint32_t)row.y * PO) % ORDER, vd = ((uint32_t)row.z * PO) % ORDER;
        const int ls = row.x, ld = G.np + row.x;
        if (vs & 1u) atomicOr(acc0 + 2 * (ls >> 5), 1u << (ls & 31));
        if (vs & 2u) atomicOr(acc0 + 2 * (ls >> 5) + 1, 1u << (ls & 31));
        if (vd & 1u) atomicOr(acc0 + 2 * (ld >> 5), 1u << (ld & 31));
        if (vd & 2u) atomicOr(acc0 + 2 * (ld >> 5) + 1, 1u << (ld & 31));
      }
    } else {
      const uint8_t* T8 = p.tab + shot * p.shot_bytes;
      for (int q = warp; q < p.n; q += NW) {
        const uint8_t* row8 = T8 + (int64_t)q * p.row_bytes;
        for (int jj = lane; jj < Wb; jj += 32) {
          XZ v{E{0u, 0u}, E{0u, 0u}};
#pragma unroll 1
          for (int b = 0; b < 32; ++b) {
            const int ln = 32 * jj + b;
            const int half = ln >= G.np, gg = half ? ln - G.np : ln;
            if (gg >= p.n) continue;
            const uint32_t xv = row8[half * p.np + gg], zv = row8[p.W + half * p.np + gg];
            v.x.l |= (xv & 1u) << b; v.x.h |= ((xv >> 1) & 1u) << b;
            v.z.l |= (zv & 1u) << b; v.z.h |= ((zv >> 1) & 1u) << b;
          }
          G.st(q, jj, v);
        }
      }
      for (int jj = tid; jj < Wb; jj += NT) {
        E ph0{0u, 0u};
#pragma unroll 1
        for (int b = 0; b < 32; ++b) {
          const int ln = 32 * jj + b;
          const int half = ln >= G.np, gg = half ? ln - G.np : ln;
          if (gg >= p.n) continue;
          const uint32_t v = T8[p.phase_off + half * p.np + gg];
          ph0.l |= (v & 1u) << b; ph0.h |= ((v >> 1) & 1u) << b;
        }
        G.stp(jj, ph0);
      }
    }
    // ---- noise pre-pass: one bit per fired event of the segment ----
    for (int t = tid; t < n_tab; t += NT) {
      const int4 row = __ldg(S + tab_base + t);                        // (event slot, qudit, noisy-layer ordinal, -)
      if (p_noise_event<D>(p, row.x, shot)) {
        atomicOr(&fired[t >> 5], 1u << (t & 31));
        atomicOr(&lfired[row.z >> 5], 1u << (row.z & 31));
      }
    }
    __syncthreads();
    // this lane's phase accumulator (warp, lane group, lane word) lives in a register while the gates run
    E ph{0u, 0u};
    { const uint2 o = *pacc; ph = E{o.x, o.y}; }
    if (!(p.flags & SDIMB_FRESH) && lane_on && gsub == 0) {     // the initial Pauli on a loaded store: n_pauli phase passes
      for (int t = warp; t < n_pauli; t += NW) {
        const int4 row = __ldg(S + tab_base + n_tab + t);
        const XZ v = G.ld_at(reinterpret_cast<const uint32_t*>(tabj + (size_t)row.x * (4 * G.RS)));
        if (D == 3) ph = add3(ph, add3(smul3(v.x, (uint32_t)row.z), smul3(v.z, (3u - (uint32_t)row.y) % 3u)));
        else ph.h ^= ((row.z & 1) ? v.x.l : 0u) ^ ((row.y & 1) ? v.z.l : 0u);
      }
    }

    // ---- this warp's stream ----
    // Rows are fetched 32 at a time, one per lane (EPB = 32 / GPW entries), two batches in flight: a row costs an
    // L2 round trip (the streams of a CTA's warps do not fit in what shared memory leaves of L1), a batch hides it
    // behind EPB wide ops.  Lane group g picks its row of entry e with three shuffles.  The stream is padded with END
    // rows so that the read-ahead stays inside it.
    const int4* fetch = mystream + lane;
    int4 cur = __ldg(fetch), ahead = __ldg(fetch + EPB * GPW);
    fetch += 2 * EPB * GPW;
    for (;;) {
      int src = g;
      for (int e = 0; e < EPB; ++e, src += GPW) {
        const int opx = __shfl_sync(0xFFFFFFFFu, cur.x, src);               // family: the same in every lane group
        const int opy = __shfl_sync(0xFFFFFFFFu, cur.y, src);
        const int opz = __shfl_sync(0xFFFFFFFFu, cur.z, src);
        const bool on = lane_on && (opx & GS_ON);
        const bool inv = (opx & GS_INV) != 0;
        switch (opx & 0xFF) {
        case GS_END: goto stream_done;
        case GS_SYNC: {
          __syncthreads();
          // N1 events of the layer that starts here: they act before its gates.  Almost always none fired (one bit per
          // layer, set by the pre-pass); if one did, warp 0 applies them — it reads rows and adds to its own phase
          // accumulator — and one more barrier keeps the layer's writers behind those reads.
          const int lay = opx >> GS_LAYER_SHIFT;
          if (opy < opz && ((lfired[lay >> 5] >> (lay & 31)) & 1u)) {
           if (warp == 0) {
            for (int t0 = opy & ~31; t0 < opz; t0 += 32) {
              uint32_t m = fired[t0 >> 5];
              if (t0 < opy) m &= ~((1u << (opy - t0)) - 1u);
              if (t0 + 32 > opz) m &= (1u << (opz - t0)) - 1u;
              while (m) {
                const int t = t0 + __ffs(m) - 1;
                m &= m - 1;
                const int4 row = __ldg(S + tab_base + t);
                const uint32_t ev = p_noise_event<D>(p, row.x, shot);     // recomputed: fired events are rare
                const uint32_t a = ev & 0xFFu, b = ev >> 8;
                if (lane_on && gsub == 0) {                              // Pauli X^a Z^b on qudit row.y: phase += po*(b*x - a*z)
                  const XZ v = G.ld_at(reinterpret_cast<const uint32_t*>(tabj + (size_t)row.y * (4 * G.RS)));
                  if (D == 3) ph = add3(ph, add3(smul3(v.x, b), smul3(v.z, (3u - a) % 3u)));
                  else ph.h ^= ((b & 1u) ? v.x.l : 0u) ^ ((a & 1u) ? v.z.l : 0u);
                }
              }
            }
           }
           __syncthreads();
          }
          break;
        }
        case GS_H:
          if (on) {
            const XZ v = R.ld(opy);
            if (D == 3) {
              ph = add3(ph, neg3(mul3(v.x, v.z)));                          // phase -= x*z
              R.st(opy, inv ? XZ{v.z, neg3(v.x)} : XZ{neg3(v.z), v.x});     // H: (x,z)<-(-z,x); H^-1: (x,z)<-(z,-x)
            } else {
              ph.h ^= v.x.l & v.z.l;                                        // phase += 2*x*z (mod 4); H == H^-1
              R.st(opy, XZ{v.z, v.x});
            }
          }
          break;
        case GS_P:
          if (on) {
            const XZ v = R.ld(opy);
            if (D == 3) {
              ph = add3(ph, inv ? E{0u, v.x.h} : E{v.x.h, 0u});             // phase +-= x(x-1)/2 = [x == 2]
              R.stz(opy, add3(v.z, inv ? neg3(v.x) : v.x));                 // z +-= x
            } else {
              if (inv) { const uint32_t borrow = ~ph.l & v.x.l; ph.l ^= v.x.l; ph.h ^= borrow; }   // phase -= x (mod 4)
              else { const uint32_t carry = ph.l & v.x.l; ph.l ^= v.x.l; ph.h ^= carry; }          // phase += x^2 = x
              R.stz(opy, E{v.z.l ^ v.x.l, 0u});
            }
          }
          break;
        case GS_CNOT:
          if (on) {
            const XZ va = R.ld(opy), vb = R.ld(opz);
            if (D == 3) {
              R.stx(opz, add3(vb.x, inv ? neg3(va.x) : va.x));              // x[t] +-= x[c]
              R.stz(opy, add3(va.z, inv ? vb.z : neg3(vb.z)));              // z[c] -+= z[t]
            } else {
              R.stx(opz, E{vb.x.l ^ va.x.l, 0u});
              R.stz(opy, E{va.z.l ^ vb.z.l, 0u});
            }
          }
          break;
        case GS_CZ:
          if (on) {
            const XZ va = R.ld(opy), vb = R.ld(opz);
            if (D == 3) {
              const E prod = mul3(va.x, vb.x);
              ph = add3(ph, inv ? neg3(prod) : prod);                       // phase +-= x[a]*x[b]
              R.stz(opy, add3(va.z, inv ? neg3(vb.x) : vb.x));
              R.stz(opz, add3(vb.z, inv ? neg3(va.x) : va.x));
            } else {
              ph.h ^= va.x.l & vb.x.l;
              R.stz(opy, E{va.z.l ^ vb.x.l, 0u});
              R.stz(opz, E{vb.z.l ^ va.x.l, 0u});
            }
          }
          break;
        case GS_SWAP:
          if (on) {
            const XZ va = R.ld(opy), vb = R.ld(opz);
            R.st(opy, vb);
            R.st(opz, va);
          }
          break;
        default: break;
        }
      }
      cur = ahead;
      ahead = __ldg(fetch);
      fetch += EPB * GPW;
    }
  stream_done:
    if (lane_on) *pacc = make_uint2(ph.l, ph.h);
    __syncthreads();
    uint2* out = reinterpret_cast<uint2*>(gimg + row_words);
    if (SM && p.gm_per_shot) {
      // the image leaves shared memory TRANSPOSED: B (generator-major) and QX, the form run_tail_kernel works on, straight
      // into this shot's slab — HBM never sees the qudit-major image, the tail kernel skips its transposition pass
      GMImg<D> M;
      M.n = G.n; M.np = G.np; M.Wq = G.np / 32; M.Wb = G.Wb;
      M.place(p.gm_slab + shot * p.gm_shot_stride_words);
      gm_transpose_out_smem<D>(G, M, tid, NT);
      out = reinterpret_cast<uint2*>(M.B + p.gm_slab_words);
    } else if (SM) {                                                         // ... or once as it is, coalesced
      for (int i = tid; i < row_words / 4; i += NT)
        reinterpret_cast<uint4*>(gimg)[i] = reinterpret_cast<const uint4*>(img)[i];
    }
    // ---- folded phase planes, for run_tail_kernel ----
    {
      for (int jj = tid; jj < Wb; jj += NT) {
        E a{0u, 0u};
        for (int w = 0; w < NW * GPW; ++w) {
          const uint2 o = acc[w * Wb + jj];
          a = (D == 3) ? add3(a, E{o.x, o.y}) : add4(a, E{o.x, o.y});
        }
        out[jj] = make_uint2(a.l, a.h);
      }
    }
    __syncthreads();
  }
}
