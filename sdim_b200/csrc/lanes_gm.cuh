// Generator-major measurement runs on uint8 lanes (any prime d <= 127) — the bytes counterpart of planes_gm.cuh.
//
// On the HBM store (include/sdimb.h: row q = X[q][0..W) | Z[q][0..W), one byte per entry) a measurement walks the
// pivot's COLUMN: one byte per row, one 32-byte DRAM sector per byte — the measurement-heavy headline shape moved 3.7x
// its algorithmic bytes that way (profiles/r1_ncu_headline_lanes_global.txt).  The run of M ops that ends a stream
// ("measure every qudit") is therefore executed by a second kernel, run_tail8_kernel, ONE WARP PER SHOT, on
//   B8   [W generators][2 nq bytes]   row g = x[0..nq) | z[0..nq) of generator g over the qudits (nq = n rounded up to 32):
//                                      built from the store by 32 x 4-byte register tiles and byte permutes, one slab per
//                                      resident warp in caller scratch
//   QX8  = the X halves of the store's own rows, updated in place (the Z halves go stale: the tableau is not kept)
//   ph8  the phase bytes, in shared memory
// so that every access is a contiguous row (full sectors): row q of QX8 gives the pivot and the factors f_i = -X[q,i];
// row p of B8 the pivot's support; rows i of B8 with f_i != 0 get row_i += f_i * row_p with the dot products Z_i . xs by
// dp4a; rows r of QX8 on the support get X[r,:] += xs_r * f plus the two column fixes; a deterministic measurement is the
// ordered product of the listed stabilizer rows.  Same closed forms as measure() in lanes.cuh
// (tableau_prime.py:262-363, exponentiate :365-380 folded in).  Needs n <= 512.
#pragma once

namespace lanesgm {

constexpr int kWarps = 4;            // shots in flight per CTA

struct GM8 {
  uint8_t* B;          // this warp's slab
  uint8_t* T;          // this shot's store
  uint8_t *xq, *f8, *ph8, *gf, *rv;   // shared memory, per warp
  uint16_t *gl, *rl;
  const uint8_t* inv;  // [d] inverses mod d
  uint8_t* stage;      // TMA staging buffer of this warp (kStageBytes), its mbarrier and the barrier's phase
  uint64_t* bar;
  uint32_t phase;
  int n, np, nq, W, nqw, Ww;          // nqw = nq / 4, Ww = W / 4
  int64_t row_bytes;
};

inline int nq_of(int n) { return (n + 31) / 32 * 32; }
inline size_t slab_bytes(int n, int W) { return (size_t)W * 2 * nq_of(n); }
inline size_t warp_smem_bytes(int n, int W) {
  const size_t nq = nq_of(n);
  return 4 * (size_t)W /* xq f8 ph8 gf */ + nq /* rv */ + 2 * (size_t)W /* gl */ + 2 * nq /* rl */;
}
inline size_t smem_bytes(int n, int W, bool tma) {
  return kWarps * (((warp_smem_bytes(n, W) + 15) & ~(size_t)15) + (tma ? 4096 + 16 : 0)) + 128;
}
inline bool shape_ok(int n, int d) { return d <= 127 && n <= 512; }

// w * s mod d on four packed lanes (s < d, every lane < d).  The slow path is one out-of-line function: inlined at its
// ~50 call sites it made the kernel 6 400 instructions (73 KB of hot code) and instruction fetch its top stall
// (ncu: stall_no_inst 43 %, gpurun_out/l8b); code size is a resource here as in round 1.
__device__ __noinline__ uint32_t smul4_slow(uint32_t d, uint32_t md, uint32_t w, uint32_t s) {
  const uint32_t a = (w & 0xFFu) * s, b = ((w >> 8) & 0xFFu) * s, c = ((w >> 16) & 0xFFu) * s, e = (w >> 24) * s;
  return (a - d * __umulhi(a, md)) | ((b - d * __umulhi(b, md)) << 8) | ((c - d * __umulhi(c, md)) << 16) |
         ((e - d * __umulhi(e, md)) << 24);
}
__device__ __forceinline__ uint32_t smul4(const Arith& A, uint32_t w, uint32_t s) {
  if (w == 0u || s == 0u) return 0u;
  if (s == 1u) return w;
  return smul4_slow(A.d, A.md, w, s);
}

// Ordered compaction of the non-zero bytes of a word vector held `words` per lane at word index lane + 32 t:
// list[k] = index of the byte, val[k] = the byte, ascending.  Returns the count (warp-uniform).
__device__ __forceinline__ int compact_bytes(const uint32_t* vec, int n_words, uint16_t* list, uint8_t* val, int lane) {
  int total = 0;
  for (int w0 = 0; w0 < n_words; w0 += 32) {
    const int w = w0 + lane;
    const uint32_t x = (w < n_words) ? vec[w] : 0u;
    const int cnt = (x & 0xFFu ? 1 : 0) + (x & 0xFF00u ? 1 : 0) + (x & 0xFF0000u ? 1 : 0) + (x >> 24 ? 1 : 0);
    if (!__any_sync(0xFFFFFFFFu, cnt)) continue;
    int incl = cnt;
#pragma unroll
    for (int d2 = 1; d2 < 32; d2 <<= 1) {
      const int o = __shfl_up_sync(0xFFFFFFFFu, incl, d2);
      if (lane >= d2) incl += o;
    }
    int pos = total + incl - cnt;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t b = (x >> (8 * k)) & 0xFFu;
      if (b) { list[pos] = (uint16_t)(4 * w + k); val[pos] = (uint8_t)b; ++pos; }
    }
    total += __shfl_sync(0xFFFFFFFFu, incl, 31);
  }
  __syncwarp();
  return total;
}

// store -> B8: tile = 32 qudits x one 4-lane word per lane; the lane then owns 32 consecutive bytes of four generator rows
__device__ __noinline__ void transpose_in(const GM8& M, int lane) {
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {                        // X block, then Z block
#pragma unroll 1
    for (int q0 = 0; q0 < M.nq; q0 += 32) {
#pragma unroll 1
      for (int w0 = 0; w0 < M.Ww; w0 += 32) {
        const int w = w0 + lane;
        if (w >= M.Ww) continue;
        uint32_t r[32];
#pragma unroll
        for (int k = 0; k < 32; ++k)
          r[k] = (q0 + k < M.n) ? reinterpret_cast<const uint32_t*>(M.T + (int64_t)(q0 + k) * M.row_bytes + (half ? M.W : 0))[w] : 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t o[8];
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            const uint32_t sel = 0x0000 | (uint32_t)j | ((uint32_t)(4 + j) << 4);       // byte j of a, byte j of b
            const uint32_t lo = __byte_perm(r[4 * m], r[4 * m + 1], sel);                // bytes: a_j, b_j, x, x
            const uint32_t hi = __byte_perm(r[4 * m + 2], r[4 * m + 3], sel);
            o[m] = __byte_perm(lo, hi, 0x5410);                                           // a_j b_j c_j d_j
          }
          uint4* dst = reinterpret_cast<uint4*>(M.B + (size_t)(4 * w + j) * 2 * M.nq + (half ? M.nq : 0) + q0);
          dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
          dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
    }
  }
}

// ---- TMA bulk loads of whole rows into a per-warp staging buffer ------------------------------------------------
// The rows of a list (generator rows of B8, qudit rows of QX8) are independent 512-byte reads, and one at a time they
// cost one dependent L2 / DRAM round trip each (ncu: stall_long_scoreboard 45 %).  With TMA the elected lane issues up to
// kStageRows `cp.async.bulk` copies at once — no registers, one instruction per row — into shared memory; an mbarrier
// with the expected byte count tells the warp when all of them have landed, and the rows are then processed from
// shared memory.  Writes stay ordinary coalesced stores.
// MEASURED (gpurun_out/l8f_probe.txt, d = 5, 4 096 shots): slower than one row at a time with an L2 prefetch of the next
// — n = 256 tail 6.25 vs 5.56 ms, n = 128 24.6 vs 19.7, n = 500 7.65 vs 6.95 (1 024 shots): the proxy fence + mbarrier
// round trip per batch and 7 instead of 8 resident CTAs cost more than the overlapped row latency buys, because the 32
// warps of an SM already overlap each other's round trips.  Kept as an opt-in form (SDIMB_TAIL8_TMA), bit-exact, tested.
constexpr int kStageBytes = 4096;     // per warp
constexpr int kStageRows = 8;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_row(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// rows list[k0 .. k0 + nb) of an array of rows (base, stride) -> M.stage; returns when they have landed
__device__ __forceinline__ void stage_rows(GM8& M, const uint16_t* list, int k0, int nb, const uint8_t* base, size_t stride,
                                           uint32_t rowbytes, int lane) {
  // earlier ordinary stores of this warp to those rows must be visible to the async proxy, and every lane must be done
  // with the previous contents of the staging buffer
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    mbar_expect_tx(M.bar, (uint32_t)nb * rowbytes);
    for (int r = 0; r < nb; ++r) tma_load_row(M.stage + (size_t)r * rowbytes, base + (size_t)list[k0 + r] * stride, rowbytes, M.bar);
  }
  mbar_wait(M.bar, M.phase);
  M.phase ^= 1u;
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) { return __reduce_add_sync(0xFFFFFFFFu, v); }

// Measurement of qudit q by one warp; returns the record byte.  NT = words per lane of one half row of B8 (1, 2, 4:
// n <= 128, 256, 512), a QX8 row has at most 2 NT words per lane.  Rows of a list are processed one at a time (the
// next one is prefetched into L2): with 32 warps per SM the other shots hide the round trip, and the code stays small.
template <int NT, bool TMA>
__device__ __forceinline__ uint32_t measure8(GM8& M, const Arith& A, const Swar& Sd, const int q, const uint32_t draw,
                                             const int lane) {
  constexpr int NQ = 2 * NT;
  const int nqw = M.nqw, Ww = M.Ww, np = M.np, nq = M.nq;
  const int KB = TMA ? min(kStageRows, kStageBytes / (2 * nq)) : 1;     // rows per staged batch (2 nq >= W)
  // ---- row q of QX8: staged, pivot = first stabilizer lane with an X component (tableau_prime.py:273-283) ----
  uint32_t piv = kNoPivot;
  {
    const uint32_t* rq = reinterpret_cast<const uint32_t*>(M.T + (int64_t)q * M.row_bytes);
#pragma unroll
    for (int t = 0; t < NQ; ++t) {
      const int w = lane + 32 * t;
      if (w < Ww) {
        const uint32_t x = rq[w];
        reinterpret_cast<uint32_t*>(M.xq)[w] = x;
        if (x && 4 * w < np && piv == kNoPivot) piv = 4u * w + ((__ffs(x) - 1) >> 3);
      }
    }
    piv = __reduce_min_sync(0xFFFFFFFFu, piv);
    __syncwarp();
  }
  uint32_t rec;
  if (piv != kNoPivot) {
    // ---- random branch (tableau_prime.py:294-334) ----
    const uint32_t e = M.inv[M.xq[piv]];
    uint32_t* const Bp = reinterpret_cast<uint32_t*>(M.B + (size_t)piv * 2 * nq);
    uint32_t* const Bd = reinterpret_cast<uint32_t*>(M.B + (size_t)(np + piv) * 2 * nq);
    uint32_t xs[NT], zs[NT], odx[NT];
    uint32_t sd_raw = 0;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int w = lane + 32 * t;
      xs[t] = zs[t] = odx[t] = 0u;
      if (w < nqw) {
        const uint32_t x = Bp[w], z = Bp[nqw + w];
        odx[t] = Bd[w];                                          // X support of the destabilizer that is overwritten below
        sd_raw = __dp4a(x, z, sd_raw);
        xs[t] = smul4(A, x, e);                                  // pivot^e (exponentiate :365-380)
        zs[t] = smul4(A, z, e);
      }
    }
    sd_raw = mod_d(A, warp_sum(sd_raw));
    const uint32_t ps_old = M.ph8[piv];
    const uint32_t ps = mod_o(A, ps_old * e + A.po * mod_d(A, sd_raw * mod_d(A, (e * (e - 1u)) >> 1)));
    const uint32_t sd = mod_d(A, mod_d(A, sd_raw * e) * e);     // x_p . z_p after exponentiation
    // factors f_i = -X[q,i] of every lane but the pivot and its destabilizer (both are replaced below)
    const int wp = (int)(piv >> 2), wd = (int)((np + piv) >> 2);
    const uint32_t mp = ~(0xFFu << (8 * (piv & 3))), md = ~(0xFFu << (8 * ((np + piv) & 3)));
#pragma unroll
    for (int t = 0; t < NQ; ++t) {
      const int w = lane + 32 * t;
      if (w < Ww) {
        uint32_t fw = swar_neg(Sd, reinterpret_cast<const uint32_t*>(M.xq)[w]);
        if (w == wp) fw &= mp;
        if (w == wd) fw &= md;
        reinterpret_cast<uint32_t*>(M.f8)[w] = fw;
      }
    }
    __syncwarp();
    const int total = compact_bytes(reinterpret_cast<const uint32_t*>(M.f8), Ww, M.gl, M.gf, lane);
    // row_i += f_i * pivot for every listed generator      (tableau_prime.py:306-321)
#pragma unroll 1
    for (int k0 = 0; k0 < total; k0 += KB) {
     const int nb = min(KB, total - k0);
     if (TMA) stage_rows(M, M.gl, k0, nb, M.B, (size_t)2 * nq, 2u * nq, lane);
#pragma unroll 1
     for (int k = k0; k < k0 + nb; ++k) {
      const int gi = M.gl[k];
      const uint32_t f = M.gf[k];
      uint32_t* const row = reinterpret_cast<uint32_t*>(M.B + (size_t)gi * 2 * nq);
      const uint32_t* const src = TMA ? reinterpret_cast<const uint32_t*>(M.stage + (size_t)(k - k0) * 2 * nq) : row;
      if (!TMA && k + 1 < total) {
        const uint32_t* nxt = reinterpret_cast<const uint32_t*>(M.B + (size_t)M.gl[k + 1] * 2 * nq);
        if (lane < 2 * NT) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + 32 * lane));
      }
      uint32_t xa[NT], za[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int w = lane + 32 * t;
        xa[t] = za[t] = 0u;
        if (w < nqw) { xa[t] = src[w]; za[t] = src[nqw + w]; }
      }
      uint32_t dot = 0;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int w = lane + 32 * t;
        dot = __dp4a(za[t], xs[t], dot);                       // Z[:,i] . xs (old Z)
        if (xs[t]) row[w] = swar_add(Sd, xa[t], smul4(A, xs[t], f));
        if (zs[t]) row[nqw + w] = swar_add(Sd, za[t], smul4(A, zs[t], f));
      }
      dot = warp_sum(dot);
      if (lane == 0) {
        // P_i += f*ps + po*((Z_i.xs)*f + sd*f(f-1)/2*po)      (tableau_prime.py:310-312,317-319)
        const uint32_t g = mod_d(A, (f * (f - 1u)) >> 1);
        const uint32_t cp = mod_d(A, mod_d(A, dot) * f + sd * g * A.po);
        M.ph8[gi] = (uint8_t)mod_o(A, (uint32_t)M.ph8[gi] + f * ps + A.po * cp);
      }
     }
    }
    // QX8: X[r,:] += xs_r * f on the pivot's X support; column p <- 0, column np + p <- xs (also where only the old
    // destabilizer had an entry).  Marker byte of a row = xs_r, or 0x80 where only the old destabilizer is non-zero.
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int w = lane + 32 * t;
      const uint32_t x = xs[t], o = odx[t];
      uint32_t m = x;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (!((x >> (8 * k)) & 0xFFu) && ((o >> (8 * k)) & 0xFFu)) m |= 0x80u << (8 * k);
      if (w < nqw) reinterpret_cast<uint32_t*>(M.rv)[w] = m;
    }
    __syncwarp();
    const int totq = compact_bytes(reinterpret_cast<const uint32_t*>(M.rv), nqw, M.rl, M.gf, lane);
#pragma unroll 1
    for (int k0 = 0; k0 < totq; k0 += KB) {
     const int nb = min(KB, totq - k0);
     if (TMA) stage_rows(M, M.rl, k0, nb, M.T, (size_t)M.row_bytes, (uint32_t)M.W, lane);
#pragma unroll 1
     for (int k = k0; k < k0 + nb; ++k) {
      const int r = M.rl[k];
      const uint32_t sv = M.gf[k] & 0x7Fu;
      uint32_t* const row = reinterpret_cast<uint32_t*>(M.T + (int64_t)r * M.row_bytes);
      const uint32_t* const src = TMA ? reinterpret_cast<const uint32_t*>(M.stage + (size_t)(k - k0) * M.W) : row;
      if (!TMA && k + 1 < totq) {
        const uint32_t* nxt = reinterpret_cast<const uint32_t*>(M.T + (int64_t)M.rl[k + 1] * M.row_bytes);
        if (lane < NQ) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + 32 * lane));
      }
      uint32_t xr[NQ];
#pragma unroll
      for (int t = 0; t < NQ; ++t) {
        const int w = lane + 32 * t;
        xr[t] = (w < Ww) ? src[w] : 0u;
      }
#pragma unroll
      for (int t = 0; t < NQ; ++t) {
        const int w = lane + 32 * t;
        if (w < Ww) {
          const uint32_t fw = reinterpret_cast<const uint32_t*>(M.f8)[w];
          uint32_t nx = (fw && sv) ? swar_add(Sd, xr[t], smul4(A, fw, sv)) : xr[t];
          if (w == wp) nx &= mp;
          if (w == wd) nx = (nx & md) | (sv << (8 * ((np + piv) & 3)));
          if (nx != xr[t]) row[w] = nx;
        }
      }
     }
    }
    // destabilizer p <- (xs, zs, ps); stabilizer p <- Z_q with phase -m*po   (tableau_prime.py:323-333)
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int w = lane + 32 * t;
      if (w < nqw) {
        Bd[w] = xs[t]; Bd[nqw + w] = zs[t];
        Bp[w] = 0u; Bp[nqw + w] = (w == (q >> 2)) ? (1u << (8 * (q & 3))) : 0u;
      }
    }
    if (lane == 0) {
      M.ph8[np + piv] = (uint8_t)ps;
      M.ph8[piv] = (uint8_t)mod_o(A, A.order - draw * A.po);
    }
    rec = draw;
  } else {
    // ---- deterministic branch (tableau_prime.py:336-363): ordered product of the stabilizers i with f_i = destab X[q,i] ----
    const int total = compact_bytes(reinterpret_cast<const uint32_t*>(M.xq) + np / 4, np / 4, M.gl, M.gf, lane);
    uint32_t az[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) az[t] = 0u;
    uint32_t cross = 0, sdg = 0, a1 = 0;
#pragma unroll 1
    for (int k0 = 0; k0 < total; k0 += KB) {
     const int nb = min(KB, total - k0);
     if (TMA) stage_rows(M, M.gl, k0, nb, M.B, (size_t)2 * nq, 2u * nq, lane);
#pragma unroll 1
     for (int k = k0; k < k0 + nb; ++k) {
      const int gi = M.gl[k];
      const uint32_t f = M.gf[k], g = mod_d(A, (f * (f - 1u)) >> 1);
      const uint32_t* const row = TMA ? reinterpret_cast<const uint32_t*>(M.stage + (size_t)(k - k0) * 2 * nq)
                                      : reinterpret_cast<const uint32_t*>(M.B + (size_t)gi * 2 * nq);
      if (!TMA && k + 1 < total) {
        const uint32_t* nxt = reinterpret_cast<const uint32_t*>(M.B + (size_t)M.gl[k + 1] * 2 * nq);
        if (lane < 2 * NT) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + 32 * lane));
      }
      uint32_t xa[NT], za[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int w = lane + 32 * t;
        xa[t] = za[t] = 0u;
        if (w < nqw) { xa[t] = row[w]; za[t] = row[nqw + w]; }
      }
      uint32_t xz = 0;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        if (xa[t]) cross = __dp4a(smul4(A, xa[t], f), az[t], cross);            // ancilla_z . (f * x_i), running ancilla
        if (za[t]) az[t] = swar_add(Sd, az[t], smul4(A, za[t], f));
        xz = __dp4a(xa[t], za[t], xz);
      }
      cross = mod_d(A, cross);
      sdg = mod_d(A, sdg + mod_d(A, xz) * g);
      a1 += f * (uint32_t)M.ph8[gi];                                              // every lane, same value
     }
    }
    const uint32_t part = mod_d(A, warp_sum(mod_d(A, cross + A.po * sdg)));
    const uint32_t ap = mod_o(A, mod_o(A, a1) + A.po * part);
    const uint32_t outcome = (A.po == 1) ? neg_d(A, ap) : (((ap + 1u) >> 1) & 1u);   // (-ap // po) % d  (:362)
    rec = outcome | SDIMB_REC_DET;
  }
  __syncwarp();
  return rec;
}

template <int NT, bool TMA>
__global__ void __launch_bounds__(32 * kWarps, 8) run_tail8_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const Arith A = p.A;
  const Swar Sd = make_swar(A.d);
  GM8 M;
  M.n = p.n; M.np = p.np; M.W = p.W; M.nq = (p.n + 31) / 32 * 32; M.nqw = M.nq / 4; M.Ww = p.W / 4;
  M.row_bytes = p.row_bytes;
  uint8_t* inv = smem;                                            // [128]
  const size_t per_warp = ((size_t)(4 * p.W + M.nq + 2 * p.W + 2 * M.nq) + 15) & ~(size_t)15;
  uint8_t* base = smem + 128 + warp * (per_warp + (TMA ? kStageBytes + 16 : 0));
  M.xq = base; M.f8 = M.xq + p.W; M.ph8 = M.f8 + p.W; M.gf = M.ph8 + p.W; M.rv = M.gf + p.W;
  M.gl = reinterpret_cast<uint16_t*>(M.rv + M.nq); M.rl = M.gl + p.W;
  M.inv = inv;
  M.stage = base + per_warp;                                      // 16-byte aligned: per_warp is a multiple of 16
  M.bar = reinterpret_cast<uint64_t*>(M.stage + kStageBytes);
  M.phase = 0u;
  if (TMA && lane == 0) mbar_init(M.bar, 1);
  if (TMA) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  for (uint32_t v = threadIdx.x; v < A.d; v += blockDim.x) {      // inverses mod d by search (d <= 127)
    uint32_t r = 0;
    for (uint32_t c = 1; c < A.d; ++c)
      if (mod_d(A, c * v) == 1u) r = c;
    inv[v] = (uint8_t)r;
  }
  __syncthreads();
  M.B = reinterpret_cast<uint8_t*>(p.gm_slab) + ((int64_t)blockIdx.x * kWarps + warp) * p.gm_slab_words * 4;
  for (;;) {
    int64_t shot = 0;
    if (lane == 0) shot = (int64_t)atomicAdd(p.shot_counter, 1u);
    shot = __shfl_sync(0xFFFFFFFFu, shot, 0);
    if (shot >= p.shots) break;
    M.T = p.tab + shot * p.shot_bytes;
    for (int w = lane; w < M.Ww; w += 32)
      reinterpret_cast<uint32_t*>(M.ph8)[w] = reinterpret_cast<const uint32_t*>(M.T + p.phase_off)[w];
    transpose_in(M, lane);
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
          mine.z = (int)__umulhi(r.x, A.d);
        }
      }
      uint32_t todo = __ballot_sync(0xFFFFFFFFu, is_m);
      uint32_t myrec = 0;
      while (todo) {
        const int k = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t rec = measure8<NT, TMA>(M, A, Sd, __shfl_sync(0xFFFFFFFFu, mine.y, k), (uint32_t)__shfl_sync(0xFFFFFFFFu, mine.z, k), lane);
        if (lane == k) myrec = rec;
      }
      if (is_m) p.records[shot * p.rec_stride + mine.w] = (uint8_t)myrec;
    }
  }
}

}  // namespace lanesgm
