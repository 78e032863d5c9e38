// uint16-lane interpreter: odd primes 127 < d < 2^15 (north_star (1): "packed uint8/uint16 lanes for odd prime d").
// Part of libsdimb (sdim_b200/csrc); included by sdimb.cu inside its anonymous namespace.
//
// Same store as the uint8 lanes (include/sdimb.h) with two bytes per entry; the order of an odd prime is d itself
// (phase_order 1, sdim/tableau/dataclasses.py:88-106).  One CTA owns one shot on the HBM store.  Values above 127 do
// not pack four to a word with a spare carry bit, so the arithmetic is per lane in 32 bits: a product of two reduced
// values stays below 2^30 and is reduced with one multiply-high, one multiply-subtract and one conditional subtract.
//   gates        a thread owns lanes tid, tid + blockDim, ... for the whole gate sequence (generator lanes are
//                independent under gates: no barrier between gates, row accesses coalesced)
//   measurement  tableau_prime.py:262-363 as in lanes.cuh: block-min pivot search, one walk down the pivot column
//                (exponentiate :365-380 folded in) -> support list, rank-1 update with lane-owning threads looping over
//                the support rows, phase update, column writes; the deterministic branch compacts the non-zero
//                factors in order and accumulates per qudit row with the running ancilla.
// Records, replayed outcomes and replayed noise exponents are uint16 (bit 15 of a record = deterministic flag).
#pragma once

namespace wide {

constexpr int kThreads = 256;
constexpr uint32_t kRecDet = 0x8000u;

struct Mod {             // x mod d for x < 2^32: q = mulhi(x, floor(2^32 / d)) is the quotient or one less
  uint32_t d, m;
  __device__ __forceinline__ uint32_t operator()(uint32_t x) const {
    uint32_t r = x - d * __umulhi(x, m);
    if (r >= d) r -= d;
    return r;
  }
  __device__ __forceinline__ uint32_t neg(uint32_t x) const { return x ? d - x : 0u; }
  __device__ __forceinline__ uint32_t add(uint32_t a, uint32_t b) const { const uint32_t s = a + b; return s >= d ? s - d : s; }
  __device__ __forceinline__ uint32_t sub(uint32_t a, uint32_t b) const { return a >= b ? a - b : a + d - b; }
  __device__ __forceinline__ uint32_t mul(uint32_t a, uint32_t b) const { return (*this)(a * b); }
  __device__ __forceinline__ uint32_t tri(uint32_t f) const { return (*this)((f * (f - 1u)) >> 1); }   // f(f-1)/2 mod d
};

inline size_t smem_bytes(int np) {   // xs, zs, list, lf (uint16 [np] each) + red[32] + cnt[8]
  return 4 * 2 * (size_t)np + 32 * 4 + 8 * 4 + 16;
}

struct Scr {
  uint16_t *xs, *zs, *list, *lf;
  uint32_t *red, *cnt;
};

__device__ __forceinline__ uint32_t inverse_mod(uint32_t v, uint32_t d) {   // extended Euclid, d prime, 0 < v < d
  int64_t t = 0, nt = 1, r = d, nr = v;
  while (nr) {
    const int64_t q = r / nr;
    int64_t tmp = t - q * nt; t = nt; nt = tmp;
    tmp = r - q * nr; r = nr; nr = tmp;
  }
  return (uint32_t)(t < 0 ? t + d : t);
}

// conjugation by X^a Z^b on qudit q: phase += b*x - a*z   (tableau_gates.py:27-137, program.py:335-339)
__device__ __forceinline__ void pauli(uint16_t* T, const KParams& p, const Mod& M, int q, uint32_t a, uint32_t b) {
  const int W = p.W;
  uint16_t* row = T + (int64_t)q * 2 * W;
  uint16_t* P = T + (int64_t)p.n * 2 * W;
  const uint32_t na = M.neg(a);
  for (int l = threadIdx.x; l < W; l += blockDim.x) {
    const uint32_t x = b ? row[l] : 0u, z = na ? row[W + l] : 0u;
    if ((x | z) == 0) continue;
    P[l] = (uint16_t)M.add(P[l], M.add(M.mul(b, x), M.mul(na, z)));
  }
}

__device__ void gate(uint16_t* T, const KParams& p, const Mod& M, int op, int a, int b) {
  const int W = p.W;
  uint16_t* ra = T + (int64_t)a * 2 * W;
  uint16_t* rb = T + (int64_t)b * 2 * W;
  uint16_t* P = T + (int64_t)p.n * 2 * W;
  for (int l = threadIdx.x; l < W; l += blockDim.x) {
    switch (op) {
      case SDIMB_OP_H: case SDIMB_OP_H_INV: {            // tableau_optimized.py:5-58
        const uint32_t x = ra[l], z = ra[W + l];
        if ((x | z) == 0) break;
        P[l] = (uint16_t)M.sub(P[l], M.mul(x, z));
        if (op == SDIMB_OP_H) { ra[l] = (uint16_t)M.neg(z); ra[W + l] = (uint16_t)x; }
        else { ra[l] = (uint16_t)z; ra[W + l] = (uint16_t)M.neg(x); }
        break;
      }
      case SDIMB_OP_P: case SDIMB_OP_P_INV: {            // tableau_optimized.py:62-96 (odd d: x(x-1)/2)
        const uint32_t x = ra[l];
        if (!x) break;
        const uint32_t inc = M.tri(x), z = ra[W + l];
        if (op == SDIMB_OP_P) { P[l] = (uint16_t)M.add(P[l], inc); ra[W + l] = (uint16_t)M.add(z, x); }
        else { P[l] = (uint16_t)M.sub(P[l], inc); ra[W + l] = (uint16_t)M.sub(z, x); }
        break;
      }
      case SDIMB_OP_CNOT: case SDIMB_OP_CNOT_INV: {      // tableau_optimized.py:99-118
        const uint32_t xc = ra[l], zt = rb[W + l];
        if ((xc | zt) == 0) break;
        if (op == SDIMB_OP_CNOT) { rb[l] = (uint16_t)M.add(rb[l], xc); ra[W + l] = (uint16_t)M.sub(ra[W + l], zt); }
        else { rb[l] = (uint16_t)M.sub(rb[l], xc); ra[W + l] = (uint16_t)M.add(ra[W + l], zt); }
        break;
      }
      case SDIMB_OP_CZ: case SDIMB_OP_CZ_INV: {          // tableau_gates.py:229-261, folded
        const uint32_t xa = ra[l], xb = rb[l];
        if ((xa | xb) == 0) break;
        const uint32_t prod = M.mul(xa, xb);
        if (op == SDIMB_OP_CZ) {
          P[l] = (uint16_t)M.add(P[l], prod);
          ra[W + l] = (uint16_t)M.add(ra[W + l], xb); rb[W + l] = (uint16_t)M.add(rb[W + l], xa);
        } else {
          P[l] = (uint16_t)M.sub(P[l], prod);
          ra[W + l] = (uint16_t)M.sub(ra[W + l], xb); rb[W + l] = (uint16_t)M.sub(rb[W + l], xa);
        }
        break;
      }
      case SDIMB_OP_SWAP: {                              // tableau_gates.py:298-329 (prime branch)
        const uint16_t x = ra[l], z = ra[W + l];
        ra[l] = rb[l]; ra[W + l] = rb[W + l]; rb[l] = x; rb[W + l] = z;
        break;
      }
      default: break;
    }
  }
}

// Measurement of qudit q (tableau_prime.py:262-363).  Called by the whole CTA behind a barrier; ends with one.
__device__ uint32_t measure(uint16_t* T, const KParams& p, const Mod& M, Scr& S, int q, uint32_t draw) {
  const int n = p.n, np = p.np, W = p.W, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t d = M.d;
  uint16_t* rowq = T + (int64_t)q * 2 * W;
  uint16_t* P = T + (int64_t)n * 2 * W;
  // pivot: first stabilizer with an X component on q (:273-283)
  uint32_t best = kNoPivot;
  for (int l = tid; l < n; l += nt)
    if (rowq[l]) { best = (uint32_t)l; break; }
  if (tid < 8) S.cnt[tid] = 0;
  const uint32_t piv = block_min(best, S.red);          // has barriers: the counters are cleared for everyone
  uint32_t rec;
  if (piv != kNoPivot) {
    // ---- random branch (:294-334), exponentiate (:365-380) folded in ----
    const uint32_t e = inverse_mod(rowq[piv], d);
    uint32_t raw = 0;
    for (int r = tid; r < n; r += nt) {
      const uint16_t* row = T + (int64_t)r * 2 * W;
      const uint32_t xr = row[piv], zr = row[W + piv];
      raw = M.add(raw, M.mul(xr, zr));
      const uint32_t xs = M.mul(xr, e), zs = M.mul(zr, e);
      S.xs[r] = (uint16_t)xs; S.zs[r] = (uint16_t)zs;
      if (xs | zs) S.list[atomicAdd(&S.cnt[0], 1u)] = (uint16_t)r;
    }
    raw = M(block_sum(raw, S.red));                      // barriers: xs / zs / list are complete
    const int nr = (int)S.cnt[0];
    const uint32_t ps = M.add(M.mul(P[piv], e), M.mul(raw, M.tri(e)));
    const uint32_t sd = M.mul(M.mul(raw, e), e);
    __syncthreads();                                      // P[piv] read before anyone rewrites it
    // row_i += f_i * pivot for every generator lane: the lane's owner walks the support rows
    for (int l = tid; l < W; l += nt) {
      uint32_t f = M.neg(rowq[l]);
      if ((uint32_t)l == piv) f = 0;                      // the pivot itself (stabilizer block only, :314-315)
      if (!f) continue;
      uint32_t dot = 0;
      for (int k = 0; k < nr; ++k) {
        const int r = S.list[k];
        uint16_t* row = T + (int64_t)r * 2 * W;
        const uint32_t s = S.xs[r], u = S.zs[r], x = row[l], z = row[W + l];
        dot = M.add(dot, M.mul(z, s));                    // Z[:,i] . x_p (old Z)
        row[l] = (uint16_t)M.add(x, M.mul(f, s));
        row[W + l] = (uint16_t)M.add(z, M.mul(f, u));
      }
      // P_i += f*ps + f*dot + sd*f(f-1)/2   (:310-312,317-319)
      P[l] = (uint16_t)M.add(P[l], M.add(M.mul(f, ps), M.add(M.mul(f, dot), M.mul(sd, M.tri(f)))));
    }
    __syncthreads();
    // destabilizer p <- old pivot; stabilizer p <- Z_q with phase -m   (:323-333)
    for (int r = tid; r < n; r += nt) {
      uint16_t* row = T + (int64_t)r * 2 * W;
      row[np + piv] = S.xs[r];
      row[W + np + piv] = S.zs[r];
      row[piv] = 0;
      row[W + piv] = (r == q) ? 1 : 0;
    }
    if (tid == 0) { P[np + piv] = (uint16_t)ps; P[piv] = (uint16_t)M.neg(draw); }
    rec = draw;
  } else {
    // ---- deterministic branch (:336-363): stabilizers i with f_i = destab X[q,i] != 0, in increasing i ----
    uint32_t a1 = 0;
    int total = 0;
    for (int base = 0; base < n; base += nt) {           // ordered compaction, one block of nt generators at a time
      const int i = base + tid;
      const uint32_t f = (i < n) ? rowq[np + i] : 0u;
      const uint32_t m = __ballot_sync(0xFFFFFFFFu, f != 0);
      if (lane == 0) S.red[warp] = (uint32_t)__popc(m);
      __syncthreads();
      int off = total, all = 0;
      for (int w = 0; w < (nt >> 5); ++w) { const int c = (int)S.red[w]; if (w < warp) off += c; all += c; }
      if (f) {
        const int pos = off + __popc(m & ((1u << lane) - 1u));
        S.list[pos] = (uint16_t)i;
        S.lf[pos] = (uint16_t)f;
        a1 = M.add(a1, M.mul(f, P[i]));
      }
      total += all;
      __syncthreads();
    }
    uint32_t part = 0;
    for (int r = tid; r < n; r += nt) {
      const uint16_t* row = T + (int64_t)r * 2 * W;
      uint32_t az = 0, cross = 0, sdg = 0;
      for (int k = 0; k < total; ++k) {
        const int i = S.list[k];
        const uint32_t f = S.lf[k], xi = row[i], zi = row[W + i];
        if ((xi | zi) == 0) continue;
        cross = M.add(cross, M.mul(az, M.mul(f, xi)));    // ancilla_z . (f * x_i), running ancilla
        az = M.add(az, M.mul(f, zi));
        sdg = M.add(sdg, M.mul(M.mul(xi, zi), M.tri(f)));
      }
      part = M.add(part, M.add(cross, sdg));
    }
    const uint32_t ap = M.add(M(block_sum(a1, S.red)), M(block_sum(part, S.red)));
    rec = M.neg(ap) | kRecDet;                            // (-ap // 1) % d   (:362)
  }
  __syncthreads();
  return rec;
}

__global__ void __launch_bounds__(kThreads) interp_wide16_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  Scr S;
  S.xs = reinterpret_cast<uint16_t*>(smem);
  S.zs = S.xs + p.np;
  S.list = S.zs + p.np;
  S.lf = S.list + p.np;
  S.red = reinterpret_cast<uint32_t*>(S.lf + p.np);
  S.cnt = S.red + 32;
  const Mod M{p.A.d, (uint32_t)((1ull << 32) / p.A.d)};
  const uint32_t d = p.A.d;
  const int W = p.W, tid = threadIdx.x, nt = blockDim.x;
  uint16_t* const records = reinterpret_cast<uint16_t*>(p.records);
  const uint16_t* const replay_meas = reinterpret_cast<const uint16_t*>(p.replay_meas);
  const uint16_t* const replay_noise = reinterpret_cast<const uint16_t*>(p.replay_noise);
  for (int64_t shot = blockIdx.x; shot < p.shots; shot += gridDim.x) {
    uint16_t* T = reinterpret_cast<uint16_t*>(p.tab + shot * p.shot_bytes);
    if (p.flags & SDIMB_FRESH) {
      uint4* v = reinterpret_cast<uint4*>(T);
      for (int64_t i = tid; i < p.shot_bytes / 16; i += nt) v[i] = make_uint4(0, 0, 0, 0);
      __syncthreads();
      for (int q = tid; q < p.n; q += nt) {
        T[(int64_t)q * 2 * W + W + q] = 1;        // stabilizer q = Z_q
        T[(int64_t)q * 2 * W + p.np + q] = 1;     // destabilizer q = X_q
      }
      __syncthreads();
    }
    const uint64_t gshot = (uint64_t)(p.shot_offset + shot);
    for (int64_t i = 0; i < p.n_ops; ++i) {
      const int4 op = __ldg(p.ops + i);
      const int code = op.x & SDIMB_OP_MASK;
      switch (code) {
        case SDIMB_OP_I: case SDIMB_OP_BARRIER: break;
        case SDIMB_OP_X: pauli(T, p, M, op.y, 1u, 0u); break;
        case SDIMB_OP_X_INV: pauli(T, p, M, op.y, d - 1u, 0u); break;
        case SDIMB_OP_Z: pauli(T, p, M, op.y, 0u, 1u); break;
        case SDIMB_OP_Z_INV: pauli(T, p, M, op.y, 0u, d - 1u); break;
        case SDIMB_OP_N1: {                       // program.py:486-507; every thread evaluates the same event
          uint32_t a = 0, b = 0;
          if (replay_noise) {
            a = replay_noise[(shot * p.n_noise + op.w) * 2]; b = replay_noise[(shot * p.n_noise + op.w) * 2 + 1];
          } else {
            const uint4 r = philox4x32((uint32_t)gshot, (uint32_t)(gshot >> 32), (uint32_t)op.w, 1u, (uint32_t)p.seed,
                                       (uint32_t)(p.seed >> 32));
            if ((r.x >> 8) >= __ldg(p.thresh + op.w)) {
              const uint32_t ch = __ldg(p.chan + op.w);
              if (ch == 0) { const uint32_t v = 1u + __umulhi(r.y, d * d - 1u); a = v % d; b = v / d; }
              else { const uint32_t ev = 1u + __umulhi(r.y, d - 1u); if (ch == 1) a = ev; else b = ev; }
            }
          }
          if (a | b) pauli(T, p, M, op.y, a, b);
          break;
        }
        case SDIMB_OP_M_X:
          gate(T, p, M, SDIMB_OP_H_INV, op.y, 0);   // tableau_gates.py:292-296
          /* fallthrough */
        case SDIMB_OP_M: case SDIMB_OP_RESET: {
          uint32_t draw;
          if (replay_meas) {
            draw = replay_meas[shot * p.n_meas + op.w];
          } else {
            const uint4 r = philox4x32((uint32_t)gshot, (uint32_t)(gshot >> 32), (uint32_t)op.w, 0u, (uint32_t)p.seed,
                                       (uint32_t)(p.seed >> 32));
            draw = __umulhi(r.x, d);
          }
          __syncthreads();                        // the gates in front of the measurement are lane-owned
          const uint32_t rec = measure(T, p, M, S, op.y, draw);
          if (tid == 0) records[shot * p.rec_stride + op.w] = (uint16_t)rec;
          const uint32_t m = rec & (kRecDet - 1u);
          if (code == SDIMB_OP_RESET && m) pauli(T, p, M, op.y, d - m, 0u);   // program.py:335-339
          break;
        }
        default: gate(T, p, M, code, op.y, op.z); break;
      }
    }
    __syncthreads();
  }
}

__global__ void init16_kernel(uint16_t* tab, int n, int np, int W, int64_t shot_elems, int64_t shots) {
  for (int64_t shot = blockIdx.x; shot < shots; shot += gridDim.x) {
    uint16_t* T = tab + shot * shot_elems;
    uint4* v = reinterpret_cast<uint4*>(T);
    for (int64_t i = threadIdx.x; i < shot_elems / 8; i += blockDim.x) v[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
      T[(int64_t)q * 2 * W + W + q] = 1;
      T[(int64_t)q * 2 * W + np + q] = 1;
    }
    __syncthreads();
  }
}

__global__ void export16_kernel(const uint16_t* T, int n, int np, int W, int64_t* x, int64_t* z, int64_t* ph,
                                int64_t* dx, int64_t* dz, int64_t* dph) {
  const int64_t total = (int64_t)n * n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i / n), g = (int)(i % n);
    const uint16_t* row = T + (int64_t)q * 2 * W;
    x[i] = row[g];
    z[i] = row[W + g];
    dx[i] = row[np + g];
    dz[i] = row[W + np + g];
    if (q == 0) {
      ph[g] = T[(int64_t)n * 2 * W + g];
      dph[g] = T[(int64_t)n * 2 * W + np + g];
    }
  }
}

}  // namespace wide
