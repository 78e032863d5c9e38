// Pauli-frame sampler kernel (simulate_frame, sdim/program.py:45-165).
// Part of libsdimb (sdim_b200/csrc); included by sdimb.cu inside its anonymous namespace.
#pragma once

// ---------------------------------------------------------------------------------------------
// Pauli-frame sampler (simulate_frame, sdim/program.py:45-165): one thread per extra shot, frames laid out
// [x|z][qudit][shot] so that every access of a warp is one contiguous 32-byte run.
// ---------------------------------------------------------------------------------------------
struct FParams {
  const int4* ops;
  int64_t n_ops;
  const uint8_t* reference;
  uint8_t* records;
  int64_t n_meas, rec_stride;
  uint8_t* frames;
  int64_t pitch;
  const uint8_t *replay_z0, *replay_zm, *replay_noise;
  const uint32_t* thresh;
  const uint8_t* chan;
  int64_t n_noise;
  uint64_t seed;
  int64_t shots, shot_offset;
  int n;
  Arith A;
};

__global__ void __launch_bounds__(128) frame_kernel(const __grid_constant__ FParams p) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= p.shots) return;
  const Arith& A = p.A;
  const uint32_t d = A.d;
  uint8_t* X = p.frames + s;
  uint8_t* Z = p.frames + (int64_t)p.n * p.pitch + s;
  const uint64_t gshot = (uint64_t)(p.shot_offset + s);
  const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
  for (int q = 0; q < p.n; ++q) {                       // x_frame = 0, z_frame uniform   (program.py:63-64)
    X[(int64_t)q * p.pitch] = 0;
    Z[(int64_t)q * p.pitch] = p.replay_z0 ? p.replay_z0[s * p.n + q]
        : (uint8_t)__umulhi(philox4x32((uint32_t)gshot, (uint32_t)(gshot >> 32), (uint32_t)q, 2u, k0, k1).x, d);
  }
  for (int64_t i = 0; i < p.n_ops; ++i) {
    const int4 op = __ldg(p.ops + i);
    const int code = op.x & SDIMB_OP_MASK;
    uint8_t* xa = X + (int64_t)op.y * p.pitch;
    uint8_t* za = Z + (int64_t)op.y * p.pitch;
    uint8_t* xb = X + (int64_t)op.z * p.pitch;
    uint8_t* zb = Z + (int64_t)op.z * p.pitch;
    switch (code) {
      case SDIMB_OP_H: { const uint32_t x = *xa, z = *za; *xa = (uint8_t)neg_d(A, z); *za = (uint8_t)x; break; }
      case SDIMB_OP_H_INV: { const uint32_t x = *xa, z = *za; *xa = (uint8_t)z; *za = (uint8_t)neg_d(A, x); break; }
      case SDIMB_OP_P: *za = (uint8_t)mod_d(A, (uint32_t)*za + *xa); break;
      case SDIMB_OP_P_INV: *za = (uint8_t)mod_d(A, (uint32_t)*za + d - *xa); break;
      case SDIMB_OP_CNOT: *xb = (uint8_t)mod_d(A, (uint32_t)*xb + *xa); *za = (uint8_t)mod_d(A, (uint32_t)*za + d - *zb); break;
      case SDIMB_OP_CNOT_INV: *xb = (uint8_t)mod_d(A, (uint32_t)*xb + d - *xa); *za = (uint8_t)mod_d(A, (uint32_t)*za + *zb); break;
      case SDIMB_OP_CZ: *zb = (uint8_t)mod_d(A, (uint32_t)*zb + *xa); *za = (uint8_t)mod_d(A, (uint32_t)*za + *xb); break;
      case SDIMB_OP_CZ_INV: *zb = (uint8_t)mod_d(A, (uint32_t)*zb + d - *xa); *za = (uint8_t)mod_d(A, (uint32_t)*za + d - *xb); break;
      case SDIMB_OP_SWAP: { const uint8_t x = *xa, z = *za; *xa = *xb; *za = *zb; *xb = x; *zb = z; break; }
      case SDIMB_OP_M:
      case SDIMB_OP_M_X:
      case SDIMB_OP_RESET: {
        if (code == SDIMB_OP_M_X) { const uint32_t x = *xa, z = *za; *xa = (uint8_t)z; *za = (uint8_t)neg_d(A, x); }
        const uint32_t ref = p.reference[op.w];
        const uint32_t val = mod_d(A, (ref & SDIMB_REC_VALUE) + *xa);       // program.py:135
        p.records[s * p.rec_stride + op.w] = (uint8_t)(val | (ref & SDIMB_REC_DET));
        if (code == SDIMB_OP_RESET) *xa = 0;                                 // program.py:155
        *za = p.replay_zm ? p.replay_zm[s * p.n_meas + op.w]                 // program.py:144,156
            : (uint8_t)__umulhi(philox4x32((uint32_t)gshot, (uint32_t)(gshot >> 32), (uint32_t)op.w, 3u, k0, k1).x, d);
        break;
      }
      case SDIMB_OP_N1: {
        uint32_t a = 0, b = 0;
        const int64_t j = op.w;
        if (p.replay_noise) {
          a = p.replay_noise[(s * p.n_noise + j) * 2]; b = p.replay_noise[(s * p.n_noise + j) * 2 + 1];
        } else {
          const uint4 r = philox4x32((uint32_t)gshot, (uint32_t)(gshot >> 32), (uint32_t)j, 1u, k0, k1);
          if ((r.x >> 8) >= __ldg(p.thresh + j)) {
            const uint32_t ch = __ldg(p.chan + j);
            if (ch == 0) { const uint32_t v = 1u + __umulhi(r.y, d * d - 1u); a = v % d; b = v / d; }
            else { const uint32_t e = 1u + __umulhi(r.y, d - 1u); if (ch == 1) a = e; else b = e; }
          }
        }
        if (a) *xa = (uint8_t)mod_d(A, (uint32_t)*xa + a);                   // program.py:160-161
        if (b) *za = (uint8_t)mod_d(A, (uint32_t)*za + b);
        break;
      }
      default: break;     // I, Paulis (frames commute with them up to phase, program.py:82-89), BARRIER
    }
  }
}
