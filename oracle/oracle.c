/*
 * CPU ORACLE (C) — TEST INFRASTRUCTURE ONLY.  Not product code.
 *
 * Plain-C restatement of the reference's prime-dimension tableau path (events555/sdim), used as the
 * fast checker at sizes where the numpy oracle (oracle/tableau_oracle.py) is too slow and as the CPU
 * baseline of bench.py.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.
 *
 * Parity status: PINNED — tests/test_oracle_golden.py checks it against every reference golden vector
 * in tests/golden/ (records and all six final arrays) and against the numpy oracle.
 *
 * Layout follows the reference: per shot six arrays, x[q*n+g] = X exponent of generator (column) g on
 * qudit (row) q (sdim/tableau/dataclasses.py:14,24-39; sdim/tableau/tableau_prime.py:24-26).  Entries
 * are int32 kept reduced (the reference: int64, reduced every 64 gates, sdim/program.py:317-318).
 * Shots are independent (sdim/program.py:308) and are spread over OpenMP threads.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int n, d, po, order;
  int32_t *x, *z, *p, *dx, *dz, *dp; /* x,z,dx,dz: n*n; p,dp: n */
  int32_t *xs, *zs, *f, *dot;        /* scratch: n each */
} Tab;

static int mod(int64_t v, int m) { int r = (int)(v % m); return r < 0 ? r + m : r; }

static int tab_alloc(Tab* t, int n, int d) {
  t->n = n; t->d = d; t->po = (d % 2 == 0) ? 2 : 1; t->order = d * t->po; /* dataclasses.py:88-106 */
  size_t nn = (size_t)n * n;
  int32_t* blk = (int32_t*)calloc(4 * nn + 6 * (size_t)n, sizeof(int32_t));
  if (!blk) return -1;
  t->x = blk; t->z = blk + nn; t->dx = blk + 2 * nn; t->dz = blk + 3 * nn;
  t->p = blk + 4 * nn; t->dp = t->p + n; t->xs = t->dp + n; t->zs = t->xs + n; t->f = t->zs + n; t->dot = t->f + n;
  return 0;
}

static void tab_reset(Tab* t) { /* |0..0>: dataclasses.py:34-39, tableau_prime.py:81-86 */
  size_t nn = (size_t)t->n * t->n;
  memset(t->x, 0, (4 * nn + 2 * (size_t)t->n) * sizeof(int32_t));
  for (int q = 0; q < t->n; ++q) { t->z[(size_t)q * t->n + q] = 1; t->dx[(size_t)q * t->n + q] = 1; }
}

/* tableau_optimized.py:5-58 */
static void hadamard(Tab* t, int a, int inverse) {
  const int n = t->n, d = t->d, po = t->po, o = t->order;
  int32_t* X[2] = {t->x + (size_t)a * n, t->dx + (size_t)a * n};
  int32_t* Z[2] = {t->z + (size_t)a * n, t->dz + (size_t)a * n};
  int32_t* P[2] = {t->p, t->dp};
  for (int h = 0; h < 2; ++h)
    for (int i = 0; i < n; ++i) {
      const int xv = X[h][i], zv = Z[h][i];
      P[h][i] = mod((int64_t)P[h][i] - (int64_t)po * xv * zv, o);
      if (inverse) { X[h][i] = zv; Z[h][i] = mod(-xv, d); }
      else { X[h][i] = mod(-zv, d); Z[h][i] = xv; }
    }
}

/* tableau_optimized.py:62-96 */
static void phase_gate(Tab* t, int a, int inverse) {
  const int n = t->n, d = t->d, o = t->order, s = inverse ? -1 : 1;
  int32_t* X[2] = {t->x + (size_t)a * n, t->dx + (size_t)a * n};
  int32_t* Z[2] = {t->z + (size_t)a * n, t->dz + (size_t)a * n};
  int32_t* P[2] = {t->p, t->dp};
  for (int h = 0; h < 2; ++h)
    for (int i = 0; i < n; ++i) {
      const int xv = X[h][i];
      const int inc = (d % 2 == 0) ? xv * xv : (xv * (xv - 1)) / 2;
      P[h][i] = mod((int64_t)P[h][i] + s * inc, o);
      Z[h][i] = mod(Z[h][i] + s * xv, d);
    }
}

/* tableau_optimized.py:99-118 */
static void cnot(Tab* t, int c, int g, int inverse) {
  const int n = t->n, d = t->d, s = inverse ? -1 : 1;
  int32_t* X[2] = {t->x, t->dx};
  int32_t* Z[2] = {t->z, t->dz};
  for (int h = 0; h < 2; ++h)
    for (int i = 0; i < n; ++i) {
      X[h][(size_t)g * n + i] = mod(X[h][(size_t)g * n + i] + s * X[h][(size_t)c * n + i], d);
      Z[h][(size_t)c * n + i] = mod(Z[h][(size_t)c * n + i] - s * Z[h][(size_t)g * n + i], d);
    }
}

/* Net effect of the composite Paulis (tableau_gates.py:27-137): conjugation by X^a Z^b. */
static void pauli(Tab* t, int q, int a, int b) {
  const int n = t->n, po = t->po, o = t->order;
  int32_t* X[2] = {t->x + (size_t)q * n, t->dx + (size_t)q * n};
  int32_t* Z[2] = {t->z + (size_t)q * n, t->dz + (size_t)q * n};
  int32_t* P[2] = {t->p, t->dp};
  for (int h = 0; h < 2; ++h)
    for (int i = 0; i < n; ++i) P[h][i] = mod((int64_t)P[h][i] + (int64_t)po * (b * X[h][i] - a * Z[h][i]), o);
}

/* CZ = H^-1(t) CNOT(c,t) H(t) (tableau_gates.py:229-261), folded. */
static void cz(Tab* t, int a, int b, int inverse) {
  const int n = t->n, d = t->d, po = t->po, o = t->order, s = inverse ? -1 : 1;
  int32_t* X[2] = {t->x, t->dx};
  int32_t* Z[2] = {t->z, t->dz};
  int32_t* P[2] = {t->p, t->dp};
  for (int h = 0; h < 2; ++h)
    for (int i = 0; i < n; ++i) {
      const int xa = X[h][(size_t)a * n + i], xb = X[h][(size_t)b * n + i];
      P[h][i] = mod((int64_t)P[h][i] + (int64_t)s * po * xa * xb, o);
      Z[h][(size_t)a * n + i] = mod(Z[h][(size_t)a * n + i] + s * xb, d);
      Z[h][(size_t)b * n + i] = mod(Z[h][(size_t)b * n + i] + s * xa, d);
    }
}

/* tableau_gates.py:298-329 (prime branch): net effect is a swap of the two qudit rows. */
static void swap_rows(Tab* t, int a, int b) {
  const int n = t->n;
  int32_t* blocks[4] = {t->x, t->z, t->dx, t->dz};
  for (int k = 0; k < 4; ++k)
    for (int i = 0; i < n; ++i) {
      const int32_t v = blocks[k][(size_t)a * n + i];
      blocks[k][(size_t)a * n + i] = blocks[k][(size_t)b * n + i];
      blocks[k][(size_t)b * n + i] = v;
    }
}

static int inverse_mod(int v, int d) {
  for (int e = 1; e < d; ++e)
    if ((v * e) % d == 1) return e;
  return 0;
}

/* tableau_prime.py:262-363.  Returns value; *det = 1 for a deterministic outcome. */
static int measure(Tab* t, int q, int draw, int* det, int* nnz) {
  const int n = t->n, d = t->d, po = t->po, o = t->order;
  int piv = -1;
  for (int i = 0; i < n; ++i)
    if (t->x[(size_t)q * n + i] != 0) { piv = i; break; } /* first anticommuting stabilizer, :273-283 */
  if (piv < 0) {
    /* _det_measurement (:336-363), restructured row-wise: running ancilla_z per qudit row */
    int64_t ap = 0, cross = 0, sdg = 0;
    const int32_t* f = t->dx + (size_t)q * n;
    int na = 0; /* generators with a non-zero factor, in increasing order (the reference skips the others, :351) */
    for (int i = 0; i < n; ++i) {
      if (!f[i]) continue;
      ap += (int64_t)f[i] * t->p[i];
      t->xs[na++] = i;
    }
    *nnz += na;
    for (int r = 0; r < n; ++r) {
      const int32_t* xr = t->x + (size_t)r * n;
      const int32_t* zr = t->z + (size_t)r * n;
      int64_t az = 0;
      for (int k = 0; k < na; ++k) {
        const int i = t->xs[k], fi = f[i];
        cross += az * (fi * xr[i]) % d;
        az = (az + fi * zr[i]) % d;
        sdg += (int64_t)xr[i] * zr[i] * (fi * (fi - 1) / 2) % d;
      }
    }
    ap = mod(ap + po * (int64_t)mod(cross + po * sdg, d), o);
    *det = 1;
    int64_t neg = -ap; /* (-ap // po) % d with floor division, :362 */
    int64_t fl = (neg >= 0) ? neg / po : -((-neg + po - 1) / po);
    return mod(fl, d);
  }
  /* exponentiate (:365-380) so that the pivot has X[q] == 1 */
  const int v = t->x[(size_t)q * n + piv];
  int64_t sd = 0;
  if (v != 1) {
    const int e = inverse_mod(v, d);
    int64_t raw = 0;
    for (int r = 0; r < n; ++r) raw += (int64_t)t->x[(size_t)r * n + piv] * t->z[(size_t)r * n + piv];
    t->p[piv] = mod((int64_t)t->p[piv] * e + (raw % d) * (e * (e - 1) / 2) * po, o);
    for (int r = 0; r < n; ++r) {
      t->x[(size_t)r * n + piv] = mod((int64_t)t->x[(size_t)r * n + piv] * e, d);
      t->z[(size_t)r * n + piv] = mod((int64_t)t->z[(size_t)r * n + piv] * e, d);
    }
  }
  for (int r = 0; r < n; ++r) {
    t->xs[r] = t->x[(size_t)r * n + piv];
    t->zs[r] = t->z[(size_t)r * n + piv];
    sd += t->xs[r] * t->zs[r];
  }
  sd %= d;
  const int ps = t->p[piv];
  /* _random_measurement (:294-334) as a rank-1 update per block; iterations only read the pivot column */
  int32_t* X[2] = {t->dx, t->x};
  int32_t* Z[2] = {t->dz, t->z};
  int32_t* P[2] = {t->dp, t->p};
  for (int h = 0; h < 2; ++h) {
    for (int i = 0; i < n; ++i) {
      t->f[i] = mod(-X[h][(size_t)q * n + i], d);
      t->dot[i] = 0;
    }
    if (h == 1) t->f[piv] = 0;
    for (int i = 0; i < n; ++i) *nnz += (t->f[i] != 0);
    for (int r = 0; r < n; ++r) {
      const int s = t->xs[r], u = t->zs[r];
      if (!s && !u) continue;
      int32_t* xr = X[h] + (size_t)r * n;
      int32_t* zr = Z[h] + (size_t)r * n;
      for (int i = 0; i < n; ++i) {
        t->dot[i] = (t->dot[i] + zr[i] * s) % d;
        xr[i] = (xr[i] + t->f[i] * s) % d;
        zr[i] = (zr[i] + t->f[i] * u) % d;
      }
    }
    for (int i = 0; i < n; ++i) {
      const int fi = t->f[i];
      const int64_t cp = (int64_t)t->dot[i] * fi + sd * (fi * (fi - 1) / 2) * po;
      P[h][i] = mod((int64_t)P[h][i] + (int64_t)fi * ps + po * cp, o);
    }
  }
  for (int r = 0; r < n; ++r) { /* destabilizer <- old pivot; stabilizer <- Z_q (:323-330) */
    t->dx[(size_t)r * n + piv] = t->xs[r];
    t->dz[(size_t)r * n + piv] = t->zs[r];
    t->x[(size_t)r * n + piv] = 0;
    t->z[(size_t)r * n + piv] = (r == q) ? 1 : 0;
  }
  t->dp[piv] = ps;
  t->p[piv] = mod(-(int64_t)draw * po, o); /* :331-333 */
  *det = 0;
  return draw;
}

/* Philox4x32-10; same streams as sdim_b200/rng.py and the device code. */
static void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int i = 0; i < 10; ++i) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    c[0] = n0; c[1] = (uint32_t)p1; c[2] = n2; c[3] = (uint32_t)p0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

/* Records and replay arrays are one byte per element for d <= 127 (bit 7 = deterministic) and uint16 for larger
 * primes (bit 15 = deterministic), as in include/sdimb.h. */
static int rd_elem(const void* a, int64_t i, int wide) {
  return wide ? ((const uint16_t*)a)[i] : ((const uint8_t*)a)[i];
}

static void run_one(Tab* t, int64_t local, int64_t gshot, const int32_t* ops, int64_t n_ops, void* records,
                    int64_t n_meas, const void* replay_meas, const void* replay_noise, const uint32_t* thresh,
                    const uint8_t* chan, int64_t n_noise, uint64_t seed, int32_t* meas_nnz) {
  const int d = t->d, wide = d > 127;
  tab_reset(t);
  for (int64_t i = 0; i < n_ops; ++i) { /* program.py:311-351 */
    const int op = ops[4 * i], a = ops[4 * i + 1], b = ops[4 * i + 2], slot = ops[4 * i + 3];
    switch (op) {
      case 0: break;
      case 1: pauli(t, a, 1, 0); break;
      case 2: pauli(t, a, d - 1, 0); break;
      case 3: pauli(t, a, 0, 1); break;
      case 4: pauli(t, a, 0, d - 1); break;
      case 5: hadamard(t, a, 0); break;
      case 6: hadamard(t, a, 1); break;
      case 7: phase_gate(t, a, 0); break;
      case 8: phase_gate(t, a, 1); break;
      case 9: cnot(t, a, b, 0); break;
      case 10: cnot(t, a, b, 1); break;
      case 11: cz(t, a, b, 0); break;
      case 12: cz(t, a, b, 1); break;
      case 13: swap_rows(t, a, b); break;
      case 14: case 15: case 16: {
        if (op == 15) hadamard(t, a, 1); /* tableau_gates.py:292-296 */
        int draw;
        if (replay_meas) draw = rd_elem(replay_meas, local * n_meas + slot, wide);
        else {
          uint32_t c[4] = {(uint32_t)gshot, (uint32_t)((uint64_t)gshot >> 32), (uint32_t)slot, 0u};
          philox(c, (uint32_t)seed, (uint32_t)(seed >> 32));
          draw = (int)(((uint64_t)c[0] * (uint64_t)d) >> 32);
        }
        int det = 0, nnz = 0;
        const int m = measure(t, a, draw, &det, &nnz);
        if (meas_nnz) meas_nnz[slot] = nnz;
        if (wide) ((uint16_t*)records)[local * n_meas + slot] = (uint16_t)((m & 0x7FFF) | (det ? 0x8000 : 0));
        else ((uint8_t*)records)[local * n_meas + slot] = (uint8_t)((m & 0x7F) | (det ? 0x80 : 0));
        if (op == 16 && m) pauli(t, a, (d - m) % d, 0); /* program.py:335-339 */
        break;
      }
      case 17: {
        int na = 0, nb = 0;
        if (replay_noise) { na = rd_elem(replay_noise, (local * n_noise + slot) * 2, wide); nb = rd_elem(replay_noise, (local * n_noise + slot) * 2 + 1, wide); }
        else {
          uint32_t c[4] = {(uint32_t)gshot, (uint32_t)((uint64_t)gshot >> 32), (uint32_t)slot, 1u};
          philox(c, (uint32_t)seed, (uint32_t)(seed >> 32));
          if ((c[0] >> 8) >= thresh[slot]) { /* program.py:486-507 */
            if (chan[slot] == 0) { const uint32_t r = 1u + (uint32_t)(((uint64_t)c[1] * (uint64_t)((uint32_t)d * (uint32_t)d - 1u)) >> 32); na = r % d; nb = r / d; }
            else { const uint32_t e = 1u + (uint32_t)(((uint64_t)c[1] * (uint64_t)(d - 1)) >> 32); if (chan[slot] == 1) na = e; else nb = e; }
          }
        }
        if (na || nb) pauli(t, a, na, nb);
        break;
      }
      default: break;
    }
  }
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* final (nullable): the LAST shot's six arrays as int64, concatenated x,z,dx,dz (n*n each) then p,dp (n each).
 * meas_nnz (nullable, [n_meas]): for the LAST shot, the number of generators with a non-zero factor in each
 * measurement (the ones the reference does not skip, tableau_prime.py:308,315,351) - used for byte accounting. */
int oracle_run(int n, int d, int64_t shots, int64_t shot_offset, const int32_t* ops, int64_t n_ops, void* records,
               int64_t n_meas, const void* replay_meas, const void* replay_noise, const uint32_t* thresh,
               const uint8_t* chan, int64_t n_noise, uint64_t seed, int64_t* final, int32_t* meas_nnz, int threads) {
  if (n < 1 || d < 2 || shots < 0) return -1;
  int failed = 0;
#ifdef _OPENMP
  if (threads < 1) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
#pragma omp parallel num_threads(threads)
  {
    Tab t;
    if (tab_alloc(&t, n, d) != 0) {
#pragma omp atomic write
      failed = 1;
    } else {
#pragma omp for schedule(dynamic, 1)
      for (int64_t s = 0; s < shots; ++s) {
        run_one(&t, s, shot_offset + s, ops, n_ops, records, n_meas, replay_meas, replay_noise, thresh, chan, n_noise,
                seed, (s == shots - 1) ? meas_nnz : 0);
        if (final && s == shots - 1) {
          const size_t nn = (size_t)n * n;
          const int32_t* src[4] = {t.x, t.z, t.dx, t.dz};
          for (int k = 0; k < 4; ++k)
            for (size_t i = 0; i < nn; ++i) final[k * nn + i] = src[k][i];
          for (int i = 0; i < n; ++i) { final[4 * nn + i] = t.p[i]; final[4 * nn + n + i] = t.dp[i]; }
        }
      }
      free(t.x);
    }
  }
  return failed ? -2 : 0;
}
