// Cluster interpreter: ONE tableau per thread-block cluster, for tableaus too large for one SM to serve.
// Part of libsdimb (sdim_b200/csrc); included by sdimb.cu inside its anonymous namespace.
//
// A single large tableau (config 5: n = 4096, d = 5 / 7, 64 MiB) gives the one-CTA interpreter of lanes.cuh
// one SM's worth of L2 bandwidth (~100 GB/s): a pivot-column walk alone moves 16 K sectors.  Here a cluster of
// C CTAs (8 portable, 16 with the non-portable opt-in; all on one GPC) owns the shot:
//
//   gates         lane-partitioned: CTA c owns lane words [c*wpc, (c+1)*wpc) of EVERY row.  Generator lanes are
//                 independent under gates, so CTAs never talk during a gate segment.  Inside a CTA the threads
//                 form `ngroups` = blockDim / wpc groups; on a scheduled stream (sdimb_schedule: layers of gates
//                 on disjoint qudits) group g executes the gates whose index in the layer is g mod ngroups, one
//                 lane word per thread, so ngroups * C gates are in flight per layer.  A thread accumulates its
//                 phase increments in a register; increments are folded into the phase words of the HBM store
//                 before the next measurement.
//   measurement   row-partitioned: CTA c owns rows [c*rpc, (c+1)*rpc) for the pivot-column walk, the rank-1
//                 update and the column writes.  Row q (pivot search, factors) is read by every CTA.  Partial
//                 dot products, x_p . z_p and the deterministic-branch sums are reduced with distributed-shared-
//                 memory atomics into the CTA that owns the lane word (phases are lane-partitioned like gates).
//
// Synchronisation is the hardware cluster barrier (release/acquire at cluster scope; 1 000 - 1 500 cycles here
// with the L1 invalidate and the skew between CTAs), used only where ownership changes hands:
//   A   before a measurement that follows gates (lane-partitioned writes -> row-partitioned reads)
//   B1  random branch only: every CTA has read row q before its owner updates it
//   B2  random branch, and deterministic branch of a RESET (every CTA needs the outcome for the correction):
//       the partial sums have arrived in their owners' shared memory
// After B2 a CTA only writes phase words it owns, which no other CTA reads before the next B2 (the old pivot phase
// travels through shared memory), so gates after a measurement need no barrier.  A deterministic M / M_X writes
// nothing to the tableau and runs with NO cluster barrier: each CTA sends its partial sums to a ring slot in CTA 0
// in one packed atomic and moves on, CTA 0 writes the record one measurement later; the row of the next
// measurement is fetched with cp.async meanwhile.  The accumulators of the barrier-carrying measurements are
// double-buffered by their parity and zeroed by their owner when consumed: a CTA that races ahead adds into the
// other set, and cannot reach the measurement after that before every CTA has finished this one.
//
// Same arithmetic as lanes.cuh (its helpers are reused); same reference behaviour (file:line citations there).
#pragma once

namespace clusters {

namespace cg = cooperative_groups;

constexpr int kClusterThreads = 1024;

// Developer build (-DSDIMB_PHASE_CLOCKS): thread 0 of CTA 0 accumulates clock64() deltas per measurement phase.
#ifdef SDIMB_PHASE_CLOCKS
__device__ unsigned long long g_phase_clk[32];
#define PHASE_BEGIN() long long ph_last = clock64()
#define PHASE(i) do { if (threadIdx.x == 0 && g.c == 0) { const long long ph_now = clock64(); \
    g_phase_clk[i] += (unsigned long long)(ph_now - ph_last); g_phase_clk[16 + (i)] += 1; ph_last = ph_now; } } while (0)
#else
#define PHASE_BEGIN() do {} while (0)
#define PHASE(i) do {} while (0)
#endif

struct CScratch : Scratch {
  uint32_t* dotg;   // [2][4*wpc] cluster-reduced dot products of the lane words this CTA owns (DSMEM target)
  uint32_t* acc;    // [blockDim] phase increments of the gate groups, staged for the fold
  uint32_t* xch;    // [2][4]     cluster accumulators: x_p . z_p | old pivot phase | det a1 | det rows (DSMEM target)
  uint32_t* rowx;   // [2][W/4]   X half of the row being measured (stabilizer lanes, then destabilizer lanes), and
                    //            of the row of the next measurement while it is being prefetched
  uint32_t* ring;   // [kRing]    CTA 0: packed partial sums of barrier-free deterministic measurements (DSMEM target)
  uint32_t* red2;   // [32]       reduction scratch of the factor-list scan (block_min / block_sum own `red`)
};

struct CGeo {
  int c, C;         // rank in cluster, cluster size
  int wpc;          // lane words per CTA (multiple of 32)
  int r0, r1;       // rows owned in measurements
  int group, ngroups, w;   // gate group of this thread, groups per CTA, lane word owned (>= W/4: none)
};

// bytes of dynamic shared memory per CTA
inline size_t cluster_smem_bytes(int np, int wpc) {
  const size_t W = 2 * (size_t)np, wz = W / 4;
  return 4 * W + 4 * wz + 8 * wz + 4 * 8 * (size_t)wpc + 4 * kClusterThreads + 4 * 32 + 4 * 4 + 4 * 8 + 4 * 64 + 4 * 32   // dot fw rowx dotg acc red cnt xch ring red2
         + 2 * (size_t)np + 2 * (size_t)np + 2 * wz + (size_t)np + (size_t)np + 128 + 64;       // ar br aw xs zs inv
}

// The launch shapes: fewer threads leave more registers per thread and fewer warps to walk the (uniform) control
// flow of a measurement; more threads give more gate groups.
typedef void (*ClusterKernel)(KParams);
template <int THREADS>
__global__ void interp_cluster_kernel(const __grid_constant__ KParams p);
inline ClusterKernel cluster_kernel_for(int threads);

// lane words per CTA for a cluster of C: ceil(wz / C) rounded up to whole warps
inline int cluster_wpc(int np, int C) {
  const int wz = np / 2;
  return ((wz + C - 1) / C + 31) / 32 * 32;
}

// Fold the gate groups' phase increments into the phase words this CTA owns.
__device__ __forceinline__ void fold_phases(uint8_t* T, const KParams& p, CScratch& S, const CGeo& g, uint32_t& pw) {
  const Swar So = make_swar(p.A.order);
  S.acc[threadIdx.x] = pw;
  pw = 0u;
  __syncthreads();
  if ((int)threadIdx.x < g.wpc && g.w < p.W / 4) {
    uint32_t sum = 0u;
    for (int k = 0; k < g.ngroups; ++k) sum = swar_add(So, sum, S.acc[k * g.wpc + threadIdx.x]);
    if (sum) {
      uint32_t* Pw = reinterpret_cast<uint32_t*>(T + p.phase_off) + g.w;
      *Pw = swar_add(So, *Pw, sum);
    }
  }
  __syncthreads();
}

// Row q is what every branch of a measurement starts from (pivot search, factors, deterministic factor list), and
// on the HBM store every dependent pass over it is a memory round trip.  It is staged ONCE: the X half of the row
// (all W lanes, stabilizers then destabilizers) goes to shared memory as 16-byte vectors, vector v by thread
// v mod blockDim — synchronously (stage_row), or ahead of time with cp.async while the previous measurement runs
// (prefetch_row; valid only if that measurement turns out not to write the tableau).  Same thread, same vector in
// both, so a thread may overwrite or read back its own vectors without a barrier.
__device__ __forceinline__ void stage_row(const uint8_t* rowq, const KParams& p, uint32_t* rowx) {
  const uint4* src = reinterpret_cast<const uint4*>(rowq);
  uint4* dst = reinterpret_cast<uint4*>(rowx);
  for (int v = threadIdx.x; v < p.W / 16; v += blockDim.x) dst[v] = src[v];
}

__device__ __forceinline__ void prefetch_row(const uint8_t* rowq, const KParams& p, uint32_t* rowx) {
  for (int v = threadIdx.x; v < p.W / 16; v += blockDim.x) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(rowx + 4 * v);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(rowq + 16 * v) : "memory");   // L2 only
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ void prefetch_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Pivot candidate — FIRST stabilizer with an X component on q (tableau_prime.py:273-283) — from the vectors this
// thread staged itself.
__device__ __forceinline__ uint32_t pivot_candidate_staged(const KParams& p, const uint32_t* rowx) {
  const int ws = p.np / 4;
  for (int v = threadIdx.x; 4 * v < ws; v += blockDim.x) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t x = rowx[4 * v + k];
      if (x && 4 * v + k < ws) return 4u * (4 * v + k) + ((__ffs(x) - 1) >> 3);
    }
  }
  return kNoPivot;
}

// Factors f = -X[q,i] mod d of the four lanes of word w; the pivot itself is skipped.
__device__ __forceinline__ uint32_t factor_word(const Arith& A, uint32_t xq_w, int w, uint32_t piv) {
  uint32_t fw = 0;
  if (xq_w) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (4u * w + k != piv) fw |= neg_d(A, byte_of(xq_w, k)) << (8 * k);
  }
  return fw;
}

// factor_words (lanes.cuh) from the staged row
__device__ __forceinline__ void factor_words_staged(const uint32_t* rowx, const KParams& p, Scratch& S, uint32_t piv) {
  for (int w = threadIdx.x; w < p.W / 4; w += blockDim.x) {
    const uint32_t fw = factor_word(p.A, rowx[w], w, piv);
    S.fw[w] = fw;
    if (fw) {
      S.aw[atomicAdd(&S.cnt[1], 1u)] = (uint16_t)w;
      *reinterpret_cast<uint4*>(S.dot + 4 * w) = make_uint4(0, 0, 0, 0);
    }
  }
}

// det_list (lanes.cuh) from the staged row: ordered compaction of the generators with f_i = destab X[q,i] != 0,
// four generators (one word) per thread and pass — one pass for n <= 4 * blockDim.  (Sixteen generators per thread
// and a single pass at n = 4096 measured slower: the unrolled 16-way emit costs more than the second pass.)  The
// phase loads behind `a1` are issued here and not waited for: the last pass leaves them in (pf, pv) and the caller
// adds sum_k pf[k] * pv[k] to a1 once its own loads are in flight.  Uses its own reduction scratch (red2), so it
// needs no barrier against block_min before it; ends with a barrier (the lists are published).
__device__ __forceinline__ int det_list_staged(const uint32_t* rowx, const uint8_t* P8, const KParams& p, CScratch& S,
                                               int pw_lo, int pw_hi, uint32_t& a1, uint32_t (&pf)[4],
                                               uint32_t (&pv)[4]) {
  const int ws = p.np / 4, nt = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int total = 0;
  a1 = 0;
  pf[0] = pf[1] = pf[2] = pf[3] = 0u;
  pv[0] = pv[1] = pv[2] = pv[3] = 0u;
  for (int base = 0; base < ws; base += nt) {
    const int j = base + tid;
    const uint32_t fq = (j < ws) ? rowx[ws + j] : 0u;          // destab X[q, 4j .. 4j+3]; padding lanes hold 0
    const int c = __popc((fq + 0x7F7F7F7Fu) & 0x80808080u);    // non-zero bytes (every byte is < 128)
    // The list is short (a handful of generators out of n): most warps hold nothing and skip the scan.
    const uint32_t any = __ballot_sync(0xFFFFFFFFu, c != 0);
    a1 += pf[0] * pv[0] + pf[1] * pv[1] + pf[2] * pv[2] + pf[3] * pv[3];   // phases loaded by the previous pass
    pf[0] = pf[1] = pf[2] = pf[3] = 0u;
    int excl = 0, wtotal = 0;
    if (any) {                                                 // exclusive scan of c (0..4) inside the warp, by ballots
      const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, c & 1), b1 = __ballot_sync(0xFFFFFFFFu, c & 2),
                     b2 = __ballot_sync(0xFFFFFFFFu, c & 4), lt = (1u << lane) - 1u;
      excl = __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt);
      wtotal = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
    }
    if (lane == 0) S.red2[warp] = (uint32_t)wtotal;
    __syncthreads();
    const uint32_t rv = lane < (nt >> 5) ? S.red2[lane] : 0u;
    const int all = (int)__reduce_add_sync(0xFFFFFFFFu, rv);
    if (any) {
      int pos = total + (int)__reduce_add_sync(0xFFFFFFFFu, lane < warp ? rv : 0u) + excl;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t f = byte_of(fq, k);
        if (f) {
          const int i = 4 * j + k;
          S.ar[pos] = (uint16_t)i;
          S.xs[pos] = (uint8_t)f;
          if (j >= pw_lo && j < pw_hi) { pf[k] = f; pv[k] = P8[i]; }   // issued here, consumed after the pass
          ++pos;
        }
      }
    }
    total += all;
    __syncthreads();
  }
  return total;
}

// Two block sums behind one barrier pair: a mod order, b mod d.  `fenced` = a barrier since the last use of `red`
// has already been passed by every thread (the leading barrier is skipped).
__device__ __forceinline__ void block_sum2(const Arith& A, uint32_t& a, uint32_t& b, uint32_t* red, bool fenced) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t wa = mod_o(A, __reduce_add_sync(0xFFFFFFFFu, a)), wb = mod_d(A, __reduce_add_sync(0xFFFFFFFFu, b));
  if (!fenced) __syncthreads();
  if (lane == 0) red[threadIdx.x >> 5] = wa | (wb << 16);
  __syncthreads();
  const uint32_t v = lane < (blockDim.x >> 5) ? red[lane] : 0u;
  a = mod_o(A, __reduce_add_sync(0xFFFFFFFFu, v & 0xFFFFu));
  b = mod_d(A, __reduce_add_sync(0xFFFFFFFFu, v >> 16));
}

// State of the measurement pipeline, identical in every thread of the cluster.
struct MState {
  uint32_t seq;        // measurements so far: parity selects the row buffer, seq mod kRing the ring slot
  uint32_t bseq;       // measurements so far that exchanged through xch / dotg (they all contain a cluster barrier):
                       // parity selects the accumulator set
  int pref_q;          // qudit whose row sits (or is arriving) in the row buffer of the NEXT measurement, -1: none
  int free_run;        // measurements since the last cluster barrier
  int pend_slot;       // ring slot of a deterministic measurement whose record CTA 0 still has to write, -1: none
  int64_t pend_rec;    // ... and where
};

// CTA 0, thread 0: collect the partial sums of an earlier barrier-free deterministic measurement and write its
// record.  The slot is complete when every CTA's packed add (a1 | rows << 12 | 1 << 24) has landed.
__device__ __forceinline__ void drain_pending(const KParams& p, CScratch& S, const CGeo& g, MState& st) {
  if (st.pend_slot >= 0 && g.c == 0 && threadIdx.x == 0) {
    const Arith& A = p.A;
    volatile uint32_t* slot = S.ring + st.pend_slot;
    uint32_t v;
    do { v = *slot; } while ((v >> 24) != (uint32_t)g.C);
    *slot = 0u;
    const uint32_t ap = mod_o(A, mod_o(A, v & 0xFFFu) + A.po * mod_d(A, (v >> 12) & 0xFFFu));
    const uint32_t outcome = (A.po == 1) ? neg_d(A, ap) : (((ap + 1u) >> 1) & 1u);   // (-ap // po) % d  (tableau_prime.py:362)
    p.records[st.pend_rec] = (uint8_t)(outcome | SDIMB_REC_DET);
  }
  st.pend_slot = -1;
}

// Measurement of qudit q in the Z basis (tableau_prime.py:262-363) by the whole cluster.  The caller has made all
// earlier writes cluster-visible.  `next_q` = row to prefetch for the measurement that follows directly (-1: none);
// `need_outcome` = the caller applies a correction that depends on the outcome (RESET).
// Dependent memory round trips: row q | pivot column | rank-1 update (random), row q | generator columns (det.).
// A deterministic measurement writes nothing to the tableau: unless its outcome is needed it runs WITHOUT a
// cluster barrier — every CTA sends its partial sums to CTA 0 in one packed shared-memory atomic and moves on;
// CTA 0 writes the record one measurement later (drain_pending).  kRing slots bound how far CTAs may drift apart:
// a cluster barrier is forced after kRing / 2 barrier-free measurements.
constexpr int kRing = 64;

__device__ uint32_t measure(cg::cluster_group& cl, uint8_t* T, const KParams& p, CScratch& S, const CGeo& g,
                            MState& st, int q, int next_q, bool need_outcome, int64_t slot, int64_t shot_local,
                            uint32_t draw) {
  const Arith& A = p.A;
  const int npad = p.np, tid = threadIdx.x;
  const int wz = p.W / 4;
  const uint32_t rpar = st.seq & 1u;
  uint8_t* rowq = T + (int64_t)q * p.row_bytes;
  uint8_t* P8 = T + p.phase_off;
  uint32_t* rowx = S.rowx + (size_t)rpar * wz;
  if (tid < 3) S.cnt[tid] = 0;
  PHASE_BEGIN();
  // every CTA stages row q and finds the pivot itself: same row, same answer, no exchange
  prefetch_wait();                                             // this thread's copies into rowx (if any) have landed
  if (st.pref_q != q) stage_row(rowq, p, rowx);
  if (next_q >= 0) prefetch_row(T + (int64_t)next_q * p.row_bytes, p, S.rowx + (size_t)(rpar ^ 1u) * wz);
  const uint32_t piv = block_min(pivot_candidate_staged(p, rowx), S.red);   // its barriers publish rowx and the counters
  st.pref_q = next_q;
  st.seq++;
  drain_pending(p, S, g, st);
  PHASE(0);

  uint32_t outcome, rec;
  if (piv != kNoPivot) {
    // -- random branch (tableau_prime.py:294-334) ---------------------------------------------------------
    st.pref_q = -1;                                            // the tableau changes: a prefetched row is stale
    st.free_run = 0;
    const uint32_t par = st.bseq++ & 1u;
    uint32_t* xch = S.xch + 4 * par;
    uint32_t* dotg = S.dotg + (size_t)par * 4 * g.wpc;
    const uint32_t e = S.inv[byte_of(rowx[piv >> 2], piv & 3)];
    factor_words_staged(rowx, p, S, piv);                      // whole row q, redundantly per CTA
    const int own_p = (int)(piv >> 2) / g.wpc, own_d = (int)((npad + piv) >> 2) / g.wpc;
    // Loads whose latency hides behind the column walk.  Phase words are only ever written by the CTA that owns
    // them, so the owner may read them before the barriers: the old pivot phase (broadcast below, every CTA needs
    // it) and this thread's own phase word if it holds a non-zero factor.
    uint32_t ps_mine = 0, fw_own = 0, ph_own = 0;
    if (g.c == own_p && tid < g.C) ps_mine = P8[piv];
    uint32_t* Pw = reinterpret_cast<uint32_t*>(P8) + g.w;
    if (tid < g.wpc && g.w < wz) {
      fw_own = factor_word(A, rowx[g.w], g.w, piv);
      if (fw_own) ph_own = *Pw;
    }
    PHASE(6);
    cl.sync();                                                 // B1: row q has been read everywhere
    PHASE(7);
    uint32_t sd_raw = column_walk(T, p, S, piv, e, g.r0, g.r1);
    sd_raw = mod_d(A, block_sum(sd_raw, S.red));               // barrier: publishes xs/zs/ar/br/fw/aw/dot/cnt
    PHASE(8);
    const int nr_a = (int)S.cnt[0], nw_a = (int)S.cnt[1];
    rank1_update(T, p, S, nr_a, nw_a);                         // own rows x all active lane words
    __syncthreads();
    PHASE(9);
    column_writes(T, p, S, q, piv, nr_a, (int)S.cnt[2]);       // own rows
    if (nr_a > 0) {                                            // partial dots -> owners of the lane words
      for (int i = tid; i < nw_a; i += blockDim.x) {
        const int w = S.aw[i], o = w / g.wpc;
        uint32_t* dst = cl.map_shared_rank(dotg, o) + 4 * (w - o * g.wpc);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t v = mod_d(A, S.dot[4 * w + k]);
          if (v) atomicAdd(dst + k, v);
        }
      }
    }
    if (tid < g.C) {
      if (sd_raw) atomicAdd(cl.map_shared_rank(xch, tid), sd_raw);
      if (g.c == own_p) *cl.map_shared_rank(xch + 1, tid) = ps_mine;
    }
    PHASE(10);
    cl.sync();                                                 // B2: sums are complete
    PHASE(11);
    const uint32_t sd_all = mod_d(A, xch[0]), ps_old = xch[1];
    const uint32_t ps = mod_o(A, ps_old * e + A.po * mod_d(A, sd_all * mod_d(A, (e * (e - 1u)) >> 1)));
    const uint32_t sd = mod_d(A, mod_d(A, sd_all * e) * e);    // x_p . z_p after exponentiation
    if (fw_own) {                                              // phases of the lane words this CTA owns
      uint32_t* dg = dotg + 4 * tid;
      *Pw = phase_word_update(A, fw_own, ph_own, dg, sd, ps);
      *reinterpret_cast<uint4*>(dg) = make_uint4(0, 0, 0, 0);
    }
    outcome = draw;
    __syncthreads();
    if (tid == 0) {
      xch[0] = 0;
      if (g.c == own_d) P8[npad + piv] = (uint8_t)ps;                           // destabilizer p <- old pivot (phase)
      if (g.c == own_p) P8[piv] = (uint8_t)mod_o(A, A.order - outcome * A.po);  // stabilizer p <- Z_q, phase -m*po
      if (g.c == 0) p.records[shot_local * p.rec_stride + slot] = (uint8_t)outcome;
    }
    // no closing barrier: the next op that reads what thread 0 just wrote is a measurement (two block barriers in
    // its pivot search come first) or a gate (phase words are not read by gates; the fold starts with a barrier)
    PHASE(12);
    return outcome;
  }
  // -- deterministic branch (tableau_prime.py:336-363): nothing is written to the tableau ------------------
  uint32_t a1, pf[4], pv[4];
  const int total = det_list_staged(rowx, P8, p, S, g.c * g.wpc, (g.c + 1) * g.wpc, a1, pf, pv);   // phases of owned words
  PHASE(1);
  uint32_t part = det_rows(T, p, S, total, g.r0, g.r1);
  PHASE(2);
  a1 = mod_o(A, a1 + pf[0] * pv[0] + pf[1] * pv[1] + pf[2] * pv[2] + pf[3] * pv[3]);
  block_sum2(A, a1, part, S.red, true);                        // fenced by the barrier that published the lists; its
                                                               // own barrier also retires them
  PHASE(3);
  if (!need_outcome) {
    // nobody but the record needs the outcome: one packed atomic to CTA 0, no barrier
    const int rs = (int)((st.seq - 1u) % kRing);
    if (tid == 0) atomicAdd(cl.map_shared_rank(S.ring + rs, 0), a1 | (part << 12) | (1u << 24));
    st.pend_slot = rs;
    st.pend_rec = shot_local * p.rec_stride + slot;
    if (++st.free_run >= kRing / 2) {                          // bound the drift between CTAs
      cl.sync();
      st.free_run = 0;
    }
    PHASE(4);
    return 0u;
  }
  uint32_t* xch = S.xch + 4 * (st.bseq++ & 1u);
  if (tid < g.C) {
    if (a1) atomicAdd(cl.map_shared_rank(xch + 2, tid), a1);
    if (part) atomicAdd(cl.map_shared_rank(xch + 3, tid), part);
  }
  cl.sync();                                                   // B2
  st.free_run = 0;
  PHASE(5);
  const uint32_t ap = mod_o(A, mod_o(A, xch[2]) + A.po * mod_d(A, xch[3]));
  outcome = (A.po == 1) ? neg_d(A, ap) : (((ap + 1u) >> 1) & 1u);   // (-ap // po) % d  (tableau_prime.py:362)
  rec = outcome | SDIMB_REC_DET;
  __syncthreads();
  if (tid == 0) {
    xch[2] = xch[3] = 0;
    if (g.c == 0) p.records[shot_local * p.rec_stride + slot] = (uint8_t)rec;
  }
  __syncthreads();
  PHASE(12);
  return outcome;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) interp_cluster_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  cg::cluster_group cl = cg::this_cluster();
  const Arith& A = p.A;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
  const int wz = p.W / 4;
  CGeo g;
  g.c = (int)cl.block_rank();
  g.C = (int)cl.num_blocks();
  g.wpc = p.wpc;
  const int rpc = (p.n + g.C - 1) / g.C;
  g.r0 = min(p.n, g.c * rpc);
  g.r1 = min(p.n, g.r0 + rpc);
  g.ngroups = nt / g.wpc;
  g.group = tid / g.wpc;
  g.w = g.group < g.ngroups ? g.c * g.wpc + tid % g.wpc : wz;
  // gate groups only help on a layered stream; on a plain stream group 0 runs every gate in order
  const bool layered = (p.flags & SDIMB_SCHEDULED) != 0 && g.ngroups > 1;
  // does this warp execute gates at all?  (warps are whole groups; the last CTAs may own no lane word)
  const bool warp_gates = __any_sync(0xFFFFFFFFu, g.w < wz) && (layered || g.group == 0);

  CScratch S;
  S.dot = reinterpret_cast<uint32_t*>(smem);
  S.fw = S.dot + p.W;
  S.rowx = S.fw + wz;
  S.dotg = S.rowx + 2 * wz;
  S.acc = S.dotg + 8 * g.wpc;
  S.red = S.acc + kClusterThreads;   // sized for the largest launch shape
  S.cnt = S.red + 32;
  S.xch = S.cnt + 4;
  S.ring = S.xch + 8;
  S.red2 = S.ring + kRing;
  S.ops = nullptr;
  S.ar = reinterpret_cast<uint16_t*>(S.red2 + 32);
  S.br = S.ar + p.np;
  S.aw = S.br + p.np;
  S.xs = reinterpret_cast<uint8_t*>(S.aw + wz);
  S.zs = S.xs + p.np;
  S.inv = S.zs + p.np;

  for (uint32_t v = tid; v < A.d; v += nt) {                  // inverse table; inv[0] unused
    uint32_t e = 0;
    for (uint32_t c2 = 1; c2 < A.d; ++c2)
      if (mod_d(A, v * c2) == 1u) { e = c2; break; }
    S.inv[v] = (uint8_t)e;
  }
  for (int i = tid; i < 8 * g.wpc; i += nt) S.dotg[i] = 0u;
  if (tid < 8) S.xch[tid] = 0u;
  if (tid < kRing) S.ring[tid] = 0u;
  cl.sync();                                                   // accumulators are zero before any CTA adds into them

  const int64_t n_clusters = gridDim.x / g.C, cluster_id = blockIdx.x / g.C;
  MState st;
  st.seq = st.bseq = 0u;
  st.pref_q = st.pend_slot = -1;
  st.free_run = 0;
  st.pend_rec = 0;
  for (int64_t shot = cluster_id; shot < p.shots; shot += n_clusters) {
    uint8_t* T = p.tab + shot * p.shot_bytes;
    if (p.flags & SDIMB_FRESH) {                               // |0...0>, split over the cluster
      uint4* v = reinterpret_cast<uint4*>(T);
      const int64_t nvec = p.shot_bytes / 16;
      for (int64_t i = (int64_t)g.c * nt + tid; i < nvec; i += (int64_t)g.C * nt) v[i] = make_uint4(0, 0, 0, 0);
      cl.sync();
      for (int q = g.c * nt + tid; q < p.n; q += g.C * nt) {
        uint8_t* row = T + (int64_t)q * p.row_bytes;
        row[p.W + q] = 1;      // Z[q][stab q]
        row[p.np + q] = 1;     // X[q][destab q]
      }
      cl.sync();
    }
    uint32_t pw = 0u;                                          // this thread's phase increments since the last fold
    bool dirty = false;                                        // gates since the last fold (uniform over the cluster)

    for (int64_t i0 = 0; i0 < p.n_ops; i0 += 32) {
      // every warp fetches the same 32 ops, one per lane, and keeps the collective ops plus its group's gates;
      // the fetching lane resolves N1 events and measurement draws
      int4 mine = make_int4(SDIMB_OP_I, 0, 0, 0);
      if (i0 + lane < p.n_ops) mine = __ldg(p.ops + i0 + lane);
      const int idx = (int)(((uint32_t)mine.x) >> SDIMB_OP_INDEX_SHIFT);
      mine.x &= SDIMB_OP_MASK;
      const bool meas = mine.x >= SDIMB_OP_M && mine.x <= SDIMB_OP_RESET;
      const bool is_gate = mine.x != SDIMB_OP_I && !meas && mine.x != SDIMB_OP_BARRIER;
      bool live = meas || (mine.x == SDIMB_OP_BARRIER && layered) ||
                  (is_gate && warp_gates && (!layered || idx % g.ngroups == g.group));
      if (live && mine.x == SDIMB_OP_N1) {
        mine.z = (int)noise_event(p, mine.w, shot);
        live = mine.z != 0;
      }
      if (meas) {                                              // outcome this measurement takes if it is random
        if (p.replay_meas) {
          mine.z = p.replay_meas[shot * p.n_meas + mine.w];
        } else {
          const uint64_t gshot = (uint64_t)(p.shot_offset + shot);
          const uint4 r = philox4x32((uint32_t)gshot, (uint32_t)(gshot >> 32), (uint32_t)mine.w, 0u, (uint32_t)p.seed,
                                     (uint32_t)(p.seed >> 32));
          mine.z = (int)__umulhi(r.x, A.d);
        }
      }
      uint32_t todo = __ballot_sync(0xFFFFFFFFu, live);
      // positions of gates (executed by ANY group) not yet covered by a fold: identical in every warp of the cluster
      const uint32_t gate_mask = __ballot_sync(0xFFFFFFFFu, is_gate), meas_mask = __ballot_sync(0xFFFFFFFFu, meas);
      uint32_t pending = gate_mask;
#pragma unroll 1
      while (todo) {
        const int k = __ffs(todo) - 1;
        todo &= todo - 1;
        int4 op;
        op.x = __shfl_sync(0xFFFFFFFFu, mine.x, k);
        op.y = __shfl_sync(0xFFFFFFFFu, mine.y, k);
        op.z = __shfl_sync(0xFFFFFFFFu, mine.z, k);
        op.w = __shfl_sync(0xFFFFFFFFu, mine.w, k);
        if (is_unitary_like(op.x)) {
          uint32_t pa, pb;
          pauli_exponents(p, op, pa, pb);
          if (g.w < wz) pw = gate_word(T, p, op.x, op.y, op.z, pa, pb, load_rows(T, p, op.x, op.y, op.z, g.w), g.w, pw);
          continue;
        }
        if (op.x == SDIMB_OP_BARRIER) { __syncthreads(); continue; }
        // collective ops: M, M_X, RESET
        const uint32_t below = (1u << k) - 1u;
        bool after_gates = dirty || (pending & below) != 0;
        pending &= ~below;
        if (op.x == SDIMB_OP_M_X) {                            // tableau_gates.py:292-296: H^-1, then measure
          if (g.group == 0 && g.w < wz)
            pw = gate_word(T, p, SDIMB_OP_H_INV, op.y, -1, 0u, 0u, load_rows(T, p, SDIMB_OP_H_INV, op.y, -1, g.w), g.w, pw);
          after_gates = true;
        }
        if (after_gates) {
          fold_phases(T, p, S, g, pw);
          cl.sync();                                           // A: lane-partitioned writes -> row-partitioned reads
          st.free_run = 0;
          st.pref_q = -1;                                      // rows changed since the prefetch was issued
        }
        // the row of the measurement that follows DIRECTLY (no gate in between) can be fetched while this one runs
        int next_q = -1;
        {
          const uint32_t above = ~((2u << k) - 1u), nm = meas_mask & above;
          if (nm) {
            const int k2 = __ffs(nm) - 1;
            const int code2 = __shfl_sync(0xFFFFFFFFu, mine.x, k2), q2 = __shfl_sync(0xFFFFFFFFu, mine.y, k2);
            if ((gate_mask & above & ((1u << k2) - 1u)) == 0 && code2 != SDIMB_OP_M_X) next_q = q2;
          }
        }
        const uint32_t m = measure(cl, T, p, S, g, st, op.y, next_q, op.x == SDIMB_OP_RESET, op.w, shot, (uint32_t)op.z);
        dirty = false;
        if (op.x == SDIMB_OP_RESET && m) {                     // program.py:335-339: X^(-m) brings the qudit to |0>
          if (g.group == 0 && g.w < wz)
            pw = gate_word(T, p, SDIMB_OP_N1, op.y, -1, A.d - m, 0u, load_rows(T, p, SDIMB_OP_N1, op.y, -1, g.w), g.w, pw);
          dirty = true;
        }
      }
      dirty = dirty || pending != 0;                           // gates behind the batch's last measurement
    }
    drain_pending(p, S, g, st);
    prefetch_wait();
    fold_phases(T, p, S, g, pw);
    cl.sync();                                                 // the shot is complete in the store
    st.free_run = 0;
    st.pref_q = -1;
  }
  cl.sync();                                                   // no CTA leaves while its shared memory may be addressed
}

inline ClusterKernel cluster_kernel_for(int threads) {
  return threads >= 1024 ? interp_cluster_kernel<1024> : threads >= 512 ? interp_cluster_kernel<512> : interp_cluster_kernel<256>;
}

}  // namespace clusters
