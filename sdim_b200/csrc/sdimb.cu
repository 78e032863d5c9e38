// libsdimb — B200 (sm_100a) stabilizer-tableau engine behind the C ABI of include/sdimb.h.
//
// One CTA owns one shot's tableau and interprets the whole op stream on it, so a shot never
// leaves the SM between gates.  This file is the host side of the C ABI (argument checks, kernel
// choice, launches, the op-stream scheduler, the host-buffer entry); the device code lives in
//   common.cuh        mod-d arithmetic, SWAR lanes, Philox4x32-10, kernel parameter block
//   lanes.cuh         uint8-lane interpreter, any prime d <= 127 (shared memory or HBM store)
//   planes.cuh        bit-plane interpreter for d = 2, 3 (shared-memory resident, 1 or 4 warps per shot)
//   clusters.cuh      uint8-lane interpreter with one shot per thread-block cluster (large tableaus, few shots)
//   util_kernels.cuh  |0...0> fill and export of the uint8 store
//   frames.cuh        Pauli-frame sampler
//
// Reference behaviour restated (file:line in events555/sdim):
//   primitives       sdim/tableau/tableau_optimized.py:5-118
//   composites       sdim/tableau/tableau_gates.py:27-137,229-261,298-329  (folded to closed forms)
//   measurement      sdim/tableau/tableau_prime.py:262-380
//   shot loop/RESET  sdim/program.py:308-351
//   noise            sdim/program.py:486-507
#include "sdimb.h"

#include <cuda_runtime.h>
#include <cooperative_groups.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace {

#include "common.cuh"
#include "lanes.cuh"
#include "clusters.cuh"
#include "planes.cuh"
#include "util_kernels.cuh"
#include "frames.cuh"
#include "wide.cuh"
#include "lanes_gm.cuh"

bool is_prime(int d) {
  if (d < 2) return false;
  for (int f = 2; f * f <= d; ++f)
    if (d % f == 0) return false;
  return true;
}

int check_dims(int n, int d) {
  if (n < 1) return SDIMB_EINVAL;
  if (d < 2 || d >= (1 << 15) || !is_prime(d)) return SDIMB_EDIM;
  if (d > 127 && n > 16384) return SDIMB_ETOOBIG;      // uint16 lanes: row indices and scratch of wide.cuh
  return SDIMB_OK;
}

int block_threads(int W) {
  int t = ((W / 4) + 31) / 32 * 32;
  if (t < 32) t = 32;
  if (const char* env = std::getenv("SDIMB_LANES_MIN_THREADS")) {   // developer knob (A/B timings)
    const int m = std::atoi(env);
    if (m > t && m <= kMaxThreads) t = m / 32 * 32;
  }
  if (t > kMaxThreads) t = t > kWideThreads ? kWideThreads : t;   // > 256 threads run interp_kernel_wide
  return t;
}

size_t scratch_bytes(int np) {
  const size_t W = 2 * (size_t)np;
  return 4 * W + W + 32 * 4 + 4 * 4 + 32 * 16 + 4 * (size_t)np + 2 * (W / 4) + 2 * (size_t)np + 128;
}

// rows of at most 8 lane words and a warp's tiles in shared memory: the tile interpreter's shapes.  The environment
// variable SDIMB_NO_TILE (developer knob, A/B timings) sends them back to one shot per warp.
bool tile_shape_ok(int n, int d) {
  static const bool off = std::getenv("SDIMB_NO_TILE") != nullptr;
  return !off && (d == 2 || d == 3) && n <= 128 && planes::tile_smem_bytes(n, d) <= (size_t)kSmemLimit;
}

// Which interpreter a (n, d, flags) call runs: 0 = uint8 lanes in global memory, 1 = uint8 lanes resident in
// shared memory, 2 = bit-plane resident (d = 2, 3), 3 = bit planes on a global image in caller scratch (d = 2, 3
// when the planes do not fit in shared memory, or FORCE_PLANES | FORCE_GLOBAL).  Negative = error code.
int plan_kernel(int n, int d, uint32_t flags, int np) {
  if ((flags & SDIMB_FORCE_GLOBAL) && (flags & SDIMB_FORCE_RESIDENT)) return SDIMB_EINVAL;
  if ((flags & SDIMB_FORCE_LANES) && (flags & SDIMB_FORCE_PLANES)) return SDIMB_EINVAL;
  SdimbLayout L;
  const int rc = sdimb_layout(n, d, &L);
  if (rc) return rc;
  if (d > 127)      // uint16 lanes (wide.cuh): one interpreter, on the HBM store
    return (flags & (SDIMB_FORCE_RESIDENT | SDIMB_FORCE_PLANES)) ? SDIMB_ETOOBIG : 4;
  const bool fits = (size_t)L.shot_bytes + scratch_bytes(np) <= (size_t)kSmemLimit;
  const bool planes_ok = d == 2 || d == 3;
  const bool planes_fit = planes_ok && planes::planes_smem_bytes(n, d) <= (size_t)kSmemLimit;
  const bool planes_global_ok = planes_ok && n <= 0xFFFF && planes::planes_scratch_bytes(n, SDIMB_SCHED_WARPS) <= (size_t)kSmemLimit;
  if ((flags & SDIMB_FORCE_PLANES) && (flags & SDIMB_FORCE_GLOBAL)) return planes_global_ok ? 3 : SDIMB_ETOOBIG;
  if ((flags & SDIMB_FORCE_PLANES) && !planes_fit) return SDIMB_ETOOBIG;
  if ((flags & SDIMB_FORCE_RESIDENT) && !fits) return SDIMB_ETOOBIG;
  const bool free_choice = !(flags & (SDIMB_FORCE_GLOBAL | SDIMB_FORCE_RESIDENT | SDIMB_FORCE_LANES | SDIMB_CLUSTER));
  // small tableaus (rows of at most 8 lane words, n <= 128): several shots per warp (planes_tile.cuh)
  if (planes_ok && tile_shape_ok(n, d) && !(flags & SDIMB_NO_TILE) &&
      (free_choice || ((flags & SDIMB_FORCE_PLANES) && !(flags & SDIMB_FORCE_GLOBAL))))
    return 5;
  if (free_choice && !(flags & SDIMB_FORCE_PLANES) && planes_ok) {
    // Shared memory holds few large images: at 3 or fewer resident CTAs per SM the same interpreter on an L2 image
    // with 8 CTAs per SM is faster (d = 3: n = 224 +19 %, 256 +22 %; d = 2: n = 320 +21 %, 400 +52 %); at 4 it is a
    // tie (d = 3, n = 200), above that shared memory wins (d = 3, n = 160: 5.0e9 against 3.7e9).
    const size_t per_cta = planes::planes_smem_bytes(n, d) + 1024;
    const int resident_ctas = planes_fit ? (int)((228u * 1024u) / per_cta) : 0;
    if (resident_ctas >= 4 || !planes_global_ok) { if (planes_fit) return 2; }
    else return 3;
  }
  if (planes_fit && free_choice) return 2;
  return (fits && !(flags & SDIMB_FORCE_GLOBAL)) ? 1 : 0;
}

// The tile interpreter (planes_tile.cuh): kernel for (d, lanes per shot).
using PlaneKernel = void (*)(const KParams);
// uni: every shot starts from |0...0> (SDIMB_FRESH): the tiles of a warp hold identical X / Z blocks, see t_measure
template <bool GLB>
PlaneKernel tile_kernel_glb(int n, int d, bool uni) {
  if (uni) {
    if (planes::tile_lps(n) == 4) return (d == 2) ? planes::interp_tile_kernel<2, 4, true, GLB> : planes::interp_tile_kernel<3, 4, true, GLB>;
    return (d == 2) ? planes::interp_tile_kernel<2, 8, true, GLB> : planes::interp_tile_kernel<3, 8, true, GLB>;
  }
  if (planes::tile_lps(n) == 4) return (d == 2) ? planes::interp_tile_kernel<2, 4, false, GLB> : planes::interp_tile_kernel<3, 4, false, GLB>;
  return (d == 2) ? planes::interp_tile_kernel<2, 8, false, GLB> : planes::interp_tile_kernel<3, 8, false, GLB>;
}
PlaneKernel tile_kernel_for(int n, int d, bool uni, bool glb = false) { return glb ? tile_kernel_glb<true>(n, d, uni) : tile_kernel_glb<false>(n, d, uni); }
// The tile interpreter keeps its images in caller scratch (L1 / L2) when the caller brought scratch for them: 20 warps per
// SM instead of the 8 that shared memory admits.  Same-box A/B (gpurun_out/t6_breakdown.txt, ms per 2e5 shots): config 3
// 34.4 -> 28.9, config 4 30.4 -> 20.7, config 2 (1e5 shots) 15.8 -> 14.5; 24 and 32 CTAs per SM (80 / 64 registers) are
// no better.  SDIMB_TILE_SMEM (developer knob) keeps the shared-memory images (tests run both).
bool tile_glb_wanted() { return std::getenv("SDIMB_TILE_SMEM") == nullptr; }
size_t tile_glb_scratch_bytes(int n, int d) {     // counter + one image per tile of every CTA a B200 keeps resident
  return 256 + (size_t)148 * SDIMB_TILE_GLB_CTAS * (32 / planes::tile_lps(n)) * 4 * planes::tile_glb_img_stride_words(n, d);
}

// The global-image plane interpreter for (d, interleaved image or not): see SDIMB_PG_IL_MIN_NP in planes.cuh.
// The environment variable of the same name moves the threshold (tests run the goldens through both images).
bool planes_interleaved(int n) {
  int min_np = SDIMB_PG_IL_MIN_NP;
  if (const char* env = std::getenv("SDIMB_PG_IL_MIN_NP")) min_np = std::atoi(env);
  return (n + 31) / 32 * 32 >= min_np;
}
PlaneKernel planes_gates_only_kernel(int d, bool il) {
  if (il) return (d == 2) ? planes::interp_planes_kernel<2, true, true, true> : planes::interp_planes_kernel<3, true, true, true>;
  return (d == 2) ? planes::interp_planes_kernel<2, true, false, true> : planes::interp_planes_kernel<3, true, false, true>;
}
PlaneKernel planes_global_kernel(int d, bool il) {
  if (il) return (d == 2) ? planes::interp_planes_kernel<2, true, true> : planes::interp_planes_kernel<3, true, true>;
  return (d == 2) ? planes::interp_planes_kernel<2, true, false> : planes::interp_planes_kernel<3, true, false>;
}

// CTAs of the bit-plane interpreter on a global image that the current device keeps resident (its grid never
// exceeds this), and the shared memory one of them needs.
int planes_global_ctas(int n, int d, size_t* smem_out) {
  const size_t smem = planes::planes_scratch_bytes(n, SDIMB_SCHED_WARPS);
  if (smem_out) *smem_out = smem;
  auto kern = planes_global_kernel(d, planes_interleaved(n));
  int dev = 0, sms = 0, per_sm = 0;
  // The image is read through L1: ask for no more shared memory per SM than the CTAs the register file admits need
  // (headline: 20 % instead of the driver's choice, +1 %; a full carve-out, as a co-resident shared-memory kernel
  // would force, costs 25 %).
  int carve = (int)((100 * (size_t)SDIMB_PLANES_GLOBAL_MIN_CTAS * (smem + 1024) + 228 * 1024 - 1) / (228 * 1024));
  if (carve > 100) carve = 100;
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * SDIMB_SCHED_WARPS, smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    return 0;
  }
  return sms * per_sm;
}

// Trailing measurement run on a generator-major image (planes_gm.cuh): kernel, resident CTAs, scratch layout.
template <int LPS, bool PRE>
PlaneKernel run_tail_kernel_lps(int d, bool il) {
  if (il) return (d == 2) ? planes::run_tail_kernel<2, true, LPS, PRE> : planes::run_tail_kernel<3, true, LPS, PRE>;
  return (d == 2) ? planes::run_tail_kernel<2, false, LPS, PRE> : planes::run_tail_kernel<3, false, LPS, PRE>;
}
template <bool PRE>
PlaneKernel run_tail_kernel_pre(int n, int d) {
  const bool il = planes_interleaved(n);
  switch (planes::run_lps(n)) {
    case 4: return run_tail_kernel_lps<4, PRE>(d, il);
    case 8: return run_tail_kernel_lps<8, PRE>(d, il);
    case 16: return run_tail_kernel_lps<16, PRE>(d, il);
    default: return run_tail_kernel_lps<32, PRE>(d, il);
  }
}
PlaneKernel run_tail_kernel_for(int n, int d, bool pre = false) { return pre ? run_tail_kernel_pre<true>(n, d) : run_tail_kernel_pre<false>(n, d); }
bool tail_run_shape_ok(int n, int d) { return (d == 2 || d == 3) && (n + 31) / 32 * 32 <= 512; }   // Wb <= 32
int run_tail_ctas(int n, int d) {
  auto kern = run_tail_kernel_for(n, d);
  const size_t smem = planes::run_smem_bytes(n);
  int dev = 0, sms = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, planes::kRunThreads, smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    return 0;
  }
  if (const char* env = std::getenv("SDIMB_RUN_CTAS_PER_SM")) {     // developer knob (A/B timings)
    const int m = std::atoi(env);
    if (m >= 1 && m < per_sm) per_sm = m;
  }
  return sms * per_sm;
}
// scratch of a call that hands its tail run to run_tail_kernel: 256 bytes of counters, one image per SHOT, one
// B + QX slab per resident tile (= shot in flight) of run_tail_kernel
size_t tail_run_scratch_bytes(int n, int d, int64_t shots, int run_ctas) {
  return 256 + (size_t)shots * 4 * planes::planes_img_stride_words(n, d) +
         (size_t)run_ctas * (planes::kRunThreads / planes::run_lps(n)) * 4 * planes::run_slab_words(n, d);
}
// ... and of one whose front kernel (gate_stream_kernel on a shared-memory image) writes B + QX per shot itself
size_t tail_run_fused_scratch_bytes(int n, int d, int64_t shots) {
  return 256 + (size_t)shots * 4 * planes::run_shot_stride_words(n, d);
}

// Gate-stream kernel (planes_stream.cuh) for the gates in front of a tail run: image in shared memory when one image,
// the accumulators of the widest launch and the largest fired-bit table fit (a rule of the shape alone, so that
// sdimb_gate_stream and sdimb_run agree), else on the per-shot global image.  SDIMB_GS_GLOBAL / SDIMB_GS_WARPS /
// SDIMB_NO_GATE_STREAM are developer knobs (A/B timings, tests of both forms).
bool gate_stream_shape_ok(int n, int d) {
  const bool off = std::getenv("SDIMB_NO_GATE_STREAM") != nullptr;
  return !off && (d == 2 || d == 3) && planes::gate_stream_gpw(n) >= 1;
}
size_t gate_stream_img_bytes(int n, int d) {
  const size_t EW = (d == 2) ? 2 : 4, np = (size_t)(n + 31) / 32 * 32, Wb = 2 * np / 32;
  return 4 * (((size_t)n * EW * (Wb + (planes_interleaved(n) ? 0 : 1)) + 3) & ~(size_t)3);
}
bool gate_stream_in_smem(int n, int d) {
  const bool force_global = std::getenv("SDIMB_GS_GLOBAL") != nullptr;
  return !force_global && gate_stream_img_bytes(n, d) + 8 * 32 * planes::kGateStreamMaxWarps + planes::kGateStreamMaxNoise / 8 + 16 <=
                              (size_t)kSmemLimit;
}
int gate_stream_warps(int n, int d) {
  if (!gate_stream_in_smem(n, d)) return SDIMB_SCHED_WARPS;     // the global-image form is compiled for 4-warp CTAs
  int nw = 8;
  if (const char* env = std::getenv("SDIMB_GS_WARPS")) nw = std::atoi(env);
  return nw < 1 ? 1 : nw > planes::kGateStreamMaxWarps ? planes::kGateStreamMaxWarps : nw;
}
PlaneKernel gate_stream_kernel_for(int n, int d) {
  const bool il = planes_interleaved(n);
  if (gate_stream_in_smem(n, d)) {
    if (il) return (d == 2) ? planes::gate_stream_kernel<2, true, true> : planes::gate_stream_kernel<3, true, true>;
    return (d == 2) ? planes::gate_stream_kernel<2, false, true> : planes::gate_stream_kernel<3, false, true>;
  }
  if (il) return (d == 2) ? planes::gate_stream_kernel<2, true, false> : planes::gate_stream_kernel<3, true, false>;
  return (d == 2) ? planes::gate_stream_kernel<2, false, false> : planes::gate_stream_kernel<3, false, false>;
}

// Trailing measurement run on uint8 lanes (lanes_gm.cuh): resident CTAs, scratch (counter + one B8 slab per resident warp)
// TMA staging of the row lists (lanes_gm.cuh): correct, but 12-25 % slower than ordinary loads + L2 prefetch on B200 (d = 5, 4096
// shots: n = 256 tail 6.25 vs 5.56 ms, n = 128 24.6 vs 19.7) — opt-in developer knob, both forms under test
bool tail8_tma() { return std::getenv("SDIMB_TAIL8_TMA") != nullptr; }
PlaneKernel run_tail8_kernel_for(int n) {
  if (tail8_tma())
    return n <= 128 ? lanesgm::run_tail8_kernel<1, true> : n <= 256 ? lanesgm::run_tail8_kernel<2, true> : lanesgm::run_tail8_kernel<4, true>;
  return n <= 128 ? lanesgm::run_tail8_kernel<1, false> : n <= 256 ? lanesgm::run_tail8_kernel<2, false> : lanesgm::run_tail8_kernel<4, false>;
}
int run_tail8_ctas(int n, int W) {
  const size_t smem = lanesgm::smem_bytes(n, W, tail8_tma());
  auto kern8 = run_tail8_kernel_for(n);
  int dev = 0, sms = 0, per_sm = 0;
  if (smem > (size_t)kSmemLimit || cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaFuncSetAttribute(kern8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern8, 32 * lanesgm::kWarps, smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    return 0;
  }
  return sms * per_sm;
}
size_t tail8_scratch_bytes(int n, int W, int ctas) { return 256 + (size_t)ctas * lanesgm::kWarps * lanesgm::slab_bytes(n, W); }
bool tail8_enabled() { return std::getenv("SDIMB_NO_TAIL8") == nullptr; }    // developer knob (A/B timings, tests)
// length of the run of plain M ops that ends an UNSCHEDULED host stream, 0 if shorter than the scheduler's run threshold
int64_t raw_tail_run(int n, const int32_t* ops, int64_t n_ops) {
  int64_t t = 0;
  while (t < n_ops && (ops[4 * (n_ops - 1 - t)] & SDIMB_OP_MASK) == SDIMB_OP_M) ++t;
  const int64_t min_run = n / 8 > 4 ? n / 8 : 4;
  return t >= min_run ? t : 0;
}

// Cluster interpreter (clusters.cuh) for a call that plan_kernel sends to the HBM store: cluster size, or 0 for
// one CTA per shot.  By default only tableaus with rows wider than one 256-thread CTA (n > 512), and the largest
// cluster size (16 = the non-portable maximum, then 8, 4, 2) of which the GPU can keep one per shot resident at the
// same time (cudaOccupancyMaxActiveClusters: a cluster lives on one GPC); more shots than that run one CTA each.
// SDIMB_CLUSTER forces it (tests; size from the environment variable SDIMB_CLUSTER_SIZE, default 8).
int cluster_threads(int wpc) {
  int t = 512;
  if (const char* env = std::getenv("SDIMB_CLUSTER_THREADS")) t = std::atoi(env);   // developer knob (A/B timings)
  t = t >= 1024 ? 1024 : t >= 512 ? 512 : 256;
  while (t < wpc) t *= 2;                          // a CTA needs one thread per lane word it owns
  return t;
}

// co-resident clusters of size C for this layout, 0 if the shape cannot run
int max_active_clusters(const SdimbLayout& L, int C) {
  const int wpc = clusters::cluster_wpc(L.np, C);
  const size_t smem = clusters::cluster_smem_bytes(L.np, wpc);
  if (wpc > clusters::kClusterThreads || smem > (size_t)kSmemLimit) return 0;
  auto kern = clusters::cluster_kernel_for(cluster_threads(wpc));
  int nc = 0;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3((unsigned)C); cfg.blockDim = dim3((unsigned)cluster_threads(wpc));
  cfg.dynamicSmemBytes = smem; cfg.attrs = attr; cfg.numAttrs = 1;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
      cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, C > 8 ? 1 : 0) != cudaSuccess ||
      cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return nc;
}

// SDIMB_TIME_KERNELS: events around the two kernels of a tail-run call, read back by sdimb_kernel_times (bench.py's
// per-kernel roofline).  One set per process: a measurement aid, not for concurrent callers.
cudaEvent_t g_time_ev[3] = {nullptr, nullptr, nullptr};
int g_time_valid = 0;
bool time_events_ready() {
  for (auto& e : g_time_ev)
    if (!e && cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return false; }
  return true;
}

int plan_cluster(const SdimbLayout& L, int64_t shots, uint32_t flags) {
  if (flags & SDIMB_NO_CLUSTER) return 0;
  if (flags & SDIMB_CLUSTER) {
    const char* env = std::getenv("SDIMB_CLUSTER_SIZE");
    int want = env ? std::atoi(env) : 8;
    if (want != 1 && want != 2 && want != 4 && want != 8 && want != 16) want = 8;
    for (int C = want; C >= 1; C >>= 1)
      if (max_active_clusters(L, C) >= 1) return C;
    return 0;
  }
  if (L.lanes / 4 <= kMaxThreads) return 0;
  for (int C = 16; C >= 2; C >>= 1)
    if (max_active_clusters(L, C) >= shots) return C;
  return 0;
}

}  // namespace

extern "C" {

int sdimb_version(void) { return SDIMB_VERSION; }

const char* sdimb_strerror(int code) {
  switch (code) {
    case SDIMB_OK: return "ok";
    case SDIMB_EINVAL: return "invalid argument";
    case SDIMB_EDIM: return "dimension must be a prime below 32768 (the frame sampler: at most 127)";
    case SDIMB_EOP: return "Invalid gate value";
    case SDIMB_ECUDA: return "CUDA error (is a GPU present? there is no CPU fallback)";
    case SDIMB_ETOOBIG: return "tableau does not fit in shared memory for the resident interpreter";
    default: return "unknown error";
  }
}

int sdimb_layout(int n, int d, SdimbLayout* out) {
  if (!out) return SDIMB_EINVAL;
  const int rc = check_dims(n, d);
  if (rc) return rc;
  const Arith A = make_arith(d);
  out->n = n;
  out->d = d;
  out->np = (n + 15) / 16 * 16;
  out->lanes = 2 * out->np;
  out->order = (int32_t)A.order;
  out->phase_order = (int32_t)A.po;
  out->elem_bytes = d > 127 ? 2 : 1;
  out->rec_bytes = out->elem_bytes;
  out->row_bytes = 2ll * out->lanes * out->elem_bytes;
  out->phase_offset = (int64_t)n * out->row_bytes;
  out->shot_bytes = out->phase_offset + (int64_t)out->lanes * out->elem_bytes;
  return SDIMB_OK;
}

int sdimb_init(void* tableau, int n, int d, int64_t shots, void* stream) {
  SdimbLayout L;
  const int rc = sdimb_layout(n, d, &L);
  if (rc) return rc;
  if (!tableau || shots < 0) return SDIMB_EINVAL;
  if (shots == 0) return SDIMB_OK;
  const int grid = (int)(shots < 148 * 16 ? shots : 148 * 16);
  if (L.elem_bytes == 2) {
    wide::init16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((uint16_t*)tableau, n, L.np, L.lanes, L.shot_bytes / 2, shots);
    g_launches++;
    return cudaGetLastError() == cudaSuccess ? SDIMB_OK : SDIMB_ECUDA;
  }
  init_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((uint8_t*)tableau, n, L.np, L.lanes, L.row_bytes, L.shot_bytes,
                                                      shots);
  g_launches++;
  return cudaGetLastError() == cudaSuccess ? SDIMB_OK : SDIMB_ECUDA;
}

int sdimb_run(const SdimbRunArgs* caller) {
  // older callers pass the struct without its last fields (tail_run_len; gate_stream, gate_stream_rows)
  if (!caller || (caller->struct_size != sizeof(SdimbRunArgs) && caller->struct_size != offsetof(SdimbRunArgs, tail_run_len) &&
                  caller->struct_size != offsetof(SdimbRunArgs, gate_stream)))
    return SDIMB_EINVAL;
  SdimbRunArgs args;
  std::memset(&args, 0, sizeof(args));
  std::memcpy(&args, caller, caller->struct_size);
  args.struct_size = sizeof(SdimbRunArgs);
  const SdimbRunArgs* a = &args;
  SdimbLayout L;
  const int rc = sdimb_layout(a->n, a->d, &L);
  if (rc) return rc;
  if (a->shots < 0 || a->n_ops < 0 || a->n_meas < 0 || a->n_noise < 0) return SDIMB_EINVAL;
  if (plan_kernel(a->n, a->d, a->flags, L.np) < 0) return plan_kernel(a->n, a->d, a->flags, L.np);
  if (a->shots == 0) return SDIMB_OK;
  if (a->n_ops > 0 && !a->ops) return SDIMB_EINVAL;
  if (a->n_meas > 0 && (!a->records || a->rec_stride < a->n_meas)) return SDIMB_EINVAL;
  if (a->n_noise > 0 && !a->replay_noise && (!a->noise_thresh24 || !a->noise_channel)) return SDIMB_EINVAL;

  const int kernel = plan_kernel(a->n, a->d, a->flags, L.np);
  if (kernel < 0) return kernel;
  const size_t scratch = scratch_bytes(L.np);
  const bool use_planes = kernel == 2, resident = kernel >= 1;   // 3 keeps its image in scratch, 5 in shared memory: no store needed either
  const bool need_tab = !resident || kernel == 4 || !(a->flags & SDIMB_FRESH) || (a->flags & SDIMB_WRITEBACK);
  if (need_tab && !a->tableau) return SDIMB_EINVAL;

  KParams p;
  std::memset(&p, 0, sizeof(p));
  p.tab = (uint8_t*)a->tableau;
  p.ops = (const int4*)a->ops;
  p.n_ops = a->n_ops;
  p.records = a->records;
  p.rec_stride = a->rec_stride;
  p.n_meas = a->n_meas;
  p.replay_meas = a->replay_meas;
  p.replay_noise = a->replay_noise;
  p.thresh = a->noise_thresh24;
  p.chan = a->noise_channel;
  p.n_noise = a->n_noise;
  p.seed = a->seed;
  p.shots = a->shots;
  p.shot_offset = a->shot_offset;
  p.n = L.n; p.np = L.np; p.W = L.lanes;
  p.row_bytes = L.row_bytes; p.phase_off = L.phase_offset; p.shot_bytes = L.shot_bytes;
  p.A = make_arith(a->d);
  p.flags = a->flags;
  p.resident = kernel == 1 ? 1 : 0;

  int dev = 0, sms = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return SDIMB_ECUDA;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return SDIMB_ECUDA;
  if (kernel == 5) {     // bit planes, several shots per warp: one-warp CTAs, 32 / LPS shots claimed at a time
    const bool no_uni = std::getenv("SDIMB_TILE_NO_UNI") != nullptr;            // developer knob (tests, A/B timings)
    const int tpw = 32 / planes::tile_lps(a->n);
    // images in caller scratch (global-image form) when asked for and the scratch holds one image per resident tile
    // ... and the call has at least three times the warps the shared-memory form keeps resident: below that the extra
    // warps do not exist and shared memory's latency wins (config 2 at its 10^4 shots: 2.10 vs 2.53 ms)
    bool glb = tile_glb_wanted() && a->scratch && a->scratch_bytes >= (int64_t)tile_glb_scratch_bytes(a->n, a->d);
    if (glb && !std::getenv("SDIMB_TILE_GLB")) {            // (SDIMB_TILE_GLB: developer knob, always scratch images)
      auto ksm = tile_kernel_for(a->n, a->d, (a->flags & SDIMB_FRESH) && !no_uni, false);
      const size_t ssm = planes::tile_smem_bytes(a->n, a->d);
      int psm = 0;
      if (cudaFuncSetAttribute(ksm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm) != cudaSuccess ||
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&psm, ksm, 32, ssm) != cudaSuccess || psm < 1) {
        cudaGetLastError();
        psm = 1;
      }
      glb = (a->shots + tpw - 1) / tpw >= 3ll * sms * psm;
    }
    auto kern = tile_kernel_for(a->n, a->d, (a->flags & SDIMB_FRESH) && !no_uni, glb);
    const size_t smem = glb ? planes::tile_glb_smem_bytes(a->n) : planes::tile_smem_bytes(a->n, a->d);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem) != cudaSuccess || per_sm < 1) {
      cudaGetLastError();
      return SDIMB_ECUDA;
    }
    int64_t grid = (int64_t)sms * per_sm;
    const int64_t groups = (a->shots + tpw - 1) / tpw;
    if (grid > groups) grid = groups;
    if (a->scratch && a->scratch_bytes >= (int64_t)sizeof(unsigned int)) {   // dynamic shot claiming
      p.shot_counter = (unsigned int*)a->scratch;
      if (cudaMemsetAsync(p.shot_counter, 0, sizeof(unsigned int), (cudaStream_t)a->stream) != cudaSuccess) return SDIMB_ECUDA;
    }
    p.tile_stride_words = (int64_t)planes::tile_stride_words(a->n, a->d);
    if (glb) {
      const int64_t cap = (a->scratch_bytes - 256) / (int64_t)(tpw * 4 * planes::tile_glb_img_stride_words(a->n, a->d));
      if (grid > cap) grid = cap;
      p.tile_stride_words = (int64_t)planes::tile_glb_lists_words(a->n);
      p.img_stride_words = (int64_t)planes::tile_glb_img_stride_words(a->n, a->d);
      p.plane_slab = (uint32_t*)((uint8_t*)a->scratch + 256);
    }
    kern<<<(unsigned)grid, 32, smem, (cudaStream_t)a->stream>>>(p);
    g_launches++;
    return cudaGetLastError() == cudaSuccess ? SDIMB_OK : SDIMB_ECUDA;
  }
  if (kernel == 4) {     // uint16 lanes, one CTA per shot on the HBM store
    const size_t smem = wide::smem_bytes(L.np);
    auto kern = wide::interp_wide16_kernel;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wide::kThreads, smem) != cudaSuccess || per_sm < 1) {
      cudaGetLastError();
      return SDIMB_ECUDA;
    }
    int64_t grid = (int64_t)sms * per_sm;
    if (grid > a->shots) grid = a->shots;
    kern<<<(unsigned)grid, wide::kThreads, smem, (cudaStream_t)a->stream>>>(p);
    g_launches++;
    return cudaGetLastError() == cudaSuccess ? SDIMB_OK : SDIMB_ECUDA;
  }
  // the global-image plane interpreter needs its slabs; a caller that brought none still gets the resident one
  // where that fits (it was the only plane interpreter of this ABI's first version)
  if (kernel == 3 && !(a->flags & SDIMB_FORCE_GLOBAL) && planes::planes_smem_bytes(a->n, a->d) <= (size_t)kSmemLimit &&
      (!a->scratch || a->scratch_bytes < (int64_t)(256 + planes::planes_row_bytes(a->n, a->d)))) {
    SdimbRunArgs b = *a;
    b.flags |= SDIMB_FORCE_PLANES;
    return sdimb_run(&b);
  }
  if (kernel == 3) {
    // bit planes on a global image: 4 warps per shot, one slab of scratch per resident CTA, shots claimed dynamically
    size_t smem = 0;
    const int max_ctas = planes_global_ctas(a->n, a->d, &smem);
    const size_t slab = planes::planes_row_bytes(a->n, a->d);
    if (max_ctas < 1) return SDIMB_ECUDA;
    if (!a->scratch || a->scratch_bytes < (int64_t)(256 + slab)) return SDIMB_EINVAL;
    int64_t grid = (a->scratch_bytes - 256) / (int64_t)slab;
    if (grid > max_ctas) grid = max_ctas;
    if (grid > a->shots) grid = a->shots;
    auto kern = planes_global_kernel(a->d, planes_interleaved(a->n));
    p.shot_counter = (unsigned int*)a->scratch;
    p.plane_slab = (uint32_t*)((uint8_t*)a->scratch + 256);
    p.pg = planes::make_plane_geo(a->n, a->d, SDIMB_SCHED_WARPS, true);
    // The run of M ops that ends the stream goes to run_tail_kernel (one warp per shot on a generator-major image)
    // when the caller says where it starts, does not want the tableau back and brought scratch for one image per shot.
    const int64_t tail = (a->flags & SDIMB_SCHEDULED) && !(a->flags & SDIMB_WRITEBACK) && tail_run_shape_ok(a->n, a->d) &&
                                 a->tail_run_len > 0 && a->tail_run_len <= a->n_ops ? a->tail_run_len : 0;
    const int run_ctas_max = tail ? run_tail_ctas(a->n, a->d) : 0;
    if (tail && run_ctas_max >= 1 && a->scratch_bytes >= (int64_t)tail_run_scratch_bytes(a->n, a->d, a->shots, run_ctas_max)) {
      if (cudaMemsetAsync(a->scratch, 0, 256, (cudaStream_t)a->stream) != cudaSuccess) return SDIMB_ECUDA;
      p.img_per_shot = 1;
      p.img_stride_words = (int64_t)planes::planes_img_stride_words(a->n, a->d);
      KParams p1 = p;
      p1.n_ops = a->n_ops - tail;
      int64_t grid1 = max_ctas < a->shots ? max_ctas : a->shots;
      // every measurement of the stream sits in the tail run: the front part runs the gates-only instantiation
      static const bool no_gates_only = std::getenv("SDIMB_NO_GATES_ONLY") != nullptr;     // developer knob (A/B timings)
      if (a->n_meas == tail && !no_gates_only) {
        kern = planes_gates_only_kernel(a->d, planes_interleaved(a->n));
        int per_sm = 0, sms = 0, dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * SDIMB_SCHED_WARPS, smem) != cudaSuccess || per_sm < 1) {
          cudaGetLastError();
          return SDIMB_ECUDA;
        }
        grid1 = (int64_t)sms * per_sm < a->shots ? (int64_t)sms * per_sm : a->shots;
      }
      // ... and as pre-decoded per-warp streams when the caller compiled them (sdimb_gate_stream)
      int threads1 = 32 * SDIMB_SCHED_WARPS;
      bool fused = false;
      if (a->n_meas == tail && a->gate_stream && a->gate_stream_rows > 0 && gate_stream_shape_ok(a->n, a->d)) {
        kern = gate_stream_kernel_for(a->n, a->d);
        const int nw = gate_stream_warps(a->n, a->d);
        threads1 = 32 * nw;
        smem = planes::gate_stream_smem_bytes(a->n, a->n_noise, nw) + (gate_stream_in_smem(a->n, a->d) ? gate_stream_img_bytes(a->n, a->d) : 0);
        int per_sm = 0, sms = 0, dev = 0;
        if (smem > (size_t)kSmemLimit || cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads1, smem) != cudaSuccess || per_sm < 1) {
          cudaGetLastError();
          return SDIMB_ECUDA;
        }
        grid1 = (int64_t)sms * per_sm < a->shots ? (int64_t)sms * per_sm : a->shots;
        p1.gate_stream = a->gate_stream;
        // image in shared memory: it leaves the SM transposed, one B + QX slab per shot (SDIMB_GS_NO_FUSE: A/B knob)
        if (gate_stream_in_smem(a->n, a->d) && !std::getenv("SDIMB_GS_NO_FUSE") &&
            a->scratch_bytes >= (int64_t)tail_run_fused_scratch_bytes(a->n, a->d, a->shots)) {
          fused = true;
          p1.gm_per_shot = 1;
          p1.gm_slab = p.plane_slab;
          p1.gm_slab_words = (int64_t)planes::run_slab_words(a->n, a->d);
          p1.gm_shot_stride_words = (int64_t)planes::run_shot_stride_words(a->n, a->d);
        }
      }
      const bool timed = (a->flags & SDIMB_TIME_KERNELS) && time_events_ready();
      g_time_valid = 0;
      if (timed) cudaEventRecord(g_time_ev[0], (cudaStream_t)a->stream);
      kern<<<(unsigned)grid1, threads1, smem, (cudaStream_t)a->stream>>>(p1);
      g_launches++;
      if (cudaGetLastError() != cudaSuccess) return SDIMB_ECUDA;
      if (timed) cudaEventRecord(g_time_ev[1], (cudaStream_t)a->stream);
      KParams p2 = p;
      p2.tail_start = a->n_ops - tail;
      p2.shot_counter = (unsigned int*)a->scratch + 1;
      p2.gm_slab = p.plane_slab + (int64_t)a->shots * p.img_stride_words;
      p2.gm_slab_words = (int64_t)planes::run_slab_words(a->n, a->d);
      if (fused) { p2.gm_per_shot = 1; p2.gm_slab = p1.gm_slab; p2.gm_shot_stride_words = p1.gm_shot_stride_words; }
      const int tiles = planes::kRunThreads / planes::run_lps(a->n);
      int64_t grid2 = (a->shots + tiles - 1) / tiles;
      if (grid2 > run_ctas_max) grid2 = run_ctas_max;
      auto kern2 = run_tail_kernel_for(a->n, a->d, fused);
      if (fused && cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)planes::run_smem_bytes(a->n)) != cudaSuccess) {
        cudaGetLastError();
        return SDIMB_ECUDA;
      }
      kern2<<<(unsigned)grid2, planes::kRunThreads, planes::run_smem_bytes(a->n), (cudaStream_t)a->stream>>>(p2);
      g_launches++;
      if (timed) { cudaEventRecord(g_time_ev[2], (cudaStream_t)a->stream); g_time_valid = 1; }
      return cudaGetLastError() == cudaSuccess ? SDIMB_OK : SDIMB_ECUDA;
    }
    if (cudaMemsetAsync(p.shot_counter, 0, sizeof(unsigned int), (cudaStream_t)a->stream) != cudaSuccess) return SDIMB_ECUDA;
    kern<<<(unsigned)grid, 32 * SDIMB_SCHED_WARPS, smem, (cudaStream_t)a->stream>>>(p);
    g_launches++;
    return cudaGetLastError() == cudaSuccess ? SDIMB_OK : SDIMB_ECUDA;
  }
  if (use_planes) {
    auto kern = (a->d == 2) ? planes::interp_planes_kernel<2, false> : planes::interp_planes_kernel<3, false>;
    // one warp per shot unless shared memory leaves the SM short of warps and the stream is scheduled
    int nw = 1;
    size_t smem = planes::planes_smem_bytes(a->n, a->d, 1);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit) != cudaSuccess)
      return SDIMB_ECUDA;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem) != cudaSuccess || per_sm < 1)
      return SDIMB_ECUDA;
    if ((a->flags & SDIMB_SCHEDULED) && per_sm < 12) {
      nw = SDIMB_SCHED_WARPS;
      smem = planes::planes_smem_bytes(a->n, a->d, nw);
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * nw, smem) != cudaSuccess || per_sm < 1)
        return SDIMB_ECUDA;
    }
    int64_t grid = (int64_t)sms * per_sm;
    if (grid > a->shots) grid = a->shots;
    if (a->scratch && a->scratch_bytes >= (int64_t)sizeof(unsigned int)) {   // dynamic shot claiming
      p.shot_counter = (unsigned int*)a->scratch;
      if (cudaMemsetAsync(p.shot_counter, 0, sizeof(unsigned int), (cudaStream_t)a->stream) != cudaSuccess)
        return SDIMB_ECUDA;
    }
    p.pg = planes::make_plane_geo(a->n, a->d, nw, false);
    kern<<<(unsigned)grid, 32 * nw, smem, (cudaStream_t)a->stream>>>(p);
    g_launches++;
    return cudaGetLastError() == cudaSuccess ? SDIMB_OK : SDIMB_ECUDA;
  }
  if (kernel == 0) {
    const int C = plan_cluster(L, a->shots, a->flags);
    if (C >= 1) {                                    // one shot per thread-block cluster
      p.wpc = clusters::cluster_wpc(L.np, C);
      const int cthreads = cluster_threads(p.wpc);
      auto kern = clusters::cluster_kernel_for(cthreads);
      const size_t smem = clusters::cluster_smem_bytes(L.np, p.wpc);
      cudaLaunchConfig_t cfg = {};
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.gridDim = dim3((unsigned)C); cfg.blockDim = dim3((unsigned)cthreads);
      cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)a->stream; cfg.attrs = attr; cfg.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) != cudaSuccess || nc < 1) { cudaGetLastError(); return SDIMB_ECUDA; }
      const int64_t n_clusters = a->shots < nc ? a->shots : nc;
      cfg.gridDim = dim3((unsigned)(n_clusters * C));
      const cudaError_t err = cudaLaunchKernelEx(&cfg, kern, p);
      g_launches++;
      if (err != cudaSuccess) { cudaGetLastError(); return SDIMB_ECUDA; }
      return SDIMB_OK;
    }
  }
  // HBM store, fresh shots, tableau not kept, the caller says where the trailing run of M ops starts and brought
  // scratch for the slabs: that run goes to run_tail8_kernel (one warp per shot on a generator-major copy, lanes_gm.cuh)
  int64_t tail8 = 0;
  int tail8_ctas = 0;
  if (kernel == 0 && (a->flags & SDIMB_FRESH) && !(a->flags & (SDIMB_WRITEBACK | SDIMB_SCHEDULED)) && tail8_enabled() &&
      lanesgm::shape_ok(a->n, a->d) && a->tail_run_len > 0 && a->tail_run_len <= a->n_ops && a->scratch) {
    tail8_ctas = run_tail8_ctas(a->n, L.lanes);
    if (tail8_ctas >= 1 && a->scratch_bytes >= (int64_t)tail8_scratch_bytes(a->n, L.lanes, tail8_ctas)) tail8 = a->tail_run_len;
  }
  if (tail8) p.n_ops = a->n_ops - tail8;
  const int64_t front_meas = a->n_meas - tail8;
  // wide rows streamed from the HBM store: 16 lanes per thread (fewer, fatter threads; 128-bit accesses)
  // ... when the stream is gate-dominated: measurements want many threads per shot, gates want few fat ones
  p.vec = (kernel == 0 && L.lanes / 4 >= 128 && L.lanes / 16 <= kMaxThreads && front_meas * 64 <= p.n_ops) ? 4 : 1;
  const int threads = block_threads(L.lanes / p.vec);
  const size_t smem = scratch + (resident ? (size_t)L.shot_bytes : 0);
  auto lane_kernel = p.vec == 4 ? interp_kernel_stream : threads > kMaxThreads ? interp_kernel_wide : interp_kernel;
  if (cudaFuncSetAttribute(lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return SDIMB_ECUDA;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lane_kernel, threads, smem) != cudaSuccess || per_sm < 1)
    return SDIMB_ECUDA;
  int64_t grid = (int64_t)sms * per_sm;
  if (grid > a->shots) grid = a->shots;
  const bool timed8 = tail8 && (a->flags & SDIMB_TIME_KERNELS) && time_events_ready();
  if (tail8) g_time_valid = 0;
  if (timed8) cudaEventRecord(g_time_ev[0], (cudaStream_t)a->stream);
  lane_kernel<<<(unsigned)grid, threads, smem, (cudaStream_t)a->stream>>>(p);
  g_launches++;
  if (cudaGetLastError() != cudaSuccess) return SDIMB_ECUDA;
  if (tail8) {
    if (timed8) cudaEventRecord(g_time_ev[1], (cudaStream_t)a->stream);
    if (cudaMemsetAsync(a->scratch, 0, 256, (cudaStream_t)a->stream) != cudaSuccess) return SDIMB_ECUDA;
    KParams p2 = p;
    p2.n_ops = a->n_ops;
    p2.tail_start = a->n_ops - tail8;
    p2.shot_counter = (unsigned int*)a->scratch;
    p2.gm_slab = (uint32_t*)((uint8_t*)a->scratch + 256);
    p2.gm_slab_words = (int64_t)(lanesgm::slab_bytes(a->n, L.lanes) / 4);
    int64_t grid2 = (a->shots + lanesgm::kWarps - 1) / lanesgm::kWarps;
    if (grid2 > tail8_ctas) grid2 = tail8_ctas;
    run_tail8_kernel_for(a->n)<<<(unsigned)grid2, 32 * lanesgm::kWarps, lanesgm::smem_bytes(a->n, L.lanes, tail8_tma()), (cudaStream_t)a->stream>>>(p2);
    g_launches++;
    if (timed8) { cudaEventRecord(g_time_ev[2], (cudaStream_t)a->stream); g_time_valid = 1; }
  }
  return cudaGetLastError() == cudaSuccess ? SDIMB_OK : SDIMB_ECUDA;
}

int sdimb_export(const void* tableau, int n, int d, int64_t shot, int64_t* x, int64_t* z, int64_t* p, int64_t* dx,
                 int64_t* dz, int64_t* dp, void* stream) {
  SdimbLayout L;
  const int rc = sdimb_layout(n, d, &L);
  if (rc) return rc;
  if (!tableau || shot < 0 || !x || !z || !p || !dx || !dz || !dp) return SDIMB_EINVAL;
  const uint8_t* T = (const uint8_t*)tableau + shot * L.shot_bytes;
  const int64_t total = (int64_t)n * n;
  const int grid = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  if (L.elem_bytes == 2) {
    wide::export16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)T, n, L.np, L.lanes, x, z, p, dx, dz, dp);
    g_launches++;
    return cudaGetLastError() == cudaSuccess ? SDIMB_OK : SDIMB_ECUDA;
  }
  export_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(T, n, L.np, L.lanes, L.row_bytes, L.phase_offset, x, z, p, dx,
                                                        dz, dp);
  g_launches++;
  return cudaGetLastError() == cudaSuccess ? SDIMB_OK : SDIMB_ECUDA;
}

// Workspace of the host-buffer entry, ONE PER DEVICE: a grow-only device arena, a grow-only pinned staging buffer, a
// stream and two events, created on first use and reused by later calls (cudaMalloc/cudaFree per call cost more than
// the simulation of a small batch).  Calls on the same device are serialised by that workspace's mutex; host threads
// driving different GPUs do not touch each other's workspace.  sdimb_release_workspace() frees all of them.
namespace {
struct HostPlan {          // see sdimb_simulate_host
  bool valid = false;
  int n = 0, d = 0, kernel = 0;
  int64_t shots = 0, n_meas = 0, up_n = 0, tail_len = 0, gs_rows = 0;
  uint32_t flags = 0, mode_flags = 0, sched_flag = 0;
  uint64_t knobs = 0;
  bool sched_used = false;
  std::vector<int32_t> ops, sched, gstream;
};
struct HostWorkspace {
  std::mutex mu;
  int device = -1;
  void* dev = nullptr;
  size_t dev_cap = 0;
  void* pin = nullptr;
  size_t pin_cap = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  HostPlan plan;
  void release() {
    if (dev) cudaFree(dev);
    if (pin) cudaFreeHost(pin);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (stream) cudaStreamDestroy(stream);
    dev = pin = nullptr; dev_cap = pin_cap = 0; stream = nullptr; e0 = e1 = nullptr; device = -1;
    plan = HostPlan();
  }
  bool prepare(int cur, size_t dev_bytes, size_t pin_bytes) {   // caller holds mu and has `cur` as current device
    device = cur;
    if (!stream && cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) return false;
    if (!e0 && cudaEventCreate(&e0) != cudaSuccess) return false;
    if (!e1 && cudaEventCreate(&e1) != cudaSuccess) return false;
    if (dev_bytes > dev_cap) {
      if (dev) cudaFree(dev);
      dev = nullptr; dev_cap = 0;
      if (cudaMalloc(&dev, dev_bytes) != cudaSuccess) return false;
      dev_cap = dev_bytes;
    }
    if (pin_bytes > pin_cap) {
      if (pin) cudaFreeHost(pin);
      pin = nullptr; pin_cap = 0;
      if (cudaHostAlloc(&pin, pin_bytes, cudaHostAllocDefault) != cudaSuccess) return false;
      pin_cap = pin_bytes;
    }
    return true;
  }
};
constexpr int kMaxDevices = 64;
HostWorkspace g_ws_by_device[kMaxDevices];
inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// everything a developer knob can change in a host plan (tests flip them between calls)
uint64_t host_plan_knobs(int n, int d) {
  uint64_t k = 0;
  auto mix = [&](uint64_t v) { k = (k ^ v) * 0x100000001B3ull; };
  mix(gate_stream_shape_ok(n, d)); mix(gate_stream_in_smem(n, d)); mix((uint64_t)gate_stream_warps(n, d));
  mix(planes_interleaved(n)); mix(tile_shape_ok(n, d)); mix(tail8_enabled()); mix(tail8_tma()); mix(tile_glb_wanted());
  for (const char* name : {"SDIMB_GM_MIN_RUN", "SDIMB_CLUSTER_SIZE", "SDIMB_CLUSTER_THREADS"}) {
    const char* v = std::getenv(name);
    mix(v ? (uint64_t)std::atoll(v) + 1 : 0);
  }
  return k;
}

int make_host_plan(int n, int d, int64_t shots, uint32_t flags, const int32_t* ops, int64_t n_ops, int64_t n_meas,
                   const SdimbLayout& L, HostPlan& hp) {
  int rc;
  uint32_t mode_flags = flags & (SDIMB_FORCE_GLOBAL | SDIMB_FORCE_RESIDENT | SDIMB_FORCE_LANES |
                                 SDIMB_FORCE_PLANES | SDIMB_CLUSTER | SDIMB_NO_CLUSTER | SDIMB_NO_TILE);
  int kernel = plan_kernel(n, d, mode_flags, L.np);
  if (kernel < 0) return kernel;
  // A few shots of a d = 2, 3 tableau too large for shared memory: one uint8 tableau per thread-block cluster beats
  // one CTA per shot on bit planes (n = 2048, 8 shots: 12.8 ms against 30 ms), as long as every shot gets a cluster
  if (kernel == 3 && !(mode_flags & SDIMB_FORCE_PLANES) && planes::planes_smem_bytes(n, d) > (size_t)kSmemLimit &&
      plan_cluster(L, shots, mode_flags) > 0) {
    mode_flags |= SDIMB_FORCE_LANES;
    kernel = plan_kernel(n, d, mode_flags, L.np);
    if (kernel < 0) return kernel;
  }
  // uint8 lanes that fit in shared memory: from 96 qudits on, the HBM store with the trailing run of M ops in the
  // generator-major kernel is faster (TableauEngine._auto_mode has the numbers)
  if (kernel == 1 && mode_flags == 0 && n >= 96 && lanesgm::shape_ok(n, d) && tail8_enabled() && raw_tail_run(n, ops, n_ops) > 0 &&
      plan_kernel(n, d, SDIMB_FORCE_GLOBAL, L.np) == 0 && plan_cluster(L, shots, SDIMB_FORCE_GLOBAL) == 0) {
    mode_flags = SDIMB_FORCE_GLOBAL;
    kernel = 0;
  }
  std::vector<int32_t>& sched = hp.sched;
  const int32_t* up_ops = ops;
  int64_t up_n = n_ops;
  uint32_t sched_flag = 0;
  // bit-plane interpreter and cluster interpreter (wide rows on the HBM store): upload the layered stream
  const bool maybe_cluster = kernel == 0 && !(flags & SDIMB_NO_CLUSTER) && (L.lanes / 4 > kMaxThreads || (flags & SDIMB_CLUSTER));
  // (a tile-interpreter call above 64 qudits is scheduled too: it may turn into a two-kernel call, see below)
  const bool tile_to_two = kernel == 5 && n > 64 && mode_flags == 0 && n_meas > 0;
  if ((kernel == 2 || kernel == 3 || maybe_cluster || tile_to_two) && n_ops > 0) {
    sched.resize((size_t)(2 * n_ops + 1) * 4);
    rc = sdimb_schedule(n, ops, n_ops, sched.data(), 2 * n_ops + 1, &up_n);
    if (rc) return rc;
    up_ops = sched.data();
    sched_flag = SDIMB_SCHEDULED;
  }
  // Every measurement in the trailing run: the two-kernel path (gate_stream_kernel + run_tail_kernel) beats the
  // shared-memory interpreters where those fit, and the tile interpreter above 64 qudits (TableauEngine._auto_mode)
  if ((kernel == 2 || tile_to_two) && mode_flags == 0 && sched_flag) {
    const int64_t t = sdimb_tail_run(up_ops, up_n);
    if (t > 0 && t == n_meas && gate_stream_shape_ok(n, d) && tail_run_shape_ok(n, d) &&
        plan_kernel(n, d, SDIMB_FORCE_PLANES | SDIMB_FORCE_GLOBAL, L.np) == 3) {
      mode_flags = SDIMB_FORCE_PLANES | SDIMB_FORCE_GLOBAL;
      kernel = 3;
    } else if (tile_to_two) {         // stays with the tiles, on the stream as the caller wrote it
      up_ops = ops; up_n = n_ops; sched_flag = 0;
    }
  }

  int64_t tail_len = sched_flag ? sdimb_tail_run(up_ops, up_n) : 0;
  // uint8 lanes on the HBM store, one CTA per shot: the trailing run of M ops of the stream as the caller wrote it
  if (kernel == 0 && !sched_flag && lanesgm::shape_ok(n, d) && tail8_enabled() && plan_cluster(L, shots, mode_flags) == 0)
    tail_len = raw_tail_run(n, ops, n_ops);
  // the gates in front of a tail run that holds every measurement: pre-decoded per-warp streams (planes_stream.cuh)
  std::vector<int32_t>& gstream = hp.gstream;
  int64_t gs_rows = 0;
  if (kernel == 3 && tail_len > 0 && tail_len == n_meas && gate_stream_shape_ok(n, d) &&
      sdimb_gate_stream(n, d, up_ops, up_n - tail_len, nullptr, 0, &gs_rows) == SDIMB_OK && gs_rows > 0) {
    gstream.resize((size_t)gs_rows * 4);
    if (sdimb_gate_stream(n, d, up_ops, up_n - tail_len, gstream.data(), gs_rows, &gs_rows) != SDIMB_OK) gs_rows = 0;
  } else {
    gs_rows = 0;
  }
  hp.mode_flags = mode_flags; hp.sched_flag = sched_flag; hp.kernel = kernel;
  hp.sched_used = up_ops != ops;
  hp.up_n = up_n; hp.tail_len = tail_len; hp.gs_rows = gs_rows;
  return SDIMB_OK;
}
}  // namespace

int sdimb_release_workspace(void) {
  int cur = 0;
  const bool have = cudaGetDevice(&cur) == cudaSuccess;
  for (int dv = 0; dv < kMaxDevices; ++dv) {
    HostWorkspace& ws = g_ws_by_device[dv];
    std::lock_guard<std::mutex> lock(ws.mu);
    if (ws.device < 0) continue;
    cudaSetDevice(ws.device);
    ws.release();
  }
  if (have) cudaSetDevice(cur);
  cudaGetLastError();
  return SDIMB_OK;
}

int sdimb_simulate_host(int n, int d, int64_t shots, int64_t shot_offset, const int32_t* ops, int64_t n_ops,
                        uint8_t* records, int64_t n_meas, const uint8_t* replay_meas, const uint8_t* replay_noise,
                        const uint32_t* noise_thresh24, const uint8_t* noise_channel, int64_t n_noise, uint64_t seed,
                        uint32_t flags, float* elapsed_ms) {
  SdimbLayout L;
  int rc = sdimb_layout(n, d, &L);
  if (rc) return rc;
  if (shots < 0 || n_ops < 0 || n_meas < 0 || n_noise < 0) return SDIMB_EINVAL;
  if ((n_ops > 0 && !ops) || (n_meas > 0 && shots > 0 && !records)) return SDIMB_EINVAL;
  for (int64_t i = 0; i < n_ops; ++i) {   // host-side validation: "Invalid gate value" (sdim/program.py:381-382)
    const int32_t* o = ops + 4 * i;
    if (o[0] < 0 || o[0] > SDIMB_OP_N1 || o[1] < 0 || o[1] >= n) return SDIMB_EOP;
    const bool two = o[0] >= SDIMB_OP_CNOT && o[0] <= SDIMB_OP_SWAP;
    if (two && (o[2] < 0 || o[2] >= n || o[2] == o[1])) return SDIMB_EOP;
    const bool meas = o[0] >= SDIMB_OP_M && o[0] <= SDIMB_OP_RESET;
    if (meas && (o[3] < 0 || o[3] >= n_meas)) return SDIMB_EOP;
    if (o[0] == SDIMB_OP_N1 && (o[3] < 0 || o[3] >= n_noise)) return SDIMB_EOP;
  }
  if (n_noise > 0 && !replay_noise && (!noise_thresh24 || !noise_channel)) return SDIMB_EINVAL;
  if (shots == 0) return SDIMB_OK;

  // ---- host-side plan of the call: interpreter choice, scheduled stream, tail run, compiled gate streams.  A pure
  // function of (n, d, shots, flags, op stream, developer knobs); the workspace keeps the plan of its last program, so
  // a caller that simulates one program batch after batch (Program.simulate's waves, bench.py's steps) pays the
  // scheduler and the gate-stream compiler once.  The op stream is compared byte for byte.
  // SDIMB_HOST_PROFILE (developer knob): wall-clock microseconds of the stages of this call on stderr
  const bool prof = std::getenv("SDIMB_HOST_PROFILE") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return (double)std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count() * 1e-3;
  };
  const auto t_in = now();
  int cur_dev = 0;
  if (cudaGetDevice(&cur_dev) != cudaSuccess || cur_dev < 0 || cur_dev >= kMaxDevices) { cudaGetLastError(); return SDIMB_ECUDA; }
  HostWorkspace& g_ws = g_ws_by_device[cur_dev];
  std::lock_guard<std::mutex> lock(g_ws.mu);
  HostPlan& hp = g_ws.plan;
  const uint64_t knobs = host_plan_knobs(n, d);
  const bool hit = hp.valid && hp.n == n && hp.d == d && hp.shots == shots && hp.flags == flags && hp.n_meas == n_meas &&
                   hp.knobs == knobs && (int64_t)hp.ops.size() == 4 * n_ops &&
                   (n_ops == 0 || std::memcmp(hp.ops.data(), ops, (size_t)n_ops * 16) == 0);
  if (!hit) {
    hp.valid = false;
    rc = make_host_plan(n, d, shots, flags, ops, n_ops, n_meas, L, hp);
    if (rc) return rc;
    hp.n = n; hp.d = d; hp.shots = shots; hp.flags = flags; hp.n_meas = n_meas; hp.knobs = knobs;
    hp.ops.assign(ops, ops + 4 * n_ops);
    hp.valid = true;
  }
  const uint32_t mode_flags = hp.mode_flags, sched_flag = hp.sched_flag;
  const int kernel = hp.kernel;
  const int32_t* up_ops = hp.sched_used ? hp.sched.data() : ops;
  const int64_t up_n = hp.up_n, tail_len = hp.tail_len, gs_rows = hp.gs_rows;
  const std::vector<int32_t>& gstream = hp.gstream;

  // carve the device arena
  const size_t eb = (size_t)L.rec_bytes;     // records / replayed outcomes / replayed exponents: 1 byte, 2 for d > 127
  const size_t b_ops = align256((size_t)up_n * 16), b_rec = align256((size_t)shots * n_meas * eb);
  const size_t b_rm = replay_meas ? b_rec : 0;
  const size_t b_rn = replay_noise ? align256((size_t)shots * n_noise * 2 * eb) : 0;
  const size_t b_th = (n_noise && noise_thresh24) ? align256((size_t)n_noise * 4) : 0;
  const size_t b_ch = (n_noise && noise_channel) ? align256((size_t)n_noise) : 0;
  const size_t b_tab = (kernel == 0 || kernel == 4) ? align256((size_t)shots * L.shot_bytes) : 0;
  const size_t b_gs = align256((size_t)gs_rows * 16);
  const size_t b_scr = align256((size_t)sdimb_scratch_bytes_shots(n, d, mode_flags, tail_len ? shots : 0));
  const size_t total = b_ops + b_rec + b_rm + b_rn + b_th + b_ch + b_tab + b_gs + b_scr + 256;

  const auto t_plan = now();
  // `records` in pinned (page-locked or registered) host memory: the device writes it directly; pageable memory goes
  // through the workspace's pinned staging buffer and one host copy
  bool direct_out = false;
  {
    cudaPointerAttributes attr;
    if (records && cudaPointerGetAttributes(&attr, records) == cudaSuccess) direct_out = attr.type == cudaMemoryTypeHost;
    else cudaGetLastError();
  }
  if (!g_ws.prepare(cur_dev, total, direct_out ? 0 : b_rec + 256)) { cudaGetLastError(); return SDIMB_ECUDA; }
  cudaStream_t st = g_ws.stream;
  auto t_h2d = t_plan, t_run = t_plan, t_sync = t_plan;
  uint8_t* base = (uint8_t*)g_ws.dev;
  uint8_t* d_ops = base; base += b_ops;
  uint8_t* d_rec = base; base += b_rec;
  uint8_t* d_rm = b_rm ? base : nullptr; base += b_rm;
  uint8_t* d_rn = b_rn ? base : nullptr; base += b_rn;
  uint8_t* d_th = b_th ? base : nullptr; base += b_th;
  uint8_t* d_ch = b_ch ? base : nullptr; base += b_ch;
  uint8_t* d_tab = b_tab ? base : nullptr; base += b_tab;
  uint8_t* d_gs = b_gs ? base : nullptr; base += b_gs;
  uint8_t* d_scr = b_scr ? base : nullptr;
  rc = SDIMB_ECUDA;
  do {
    if (cudaEventRecord(g_ws.e0, st) != cudaSuccess) break;
    if (up_n && cudaMemcpyAsync(d_ops, up_ops, (size_t)up_n * 16, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
    if (d_rm && cudaMemcpyAsync(d_rm, replay_meas, (size_t)shots * n_meas * eb, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
    if (d_rn && cudaMemcpyAsync(d_rn, replay_noise, (size_t)shots * n_noise * 2 * eb, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
    if (d_th && cudaMemcpyAsync(d_th, noise_thresh24, (size_t)n_noise * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
    if (d_ch && cudaMemcpyAsync(d_ch, noise_channel, (size_t)n_noise, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
    if (d_gs && cudaMemcpyAsync(d_gs, gstream.data(), (size_t)gs_rows * 16, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
    t_h2d = now();
    SdimbRunArgs a;
    std::memset(&a, 0, sizeof(a));
    a.struct_size = sizeof(a);
    a.flags = mode_flags | SDIMB_FRESH | sched_flag;
    a.n = n; a.d = d; a.shots = shots; a.shot_offset = shot_offset;
    a.tableau = d_tab;
    a.ops = (const int32_t*)d_ops; a.n_ops = up_n;
    a.records = d_rec; a.n_meas = n_meas; a.rec_stride = n_meas;
    a.replay_meas = d_rm; a.replay_noise = d_rn;
    a.noise_thresh24 = (const uint32_t*)d_th; a.noise_channel = d_ch; a.n_noise = n_noise;
    a.seed = seed; a.stream = st;
    a.scratch = d_scr; a.scratch_bytes = (int64_t)b_scr;
    a.tail_run_len = tail_len;
    a.gate_stream = (const int32_t*)d_gs; a.gate_stream_rows = gs_rows;
    rc = sdimb_run(&a);
    if (rc) break;
    t_run = now();
    rc = SDIMB_ECUDA;
    const size_t rec_bytes = (size_t)shots * n_meas * eb;
    if (rec_bytes && cudaMemcpyAsync(direct_out ? (void*)records : g_ws.pin, d_rec, rec_bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
    if (cudaEventRecord(g_ws.e1, st) != cudaSuccess) break;
    if (cudaStreamSynchronize(st) != cudaSuccess) break;
    t_sync = now();
    if (rec_bytes && !direct_out) std::memcpy(records, g_ws.pin, rec_bytes);
    if (elapsed_ms && cudaEventElapsedTime(elapsed_ms, g_ws.e0, g_ws.e1) != cudaSuccess) break;
    rc = SDIMB_OK;
    if (prof)
      std::fprintf(stderr, "sdimb_simulate_host: plan %.0f us, arena %.0f, h2d enqueue %.0f, sdimb_run %.0f, wait %.0f, copy out %.0f\n",
                   us(t_in, t_plan), us(t_plan, t_plan), us(t_plan, t_h2d), us(t_h2d, t_run), us(t_run, t_sync), us(t_sync, now()));
  } while (0);
  if (rc == SDIMB_ECUDA) cudaGetLastError();
  return rc;
}

int64_t sdimb_launch_count(void) { return g_launches.load(); }

int sdimb_kernel_times(float* front_ms, float* tail_ms) {
  if (!front_ms || !tail_ms) return SDIMB_EINVAL;
  if (!g_time_valid) return SDIMB_EINVAL;
  if (cudaEventSynchronize(g_time_ev[2]) != cudaSuccess || cudaEventElapsedTime(front_ms, g_time_ev[0], g_time_ev[1]) != cudaSuccess ||
      cudaEventElapsedTime(tail_ms, g_time_ev[1], g_time_ev[2]) != cudaSuccess) {
    cudaGetLastError();
    return SDIMB_ECUDA;
  }
  return SDIMB_OK;
}

#ifdef SDIMB_PHASE_CLOCKS
// developer build only: read (and clear) the cluster interpreter's per-phase clock accumulators
int sdimb_debug_phase_clocks(unsigned long long* out32) {
  unsigned long long zero[32] = {0};
  if (cudaDeviceSynchronize() != cudaSuccess) return SDIMB_ECUDA;
  if (cudaMemcpyFromSymbol(out32, clusters::g_phase_clk, sizeof(zero)) != cudaSuccess) return SDIMB_ECUDA;
  if (cudaMemcpyToSymbol(clusters::g_phase_clk, zero, sizeof(zero)) != cudaSuccess) return SDIMB_ECUDA;
  return SDIMB_OK;
}
#endif

int sdimb_cluster_size(int n, int d, int64_t shots, uint32_t flags) {
  SdimbLayout L;
  if (sdimb_layout(n, d, &L) || shots < 1) return 0;
  if (plan_kernel(n, d, flags, L.np) != 0) return 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return plan_cluster(L, shots, flags);
}

int sdimb_schedule(int n, const int32_t* ops, int64_t n_ops, int32_t* out, int64_t out_cap, int64_t* out_n) {
  if (n < 1 || n_ops < 0 || (n_ops > 0 && !ops) || !out || !out_n || out_cap < 2 * n_ops + 1) return SDIMB_EINVAL;
  // lw[q]: first layer in which row q may be read again (last writer + 1); lr[q]: first layer in which row q
  // may be written again (last reader or writer + 1).  Layers restart after every collective op.
  std::vector<int> lw(n, 0), lr(n, 0);
  std::vector<std::vector<int64_t>> layers;
  int64_t w = 0;
  // Runs of consecutive M ops (nothing but I / BARRIER between them) of at least min_run measurements are marked
  // SDIMB_GM_*: the global-image bit-plane interpreter runs them on a generator-major copy of its image, which costs
  // two transpositions (~n^2 / 8 bits each way) and saves a column walk over n rows per measurement.
  int64_t min_run = n / 8 > 4 ? n / 8 : 4;
  if (const char* env = std::getenv("SDIMB_GM_MIN_RUN")) min_run = std::atoll(env) > 2 ? std::atoll(env) : 2;
  std::vector<uint8_t> gm_mark(n_ops, 0);
  {
    int64_t run_first = -1, run_last = -1, run_len = 0;
    auto close_run = [&](bool follow) {
      if (run_len >= min_run) {
        for (int64_t i = run_first; i <= run_last; ++i)
          if ((ops[4 * i] & SDIMB_OP_MASK) == SDIMB_OP_M) gm_mark[i] = SDIMB_GM_IN;
        gm_mark[run_first] |= SDIMB_GM_FIRST;
        gm_mark[run_last] |= SDIMB_GM_LAST | (follow ? SDIMB_GM_FOLLOW : 0);
      }
      run_len = 0;
    };
    for (int64_t i = 0; i < n_ops; ++i) {
      const int op = ops[4 * i] & SDIMB_OP_MASK;
      if (op == SDIMB_OP_I || op == SDIMB_OP_BARRIER) continue;
      if (op == SDIMB_OP_M) {
        if (run_len == 0) run_first = i;
        run_last = i;
        ++run_len;
      } else {
        close_run(true);
      }
    }
    close_run(false);
  }
  auto flush = [&]() {
    for (auto& layer : layers) {
      if (layer.empty()) continue;
      // ops of a layer commute: order them by opcode, so that a warp (which takes every SDIMB_SCHED_WARPS-th op)
      // runs gates of one kind back to back and stays in the same stretch of the interpreter's code
      std::stable_sort(layer.begin(), layer.end(), [&](int64_t x, int64_t y) {
        return (ops[4 * x] & SDIMB_OP_MASK) < (ops[4 * y] & SDIMB_OP_MASK);
      });
      int k = 0;
      for (int64_t i : layer) {
        const int32_t* o = ops + 4 * i;
        int32_t* r = out + 4 * w++;
        r[0] = (o[0] & SDIMB_OP_MASK) | ((k % SDIMB_SCHED_WARPS) << SDIMB_OP_WARP_SHIFT) |
               ((k & 0x7FFF) << SDIMB_OP_INDEX_SHIFT);
        ++k;
        r[1] = o[1]; r[2] = o[2]; r[3] = o[3];
      }
      int32_t* b = out + 4 * w++;
      b[0] = SDIMB_OP_BARRIER; b[1] = 0; b[2] = -1; b[3] = -1;
    }
    layers.clear();
    std::fill(lw.begin(), lw.end(), 0);
    std::fill(lr.begin(), lr.end(), 0);
  };
  for (int64_t i = 0; i < n_ops; ++i) {
    const int32_t* o = ops + 4 * i;
    const int op = o[0] & SDIMB_OP_MASK, a = o[1], b = o[2];
    if (op < 0 || op > SDIMB_OP_BARRIER || a < 0 || a >= n) return SDIMB_EOP;
    if (op == SDIMB_OP_I || op == SDIMB_OP_BARRIER) continue;
    if (op >= SDIMB_OP_M && op <= SDIMB_OP_RESET) {     // collective: everything before it completes first
      flush();
      int32_t* r = out + 4 * w++;
      r[0] = op | ((int32_t)gm_mark[i] << SDIMB_OP_WARP_SHIFT); r[1] = a; r[2] = -1; r[3] = o[3];
      continue;
    }
    const bool two = op >= SDIMB_OP_CNOT && op <= SDIMB_OP_SWAP;
    if (two && (b < 0 || b >= n || b == a)) return SDIMB_EOP;
    const bool reads_only = (op >= SDIMB_OP_X && op <= SDIMB_OP_Z_INV) || op == SDIMB_OP_N1;
    int level;
    if (reads_only) {
      level = lw[a];
      lr[a] = std::max(lr[a], level + 1);
    } else {
      level = std::max(lw[a], lr[a]);
      if (two) level = std::max(level, std::max(lw[b], lr[b]));
      lw[a] = lr[a] = level + 1;
      if (two) lw[b] = lr[b] = level + 1;
    }
    if ((size_t)level >= layers.size()) layers.resize(level + 1);
    layers[level].push_back(i);
  }
  flush();
  *out_n = w;
  return SDIMB_OK;
}

int sdimb_gate_stream(int n, int d, const int32_t* ops, int64_t n_ops, int32_t* out, int64_t out_cap_rows, int64_t* out_rows) {
  if (n < 1 || (d != 2 && d != 3) || n_ops < 0 || (n_ops > 0 && !ops) || !out_rows) return SDIMB_EINVAL;
  const int gpw = planes::gate_stream_gpw(n);
  if (gpw < 1) return SDIMB_EINVAL;
  const int nw = gate_stream_warps(n, d);
  const bool il = planes_interleaved(n);
  const int32_t rs_bytes = (int32_t)(4 * ((d == 2) ? 2 : 4) * (2 * ((n + 31) / 32 * 32) / 32 + (il ? 0 : 1)));   // Geo<D, IL>::RS words
  struct Row { int32_t x, y, z, w; };
  std::vector<std::vector<Row>> ws(nw);          // per-warp streams: entries of gpw rows
  std::vector<Row> tab;                          // N1 events (slot, qudit, noisy-layer ordinal), layer by layer
  int64_t max_slot = -1;
  auto family = [&](int op, int32_t& flags) -> int {
    switch (op) {
      case SDIMB_OP_H: return planes::GS_H;
      case SDIMB_OP_H_INV: flags = GS_INV; return planes::GS_H;
      case SDIMB_OP_P: return planes::GS_P;
      case SDIMB_OP_P_INV: flags = GS_INV; return planes::GS_P;
      case SDIMB_OP_CNOT: return planes::GS_CNOT;
      case SDIMB_OP_CNOT_INV: flags = GS_INV; return planes::GS_CNOT;
      case SDIMB_OP_CZ: return planes::GS_CZ;
      case SDIMB_OP_CZ_INV: flags = GS_INV; return planes::GS_CZ;
      case SDIMB_OP_SWAP: return planes::GS_SWAP;
      default: return -1;
    }
  };
  // Layers of THIS kernel: ASAP over the gates that write rows.  The input order is one valid sequential order of the
  // stretch (sdimb_schedule's, or the program's own); its layering is not reused, because there every N1 and Pauli
  // reader sits in a layer of its own kind between two writers of its row (gate, N1, gate ... costs two layers per
  // gate).  Here Pauli gates leave the stream (below) and an N1 event belongs to the START of the layer of the next
  // writer of its row: the kernel applies the fired events of a layer before any of its gates runs.
  struct Layer { std::vector<Row> fam[8]; std::vector<Row> noise; };
  std::vector<Layer> L;
  std::vector<int64_t> lvl(n, 0);                // first layer in which row q may be written again
  for (int64_t i = 0; i < n_ops; ++i) {
    const int32_t* o = ops + 4 * i;
    const int op = o[0] & SDIMB_OP_MASK;
    if (op == SDIMB_OP_BARRIER || op == SDIMB_OP_I) continue;
    if (o[1] < 0 || o[1] >= n) return SDIMB_EOP;
    if (op >= SDIMB_OP_X && op <= SDIMB_OP_Z_INV) continue;         // Pauli gates: pushed back to the start, below
    if (op == SDIMB_OP_N1) {
      if (o[3] < 0) return SDIMB_EINVAL;
      if ((int64_t)L.size() <= lvl[o[1]]) L.resize(lvl[o[1]] + 1);
      L[lvl[o[1]]].noise.push_back(Row{o[3], o[1], 0, 0});
      if (o[3] > max_slot) max_slot = o[3];
      continue;
    }
    int32_t flags = 0;
    const int fam = family(op, flags);
    if (fam < 0) return SDIMB_EINVAL;                                // a measurement: not a gate-only stretch
    const bool two = fam >= planes::GS_CNOT;
    if (two && (o[2] < 0 || o[2] >= n || o[2] == o[1])) return SDIMB_EOP;
    const int64_t level = two ? std::max(lvl[o[1]], lvl[o[2]]) : lvl[o[1]];
    lvl[o[1]] = level + 1;
    if (two) lvl[o[2]] = level + 1;
    if ((int64_t)L.size() <= level) L.resize(level + 1);
    L[level].fam[fam].push_back(Row{fam | flags | GS_ON, o[1] * rs_bytes, two ? o[2] * rs_bytes : 0, 0});
  }
  const int64_t layers = (int64_t)L.size();
  int32_t noisy = 0;                             // layers with N1 events so far: index of the layer's "some event fired" bit
  for (auto& layer : L) {
    const int32_t tab_lo = (int32_t)tab.size();
    for (Row r : layer.noise) { r.z = noisy; tab.push_back(r); }
    const int32_t sync_x = planes::GS_SYNC | (noisy << GS_LAYER_SHIFT);
    if (!layer.noise.empty()) ++noisy;
    for (int w = 0; w < nw; ++w)
      for (int g = 0; g < gpw; ++g) ws[w].push_back(Row{sync_x, tab_lo, (int32_t)tab.size(), 0});
    int64_t k = 0;
    for (int fam = planes::GS_H; fam <= planes::GS_SWAP; ++fam) {
      const auto& v = layer.fam[fam];
      for (size_t c = 0; c < v.size(); c += gpw) {
        auto& dst = ws[k++ % nw];
        for (int g = 0; g < gpw; ++g) dst.push_back(c + g < v.size() ? v[c + g] : Row{fam, 0, 0, 0});
      }
    }
  }
  if (noisy >= (1 << (31 - GS_LAYER_SHIFT))) return SDIMB_EINVAL;
  if ((int64_t)tab.size() > planes::kGateStreamMaxNoise || (int64_t)tab.size() > max_slot + 1) return SDIMB_EINVAL;
  // Pauli gates, pushed BACK through the gates in front of them (P then U == U then U P U^-1, so walking backwards the
  // pending Pauli takes the rule of the INVERSE gate on its exponents; the scheduled order is one valid sequential
  // order of the stretch) and merged into one X^a Z^b per qudit at the start.  Global phases do not reach a tableau.
  std::vector<int> pa(n, 0), pb(n, 0);
  for (int64_t i = n_ops - 1; i >= 0; --i) {
    const int32_t* o = ops + 4 * i;
    const int q = o[1], t = o[2];
    switch (o[0] & SDIMB_OP_MASK) {
      case SDIMB_OP_X: pa[q] = (pa[q] + 1) % d; break;
      case SDIMB_OP_X_INV: pa[q] = (pa[q] + d - 1) % d; break;
      case SDIMB_OP_Z: pb[q] = (pb[q] + 1) % d; break;
      case SDIMB_OP_Z_INV: pb[q] = (pb[q] + d - 1) % d; break;
      case SDIMB_OP_H: { const int a = pa[q]; pa[q] = pb[q]; pb[q] = (d - a) % d; break; }          // H^-1: (x,z) <- (z,-x)
      case SDIMB_OP_H_INV: { const int a = pa[q]; pa[q] = (d - pb[q]) % d; pb[q] = a; break; }      // H: (x,z) <- (-z,x)
      case SDIMB_OP_P: pb[q] = (pb[q] + d - pa[q]) % d; break;                                      // P^-1: z -= x
      case SDIMB_OP_P_INV: pb[q] = (pb[q] + pa[q]) % d; break;
      case SDIMB_OP_CNOT: pa[t] = (pa[t] + d - pa[q]) % d; pb[q] = (pb[q] + pb[t]) % d; break;      // CNOT^-1: x[t] -= x[c], z[c] += z[t]
      case SDIMB_OP_CNOT_INV: pa[t] = (pa[t] + pa[q]) % d; pb[q] = (pb[q] + d - pb[t]) % d; break;
      case SDIMB_OP_CZ: { const int bq = (pb[q] + d - pa[t]) % d, bt = (pb[t] + d - pa[q]) % d; pb[q] = bq; pb[t] = bt; break; }   // CZ^-1: z[a] -= x[b], z[b] -= x[a]
      case SDIMB_OP_CZ_INV: { const int bq = (pb[q] + pa[t]) % d, bt = (pb[t] + pa[q]) % d; pb[q] = bq; pb[t] = bt; break; }
      case SDIMB_OP_SWAP: std::swap(pa[q], pa[t]); std::swap(pb[q], pb[t]); break;
      default: break;
    }
  }
  std::vector<Row> pauli;
  for (int q = 0; q < n; ++q)
    if (pa[q] || pb[q]) pauli.push_back(Row{q, pa[q], pb[q], 0});
  for (int w = 0; w < nw; ++w)
    for (int g = 0; g < planes::kGateStreamPadRows; ++g) ws[w].push_back(Row{planes::GS_END, 0, 0, 0});
  int64_t total = planes::kGateStreamHeaderRows + nw;
  for (int w = 0; w < nw; ++w) total += (int64_t)ws[w].size();
  const int64_t tab_base = total;
  total += (int64_t)tab.size() + (int64_t)pauli.size();
  *out_rows = total;
  if (!out) return SDIMB_OK;
  if (out_cap_rows < total || total > 0x7FFFFFFF) return SDIMB_EINVAL;
  Row* R = reinterpret_cast<Row*>(out);
  R[0] = Row{planes::kGateStreamMagic, gpw, nw, (int32_t)tab.size()};
  R[1] = Row{(int32_t)tab_base, (int32_t)total, (int32_t)layers, (int32_t)pauli.size()};
  R[2] = Row{rs_bytes, il ? 1 : 0, 0, 0};
  int64_t at = planes::kGateStreamHeaderRows + nw;
  for (int w = 0; w < nw; ++w) {
    // entries up to and including the first END
    R[planes::kGateStreamHeaderRows + w] = Row{(int32_t)at, (int32_t)((ws[w].size() - planes::kGateStreamPadRows) / gpw + 1), 0, 0};
    std::memcpy(R + at, ws[w].data(), ws[w].size() * sizeof(Row));
    at += (int64_t)ws[w].size();
  }
  if (!tab.empty()) std::memcpy(R + at, tab.data(), tab.size() * sizeof(Row));
  at += (int64_t)tab.size();
  if (!pauli.empty()) std::memcpy(R + at, pauli.data(), pauli.size() * sizeof(Row));
  return SDIMB_OK;
}

int sdimb_frames(int n, int d, int64_t shots, int64_t shot_offset, const int32_t* ops, int64_t n_ops,
                 const uint8_t* reference, uint8_t* records, int64_t n_meas, int64_t rec_stride, uint8_t* frames,
                 const uint8_t* replay_z0, const uint8_t* replay_zm, const uint8_t* replay_noise,
                 const uint32_t* noise_thresh24, const uint8_t* noise_channel, int64_t n_noise, uint64_t seed,
                 void* stream) {
  const int rc = check_dims(n, d);
  if (rc) return rc;
  if (d > 127) return SDIMB_EDIM;       // frames and packed record bytes are uint8 here
  if (shots < 0 || n_ops < 0 || n_meas < 0 || n_noise < 0) return SDIMB_EINVAL;
  if (shots == 0) return SDIMB_OK;
  if ((n_ops > 0 && !ops) || !frames) return SDIMB_EINVAL;
  if (n_meas > 0 && (!reference || !records || rec_stride < n_meas)) return SDIMB_EINVAL;
  if (n_noise > 0 && !replay_noise && (!noise_thresh24 || !noise_channel)) return SDIMB_EINVAL;
  FParams p;
  std::memset(&p, 0, sizeof(p));
  p.ops = (const int4*)ops; p.n_ops = n_ops;
  p.reference = reference; p.records = records; p.n_meas = n_meas; p.rec_stride = rec_stride;
  p.frames = frames; p.pitch = (shots + 127) / 128 * 128;
  p.replay_z0 = replay_z0; p.replay_zm = replay_zm; p.replay_noise = replay_noise;
  p.thresh = noise_thresh24; p.chan = noise_channel; p.n_noise = n_noise;
  p.seed = seed; p.shots = shots; p.shot_offset = shot_offset; p.n = n;
  p.A = make_arith(d);
  const int64_t grid = (shots + 127) / 128;
  frame_kernel<<<(unsigned)grid, 128, 0, (cudaStream_t)stream>>>(p);
  g_launches++;
  return cudaGetLastError() == cudaSuccess ? SDIMB_OK : SDIMB_ECUDA;
}

int64_t sdimb_scratch_bytes(int n, int d, uint32_t flags) {
  SdimbLayout L;
  if (sdimb_layout(n, d, &L)) return 0;
  const int k = plan_kernel(n, d, flags, L.np);
  if (k == 5 && tile_glb_wanted()) return (int64_t)tile_glb_scratch_bytes(n, d);
  if (k == 2 || k == 5) return 256;   // the shot counter
  if (k == 3) {             // the shot counter + one image per CTA the device keeps resident (148 x 8 without a device)
    int ctas = planes_global_ctas(n, d, nullptr);
    if (ctas < 1) ctas = 148 * 8;
    return 256 + (int64_t)ctas * (int64_t)planes::planes_row_bytes(n, d);
  }
  return 0;
}

int64_t sdimb_scratch_bytes_shots(int n, int d, uint32_t flags, int64_t shots) {
  const int64_t base = sdimb_scratch_bytes(n, d, flags);
  SdimbLayout L;
  if (shots > 0 && !sdimb_layout(n, d, &L) && plan_kernel(n, d, flags, L.np) == 0 && lanesgm::shape_ok(n, d) && tail8_enabled()) {
    int c8 = run_tail8_ctas(n, L.lanes);                 // uint8 lanes on the HBM store: one B8 slab per resident warp
    if (c8 < 1) c8 = 148 * 8;
    return (int64_t)tail8_scratch_bytes(n, L.lanes, c8);
  }
  if (shots <= 0 || sdimb_layout(n, d, &L) || plan_kernel(n, d, flags, L.np) != 3 || !tail_run_shape_ok(n, d)) return base;
  int ctas = run_tail_ctas(n, d);
  if (ctas < 1) ctas = 148 * planes::kRunCtasPerSm;     // no device: the size a B200 would need
  int64_t need = (int64_t)tail_run_scratch_bytes(n, d, shots, ctas);
  if (gate_stream_shape_ok(n, d) && gate_stream_in_smem(n, d) && (int64_t)tail_run_fused_scratch_bytes(n, d, shots) > need)
    need = (int64_t)tail_run_fused_scratch_bytes(n, d, shots);
  return need > base ? need : base;
}

int64_t sdimb_tail_run(const int32_t* ops, int64_t n_ops) {
  if (!ops || n_ops < 1) return 0;
  auto mark = [&](int64_t i) { return (ops[4 * i] >> SDIMB_OP_WARP_SHIFT) & 0xFF; };
  auto code = [&](int64_t i) { return ops[4 * i] & SDIMB_OP_MASK; };
  const int64_t last = n_ops - 1;
  if (code(last) != SDIMB_OP_M || !(mark(last) & SDIMB_GM_LAST) || (mark(last) & SDIMB_GM_FOLLOW)) return 0;
  for (int64_t i = last; i >= 0; --i) {
    if (code(i) != SDIMB_OP_M || !(mark(i) & SDIMB_GM_IN)) return 0;
    if (mark(i) & SDIMB_GM_FIRST) return n_ops - i;
  }
  return 0;
}

int sdimb_plan(int n, int d, uint32_t flags, int* kernel, int* needs_tableau) {
  SdimbLayout L;
  const int rc = sdimb_layout(n, d, &L);
  if (rc) return rc;
  const int k = plan_kernel(n, d, flags, L.np);
  if (k < 0) return k;
  if (kernel) *kernel = k;
  if (needs_tableau) *needs_tableau = (k == 0 || k == 4 || !(flags & SDIMB_FRESH) || (flags & SDIMB_WRITEBACK)) ? 1 : 0;   // k = 3 keeps its image in scratch
  return SDIMB_OK;
}

}  // extern "C"
