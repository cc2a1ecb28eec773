// Window solve with TMA-staged stage tiles (sm_100a).  Included by dekf_api.cu (product) and tools/tune_solve.cu
// (kernel tuning harness).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <type_traits>

#include "estimator_core.cuh"

namespace dekf {

// ------------------------------------------------------------------------------------------------
// window solve with TMA-staged stage tiles
// ------------------------------------------------------------------------------------------------
// One CTA = 128 consecutive instances.  The window ring is a 2-D tensor [NW*25 rows][ns instances]; the record of
// stage k for the CTA's instances is the box {128 instances, 25 rows} at (i0, slot*25).  One elected thread fetches
// each stage with ONE cp.async.bulk.tensor.2d (SASS: UTMALDG) into a kStages-deep shared-memory ring; completion
// is signalled on a "full" mbarrier (expect_tx), consumers hand the buffer back through an "empty" mbarrier.
// The stage record is read from shared memory at the point of use instead of being parked in registers.
constexpr int kTile = 128;
// Ring depth, CTAs per SM and pipeline granularity per element type (measured, profiles/r02_tune_solve.md).  fp64: the sweep
// needs ~250 registers, i.e. 2 CTAs/SM; 3 stage tiles of 25.6 KB each per CTA (2: 95 us, 3: 89 us, 4: 91 us per launch).
// fp32: 128 registers and 4 CTAs/SM without spills.  One pipeline per CTA or one per warp makes no difference (89.2 / 89.6 us).
// (the DEKF_SOLVE_* macros exist for the diagnosis builds of DESIGN.md section 10.1; the product build defines none of them)
#ifndef DEKF_SOLVE_STAGES
#define DEKF_SOLVE_STAGES 3
#endif
#ifndef DEKF_SOLVE_PW
#define DEKF_SOLVE_PW false
#endif
template <typename T>
struct SolveCfg {
  static constexpr int kStages = DEKF_SOLVE_STAGES, kMinB = 1;
  static constexpr bool kXS = false;
  static constexpr bool kPerWarp = DEKF_SOLVE_PW;
};
template <>
struct SolveCfg<float> {
  static constexpr int kStages = 3, kMinB = 4;
  static constexpr bool kXS = false;
  static constexpr bool kPerWarp = false;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}

template <typename T, int TILE, bool XS>
__device__ __forceinline__ typename std::conditional<XS, MemVec9<T, TILE>, Vec9<T>>::type make_x(T *col) {
  if constexpr (XS)
    return MemVec9<T, TILE>(col);
  else
    return Vec9<T>();
}

template <typename T, int TILE = kTile, int STAGES = SolveCfg<T>::kStages>
struct SmemStageSource {
  T *tiles;         // [STAGES][REC_SIZE][TILE]
  uint64_t *full;   // [STAGES]
  uint64_t *empty;  // [STAGES]
  const CUtensorMap *map;
  int NW, i0, tid, k0, nst;

  __device__ __forceinline__ void issue(int j) const {  // elected thread only
    const int buf = j % STAGES;
    const int slot = (k0 + j) % NW;
    mbar_expect_tx(&full[buf], (uint32_t)(REC_SIZE * TILE * sizeof(T)));
    tma_load_2d(tiles + (size_t)buf * REC_SIZE * TILE, map, i0, slot * REC_SIZE, &full[buf]);
  }
  __device__ __forceinline__ void acquire(int j) const { mbar_wait(&full[j % STAGES], (uint32_t)((j / STAGES) & 1)); }
  __device__ __forceinline__ void release(int j) const {
#if defined(DEKF_SOLVE_SYNC_EVERY_STAGE)
    __syncthreads();  // diagnosis: every thread of the CTA has finished reading the stage before anybody goes on
#endif
#if !defined(DEKF_SOLVE_NO_PROXY_FENCE)
    // This thread's LDS of the stage record (generic proxy) must be performed before the refill of the buffer -- a TMA write, i.e. the
    // ASYNC proxy, issued by ANOTHER warp once all arrivals are in.  mbarrier.arrive alone does not order the two proxies: with the
    // load / store queues of the SM backed up by a co-resident kernel, a warp's loads were still pending when the refill landed
    // (a transient wrong stage record for one warp, only under ragged VO arrival).
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
    mbar_arrive(&empty[j % STAGES]);
    // the elected thread refills the buffer of the PREVIOUS stage (everybody has long released it) with the
    // stage STAGES-1 ahead of the current one
    if (tid == 0 && j >= 1 && (j - 1 + STAGES) < nst) {
      mbar_wait(&empty[(j - 1) % STAGES], (uint32_t)(((j - 1) / STAGES) & 1));
      issue(j - 1 + STAGES);
    }
  }
  __device__ __forceinline__ const T *row(int j, int f) const {
    return tiles + ((size_t)(j % STAGES) * REC_SIZE + f) * TILE + tid;
  }
  __device__ __forceinline__ void meas(int j, int, S3<T> &Lam, V3<T> &eta) const {
#pragma unroll
    for (int f = 0; f < 6; ++f) Lam.a[f] = *row(j, REC_LAM + f);
#pragma unroll
    for (int f = 0; f < 3; ++f) eta[f] = *row(j, REC_ETA + f);
  }
  __device__ __forceinline__ void rot(int j, int, M3<T> &R) const {
#pragma unroll
    for (int f = 0; f < 9; ++f) R.a[f] = *row(j, REC_R + f);
  }
  __device__ __forceinline__ void dyn(int j, int, V3<T> &as, V3<T> &dlt, bool &vo) const {
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      as[f] = *row(j, REC_AS + f);
      dlt[f] = *row(j, REC_DLT + f);
    }
    vo = *row(j, REC_FLAG) != T(0);
  }
};

template <typename T, int TILE = kTile, int STAGES = SolveCfg<T>::kStages, bool XS = SolveCfg<T>::kXS>
constexpr size_t solve_tma_smem_bytes() {
  // stage ring + full/empty barriers of up to TILE/32 pipelines (+ x when it lives in shared memory)
  return (size_t)STAGES * REC_SIZE * TILE * sizeof(T) + (size_t)(TILE / 32) * 2 * STAGES * sizeof(uint64_t) + (XS ? (size_t)9 * TILE * sizeof(T) : 0);
}

// XS: keep the state mean x (9 scalars per instance) in shared memory instead of registers ([9][TILE] behind the ring).
// PW: one pipeline PER WARP (TMA box {32 instances, 25 rows}, the warp's own full / empty barriers, lane 0 refills): no warp
// ever waits for another warp of its CTA -- with PW = false the whole CTA shares one pipeline fed by thread 0.
template <typename T, int TILE = kTile, int STAGES = SolveCfg<T>::kStages, int MINB = SolveCfg<T>::kMinB, typename Math = DefaultMath<T>,
          bool XS = SolveCfg<T>::kXS, bool PW = SolveCfg<T>::kPerWarp>
__global__ void __launch_bounds__(TILE, MINB) k_solve_tma(const __grid_constant__ CUtensorMap tmap, const MheConst<T> c, const Dims dm,
                                                     const Buffers<T> b, const Inputs in, const Outputs out, int Tk,
                                                     int32_t *status_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int PT = PW ? 32 : TILE;           // instances per pipeline
  constexpr int NP = TILE / PT;                // pipelines per CTA
  const int pipe = PW ? (int)(threadIdx.x >> 5) : 0;
  const int ptid = PW ? (int)(threadIdx.x & 31) : (int)threadIdx.x;
  SmemStageSource<T, PT, STAGES> src;
  src.tiles = reinterpret_cast<T *>(smem_raw) + (size_t)pipe * STAGES * REC_SIZE * PT;
  src.full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * REC_SIZE * TILE * sizeof(T)) + (size_t)pipe * 2 * STAGES;
  src.empty = src.full + STAGES;
  src.map = &tmap;
  src.NW = dm.NW;
  src.i0 = (blockIdx.x + dm.tile0) * TILE + pipe * PT;
  src.tid = ptid;
  src.k0 = (Tk < dm.N) ? 0 : Tk - dm.N;
  src.nst = Tk - src.k0 + 1;
  const int active = min(PT, dm.n - src.i0);  // consumers of this pipeline (<= 0: the pipeline has no live instance)
  // the arrival-cost loads are in flight while the barriers are set up and the first tiles are requested
  const int i = src.i0 + ptid;
  Cov9<T> P;
  using XT = typename std::conditional<XS, MemVec9<T, TILE>, Vec9<T>>::type;
  XT x = make_x<T, TILE, XS>(reinterpret_cast<T *>(smem_raw + (size_t)STAGES * REC_SIZE * TILE * sizeof(T) + (size_t)NP * 2 * STAGES * sizeof(uint64_t)) + threadIdx.x);
  if (i < dm.n) mhe_solve_start(c, dm, b, Tk, i, P, x);
  if (ptid == 0 && active > 0) {
    if (threadIdx.x == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&src.full[s], 1);
      mbar_init(&src.empty[s], (uint32_t)active);
    }
    mbar_fence_init();
  }
  if constexpr (PW)
    __syncwarp();
  else
    __syncthreads();
  if (ptid == 0 && active > 0) {
    const int pre = src.nst < STAGES ? src.nst : STAGES;
    for (int j = 0; j < pre; ++j) src.issue(j);
  }
  if (i >= dm.n) return;
  int st = tick_status(dm, b, Tk, i);
  st |= mhe_solve_sweep<T, SmemStageSource<T, PT, STAGES>, Math, XT>(c, dm, b, in, out, Tk, i, src, P, x, src.k0);
  tick_status(dm, b, Tk, i) = st;
  if (status_out != nullptr) status_out[i] = st;
}

// Incremental window solve on a tick that carries VO messages: the CTA agrees on the earliest restart stage of its 128
// instances (re-sweeping from an earlier valid checkpoint gives the same bits) and streams the stage records
// ks .. T through the same TMA ring as k_solve_tma.
template <typename T, int TILE = kTile, int STAGES = SolveCfg<T>::kStages, typename Math = DefaultMath<T>, bool PW = SolveCfg<T>::kPerWarp>
__global__ void __launch_bounds__(TILE, 1) k_solve_incr_tma(const __grid_constant__ CUtensorMap tmap, const MheConst<T> c,
                                                            const Dims dm, const Buffers<T> b, const Inputs in,
                                                            const Outputs out, int Tk, int32_t *status_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ int s_ks;
  constexpr int PT = PW ? 32 : TILE;
  const int pipe = PW ? (int)(threadIdx.x >> 5) : 0;
  const int ptid = PW ? (int)(threadIdx.x & 31) : (int)threadIdx.x;
  SmemStageSource<T, PT, STAGES> src;
  src.tiles = reinterpret_cast<T *>(smem_raw) + (size_t)pipe * STAGES * REC_SIZE * PT;
  src.full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * REC_SIZE * TILE * sizeof(T)) + (size_t)pipe * 2 * STAGES;
  src.empty = src.full + STAGES;
  src.map = &tmap;
  src.NW = dm.NW;
  src.i0 = blockIdx.x * TILE + pipe * PT;
  src.tid = ptid;
  const int i = src.i0 + ptid;
  const int active = min(PT, dm.n - src.i0);
  if (ptid == 0 && active > 0) {
    if (threadIdx.x == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&src.full[s], 1);
      mbar_init(&src.empty[s], (uint32_t)active);
    }
    mbar_fence_init();
  }
  // the pipeline agrees on the earliest restart stage of its instances (re-sweeping from an earlier valid checkpoint gives
  // the same bits): per warp with PW, per CTA otherwise
  int ks;
  if constexpr (PW) {
    const int mine = (i < dm.n) ? incr_restart_stage(dm, b, Tk, i) : Tk - 1;
    ks = __reduce_min_sync(0xffffffffu, mine);
  } else {
    if (threadIdx.x == 0) s_ks = Tk - 1;
    __syncthreads();
    if (i < dm.n) atomicMin(&s_ks, incr_restart_stage(dm, b, Tk, i));
    __syncthreads();
    ks = s_ks;
  }
  src.k0 = ks;
  src.nst = Tk - ks + 1;
  if (ptid == 0 && active > 0) {
    const int pre = src.nst < STAGES ? src.nst : STAGES;
    for (int j = 0; j < pre; ++j) src.issue(j);
  }
  if (i >= dm.n) return;
  int st = tick_status(dm, b, Tk, i);
  st |= mhe_solve_incr<T, SmemStageSource<T, PT, STAGES>, Math>(c, dm, b, in, out, Tk, i, src, ks);
  tick_status(dm, b, Tk, i) = st;
  if (status_out != nullptr) status_out[i] = st;
}

}  // namespace dekf
