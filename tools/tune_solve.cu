// Kernel tuning harness for the window-solve kernel (dev tool, not a product path, not shipped in libdekf_b200.so).
// Builds synthetic but well-conditioned window state for n instances and times launch-configuration variants of
// k_solve_tma / k_solve with CUDA events.  Usage: tune_solve [n=65536] [N=20] [reps=30]
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/_build/tune_solve tools/tune_solve.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../decentralized_ekf_mhe_b200/csrc/host_setup.hpp"
#include "../decentralized_ekf_mhe_b200/csrc/solve_tma.cuh"

using namespace dekf;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      std::exit(1);                                                                \
    }                                                                              \
  } while (0)

template <typename T>
__global__ void k_fill(Dims dm, Buffers<T> b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dm.ns) return;
  const int ns = dm.ns;
  unsigned s = 1234567u + 747796405u * (unsigned)i;
  auto rnd = [&]() {
    s = s * 1664525u + 1013904223u;
    return (double)(s >> 8) * (1.0 / 16777216.0) - 0.5;
  };
  for (int f = 0; f < 45; ++f) b.arr_P[(size_t)f * ns + i] = T(0);
  const int diag[3] = {0, 3, 5};
  for (int f = 0; f < 3; ++f) {
    b.arr_P[(size_t)(0 + diag[f]) * ns + i] = T(1e-4);
    b.arr_P[(size_t)(6 + diag[f]) * ns + i] = T(1e-3);
    b.arr_P[(size_t)(12 + diag[f]) * ns + i] = T(1e-4);
  }
  for (int f = 0; f < 9; ++f) b.arr_x[(size_t)f * ns + i] = T(0.1 * rnd());
  for (int k = 0; k < dm.NW; ++k) {
    T *rec = b.win + (size_t)k * REC_SIZE * ns + i;
    double q[4] = {1.0 + 0.1 * rnd(), 0.2 * rnd(), 0.2 * rnd(), 0.5 * rnd()};
    const M3<T> R = quat_to_rot<T>((T)q[0], (T)q[1], (T)q[2], (T)q[3]);
    for (int f = 0; f < 9; ++f) rec[(size_t)(REC_R + f) * ns] = R.a[f];
    for (int f = 0; f < 3; ++f) rec[(size_t)(REC_AS + f) * ns] = T(0.5 * rnd());
    const double lam = 4.0e3 * (1.0 + rnd());
    rec[(size_t)(REC_LAM + 0) * ns] = T(lam);
    rec[(size_t)(REC_LAM + 1) * ns] = T(0.1 * lam * rnd());
    rec[(size_t)(REC_LAM + 2) * ns] = T(0.1 * lam * rnd());
    rec[(size_t)(REC_LAM + 3) * ns] = T(lam * 1.1);
    rec[(size_t)(REC_LAM + 4) * ns] = T(0.1 * lam * rnd());
    rec[(size_t)(REC_LAM + 5) * ns] = T(lam * 0.9);
    for (int f = 0; f < 3; ++f) rec[(size_t)(REC_ETA + f) * ns] = T(lam * 0.5 * rnd());
    for (int f = 0; f < 3; ++f) rec[(size_t)(REC_DLT + f) * ns] = T(0.0025 + 1e-4 * rnd());
    rec[(size_t)REC_FLAG * ns] = (k % 10) < 3 ? T(1) : T(0);  // 6 of 20 stages carry a VO equality row
  }
  b.status[i] = 0;
}

typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                              const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename T>
CUtensorMap make_map(const Dims &dm, T *win, int tile) {
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  CUtensorMap m;
  const cuuint64_t gdim[2] = {(cuuint64_t)dm.ns, (cuuint64_t)dm.NW * REC_SIZE};
  const cuuint64_t gstride[1] = {(cuuint64_t)dm.ns * sizeof(T)};
  const cuuint32_t box[2] = {(cuuint32_t)tile, (cuuint32_t)REC_SIZE};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = ((encode_fn)fn)(&m, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, win, gdim,
                               gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::fprintf(stderr, "encode failed %d\n", (int)r);
    std::exit(1);
  }
  return m;
}

template <typename T>
struct Ctx {
  Dims dm;
  MheConst<T> mc;
  Buffers<T> b;
  Inputs in;
  Outputs out;
  int Tk;
  int reps;
};

template <typename T, int TILE, int STAGES, int MINB, typename Math = DefaultMath<T>, bool XS = false, bool PW = false>
void run_variant(Ctx<T> &c, const char *name) {
  auto kern = k_solve_tma<T, TILE, STAGES, MINB, Math, XS, PW>;
  const size_t smem = solve_tma_smem_bytes<T, TILE, STAGES, XS>();
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, kern));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TILE, smem));
  CUtensorMap map = make_map<T>(c.dm, c.b.win, PW ? 32 : TILE);
  const int grid = c.dm.ns / TILE;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; ++w) kern<<<grid, TILE, smem>>>(map, c.mc, c.dm, c.b, c.in, c.out, c.Tk, nullptr);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < c.reps; ++r) kern<<<grid, TILE, smem>>>(map, c.mc, c.dm, c.b, c.in, c.out, c.Tk, nullptr);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  CK(cudaGetLastError());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<double> x(9);
  CK(cudaMemcpy(x.data(), c.out.x, sizeof(double), cudaMemcpyDeviceToHost));
  std::printf("%-28s T=%zu tile=%3d stages=%d minb=%d regs=%3d local=%4zu smem=%6zu occ=%d CTA/SM  %8.2f us/launch  x0=%.6e\n", name,
              sizeof(T), TILE, STAGES, MINB, fa.numRegs, (size_t)fa.localSizeBytes, smem, occ, 1e3 * ms / c.reps, x[0]);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
}

template <typename T>
__global__ void __launch_bounds__(128) k_solve_plain(const MheConst<T> c, const Dims dm, const Buffers<T> b, const Inputs in,
                                                     const Outputs out, int Tk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dm.n) return;
  b.status[i] |= mhe_solve<T>(c, dm, b, in, out, Tk, i);
}

template <typename T>
void run_all(int n, int N, int reps) {
  dekf_config cfg;
  fill_go1_defaults(&cfg);
  cfg.n_instances = n;
  cfg.N = N;
  Ctx<T> c;
  c.dm = make_dims(cfg);
  c.mc = make_mhe_const<T>(cfg);
  c.reps = reps;
  c.Tk = 5 * N + 3;
  const StateSizes s = state_sizes(c.dm);
  std::memset(&c.b, 0, sizeof(c.b));
  CK(cudaMalloc((void **)&c.b.arr_P, s.arr_P * sizeof(T)));
  CK(cudaMalloc((void **)&c.b.arr_x, s.arr_x * sizeof(T)));
  CK(cudaMalloc((void **)&c.b.win, s.win * sizeof(T)));
  CK(cudaMalloc((void **)&c.b.status, s.status * sizeof(int32_t)));
  std::memset(&c.in, 0, sizeof(c.in));
  std::memset(&c.out, 0, sizeof(c.out));
  double *gy, *ox, *ov;
  CK(cudaMalloc((void **)&gy, (size_t)3 * n * sizeof(double)));
  CK(cudaMemset(gy, 0, (size_t)3 * n * sizeof(double)));
  CK(cudaMalloc((void **)&ox, (size_t)9 * n * sizeof(double)));
  CK(cudaMalloc((void **)&ov, (size_t)3 * n * sizeof(double)));
  c.in.gyro = gy;
  c.out.x = ox;
  c.out.v_body = ov;
  k_fill<T><<<(c.dm.ns + 127) / 128, 128>>>(c.dm, c.b);
  CK(cudaDeviceSynchronize());
  std::printf("# n=%d N=%d elt=%zu reps=%d\n", n, N, sizeof(T), reps);
  {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) k_solve_plain<T><<<(n + 127) / 128, 128>>>(c.mc, c.dm, c.b, c.in, c.out, c.Tk);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) k_solve_plain<T><<<(n + 127) / 128, 128>>>(c.mc, c.dm, c.b, c.in, c.out, c.Tk);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    std::printf("%-28s %8.2f us/launch\n", "global-load kernel", 1e3 * ms / reps);
  }
  run_variant<T, 128, SolveCfg<T>::kStages, SolveCfg<T>::kMinB, DefaultMath<T>, false, SolveCfg<T>::kPerWarp>(c, "product config");
  run_variant<T, 128, 3, SolveCfg<T>::kMinB, DefaultMath<T>, false, false>(c, "CTA pipeline x3");
  run_variant<T, 128, 3, SolveCfg<T>::kMinB, DefaultMath<T>, false, true>(c, "warp pipelines x3");
  run_variant<T, 128, 4, SolveCfg<T>::kMinB, DefaultMath<T>, false, true>(c, "warp pipelines x4");
  run_variant<T, 128, 2, SolveCfg<T>::kMinB, DefaultMath<T>, false, true>(c, "warp pipelines x2");
  run_variant<T, 64, 3, SolveCfg<T>::kMinB == 1 ? 1 : 8, DefaultMath<T>, false, true>(c, "warp pipelines x3, 64-thread CTAs");
  run_variant<T, 32, 3, SolveCfg<T>::kMinB == 1 ? 1 : 16, DefaultMath<T>, false, true>(c, "warp pipelines x3, 32-thread CTAs");
  run_variant<T, 128, 3, 1, MathSel<2, 1>>(c, "tma 128x3 meas2 prop1 (r01)");
  cudaFree(c.b.arr_P);
  cudaFree(c.b.arr_x);
  cudaFree(c.b.win);
  cudaFree(c.b.status);
  cudaFree(gy);
  cudaFree(ox);
  cudaFree(ov);
}

int main(int argc, char **argv) {
  const int n = argc > 1 ? std::atoi(argv[1]) : 65536;
  const int N = argc > 2 ? std::atoi(argv[2]) : 20;
  const int reps = argc > 3 ? std::atoi(argv[3]) : 30;
  run_all<double>(n, N, reps);
  run_all<float>(n, N, reps);
  return 0;
}
