// FP64 pipe micro-benchmark (dev tool): DFMA issue rate per SM as a function of resident warps and of the number of
// independent dependency chains per thread.  Tells how much ILP x TLP the window-solve kernel needs to fill the pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/ubench_fp64 tools/ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CH, typename T>
__global__ void k_chain(T *out, int iters, T a, T b, long long *cycles) {
  T x[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) x[j] = (T)(threadIdx.x + j);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int j = 0; j < CH; ++j) x[j] = x[j] * a + b;
    }
  }
  const long long t1 = clock64();
  T s = (T)0;
#pragma unroll
  for (int j = 0; j < CH; ++j) s += x[j];
  if (s == (T)123456789) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int CH, typename T>
void run(int warps_per_sm, int sms, T *out, long long *dcyc) {
  const int iters = 2048;
  const int threads = 32 * warps_per_sm;
  k_chain<CH, T><<<sms, threads>>>(out, 16, (T)1.0000001, (T)1e-7, dcyc);
  cudaDeviceSynchronize();
  k_chain<CH, T><<<sms, threads>>>(out, iters, (T)1.0000001, (T)1e-7, dcyc);
  cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, dcyc, sizeof(cyc), cudaMemcpyDeviceToHost);
  const double fma_per_warp = (double)iters * 8 * CH;
  const double warp_fma_per_clk_sm = fma_per_warp * warps_per_sm / (double)cyc;
  std::printf("%s warps/SM=%2d chains=%d : %7.3f warp-FMA/clk/SM (%6.1f lane-FMA/clk/SM), %6.2f clk per dependent FMA\n",
              sizeof(T) == 8 ? "fp64" : "fp32", warps_per_sm, CH, warp_fma_per_clk_sm, 32 * warp_fma_per_clk_sm,
              (double)cyc / ((double)iters * 8));
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out;
  long long *dcyc;
  cudaMalloc(&out, sizeof(double) * sms * 1024);
  cudaMalloc(&dcyc, sizeof(long long));
  const int ws[] = {1, 4, 8, 12, 16, 32};
  for (int w : ws) {
    run<1, double>(w, sms, out, dcyc);
    run<2, double>(w, sms, out, dcyc);
    run<4, double>(w, sms, out, dcyc);
    run<8, double>(w, sms, out, dcyc);
  }
  for (int w : ws) {
    run<1, float>(w, sms, (float *)out, dcyc);
    run<4, float>(w, sms, (float *)out, dcyc);
    run<8, float>(w, sms, (float *)out, dcyc);
  }
  return 0;
}
