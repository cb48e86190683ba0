// tools/micro/dfma_forms.cu -- FP64 pipe cost of one instruction by operand kind (vector registers, a
// uniform-register / constant-bank coefficient, an immediate), and of DADD / DMUL, measured as cycles per
// instruction per scheduler with 8 warps x 8 independent chains.  nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cuda_runtime.h>

__constant__ double cC[16] = {0.5000001, 0.1666667, 0.0416667, 0.0083334, 0.0013889, 1.98e-4, 2.48e-5, 2.7e-6,
                              0.9999991, 0.9999992, 0.9999993, 0.9999994, 0.9999995, 0.9999996, 0.9999997, 0.9999998};

template <int MODE> __global__ void __launch_bounds__(1024, 1) k(int iters, double* sink, long long* cyc) {
  double a[8], x[8], y[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = 1.0 + 1e-6 * (threadIdx.x + j);
    x[j] = 0.999999 - 1e-9 * (threadIdx.x + 3 * j);
    y[j] = 1e-7 * (1 + j + threadIdx.x);
  }
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) a[j] = fma(a[j], x[j], y[j]);                 // three vector registers
      if (MODE == 1) a[j] = fma(a[j], x[j], cC[j]);                // coefficient from constant memory (UR / c[] operand)
      if (MODE == 2) a[j] = fma(a[j], x[j], 0.5);                  // immediate
      if (MODE == 3) a[j] = a[j] + y[j];                           // DADD
      if (MODE == 4) a[j] = a[j] * x[j];                           // DMUL
      if (MODE == 5) a[j] = fma(a[j], cC[8 + j], cC[j]);           // two constant operands
      if (MODE == 6) a[j] = fma(a[j], x[(j + 1) & 7], y[(j + 3) & 7]);  // three vector registers, operands shared between chains
      if (MODE == 7) a[j] = fma(a[j], a[j], a[j]);                 // one register three times (the rcp refinement's shape)
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += a[j];
  if (s == -1.0) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// dependent-issue latency: ONE warp per scheduler, one chain of NCHAIN-way interleaved dependent DFMAs
template <int NCHAIN> __global__ void __launch_bounds__(128, 1) lat(int iters, double* sink, long long* cyc) {
  double a[8], x = 0.999999 - 1e-9 * threadIdx.x, y = 1e-7 * (1 + threadIdx.x);
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 1.0 + 1e-6 * (threadIdx.x + j);
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int j = 0; j < NCHAIN; ++j) a[j] = fma(a[j], x, y);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += a[j];
  if (s == -1.0) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int NCHAIN> void runlat() {
  double* sink; long long* cyc; long long h;
  cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
  const int iters = 20000;
  lat<NCHAIN><<<148, 128>>>(iters, sink, cyc);
  lat<NCHAIN><<<148, 128>>>(iters, sink, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("one warp per scheduler, %d independent chain(s): %.2f cycles per DFMA (1 chain = dependent-issue latency)\n", NCHAIN,
         (double)h / iters / (8.0 * NCHAIN));
  cudaFree(sink); cudaFree(cyc);
}

template <int MODE> void run(const char* what) {
  double* sink; long long* cyc; long long h;
  cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
  const int iters = 20000;
  k<MODE><<<148, 1024>>>(iters, sink, cyc);
  k<MODE><<<148, 1024>>>(iters, sink, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-46s %.3f cycles per instruction per scheduler\n", what, (double)h / iters / 64.0);
  cudaFree(sink); cudaFree(cyc);
}

int main() {
  run<0>("DFMA R, R, R (own operands per chain)");
  run<6>("DFMA R, R, R (operands shared between chains)");
  run<1>("DFMA R, R, const");
  run<2>("DFMA R, R, imm");
  run<5>("DFMA R, const, const");
  run<7>("DFMA R, R, R (same register x3)");
  run<3>("DADD R, R");
  run<4>("DMUL R, R");
  runlat<1>(); runlat<2>(); runlat<3>(); runlat<4>(); runlat<6>(); runlat<8>();
  return 0;
}
