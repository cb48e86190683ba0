// tools/micro/issue_model.cu -- does an FP64 warp-instruction hold the scheduler's dispatch port for two
// cycles (so that non-FP64 instructions ADD to the loop time) or can integer / shared-memory / FP32 work issue in
// the FP64 pipe's shadow?  Measures cycles per loop iteration per scheduler for NF DFMAs + NI others, all
// independent chains, 8 warps per scheduler.   nvcc -arch=sm_100a -O3 -o issue_model issue_model.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NF, int NI, int KIND>
__global__ void __launch_bounds__(1024, 1) k(int iters, double* sink, long long* cyc) {
  double a[8];
  unsigned b[8];
  float f[8];
  __shared__ double sh[1024];
  sh[threadIdx.x] = threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = 1.0 + 1e-6 * (threadIdx.x + j); b[j] = threadIdx.x * 2654435761u + j; f[j] = 1.0f + j; }
  const double m = 0.999999, c = 1e-7;
  __syncthreads();
  const unsigned shbase = (unsigned)__cvta_generic_to_shared(sh) + 8u * (threadIdx.x & 511);
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < NF; ++j) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[j & 7]) : "d"(m), "d"(c));
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      if (KIND == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[j & 7]) : "r"(b[(j + 1) & 7]), "r"(0x9e3779b9u));
      if (KIND == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[j & 7]) : "f"(0.999f), "f"(1e-3f));
      if (KIND == 2) { unsigned lo, hi; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(shbase + 8u * (unsigned)j) : "memory"); b[j & 7] ^= lo ^ hi; }
      if (KIND == 4) { float t; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(t) : "d"(a[j & 7])); f[j & 7] += t; }
      if (KIND == 5) { double t; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(f[j & 7])); asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(b[j & 7]), "=r"(b[(j + 1) & 7]) : "d"(t)); }
      if (KIND == 3) asm volatile("mov.b32 %0, %1;" : "=r"(b[j & 7]) : "r"(b[(j + 3) & 7] ));
    }
  }
  long long t1 = clock64();
  double s = 0;
  unsigned u = 0;
  float g = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { s += a[j]; u ^= b[j]; g += f[j]; }
  if (s == -1.0 || u == 0x12345u || g == -3.f) sink[0] = s + u + g;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int NF, int NI, int KIND> void run(const char* what) {
  double* sink; long long* cyc; long long h;
  cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
  const int iters = 20000;
  k<NF, NI, KIND><<<148, 1024>>>(iters, sink, cyc);  // 8 warps per scheduler
  k<NF, NI, KIND><<<148, 1024>>>(iters, sink, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  // per scheduler: 8 warps x (NF + NI) instructions per iteration
  double per_iter = (double)h / iters;
  printf("%-6s NF=%2d NI=%2d: %.2f cycles / iteration / scheduler (8 warps)  -> per warp-iteration %.2f; 2*NF=%d, 2*NF+NI=%d, NF+NI=%d\n", what, NF, NI,
         per_iter, per_iter / 8, 2 * NF, 2 * NF + NI, NF + NI);
  cudaFree(sink); cudaFree(cyc);
}

int main() {
  run<8, 0, 0>("lop3"); run<8, 4, 0>("lop3"); run<8, 8, 0>("lop3"); run<8, 16, 0>("lop3"); run<0, 16, 0>("lop3");
  run<8, 8, 1>("ffma"); run<8, 16, 1>("ffma");
  run<8, 4, 2>("lds"); run<8, 8, 2>("lds"); run<0, 8, 2>("lds");
  run<0, 8, 4>("cvt64>32"); run<8, 4, 4>("cvt64>32"); run<0, 8, 5>("cvt32>64"); run<8, 4, 5>("cvt32>64");
  run<8, 8, 3>("mov"); run<8, 16, 3>("mov");
  return 0;
}
