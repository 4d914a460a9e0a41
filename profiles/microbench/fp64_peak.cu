// FP64 issue-rate microbenchmark for sm_100a: how many DFMA / DMUL / DADD / DSETP+FSEL warp-instructions per
// cycle per SM does a B200 sustain?  (The step kernels sit at ~55 % "sm__pipe_fp64_cycles_active"; this tells
// what 100 % means in instructions per second.)   nvcc -arch=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int ILP>
__global__ void __launch_bounds__(256) k(double* out, int iters, double a, double b)
{
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i)
    x[i] = a + i + threadIdx.x * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (OP == 0) x[i] = fma(x[i], a, b);
        if (OP == 1) x[i] = x[i] * a;
        if (OP == 2) x[i] = x[i] + b;
        if (OP == 3) x[i] = (x[i] > b) ? x[i] - a : x[i] + a; // DSETP + DADDs + FSELs
        if (OP == 4) { x[i] = fma(x[i], a, b); x[i] = x[i] * b; } // DFMA + DMUL
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i)
    s += x[i];
  if (s == 123.456)
    out[threadIdx.x] = s;
}

template <int OP, int ILP>
void run(const char* name, int warps_per_sm, double per_iter_fp64)
{
  int dev = 0;
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, dev);
  double* out;
  cudaMalloc(&out, 4096);
  const int threads = 256, blocks = p.multiProcessorCount * (warps_per_sm * 32 / threads);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<OP, ILP><<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<OP, ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  const double inst = (double)blocks * threads / 32 * iters * 8.0 * ILP * per_iter_fp64; // FP64 warp-instructions
  const double per_s = inst / (ms * 1e-3);
  printf("%-28s ILP=%d warps/SM=%2d : %.3f ms  %.2f T FP64-thread-inst/s  = %.2f FP64 warp-inst/clk/SM at %d MHz (max clock)\n", name, ILP,
         warps_per_sm, ms, per_s * 32 / 1e12, per_s / p.multiProcessorCount / (clk * 1e3), clk / 1000);
  cudaFree(out);
}

int main()
{
  run<0, 4>("DFMA", 16, 1);
  run<0, 8>("DFMA", 16, 1);
  run<0, 8>("DFMA", 32, 1);
  run<0, 2>("DFMA", 16, 1);
  run<0, 1>("DFMA", 16, 1);
  run<0, 1>("DFMA", 64, 1);
  run<1, 8>("DMUL", 16, 1);
  run<2, 8>("DADD", 16, 1);
  run<4, 8>("DFMA+DMUL", 16, 2);
  run<3, 8>("DSETP+2DADD+2FSEL (3 fp64)", 16, 3);
  return 0;
}
