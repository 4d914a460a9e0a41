// Micro-benchmark (evidence for DESIGN.md): achievable HBM bandwidth of the access patterns of the MHD
// trace stage on B200 -- NW tiled (AoSoA, 32-lane) store streams per cell + NR load streams.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_aosoa stream_aosoa.cu && ./stream_aosoa
#include <cstdio>
#include <cuda_runtime.h>
template <int NR, int NW>
__global__ void __launch_bounds__(256) k(const double* __restrict__ in, double* __restrict__ out, unsigned n)
{
  const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const double* p = in + (size_t)(c >> 5) * (NR * 32) + (c & 31);
  double s = 0;
#pragma unroll
  for (int v = 0; v < NR; ++v) s += p[32 * v];
  double* q = out + (size_t)(c >> 5) * (NW * 32) + (c & 31);
#pragma unroll
  for (int v = 0; v < NW; ++v) q[32 * v] = s + v;
}
template <int NR, int NW>
void run(unsigned n)
{
  double *in, *out;
  cudaMalloc(&in, (size_t)n * (NR ? NR : 1) * 8);
  cudaMalloc(&out, (size_t)n * NW * 8);
  cudaMemset(in, 0, (size_t)n * (NR ? NR : 1) * 8);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int it = 0; it < 3; ++it) k<NR, NW><<<(n + 255) / 256, 256>>>(in, out, n);
  cudaEventRecord(a);
  const int reps = 10;
  for (int it = 0; it < reps; ++it) k<NR, NW><<<(n + 255) / 256, 256>>>(in, out, n);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= reps;
  printf("read %2d + write %2d streams/cell: %.3f ms, %.0f GB/s\n", NR, NW, ms, (double)n * (NR + NW) * 8 / ms / 1e6);
  cudaFree(in); cudaFree(out);
}
int main()
{
  const unsigned n = 262u * 262u * 262u; // 256^3 + ghosts
  run<8, 8>(n);
  run<19, 19>(n);
  run<0, 38>(n);
  run<14, 38>(n);
  run<38, 18>(n);
  run<38, 0 + 1>(n);
  run<26, 8>(n);
  return 0;
}
