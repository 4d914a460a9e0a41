// Probe: cp.async.bulk.tensor box loads of FP64 data as the TMA-fed MHD kernel issues them (4-D tensor x,y,z,var;
// box 36 x 19 x 1 x 1; zero fill outside the array).  Usage: tma_probe <variant>
//   0: 4-D f64 box 36x19x1x1 (the kernel's load)   1: same as 2-D   2: 4-D with L2 promotion none   3: 4-D f64 as UINT64
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap pmap, const CUtensorMap* gmap, double* out, int c0, int c1, int c2, int c3, unsigned bytes)
{
  const CUtensorMap& map = gmap ? *gmap : pmap;
  extern __shared__ __align__(1024) unsigned char raw[];
  double* dst = reinterpret_cast<double*>(raw);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(raw + 8192);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
    if (RANK == 4)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(s32(dst)),
                   "l"((unsigned long long)&map), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                   : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(dst)),
                   "l"((unsigned long long)&map), "r"(s32(bar)), "r"(c0), "r"(c1)
                   : "memory");
  }
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(s32(bar)) : "memory");
  for (int i = threadIdx.x; i < 36 * 19; i += blockDim.x)
    out[i] = dst[i];
}

int main(int argc, char** argv)
{
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int NX = 22, NY = 18, NZ = 14, NV = 8;
  std::vector<double> h((size_t)NX * NY * NZ * NV);
  for (size_t i = 0; i < h.size(); ++i)
    h[i] = (double)i;
  double *d, *out;
  cudaMalloc(&d, h.size() * 8);
  cudaMalloc(&out, 36 * 19 * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  alignas(64) CUtensorMap map;
  const cuuint64_t dims[4] = { NX, NY, NZ, NV };
  const cuuint64_t strides[3] = { NX * 8, NX * NY * 8, (cuuint64_t)NX * NY * NZ * 8 };
  const cuuint32_t bx = (variant == 4) ? 32 : (variant == 6) ? 16 : 36;
  const cuuint32_t box[4] = { bx, 19, 1, 1 }, es[4] = { 1, 1, 1, 1 };
  const int rank = (variant == 1) ? 2 : 4;
  CUresult r = ((EncodeFn)fn)(&map, variant == 3 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, rank, d, dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              variant == 2 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d encode rc=%d\n", variant, (int)r);
  const unsigned bytes = bx * 19 * 8;
  CUtensorMap* gmap = nullptr;
  if (variant == 5) {
    cudaMalloc(&gmap, sizeof(CUtensorMap));
    cudaMemcpy(gmap, &map, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
  }
  cudaFuncSetAttribute(probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  if (rank == 4)
    probe<4><<<1, 512, 16384>>>(map, gmap, out, 1, 1, 3, 2, bytes);
  else
    probe<2><<<1, 512, 16384>>>(map, gmap, out, 1, 1, 0, 0, bytes);
  cudaError_t e = cudaDeviceSynchronize();
  printf("  kernel: %s\n", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<double> o(36 * 19);
    cudaMemcpy(o.data(), out, o.size() * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int j = 0; j < 19; ++j)
      for (int i = 0; i < 36; ++i) {
        const int x = 1 + i, y = 1 + j, z = (rank == 4) ? 3 : 0, v = (rank == 4) ? 2 : 0;
        const double want = (x < NX && y < NY) ? (double)(x + NX * (y + NY * (z + (size_t)NZ * v))) : 0.0;
        if (i < (int)bx) bad += (o[j * bx + i] != want);
      }
    printf("  mismatches: %d  (o[0]=%g o[1]=%g o[36]=%g)\n", bad, o[0], o[1], o[36]);
  }
  return 0;
}
