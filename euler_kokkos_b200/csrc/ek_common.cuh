// Shared device-side definitions for the B200 MUSCL-Hancock kernels (sm_100a, FP64).
#pragma once
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define EK_HD __host__ __device__ __forceinline__
#else
// host-only inclusion (tests/emul: the per-cell device functions compiled by g++ for CPU-side checks)
#include <cmath>
#define EK_HD inline
#endif
#include <stdint.h>

namespace ek {

enum : int { ID = 0, IP = 1, IU = 2, IV = 3, IW = 4, IA = 5, IB = 6, IC = 7 };
enum : int { RIEMANN_APPROX = 0, RIEMANN_LLF = 1, RIEMANN_HLL = 2, RIEMANN_HLLC = 3, RIEMANN_HLLD = 4 };
enum : int { BC_UNDEFINED = 0, BC_DIRICHLET = 1, BC_NEUMANN = 2, BC_PERIODIC = 3, BC_COPY = 4 };

// Device-resident scalars of the sync-free time loop (one per context).
struct TimeState {
  double dt;                     // time step used by the step kernels
  double t;                      // current time
  double t_end;
  unsigned long long invdt_bits; // running max of invDt (bit pattern of a non-negative double)
  int steps_done;
  int finished;
};

// Geometry + physics constants handed to every kernel by value (constant bank).
struct Grid {
  int nx, ny, nz;          // interior sizes (nz = 1 in 2-D)
  int gw;                  // ghost width
  int isize, jsize, ksize; // ghosted sizes (ksize = 1 in 2-D)
  int nbvar;
  long long sj, sk, sv;    // strides (in doubles) of j, k, var
  double dx, dy, dz;
  double gamma0, gamma6, smallr, smallc, smallp, smallpp, cfl, slope_type;
  int bc[6];
  int impl_version;
  int gravity_hancock;
};

EK_HD double dmax(double a, double b) { return fmax(a, b); }
EK_HD double dmin(double a, double b) { return fmin(a, b); }

// ---------------------------------------------------------------------------------------------
// Division / square root.  EK_PRECISE_MATH=1 uses IEEE-rounded '/' and sqrt() everywhere (parity
// bisecting build); the default uses MUFU seeds + FMA Newton steps without the IEEE special-case
// slow paths (relative error <= ~1.5 ulp; operands here are always finite, normal and non-zero
// wherever these are used -- the callers floor densities/pressures first).
// ---------------------------------------------------------------------------------------------
#ifndef EK_PRECISE_MATH
#define EK_PRECISE_MATH 0
#endif

EK_HD double rcp_fast(double x)
{
#if EK_PRECISE_MATH || !defined(__CUDA_ARCH__)
  return 1.0 / x;
#else
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
#endif
}

EK_HD double div_fast(double a, double b)
{
#if EK_PRECISE_MATH || !defined(__CUDA_ARCH__)
  return a / b;
#else
  double y = rcp_fast(b);
  double q = a * y;
  double r = fma(-b, q, a); // one residual correction: q is then within 1 ulp
  return fma(r, y, q);
#endif
}

// sqrt for x >= 0 (x == 0 handled)
EK_HD double sqrt_fast(double x)
{
  return sqrt(x); // nvcc's FP64 sqrt is already MUFU.RSQ64H + Newton; kept IEEE for the floors
}

// 1/sqrt(x) for finite normal x > 0: MUFU.RSQ64H seed (rel. error 2^-22.9) + two Newton steps
EK_HD double rsqrt_fast(double x)
{
#if EK_PRECISE_MATH || !defined(__CUDA_ARCH__)
  return 1.0 / sqrt(x);
#else
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x * y, y, 1.0);
  y = fma(y, 0.5 * e, y);
  e = fma(-x * y, y, 1.0);
  y = fma(y, 0.5 * e, y);
  return y;
#endif
}

#ifdef __CUDACC__
// order-preserving atomic max for non-negative doubles
__device__ __forceinline__ void atomic_max_nonneg(unsigned long long* addr, double v)
{
  atomicMax(addr, (unsigned long long)__double_as_longlong(v));
}

__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

#endif // __CUDACC__

} // namespace ek
