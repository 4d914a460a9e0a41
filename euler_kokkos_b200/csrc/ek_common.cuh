// Shared device-side definitions for the B200 MUSCL-Hancock kernels (sm_100a, FP64).
#pragma once
#include <cuda_runtime.h>
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

__device__ __forceinline__ double dmax(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ double dmin(double a, double b) { return fmin(a, b); }

// ---------------------------------------------------------------------------------------------
// Division / square root.  EK_PRECISE_MATH=1 uses IEEE-rounded '/' and sqrt() everywhere (parity
// bisecting build); the default uses MUFU seeds + FMA Newton steps without the IEEE special-case
// slow paths (relative error <= ~1.5 ulp; operands here are always finite, normal and non-zero
// wherever these are used -- the callers floor densities/pressures first).
// ---------------------------------------------------------------------------------------------
#ifndef EK_PRECISE_MATH
#define EK_PRECISE_MATH 0
#endif

__device__ __forceinline__ double rcp_fast(double x)
{
#if EK_PRECISE_MATH
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

__device__ __forceinline__ double div_fast(double a, double b)
{
#if EK_PRECISE_MATH
  return a / b;
#else
  double y = rcp_fast(b);
  double q = a * y;
  double r = fma(-b, q, a); // one residual correction: q is then within 1 ulp
  return fma(r, y, q);
#endif
}

// sqrt for x >= 0 (x == 0 handled)
__device__ __forceinline__ double sqrt_fast(double x)
{
#if EK_PRECISE_MATH
  return sqrt(x);
#else
  return sqrt(x); // nvcc's FP64 sqrt is already MUFU.RSQ64H + Newton; kept IEEE for the floors
#endif
}

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

} // namespace ek
