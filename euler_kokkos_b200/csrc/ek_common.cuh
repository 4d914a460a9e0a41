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
  int max_steps; // iterations after which the loop is finished (device resident so that a captured graph does not bake it in)
};

// Geometry + physics constants handed to every kernel by value (constant bank).
struct Grid {
  int nx, ny, nz;          // interior sizes (nz = 1 in 2-D)
  int gw;                  // ghost width
  int isize, jsize, ksize; // ghosted sizes (ksize = 1 in 2-D)
  int nbvar;
  long long sj, sk, sv;    // strides (in doubles) of j, k, var
  double dx, dy, dz;
  double invdx, invdy, invdz; // 1/dx, 1/dy, 1/dz (IEEE, computed on the host)
  double gamma0, gamma6, smallr, smallc, smallp, smallpp, cfl, slope_type;
  double entho;               // 1/(gamma0-1)
  double slope_k;             // 0.5/slope_type (limited_slope_ref)
  int bc[6];
  int impl_version;
  int gravity_hancock;
};

// max/min as compare+select (3 instructions; fmax/fmin expand to a 6-instruction NaN-aware sequence
// on sm_100a).  NaN semantics differ from fmax/fmin only when an operand IS a NaN.
EK_HD double dmax(double a, double b) { return (a > b) ? a : b; }
EK_HD double dmin(double a, double b) { return (a < b) ? a : b; }

// max(a, 0) / min(a, 0).  With a literal 0.0 the compiler proves one operand non-NaN, turns the compare+select above
// back into max.f64 / min.f64 and ptxas expands that to the NaN-aware sequence again (DSETP.MAX + FSEL + SEL +
// predicated LOP3 + moves: 38 of them per thread in MHD stage R, cuobjdump).  The comparison is therefore written in
// PTX: setp + selp = one DSETP and two FSEL.  Same result for every non-NaN a (and 0 for a NaN, as before).
#ifndef EK_ASM_MAX0
#define EK_ASM_MAX0 1
#endif
EK_HD double dmax0(double a)
{
#if EK_ASM_MAX0 && defined(__CUDA_ARCH__)
  double r;
  asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, 0d0000000000000000;\n\tselp.f64 %0, %1, 0d0000000000000000, p;\n\t}" : "=d"(r) : "d"(a));
  return r;
#else
  return (a > 0.0) ? a : 0.0;
#endif
}
EK_HD double dmin0(double a)
{
#if EK_ASM_MAX0 && defined(__CUDA_ARCH__)
  double r;
  asm("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %1, 0d0000000000000000;\n\tselp.f64 %0, %1, 0d0000000000000000, p;\n\t}" : "=d"(r) : "d"(a));
  return r;
#else
  return (a < 0.0) ? a : 0.0;
#endif
}

// Limited slope of the reference (slope_unsplit_hydro_{2d,3d}_scalar, HydroBaseFunctor3D.h:413-461;
// same formula in MHDBaseFunctor3D.h:262-337 and for the face field in :563-668):
//   dlft = st*(q-qm), drgt = st*(qp-q), dcen = (qp-qm)/2, dlim = (dlft*drgt <= 0) ? 0 : min(|dlft|,|drgt|),
//   slope = sign(dcen) * min(dlim, |dcen|)
// evaluated with 8 FP64-pipe operations: "dlft*drgt <= 0" is decided on the sign bits (a zero factor
// makes the min zero anyway) and sign(dcen) is transplanted with copysign (when the signs of dlft and
// drgt agree, dcen has that sign too; the sign of a zero result is immaterial).
// sk = 0.5/st.  For st in {1, 2} (the only non-zero values the reference supports) the scalings by st
// and sk are exact, so st*min(|a|,|b|,sk*|c|) == min(|st*a|,|st*b|,|c/2|) bit for bit: 7 FP64 operations.
EK_HD double limited_slope_ref(double st, double sk, double qm, double q, double qp)
{
  const double a = q - qm;
  const double b = qp - q;
  const double c = qp - qm;
  const double m = st * dmin(dmin(fabs(a), fabs(b)), sk * fabs(c));
#ifdef __CUDA_ARCH__
  const bool opposite = (__double2hiint(a) ^ __double2hiint(b)) < 0;
#else
  const bool opposite = std::signbit(a) != std::signbit(b);
#endif
  return opposite ? 0.0 : copysign(m, c);
}

// ---------------------------------------------------------------------------------------------
// Division / square root.  EK_PRECISE_MATH=1 uses IEEE-rounded '/' and sqrt() everywhere (parity
// bisecting build); the default uses MUFU seeds + ONE cubically convergent correction without the IEEE
// special-case slow paths (relative error <= ~1.5 ulp; operands here are always finite, normal and non-zero
// wherever these are used -- the callers floor densities/pressures first).
// The 64-bit MUFU seeds carry ~23 bits (relative error e0 ~ 2^-23).  Two Newton steps (e -> e^2 -> e^4) were used in
// round 1; one third-order step (e -> e^3 ~ 2^-69) is enough for a double and is one to three FP64 instructions
// shorter: 1/x = y0 (1 + e + e^2 + O(e^3)) with e = 1 - x y0;  x^-1/2 = y0 (1 + e/2 + 3 e^2/8 + O(e^3)) with
// e = 1 - x y0^2.  The MHD stage R executes ~84 reciprocals, ~51 square roots and ~6 reciprocal square roots per cell.
#ifndef EK_CUBIC_SEEDS
#define EK_CUBIC_SEEDS 1
#endif
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
#if EK_CUBIC_SEEDS
  const double e = fma(-x, y, 1.0);
  return fma(y, fma(e, e, e), y);
#else
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
#endif
#endif
}

EK_HD double div_fast(double a, double b)
{
#if EK_PRECISE_MATH || !defined(__CUDA_ARCH__)
  return a / b;
#else
#if EK_CUBIC_SEEDS
  return a * rcp_fast(b); // rcp_fast is within ~0.6 ulp after the cubic step: the product is within ~1.2 ulp
#else
  double y = rcp_fast(b);
  double q = a * y;
  double r = fma(-b, q, a); // one residual correction: q is then within 1 ulp
  return fma(r, y, q);
#endif
#endif
}

// sqrt for x >= 0 (zero and subnormal inputs give 0): MUFU.RSQ64H seed, one Goldschmidt step and a
// Newton correction (7 FP64 ops, error <= ~1 ulp), without the IEEE slow-path branch of sqrt()
EK_HD double sqrt_fast(double x)
{
#if EK_PRECISE_MATH || !defined(__CUDA_ARCH__)
  return sqrt(x);
#else
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double g = x * y;
  double h = 0.5 * y;
  const double r = fma(-h, g, 0.5); // = e/2
#if EK_CUBIC_SEEDS
  g = fma(g, fma(1.5, r, 1.0) * r, g); // sqrt(x) = g0 (1 + r + 3 r^2 / 2 + O(r^3))
#else
  g = fma(g, r, g);
  h = fma(h, r, h);
  const double d = fma(-g, g, x);
  g = fma(d, h, g);
#endif
  return (x >= 2.2250738585072014e-308) ? g : 0.0;
#endif
}

// sqrt for finite normal x > 0 (no zero guard)
EK_HD double sqrt_pos_fast(double x)
{
#if EK_PRECISE_MATH || !defined(__CUDA_ARCH__)
  return sqrt(x);
#else
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double g = x * y;
  double h = 0.5 * y;
  const double r = fma(-h, g, 0.5);
#if EK_CUBIC_SEEDS
  return fma(g, fma(1.5, r, 1.0) * r, g);
#else
  g = fma(g, r, g);
  h = fma(h, r, h);
  const double d = fma(-g, g, x);
  return fma(d, h, g);
#endif
#endif
}

// 1/sqrt(x) for finite normal x > 0: MUFU.RSQ64H seed (rel. error 2^-22.9) + two Newton steps
EK_HD double rsqrt_fast(double x)
{
#if EK_PRECISE_MATH || !defined(__CUDA_ARCH__)
  return 1.0 / sqrt(x);
#else
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#if EK_CUBIC_SEEDS
  const double e = fma(-x * y, y, 1.0);
  return fma(y, fma(0.375, e, 0.5) * e, y);
#else
  double e = fma(-x * y, y, 1.0);
  y = fma(y, 0.5 * e, y);
  e = fma(-x * y, y, 1.0);
  y = fma(y, 0.5 * e, y);
  return y;
#endif
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
