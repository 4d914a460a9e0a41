// Host-side dispatch of the fused hydro step: picks the template instance (dimension, Riemann
// solver, gravity) and sizes the grid for the 148 SMs of a B200.
#include "ek_internal.h"
#include <cstdlib>
#include "hydro_kernels.cuh"

namespace ek {

#ifndef EK_HYDRO_TY
// interior rows per tile (TY+4 warps per CTA).  11 -> 15 warps: the 12 rows that solve Riemann problems
// and the 11 that update spread 3/3/3/3 over the four warp schedulers of an SM; with TY = 12 one
// scheduler carries 4 rows and every barrier waits for it (measured: 14.0 -> 12.9 ms at 512^3).
#define EK_HYDRO_TY 11
#endif
constexpr int TY = EK_HYDRO_TY;
#ifndef EK_HYDRO_TYF
// interior rows per tile of the 3-D "folded halo rows" layout (hydro_kernels.cuh): TYF+2 = 16 warps per CTA, the 15
// that solve Riemann problems spread 4/4/4/3 over the warp schedulers and the row-1 warp rides with the 3
#define EK_HYDRO_TYF 14
#endif
constexpr int TYF = EK_HYDRO_TYF;
// (Measured in round 2 and removed: folded rows on a 28 x 6 tile, 8 warps, TWO resident CTAs per SM whose barriers are
//  independent -- 11.96 ms against 10.88 on blast 512^3, 36.1 against 35.0 on Rayleigh-Taylor: the de-phasing of the two
//  CTAs does not pay for 13 % more traces and 7 % more Riemann problems per cell of the smaller tile.)

// 3-D tile layout: 1 = folded halo rows, 28 x 14 tile, 16 warps; 0 = one warp per tile row incl. halo rows, 28 x 11,
// 15 warps.  Measured at 512^3 (profiles/r01_hydro_fold_ncu.txt): the folded layout executes 3.4 % fewer
// instructions and halves the barrier stalls, but the three lock-step HLLC problems come out less interleaved (128
// registers used instead of 122; fixed-latency "wait" stalls 0.98 -> 1.93 per issue) and it ends 1.6 %
// slower (10.90 vs 10.73 ms); with the iterative two-shock solver + gravity it is 8 % faster (34.5 vs 37.5 ms).
// Default: folded for riemann = approx, unfolded otherwise; EK_HYDRO_FOLD=0/1 forces one.
static int fold_rows(int riemann)
{
  const int v = getenv("EK_HYDRO_FOLD") ? atoi(getenv("EK_HYDRO_FOLD")) : -1;
  if (v < 0)
    return riemann == RIEMANN_APPROX ? TYF : 0;
  return v == 0 ? 0 : TYF;
}

// z-chunk length: enough CTAs to fill whole waves over the SMs, long enough to amortise the
// 3-plane pipeline warm-up of every chunk.
static int pick_zchunk(int nz, long long xy_ctas, int sm_count, int min_chunks = 1, int min_len = 8)
{
  int best = nz;
  double best_cost = 1e300;
  for (int nchunk = min_chunks; nchunk <= nz; ++nchunk) {
    const int zc = (nz + nchunk - 1) / nchunk;
    if (zc < min_len && nchunk > min_chunks)
      break;
    const long long real_chunks = (nz + zc - 1) / zc;
    const long long ctas = xy_ctas * real_chunks;
    const long long waves = (ctas + sm_count - 1) / sm_count;
    const double tail = (double)(waves * sm_count) / (double)ctas; // idle-SM factor of the last wave
    const double warm = (double)(zc + 2.5) / (double)zc; // pipeline warm-up per chunk
    const double cost = tail * warm;
    if (cost < best_cost - 1e-12) {
      best_cost = cost;
      best = zc;
    }
  }
  return best;
}

// z-chunking of the 3-D step.  With split = true the chunks are laid out so that the launch can be cut in two:
// the two END chunks (they produce the planes the z neighbours need as ghosts) and the MIDDLE chunks.
struct ZPlan {
  int zchunk, nchunk;
  bool splittable;
};
static ZPlan z_plan(const Ctx& c, long long xy_ctas, bool want_split)
{
  ZPlan z;
  const int nz = c.g.nz, gw = c.g.gw;
  const int slots = c.sm_count; // one resident CTA per SM
  z.zchunk = pick_zchunk(nz, xy_ctas, slots);
  z.nchunk = (nz + z.zchunk - 1) / z.zchunk;
  z.splittable = false;
  if (!want_split)
    return z;
  const int zc = (z.nchunk >= 3) ? z.zchunk : pick_zchunk(nz, xy_ctas, slots, 3, gw);
  const int nc = (nz + zc - 1) / zc;
  const int last = nz - (nc - 1) * zc;
  if (nc >= 3 && zc >= gw && last >= gw) {
    z.zchunk = zc;
    z.nchunk = nc;
    z.splittable = true;
  }
  return z;
}

static int tile_rows(int dim, int riemann) { return (dim == 3 && fold_rows(riemann) > 0) ? fold_rows(riemann) : TY; }
static long long hydro_xy_ctas(const Grid& g, int riemann)
{
  const int ty = tile_rows(g.nz > 1 ? 3 : 2, riemann);
  return (long long)((g.nx + HTX - 1) / HTX) * ((g.ny + ty - 1) / ty);
}

bool hydro_z_splittable(const Ctx& c) { return c.p.dim == 3 && z_plan(c, hydro_xy_ctas(c.g, c.p.riemann), true).splittable; }

// part: 0 = every chunk, 1 = the two end chunks, 2 = the middle chunks (1 and 2 only when splittable)
template <int DIM, int RS, bool GRAV>
static cudaError_t launch_one(Ctx& c, const double* Uin, double* Uout, int part, cudaStream_t stream)
{
  using Smem = HydroSmem<DIM, TY>;
  void (*kern)(const HydroArgs) = hydro_step_kernel<DIM, RS, TY, GRAV>;
  size_t smem = sizeof(Smem);
  int ty = TY, warps = TY + 4;
  if constexpr (DIM == 3) {
    const int tr = tile_rows(3, RS);
    if (tr == TYF) {
      kern = hydro_step_kernel<DIM, RS, TYF, GRAV, 1>;
      smem = sizeof(HydroSmem<DIM, TYF>);
      ty = TYF;
      warps = TYF + 2;
    }
  }
  // (idempotent; set on every launch path rather than cached in a static, so that contexts on several host threads
  //  cannot race on a flag)
  if (c.smem_attr_for != (const void*)kern) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
      return e;
    c.smem_attr_for = (const void*)kern;
  }
  HydroArgs a;
  a.g = c.g;
  a.Uin = Uin;
  a.Uout = Uout;
  a.grav = c.grav;
  a.ts = c.ts;
  dim3 block(32, warps, 1);
  dim3 grid((c.g.nx + HTX - 1) / HTX, (c.g.ny + ty - 1) / ty, 1);
  a.zchunk = 1;
  a.zfirst = 0;
  a.zstep = 1;
  if (DIM == 3) {
    const ZPlan z = z_plan(c, (long long)grid.x * grid.y, part != 0);
    if (part != 0 && !z.splittable)
      return cudaErrorInvalidValue;
    a.zchunk = z.zchunk;
    grid.z = z.nchunk;
    if (part == 1) {
      a.zstep = z.nchunk - 1;
      grid.z = 2;
    } else if (part == 2) {
      a.zfirst = 1;
      grid.z = z.nchunk - 2;
    }
  }
  kern<<<grid, block, smem, stream>>>(a);
  c.launches++;
  return cudaGetLastError();
}

template <int DIM, int RS>
static cudaError_t launch_grav(Ctx& c, const double* Uin, double* Uout, int part, cudaStream_t stream)
{
  if (c.p.gravity_enabled)
    return launch_one<DIM, RS, true>(c, Uin, Uout, part, stream);
  return launch_one<DIM, RS, false>(c, Uin, Uout, part, stream);
}

cudaError_t launch_hydro_step(Ctx& c, const double* Uin, double* Uout, int part, cudaStream_t stream)
{
  if (!stream)
    stream = c.stream;
  if (c.p.dim == 3) {
    switch (c.p.riemann) {
      case RIEMANN_HLLC: return launch_grav<3, RIEMANN_HLLC>(c, Uin, Uout, part, stream);
      case RIEMANN_APPROX: return launch_grav<3, RIEMANN_APPROX>(c, Uin, Uout, part, stream);
      case RIEMANN_LLF: return launch_grav<3, RIEMANN_LLF>(c, Uin, Uout, part, stream);
      default: return cudaErrorInvalidValue;
    }
  }
  switch (c.p.riemann) {
    case RIEMANN_HLLC: return launch_grav<2, RIEMANN_HLLC>(c, Uin, Uout, part, stream);
    case RIEMANN_APPROX: return launch_grav<2, RIEMANN_APPROX>(c, Uin, Uout, part, stream);
    case RIEMANN_LLF: return launch_grav<2, RIEMANN_LLF>(c, Uin, Uout, part, stream);
    case RIEMANN_HLL: return launch_grav<2, RIEMANN_HLL>(c, Uin, Uout, part, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_hydro_invdt(Ctx& c, const double* U)
{
  const long long ncell = (long long)c.g.nx * c.g.ny * c.g.nz;
  long long nb = (ncell + 255) / 256;
  if (nb > (long long)c.sm_count * 8)
    nb = (long long)c.sm_count * 8;
  const bool grav = c.p.gravity_enabled != 0;
  if (c.p.dim == 3) {
    if (grav)
      hydro_invdt_kernel<3, true><<<(int)nb, 256, 0, c.stream>>>(c.g, U, c.grav, &c.ts->invdt_bits);
    else
      hydro_invdt_kernel<3, false><<<(int)nb, 256, 0, c.stream>>>(c.g, U, c.grav, &c.ts->invdt_bits);
  } else {
    if (grav)
      hydro_invdt_kernel<2, true><<<(int)nb, 256, 0, c.stream>>>(c.g, U, c.grav, &c.ts->invdt_bits);
    else
      hydro_invdt_kernel<2, false><<<(int)nb, 256, 0, c.stream>>>(c.g, U, c.grav, &c.ts->invdt_bits);
  }
  c.launches++;
  return cudaGetLastError();
}

} // namespace ek
