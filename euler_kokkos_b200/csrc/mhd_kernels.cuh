// CUDA kernels of the CT-MHD step: thin thread->cell mappings over the per-cell stages of mhd_cells.cuh (3-D: stage U
// and the CFL passes here, stages P+E+T in mhd_pet.cuh, stage R in mhd_stage_r.cuh; 2-D: all five stages here).
// One thread per cell of the ghosted box, flattened index (i fastest => coalesced, no partially filled blocks along x
// for sizes like 518).
#pragma once
#include "mhd_cells2d.cuh"

namespace ek {

struct MhdArgs {
  Grid g;
  const double* __restrict__ Uin;
  double* __restrict__ Uout;
  double* __restrict__ work;
  TimeState* ts;
  MhdChunkMap cm; // thread -> cell order of the 3-D stencil stages (yc = 0: flat)
  unsigned t0 = 0; // offset added to the linear thread index (launches restricted to a range of planes / of y-slabs)
  unsigned t1 = 0xffffffffu; // one past the last linear thread index this launch may take
  int upd_klo = 0, upd_khi = 1 << 30; // planes [upd_klo, upd_khi) the update kernel may write
  int pf_planes = -1; // stage R: every warp asks L2 for the tile of S the warp this many planes further along z will read (-1: off)
};

// cp.async.bulk.prefetch.L2 (sm_90+): one instruction of one lane brings a contiguous block (a whole AoSoA tile) into L2
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, unsigned bytes)
{
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mhd_cell_of_thread(const MhdArgs& a, MhdIdx& x)
{
  const unsigned t = a.t0 + blockIdx.x * blockDim.x + threadIdx.x;
  return t < a.t1 && mhd_chunk_cell(a.g, a.cm, t, x);
}

__device__ __forceinline__ bool mhd_cell_of_thread(const Grid& g, MhdIdx& x)
{
  const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned sj = (unsigned)g.isize, sk = (unsigned)g.sk;
  x.c = c;
  x.sj = sj;
  x.sk = sk;
  if (c >= (unsigned)g.sv)
    return false;
  const unsigned k = c / sk;
  const unsigned r = c - k * sk;
  const unsigned j = r / sj;
  x.k = (int)k;
  x.j = (int)j;
  x.i = (int)(r - j * sj);
  return true;
}

// CFL max-reduction of a ghost-filled level (first step / ek_compute_dt; inside the time loop the update kernel carries
// it): primitives of every interior cell on the fly, warp shuffles, then one atomic per CTA (grid reduce)
__global__ void __launch_bounds__(256) mhd_invdt_kernel(const MhdArgs a)
{
  if (a.ts->finished)
    return;
  MhdIdx x;
  double m = 0.0;
  if (mhd_cell_of_thread(a.g, x)) {
    const int gw = a.g.gw;
    if (x.i >= gw && x.i < a.g.isize - gw && x.j >= gw && x.j < a.g.jsize - gw && x.k >= gw && x.k < a.g.ksize - gw) {
      double q[8];
      mhd_prims_at(a.g, a.Uin, x, q);
      m = mhd_cell_invdt(a.g, q);
    }
  }
  __shared__ double red[8];
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0)
    red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    double r = (threadIdx.x < 8) ? red[threadIdx.x] : 0.0;
    r = warp_max(r);
    if (threadIdx.x == 0 && r > 0.0)
      atomic_max_nonneg(&a.ts->invdt_bits, r);
  }
}

// block maximum of m (256-thread CTAs) into the running max of the inverse dt: warp shuffles, one atomic per CTA
__device__ __forceinline__ void mhd_block_max_invdt(double m, TimeState* ts)
{
  __shared__ double red[8];
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0)
    red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    double r = (threadIdx.x < 8) ? red[threadIdx.x] : 0.0;
    r = warp_max(r);
    if (threadIdx.x == 0 && r > 0.0)
      atomic_max_nonneg(&ts->invdt_bits, r);
  }
}

// stage U.  INVDT: the CFL reduction of the NEXT step rides along (every interior cell except the last layer along
// x, y, z, see mhd_stage_update_t); pre_step_kernel has zeroed ts->invdt_bits before the step's kernels run.
#ifndef EK_MHD_U_MINB
#define EK_MHD_U_MINB 5 // resident CTAs of 256 threads the register budget is cut for (measured: 4 and 6 are slower)
#endif
template <bool INVDT>
__global__ void __launch_bounds__(256, EK_MHD_U_MINB) mhd_update_kernel(const MhdArgs a)
{
  if (a.ts->finished)
    return;
  const MhdWork w = mhd_work_split(a.work, a.g.sv);
  const double dt = a.ts->dt;
  MhdIdx x;
  double m = 0.0;
  if (mhd_cell_of_thread(a, x) && x.k >= a.upd_klo && x.k < a.upd_khi)
    m = mhd_stage_update_t<INVDT>(a.g, a.Uin, w.F, w.EMF, a.Uout, dt / a.g.dx, dt / a.g.dy, dt / a.g.dz, x);
  if constexpr (INVDT)
    mhd_block_max_invdt(m, a.ts);
}

// CFL contribution of the interior cells whose upper faces are ghosts: the last interior layer along x, y and z
// (after the boundary fill / halo exchange of the new state).  Thread t -> cell of layer x (ny*nz cells), then y
// (nx*nz), then z (nx*ny); edge cells are visited more than once, which a maximum does not mind.
__global__ void __launch_bounds__(256) mhd_invdt_layers_kernel(const MhdArgs a)
{
  if (a.ts->finished)
    return;
  const Grid& g = a.g;
  const unsigned nx = (unsigned)g.nx, ny = (unsigned)g.ny, nz = (unsigned)g.nz;
  const unsigned n0 = ny * nz, n1 = nx * nz, n2 = nx * ny;
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  double m = 0.0;
  if (t < n0 + n1 + n2) {
    unsigned i, j, k;
    if (t < n0) {
      i = nx - 1; j = t % ny; k = t / ny;
    } else if (t < n0 + n1) {
      t -= n0;
      i = t % nx; j = ny - 1; k = t / nx;
    } else {
      t -= n0 + n1;
      i = t % nx; j = t / nx; k = nz - 1;
    }
    MhdIdx x;
    x.i = (int)i + g.gw;
    x.j = (int)j + g.gw;
    x.k = (int)k + g.gw;
    x.sj = (unsigned)g.isize;
    x.sk = (unsigned)g.sk;
    x.c = (unsigned)x.i + x.sj * (unsigned)x.j + x.sk * (unsigned)x.k;
    double q[8];
    mhd_prims_at(g, a.Uin, x, q);
    m = mhd_cell_invdt(g, q);
  }
  mhd_block_max_invdt(m, a.ts);
}

// ---- MHD_Muscl_2D: same stages on the 2-D box (one thread per cell, flattened index) --------------
__device__ __forceinline__ bool mhd2d_cell_of_thread(const Grid& g, MhdIdx& x)
{
  const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned sj = (unsigned)g.isize;
  x.c = c;
  x.sj = sj;
  x.sk = 0;
  x.k = 0;
  if (c >= (unsigned)g.sv)
    return false;
  const unsigned j = c / sj;
  x.j = (int)j;
  x.i = (int)(c - j * sj);
  return true;
}

__global__ void __launch_bounds__(256) mhd2d_prims_kernel(const MhdArgs a, const int do_invdt)
{
  if (a.ts->finished)
    return;
  const MhdWork2D w = mhd2d_work_split(a.work, a.g.sv);
  MhdIdx x;
  double m = 0.0;
  if (mhd2d_cell_of_thread(a.g, x))
    m = mhd2d_stage_prims(a.g, a.Uin, w.Q, x);
  if (!do_invdt)
    return;
  __shared__ double red[8];
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0)
    red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    double r = (threadIdx.x < 8) ? red[threadIdx.x] : 0.0;
    r = warp_max(r);
    if (threadIdx.x == 0 && r > 0.0)
      atomic_max_nonneg(&a.ts->invdt_bits, r);
  }
}

__global__ void __launch_bounds__(256) mhd2d_elec_kernel(const MhdArgs a)
{
  if (a.ts->finished)
    return;
  const MhdWork2D w = mhd2d_work_split(a.work, a.g.sv);
  MhdIdx x;
  if (mhd2d_cell_of_thread(a.g, x))
    mhd2d_stage_elec(a.g, a.Uin, w.Q, w.E, x);
}

__global__ void __launch_bounds__(256) mhd2d_trace_kernel(const MhdArgs a)
{
  if (a.ts->finished)
    return;
  const MhdWork2D w = mhd2d_work_split(a.work, a.g.sv);
  const double dt = a.ts->dt;
  MhdIdx x;
  if (mhd2d_cell_of_thread(a.g, x))
    mhd2d_stage_trace(a.g, a.Uin, w.Q, w.E, w.S, dt / a.g.dx, dt / a.g.dy, x);
}

template <int RS>
__global__ void __launch_bounds__(128) mhd2d_faces_edge_kernel(const MhdArgs a)
{
  if (a.ts->finished)
    return;
  const MhdWork2D w = mhd2d_work_split(a.work, a.g.sv);
  MhdIdx x;
  if (mhd2d_cell_of_thread(a.g, x))
    mhd2d_stage_faces_edge<RS>(a.g, w.S, w.F, w.EMF, x);
}

__global__ void __launch_bounds__(256) mhd2d_update_kernel(const MhdArgs a)
{
  if (a.ts->finished)
    return;
  const MhdWork2D w = mhd2d_work_split(a.work, a.g.sv);
  const double dt = a.ts->dt;
  MhdIdx x;
  if (mhd2d_cell_of_thread(a.g, x))
    mhd2d_stage_update(a.g, a.Uin, w.F, w.EMF, a.Uout, dt / a.g.dx, dt / a.g.dy, x);
}

} // namespace ek

#include "mhd_stage_r.cuh"
#include "mhd_ru.cuh"
#include "mhd_pet.cuh"
#include "mhd_pet_tma.cuh"
