// Stages R + U of the CT-MHD step fused in ONE kernel: the face fluxes and edge EMFs never exist in HBM.
//
// Before: mhd_faces_edges_call_kernel wrote 15 flux + 3 EMF values per cell (19.5 GB at 512^3) and mhd_update_kernel
// read them back with their +x / +y / +z neighbours (29 GB) -- 48.5 GB of the step's 155 GB of DRAM traffic and a
// 7.8 ms kernel during which the FP64 pipe idles.  Here a CTA of 32 x 16 threads owns a 31 x 15 column of cells and
// marches along z (the scheme of the hydro kernel, hydro_kernels.cuh):
//   * every thread solves the 3 Riemann problems at the lower faces and the 3 edge problems at the lower edges of its
//     cell in plane k (the code of mhd_stage_r.cuh: same reconstruction, same two non-inlined solvers) and leaves the
//     18 results in its slot of a shared-memory plane;
//   * lane 31 and row 15 are halo producers: they own no updated cell, their results are the +x / +y neighbour values of
//     the last tile column / row (row 15 skips the three problems nobody reads: Fx, Fz, Ey);
//   * after a barrier the cells of plane k-1 are updated from the two planes in shared memory (their own plane k-1
//     values and neighbours, and the z-face flux / Ex / Ey of plane k) -- mhd_update_core, the arithmetic and order of
//     UpdateFunctor3D_MHD + UpdateEmfFunctor3D -- together with the CFL reduction of the next step.
// The redundant work is the halo lanes: 32 x 16 threads for 31 x 15 cells (+ the tile rounding of the grid).
#pragma once
#include "mhd_stage_r.cuh"

namespace ek {

constexpr int RU_TX = 32, RU_TY = 16, RU_NT = RU_TX * RU_TY; // threads; the tile is (RU_TX-1) x (RU_TY-1) cells
constexpr int RU_NC = 18;                                    // Fx[5], Fy[5], Fz[5], Ex, Ey, Ez

struct RuSmem {
  double f[2][RU_NC][RU_NT]; // planes k-1 / k by plane parity: component, thread slot
  double inv[RU_NT];         // running CFL maximum of each thread (kept out of the registers that live across the solver calls)
  double dtd[4];             // dt/dx, dt/dy, dt/dz
};

// flux / EMF sources of mhd_update_core for the cell of thread slot t: plane k-1 in `lo`, plane k in `hi`
struct MhdUpdSmem {
  const double* lo;
  const double* hi;
  int t;
  __device__ __forceinline__ double fx0(int v) const { return lo[(0 + v) * RU_NT + t]; }
  __device__ __forceinline__ double fy0(int v) const { return lo[(5 + v) * RU_NT + t]; }
  __device__ __forceinline__ double fz0(int v) const { return lo[(10 + v) * RU_NT + t]; }
  __device__ __forceinline__ double fxi(int v) const { return lo[(0 + v) * RU_NT + t + 1]; }
  __device__ __forceinline__ double fyj(int v) const { return lo[(5 + v) * RU_NT + t + RU_TX]; }
  __device__ __forceinline__ double fzk(int v) const { return hi[(10 + v) * RU_NT + t]; }
  __device__ __forceinline__ double e0(int c) const { return lo[(15 + c) * RU_NT + t]; }
  __device__ __forceinline__ double ei(int c) const { return lo[(15 + c) * RU_NT + t + 1]; }
  __device__ __forceinline__ double ej(int c) const { return lo[(15 + c) * RU_NT + t + RU_TX]; }
  __device__ __forceinline__ double ek(int c) const { return hi[(15 + c) * RU_NT + t]; }
  __device__ __forceinline__ double eij(int c) const { return lo[(15 + c) * RU_NT + t + RU_TX + 1]; }
  __device__ __forceinline__ double eik(int c) const { return hi[(15 + c) * RU_NT + t + 1]; }
  __device__ __forceinline__ double ejk(int c) const { return hi[(15 + c) * RU_NT + t + RU_TX]; }
};

struct MhdRuArgs {
  MhdArgs a;
  int zchunk; // interior planes per CTA along z
  int zfirst; // z-chunk of blockIdx.z = 0 ...
  int zstep;  // ... and chunk stride per blockIdx.z (one launch for the two end chunks, another for the middle ones)
};

template <int D, int RS>
__device__ __forceinline__ void ru_face(const Grid& g, const double* __restrict__ Sm, const double* __restrict__ Sc, double* __restrict__ slot)
{
  constexpr int on = (D == 0) ? 2 : (D == 1) ? 3 : 4;  // normal momentum = mom_D
  constexpr int ot1 = (D == 0) ? 3 : (D == 1) ? 2 : 3; // first transverse: y, x, y
  constexpr int ot2 = (D == 0) ? 4 : (D == 1) ? 4 : 2; // second transverse: z, z, x
  FaceProblem p;
  p.L = mhd_face_side<D>(g, Sm, +1.0);
  p.R = mhd_face_side<D>(g, Sc, -1.0);
  p.a = Sc[TL * (SFA + D)];
  stg_face_solver<RS>(g.gamma0, g.entho, p.L, p.R, p.a, slot + 5 * D * RU_NT, RU_NT, RU_NT * on, RU_NT * ot1, RU_NT * ot2);
}

// Everything that is cheap to recompute (cell coordinates, flags, dt/dx) is recomputed where it is used instead of being
// carried across the six solver calls of a plane: the calls leave ~100 registers to their caller, and a value that lives
// across them is either in one of those or spilled (first version: 404 bytes of spills per thread).
template <int RS, bool INVDT>
__global__ void __launch_bounds__(RU_NT, 1) mhd_ru_kernel(const MhdRuArgs ra)
{
  extern __shared__ __align__(16) unsigned char ru_raw[];
  RuSmem& sm = *reinterpret_cast<RuSmem*>(ru_raw);
  const MhdArgs& a = ra.a;
  if (a.ts->finished)
    return;
  const Grid& g = a.g;
  const int gw = g.gw;
  const int t = threadIdx.y * RU_TX + threadIdx.x;
  if (t == 0) {
    const double dt = a.ts->dt;
    sm.dtd[0] = dt / g.dx;
    sm.dtd[1] = dt / g.dy;
    sm.dtd[2] = dt / g.dz;
  }
  sm.inv[t] = 0.0;
  const int kbeg = gw + (ra.zfirst + (int)blockIdx.z * ra.zstep) * ra.zchunk; // first plane this CTA updates
  const int kend = min(kbeg + ra.zchunk, g.nz + gw);                          // one past the last
  __syncthreads();

#pragma unroll 1
  for (int k = kbeg; k <= kend; ++k) {
    {
      // an opaque zero: without it the compiler hoists the (k-independent) coordinates, tile pointers, flags and constants
      // out of the loop and then has to carry them across every solver call of every plane
      int z0;
      asm volatile("mov.u32 %0, 0;" : "=r"(z0));
      const int i = gw + (RU_TX - 1) * blockIdx.x + threadIdx.x + z0, j = gw + (RU_TY - 1) * blockIdx.y + threadIdx.y + z0;
      // lower faces / edges of cell (i,j,.) exist for gw <= i <= isize-gw (the last one is the upper boundary face)
      if (i <= g.isize - gw && j <= g.jsize - gw) {
        const bool halo_row = (int)threadIdx.y + z0 == RU_TY - 1; // warp-uniform: read as +y neighbour only (Fy, Ez, Ex)
        const MhdWork w = mhd_work_split(a.work, g.sv);
        const double* __restrict__ S = w.S;
        const unsigned sj = (unsigned)g.isize, sk = (unsigned)g.sk;
        const unsigned c = (unsigned)i + sj * (unsigned)j + sk * (unsigned)k;
        double* __restrict__ slot = &sm.f[k & 1][0][t];
        const double* __restrict__ S000 = S + tile_off<S_TOTAL>(c);
        const double* __restrict__ S100 = S + tile_off<S_TOTAL>(c - 1);
        const double* __restrict__ S010 = S + tile_off<S_TOTAL>(c - sj);
        const double* __restrict__ S001 = S + tile_off<S_TOTAL>(c - sk);
        if (!halo_row)
          ru_face<0, RS>(g, S100, S000, slot);
        ru_face<1, RS>(g, S010, S000, slot);
        if (!halo_row)
          ru_face<2, RS>(g, S001, S000, slot);
        call_edge<2>(g, S + tile_off<S_TOTAL>(c - 1 - sj), S010, S100, S000, slot + 17 * RU_NT);
        call_edge<0>(g, S + tile_off<S_TOTAL>(c - sj - sk), S001, S010, S000, slot + 15 * RU_NT);
        if (!halo_row)
          call_edge<1>(g, S + tile_off<S_TOTAL>(c - sk - 1), S100, S001, S000, slot + 16 * RU_NT);
      }
    }
    __syncthreads(); // plane k complete in shared memory
    if (k > kbeg) {
      int z1;
      asm volatile("mov.u32 %0, 0;" : "=r"(z1));
      MhdIdx x;
      x.i = gw + (RU_TX - 1) * blockIdx.x + threadIdx.x + z1;
      x.j = gw + (RU_TY - 1) * blockIdx.y + threadIdx.y + z1;
      // cells this thread updates: not the halo lane / row, inside the interior
      if (threadIdx.x < RU_TX - 1 && threadIdx.y < RU_TY - 1 && x.i < g.isize - gw && x.j < g.jsize - gw) {
        x.k = k - 1;
        x.sj = (unsigned)g.isize;
        x.sk = (unsigned)g.sk;
        x.c = (unsigned)x.i + x.sj * (unsigned)x.j + x.sk * (unsigned)(k - 1);
        MhdUpdSmem P;
        P.lo = &sm.f[(k - 1) & 1][0][0];
        P.hi = &sm.f[k & 1][0][0];
        P.t = t;
        const double m = mhd_update_core<INVDT>(g, a.Uin, a.Uout, P, sm.dtd[0], sm.dtd[1], sm.dtd[2], x);
        if constexpr (INVDT)
          sm.inv[t] = dmax(sm.inv[t], m);
      }
    }
    __syncthreads(); // plane k-1 is free: the next iteration writes plane k+1 over it
  }

  if constexpr (INVDT) {
    // CFL reduction of the next step: warp shuffles, one atomic per CTA (grid reduce)
    double m = warp_max(sm.inv[t]);
    __syncthreads();
    if (threadIdx.x == 0)
      sm.inv[threadIdx.y] = m;
    __syncthreads();
    if (threadIdx.y == 0) {
      double r = (threadIdx.x < RU_TY) ? sm.inv[threadIdx.x] : 0.0;
      r = warp_max(r);
      if (threadIdx.x == 0 && r > 0.0)
        atomic_max_nonneg(&a.ts->invdt_bits, r);
    }
  }
}

} // namespace ek
