// Stage R of the CT-MHD step: per cell, the 3 HLLD face fluxes at its lower faces and the 3 two-dimensional HLLD edge
// EMFs at its lower edges, states rebuilt on the fly from the compact half-step state S (mhd_cells.cuh).
// The two Riemann solvers are real (non-inlined) functions: one copy of the HLLD flux and one of the 2-D HLLD EMF
// instead of three each keeps the hot code inside the instruction cache (ncu on the fully inlined version: 72 KB of
// SASS, 0.93 stall cycles per issued instruction waiting for instruction fetch, 0.13 with the calls).
// Variants that were measured and lost (round 1, recorded in DESIGN.md section 4.2): a cp.async software pipeline that
// staged the next problem's inputs in shared memory (+25 % instructions, 38.3 vs 29.5 ms), per-line L2 prefetch
// instructions for S some planes ahead (the bulk form below, one instruction per warp, is what pays), warps of a CTA on
// consecutive planes, fully inlined solvers.
#pragma once
#include "mhd_cells.cuh"

namespace ek {

struct FaceProblem {
  MhdSide L, R;
  double a;
};

#define STG_SOLVER __device__ __noinline__

// solve and store the flux through a lower face: mass at Fd[0], energy at Fd[oe], normal / transverse momentum fluxes
// at Fd[on], Fd[ot1], Fd[ot2] (the caller knows the layout: AoSoA tile of the F work array, or a shared-memory plane)
template <int RS>
STG_SOLVER void stg_face_solver(double gamma0, double entho, MhdSide L, MhdSide R, double a, double* __restrict__ Fd, int oe, int on, int ot1,
                                int ot2)
{
  Grid g{};
  g.gamma0 = gamma0;
  g.entho = entho;
  double f[5];
  mhd_riemann<RS>(g, L, R, a, f);
  Fd[0] = f[0];
  Fd[oe] = f[1];
  Fd[on] = f[2];
  Fd[ot1] = f[3];
  Fd[ot2] = f[4];
}

STG_SOLVER void stg_edge_solver(double gamma0, double smallc, MhdCorner LL, MhdCorner RL, MhdCorner LR, MhdCorner RR, double* __restrict__ out)
{
  Grid g{};
  g.gamma0 = gamma0;
  g.smallc = smallc;
  *out = mhd_emf2d_hlld(g, LL, RL, LR, RR);
}

// component order of mhd_face_flux: (mass, energy, mom_x, mom_y, mom_z) at Fc[TL*(5*D + .)]
template <int D, int RS>
__device__ __forceinline__ void stg_solve_face(const Grid& g, const FaceProblem& p, double* __restrict__ Fc)
{
  constexpr int on = (D == 0) ? 2 : (D == 1) ? 3 : 4;  // normal momentum = mom_D
  constexpr int ot1 = (D == 0) ? 3 : (D == 1) ? 2 : 3; // first transverse: y, x, y
  constexpr int ot2 = (D == 0) ? 4 : (D == 1) ? 4 : 2; // second transverse: z, z, x
  stg_face_solver<RS>(g.gamma0, g.entho, p.L, p.R, p.a, Fc + TL * 5 * D, TL, TL * on, TL * ot1, TL * ot2);
}

// ---- the kernel: plain loads, reconstruction, solver calls ---------------------------------------------------------
template <int D, int RS>
__device__ __forceinline__ void call_face(const Grid& g, const double* __restrict__ Sm, const double* __restrict__ Sc, double* __restrict__ Fc)
{
  FaceProblem p;
  p.L = mhd_face_side<D>(g, Sm, +1.0);
  p.R = mhd_face_side<D>(g, Sc, -1.0);
  p.a = Sc[TL * (SFA + D)];
  stg_solve_face<D, RS>(g, p, Fc);
}

template <int D3>
__device__ __forceinline__ void call_edge(const Grid& g, const double* __restrict__ S11, const double* __restrict__ S01,
                                          const double* __restrict__ S10, const double* __restrict__ S00, double* __restrict__ out)
{
  constexpr int D1 = (D3 + 1) % 3, D2 = (D3 + 2) % 3;
  constexpr int F1 = SFA + D1, F2 = SFA + D2;
  constexpr int DF1 = (D1 == 0) ? (D2 == 1 ? SDAY : SDAZ) : (D1 == 1) ? (D2 == 0 ? SDBX : SDBZ) : (D2 == 0 ? SDCX : SDCY);
  constexpr int DF2 = (D2 == 0) ? (D1 == 1 ? SDAY : SDAZ) : (D2 == 1) ? (D1 == 0 ? SDBX : SDBZ) : (D1 == 0 ? SDCX : SDCY);
  const double aLow = S01[TL * F1] + S01[TL * DF1];
  const double aHigh = S00[TL * F1] - S00[TL * DF1];
  const double bLow = S10[TL * F2] + S10[TL * DF2];
  const double bHigh = S00[TL * F2] - S00[TL * DF2];
  const MhdCorner LL = mhd_edge_corner<D3>(g, S11, +1.0, +1.0, aLow, bLow);
  const MhdCorner RL = mhd_edge_corner<D3>(g, S01, -1.0, +1.0, aLow, bHigh);
  const MhdCorner LR = mhd_edge_corner<D3>(g, S10, +1.0, -1.0, aHigh, bLow);
  const MhdCorner RR = mhd_edge_corner<D3>(g, S00, -1.0, -1.0, aHigh, bHigh);
  stg_edge_solver(g.gamma0, g.smallc, LL, RL, LR, RR, out);
}

#ifndef EK_MHD_R_MINB
#define EK_MHD_R_MINB 4 // resident CTAs of 128 threads the register budget is cut for (128 registers)
#endif
template <int RS>
__global__ void __launch_bounds__(128, EK_MHD_R_MINB) mhd_faces_edges_call_kernel(const MhdArgs a)
{
  if (a.ts->finished)
    return;
  MhdIdx x;
  if (!mhd_cell_of_thread(a, x))
    return;
  const Grid& g = a.g;
  const MhdWork w = mhd_work_split(a.work, g.sv);
  const double* __restrict__ S = w.S;
  // L2 prefetch: this warp's lane 0 asks for the tile of S the warp pf_planes further along z will read first (one
  // cp.async.bulk.prefetch.L2 of 9.7 KB: no registers, one instruction per warp), so that the first touch of a line of
  // S is an L2 hit instead of a DRAM access (each of the six load batches of a thread waits for its slowest line)
  if (a.pf_planes >= 0 && (threadIdx.x & 31) == 0 && x.k + a.pf_planes < g.ksize)
    l2_prefetch_bulk(S + (size_t)((x.c + (unsigned)a.pf_planes * x.sk) >> 5) * (S_TOTAL * 32), S_TOTAL * 32 * sizeof(double));
  if (!mhd_stage_r_active(g, x))
    return;
  const unsigned c = x.c, sj = x.sj, sk = x.sk;
  const double* __restrict__ S000 = S + tile_off<S_TOTAL>(c);
  const double* __restrict__ S100 = S + tile_off<S_TOTAL>(c - 1);
  const double* __restrict__ S010 = S + tile_off<S_TOTAL>(c - sj);
  const double* __restrict__ S001 = S + tile_off<S_TOTAL>(c - sk);
  double* __restrict__ Fc = w.F + tile_off<15>(c);
  double* __restrict__ Ec = w.EMF + tile_off<3>(c);
  // consecutive problems share a neighbour cell (x face, z edge: c-1;  z edge, y face, x edge: c-sj;  x edge, z face,
  // y edge: c-sk), so that the rows one problem pulled into L1 are still there for the next: a warp touches ~170 rows
  // of 256 bytes per cell and has ~16 KB of L1 to itself (measured against faces-then-edges: -0.55 ms per step;
  // streaming loads for the once-read diagonal corner: no difference)
  call_face<0, RS>(g, S100, S000, Fc);
  call_edge<2>(g, S + tile_off<S_TOTAL>(c - 1 - sj), S010, S100, S000, Ec + 2 * TL);
  call_face<1, RS>(g, S010, S000, Fc);
  call_edge<0>(g, S + tile_off<S_TOTAL>(c - sj - sk), S001, S010, S000, Ec);
  call_face<2, RS>(g, S001, S000, Fc);
  call_edge<1>(g, S + tile_off<S_TOTAL>(c - sk - 1), S100, S001, S000, Ec + TL);
}

} // namespace ek
