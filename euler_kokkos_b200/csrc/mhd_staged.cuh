// Stage R of the CT-MHD step (3 HLLD face fluxes + 3 two-dimensional HLLD edge EMFs per cell) as a
// software pipeline: while a warp solves one Riemann problem out of registers, the compact-state
// values of its NEXT problem are already in flight (cp.async, LDGSTS) towards a private staging
// column in shared memory.  ncu on the plain version (mhd_faces_edges_kernel): 33 % of the warp
// stall samples are long-scoreboard waits on the ~260 global loads of a thread, with only 16 warps
// per SM (128 registers) to hide them; DRAM and L2 throughput are far from their limits.
//
// Staging area: [warp][slot][lane] doubles, one 32-lane row (256 B) per staged value.  A lane only
// ever reads the column it filled itself, so the pipeline needs no barrier of any kind, only
// cp.async.wait_group.  The physics is the code of mhd_cells.cuh / mhd_physics.cuh, reading through
// a "view" (slot map) instead of a tile pointer.
#pragma once
#include "mhd_cells.cuh"

namespace ek {

// An edge problem needs 4 corners x 15 values + 8 face-field values = 68 rows (17 KB per warp: 12 warps per
// SM).  EK_MHD_STAGED_SPLIT = 1 stages the own cell's corner (RR) late, into the rows of the first corner once
// that one has been formed: 53 rows (13.6 KB per warp) => 16 warps per SM in 128 registers.
#ifndef EK_MHD_STAGED_SPLIT
#define EK_MHD_STAGED_SPLIT 1
#endif
constexpr int STG_FACE_SLOTS = 31;  // 2 x (8 cell values + 7 slopes) + the face field
constexpr int STG_EDGE_SLOTS = EK_MHD_STAGED_SPLIT ? 53 : 68;
constexpr int STG_SLOTS = STG_EDGE_SLOTS;
constexpr int STG_FACEVALS = EK_MHD_STAGED_SPLIT ? 45 : 60; // first row of the 8 face-field values
constexpr int STG_RR = EK_MHD_STAGED_SPLIT ? 0 : 45;        // first row of the own cell's corner

__device__ __forceinline__ void stg_copy8(double* smem_dst, const double* gsrc)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void stg_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void stg_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- face problem: slots BASE+[0,8) = cell values, BASE+[8,15) = slope group of direction D ------
template <int D>
struct FaceView {
  static constexpr int SD = (D == 0) ? SX : (D == 1) ? SY : SZ;
  const double* col;
  __device__ __forceinline__ double operator()(int var) const { return col[32 * (var < 8 ? var : 8 + var - SD)]; }
};

template <int D>
__device__ __forceinline__ void stg_issue_face(double* col, const double* __restrict__ Sm, const double* __restrict__ Sc)
{
  constexpr int SD = FaceView<D>::SD;
#pragma unroll
  for (int v = 0; v < 8; ++v) {
    if (v == IA + D)
      continue; // the cell-centred normal field is not part of a face state
    stg_copy8(col + 32 * v, Sm + TL * (SC + v));
    stg_copy8(col + 32 * (15 + v), Sc + TL * (SC + v));
  }
#pragma unroll
  for (int v = 0; v < 7; ++v) {
    stg_copy8(col + 32 * (8 + v), Sm + TL * (SD + v));
    stg_copy8(col + 32 * (23 + v), Sc + TL * (SD + v));
  }
  stg_copy8(col + 32 * 30, Sc + TL * (SFA + D));
  stg_commit();
}

struct FaceProblem {
  MhdSide L, R;
  double a;
};

template <int D>
__device__ __forceinline__ FaceProblem stg_form_face(const Grid& g, const double* col)
{
  FaceProblem p;
  p.L = mhd_face_side_v<D>(g, FaceView<D>{ col }, +1.0);
  p.R = mhd_face_side_v<D>(g, FaceView<D>{ col + 32 * 15 }, -1.0);
  p.a = col[32 * 30];
  return p;
}

// The Riemann solvers are real (non-inlined) functions in this kernel: one copy of the HLLD flux and one of
// the 2-D HLLD EMF instead of three each keeps the hot code inside the instruction cache (ncu on the fully
// inlined version: 72 KB of SASS, "no instruction" = 1.5 stall cycles per issued instruction).
#ifndef EK_MHD_NOINLINE
#define EK_MHD_NOINLINE 1
#endif
#if EK_MHD_NOINLINE
#define STG_SOLVER __device__ __noinline__
#else
#define STG_SOLVER __device__ __forceinline__
#endif

// solve and store the flux through a lower face; Fn/Ft1/Ft2: where the normal / transverse momentum fluxes go
template <int RS>
STG_SOLVER void stg_face_solver(double gamma0, double entho, MhdSide L, MhdSide R, double a, double* __restrict__ Fd, int on, int ot1, int ot2)
{
  Grid g{};
  g.gamma0 = gamma0;
  g.entho = entho;
  double f[5];
  mhd_riemann<RS>(g, L, R, a, f);
  Fd[0] = f[0];
  Fd[TL] = f[1];
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
  stg_face_solver<RS>(g.gamma0, g.entho, p.L, p.R, p.a, Fc + TL * 5 * D, TL * on, TL * ot1, TL * ot2);
}

// ---- edge problem along D3: 15 values per corner cell -------------------------------------------
template <int D3>
struct CornerVars {
  static constexpr int D1 = (D3 + 1) % 3, D2 = (D3 + 2) % 3;
  static constexpr int S1 = (D1 == 0) ? SX : (D1 == 1) ? SY : SZ;
  static constexpr int S2 = (D2 == 0) ? SX : (D2 == 1) ? SY : SZ;
  // same slots as mhd_edge_corner_v
  static constexpr int K1 = (D1 == 0) ? (D3 == 1 ? 5 : 6) : (D1 == 1) ? (D3 == 0 ? 5 : 6) : (D3 == 0 ? 5 : 6);
  static constexpr int K2 = (D2 == 0) ? (D3 == 1 ? 5 : 6) : (D2 == 1) ? (D3 == 0 ? 5 : 6) : (D3 == 0 ? 5 : 6);
  __host__ __device__ static constexpr int var(int i)
  {
    return i == 0 ? SC + ID : i == 1 ? SC + IP : i == 2 ? SC + IU + D1 : i == 3 ? SC + IU + D2 : i == 4 ? SC + IA + D3
         : i == 5 ? S1 + 0 : i == 6 ? S1 + 1 : i == 7 ? S1 + 2 + D1 : i == 8 ? S1 + 2 + D2 : i == 9 ? S1 + K1
         : i == 10 ? S2 + 0 : i == 11 ? S2 + 1 : i == 12 ? S2 + 2 + D1 : i == 13 ? S2 + 2 + D2 : S2 + K2;
  }
  __host__ __device__ static constexpr int slot(int v)
  {
    int s = -1;
    for (int i = 0; i < 15; ++i)
      if (var(i) == v)
        s = i;
    return s;
  }
};

template <int D3>
struct CornerView {
  const double* col;
  __device__ __forceinline__ double operator()(int var) const { return col[32 * CornerVars<D3>::slot(var)]; }
};

// S00 = cell c, S10 = c - stride(D1), S01 = c - stride(D2), S11 = c - both
template <int D3>
__device__ __forceinline__ void stg_issue_edge(double* col, const double* __restrict__ S11, const double* __restrict__ S01,
                                               const double* __restrict__ S10, const double* __restrict__ S00)
{
  using CV = CornerVars<D3>;
  constexpr int D1 = CV::D1, D2 = CV::D2;
  constexpr int F1 = SFA + D1, F2 = SFA + D2;
  constexpr int DF1 = (D1 == 0) ? (D2 == 1 ? SDAY : SDAZ) : (D1 == 1) ? (D2 == 0 ? SDBX : SDBZ) : (D2 == 0 ? SDCX : SDCY);
  constexpr int DF2 = (D2 == 0) ? (D1 == 1 ? SDAY : SDAZ) : (D2 == 1) ? (D1 == 0 ? SDBX : SDBZ) : (D1 == 0 ? SDCX : SDCY);
#pragma unroll
  for (int i = 0; i < 15; ++i) {
    stg_copy8(col + 32 * i, S11 + TL * CV::var(i));
    stg_copy8(col + 32 * (15 + i), S01 + TL * CV::var(i));
    stg_copy8(col + 32 * (30 + i), S10 + TL * CV::var(i));
    if (!EK_MHD_STAGED_SPLIT)
      stg_copy8(col + 32 * (STG_RR + i), S00 + TL * CV::var(i));
  }
  constexpr int FV = STG_FACEVALS;
  stg_copy8(col + 32 * (FV + 0), S01 + TL * F1);
  stg_copy8(col + 32 * (FV + 1), S01 + TL * DF1);
  stg_copy8(col + 32 * (FV + 2), S00 + TL * F1);
  stg_copy8(col + 32 * (FV + 3), S00 + TL * DF1);
  stg_copy8(col + 32 * (FV + 4), S10 + TL * F2);
  stg_copy8(col + 32 * (FV + 5), S10 + TL * DF2);
  stg_copy8(col + 32 * (FV + 6), S00 + TL * F2);
  stg_copy8(col + 32 * (FV + 7), S00 + TL * DF2);
  stg_commit();
}

struct EdgeProblem {
  MhdCorner LL, RL, LR, RR;
};

// corner assignment and face-field reconstruction of mhd_edge_emf
template <int D3>
__device__ __forceinline__ EdgeProblem stg_form_edge(const Grid& g, double* col, const double* __restrict__ S00)
{
  constexpr int FV = STG_FACEVALS;
  const double aLow = col[32 * (FV + 0)] + col[32 * (FV + 1)];
  const double aHigh = col[32 * (FV + 2)] - col[32 * (FV + 3)];
  const double bLow = col[32 * (FV + 4)] + col[32 * (FV + 5)];
  const double bHigh = col[32 * (FV + 6)] - col[32 * (FV + 7)];
  EdgeProblem p;
  p.LL = mhd_edge_corner_v<D3>(g, CornerView<D3>{ col }, +1.0, +1.0, aLow, bLow);
  if (EK_MHD_STAGED_SPLIT) {
    // the rows of LL are free now: fetch the own cell's corner into them while RL and LR are formed
#pragma unroll
    for (int i = 0; i < 15; ++i)
      stg_copy8(col + 32 * (STG_RR + i), S00 + TL * CornerVars<D3>::var(i));
    stg_commit();
  }
  p.RL = mhd_edge_corner_v<D3>(g, CornerView<D3>{ col + 32 * 15 }, -1.0, +1.0, aLow, bHigh);
  p.LR = mhd_edge_corner_v<D3>(g, CornerView<D3>{ col + 32 * 30 }, +1.0, -1.0, aHigh, bLow);
  if (EK_MHD_STAGED_SPLIT)
    stg_wait_all();
  p.RR = mhd_edge_corner_v<D3>(g, CornerView<D3>{ col + 32 * STG_RR }, -1.0, -1.0, aHigh, bHigh);
  return p;
}

#ifndef EK_MHD_STAGED_MINB
#define EK_MHD_STAGED_MINB (EK_MHD_STAGED_SPLIT ? 4 : 3)
#endif
constexpr int STG_THREADS = 128;
constexpr size_t STG_SMEM_BYTES = (size_t)(STG_THREADS / 32) * STG_SLOTS * 32 * sizeof(double);

template <int RS>
__global__ void __launch_bounds__(STG_THREADS, EK_MHD_STAGED_MINB) mhd_faces_edges_staged_kernel(const MhdArgs a)
{
  extern __shared__ __align__(16) double stg_area[];
  if (a.ts->finished)
    return;
  MhdIdx x;
  if (!mhd_cell_of_thread(a, x) || !mhd_stage_r_active(a.g, x))
    return;
  const Grid& g = a.g;
  const MhdWork w = mhd_work_split(a.work, g.sv);
  double* col = stg_area + (threadIdx.x >> 5) * (STG_SLOTS * 32) + (threadIdx.x & 31);
  const double* __restrict__ S = w.S;
  const unsigned c = x.c, sj = x.sj, sk = x.sk;
  const double* __restrict__ S000 = S + tile_off<S_TOTAL>(c);
  const double* __restrict__ S100 = S + tile_off<S_TOTAL>(c - 1);
  const double* __restrict__ S010 = S + tile_off<S_TOTAL>(c - sj);
  const double* __restrict__ S001 = S + tile_off<S_TOTAL>(c - sk);
  double* __restrict__ Fc = w.F + tile_off<15>(c);
  double* __restrict__ Ec = w.EMF + tile_off<3>(c);

  stg_issue_face<0>(col, S100, S000);
  stg_wait_all();
  {
    const FaceProblem p = stg_form_face<0>(g, col);
    stg_issue_face<1>(col, S010, S000);
    stg_solve_face<0, RS>(g, p, Fc);
  }
  stg_wait_all();
  {
    const FaceProblem p = stg_form_face<1>(g, col);
    stg_issue_face<2>(col, S001, S000);
    stg_solve_face<1, RS>(g, p, Fc);
  }
  stg_wait_all();
  {
    const FaceProblem p = stg_form_face<2>(g, col);
    // edge along z: D1 = x, D2 = y
    stg_issue_edge<2>(col, S + tile_off<S_TOTAL>(c - 1 - sj), S010, S100, S000);
    stg_solve_face<2, RS>(g, p, Fc);
  }
  stg_wait_all();
  {
    const EdgeProblem p = stg_form_edge<2>(g, col, S000);
    // edge along x: D1 = y, D2 = z
    stg_issue_edge<0>(col, S + tile_off<S_TOTAL>(c - sj - sk), S001, S010, S000);
    stg_edge_solver(g.gamma0, g.smallc, p.LL, p.RL, p.LR, p.RR, Ec + 2 * TL);
  }
  stg_wait_all();
  {
    const EdgeProblem p = stg_form_edge<0>(g, col, S000);
    // edge along y: D1 = z, D2 = x
    stg_issue_edge<1>(col, S + tile_off<S_TOTAL>(c - sk - 1), S100, S001, S000);
    stg_edge_solver(g.gamma0, g.smallc, p.LL, p.RL, p.LR, p.RR, Ec);
  }
  stg_wait_all();
  {
    const EdgeProblem p = stg_form_edge<1>(g, col, S000);
    stg_edge_solver(g.gamma0, g.smallc, p.LL, p.RL, p.LR, p.RR, Ec + TL);
  }
}

// ---- plain loads, but the solvers as real functions (same instruction-cache argument, no staging) ----
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

template <int RS>
__global__ void __launch_bounds__(128, 4) mhd_faces_edges_call_kernel(const MhdArgs a)
{
  if (a.ts->finished)
    return;
  MhdIdx x;
  if (!mhd_cell_of_thread(a, x) || !mhd_stage_r_active(a.g, x))
    return;
  const Grid& g = a.g;
  const MhdWork w = mhd_work_split(a.work, g.sv);
  const double* __restrict__ S = w.S;
  const unsigned c = x.c, sj = x.sj, sk = x.sk;
  const double* __restrict__ S000 = S + tile_off<S_TOTAL>(c);
  const double* __restrict__ S100 = S + tile_off<S_TOTAL>(c - 1);
  const double* __restrict__ S010 = S + tile_off<S_TOTAL>(c - sj);
  const double* __restrict__ S001 = S + tile_off<S_TOTAL>(c - sk);
  double* __restrict__ Fc = w.F + tile_off<15>(c);
  double* __restrict__ Ec = w.EMF + tile_off<3>(c);
  if (a.r_pf > 0 && x.k + a.r_pf <= g.ksize - g.gw) {
    // the same column a.r_pf planes ahead is taken by a later wave of this y-slab: have its compact state brought
    // from DRAM to L2 now (each S row is touched for the first time by the thread that owns it)
    const double* __restrict__ Sn = S + tile_off<S_TOTAL>(c + (unsigned)a.r_pf * sk);
#pragma unroll
    for (int v = 0; v < S_TOTAL; ++v)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(Sn + TL * v));
  }
  call_face<0, RS>(g, S100, S000, Fc);
  call_face<1, RS>(g, S010, S000, Fc);
  call_face<2, RS>(g, S001, S000, Fc);
  call_edge<2>(g, S + tile_off<S_TOTAL>(c - 1 - sj), S010, S100, S000, Ec + 2 * TL);
  call_edge<0>(g, S + tile_off<S_TOTAL>(c - sj - sk), S001, S010, S000, Ec);
  call_edge<1>(g, S + tile_off<S_TOTAL>(c - sk - 1), S100, S001, S000, Ec + TL);
}

} // namespace ek
