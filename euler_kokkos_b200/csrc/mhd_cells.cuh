// Per-cell stages of the CT-MHD MUSCL-Hancock step (MHD_Muscl_3D, implementationVersion 0 semantics,
// muscl/SolverMHDMuscl.cpp:463-513), restructured around a COMPACT half-step state instead of the
// reference's 18 stored face/edge states per cell (144 doubles/cell, ComputeTraceFunctor3D_MHD):
//
//   stage P  primitives Q (8/cell) + CFL reduction               (ConvertToPrimitives + ComputeDt)
//   stage E  edge electric field v x B (3/cell)                   (ComputeElecFieldFunctor3D)
//   stage T  Hancock half step -> S (38/cell): cell-centred half-step state (8), halved limited
//            slopes actually used by the reconstruction (3 x 7), half-step FACE fields A',B',C' (3)
//            and their halved transverse slopes (6)              (ComputeMagSlopes + ComputeTrace)
//   stage R  per cell: 3 Riemann fluxes at its lower faces + 3 EMFs at its lower edges, states
//            rebuilt on the fly from S                            (ComputeFluxesAndStore + ComputeEmfAndStore)
//   stage U  conservative update + CT update of the face field   (UpdateFunctor3D_MHD + UpdateEmfFunctor3D)
//
// Exact identities used (bitwise in the reference too): the reference's per-cell "right face"
// values AR/BR/CR and their slopes of cell c are the "left face" values of cell c+1, and the
// 0.5*(x+x) "continuity" averages of compute_emf collapse to x.  Hence the face-centred arrays.
//
// All functions are EK_HD: the CUDA kernels (mhd_kernels.cuh) and the CPU emulation used by the
// tests (tests/emul) call the same code.
#pragma once
#include "mhd_physics.cuh"

namespace ek {

// work-array layout, in units of one cell-array (g.sv doubles)
enum : int {
  WQ = 0,         // primitives, 8
  WE = 8,         // edge electric field Ex,Ey,Ez at the lower edges, 3
  WS = 11,        // compact half-step state, 38
  WF = 49,        // face fluxes Fx,Fy,Fz x (mass, energy, mom_x, mom_y, mom_z), 15
  WEMF = 64,      // EMFs at the lower edges: Ex, Ey, Ez, 3
  W_TOTAL = 67
};
// compact state S
enum : int {
  SC = 0,   // r,p,u,v,w,A,B,C at t^{n+1/2} (index = ID..IC)
  SX = 8,   // halved x slopes of r,p,u,v,w,B,C
  SY = 15,  // halved y slopes of r,p,u,v,w,A,C
  SZ = 22,  // halved z slopes of r,p,u,v,w,A,B
  SFA = 29, SFB = 30, SFC = 31,            // half-step face fields A', B', C' (lower faces)
  SDAY = 32, SDAZ = 33, SDBX = 34, SDBZ = 35, SDCX = 36, SDCY = 37, // halved transverse slopes of the face fields
  S_TOTAL = 38
};

struct MhdWork {
  double* Q;
  double* E;
  double* S;
  double* F;
  double* EMF;
};

EK_HD MhdWork mhd_work_split(double* base, long long sv)
{
  MhdWork w;
  w.Q = base + (long long)WQ * sv;
  w.E = base + (long long)WE * sv;
  w.S = base + (long long)WS * sv;
  w.F = base + (long long)WF * sv;
  w.EMF = base + (long long)WEMF * sv;
  return w;
}

// ---- stage P ------------------------------------------------------------------------------------
// Q(i,j,k) for 0 <= i < isize-1 (the last layer is never written, MHDRunFunctors3D.h:120-122);
// returns the cell's inverse-dt contribution (0 outside the interior).
EK_HD double mhd_stage_prims(const Grid& g, const double* __restrict__ U, double* __restrict__ Q, int i, int j, int k)
{
  if (i >= g.isize - 1 || j >= g.jsize - 1 || k >= g.ksize - 1)
    return 0.0;
  const long long o = i + g.sj * j + g.sk * k;
  double u[8], q[8];
#pragma unroll
  for (int v = 0; v < 8; ++v)
    u[v] = U[o + g.sv * v];
  mhd_cons_to_prim(g, u, U[o + 1 + g.sv * IA], U[o + g.sj + g.sv * IB], U[o + g.sk + g.sv * IC], q);
#pragma unroll
  for (int v = 0; v < 8; ++v)
    Q[o + g.sv * v] = q[v];
  const int gw = g.gw;
  if (i >= gw && i < g.isize - gw && j >= gw && j < g.jsize - gw && k >= gw && k < g.ksize - gw)
    return mhd_cell_invdt(g, q);
  return 0.0;
}

// ---- stage E ------------------------------------------------------------------------------------
// E at the three lower edges of cell (i,j,k), 1 <= i <= isize-2 (MHDRunFunctors3D.h:303-352)
EK_HD void mhd_stage_elec(const Grid& g, const double* __restrict__ U, const double* __restrict__ Q, double* __restrict__ E,
                          int i, int j, int k)
{
  if (i < 1 || j < 1 || k < 1 || i > g.isize - 2 || j > g.jsize - 2 || k > g.ksize - 2)
    return;
  const long long o = i + g.sj * j + g.sk * k, sj = g.sj, sk = g.sk, sv = g.sv;
  const double* __restrict__ Qu = Q + sv * IU;
  const double* __restrict__ Qv = Q + sv * IV;
  const double* __restrict__ Qw = Q + sv * IW;
  const double* __restrict__ UA = U + sv * IA;
  const double* __restrict__ UB = U + sv * IB;
  const double* __restrict__ UC = U + sv * IC;
  double u, v, w, A, B, C;
  v = 0.25 * (Qv[o - sj - sk] + Qv[o - sj] + Qv[o - sk] + Qv[o]);
  w = 0.25 * (Qw[o - sj - sk] + Qw[o - sj] + Qw[o - sk] + Qw[o]);
  B = 0.5 * (UB[o - sk] + UB[o]);
  C = 0.5 * (UC[o - sj] + UC[o]);
  E[o] = v * C - w * B;
  u = 0.25 * (Qu[o - 1 - sk] + Qu[o - 1] + Qu[o - sk] + Qu[o]);
  w = 0.25 * (Qw[o - 1 - sk] + Qw[o - 1] + Qw[o - sk] + Qw[o]);
  A = 0.5 * (UA[o - sk] + UA[o]);
  C = 0.5 * (UC[o - 1] + UC[o]);
  E[o + sv] = w * A - u * C;
  u = 0.25 * (Qu[o - 1 - sj] + Qu[o - 1] + Qu[o - sj] + Qu[o]);
  v = 0.25 * (Qv[o - 1 - sj] + Qv[o - 1] + Qv[o - sj] + Qv[o]);
  A = 0.5 * (UA[o - sj] + UA[o]);
  B = 0.5 * (UB[o - 1] + UB[o]);
  E[o + 2 * sv] = u * B - v * A;
}

// ---- stage T ------------------------------------------------------------------------------------
// Hancock predictor of cell (i,j,k) (trace_unsplit_mhd_3d_simpler, MHDBaseFunctor3D.h:690-914) and of
// its three lower faces.  Cells in [2, isize-3], faces in [2, isize-2] (what stage R consumes).
EK_HD void mhd_stage_trace(const Grid& g, const double* __restrict__ U, const double* __restrict__ Q, const double* __restrict__ E,
                           double* __restrict__ S, double dtdx, double dtdy, double dtdz, int i, int j, int k)
{
  if (i < 2 || j < 2 || k < 2 || i > g.isize - 2 || j > g.jsize - 2 || k > g.ksize - 2)
    return;
  const long long o = i + g.sj * j + g.sk * k, sj = g.sj, sk = g.sk, sv = g.sv;
  const double* __restrict__ UA = U + sv * IA;
  const double* __restrict__ UB = U + sv * IB;
  const double* __restrict__ UC = U + sv * IC;
  const double* __restrict__ Ex = E;
  const double* __restrict__ Ey = E + sv;
  const double* __restrict__ Ez = E + 2 * sv;
  const bool in_i = i <= g.isize - 3, in_j = j <= g.jsize - 3, in_k = k <= g.ksize - 3;
  const double xst = dmin(g.slope_type, 2.0); // slope_unsplit_mhd_3d, MHDBaseFunctor3D.h:600

  const double AL = UA[o], BL = UB[o], CL = UC[o];
  const double ELL = Ex[o], FLL = Ey[o], GLL = Ez[o];
  // face-centred half step: AL + sAL0 etc. (:872-877, :889-894) and halved transverse slopes (:810-822)
  if (in_j && in_k) {
    const double GLR = Ez[o + sj], FLR = Ey[o + sk];
    S[o + sv * SFA] = AL + (+(GLR - GLL) * dtdy * 0.5 - (FLR - FLL) * dtdz * 0.5);
    S[o + sv * SDAY] = 0.5 * mhd_slope(xst, UA[o - sj], AL, UA[o + sj]);
    S[o + sv * SDAZ] = 0.5 * mhd_slope(xst, UA[o - sk], AL, UA[o + sk]);
  }
  if (in_i && in_k) {
    const double GRL = Ez[o + 1], ELR = Ex[o + sk];
    S[o + sv * SFB] = BL + (-(GRL - GLL) * dtdx * 0.5 + (ELR - ELL) * dtdz * 0.5);
    S[o + sv * SDBX] = 0.5 * mhd_slope(xst, UB[o - 1], BL, UB[o + 1]);
    S[o + sv * SDBZ] = 0.5 * mhd_slope(xst, UB[o - sk], BL, UB[o + sk]);
  }
  if (in_i && in_j) {
    const double FRL = Ey[o + 1], ERL = Ex[o + sj];
    S[o + sv * SFC] = CL + (+(FRL - FLL) * dtdx * 0.5 - (ERL - ELL) * dtdy * 0.5);
    S[o + sv * SDCX] = 0.5 * mhd_slope(xst, UC[o - 1], CL, UC[o + 1]);
    S[o + sv * SDCY] = 0.5 * mhd_slope(xst, UC[o - sj], CL, UC[o + sj]);
  }
  if (!(in_i && in_j && in_k))
    return;

  // cell-centred half step
  const double st = g.slope_type;
  const bool use_slopes = (st == 1.0 || st == 2.0); // slope_unsplit_hydro_3d, :379,:470
  double q[8], dx_[8], dy_[8], dz_[8];
#pragma unroll
  for (int v = 0; v < 8; ++v) {
    const double* __restrict__ Qv = Q + sv * v;
    const double qc = Qv[o];
    q[v] = qc;
    // halved limited slopes (the reference halves them first thing in the trace, :765-806)
    dx_[v] = (use_slopes && v != IA) ? 0.5 * mhd_slope(st, Qv[o - 1], qc, Qv[o + 1]) : 0.0;
    dy_[v] = (use_slopes && v != IB) ? 0.5 * mhd_slope(st, Qv[o - sj], qc, Qv[o + sj]) : 0.0;
    dz_[v] = (use_slopes && v != IC) ? 0.5 * mhd_slope(st, Qv[o - sk], qc, Qv[o + sk]) : 0.0;
  }
  const double AR = UA[o + 1], BR = UB[o + sj], CR = UC[o + sk];
  const double r = q[ID], p = q[IP], u = q[IU], v = q[IV], w = q[IW], A = q[IA], B = q[IB], C = q[IC];
  const double drx = dx_[ID], dpx = dx_[IP], dux = dx_[IU], dvx = dx_[IV], dwx = dx_[IW], dBx = dx_[IB], dCx = dx_[IC];
  const double dry = dy_[ID], dpy = dy_[IP], duy = dy_[IU], dvy = dy_[IV], dwy = dy_[IW], dAy = dy_[IA], dCy = dy_[IC];
  const double drz = dz_[ID], dpz = dz_[IP], duz = dz_[IU], dvz = dz_[IV], dwz = dz_[IW], dAz = dz_[IA], dBz = dz_[IB];
  const double dAx = 0.5 * (AR - AL), dBy = 0.5 * (BR - BL), dCz = 0.5 * (CR - CL);
  const double ir = rcp_fast(r);
  const double gamma = g.gamma0;

  const double sr0 = (-u * drx - dux * r) * dtdx + (-v * dry - dvy * r) * dtdy + (-w * drz - dwz * r) * dtdz;
  const double su0 = (-u * dux - (dpx + B * dBx + C * dCx) * ir) * dtdx + (-v * duy + B * dAy * ir) * dtdy + (-w * duz + C * dAz * ir) * dtdz;
  const double sv0 = (-u * dvx + A * dBx * ir) * dtdx + (-v * dvy - (dpy + A * dAy + C * dCy) * ir) * dtdy + (-w * dvz + C * dBz * ir) * dtdz;
  const double sw0 = (-u * dwx + A * dCx * ir) * dtdx + (-v * dwy + B * dCy * ir) * dtdy + (-w * dwz - (dpz + A * dAz + B * dBz) * ir) * dtdz;
  const double sp0 = (-u * dpx - dux * gamma * p) * dtdx + (-v * dpy - dvy * gamma * p) * dtdy + (-w * dpz - dwz * gamma * p) * dtdz;
  const double sA0 = (u * dBy + B * duy - v * dAy - A * dvy) * dtdy + (u * dCz + C * duz - w * dAz - A * dwz) * dtdz;
  const double sB0 = (v * dAx + A * dvx - u * dBx - B * dux) * dtdx + (v * dCz + C * dvz - w * dBz - B * dwz) * dtdz;
  const double sC0 = (w * dAx + A * dwx - u * dCx - C * dux) * dtdx + (w * dBy + B * dwy - v * dCy - C * dvy) * dtdy;

  double* __restrict__ Sc = S + o;
  Sc[sv * (SC + ID)] = r + sr0;
  Sc[sv * (SC + IP)] = p + sp0;
  Sc[sv * (SC + IU)] = u + su0;
  Sc[sv * (SC + IV)] = v + sv0;
  Sc[sv * (SC + IW)] = w + sw0;
  Sc[sv * (SC + IA)] = A + sA0;
  Sc[sv * (SC + IB)] = B + sB0;
  Sc[sv * (SC + IC)] = C + sC0;
  Sc[sv * (SX + 0)] = drx; Sc[sv * (SX + 1)] = dpx; Sc[sv * (SX + 2)] = dux; Sc[sv * (SX + 3)] = dvx; Sc[sv * (SX + 4)] = dwx;
  Sc[sv * (SX + 5)] = dBx; Sc[sv * (SX + 6)] = dCx;
  Sc[sv * (SY + 0)] = dry; Sc[sv * (SY + 1)] = dpy; Sc[sv * (SY + 2)] = duy; Sc[sv * (SY + 3)] = dvy; Sc[sv * (SY + 4)] = dwy;
  Sc[sv * (SY + 5)] = dAy; Sc[sv * (SY + 6)] = dCy;
  Sc[sv * (SZ + 0)] = drz; Sc[sv * (SZ + 1)] = dpz; Sc[sv * (SZ + 2)] = duz; Sc[sv * (SZ + 3)] = dvz; Sc[sv * (SZ + 4)] = dwz;
  Sc[sv * (SZ + 5)] = dAz; Sc[sv * (SZ + 6)] = dBz;
}

// ---- stage R ------------------------------------------------------------------------------------
// one side of a face: cell state +/- halved slope along the face normal, floors as :896-975
// D: normal direction; sgn = +1 (upper face of the cell = left state of the face), -1 (lower face)
template <int D>
EK_HD MhdSide mhd_face_side(const Grid& g, const double* __restrict__ S, long long o, double sgn)
{
  const long long sv = g.sv;
  constexpr int SD = (D == 0) ? SX : (D == 1) ? SY : SZ;
  // velocity / field components in the face-normal frame (the reference swaps IU<->IV, IA<->IB for y
  // and IU<->IW, IA<->IC for z, MHDRunFunctors3D.h:1857-1890)
  constexpr int VN = (D == 0) ? IU : (D == 1) ? IV : IW;
  constexpr int VT1 = (D == 0) ? IV : (D == 1) ? IU : IV;
  constexpr int VT2 = (D == 0) ? IW : (D == 1) ? IW : IU;
  constexpr int BT1 = (D == 0) ? IB : (D == 1) ? IA : IB;
  constexpr int BT2 = (D == 0) ? IC : (D == 1) ? IC : IA;
  // slot of the two transverse-field slopes inside the 7-slope group of direction D:
  //   x: (.., dBx, dCx) -> B at 5, C at 6;  y: (.., dAy, dCy) -> A at 5, C at 6;  z: (.., dAz, dBz) -> A at 5, B at 6
  constexpr int KT1 = (D == 2) ? 6 : 5; // z: bt1 = B (slot 6);  x: bt1 = B (slot 5);  y: bt1 = A (slot 5)
  constexpr int KT2 = (D == 2) ? 5 : 6; // z: bt2 = A (slot 5);  x: bt2 = C (slot 6);  y: bt2 = C (slot 6)
  const double* __restrict__ Sc = S + o;
  MhdSide s;
  s.r = dmax(g.smallr, Sc[sv * (SC + ID)] + sgn * Sc[sv * (SD + 0)]);
  s.p = dmax(g.smallp, Sc[sv * (SC + IP)] + sgn * Sc[sv * (SD + 1)]);
  s.un = Sc[sv * (SC + VN)] + sgn * Sc[sv * (SD + 2 + (VN - IU))];
  s.ut1 = Sc[sv * (SC + VT1)] + sgn * Sc[sv * (SD + 2 + (VT1 - IU))];
  s.ut2 = Sc[sv * (SC + VT2)] + sgn * Sc[sv * (SD + 2 + (VT2 - IU))];
  s.bt1 = Sc[sv * (SC + BT1)] + sgn * Sc[sv * (SD + KT1)];
  s.bt2 = Sc[sv * (SC + BT2)] + sgn * Sc[sv * (SD + KT2)];
  return s;
}

// flux through the lower D-face of cell o, stored as (mass, energy, mom_x, mom_y, mom_z)
template <int D, int RS>
EK_HD void mhd_face_flux(const Grid& g, const double* __restrict__ S, long long o, double* __restrict__ F)
{
  const long long sv = g.sv;
  const long long sd = (D == 0) ? 1 : (D == 1) ? g.sj : g.sk;
  const MhdSide L = mhd_face_side<D>(g, S, o - sd, +1.0);
  const MhdSide R = mhd_face_side<D>(g, S, o, -1.0);
  const double a = S[o + sv * (SFA + D)]; // both sides carry the same half-step face field
  double f[5];
  mhd_riemann<RS>(g, L, R, a, f);
  double* __restrict__ Fd = F + o + sv * (5 * D);
  Fd[0] = f[0];
  Fd[sv] = f[1];
  if (D == 0) {
    Fd[2 * sv] = f[2]; Fd[3 * sv] = f[3]; Fd[4 * sv] = f[4];
  } else if (D == 1) {
    Fd[2 * sv] = f[3]; Fd[3 * sv] = f[2]; Fd[4 * sv] = f[4];
  } else {
    Fd[2 * sv] = f[4]; Fd[3 * sv] = f[3]; Fd[4 * sv] = f[2];
  }
}

// one corner state of the edge along D3: cell at o, slopes applied with signs s1 (along D1), s2 (along D2)
template <int D3>
EK_HD MhdCorner mhd_edge_corner(const Grid& g, const double* __restrict__ S, long long o, double s1, double s2, double a, double b)
{
  constexpr int D1 = (D3 + 1) % 3, D2 = (D3 + 2) % 3;
  constexpr int S1 = (D1 == 0) ? SX : (D1 == 1) ? SY : SZ;
  constexpr int S2 = (D2 == 0) ? SX : (D2 == 1) ? SY : SZ;
  // slot of the slope of the orthogonal field B_D3 in the groups of D1 and D2:
  //   group x = (B,C) at (5,6), group y = (A,C) at (5,6), group z = (A,B) at (5,6)
  constexpr int K1 = (D1 == 0) ? (D3 == 1 ? 5 : 6) : (D1 == 1) ? (D3 == 0 ? 5 : 6) : (D3 == 0 ? 5 : 6);
  constexpr int K2 = (D2 == 0) ? (D3 == 1 ? 5 : 6) : (D2 == 1) ? (D3 == 0 ? 5 : 6) : (D3 == 0 ? 5 : 6);
  const long long sv = g.sv;
  const double* __restrict__ Sc = S + o;
  MhdCorner c;
  c.r = dmax(g.smallr, Sc[sv * (SC + ID)] + (s1 * Sc[sv * (S1 + 0)] + s2 * Sc[sv * (S2 + 0)]));
  c.p = dmax(g.smallp, Sc[sv * (SC + IP)] + (s1 * Sc[sv * (S1 + 1)] + s2 * Sc[sv * (S2 + 1)]));
  c.u = Sc[sv * (SC + IU + D1)] + (s1 * Sc[sv * (S1 + 2 + D1)] + s2 * Sc[sv * (S2 + 2 + D1)]);
  c.v = Sc[sv * (SC + IU + D2)] + (s1 * Sc[sv * (S1 + 2 + D2)] + s2 * Sc[sv * (S2 + 2 + D2)]);
  c.c = Sc[sv * (SC + IA + D3)] + (s1 * Sc[sv * (S1 + K1)] + s2 * Sc[sv * (S2 + K2)]);
  c.a = a;
  c.b = b;
  return c;
}

// EMF at the lower edge of cell o that runs along D3 (compute_emf<EMFX|EMFY|EMFZ> with the corner
// assignment of ComputeEmfAndStoreFunctor3D, MHDRunFunctors3D.h:2199-2226)
template <int D3>
EK_HD double mhd_edge_emf(const Grid& g, const double* __restrict__ S, long long o)
{
  constexpr int D1 = (D3 + 1) % 3, D2 = (D3 + 2) % 3;
  const long long sv = g.sv;
  const long long s1 = (D1 == 0) ? 1 : (D1 == 1) ? g.sj : g.sk;
  const long long s2 = (D2 == 0) ? 1 : (D2 == 1) ? g.sj : g.sk;
  // face field B_D1 lives on D1-faces and varies along D2; its transverse slope along D2
  constexpr int F1 = SFA + D1, F2 = SFA + D2;
  constexpr int DF1 = (D1 == 0) ? (D2 == 1 ? SDAY : SDAZ) : (D1 == 1) ? (D2 == 0 ? SDBX : SDBZ) : (D2 == 0 ? SDCX : SDCY);
  constexpr int DF2 = (D2 == 0) ? (D1 == 1 ? SDAY : SDAZ) : (D2 == 1) ? (D1 == 0 ? SDBX : SDBZ) : (D1 == 0 ? SDCX : SDCY);
  const double aLow = S[o - s2 + sv * F1] + S[o - s2 + sv * DF1]; // aLL = aRL
  const double aHigh = S[o + sv * F1] - S[o + sv * DF1];          // aLR = aRR
  const double bLow = S[o - s1 + sv * F2] + S[o - s1 + sv * DF2]; // bLL = bLR
  const double bHigh = S[o + sv * F2] - S[o + sv * DF2];          // bRL = bRR
  const MhdCorner LL = mhd_edge_corner<D3>(g, S, o - s1 - s2, +1.0, +1.0, aLow, bLow);
  const MhdCorner RL = mhd_edge_corner<D3>(g, S, o - s2, -1.0, +1.0, aLow, bHigh);
  const MhdCorner LR = mhd_edge_corner<D3>(g, S, o - s1, +1.0, -1.0, aHigh, bLow);
  const MhdCorner RR = mhd_edge_corner<D3>(g, S, o, -1.0, -1.0, aHigh, bHigh);
  return mhd_emf2d_hlld(g, LL, RL, LR, RR);
}

EK_HD bool mhd_stage_r_active(const Grid& g, int i, int j, int k)
{
  const int gw = g.gw;
  return i >= gw && j >= gw && k >= gw && i <= g.isize - gw && j <= g.jsize - gw && k <= g.ksize - gw;
}

template <int RS>
EK_HD void mhd_stage_fluxes(const Grid& g, const double* __restrict__ S, double* __restrict__ F, int i, int j, int k)
{
  if (!mhd_stage_r_active(g, i, j, k))
    return;
  const long long o = i + g.sj * j + g.sk * k;
  mhd_face_flux<0, RS>(g, S, o, F);
  mhd_face_flux<1, RS>(g, S, o, F);
  mhd_face_flux<2, RS>(g, S, o, F);
}

EK_HD void mhd_stage_emfs(const Grid& g, const double* __restrict__ S, double* __restrict__ EMF, int i, int j, int k)
{
  if (!mhd_stage_r_active(g, i, j, k))
    return;
  const long long o = i + g.sj * j + g.sk * k;
  EMF[o] = mhd_edge_emf<0>(g, S, o);
  EMF[o + g.sv] = mhd_edge_emf<1>(g, S, o);
  EMF[o + 2 * g.sv] = mhd_edge_emf<2>(g, S, o);
}

// ---- stage U ------------------------------------------------------------------------------------
// UpdateFunctor3D_MHD + UpdateEmfFunctor3D on one interior cell, same accumulation order
EK_HD void mhd_stage_update(const Grid& g, const double* __restrict__ Uin, const double* __restrict__ F, const double* __restrict__ EMF,
                            double* __restrict__ Uout, double dtdx, double dtdy, double dtdz, int i, int j, int k)
{
  const int gw = g.gw;
  if (i < gw || j < gw || k < gw || i >= g.isize - gw || j >= g.jsize - gw || k >= g.ksize - gw)
    return;
  const long long o = i + g.sj * j + g.sk * k, sj = g.sj, sk = g.sk, sv = g.sv;
#pragma unroll
  for (int v = 0; v < 5; ++v) {
    double u = Uin[o + sv * v];
    const double* __restrict__ Fx = F + sv * v;
    const double* __restrict__ Fy = F + sv * (5 + v);
    const double* __restrict__ Fz = F + sv * (10 + v);
    u += Fx[o] * dtdx;
    u -= Fx[o + 1] * dtdx;
    u += Fy[o] * dtdy;
    u -= Fy[o + sj] * dtdy;
    u += Fz[o] * dtdz;
    u -= Fz[o + sk] * dtdz;
    Uout[o + sv * v] = u;
  }
  const double* __restrict__ Ex = EMF;
  const double* __restrict__ Ey = EMF + sv;
  const double* __restrict__ Ez = EMF + 2 * sv;
  double A = Uin[o + sv * IA], B = Uin[o + sv * IB], C = Uin[o + sv * IC];
  const double ex = Ex[o], ey = Ey[o], ez = Ez[o];
  A += (Ez[o + sj] - ez) * dtdy;
  B -= (Ez[o + 1] - ez) * dtdx;
  A -= (Ey[o + sk] - ey) * dtdz;
  B += (Ex[o + sk] - ex) * dtdz;
  C += (Ey[o + 1] - ey) * dtdx;
  C -= (Ex[o + sj] - ex) * dtdy;
  Uout[o + sv * IA] = A;
  Uout[o + sv * IB] = B;
  Uout[o + sv * IC] = C;
}

} // namespace ek
