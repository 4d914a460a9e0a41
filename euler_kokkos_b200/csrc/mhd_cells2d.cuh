// Per-cell stages of the 2-D CT-MHD step (MHD_Muscl_2D, implementationVersion 0 semantics,
// muscl/SolverMHDMuscl.cpp:371-454, muscl/MHDRunFunctors2D.h, muscl/MHDBaseFunctor2D.h:776-1109).
// Same staged design as the 3-D path (mhd_cells.cuh): primitives + CFL -> edge Ez -> Hancock half step
// into the compact state S (the 3-D slot numbering is reused, the z slots stay empty) -> per cell 2 face
// fluxes + 1 edge EMF -> update.  Differences of the reference's 2-D code that are kept:
//   * cell-centred Bz = U[IC]/2 (magFieldNeighbors[IZ] = 0, MHDRunFunctors2D.h:138);
//   * the trace floors the pressure with smallp*rho (3-D: smallp) and limits the face-field slopes
//     with slope_type itself (3-D: min(slope_type, 2));
//   * Bz is advanced with the IC component of the face fluxes (UpdateFunctor2D_MHD).
#pragma once
#include "mhd_cells.cuh"

namespace ek {

enum : int {
  W2Q = 0,     // primitives, 8
  W2E = 8,     // Ez at the lower-left corner, 1
  W2S = 9,     // compact half-step state (3-D slot numbering), 38
  W2F = 47,    // face fluxes Fx, Fy x (mass, energy, mom_x, mom_y, mom_z, Bz), 12
  W2EMF = 59,  // EMF_z at the lower-left corner, 1
  W2_TOTAL = 60
};

struct MhdWork2D {
  double* Q;
  double* E;
  double* S;
  double* F;
  double* EMF;
};

EK_HD MhdWork2D mhd2d_work_split(double* base, long long sv)
{
  const long long T = mhd_padded_cells(sv);
  MhdWork2D w;
  w.Q = base + (long long)W2Q * T;
  w.E = base + (long long)W2E * T;
  w.S = base + (long long)W2S * T;
  w.F = base + (long long)W2F * T;
  w.EMF = base + (long long)W2EMF * T;
  return w;
}

// ---- stage P ------------------------------------------------------------------------------------
EK_HD double mhd2d_stage_prims(const Grid& g, const double* __restrict__ U, double* __restrict__ Q, const MhdIdx& x)
{
  if (x.i >= g.isize - 1 || x.j >= g.jsize - 1)
    return 0.0;
  const double* __restrict__ Uc = U + x.c;
  const long long sv = g.sv;
  double u[8], q[8];
#pragma unroll
  for (int v = 0; v < 8; ++v)
    u[v] = Uc[sv * v];
  mhd_cons_to_prim(g, u, Uc[sv * IA + 1], Uc[sv * IB + x.sj], 0.0, q);
  double* __restrict__ Qc = Q + tile_off<8>(x.c);
#pragma unroll
  for (int v = 0; v < 8; ++v)
    Qc[TL * v] = q[v];
  const int gw = g.gw;
  if (x.i >= gw && x.i < g.isize - gw && x.j >= gw && x.j < g.jsize - gw) {
    // find_speed_info<TWO_D> (mhd_utils.h:321-366)
    const double rd = rcp_fast(q[ID]);
    const double a = q[IA], b = q[IB], c = q[IC];
    const double b2 = a * a + b * b + c * c;
    const double c2 = g.gamma0 * q[IP] * rd;
    const double d2 = 0.5 * (b2 * rd + c2);
    const double d22 = d2 * d2, c2r = c2 * rd;
    const double cfx = sqrt_fast(d2 + sqrt_fast(dmax0(d22 - c2r * a * a)));
    const double cfy = sqrt_fast(d2 + sqrt_fast(dmax0(d22 - c2r * b * b)));
    return (cfx + fabs(q[IU])) * g.invdx + (cfy + fabs(q[IV])) * g.invdy;
  }
  return 0.0;
}

// ---- stage E: Ez at the lower-left corner of cell (i,j), 1 <= i <= isize-2 (MHDBaseFunctor2D.h:806-823)
EK_HD void mhd2d_stage_elec(const Grid& g, const double* __restrict__ U, const double* __restrict__ Q, double* __restrict__ E,
                            const MhdIdx& x)
{
  if (x.i < 1 || x.j < 1 || x.i > g.isize - 2 || x.j > g.jsize - 2)
    return;
  const unsigned c = x.c, sj = x.sj;
  const long long sv = g.sv;
  const double* __restrict__ q00 = Q + tile_off<8>(c);
  const double* __restrict__ q10 = Q + tile_off<8>(c - 1);
  const double* __restrict__ q01 = Q + tile_off<8>(c - sj);
  const double* __restrict__ q11 = Q + tile_off<8>(c - 1 - sj);
  const double* __restrict__ UA = U + sv * IA + c;
  const double* __restrict__ UB = U + sv * IB + c;
  const double u = 0.25 * (q11[TL * IU] + q10[TL * IU] + q01[TL * IU] + q00[TL * IU]);
  const double v = 0.25 * (q11[TL * IV] + q10[TL * IV] + q01[TL * IV] + q00[TL * IV]);
  const double A = 0.5 * (UA[-(long long)sj] + UA[0]);
  const double B = 0.5 * (UB[-1] + UB[0]);
  E[tile_off<1>(c)] = u * B - v * A;
}

// ---- stage T: trace_unsplit_mhd_2d, cells in [2, isize-3], faces in [2, isize-2] ------------------
EK_HD void mhd2d_stage_trace(const Grid& g, const double* __restrict__ U, const double* __restrict__ Q, const double* __restrict__ E,
                             double* __restrict__ S, double dtdx, double dtdy, const MhdIdx& x)
{
  if (x.i < 2 || x.j < 2 || x.i > g.isize - 2 || x.j > g.jsize - 2)
    return;
  const unsigned c = x.c;
  const long long sj = x.sj, sv = g.sv;
  const double* __restrict__ UA = U + sv * IA + c;
  const double* __restrict__ UB = U + sv * IB + c;
  const bool in_i = x.i <= g.isize - 3, in_j = x.j <= g.jsize - 3;
  const double st = g.slope_type;
  const double sth = 0.5 * st;
  double* __restrict__ Sc = S + tile_off<S_TOTAL>(c);
  const double AL = UA[0], BL = UB[0];
  const double ELL = E[tile_off<1>(c)];
  if (in_j) {
    const double ELR = E[tile_off<1>(c + x.sj)];
    Sc[TL * SFA] = AL + (+(ELR - ELL) * 0.5 * dtdy);
    Sc[TL * SDAY] = mhd_half_slope(sth, g.slope_k, UA[-sj], AL, UA[sj]);
  }
  if (in_i) {
    const double ERL = E[tile_off<1>(c + 1)];
    Sc[TL * SFB] = BL + (-(ERL - ELL) * 0.5 * dtdx);
    Sc[TL * SDBX] = mhd_half_slope(sth, g.slope_k, UB[-1], BL, UB[1]);
  }
  if (!(in_i && in_j))
    return;

  const bool use_slopes = (st == 1.0 || st == 2.0);
  const double* __restrict__ Q0 = Q + tile_off<8>(c);
  double q[8];
#pragma unroll
  for (int v = 0; v < 8; ++v)
    q[v] = Q0[TL * v];
  const double r = q[ID], p = q[IP], u = q[IU], v = q[IV], w = q[IW], A = q[IA], B = q[IB], C = q[IC];
  const double ir = rcp_fast(r);
  const double gp = g.gamma0 * p;
  double sr0, sp0, su0, sv0, sw0, sA0, sB0, sC0;
  {
    const double* __restrict__ Qm = Q + tile_off<8>(c - 1);
    const double* __restrict__ Qp = Q + tile_off<8>(c + 1);
    double d[8];
#pragma unroll
    for (int n = 0; n < 8; ++n)
      d[n] = (use_slopes && n != IA) ? mhd_half_slope(sth, g.slope_k, Qm[TL * n], q[n], Qp[TL * n]) : 0.0;
    const double drx = d[ID], dpx = d[IP], dux = d[IU], dvx = d[IV], dwx = d[IW], dBx = d[IB], dCx = d[IC];
    const double dAx = 0.5 * (UA[1] - AL);
    sr0 = (-u * drx - dux * r) * dtdx;
    su0 = (-u * dux - dpx * ir - B * dBx * ir - C * dCx * ir) * dtdx;
    sv0 = (-u * dvx + A * dBx * ir) * dtdx;
    sw0 = (-u * dwx + A * dCx * ir) * dtdx;
    sp0 = (-u * dpx - dux * gp) * dtdx;
    sB0 = (-u * dBx - B * dux + v * dAx + A * dvx) * dtdx;
    sC0 = (w * dAx + A * dwx - u * dCx - C * dux) * dtdx;
    Sc[TL * (SX + 0)] = drx; Sc[TL * (SX + 1)] = dpx; Sc[TL * (SX + 2)] = dux; Sc[TL * (SX + 3)] = dvx; Sc[TL * (SX + 4)] = dwx;
    Sc[TL * (SX + 5)] = dBx; Sc[TL * (SX + 6)] = dCx;
  }
  {
    const double* __restrict__ Qm = Q + tile_off<8>(c - x.sj);
    const double* __restrict__ Qp = Q + tile_off<8>(c + x.sj);
    double d[8];
#pragma unroll
    for (int n = 0; n < 8; ++n)
      d[n] = (use_slopes && n != IB) ? mhd_half_slope(sth, g.slope_k, Qm[TL * n], q[n], Qp[TL * n]) : 0.0;
    const double dry = d[ID], dpy = d[IP], duy = d[IU], dvy = d[IV], dwy = d[IW], dAy = d[IA], dCy = d[IC];
    const double dBy = 0.5 * (UB[sj] - BL);
    sr0 += (-v * dry - dvy * r) * dtdy;
    su0 += (-v * duy + B * dAy * ir) * dtdy;
    sv0 += (-v * dvy - dpy * ir - A * dAy * ir - C * dCy * ir) * dtdy;
    sw0 += (-v * dwy + B * dCy * ir) * dtdy;
    sp0 += (-v * dpy - dvy * gp) * dtdy;
    sA0 = (u * dBy + B * duy - v * dAy - A * dvy) * dtdy;
    sC0 += (-v * dCy - C * dvy + w * dBy + B * dwy) * dtdy;
    Sc[TL * (SY + 0)] = dry; Sc[TL * (SY + 1)] = dpy; Sc[TL * (SY + 2)] = duy; Sc[TL * (SY + 3)] = dvy; Sc[TL * (SY + 4)] = dwy;
    Sc[TL * (SY + 5)] = dAy; Sc[TL * (SY + 6)] = dCy;
  }
  Sc[TL * (SC + ID)] = r + sr0;
  Sc[TL * (SC + IP)] = p + sp0;
  Sc[TL * (SC + IU)] = u + su0;
  Sc[TL * (SC + IV)] = v + sv0;
  Sc[TL * (SC + IW)] = w + sw0;
  Sc[TL * (SC + IA)] = A + sA0;
  Sc[TL * (SC + IB)] = B + sB0;
  Sc[TL * (SC + IC)] = C + sC0;
}

// ---- stage R: 2 face fluxes (6 components) + EMF_z at the lower-left corner ----------------------
template <int RS>
EK_HD void mhd2d_stage_faces_edge(const Grid& g, const double* __restrict__ S, double* __restrict__ F, double* __restrict__ EMF,
                                  const MhdIdx& x)
{
  const int gw = g.gw;
  if (x.i < gw || x.j < gw || x.i > g.isize - gw || x.j > g.jsize - gw)
    return;
  const double* __restrict__ Sc = S + tile_off<S_TOTAL>(x.c);
  double* __restrict__ Fc = F + tile_off<12>(x.c);
  double f[6];
  {
    const double* __restrict__ Sm = S + tile_off<S_TOTAL>(x.c - 1);
    const MhdSide L = mhd_face_side<0, true>(g, Sm, +1.0);
    const MhdSide R = mhd_face_side<0, true>(g, Sc, -1.0);
    mhd_riemann<RS, true>(g, L, R, Sc[TL * SFA], f);
#pragma unroll
    for (int v = 0; v < 6; ++v)
      Fc[TL * v] = f[v];
  }
  {
    const double* __restrict__ Sm = S + tile_off<S_TOTAL>(x.c - x.sj);
    const MhdSide L = mhd_face_side<1, true>(g, Sm, +1.0);
    const MhdSide R = mhd_face_side<1, true>(g, Sc, -1.0);
    mhd_riemann<RS, true>(g, L, R, Sc[TL * SFB], f);
    Fc[TL * 6] = f[0];
    Fc[TL * 7] = f[1];
    Fc[TL * 8] = f[3]; // mom_x = first transverse component of the y face
    Fc[TL * 9] = f[2];
    Fc[TL * 10] = f[4];
    Fc[TL * 11] = f[5];
  }
  EMF[tile_off<1>(x.c)] = mhd_edge_emf<2, true>(g, S, x);
}

// ---- stage U: UpdateFunctor2D_MHD + UpdateEmfFunctor2D ---------------------------------------------
EK_HD void mhd2d_stage_update(const Grid& g, const double* __restrict__ Uin, const double* __restrict__ F, const double* __restrict__ EMF,
                              double* __restrict__ Uout, double dtdx, double dtdy, const MhdIdx& x)
{
  const int gw = g.gw;
  if (x.i < gw || x.j < gw || x.i >= g.isize - gw || x.j >= g.jsize - gw)
    return;
  const unsigned c = x.c;
  const long long sv = g.sv;
  const double* __restrict__ F0 = F + tile_off<12>(c);
  const double* __restrict__ Fi = F + tile_off<12>(c + 1);
  const double* __restrict__ Fj = F + tile_off<12>(c + x.sj);
  const double* __restrict__ Uc = Uin + c;
  double* __restrict__ Oc = Uout + c;
  const int var_of[6] = { ID, IP, IU, IV, IW, IC };
#pragma unroll
  for (int n = 0; n < 6; ++n) {
    double u = Uc[sv * var_of[n]];
    u += F0[TL * n] * dtdx;
    u -= Fi[TL * n] * dtdx;
    u += F0[TL * (6 + n)] * dtdy;
    u -= Fj[TL * (6 + n)] * dtdy;
    Oc[sv * var_of[n]] = u;
  }
  const double e0 = EMF[tile_off<1>(c)];
  double A = Uc[sv * IA], B = Uc[sv * IB];
  A += (EMF[tile_off<1>(c + x.sj)] - e0) * dtdy;
  B -= (EMF[tile_off<1>(c + 1)] - e0) * dtdx;
  Oc[sv * IA] = A;
  Oc[sv * IB] = B;
}

} // namespace ek
