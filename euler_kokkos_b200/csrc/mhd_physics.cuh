// Constrained-transport MHD per-face / per-edge device functions, FP64, written for instruction
// count: one reciprocal per density, shared reciprocals for every repeated denominator, rsqrt-based
// Alfven speeds, only the five hydro components of the HLLD flux (the reference's update never reads
// the magnetic components of the face flux: B is advanced by the EMFs, MHDRunFunctors3D.h:2430-2628).
//
// Semantics follow the reference (citations relative to /root/reference/src):
//   constoprim_mhd                 muscl/MHDBaseFunctor3D.h:194-242
//   find_speed_fast / _info        shared/mhd_utils.h:91-117, 321-366
//   riemann_hlld / hll / llf       shared/RiemannSolvers_MHD.h:29-364
//   mag_riemann2d_hlld             shared/RiemannSolvers_MHD.h:397-627
//   compute_emf<dir>               shared/RiemannSolvers_MHD.h:650-871
//
// Everything here is EK_HD (host + device) so that tests/emul can run the very same code on the CPU
// against the oracle; the product path only ever calls it from the CUDA kernels.
#pragma once
#include "ek_common.cuh"

namespace ek {

EK_HD bool sign_is_neg(double x)
{
#ifdef __CUDA_ARCH__
  return signbit(x);
#else
  return std::signbit(x);
#endif
}

// state of one side of a face in the face-normal frame
struct MhdSide {
  double r, p, un, ut1, ut2, bt1, bt2;
};

// state of one corner of an edge in the edge-normal plane (X = first, Y = second in-plane direction)
struct MhdCorner {
  double r, p, u, v, a, b, c;
};

// limited slope (slope_unsplit_hydro_3d_scalar, MHDBaseFunctor3D.h:320-337; the same formula is used
// for the face-centred field in slope_unsplit_mhd_3d, :563-668)
// HALF of the limited slope (what the trace uses): sth = slope_type/2, sk = 0.5/slope_type
EK_HD double mhd_half_slope(double sth, double sk, double qm, double q, double qp)
{
  return limited_slope_ref(sth, sk, qm, q, qp);
}

// ---- conservative -> primitive (cell-centred B = mean of the two faces) -----------------------
EK_HD void mhd_cons_to_prim(const Grid& g, const double* __restrict__ u, double bxp, double byp, double bzp,
                            double* __restrict__ q)
{
  const double d = dmax(u[ID], g.smallr);
  const double rd = rcp_fast(d);
  const double vx = u[IU] * rd, vy = u[IV] * rd, vz = u[IW] * rd;
  const double A = 0.5 * (u[IA] + bxp), B = 0.5 * (u[IB] + byp), C = 0.5 * (u[IC] + bzp);
  const double eken = 0.5 * (vx * vx + vy * vy + vz * vz);
  const double emag = 0.5 * (A * A + B * B + C * C);
  const double eint = (u[IP] - emag) * rd - eken;
  q[ID] = d;
  q[IP] = dmax((g.gamma0 - 1.0) * d * eint, d * g.smallp);
  q[IU] = vx;
  q[IV] = vy;
  q[IW] = vz;
  q[IA] = A;
  q[IB] = B;
  q[IC] = C;
}

// inverse-dt contribution of one cell from its primitive state (find_speed_info<THREE_D>)
EK_HD double mhd_cell_invdt(const Grid& g, const double* __restrict__ q)
{
  const double rd = rcp_fast(q[ID]);
  const double a = q[IA], b = q[IB], c = q[IC];
  const double b2 = a * a + b * b + c * c;
  const double c2 = g.gamma0 * q[IP] * rd;
  const double d2 = 0.5 * (b2 * rd + c2);
  const double d22 = d2 * d2, c2r = c2 * rd;
  const double cfx = sqrt_pos_fast(d2 + sqrt_fast(dmax0(d22 - c2r * a * a))); // outer argument >= d2 > 0: no zero guard
  const double cfy = sqrt_pos_fast(d2 + sqrt_fast(dmax0(d22 - c2r * b * b)));
  const double cfz = sqrt_pos_fast(d2 + sqrt_fast(dmax0(d22 - c2r * c * c)));
  return (cfx + fabs(q[IU])) * g.invdx + (cfy + fabs(q[IV])) * g.invdy + (cfz + fabs(q[IW])) * g.invdz;
}

// SQUARE of the fast magnetosonic speed along the face normal: bn2 = Bn^2, bt2sum = Bt1^2 + Bt2^2.
// Callers take the max over states first and the square root once (sqrt is monotonic).
EK_HD double mhd_fast_speed2(double gamma0, double r, double p, double bn2, double bt2sum)
{
  const double rd = rcp_fast(r);
  const double c2 = gamma0 * p * rd;
  const double d2 = 0.5 * ((bn2 + bt2sum) * rd + c2);
  return d2 + sqrt_fast(dmax0(d2 * d2 - c2 * bn2 * rd));
}

// ---- HLLD (Miyoshi & Kusano 2005), hydro components of the flux only --------------------------
// a = normal field (already the mean of both sides).  f = (mass, energy, normal, t1, t2 momentum
// [, flux of the second transverse field: only the 2-D solver advances Bz with it, MHDRunFunctors2D.h:1593-1625]).
template <bool BT2>
EK_HD void mhd_hlld(const Grid& g, const MhdSide& L, const MhdSide& R, double a, double* __restrict__ f)
{
  const double entho = g.entho;
  const double sgnm = (a >= 0.0) ? 1.0 : -1.0;
  const double a2 = a * a;

  const double rl = L.r, pl = L.p, ul = L.un, vl = L.ut1, wl = L.ut2, bl = L.bt1, cl = L.bt2;
  const double bt2l = bl * bl + cl * cl;
  const double ecinl = 0.5 * (ul * ul + vl * vl + wl * wl) * rl;
  const double emagl = 0.5 * (a2 + bt2l);
  const double etotl = pl * entho + ecinl + emagl;
  const double ptotl = pl + emagl;
  const double vdotbl = ul * a + vl * bl + wl * cl;

  const double rr = R.r, pr = R.p, ur = R.un, vr = R.ut1, wr = R.ut2, br = R.bt1, cr = R.bt2;
  const double bt2r = br * br + cr * cr;
  const double ecinr = 0.5 * (ur * ur + vr * vr + wr * wr) * rr;
  const double emagr = 0.5 * (a2 + bt2r);
  const double etotr = pr * entho + ecinr + emagr;
  const double ptotr = pr + emagr;
  const double vdotbr = ur * a + vr * br + wr * cr;

  const double cfmax = sqrt_pos_fast(dmax(mhd_fast_speed2(g.gamma0, rl, pl, a2, bt2l), mhd_fast_speed2(g.gamma0, rr, pr, a2, bt2r)));
  const double sl = dmin(ul, ur) - cfmax;
  const double sr = dmax(ul, ur) + cfmax;

  const double rcl = rl * (ul - sl);
  const double rcr = rr * (sr - ur);
  const double irc = rcp_fast(rcr + rcl);
  const double ustar = (rcr * ur + rcl * ul + (ptotl - ptotr)) * irc;
  const double ptotstar = (rcr * ptotl + rcl * ptotr + rcl * rcr * (ul - ur)) * irc;

  // left star region
  const double dsl = sl - ul, dsls = sl - ustar;
  const double idsls = rcp_fast(dsls);
  const double rstarl = rl * dsl * idsls;
  const double estarl = rl * dsl * dsls - a2;
  const double el = rl * dsl * dsl - a2;
  double vstarl = vl, bstarl = bl, wstarl = wl, cstarl = cl;
  if (!(a2 > 0.0 && fabs(estarl - a2) <= 1e-8 * a2)) { // reference: fabs(estar/a^2 - 1) <= 1e-8
    const double ie = rcp_fast(estarl);
    const double k = a * (ustar - ul) * ie, m = el * ie;
    vstarl = vl - k * bl;
    bstarl = bl * m;
    wstarl = wl - k * cl;
    cstarl = cl * m;
  }
  const double vdotbstarl = ustar * a + vstarl * bstarl + wstarl * cstarl;
  const double etotstarl = (dsl * etotl - ptotl * ul + ptotstar * ustar + a * (vdotbl - vdotbstarl)) * idsls;
  const double irsl = rsqrt_fast(rstarl);
  const double sqrrstarl = rstarl * irsl;
  const double sal = ustar - fabs(a) * irsl;

  // right star region
  const double dsr = sr - ur, dsrs = sr - ustar;
  const double idsrs = rcp_fast(dsrs);
  const double rstarr = rr * dsr * idsrs;
  const double estarr = rr * dsr * dsrs - a2;
  const double er = rr * dsr * dsr - a2;
  double vstarr = vr, bstarr = br, wstarr = wr, cstarr = cr;
  if (!(a2 > 0.0 && fabs(estarr - a2) <= 1e-8 * a2)) {
    const double ie = rcp_fast(estarr);
    const double k = a * (ustar - ur) * ie, m = er * ie;
    vstarr = vr - k * br;
    bstarr = br * m;
    wstarr = wr - k * cr;
    cstarr = cr * m;
  }
  const double vdotbstarr = ustar * a + vstarr * bstarr + wstarr * cstarr;
  const double etotstarr = (dsr * etotr - ptotr * ur + ptotstar * ustar + a * (vdotbr - vdotbstarr)) * idsrs;
  const double irsr = rsqrt_fast(rstarr);
  const double sqrrstarr = rstarr * irsr;
  const double sar = ustar + fabs(a) * irsr;

  // double star region
  const double isq = rcp_fast(sqrrstarl + sqrrstarr);
  const double ss = sgnm * sqrrstarl * sqrrstarr;
  const double vstarstar = (sqrrstarl * vstarl + sqrrstarr * vstarr + sgnm * (bstarr - bstarl)) * isq;
  const double wstarstar = (sqrrstarl * wstarl + sqrrstarr * wstarr + sgnm * (cstarr - cstarl)) * isq;
  const double bstarstar = (sqrrstarl * bstarr + sqrrstarr * bstarl + ss * (vstarr - vstarl)) * isq;
  const double cstarstar = (sqrrstarl * cstarr + sqrrstarr * cstarl + ss * (wstarr - wstarl)) * isq;
  const double vdotbstarstar = ustar * a + vstarstar * bstarstar + wstarstar * cstarstar;
  const double etotstarstarl = etotstarl - sgnm * sqrrstarl * (vdotbstarl - vdotbstarstar);
  const double etotstarstarr = etotstarr + sgnm * sqrrstarr * (vdotbstarr - vdotbstarstar);

  // sample the fan at x/t = 0
  double ro, uo, vo, wo, bo, co, ptoto, etoto, vdotbo;
  if (sl > 0.0) {
    ro = rl; uo = ul; vo = vl; wo = wl; bo = bl; co = cl; ptoto = ptotl; etoto = etotl; vdotbo = vdotbl;
  } else if (sal > 0.0) {
    ro = rstarl; uo = ustar; vo = vstarl; wo = wstarl; bo = bstarl; co = cstarl; ptoto = ptotstar; etoto = etotstarl; vdotbo = vdotbstarl;
  } else if (ustar > 0.0) {
    ro = rstarl; uo = ustar; vo = vstarstar; wo = wstarstar; bo = bstarstar; co = cstarstar; ptoto = ptotstar; etoto = etotstarstarl; vdotbo = vdotbstarstar;
  } else if (sar > 0.0) {
    ro = rstarr; uo = ustar; vo = vstarstar; wo = wstarstar; bo = bstarstar; co = cstarstar; ptoto = ptotstar; etoto = etotstarstarr; vdotbo = vdotbstarstar;
  } else if (sr > 0.0) {
    ro = rstarr; uo = ustar; vo = vstarr; wo = wstarr; bo = bstarr; co = cstarr; ptoto = ptotstar; etoto = etotstarr; vdotbo = vdotbstarr;
  } else {
    ro = rr; uo = ur; vo = vr; wo = wr; bo = br; co = cr; ptoto = ptotr; etoto = etotr; vdotbo = vdotbr;
  }
  const double fm = ro * uo;
  f[0] = fm;
  f[1] = (etoto + ptoto) * uo - a * vdotbo;
  f[2] = fm * uo - a2 + ptoto;
  f[3] = fm * vo - a * bo;
  f[4] = fm * wo - a * co;
  if constexpr (BT2)
    f[5] = co * uo - a * wo;
}

// conservative variables and physical flux of one side (find_mhd_flux, mhd_utils.h:175-231), hydro comps
EK_HD void mhd_side_flux(double entho, const MhdSide& s, double a, double* __restrict__ cv, double* __restrict__ ff)
{
  const double ecin = 0.5 * (s.un * s.un + s.ut1 * s.ut1 + s.ut2 * s.ut2) * s.r;
  const double emag = 0.5 * (a * a + s.bt1 * s.bt1 + s.bt2 * s.bt2);
  const double etot = s.p * entho + ecin + emag;
  const double ptot = s.p + emag;
  cv[0] = s.r;
  cv[1] = etot;
  cv[2] = s.r * s.un;
  cv[3] = s.r * s.ut1;
  cv[4] = s.r * s.ut2;
  ff[0] = s.r * s.un;
  ff[1] = (etot + ptot) * s.un - a * (a * s.un + s.bt1 * s.ut1 + s.bt2 * s.ut2);
  ff[2] = s.r * s.un * s.un - a * a + ptot;
  ff[3] = s.r * s.un * s.ut1 - a * s.bt1;
  ff[4] = s.r * s.un * s.ut2 - a * s.bt2;
  cv[5] = s.bt2;
  ff[5] = s.bt2 * s.un - a * s.ut2;
}

// MHD HLL (RiemannSolvers_MHD.h:29-64)
template <bool BT2>
EK_HD void mhd_hll(const Grid& g, const MhdSide& L, const MhdSide& R, double a, double* __restrict__ f)
{
  const double entho = g.entho;
  double ul[6], fl[6], ur[6], fr[6];
  mhd_side_flux(entho, L, a, ul, fl);
  mhd_side_flux(entho, R, a, ur, fr);
  const double a2 = a * a;
  const double cmax = sqrt_pos_fast(dmax(mhd_fast_speed2(g.gamma0, L.r, L.p, a2, L.bt1 * L.bt1 + L.bt2 * L.bt2),
                                     mhd_fast_speed2(g.gamma0, R.r, R.p, a2, R.bt1 * R.bt1 + R.bt2 * R.bt2)));
  const double sl = dmin0(dmin(L.un, R.un) - cmax);
  const double sr = dmax0(dmax(L.un, R.un) + cmax);
  const double inv = rcp_fast(sr - sl);
#pragma unroll
  for (int v = 0; v < (BT2 ? 6 : 5); ++v)
    f[v] = (sr * fl[v] - sl * fr[v] + sr * sl * (ur[v] - ul[v])) * inv;
}

// MHD local Lax-Friedrichs (RiemannSolvers_MHD.h:80-108)
template <bool BT2>
EK_HD void mhd_llf(const Grid& g, const MhdSide& L, const MhdSide& R, double a, double* __restrict__ f)
{
  const double entho = g.entho;
  double ul[6], fl[6], ur[6], fr[6];
  mhd_side_flux(entho, L, a, ul, fl);
  mhd_side_flux(entho, R, a, ur, fr);
  const double a2 = a * a;
  const double cl = sqrt_pos_fast(mhd_fast_speed2(g.gamma0, L.r, L.p, a2, L.bt1 * L.bt1 + L.bt2 * L.bt2)) + fabs(L.un);
  const double cr = sqrt_pos_fast(mhd_fast_speed2(g.gamma0, R.r, R.p, a2, R.bt1 * R.bt1 + R.bt2 * R.bt2)) + fabs(R.un);
  const double vel = dmax(cl, cr);
#pragma unroll
  for (int v = 0; v < (BT2 ? 6 : 5); ++v)
    f[v] = (fl[v] + fr[v]) * 0.5 - vel * (ur[v] - ul[v]) * 0.5;
}

template <int RS, bool BT2 = false>
EK_HD void mhd_riemann(const Grid& g, const MhdSide& L, const MhdSide& R, double a, double* __restrict__ f)
{
  if constexpr (RS == RIEMANN_HLLD)
    mhd_hlld<BT2>(g, L, R, a, f);
  else if constexpr (RS == RIEMANN_HLL)
    mhd_hll<BT2>(g, L, R, a, f);
  else
    mhd_llf<BT2>(g, L, R, a, f);
}

// ---- 2-D HLLD magnetic Riemann solver at an edge ----------------------------------------------
// Corner naming as the reference: first letter = side along X (L = low), second = side along Y.
EK_HD double mhd_emf2d_hlld(const Grid& g, const MhdCorner& LL, const MhdCorner& RL, const MhdCorner& LR, const MhdCorner& RR)
{
  const double gamma0 = g.gamma0;
  // fast speeds along X and Y of the four states + total pressures
  double cfx[4], cfy[4], ptot[4];
  const MhdCorner* S[4] = { &LL, &RL, &LR, &RR };
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    const MhdCorner& s = *S[n];
    const double rd = rcp_fast(s.r);
    const double a2 = s.a * s.a, b2 = s.b * s.b;
    const double bb = a2 + b2 + s.c * s.c;
    const double c2 = gamma0 * s.p * rd;
    const double d2 = 0.5 * (bb * rd + c2);
    const double d22 = d2 * d2, c2r = c2 * rd;
    cfx[n] = d2 + sqrt_fast(dmax0(d22 - c2r * a2)); // squares: the square root is taken after the max
    cfy[n] = d2 + sqrt_fast(dmax0(d22 - c2r * b2));
    ptot[n] = s.p + 0.5 * bb;
  }
  const double cxm = sqrt_pos_fast(dmax(dmax(cfx[0], cfx[1]), dmax(cfx[2], cfx[3]))); // fast speeds squared are > 0
  const double cym = sqrt_pos_fast(dmax(dmax(cfy[0], cfy[1]), dmax(cfy[2], cfy[3])));
  const double SL = dmin(dmin(LL.u, LR.u), dmin(RL.u, RR.u)) - cxm;
  const double SR = dmax(dmax(LL.u, LR.u), dmax(RL.u, RR.u)) + cxm;
  const double SB = dmin(dmin(LL.v, LR.v), dmin(RL.v, RR.v)) - cym;
  const double ST = dmax(dmax(LL.v, LR.v), dmax(RL.v, RR.v)) + cym;
  const double PtotLL = ptot[0], PtotRL = ptot[1], PtotLR = ptot[2], PtotRR = ptot[3];

  const double rcLLx = LL.r * (LL.u - SL), rcRLx = RL.r * (SR - RL.u), rcLRx = LR.r * (LR.u - SL), rcRRx = RR.r * (SR - RR.u);
  const double rcLLy = LL.r * (LL.v - SB), rcLRy = LR.r * (ST - LR.v), rcRLy = RL.r * (RL.v - SB), rcRRy = RR.r * (ST - RR.v);

  const double ustar = (rcLLx * LL.u + rcLRx * LR.u + rcRLx * RL.u + rcRRx * RR.u + (PtotLL - PtotRL + PtotLR - PtotRR)) *
                       rcp_fast(rcLLx + rcLRx + rcRLx + rcRRx);
  const double vstar = (rcLLy * LL.v + rcLRy * LR.v + rcRLy * RL.v + rcRRy * RR.v + (PtotLL - PtotLR + PtotRL - PtotRR)) *
                       rcp_fast(rcLLy + rcLRy + rcRLy + rcRRy);

  const double iSLu = rcp_fast(SL - ustar), iSRu = rcp_fast(SR - ustar);
  const double iSBv = rcp_fast(SB - vstar), iSTv = rcp_fast(ST - vstar);

  // compression factors of each state across its X and Y outer waves
  const double fxLL = (SL - LL.u) * iSLu, fyLL = (SB - LL.v) * iSBv;
  const double fxLR = (SL - LR.u) * iSLu, fyLR = (ST - LR.v) * iSTv;
  const double fxRL = (SR - RL.u) * iSRu, fyRL = (SB - RL.v) * iSBv;
  const double fxRR = (SR - RR.u) * iSRu, fyRR = (ST - RR.v) * iSTv;

  const double BstarLL = LL.b * fxLL, AstarLL = LL.a * fyLL;
  const double BstarLR = LR.b * fxLR, AstarLR = LR.a * fyLR;
  const double BstarRL = RL.b * fxRL, AstarRL = RL.a * fyRL;
  const double BstarRR = RR.b * fxRR, AstarRR = RR.a * fyRR;

#ifndef EK_MHD_ALFVEN_SQ
#define EK_MHD_ALFVEN_SQ 1
#endif
#if !EK_MHD_ALFVEN_SQ
  const double rstarLLx = LL.r * fxLL, rstarLLy = LL.r * fyLL, rstarLL = rstarLLx * fyLL;
  const double rstarLRx = LR.r * fxLR, rstarLRy = LR.r * fyLR, rstarLR = rstarLRx * fyLR;
  const double rstarRLx = RL.r * fxRL, rstarRLy = RL.r * fyRL, rstarRL = rstarRLx * fyRL;
  const double rstarRRx = RR.r * fxRR, rstarRRy = RR.r * fyRR, rstarRR = rstarRRx * fyRR;

  const double isLL = rsqrt_fast(rstarLL), isLR = rsqrt_fast(rstarLR), isRL = rsqrt_fast(rstarRL), isRR = rsqrt_fast(rstarRR);
  const double calfvenL = dmax(dmax(dmax(fabs(LR.a) * rsqrt_fast(rstarLRx), fabs(AstarLR) * isLR),
                                    dmax(fabs(LL.a) * rsqrt_fast(rstarLLx), fabs(AstarLL) * isLL)), g.smallc);
  const double calfvenR = dmax(dmax(dmax(fabs(RR.a) * rsqrt_fast(rstarRRx), fabs(AstarRR) * isRR),
                                    dmax(fabs(RL.a) * rsqrt_fast(rstarRLx), fabs(AstarRL) * isRL)), g.smallc);
  const double calfvenB = dmax(dmax(dmax(fabs(LL.b) * rsqrt_fast(rstarLLy), fabs(BstarLL) * isLL),
                                    dmax(fabs(RL.b) * rsqrt_fast(rstarRLy), fabs(BstarRL) * isRL)), g.smallc);
  const double calfvenT = dmax(dmax(dmax(fabs(LR.b) * rsqrt_fast(rstarLRy), fabs(BstarLR) * isLR),
                                    dmax(fabs(RR.b) * rsqrt_fast(rstarRRy), fabs(BstarRR) * isRR)), g.smallc);

#else
  // Alfven speeds of the four outer states.  The reference takes, per state and direction, the larger of
  //   |a| / sqrt(rho*)   with rho* = r fx            (compressed across the X wave only)   and
  //   |a fy| / sqrt(rho**) with rho** = r fx fy      (compressed across both waves)
  // (RiemannSolvers_MHD.h:480-560).  The second is the first times sqrt(fy) (fy > 0: it is a density ratio), so the pair
  // is |a| sqrt(max(1, fy)) / sqrt(r fx); the maxima are taken over the SQUARES a^2 max(1, fy) / (r fx) and one square
  // root per side follows: 8 reciprocals + 4 square roots instead of 12 reciprocal square roots.
  const double sc2 = g.smallc * g.smallc;
  const double tLLx = LL.a * LL.a * dmax(1.0, fyLL) * rcp_fast(LL.r * fxLL), tLLy = LL.b * LL.b * dmax(1.0, fxLL) * rcp_fast(LL.r * fyLL);
  const double tLRx = LR.a * LR.a * dmax(1.0, fyLR) * rcp_fast(LR.r * fxLR), tLRy = LR.b * LR.b * dmax(1.0, fxLR) * rcp_fast(LR.r * fyLR);
  const double tRLx = RL.a * RL.a * dmax(1.0, fyRL) * rcp_fast(RL.r * fxRL), tRLy = RL.b * RL.b * dmax(1.0, fxRL) * rcp_fast(RL.r * fyRL);
  const double tRRx = RR.a * RR.a * dmax(1.0, fyRR) * rcp_fast(RR.r * fxRR), tRRy = RR.b * RR.b * dmax(1.0, fxRR) * rcp_fast(RR.r * fyRR);
  const double calfvenL = sqrt_pos_fast(dmax(dmax(tLRx, tLLx), sc2));
  const double calfvenR = sqrt_pos_fast(dmax(dmax(tRRx, tRLx), sc2));
  const double calfvenB = sqrt_pos_fast(dmax(dmax(tLLy, tRLy), sc2));
  const double calfvenT = sqrt_pos_fast(dmax(dmax(tLRy, tRRy), sc2));
#endif

  const double SAL = dmin0(ustar - calfvenL), SAR = dmax0(ustar + calfvenR);
  const double SAB = dmin0(vstar - calfvenB), SAT = dmax0(vstar + calfvenT);
  const double iSA_x = rcp_fast(SAR - SAL), iSA_y = rcp_fast(SAT - SAB);

  // the reference blends the five cases with 0/1 weights built from copysign(1, S): the sign BIT
  // decides (+0.0 counts as positive), and exactly one weight product is 1
  const bool SB_pos = !sign_is_neg(SB), ST_pos = !sign_is_neg(ST), SL_pos = !sign_is_neg(SL), SR_pos = !sign_is_neg(SR);
  double E;
  if (SB_pos || !ST_pos) {
    // bottom (SB >= 0) or top (ST < 0) supersonic in Y: 1-D HLLD along X with the lower / upper pair
    const MhdCorner& l = SB_pos ? LL : LR;
    const MhdCorner& r = SB_pos ? RL : RR;
    if (SL_pos) {
      E = l.u * l.b - l.v * l.a;
    } else if (!SR_pos) {
      E = r.u * r.b - r.v * r.a;
    } else {
      const double EstarLx = ustar * (SB_pos ? BstarLL : BstarLR) - l.v * l.a;
      const double EstarRx = ustar * (SB_pos ? BstarRL : BstarRR) - r.v * r.a;
      E = (SAR * EstarLx - SAL * EstarRx + SAR * SAL * (r.b - l.b)) * iSA_x;
    }
  } else if (SL_pos) {
    const double EstarLLy = LL.u * LL.b - vstar * AstarLL, EstarLRy = LR.u * LR.b - vstar * AstarLR;
    E = (SAT * EstarLLy - SAB * EstarLRy - SAT * SAB * (LR.a - LL.a)) * iSA_y;
  } else if (!SR_pos) {
    const double EstarRLy = RL.u * RL.b - vstar * AstarRL, EstarRRy = RR.u * RR.b - vstar * AstarRR;
    E = (SAT * EstarRLy - SAB * EstarRRy - SAT * SAB * (RR.a - RL.a)) * iSA_y;
  } else {
    const double EstarLL = ustar * BstarLL - vstar * AstarLL, EstarLR = ustar * BstarLR - vstar * AstarLR;
    const double EstarRL = ustar * BstarRL - vstar * AstarRL, EstarRR = ustar * BstarRR - vstar * AstarRR;
    const double AstarT = (SAR * AstarRR - SAL * AstarLR) * iSA_x, AstarB = (SAR * AstarRL - SAL * AstarLL) * iSA_x;
    const double BstarR = (SAT * BstarRR - SAB * BstarRL) * iSA_y, BstarL = (SAT * BstarLR - SAB * BstarLL) * iSA_y;
    E = (SAL * SAB * EstarRR - SAL * SAT * EstarRL - SAR * SAB * EstarLR + SAR * SAT * EstarLL) * iSA_x * iSA_y -
        SAT * SAB * iSA_y * (AstarT - AstarB) + SAR * SAL * iSA_x * (BstarR - BstarL);
  }
  return E;
}

} // namespace ek
