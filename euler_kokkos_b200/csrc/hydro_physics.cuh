// Euler (hydro) per-cell / per-face device functions, FP64, restructured for instruction count:
// one reciprocal per cell in cons->prim and trace, shared reciprocals in the Riemann solvers,
// the Riemann solver selected at compile time.
//
// Semantics follow the reference (citations relative to /root/reference/src):
//   computePrimitives                      muscl/HydroBaseFunctor3D.h:81-107, HydroBaseFunctor2D.h:81-105
//   slope_unsplit_hydro_{2d,3d}_scalar     muscl/HydroBaseFunctor3D.h:413-461
//   trace_unsplit_{2d,3d}_along_dir        muscl/HydroBaseFunctor3D.h:276-389, HydroBaseFunctor2D.h:218-286
//   riemann_approx/llf/hll/hllc, cmpflx    shared/RiemannSolvers.h:21-568
#pragma once
#include "ek_common.cuh"

namespace ek {

template <int DIM>
struct HydroTraits {
  static constexpr int NV = (DIM == 3) ? 5 : 4;
};

// ---- conservative -> primitive ---------------------------------------------------------------
template <int DIM>
EK_HD void hydro_cons_to_prim(const Grid& g, const double* __restrict__ u, double* __restrict__ q,
                                                   double& rd_out)
{
  const double d = dmax(u[ID], g.smallr);
  const double rd = rcp_fast(d);
  const double ux = u[IU] * rd;
  const double uy = u[IV] * rd;
  double eken;
  if constexpr (DIM == 3) {
    const double uz = u[IW] * rd;
    eken = 0.5 * (ux * ux + uy * uy + uz * uz);
    q[IW] = uz;
  } else {
    eken = 0.5 * (ux * ux + uy * uy);
  }
  const double e = u[IP] * rd - eken;
  q[ID] = d;
  q[IP] = dmax((g.gamma0 - 1.0) * d * e, d * g.smallp);
  q[IU] = ux;
  q[IV] = uy;
  rd_out = rd;
}

// inverse-dt contribution of one cell (ComputeDtFunctor{2D,3D}, HydroRunFunctors3D.h:44-82;
// ComputeDtGravityFunctor{2D,3D}, :133-189)
template <int DIM, bool GRAV>
EK_HD double hydro_cell_invdt(const Grid& g, const double* __restrict__ u, const double* __restrict__ gvec)
{
  double q[5], rd;
  hydro_cons_to_prim<DIM>(g, u, q, rd);
  const double c = sqrt_fast(g.gamma0 * q[IP] * rd);
  if constexpr (!GRAV) {
    double s = (c + fabs(q[IU])) * g.invdx + (c + fabs(q[IV])) * g.invdy;
    if constexpr (DIM == 3)
      s += (c + fabs(q[IW])) * g.invdz;
    return s;
  } else {
    double dxm = dmin(g.dx, g.dy);
    if constexpr (DIM == 3)
      dxm = dmin(dxm, g.dz);
    double vel = (c + fabs(q[IU])) + (c + fabs(q[IV]));
    double kk = fabs(gvec[0]) + fabs(gvec[1]);
    if constexpr (DIM == 3) {
      vel += c + fabs(q[IW]);
      kk += fabs(gvec[2]);
    }
    kk *= g.cfl * dxm / (vel * vel);
    kk = dmax(kk, 1e-4);
    vel *= kk / (sqrt(1.0 + 2.0 * kk) - 1.0);
    return vel / dxm;
  }
}

// ---- limited slope ---------------------------------------------------------------------------
EK_HD double limited_slope(double st, double sk, double qm, double q, double qp)
{
  return limited_slope_ref(st, sk, qm, q, qp);
}

// ---- Hancock predictor: source terms of one cell ---------------------------------------------
// s[v] = time derivative of primitive variable v (to be multiplied by dt/dx_d * 0.5 per face)
template <int DIM>
EK_HD void hydro_sources(const Grid& g, const double* __restrict__ q, const double (*dq)[5], double* __restrict__ s)
{
  const double r = q[ID], p = q[IP], u = q[IU], v = q[IV];
  const double rr = rcp_fast(r);
  if constexpr (DIM == 3) {
    const double w = q[IW];
    const double div = dq[0][IU] + dq[1][IV] + dq[2][IW];
    s[ID] = -u * dq[0][ID] - v * dq[1][ID] - w * dq[2][ID] - div * r;
    s[IP] = -u * dq[0][IP] - v * dq[1][IP] - w * dq[2][IP] - div * g.gamma0 * p;
    s[IU] = -u * dq[0][IU] - v * dq[1][IU] - w * dq[2][IU] - dq[0][IP] * rr;
    s[IV] = -u * dq[0][IV] - v * dq[1][IV] - w * dq[2][IV] - dq[1][IP] * rr;
    s[IW] = -u * dq[0][IW] - v * dq[1][IW] - w * dq[2][IW] - dq[2][IP] * rr;
  } else {
    const double div = dq[0][IU] + dq[1][IV];
    s[ID] = -u * dq[0][ID] - v * dq[1][ID] - div * r;
    s[IP] = -u * dq[0][IP] - v * dq[1][IP] - div * g.gamma0 * p;
    s[IU] = -u * dq[0][IU] - v * dq[1][IU] - dq[0][IP] * rr;
    s[IV] = -u * dq[0][IV] - v * dq[1][IV] - dq[1][IP] * rr;
  }
}

// ---- Riemann solvers: q = (rho, p, un, ut1[, ut2]) in the face-normal frame -------------------
template <int DIM>
EK_HD void riemann_hllc(const Grid& g, const double* __restrict__ ql, const double* __restrict__ qr, double* __restrict__ flux)
{
  const double entho = g.entho;
  const double rl = dmax(ql[ID], g.smallr);
  const double pl = dmax(ql[IP], rl * g.smallp);
  const double ul = ql[IU];
  double ecinl = 0.5 * rl * ul * ul;
  ecinl += 0.5 * rl * ql[IV] * ql[IV];
  if constexpr (DIM == 3)
    ecinl += 0.5 * rl * ql[IW] * ql[IW];
  const double etotl = pl * entho + ecinl;

  const double rr = dmax(qr[ID], g.smallr);
  const double pr = dmax(qr[IP], rr * g.smallp);
  const double ur = qr[IU];
  double ecinr = 0.5 * rr * ur * ur;
  ecinr += 0.5 * rr * qr[IV] * qr[IV];
  if constexpr (DIM == 3)
    ecinr += 0.5 * rr * qr[IW] * qr[IW];
  const double etotr = pr * entho + ecinr;

  const double sc2 = g.smallc * g.smallc;
  // max(sqrt(x), sqrt(y)) == sqrt(max(x, y)): one square root per face
  const double cmax = sqrt_fast(dmax(dmax(div_fast(g.gamma0 * pl, rl), div_fast(g.gamma0 * pr, rr)), sc2));
  const double SL = dmin(ul, ur) - cmax;
  const double SR = dmax(ul, ur) + cmax;

  const double rcl = rl * (ul - SL);
  const double rcr = rr * (SR - ur);
  const double irc = rcp_fast(rcr + rcl);
  const double ustar = (rcr * ur + rcl * ul + (pl - pr)) * irc;
  const double pstar = (rcr * pl + rcl * pr + rcl * rcr * (ul - ur)) * irc;

  // sample at x/t = 0: pick the side first, then evaluate only that star state
  const bool left = (ustar > 0.0);
  const double S = left ? SL : SR;
  const double ro_ = left ? rl : rr;
  const double uo_ = left ? ul : ur;
  const double po_ = left ? pl : pr;
  const double eo_ = left ? etotl : etotr;
  const double is = rcp_fast(S - ustar);
  const double rstar = ro_ * (S - uo_) * is;
  const double estar = ((S - uo_) * eo_ - po_ * uo_ + pstar * ustar) * is;

  double ro, uo, po, eo;
  if (SL > 0.0) {
    ro = rl; uo = ul; po = pl; eo = etotl;
  } else if (ustar > 0.0) {
    ro = rstar; uo = ustar; po = pstar; eo = estar;
  } else if (SR > 0.0) {
    ro = rstar; uo = ustar; po = pstar; eo = estar;
  } else {
    ro = rr; uo = ur; po = pr; eo = etotr;
  }
  const double fm = ro * uo;
  flux[ID] = fm;
  flux[IU] = fm * uo + po;
  flux[IP] = (eo + po) * uo;
  const bool up = fm > 0.0;
  flux[IV] = fm * (up ? ql[IV] : qr[IV]);
  if constexpr (DIM == 3)
    flux[IW] = fm * (up ? ql[IW] : qr[IW]);
}

template <int DIM>
EK_HD void hydro_cmpflx(const Grid& g, const double* __restrict__ qg, double* __restrict__ flux)
{
  const double fm = qg[ID] * qg[IU];
  flux[ID] = fm;
  flux[IU] = fm * qg[IU] + qg[IP];
  flux[IV] = fm * qg[IV];
  const double entho = g.entho;
  double ekin = 0.5 * qg[ID] * (qg[IU] * qg[IU] + qg[IV] * qg[IV]);
  if constexpr (DIM == 3) {
    flux[IW] = fm * qg[IW];
    ekin += 0.5 * qg[ID] * (qg[IW] * qg[IW]);
  }
  const double etot = qg[IP] * entho + ekin;
  flux[IP] = qg[IU] * (etot + qg[IP]);
}

// Approximate (two-shock, Newton-Raphson) Riemann solver, N independent problems advanced in lock step so that the
// serial chains of the Newton iterations interleave.  Same algorithm and same convergence rule as the reference
// (shared/RiemannSolvers.h:61-199: at most 10 iterations -- the ini's niter_riemann is ignored there, :94 -- while
// |delp / (pold + smallpp)| > 1e-6), restructured for the FP64 pipe:
//   * the Lagrangian wave speeds come from rsqrt (ww = x * rsqrt(x), 1/ww = rsqrt(x)), so (p - pl)/ww is a product;
//   * ql*qr/(ql+qr) with ql = al/bl, qr = ar/br is (al*ar)/(al*br + ar*bl): one reciprocal instead of three divisions;
//   * the convergence test is |delp| > 1e-6 * (pold + smallpp) (pold >= 0, smallpp > 0): no division;
//   * a problem that has converged keeps its state (select), exactly like leaving the reference's loop.
// Per iteration and problem about 50 FP64-pipe instructions instead of about 280 with IEEE '/' and sqrt().
template <int DIM, int N>
EK_HD void riemann_approx_n(const Grid& g, const double (*ql)[5], const double (*qr)[5], double (*flux)[5])
{
  const double gamma0 = g.gamma0, gamma6 = g.gamma6;
  double rl[N], ul[N], pl[N], rr[N], ur[N], pr[N], cl[N], cr[N], ipl[N], ipr[N], pold[N];
  bool active[N];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    rl[n] = dmax(ql[n][ID], g.smallr); ul[n] = ql[n][IU]; pl[n] = dmax(ql[n][IP], rl[n] * g.smallp);
    rr[n] = dmax(qr[n][ID], g.smallr); ur[n] = qr[n][IU]; pr[n] = dmax(qr[n][IP], rr[n] * g.smallp);
    cl[n] = gamma0 * pl[n] * rl[n];
    cr[n] = gamma0 * pr[n] * rr[n];
    const double wl = sqrt_pos_fast(cl[n]), wr = sqrt_pos_fast(cr[n]);
    pold[n] = dmax0(((wr * pl[n] + wl * pr[n]) + wl * wr * (ul[n] - ur[n])) * rcp_fast(wl + wr));
    ipl[n] = rcp_fast(pl[n]);
    ipr[n] = rcp_fast(pr[n]);
    active[n] = true; // conv = 1 > 1e-6
  }
  for (int iter = 0; iter < 10; ++iter) {
    bool any = false;
#pragma unroll
    for (int n = 0; n < N; ++n)
      any = any || active[n];
    if (!any)
      break;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const double wl2 = cl[n] * (1.0 + gamma6 * (pold[n] - pl[n]) * ipl[n]);
      const double wr2 = cr[n] * (1.0 + gamma6 * (pold[n] - pr[n]) * ipr[n]);
      const double iwl = rsqrt_fast(wl2), iwr = rsqrt_fast(wr2);
      const double wwl = wl2 * iwl, wwr = wr2 * iwr;
      const double al = 2.0 * wwl * wl2, bl = wl2 + cl[n];
      const double ar = 2.0 * wwr * wr2, br = wr2 + cr[n];
      const double usl = ul[n] - (pold[n] - pl[n]) * iwl;
      const double usr = ur[n] + (pold[n] - pr[n]) * iwr;
      const double delp = dmax((al * ar) * rcp_fast(al * br + ar * bl) * (usl - usr), -pold[n]);
      const double pnew = pold[n] + delp;
      const bool was = active[n];
      pold[n] = was ? pnew : pold[n];
      active[n] = was && (fabs(delp) > 1e-6 * (pnew + g.smallpp));
    }
  }
#pragma unroll
  for (int n = 0; n < N; ++n) {
    const double pstar = pold[n];
    const double wl2 = cl[n] * (1.0 + gamma6 * (pstar - pl[n]) * ipl[n]);
    const double wr2 = cr[n] * (1.0 + gamma6 * (pstar - pr[n]) * ipr[n]);
    const double iwl = rsqrt_fast(wl2), iwr = rsqrt_fast(wr2);
    const double ustar = 0.5 * (ul[n] + (pl[n] - pstar) * iwl + ur[n] - (pr[n] - pstar) * iwr);
    const double sgnm = copysign(1.0, ustar);
    const bool left = sgnm > 0.0;
    const double ro = left ? rl[n] : rr[n], uo = left ? ul[n] : ur[n], po = left ? pl[n] : pr[n];
    const double wo2 = left ? wl2 : wr2, iwo = left ? iwl : iwr;
    const double wo = wo2 * iwo;
    const double iro = rcp_fast(ro);
    const double co = dmax(g.smallc, sqrt_fast(fabs(gamma0 * po * iro)));
    const double rstar = dmax(ro * rcp_fast(1.0 + ro * (po - pstar) * (iwo * iwo)), g.smallr);
    const double cstar = dmax(g.smallc, sqrt_fast(fabs(gamma0 * pstar * rcp_fast(rstar))));
    double spout = co - sgnm * uo;
    double spin = cstar - sgnm * ustar;
    const double ushock = wo * iro - sgnm * uo;
    if (pstar >= po) {
      spin = ushock;
      spout = ushock;
    }
    const double scr = dmax(spout - spin, g.smallc + fabs(spout + spin));
    double frac = 0.5 * (1.0 + (spout + spin) * rcp_fast(scr));
    if (frac != frac)
      frac = 0.0;
    else
      frac = frac >= 1.0 ? 1.0 : frac <= 0.0 ? 0.0 : frac;
    double qg[5];
    qg[ID] = frac * rstar + (1.0 - frac) * ro;
    qg[IU] = frac * ustar + (1.0 - frac) * uo;
    qg[IP] = frac * pstar + (1.0 - frac) * po;
    if (spout < 0.0) {
      qg[ID] = ro; qg[IU] = uo; qg[IP] = po;
    }
    if (spin > 0.0) {
      qg[ID] = rstar; qg[IU] = ustar; qg[IP] = pstar;
    }
    qg[IV] = left ? ql[n][IV] : qr[n][IV];
    if constexpr (DIM == 3)
      qg[IW] = left ? ql[n][IW] : qr[n][IW];
    hydro_cmpflx<DIM>(g, qg, flux[n]);
  }
}

template <int DIM>
EK_HD void riemann_approx(const Grid& g, const double* __restrict__ ql, const double* __restrict__ qr, double* __restrict__ flux)
{
  double l[1][5], r[1][5], f[1][5];
#pragma unroll
  for (int v = 0; v < 5; ++v) {
    l[0][v] = ql[v];
    r[0][v] = qr[v];
  }
  riemann_approx_n<DIM, 1>(g, l, r, f);
#pragma unroll
  for (int v = 0; v < 5; ++v)
    flux[v] = f[0][v];
}

// conservative state + physical flux of one side (LLF / HLL)
template <int DIM, bool HLL>
EK_HD void hydro_cons_flux(const double* __restrict__ q, double entho, double* __restrict__ u, double* __restrict__ f)
{
  u[ID] = q[ID];
  double e = q[IP] * entho + 0.5 * q[ID] * q[IU] * q[IU];
  e += 0.5 * q[ID] * q[IV] * q[IV];
  if constexpr (DIM == 3)
    e += 0.5 * q[ID] * q[IW] * q[IW];
  u[IP] = e;
  u[IU] = q[ID] * q[IU];
  u[IV] = q[ID] * q[IV];
  f[ID] = HLL ? u[IU] : u[ID] * q[IU];
  f[IP] = q[IU] * (u[IP] + q[IP]);
  f[IU] = q[IP] + u[IU] * q[IU];
  f[IV] = f[ID] * q[IV];
  if constexpr (DIM == 3) {
    u[IW] = q[ID] * q[IW];
    f[IW] = f[ID] * q[IW];
  }
}

template <int DIM>
EK_HD void riemann_llf(const Grid& g, const double* __restrict__ ql, const double* __restrict__ qr, double* __restrict__ flux)
{
  constexpr int NV = HydroTraits<DIM>::NV;
  const double entho = g.entho;
  const double rl = dmax(ql[ID], g.smallr), ul = ql[IU], pl = dmax(ql[IP], rl * g.smallp);
  const double rr = dmax(qr[ID], g.smallr), ur = qr[IU], pr = dmax(qr[IP], rr * g.smallp);
  const double cl = sqrt(g.gamma0 * pl / rl);
  const double cr = sqrt(g.gamma0 * pr / rr);
  const double cmax = dmax(fabs(ul) + cl, fabs(ur) + cr);
  double uL[5], uR[5], fL[5], fR[5];
  hydro_cons_flux<DIM, false>(ql, entho, uL, fL);
  hydro_cons_flux<DIM, false>(qr, entho, uR, fR);
#pragma unroll
  for (int v = 0; v < NV; ++v)
    flux[v] = 0.5 * (fL[v] + fR[v] - cmax * (uR[v] - uL[v]));
}

template <int DIM>
EK_HD void riemann_hll(const Grid& g, const double* __restrict__ ql, const double* __restrict__ qr, double* __restrict__ flux)
{
  constexpr int NV = HydroTraits<DIM>::NV;
  const double entho = g.entho;
  const double rl = dmax(ql[ID], g.smallr), ul = ql[IU], pl = dmax(ql[IP], rl * g.smallp);
  const double rr = dmax(qr[ID], g.smallr), ur = qr[IU], pr = dmax(qr[IP], rr * g.smallp);
  const double cl = sqrt(g.gamma0 * pl / rl);
  const double cr = sqrt(g.gamma0 * pr / rr);
  const double SL = dmin0(dmin(ul, ur) - dmax(cl, cr));
  const double SR = dmax0(dmax(ul, ur) + dmax(cl, cr));
  double uL[5], uR[5], fL[5], fR[5];
  hydro_cons_flux<DIM, true>(ql, entho, uL, fL);
  hydro_cons_flux<DIM, true>(qr, entho, uR, fR);
  const double inv = 1.0 / (SR - SL);
#pragma unroll
  for (int v = 0; v < NV; ++v)
    flux[v] = (SR * fL[v] - SL * fR[v] + SR * SL * (uR[v] - uL[v])) * inv;
}

template <int DIM, int RS>
EK_HD void riemann_hydro(const Grid& g, const double* __restrict__ ql, const double* __restrict__ qr, double* __restrict__ flux)
{
  if constexpr (RS == RIEMANN_HLLC)
    riemann_hllc<DIM>(g, ql, qr, flux);
  else if constexpr (RS == RIEMANN_APPROX)
    riemann_approx<DIM>(g, ql, qr, flux);
  else if constexpr (RS == RIEMANN_LLF)
    riemann_llf<DIM>(g, ql, qr, flux);
  else
    riemann_hll<DIM>(g, ql, qr, flux);
}

// N independent HLLC problems advanced in lock step (statement by statement) so that the FP64
// dependency chains of the N faces interleave: the solver is a long serial chain (two reciprocal
// Newton sequences and a square root), and a CTA only has 4 warps per scheduler to hide it.
template <int DIM, int N>
EK_HD void riemann_hllc_n(const Grid& g, const double (*ql)[5], const double (*qr)[5], double (*flux)[5])
{
  const double entho = g.entho;
  const double sc2 = g.smallc * g.smallc;
  double rl[N], pl[N], rr[N], pr[N], etotl[N], etotr[N], cmax[N], SL[N], SR[N], ustar[N], pstar[N];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    // the traced states already carry the density floor (trace_unsplit_*: rho = max(smallr, .)), so the
    // reference's max(rho, smallr) here is the identity
    rl[n] = ql[n][ID];
    pl[n] = dmax(ql[n][IP], rl[n] * g.smallp);
    rr[n] = qr[n][ID];
    pr[n] = dmax(qr[n][IP], rr[n] * g.smallp);
  }
#pragma unroll
  for (int n = 0; n < N; ++n) {
    double el = 0.5 * rl[n] * ql[n][IU] * ql[n][IU];
    el += 0.5 * rl[n] * ql[n][IV] * ql[n][IV];
    double er = 0.5 * rr[n] * qr[n][IU] * qr[n][IU];
    er += 0.5 * rr[n] * qr[n][IV] * qr[n][IV];
    if constexpr (DIM == 3) {
      el += 0.5 * rl[n] * ql[n][IW] * ql[n][IW];
      er += 0.5 * rr[n] * qr[n][IW] * qr[n][IW];
    }
    etotl[n] = pl[n] * entho + el;
    etotr[n] = pr[n] * entho + er;
  }
  double c2l[N], c2r[N];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    c2l[n] = div_fast(g.gamma0 * pl[n], rl[n]);
    c2r[n] = div_fast(g.gamma0 * pr[n], rr[n]);
  }
#pragma unroll
  for (int n = 0; n < N; ++n)
    cmax[n] = sqrt_pos_fast(dmax(dmax(c2l[n], c2r[n]), sc2)); // argument >= smallc^2 > 0
  double irc[N], rcl[N], rcr[N];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    SL[n] = dmin(ql[n][IU], qr[n][IU]) - cmax[n];
    SR[n] = dmax(ql[n][IU], qr[n][IU]) + cmax[n];
    rcl[n] = rl[n] * (ql[n][IU] - SL[n]);
    rcr[n] = rr[n] * (SR[n] - qr[n][IU]);
  }
#pragma unroll
  for (int n = 0; n < N; ++n)
    irc[n] = rcp_fast(rcr[n] + rcl[n]);
#pragma unroll
  for (int n = 0; n < N; ++n) {
    const double ul = ql[n][IU], ur = qr[n][IU];
    ustar[n] = (rcr[n] * ur + rcl[n] * ul + (pl[n] - pr[n])) * irc[n];
    pstar[n] = (rcr[n] * pl[n] + rcl[n] * pr[n] + rcl[n] * rcr[n] * (ul - ur)) * irc[n];
  }
  double S[N], ro_[N], uo_[N], po_[N], eo_[N], is[N];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    const bool left = (ustar[n] > 0.0);
    S[n] = left ? SL[n] : SR[n];
    ro_[n] = left ? rl[n] : rr[n];
    uo_[n] = left ? ql[n][IU] : qr[n][IU];
    po_[n] = left ? pl[n] : pr[n];
    eo_[n] = left ? etotl[n] : etotr[n];
  }
#pragma unroll
  for (int n = 0; n < N; ++n)
    is[n] = rcp_fast(S[n] - ustar[n]);
#pragma unroll
  for (int n = 0; n < N; ++n) {
    const double rstar = ro_[n] * (S[n] - uo_[n]) * is[n];
    const double estar = ((S[n] - uo_[n]) * eo_[n] - po_[n] * uo_[n] + pstar[n] * ustar[n]) * is[n];
    // the state sampled at x/t = 0: upwind state if the fan does not contain 0, else the star state
    // Sampling at x/t = 0.  The reference tests SL > 0, ustar > 0, SR > 0 in turn (RiemannSolvers.h:528-553).
    // Because rcl*(ul-SL) >= gamma*pl and rcr*(ur-SL) >= gamma*pr, SL > 0 implies ustar > 0 (and SR <= 0 implies
    // ustar <= 0), so the side picked by the sign of ustar already is the upwind side of the supersonic cases:
    // star state iff SL <= 0 < SR, otherwise the state of that side.
    const bool star = !(SL[n] > 0.0) && (SR[n] > 0.0);
    const double ro = star ? rstar : ro_[n];
    const double uo = star ? ustar[n] : uo_[n];
    const double po = star ? pstar[n] : po_[n];
    const double eo = star ? estar : eo_[n];
    const double fm = ro * uo;
    flux[n][ID] = fm;
    flux[n][IU] = fm * uo + po;
    flux[n][IP] = (eo + po) * uo;
    const bool up = fm > 0.0;
    flux[n][IV] = fm * (up ? ql[n][IV] : qr[n][IV]);
    if constexpr (DIM == 3)
      flux[n][IW] = fm * (up ? ql[n][IW] : qr[n][IW]);
  }
}

template <int DIM, int RS, int N>
EK_HD void riemann_hydro_n(const Grid& g, const double (*ql)[5], const double (*qr)[5], double (*flux)[5])
{
  if constexpr (RS == RIEMANN_HLLC) {
    riemann_hllc_n<DIM, N>(g, ql, qr, flux);
  } else if constexpr (RS == RIEMANN_APPROX) {
    riemann_approx_n<DIM, N>(g, ql, qr, flux);
  } else {
#pragma unroll
    for (int n = 0; n < N; ++n)
      riemann_hydro<DIM, RS>(g, ql[n], qr[n], flux[n]);
  }
}

} // namespace ek
