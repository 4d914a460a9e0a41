/*
 * TEST INFRASTRUCTURE ONLY (see muscl_oracle.h) -- constrained-transport MHD part of the CPU
 * restatement: MHD_Muscl_3D and MHD_Muscl_2D, implementationVersion 0, staged exactly like the reference
 * (muscl/SolverMHDMuscl.cpp:463-513): ghost fill, copy, primitives, [dt], electric field,
 * magnetic slopes, trace (6 face + 12 edge states per cell), 3 HLLD fluxes per cell, 3 EMFs per
 * cell, flux update, CT update.  Same floating-point order as the reference functors, cell by cell.
 * Citations are relative to /root/reference/src.
 *
 * Parity pin: checked against dumps of the reference binary (tests/golden/, tests/test_oracle_*.py).
 */
#include "muscl_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { ID = 0, IP = 1, IU = 2, IV = 3, IW = 4, IA = 5, IB = 6, IC = 7, NV = 8 };
enum { RIEMANN_APPROX = 0, RIEMANN_LLF = 1, RIEMANN_HLL = 2, RIEMANN_HLLC = 3, RIEMANN_HLLD = 4 };
enum { IRT = 0, IRB = 1, ILT = 2, ILB = 3 };        /* shared/enums.h:195-198 */
enum { ILL = 0, IRL = 1, ILR = 2, IRR = 3 };        /* shared/enums.h:203-206 */
enum { EMFX = 0, EMFY = 1, EMFZ = 2 };              /* shared/enums.h:212-214 */

typedef struct {
  int isize, jsize, ksize;
  size_t sj, sk, sv;
} geom_t;

static geom_t geom(const orc_params* p)
{
  geom_t g;
  g.isize = p->nx + 2 * p->ghost;
  g.jsize = p->ny + 2 * p->ghost;
  g.ksize = (p->dim == 3) ? p->nz + 2 * p->ghost : 1;
  g.sj = (size_t)g.isize;
  g.sk = (size_t)g.isize * g.jsize;
  g.sv = g.sk * g.ksize;
  return g;
}
#define AT(A, i, j, k, v) (A)[(size_t)(i) + g.sj * (size_t)(j) + g.sk * (size_t)(k) + g.sv * (size_t)(v)]

static void get_state(const geom_t g, const double* A, int i, int j, int k, double* q)
{
  for (int v = 0; v < NV; ++v)
    q[v] = AT(A, i, j, k, v);
}
static void set_state(const geom_t g, double* A, int i, int j, int k, const double* q)
{
  for (int v = 0; v < NV; ++v)
    AT(A, i, j, k, v) = q[v];
}
static void swapv(double* a, double* b)
{
  double t = *a;
  *a = *b;
  *b = t;
}

/* ------------------------------------------------------------------------------------------- */
/* muscl/MHDBaseFunctor3D.h:194-242 constoprim_mhd (cIso == 0 branch: cIso is never read from the ini) */
static void constoprim_mhd(const orc_params* p, const double* u, const double* magNb, double* c, double* q)
{
  q[ID] = fmax(u[ID], p->smallr);
  q[IU] = u[IU] / q[ID];
  q[IV] = u[IV] / q[ID];
  q[IW] = u[IW] / q[ID];
  q[IA] = 0.5 * (u[IA] + magNb[0]);
  q[IB] = 0.5 * (u[IB] + magNb[1]);
  q[IC] = 0.5 * (u[IC] + magNb[2]);
  double eken = 0.5 * (q[IU] * q[IU] + q[IV] * q[IV] + q[IW] * q[IW]);
  double emag = 0.5 * (q[IA] * q[IA] + q[IB] * q[IB] + q[IC] * q[IC]);
  double eint = (u[IP] - emag) / q[ID] - eken;
  q[IP] = fmax((p->gamma0 - 1.0) * q[ID] * eint, q[ID] * p->smallp);
  *c = sqrt(p->gamma0 * q[IP] / q[ID]);
}

/* muscl/MHDRunFunctors3D.h:88-163 ConvertToPrimitivesFunctor3D_MHD: the last layer is never written */
void orc_mhd_prims(const orc_params* p, const double* U, double* Q)
{
  const geom_t g = geom(p);
  if (p->dim == 2) { /* muscl/MHDRunFunctors2D.h:85-160: magFieldNeighbors[IZ] = 0 */
    for (int j = 0; j < g.jsize - 1; ++j)
      for (int i = 0; i < g.isize - 1; ++i) {
        double u[NV], q[NV], nb[3], c;
        get_state(g, U, i, j, 0, u);
        nb[0] = AT(U, i + 1, j, 0, IA);
        nb[1] = AT(U, i, j + 1, 0, IB);
        nb[2] = 0.0;
        constoprim_mhd(p, u, nb, &c, q);
        set_state(g, Q, i, j, 0, q);
      }
    return;
  }
  for (int k = 0; k < g.ksize - 1; ++k)
    for (int j = 0; j < g.jsize - 1; ++j)
      for (int i = 0; i < g.isize - 1; ++i) {
        double u[NV], q[NV], nb[3], c;
        get_state(g, U, i, j, k, u);
        nb[0] = AT(U, i + 1, j, k, IA);
        nb[1] = AT(U, i, j + 1, k, IB);
        nb[2] = AT(U, i, j, k + 1, IC);
        constoprim_mhd(p, u, nb, &c, q);
        set_state(g, Q, i, j, k, q);
      }
}

/* shared/mhd_utils.h:321-366 find_speed_info<THREE_D>; muscl/MHDRunFunctors3D.h:16-83 */
double orc_mhd_invdt(const orc_params* p, const double* Q)
{
  const geom_t g = geom(p);
  const int gw = p->ghost;
  double invdt = 0.0;
  if (p->dim == 2) { /* muscl/MHDRunFunctors2D.h:16-80, find_speed_info<TWO_D> */
    for (int j = gw; j < g.jsize - gw; ++j)
      for (int i = gw; i < g.isize - gw; ++i) {
        double q[NV];
        get_state(g, Q, i, j, 0, q);
        const double d = q[ID], pp = q[IP], a = q[IA], b = q[IB], c = q[IC];
        const double b2 = a * a + b * b + c * c;
        const double c2 = p->gamma0 * pp / d;
        const double d2 = 0.5 * (b2 / d + c2);
        double cf = sqrt(d2 + sqrt(d2 * d2 - c2 * a * a / d));
        const double vx = cf + fabs(q[IU]);
        cf = sqrt(d2 + sqrt(d2 * d2 - c2 * b * b / d));
        const double vy = cf + fabs(q[IV]);
        invdt = fmax(invdt, vx / p->dx + vy / p->dy);
      }
    return invdt;
  }
  for (int k = gw; k < g.ksize - gw; ++k)
    for (int j = gw; j < g.jsize - gw; ++j)
      for (int i = gw; i < g.isize - gw; ++i) {
        double q[NV];
        get_state(g, Q, i, j, k, q);
        const double d = q[ID], pp = q[IP], a = q[IA], b = q[IB], c = q[IC];
        const double b2 = a * a + b * b + c * c;
        const double c2 = p->gamma0 * pp / d;
        const double d2 = 0.5 * (b2 / d + c2);
        double cf = sqrt(d2 + sqrt(d2 * d2 - c2 * a * a / d));
        const double vx = cf + fabs(q[IU]);
        cf = sqrt(d2 + sqrt(d2 * d2 - c2 * b * b / d));
        const double vy = cf + fabs(q[IV]);
        cf = sqrt(d2 + sqrt(d2 * d2 - c2 * c * c / d));
        const double vz = cf + fabs(q[IW]);
        invdt = fmax(invdt, vx / p->dx + vy / p->dy + vz / p->dz);
      }
  return invdt;
}

/* ------------------------------------------------------------------------------------------- */
/* muscl/MHDBaseFunctor3D.h:262-337 slope_unsplit_hydro_3d_scalar */
static double slope1(double st, double q, double qp, double qm)
{
  double dlft = st * (q - qm);
  double drgt = st * (qp - q);
  double dcen = 0.5 * (qp - qm);
  double dsgn = (dcen >= 0.0) ? 1.0 : -1.0;
  double slop = fmin(fabs(dlft), fabs(drgt));
  double dlim = slop;
  if ((dlft * drgt) <= 0.0)
    dlim = 0.0;
  return dsgn * fmin(dlim, fabs(dcen));
}

/* muscl/MHDBaseFunctor3D.h:364-495 slope_unsplit_hydro_3d, including the qMinusY[IC] slip at :458 */
static void slope_unsplit_hydro_3d(const orc_params* p, const double* q, const double* qpx, const double* qmx, const double* qpy,
                                   const double* qmy, const double* qpz, const double* qmz, double dq[3][NV])
{
  const double st = p->slope_type;
  if (st == 1 || st == 2) {
    for (int v = 0; v < NV; ++v) {
      dq[0][v] = slope1(st, q[v], qpx[v], qmx[v]);
      dq[1][v] = slope1(st, q[v], qpy[v], qmy[v]);
      dq[2][v] = slope1(st, q[v], qpz[v], (v == IC) ? qmy[v] : qmz[v]);
    }
  } else {
    for (int d = 0; d < 3; ++d)
      for (int v = 0; v < NV; ++v)
        dq[d][v] = 0.0;
  }
}

/* ------------------------------------------------------------------------------------------- */
/* shared/mhd_utils.h:91-117 find_speed_fast<IX> */
static double find_speed_fast_x(const orc_params* p, const double* q)
{
  const double d = q[ID], pp = q[IP], a = q[IA], b = q[IB], c = q[IC];
  const double b2 = a * a + b * b + c * c;
  const double c2 = p->gamma0 * pp / d;
  const double d2 = 0.5 * (b2 / d + c2);
  return sqrt(d2 + sqrt(d2 * d2 - c2 * a * a / d));
}
static double find_speed_fast_y(const orc_params* p, const double* q)
{
  const double d = q[ID], pp = q[IP], a = q[IA], b = q[IB], c = q[IC];
  const double b2 = a * a + b * b + c * c;
  const double c2 = p->gamma0 * pp / d;
  const double d2 = 0.5 * (b2 / d + c2);
  return sqrt(d2 + sqrt(d2 * d2 - c2 * b * b / d));
}

/* shared/mhd_utils.h:175-231 find_mhd_flux */
static void find_mhd_flux(const orc_params* p, const double* q, double* cv, double* ff)
{
  const double entho = 1.0 / (p->gamma0 - 1.0);
  const double pr = q[IP];
  const double d = q[ID], u = q[IU], v = q[IV], w = q[IW], a = q[IA], b = q[IB], c = q[IC];
  const double ecin = 0.5 * (u * u + v * v + w * w) * d;
  const double emag = 0.5 * (a * a + b * b + c * c);
  const double etot = pr * entho + ecin + emag;
  const double ptot = pr + emag;
  cv[ID] = d; cv[IP] = etot; cv[IU] = d * u; cv[IV] = d * v; cv[IW] = d * w; cv[IA] = a; cv[IB] = b; cv[IC] = c;
  ff[ID] = d * u;
  ff[IP] = (etot + ptot) * u - a * (a * u + b * v + c * w);
  ff[IU] = d * u * u - a * a + ptot;
  ff[IV] = d * u * v - a * b;
  ff[IW] = d * u * w - a * c;
  ff[IA] = 0.0;
  ff[IB] = b * u - a * v;
  ff[IC] = c * u - a * w;
}

/* shared/RiemannSolvers_MHD.h:29-64 */
static void riemann_hll_mhd(const orc_params* p, double* ql, double* qr, double* flux)
{
  double bx = 0.5 * (ql[IA] + qr[IA]);
  ql[IA] = bx;
  qr[IA] = bx;
  double ul[NV], fl[NV], ur[NV], fr[NV];
  find_mhd_flux(p, ql, ul, fl);
  find_mhd_flux(p, qr, ur, fr);
  double cfl_ = find_speed_fast_x(p, ql), cfr = find_speed_fast_x(p, qr);
  double vl = ql[IU], vr = qr[IU];
  double sl = fmin(fmin(vl, vr) - fmax(cfl_, cfr), 0.0);
  double sr = fmax(fmax(vl, vr) + fmax(cfl_, cfr), 0.0);
  for (int v = 0; v < NV; ++v)
    flux[v] = (sr * fl[v] - sl * fr[v] + sr * sl * (ur[v] - ul[v])) / (sr - sl);
}

/* shared/RiemannSolvers_MHD.h:80-108 (find_speed_info 1-D variant, mhd_utils.h:377-403) */
static void riemann_llf_mhd(const orc_params* p, double* ql, double* qr, double* flux)
{
  double bx = 0.5 * (ql[IA] + qr[IA]);
  ql[IA] = bx;
  qr[IA] = bx;
  double ul[NV], fl[NV], ur[NV], fr[NV];
  find_mhd_flux(p, ql, ul, fl);
  find_mhd_flux(p, qr, ur, fr);
  for (int v = 0; v < NV; ++v)
    flux[v] = (fl[v] + fr[v]) / 2;
  double cl = find_speed_fast_x(p, ql) + fabs(ql[IU]);
  double cr = find_speed_fast_x(p, qr) + fabs(qr[IU]);
  double vel = fmax(cl, cr);
  for (int v = 0; v < NV; ++v)
    flux[v] -= vel * (ur[v] - ul[v]) / 2;
}

/* shared/RiemannSolvers_MHD.h:132-364 riemann_hlld (Miyoshi & Kusano 2005) */
static void riemann_hlld(const orc_params* p, double* qleft, double* qright, double* flux)
{
  const double gamma0 = p->gamma0;
  const double entho = 1.0 / (gamma0 - 1.0);
  double a = 0.5 * (qleft[IA] + qright[IA]);
  double sgnm = (a >= 0) ? 1.0 : -1.0;
  qleft[IA] = a;
  qright[IA] = a;

  double rl = qleft[ID], pl = qleft[IP], ul = qleft[IU], vl = qleft[IV], wl = qleft[IW], bl = qleft[IB], cl = qleft[IC];
  double ecinl = 0.5 * (ul * ul + vl * vl + wl * wl) * rl;
  double emagl = 0.5 * (a * a + bl * bl + cl * cl);
  double etotl = pl * entho + ecinl + emagl;
  double ptotl = pl + emagl;
  double vdotbl = ul * a + vl * bl + wl * cl;

  double rr = qright[ID], pr = qright[IP], ur = qright[IU], vr = qright[IV], wr = qright[IW], br = qright[IB], cr = qright[IC];
  double ecinr = 0.5 * (ur * ur + vr * vr + wr * wr) * rr;
  double emagr = 0.5 * (a * a + br * br + cr * cr);
  double etotr = pr * entho + ecinr + emagr;
  double ptotr = pr + emagr;
  double vdotbr = ur * a + vr * br + wr * cr;

  double cfastl = find_speed_fast_x(p, qleft);
  double cfastr = find_speed_fast_x(p, qright);

  double sl = fmin(ul, ur) - fmax(cfastl, cfastr);
  double sr = fmax(ul, ur) + fmax(cfastl, cfastr);

  double rcl = rl * (ul - sl);
  double rcr = rr * (sr - ur);

  double ustar = (rcr * ur + rcl * ul + (ptotl - ptotr)) / (rcr + rcl);
  double ptotstar = (rcr * ptotl + rcl * ptotr + rcl * rcr * (ul - ur)) / (rcr + rcl);

  double estar, rstarl, el;
  rstarl = rl * (sl - ul) / (sl - ustar);
  estar = rl * (sl - ul) * (sl - ustar) - a * a;
  el = rl * (sl - ul) * (sl - ul) - a * a;
  double vstarl, wstarl, bstarl, cstarl;
  if (a * a > 0 && fabs(estar / (a * a) - 1.0) <= 1e-8) {
    vstarl = vl; bstarl = bl; wstarl = wl; cstarl = cl;
  } else {
    vstarl = vl - a * bl * (ustar - ul) / estar;
    bstarl = bl * el / estar;
    wstarl = wl - a * cl * (ustar - ul) / estar;
    cstarl = cl * el / estar;
  }
  double vdotbstarl = ustar * a + vstarl * bstarl + wstarl * cstarl;
  double etotstarl = ((sl - ul) * etotl - ptotl * ul + ptotstar * ustar + a * (vdotbl - vdotbstarl)) / (sl - ustar);
  double sqrrstarl = sqrt(rstarl);
  double calfvenl = fabs(a) / sqrrstarl;
  double sal = ustar - calfvenl;

  double rstarr, er;
  rstarr = rr * (sr - ur) / (sr - ustar);
  estar = rr * (sr - ur) * (sr - ustar) - a * a;
  er = rr * (sr - ur) * (sr - ur) - a * a;
  double vstarr, wstarr, bstarr, cstarr;
  if (a * a > 0 && fabs(estar / (a * a) - 1.0) <= 1e-8) {
    vstarr = vr; bstarr = br; wstarr = wr; cstarr = cr;
  } else {
    vstarr = vr - a * br * (ustar - ur) / estar;
    bstarr = br * er / estar;
    wstarr = wr - a * cr * (ustar - ur) / estar;
    cstarr = cr * er / estar;
  }
  double vdotbstarr = ustar * a + vstarr * bstarr + wstarr * cstarr;
  double etotstarr = ((sr - ur) * etotr - ptotr * ur + ptotstar * ustar + a * (vdotbr - vdotbstarr)) / (sr - ustar);
  double sqrrstarr = sqrt(rstarr);
  double calfvenr = fabs(a) / sqrrstarr;
  double sar = ustar + calfvenr;

  double vstarstar = (sqrrstarl * vstarl + sqrrstarr * vstarr + sgnm * (bstarr - bstarl)) / (sqrrstarl + sqrrstarr);
  double wstarstar = (sqrrstarl * wstarl + sqrrstarr * wstarr + sgnm * (cstarr - cstarl)) / (sqrrstarl + sqrrstarr);
  double bstarstar = (sqrrstarl * bstarr + sqrrstarr * bstarl + sgnm * sqrrstarl * sqrrstarr * (vstarr - vstarl)) / (sqrrstarl + sqrrstarr);
  double cstarstar = (sqrrstarl * cstarr + sqrrstarr * cstarl + sgnm * sqrrstarl * sqrrstarr * (wstarr - wstarl)) / (sqrrstarl + sqrrstarr);
  double vdotbstarstar = ustar * a + vstarstar * bstarstar + wstarstar * cstarstar;
  double etotstarstarl = etotstarl - sgnm * sqrrstarl * (vdotbstarl - vdotbstarstar);
  double etotstarstarr = etotstarr + sgnm * sqrrstarr * (vdotbstarr - vdotbstarstar);

  double ro, uo, vo, wo, bo, co, ptoto, etoto, vdotbo;
  if (sl > 0) {
    ro = rl; uo = ul; vo = vl; wo = wl; bo = bl; co = cl; ptoto = ptotl; etoto = etotl; vdotbo = vdotbl;
  } else if (sal > 0) {
    ro = rstarl; uo = ustar; vo = vstarl; wo = wstarl; bo = bstarl; co = cstarl; ptoto = ptotstar; etoto = etotstarl; vdotbo = vdotbstarl;
  } else if (ustar > 0) {
    ro = rstarl; uo = ustar; vo = vstarstar; wo = wstarstar; bo = bstarstar; co = cstarstar; ptoto = ptotstar; etoto = etotstarstarl; vdotbo = vdotbstarstar;
  } else if (sar > 0) {
    ro = rstarr; uo = ustar; vo = vstarstar; wo = wstarstar; bo = bstarstar; co = cstarstar; ptoto = ptotstar; etoto = etotstarstarr; vdotbo = vdotbstarstar;
  } else if (sr > 0) {
    ro = rstarr; uo = ustar; vo = vstarr; wo = wstarr; bo = bstarr; co = cstarr; ptoto = ptotstar; etoto = etotstarr; vdotbo = vdotbstarr;
  } else {
    ro = rr; uo = ur; vo = vr; wo = wr; bo = br; co = cr; ptoto = ptotr; etoto = etotr; vdotbo = vdotbr;
  }

  flux[ID] = ro * uo;
  flux[IP] = (etoto + ptoto) * uo - a * vdotbo;
  flux[IU] = ro * uo * uo - a * a + ptoto;
  flux[IV] = ro * uo * vo - a * bo;
  flux[IW] = ro * uo * wo - a * co;
  flux[IA] = 0.0;
  flux[IB] = bo * uo - a * vo;
  flux[IC] = co * uo - a * wo;
}

/* shared/RiemannSolvers_MHD.h:371-389: no solver matches -> flux left untouched (we zero it) */
static void riemann_mhd(const orc_params* p, double* ql, double* qr, double* flux)
{
  if (p->riemann == RIEMANN_HLLD)
    riemann_hlld(p, ql, qr, flux);
  else if (p->riemann == RIEMANN_HLL)
    riemann_hll_mhd(p, ql, qr, flux);
  else if (p->riemann == RIEMANN_LLF)
    riemann_llf_mhd(p, ql, qr, flux);
}

void orc_riemann_mhd(const orc_params* p, const double* ql, const double* qr, double* flux)
{
  double l[NV], r[NV];
  memcpy(l, ql, sizeof l);
  memcpy(r, qr, sizeof r);
  memset(flux, 0, NV * sizeof(double));
  riemann_mhd(p, l, r, flux);
}

/* ------------------------------------------------------------------------------------------- */
static double fmax4(double a0, double a1, double a2, double a3)
{
  double r = a0;
  r = (a1 > r) ? a1 : r;
  r = (a2 > r) ? a2 : r;
  r = (a3 > r) ? a3 : r;
  return r;
}
static double fmin4(double a0, double a1, double a2, double a3)
{
  double r = a0;
  r = (a1 < r) ? a1 : r;
  r = (a2 < r) ? a2 : r;
  r = (a3 < r) ? a3 : r;
  return r;
}
static double fmax5(double a0, double a1, double a2, double a3, double a4)
{
  double r = a0;
  r = (a1 > r) ? a1 : r;
  r = (a2 > r) ? a2 : r;
  r = (a3 > r) ? a3 : r;
  r = (a4 > r) ? a4 : r;
  return r;
}

/* shared/RiemannSolvers_MHD.h:397-627 mag_riemann2d_hlld */
static double mag_riemann2d_hlld(const orc_params* p, const double qLLRR[4][NV], const double eLLRR[4])
{
  const double* qLL = qLLRR[ILL];
  const double* qRL = qLLRR[IRL];
  const double* qLR = qLLRR[ILR];
  const double* qRR = qLLRR[IRR];
  const double ELL = eLLRR[ILL], ERL = eLLRR[IRL], ELR = eLLRR[ILR], ERR = eLLRR[IRR];

  const double rLL = qLL[ID], pLL = qLL[IP], uLL = qLL[IU], vLL = qLL[IV], aLL = qLL[IA], bLL = qLL[IB], cLL = qLL[IC];
  const double rLR = qLR[ID], pLR = qLR[IP], uLR = qLR[IU], vLR = qLR[IV], aLR = qLR[IA], bLR = qLR[IB], cLR = qLR[IC];
  const double rRL = qRL[ID], pRL = qRL[IP], uRL = qRL[IU], vRL = qRL[IV], aRL = qRL[IA], bRL = qRL[IB], cRL = qRL[IC];
  const double rRR = qRR[ID], pRR = qRR[IP], uRR = qRR[IU], vRR = qRR[IV], aRR = qRR[IA], bRR = qRR[IB], cRR = qRR[IC];

  double cFastLLx = find_speed_fast_x(p, qLL), cFastLRx = find_speed_fast_x(p, qLR);
  double cFastRLx = find_speed_fast_x(p, qRL), cFastRRx = find_speed_fast_x(p, qRR);
  double cFastLLy = find_speed_fast_y(p, qLL), cFastLRy = find_speed_fast_y(p, qLR);
  double cFastRLy = find_speed_fast_y(p, qRL), cFastRRy = find_speed_fast_y(p, qRR);

  double SL = fmin4(uLL, uLR, uRL, uRR) - fmax4(cFastLLx, cFastLRx, cFastRLx, cFastRRx);
  double SR = fmax4(uLL, uLR, uRL, uRR) + fmax4(cFastLLx, cFastLRx, cFastRLx, cFastRRx);
  double SB = fmin4(vLL, vLR, vRL, vRR) - fmax4(cFastLLy, cFastLRy, cFastRLy, cFastRRy);
  double ST = fmax4(vLL, vLR, vRL, vRR) + fmax4(cFastLLy, cFastLRy, cFastRLy, cFastRRy);

  double PtotLL = pLL + 0.5 * (aLL * aLL + bLL * bLL + cLL * cLL);
  double PtotLR = pLR + 0.5 * (aLR * aLR + bLR * bLR + cLR * cLR);
  double PtotRL = pRL + 0.5 * (aRL * aRL + bRL * bRL + cRL * cRL);
  double PtotRR = pRR + 0.5 * (aRR * aRR + bRR * bRR + cRR * cRR);

  double rcLLx = rLL * (uLL - SL);
  double rcRLx = rRL * (SR - uRL);
  double rcLRx = rLR * (uLR - SL);
  double rcRRx = rRR * (SR - uRR);
  double rcLLy = rLL * (vLL - SB);
  double rcLRy = rLR * (ST - vLR);
  double rcRLy = rRL * (vRL - SB);
  double rcRRy = rRR * (ST - vRR);

  double ustar = (rcLLx * uLL + rcLRx * uLR + rcRLx * uRL + rcRRx * uRR + (PtotLL - PtotRL + PtotLR - PtotRR)) /
                 (rcLLx + rcLRx + rcRLx + rcRRx);
  double vstar = (rcLLy * vLL + rcLRy * vLR + rcRLy * vRL + rcRRy * vRR + (PtotLL - PtotLR + PtotRL - PtotRR)) /
                 (rcLLy + rcLRy + rcRLy + rcRRy);

  double rstarLLx = rLL * (SL - uLL) / (SL - ustar);
  double BstarLL = bLL * (SL - uLL) / (SL - ustar);
  double rstarLLy = rLL * (SB - vLL) / (SB - vstar);
  double AstarLL = aLL * (SB - vLL) / (SB - vstar);
  double rstarLL = rLL * (SL - uLL) / (SL - ustar) * (SB - vLL) / (SB - vstar);
  double EstarLLx = ustar * BstarLL - vLL * aLL;
  double EstarLLy = uLL * bLL - vstar * AstarLL;
  double EstarLL = ustar * BstarLL - vstar * AstarLL;

  double rstarLRx = rLR * (SL - uLR) / (SL - ustar);
  double BstarLR = bLR * (SL - uLR) / (SL - ustar);
  double rstarLRy = rLR * (ST - vLR) / (ST - vstar);
  double AstarLR = aLR * (ST - vLR) / (ST - vstar);
  double rstarLR = rLR * (SL - uLR) / (SL - ustar) * (ST - vLR) / (ST - vstar);
  double EstarLRx = ustar * BstarLR - vLR * aLR;
  double EstarLRy = uLR * bLR - vstar * AstarLR;
  double EstarLR = ustar * BstarLR - vstar * AstarLR;

  double rstarRLx = rRL * (SR - uRL) / (SR - ustar);
  double BstarRL = bRL * (SR - uRL) / (SR - ustar);
  double rstarRLy = rRL * (SB - vRL) / (SB - vstar);
  double AstarRL = aRL * (SB - vRL) / (SB - vstar);
  double rstarRL = rRL * (SR - uRL) / (SR - ustar) * (SB - vRL) / (SB - vstar);
  double EstarRLx = ustar * BstarRL - vRL * aRL;
  double EstarRLy = uRL * bRL - vstar * AstarRL;
  double EstarRL = ustar * BstarRL - vstar * AstarRL;

  double rstarRRx = rRR * (SR - uRR) / (SR - ustar);
  double BstarRR = bRR * (SR - uRR) / (SR - ustar);
  double rstarRRy = rRR * (ST - vRR) / (ST - vstar);
  double AstarRR = aRR * (ST - vRR) / (ST - vstar);
  double rstarRR = rRR * (SR - uRR) / (SR - ustar) * (ST - vRR) / (ST - vstar);
  double EstarRRx = ustar * BstarRR - vRR * aRR;
  double EstarRRy = uRR * bRR - vstar * AstarRR;
  double EstarRR = ustar * BstarRR - vstar * AstarRR;

  double calfvenL = fmax5(fabs(aLR) / sqrt(rstarLRx), fabs(AstarLR) / sqrt(rstarLR), fabs(aLL) / sqrt(rstarLLx),
                          fabs(AstarLL) / sqrt(rstarLL), p->smallc);
  double calfvenR = fmax5(fabs(aRR) / sqrt(rstarRRx), fabs(AstarRR) / sqrt(rstarRR), fabs(aRL) / sqrt(rstarRLx),
                          fabs(AstarRL) / sqrt(rstarRL), p->smallc);
  double calfvenB = fmax5(fabs(bLL) / sqrt(rstarLLy), fabs(BstarLL) / sqrt(rstarLL), fabs(bRL) / sqrt(rstarRLy),
                          fabs(BstarRL) / sqrt(rstarRL), p->smallc);
  double calfvenT = fmax5(fabs(bLR) / sqrt(rstarLRy), fabs(BstarLR) / sqrt(rstarLR), fabs(bRR) / sqrt(rstarRRy),
                          fabs(BstarRR) / sqrt(rstarRR), p->smallc);

  double SAL = fmin(ustar - calfvenL, 0.0);
  double SAR = fmax(ustar + calfvenR, 0.0);
  double SAB = fmin(vstar - calfvenB, 0.0);
  double SAT = fmax(vstar + calfvenT, 0.0);

  double AstarT = (SAR * AstarRR - SAL * AstarLR) / (SAR - SAL);
  double AstarB = (SAR * AstarRL - SAL * AstarLL) / (SAR - SAL);
  double BstarR = (SAT * BstarRR - SAB * BstarRL) / (SAT - SAB);
  double BstarL = (SAT * BstarLR - SAB * BstarLL) / (SAT - SAB);

  double E = 0, tmpE = 0;
  int SB_pos = (int)(1 + copysign(1.0, SB)) / 2, SB_neg = 1 - SB_pos;
  int ST_pos = (int)(1 + copysign(1.0, ST)) / 2, ST_neg = 1 - ST_pos;
  int SL_pos = (int)(1 + copysign(1.0, SL)) / 2, SL_neg = 1 - SL_pos;
  int SR_pos = (int)(1 + copysign(1.0, SR)) / 2, SR_neg = 1 - SR_pos;

  tmpE = (SAL * SAB * EstarRR - SAL * SAT * EstarRL - SAR * SAB * EstarLR + SAR * SAT * EstarLL) / (SAR - SAL) / (SAT - SAB) -
         SAT * SAB / (SAT - SAB) * (AstarT - AstarB) + SAR * SAL / (SAR - SAL) * (BstarR - BstarL);
  E += (SB_neg * ST_pos * SL_neg * SR_pos) * tmpE;

  tmpE = (SAR * EstarLLx - SAL * EstarRLx + SAR * SAL * (bRL - bLL)) / (SAR - SAL);
  tmpE = SL_pos * ELL + SL_neg * SR_neg * ERL + SL_neg * SR_pos * tmpE;
  E += SB_pos * tmpE;

  tmpE = (SAR * EstarLRx - SAL * EstarRRx + SAR * SAL * (bRR - bLR)) / (SAR - SAL);
  tmpE = SL_pos * ELR + SL_neg * SR_neg * ERR + SL_neg * SR_pos * tmpE;
  E += (SB_neg * ST_neg) * tmpE;

  tmpE = (SAT * EstarLLy - SAB * EstarLRy - SAT * SAB * (aLR - aLL)) / (SAT - SAB);
  E += (SB_neg * ST_pos * SL_pos) * tmpE;

  tmpE = (SAT * EstarRLy - SAB * EstarRRy - SAT * SAB * (aRR - aRL)) / (SAT - SAB);
  E += (SB_neg * ST_pos * SL_neg * SR_neg) * tmpE;

  (void)rstarLLx; (void)rstarLLy; (void)rstarLRx; (void)rstarLRy; (void)rstarRLx; (void)rstarRLy; (void)rstarRRx; (void)rstarRRy;
  return E;
}

/* shared/RiemannSolvers_MHD.h:650-871 compute_emf<emfDir>; qEdge indexed IRT, IRB, ILT, ILB */
static double compute_emf(const orc_params* p, int emfDir, double qEdge[4][NV])
{
  const double* qRT = qEdge[IRT];
  const double* qLT = qEdge[ILT];
  const double* qRB = qEdge[IRB];
  const double* qLB = qEdge[ILB];
  double qLLRR[4][NV];
  double* qLL = qLLRR[ILL];
  double* qRL = qLLRR[IRL];
  double* qLR = qLLRR[ILR];
  double* qRR = qLLRR[IRR];
  int iu, iv, iw, ia, ib, ic;
  if (emfDir == EMFZ) {
    iu = IU; iv = IV; iw = IW; ia = IA; ib = IB; ic = IC;
  } else if (emfDir == EMFY) {
    iu = IW; iv = IU; iw = IV; ia = IC; ib = IA; ic = IB;
  } else {
    iu = IV; iv = IW; iw = IU; ia = IB; ib = IC; ic = IA;
  }
  qLL[ID] = qRT[ID]; qRL[ID] = qLT[ID]; qLR[ID] = qRB[ID]; qRR[ID] = qLB[ID];
  qLL[IP] = qRT[IP]; qRL[IP] = qLT[IP]; qLR[IP] = qRB[IP]; qRR[IP] = qLB[IP];
  qLL[IU] = qRT[iu]; qRL[IU] = qLT[iu]; qLR[IU] = qRB[iu]; qRR[IU] = qLB[iu];
  qLL[IV] = qRT[iv]; qRL[IV] = qLT[iv]; qLR[IV] = qRB[iv]; qRR[IV] = qLB[iv];
  qLL[IA] = 0.5 * (qRT[ia] + qLT[ia]);
  qRL[IA] = 0.5 * (qRT[ia] + qLT[ia]);
  qLR[IA] = 0.5 * (qRB[ia] + qLB[ia]);
  qRR[IA] = 0.5 * (qRB[ia] + qLB[ia]);
  qLL[IB] = 0.5 * (qRT[ib] + qRB[ib]);
  qRL[IB] = 0.5 * (qLT[ib] + qLB[ib]);
  qLR[IB] = 0.5 * (qRT[ib] + qRB[ib]);
  qRR[IB] = 0.5 * (qLT[ib] + qLB[ib]);
  qLL[IW] = qRT[iw]; qRL[IW] = qLT[iw]; qLR[IW] = qRB[iw]; qRR[IW] = qLB[iw];
  qLL[IC] = qRT[ic]; qRL[IC] = qLT[ic]; qLR[IC] = qRB[ic]; qRR[IC] = qLB[ic];

  double eLLRR[4];
  eLLRR[ILL] = qLL[IU] * qLL[IB] - qLL[IV] * qLL[IA];
  eLLRR[IRL] = qRL[IU] * qRL[IB] - qRL[IV] * qRL[IA];
  eLLRR[ILR] = qLR[IU] * qLR[IB] - qLR[IV] * qLR[IA];
  eLLRR[IRR] = qRR[IU] * qRR[IB] - qRR[IV] * qRR[IA];
  return mag_riemann2d_hlld(p, (const double(*)[NV])qLLRR, eLLRR);
}

/* KAT entry: qEdge = 4 states in the order IRT, IRB, ILT, ILB */
double orc_compute_emf(const orc_params* p, int emfDir, const double* qEdge4x8)
{
  double q[4][NV];
  memcpy(q, qEdge4x8, sizeof q);
  return compute_emf(p, emfDir, q);
}

/* ------------------------------------------------------------------------------------------- */
/* muscl/MHDBaseFunctor3D.h:690-1114 trace_unsplit_mhd_3d_simpler (Omega0 == 0: never read from the ini) */
static void trace_unsplit_mhd_3d_simpler(const orc_params* pr_, const double* q, double dq[3][NV], const double bfNb[6],
                                         const double dbf[12], const double elecFields[3][2][2], double dtdx, double dtdy,
                                         double dtdz, double qm[3][NV], double qp[3][NV], double qEdge[4][3][NV])
{
  const double(*Ex)[2] = elecFields[0];
  const double(*Ey)[2] = elecFields[1];
  const double(*Ez)[2] = elecFields[2];
  double* qRT_X = qEdge[0][0]; double* qRB_X = qEdge[1][0]; double* qLT_X = qEdge[2][0]; double* qLB_X = qEdge[3][0];
  double* qRT_Y = qEdge[0][1]; double* qRB_Y = qEdge[1][1]; double* qLT_Y = qEdge[2][1]; double* qLB_Y = qEdge[3][1];
  double* qRT_Z = qEdge[0][2]; double* qRB_Z = qEdge[1][2]; double* qLT_Z = qEdge[2][2]; double* qLB_Z = qEdge[3][2];

  const double gamma = pr_->gamma0, smallR = pr_->smallr, smallp = pr_->smallp;

  const double ELL = Ex[0][0], ELR = Ex[0][1], ERL = Ex[1][0], ERR = Ex[1][1];
  const double FLL = Ey[0][0], FLR = Ey[0][1], FRL = Ey[1][0], FRR = Ey[1][1];
  const double GLL = Ez[0][0], GLR = Ez[0][1], GRL = Ez[1][0], GRR = Ez[1][1];

  double r = q[ID], p = q[IP], u = q[IU], v = q[IV], w = q[IW], A = q[IA], B = q[IB], C = q[IC];
  double AL = bfNb[0], AR = bfNb[1], BL = bfNb[2], BR = bfNb[3], CL = bfNb[4], CR = bfNb[5];

  double drx = dq[0][ID] * 0.5, dpx = dq[0][IP] * 0.5, dux = dq[0][IU] * 0.5, dvx = dq[0][IV] * 0.5, dwx = dq[0][IW] * 0.5;
  double dCx = dq[0][IC] * 0.5, dBx = dq[0][IB] * 0.5;
  double dry = dq[1][ID] * 0.5, dpy = dq[1][IP] * 0.5, duy = dq[1][IU] * 0.5, dvy = dq[1][IV] * 0.5, dwy = dq[1][IW] * 0.5;
  double dCy = dq[1][IC] * 0.5, dAy = dq[1][IA] * 0.5;
  double drz = dq[2][ID] * 0.5, dpz = dq[2][IP] * 0.5, duz = dq[2][IU] * 0.5, dvz = dq[2][IV] * 0.5, dwz = dq[2][IW] * 0.5;
  double dAz = dq[2][IA] * 0.5, dBz = dq[2][IB] * 0.5;

  double dALy = 0.5 * dbf[0], dALz = 0.5 * dbf[1], dBLx = 0.5 * dbf[2], dBLz = 0.5 * dbf[3], dCLx = 0.5 * dbf[4], dCLy = 0.5 * dbf[5];
  double dARy = 0.5 * dbf[6], dARz = 0.5 * dbf[7], dBRx = 0.5 * dbf[8], dBRz = 0.5 * dbf[9], dCRx = 0.5 * dbf[10], dCRy = 0.5 * dbf[11];

  double dAx = 0.5 * (AR - AL);
  double dBy = 0.5 * (BR - BL);
  double dCz = 0.5 * (CR - CL);

  double sr0, su0, sv0, sw0, sp0, sA0, sB0, sC0, sAL0, sAR0, sBL0, sBR0, sCL0, sCR0;
  sr0 = (-u * drx - dux * r) * dtdx + (-v * dry - dvy * r) * dtdy + (-w * drz - dwz * r) * dtdz;
  su0 = (-u * dux - (dpx + B * dBx + C * dCx) / r) * dtdx + (-v * duy + B * dAy / r) * dtdy + (-w * duz + C * dAz / r) * dtdz;
  sv0 = (-u * dvx + A * dBx / r) * dtdx + (-v * dvy - (dpy + A * dAy + C * dCy) / r) * dtdy + (-w * dvz + C * dBz / r) * dtdz;
  sw0 = (-u * dwx + A * dCx / r) * dtdx + (-v * dwy + B * dCy / r) * dtdy + (-w * dwz - (dpz + A * dAz + B * dBz) / r) * dtdz;
  sp0 = (-u * dpx - dux * gamma * p) * dtdx + (-v * dpy - dvy * gamma * p) * dtdy + (-w * dpz - dwz * gamma * p) * dtdz;
  sA0 = (u * dBy + B * duy - v * dAy - A * dvy) * dtdy + (u * dCz + C * duz - w * dAz - A * dwz) * dtdz;
  sB0 = (v * dAx + A * dvx - u * dBx - B * dux) * dtdx + (v * dCz + C * dvz - w * dBz - B * dwz) * dtdz;
  sC0 = (w * dAx + A * dwx - u * dCx - C * dux) * dtdx + (w * dBy + B * dwy - v * dCy - C * dvy) * dtdy;

  sAL0 = +(GLR - GLL) * dtdy * 0.5 - (FLR - FLL) * dtdz * 0.5;
  sAR0 = +(GRR - GRL) * dtdy * 0.5 - (FRR - FRL) * dtdz * 0.5;
  sBL0 = -(GRL - GLL) * dtdx * 0.5 + (ELR - ELL) * dtdz * 0.5;
  sBR0 = -(GRR - GLR) * dtdx * 0.5 + (ERR - ERL) * dtdz * 0.5;
  sCL0 = +(FRL - FLL) * dtdx * 0.5 - (ERL - ELL) * dtdy * 0.5;
  sCR0 = +(FRR - FLR) * dtdx * 0.5 - (ERR - ELR) * dtdy * 0.5;

  r = r + sr0; u = u + su0; v = v + sv0; w = w + sw0; p = p + sp0; A = A + sA0; B = B + sB0; C = C + sC0;
  AL = AL + sAL0; AR = AR + sAR0; BL = BL + sBL0; BR = BR + sBR0; CL = CL + sCL0; CR = CR + sCR0;

#define FLOOR(s)                       \
  (s)[ID] = fmax(smallR, (s)[ID]);     \
  (s)[IP] = fmax(smallp, (s)[IP])

  qp[0][ID] = r - drx; qp[0][IU] = u - dux; qp[0][IV] = v - dvx; qp[0][IW] = w - dwx; qp[0][IP] = p - dpx;
  qp[0][IA] = AL; qp[0][IB] = B - dBx; qp[0][IC] = C - dCx; FLOOR(qp[0]);
  qm[0][ID] = r + drx; qm[0][IU] = u + dux; qm[0][IV] = v + dvx; qm[0][IW] = w + dwx; qm[0][IP] = p + dpx;
  qm[0][IA] = AR; qm[0][IB] = B + dBx; qm[0][IC] = C + dCx; FLOOR(qm[0]);

  qp[1][ID] = r - dry; qp[1][IU] = u - duy; qp[1][IV] = v - dvy; qp[1][IW] = w - dwy; qp[1][IP] = p - dpy;
  qp[1][IA] = A - dAy; qp[1][IB] = BL; qp[1][IC] = C - dCy; FLOOR(qp[1]);
  qm[1][ID] = r + dry; qm[1][IU] = u + duy; qm[1][IV] = v + dvy; qm[1][IW] = w + dwy; qm[1][IP] = p + dpy;
  qm[1][IA] = A + dAy; qm[1][IB] = BR; qm[1][IC] = C + dCy; FLOOR(qm[1]);

  qp[2][ID] = r - drz; qp[2][IU] = u - duz; qp[2][IV] = v - dvz; qp[2][IW] = w - dwz; qp[2][IP] = p - dpz;
  qp[2][IA] = A - dAz; qp[2][IB] = B - dBz; qp[2][IC] = CL; FLOOR(qp[2]);
  qm[2][ID] = r + drz; qm[2][IU] = u + duz; qm[2][IV] = v + dvz; qm[2][IW] = w + dwz; qm[2][IP] = p + dpz;
  qm[2][IA] = A + dAz; qm[2][IB] = B + dBz; qm[2][IC] = CR; FLOOR(qm[2]);

  /* X edges */
  qRT_X[ID] = r + (+dry + drz); qRT_X[IU] = u + (+duy + duz); qRT_X[IV] = v + (+dvy + dvz); qRT_X[IW] = w + (+dwy + dwz);
  qRT_X[IP] = p + (+dpy + dpz); qRT_X[IA] = A + (+dAy + dAz); qRT_X[IB] = BR + (+dBRz); qRT_X[IC] = CR + (+dCRy); FLOOR(qRT_X);
  qRB_X[ID] = r + (+dry - drz); qRB_X[IU] = u + (+duy - duz); qRB_X[IV] = v + (+dvy - dvz); qRB_X[IW] = w + (+dwy - dwz);
  qRB_X[IP] = p + (+dpy - dpz); qRB_X[IA] = A + (+dAy - dAz); qRB_X[IB] = BR + (-dBRz); qRB_X[IC] = CL + (+dCLy); FLOOR(qRB_X);
  qLT_X[ID] = r + (-dry + drz); qLT_X[IU] = u + (-duy + duz); qLT_X[IV] = v + (-dvy + dvz); qLT_X[IW] = w + (-dwy + dwz);
  qLT_X[IP] = p + (-dpy + dpz); qLT_X[IA] = A + (-dAy + dAz); qLT_X[IB] = BL + (+dBLz); qLT_X[IC] = CR + (-dCRy); FLOOR(qLT_X);
  qLB_X[ID] = r + (-dry - drz); qLB_X[IU] = u + (-duy - duz); qLB_X[IV] = v + (-dvy - dvz); qLB_X[IW] = w + (-dwy - dwz);
  qLB_X[IP] = p + (-dpy - dpz); qLB_X[IA] = A + (-dAy - dAz); qLB_X[IB] = BL + (-dBLz); qLB_X[IC] = CL + (-dCLy); FLOOR(qLB_X);

  /* Y edges */
  qRT_Y[ID] = r + (+drx + drz); qRT_Y[IU] = u + (+dux + duz); qRT_Y[IV] = v + (+dvx + dvz); qRT_Y[IW] = w + (+dwx + dwz);
  qRT_Y[IP] = p + (+dpx + dpz); qRT_Y[IA] = AR + (+dARz); qRT_Y[IB] = B + (+dBx + dBz); qRT_Y[IC] = CR + (+dCRx); FLOOR(qRT_Y);
  qRB_Y[ID] = r + (+drx - drz); qRB_Y[IU] = u + (+dux - duz); qRB_Y[IV] = v + (+dvx - dvz); qRB_Y[IW] = w + (+dwx - dwz);
  qRB_Y[IP] = p + (+dpx - dpz); qRB_Y[IA] = AR + (-dARz); qRB_Y[IB] = B + (+dBx - dBz); qRB_Y[IC] = CL + (+dCLx); FLOOR(qRB_Y);
  qLT_Y[ID] = r + (-drx + drz); qLT_Y[IU] = u + (-dux + duz); qLT_Y[IV] = v + (-dvx + dvz); qLT_Y[IW] = w + (-dwx + dwz);
  qLT_Y[IP] = p + (-dpx + dpz); qLT_Y[IA] = AL + (+dALz); qLT_Y[IB] = B + (-dBx + dBz); qLT_Y[IC] = CR + (-dCRx); FLOOR(qLT_Y);
  qLB_Y[ID] = r + (-drx - drz); qLB_Y[IU] = u + (-dux - duz); qLB_Y[IV] = v + (-dvx - dvz); qLB_Y[IW] = w + (-dwx - dwz);
  qLB_Y[IP] = p + (-dpx - dpz); qLB_Y[IA] = AL + (-dALz); qLB_Y[IB] = B + (-dBx - dBz); qLB_Y[IC] = CL + (-dCLx); FLOOR(qLB_Y);

  /* Z edges */
  qRT_Z[ID] = r + (+drx + dry); qRT_Z[IU] = u + (+dux + duy); qRT_Z[IV] = v + (+dvx + dvy); qRT_Z[IW] = w + (+dwx + dwy);
  qRT_Z[IP] = p + (+dpx + dpy); qRT_Z[IA] = AR + (+dARy); qRT_Z[IB] = BR + (+dBRx); qRT_Z[IC] = C + (+dCx + dCy); FLOOR(qRT_Z);
  qRB_Z[ID] = r + (+drx - dry); qRB_Z[IU] = u + (+dux - duy); qRB_Z[IV] = v + (+dvx - dvy); qRB_Z[IW] = w + (+dwx - dwy);
  qRB_Z[IP] = p + (+dpx - dpy); qRB_Z[IA] = AR + (-dARy); qRB_Z[IB] = BL + (+dBLx); qRB_Z[IC] = C + (+dCx - dCy); FLOOR(qRB_Z);
  qLT_Z[ID] = r + (-drx + dry); qLT_Z[IU] = u + (-dux + duy); qLT_Z[IV] = v + (-dvx + dvy); qLT_Z[IW] = w + (-dwx + dwy);
  qLT_Z[IP] = p + (-dpx + dpy); qLT_Z[IA] = AL + (+dALy); qLT_Z[IB] = BR + (-dBRx); qLT_Z[IC] = C + (-dCx + dCy); FLOOR(qLT_Z);
  qLB_Z[ID] = r + (-drx - dry); qLB_Z[IU] = u + (-dux - duy); qLB_Z[IV] = v + (-dvx - dvy); qLB_Z[IW] = w + (-dwx - dwy);
  qLB_Z[IP] = p + (-dpx - dpy); qLB_Z[IA] = AL + (-dALy); qLB_Z[IB] = BL + (-dBLx); qLB_Z[IC] = C + (-dCx - dCy); FLOOR(qLB_Z);
#undef FLOOR
}

/* ------------------------------------------------------------------------------------------- */
/* muscl/MHDBaseFunctor2D.h:776-1109 trace_unsplit_mhd_2d (Omega0 == 0); qNb[di][dj], bfNb[di][dj][3] */
static void trace_unsplit_mhd_2d(const orc_params* pr_, double qNb[3][3][NV], double bfNb[4][4][3], double dtdx, double dtdy,
                                 double qm[2][NV], double qp[2][NV], double qEdge[4][NV])
{
  enum { CENTER = 1 };
  double* qRT = qEdge[0];
  double* qRB = qEdge[1];
  double* qLT = qEdge[2];
  double* qLB = qEdge[3];
  const double smallR = pr_->smallr, smallp = pr_->smallp, gamma = pr_->gamma0;
  const double* q = qNb[CENTER][CENTER];
  double Ez[2][2];
  for (int di = 0; di < 2; di++)
    for (int dj = 0; dj < 2; dj++) {
      int cx = CENTER + di, cy = CENTER + dj;
      double u = 0.25 * (qNb[cx - 1][cy - 1][IU] + qNb[cx - 1][cy][IU] + qNb[cx][cy - 1][IU] + qNb[cx][cy][IU]);
      double v = 0.25 * (qNb[cx - 1][cy - 1][IV] + qNb[cx - 1][cy][IV] + qNb[cx][cy - 1][IV] + qNb[cx][cy][IV]);
      double A = 0.5 * (bfNb[cx][cy - 1][0] + bfNb[cx][cy][0]);
      double B = 0.5 * (bfNb[cx - 1][cy][1] + bfNb[cx][cy][1]);
      Ez[di][dj] = u * B - v * A;
    }
  const double ELL = Ez[0][0], ELR = Ez[0][1], ERL = Ez[1][0], ERR = Ez[1][1];
  double r = q[ID], p = q[IP], u = q[IU], v = q[IV], w = q[IW], A = q[IA], B = q[IB], C = q[IC];
  double AL = bfNb[CENTER][CENTER][0], AR = bfNb[CENTER + 1][CENTER][0];
  double BL = bfNb[CENTER][CENTER][1], BR = bfNb[CENTER][CENTER + 1][1];

  /* slope_unsplit_hydro_2d (:347-390); slope_type other than 1, 2 leaves dq uninitialised in the reference: zero here */
  double dq[2][NV];
  const double st = pr_->slope_type;
  for (int n = 0; n < NV; ++n) {
    if (st == 1 || st == 2) {
      dq[0][n] = slope1(st, q[n], qNb[CENTER + 1][CENTER][n], qNb[CENTER - 1][CENTER][n]);
      dq[1][n] = slope1(st, q[n], qNb[CENTER][CENTER + 1][n], qNb[CENTER][CENTER - 1][n]);
    } else {
      dq[0][n] = dq[1][n] = 0.0;
    }
  }
  double drx = dq[0][ID] * 0.5, dpx = dq[0][IP] * 0.5, dux = dq[0][IU] * 0.5, dvx = dq[0][IV] * 0.5, dwx = dq[0][IW] * 0.5;
  double dCx = dq[0][IC] * 0.5, dBx = dq[0][IB] * 0.5;
  double dry = dq[1][ID] * 0.5, dpy = dq[1][IP] * 0.5, duy = dq[1][IU] * 0.5, dvy = dq[1][IV] * 0.5, dwy = dq[1][IW] * 0.5;
  double dCy = dq[1][IC] * 0.5, dAy = dq[1][IA] * 0.5;

  /* slope_unsplit_mhd_2d (:456-509): slope_type used as is */
  double dALy = 0.5 * slope1(st, bfNb[CENTER][CENTER][0], bfNb[CENTER][CENTER + 1][0], bfNb[CENTER][CENTER - 1][0]);
  double dBLx = 0.5 * slope1(st, bfNb[CENTER][CENTER][1], bfNb[CENTER + 1][CENTER][1], bfNb[CENTER - 1][CENTER][1]);
  double dARy = 0.5 * slope1(st, bfNb[CENTER + 1][CENTER][0], bfNb[CENTER + 1][CENTER + 1][0], bfNb[CENTER + 1][CENTER - 1][0]);
  double dBRx = 0.5 * slope1(st, bfNb[CENTER][CENTER + 1][1], bfNb[CENTER + 1][CENTER + 1][1], bfNb[CENTER - 1][CENTER + 1][1]);

  double dAx = 0.5 * (AR - AL);
  double dBy = 0.5 * (BR - BL);
  double sr0, su0, sv0, sw0, sp0, sA0, sB0, sC0, sAL0, sAR0, sBL0, sBR0;
  sr0 = (-u * drx - dux * r) * dtdx + (-v * dry - dvy * r) * dtdy;
  su0 = (-u * dux - dpx / r - B * dBx / r - C * dCx / r) * dtdx + (-v * duy + B * dAy / r) * dtdy;
  sv0 = (-u * dvx + A * dBx / r) * dtdx + (-v * dvy - dpy / r - A * dAy / r - C * dCy / r) * dtdy;
  sw0 = (-u * dwx + A * dCx / r) * dtdx + (-v * dwy + B * dCy / r) * dtdy;
  sp0 = (-u * dpx - dux * gamma * p) * dtdx + (-v * dpy - dvy * gamma * p) * dtdy;
  sA0 = (u * dBy + B * duy - v * dAy - A * dvy) * dtdy;
  sB0 = (-u * dBx - B * dux + v * dAx + A * dvx) * dtdx;
  sC0 = (w * dAx + A * dwx - u * dCx - C * dux) * dtdx + (-v * dCy - C * dvy + w * dBy + B * dwy) * dtdy;
  sAL0 = +(ELR - ELL) * 0.5 * dtdy;
  sAR0 = +(ERR - ERL) * 0.5 * dtdy;
  sBL0 = -(ERL - ELL) * 0.5 * dtdx;
  sBR0 = -(ERR - ELR) * 0.5 * dtdx;

  r = r + sr0; u = u + su0; v = v + sv0; w = w + sw0; p = p + sp0; A = A + sA0; B = B + sB0; C = C + sC0;
  AL = AL + sAL0; AR = AR + sAR0; BL = BL + sBL0; BR = BR + sBR0;

#define FLOOR2(s)                       \
  (s)[ID] = fmax(smallR, (s)[ID]);      \
  (s)[IP] = fmax(smallp * (s)[ID], (s)[IP])
  qp[0][ID] = r - drx; qp[0][IU] = u - dux; qp[0][IV] = v - dvx; qp[0][IW] = w - dwx; qp[0][IP] = p - dpx;
  qp[0][IA] = AL; qp[0][IB] = B - dBx; qp[0][IC] = C - dCx; FLOOR2(qp[0]);
  qm[0][ID] = r + drx; qm[0][IU] = u + dux; qm[0][IV] = v + dvx; qm[0][IW] = w + dwx; qm[0][IP] = p + dpx;
  qm[0][IA] = AR; qm[0][IB] = B + dBx; qm[0][IC] = C + dCx; FLOOR2(qm[0]);
  qp[1][ID] = r - dry; qp[1][IU] = u - duy; qp[1][IV] = v - dvy; qp[1][IW] = w - dwy; qp[1][IP] = p - dpy;
  qp[1][IA] = A - dAy; qp[1][IB] = BL; qp[1][IC] = C - dCy; FLOOR2(qp[1]);
  qm[1][ID] = r + dry; qm[1][IU] = u + duy; qm[1][IV] = v + dvy; qm[1][IW] = w + dwy; qm[1][IP] = p + dpy;
  qm[1][IA] = A + dAy; qm[1][IB] = BR; qm[1][IC] = C + dCy; FLOOR2(qm[1]);

  qRT[ID] = r + (+drx + dry); qRT[IU] = u + (+dux + duy); qRT[IV] = v + (+dvx + dvy); qRT[IW] = w + (+dwx + dwy);
  qRT[IP] = p + (+dpx + dpy); qRT[IA] = AR + (+dARy); qRT[IB] = BR + (+dBRx); qRT[IC] = C + (+dCx + dCy); FLOOR2(qRT);
  qRB[ID] = r + (+drx - dry); qRB[IU] = u + (+dux - duy); qRB[IV] = v + (+dvx - dvy); qRB[IW] = w + (+dwx - dwy);
  qRB[IP] = p + (+dpx - dpy); qRB[IA] = AR + (-dARy); qRB[IB] = BL + (+dBLx); qRB[IC] = C + (+dCx - dCy); FLOOR2(qRB);
  qLB[ID] = r + (-drx - dry); qLB[IU] = u + (-dux - duy); qLB[IV] = v + (-dvx - dvy); qLB[IW] = w + (-dwx - dwy);
  qLB[IP] = p + (-dpx - dpy); qLB[IA] = AL + (-dALy); qLB[IB] = BL + (-dBLx); qLB[IC] = C + (-dCx - dCy); FLOOR2(qLB);
  qLT[ID] = r + (-drx + dry); qLT[IU] = u + (-dux + duy); qLT[IV] = v + (-dvx + dvy); qLT[IW] = w + (-dwx + dwy);
  qLT[IP] = p + (-dpx + dpy); qLT[IA] = AL + (+dALy); qLT[IB] = BR + (-dBRx); qLT[IC] = C + (-dCx + dCy); FLOOR2(qLT);
#undef FLOOR2
}

/* muscl/SolverMHDMuscl.cpp:371-454 godunov_unsplit_impl<2>, implementationVersion 0 */
static void mhd_step_2d(const orc_params* p, double* Uin, double* Uout, double dt, double t, double t_end, double* dt_used)
{
  const geom_t g = geom(p);
  const int gw = p->ghost;
  const int isize = g.isize, jsize = g.jsize;
  const size_t ncell = g.sv, nstate = g.sv * NV;

  orc_make_boundaries(p, Uin);
  memcpy(Uout, Uin, nstate * sizeof(double));
  double* Q = calloc(nstate, sizeof(double));
  orc_mhd_prims(p, Uin, Q);
  if (dt < 0) {
    dt = p->cfl / orc_mhd_invdt(p, Q);
    if (t + dt > t_end)
      dt = t_end - t;
  }
  if (dt_used)
    *dt_used = dt;
  const double dtdx = dt / p->dx, dtdy = dt / p->dy;

  /* ComputeTraceFunctor2D_MHD, muscl/MHDRunFunctors2D.h:782-905 */
  double* Qm[2];
  double* Qp[2];
  double* QE[4];
  for (int d = 0; d < 2; ++d) {
    Qm[d] = calloc(nstate, sizeof(double));
    Qp[d] = calloc(nstate, sizeof(double));
  }
  for (int e = 0; e < 4; ++e)
    QE[e] = calloc(nstate, sizeof(double));
  for (int j = gw - 2; j < jsize - gw + 1; ++j)
    for (int i = gw - 2; i < isize - gw + 1; ++i) {
      double qNb[3][3][NV], bfNb[4][4][3], qm[2][NV], qp[2][NV], qEdge[4][NV];
      for (int di = 0; di < 3; di++)
        for (int dj = 0; dj < 3; dj++)
          get_state(g, Q, i + di - 1, j + dj - 1, 0, qNb[di][dj]);
      for (int di = 0; di < 4; di++)
        for (int dj = 0; dj < 4; dj++) {
          bfNb[di][dj][0] = AT(Uin, i + di - 1, j + dj - 1, 0, IA);
          bfNb[di][dj][1] = AT(Uin, i + di - 1, j + dj - 1, 0, IB);
          bfNb[di][dj][2] = AT(Uin, i + di - 1, j + dj - 1, 0, IC);
        }
      trace_unsplit_mhd_2d(p, qNb, bfNb, dtdx, dtdy, qm, qp, qEdge);
      for (int d = 0; d < 2; ++d) {
        set_state(g, Qm[d], i, j, 0, qm[d]);
        set_state(g, Qp[d], i, j, 0, qp[d]);
      }
      for (int e = 0; e < 4; ++e)
        set_state(g, QE[e], i, j, 0, qEdge[e]);
    }

  /* ComputeFluxesAndStoreFunctor2D_MHD (:375-470) + ComputeEmfAndStoreFunctor2D (:612-686) */
  double* F[2];
  for (int d = 0; d < 2; ++d)
    F[d] = calloc(nstate, sizeof(double));
  double* Emf = calloc(ncell, sizeof(double));
  for (int j = gw; j < jsize - gw + 1; ++j)
    for (int i = gw; i < isize - gw + 1; ++i) {
      double ql[NV], qr[NV], flux[NV];
      memset(flux, 0, sizeof flux);
      get_state(g, Qm[0], i - 1, j, 0, ql);
      get_state(g, Qp[0], i, j, 0, qr);
      riemann_mhd(p, ql, qr, flux);
      set_state(g, F[0], i, j, 0, flux);
      get_state(g, Qm[1], i, j - 1, 0, ql);
      swapv(&ql[IU], &ql[IV]);
      swapv(&ql[IA], &ql[IB]);
      get_state(g, Qp[1], i, j, 0, qr);
      swapv(&qr[IU], &qr[IV]);
      swapv(&qr[IA], &qr[IB]);
      riemann_mhd(p, ql, qr, flux);
      set_state(g, F[1], i, j, 0, flux);

      double qe[4][NV];
      get_state(g, QE[IRT], i - 1, j - 1, 0, qe[IRT]);
      get_state(g, QE[IRB], i - 1, j, 0, qe[IRB]);
      get_state(g, QE[ILT], i, j - 1, 0, qe[ILT]);
      get_state(g, QE[ILB], i, j, 0, qe[ILB]);
      AT(Emf, i, j, 0, 0) = compute_emf(p, EMFZ, qe);
    }

  /* UpdateFunctor2D_MHD (:1539-1640) then UpdateEmfFunctor2D (:1646-1700) */
  for (int j = gw; j < jsize - gw; ++j)
    for (int i = gw; i < isize - gw; ++i) {
      double u[NV], f[NV];
      get_state(g, Uout, i, j, 0, u);
      get_state(g, F[0], i, j, 0, f);
      u[ID] += f[ID] * dtdx; u[IP] += f[IP] * dtdx; u[IU] += f[IU] * dtdx; u[IV] += f[IV] * dtdx; u[IW] += f[IW] * dtdx; u[IC] += f[IC] * dtdx;
      get_state(g, F[0], i + 1, j, 0, f);
      u[ID] -= f[ID] * dtdx; u[IP] -= f[IP] * dtdx; u[IU] -= f[IU] * dtdx; u[IV] -= f[IV] * dtdx; u[IW] -= f[IW] * dtdx; u[IC] -= f[IC] * dtdx;
      get_state(g, F[1], i, j, 0, f);
      u[ID] += f[ID] * dtdy; u[IP] += f[IP] * dtdy; u[IU] += f[IV] * dtdy; u[IV] += f[IU] * dtdy; u[IW] += f[IW] * dtdy; u[IC] += f[IC] * dtdy;
      get_state(g, F[1], i, j + 1, 0, f);
      u[ID] -= f[ID] * dtdy; u[IP] -= f[IP] * dtdy; u[IU] -= f[IV] * dtdy; u[IV] -= f[IU] * dtdy; u[IW] -= f[IW] * dtdy; u[IC] -= f[IC] * dtdy;
      u[IA] += (AT(Emf, i, j + 1, 0, 0) - AT(Emf, i, j, 0, 0)) * dtdy;
      u[IB] -= (AT(Emf, i + 1, j, 0, 0) - AT(Emf, i, j, 0, 0)) * dtdx;
      set_state(g, Uout, i, j, 0, u);
    }

  free(Q); free(Emf);
  for (int d = 0; d < 2; ++d) {
    free(Qm[d]); free(Qp[d]); free(F[d]);
  }
  for (int e = 0; e < 4; ++e)
    free(QE[e]);
}

/* ------------------------------------------------------------------------------------------- */
/* muscl/SolverMHDMuscl.cpp:463-513 godunov_unsplit_impl<3>, implementationVersion 0 */
void orc_mhd_step(const orc_params* p, double* Uin, double* Uout, double dt, double t, double t_end, double* dt_used)
{
  if (p->dim == 2) {
    mhd_step_2d(p, Uin, Uout, dt, t, t_end, dt_used);
    return;
  }
  const geom_t g = geom(p);
  const int gw = p->ghost;
  const int isize = g.isize, jsize = g.jsize, ksize = g.ksize;
  const size_t ncell = g.sv, nstate = g.sv * NV;

  orc_make_boundaries(p, Uin);
  memcpy(Uout, Uin, nstate * sizeof(double));

  double* Q = calloc(nstate, sizeof(double));
  orc_mhd_prims(p, Uin, Q);

  if (dt < 0) { /* SolverBase::compute_dt, shared/SolverBase.cpp:125-154 */
    dt = p->cfl / orc_mhd_invdt(p, Q);
    if (t + dt > t_end)
      dt = t_end - t;
  }
  if (dt_used)
    *dt_used = dt;
  const double dtdx = dt / p->dx, dtdy = dt / p->dy, dtdz = dt / p->dz;

  /* geometry of 3-component arrays: same box, 3 "variables" */
  double* Elec = calloc(ncell * 3, sizeof(double));
  double* DeltaA = calloc(ncell * 3, sizeof(double));
  double* DeltaB = calloc(ncell * 3, sizeof(double));
  double* DeltaC = calloc(ncell * 3, sizeof(double));

  /* ComputeElecFieldFunctor3D, muscl/MHDRunFunctors3D.h:278-362 */
  for (int k = 1; k < ksize - 1; ++k)
    for (int j = 1; j < jsize - 1; ++j)
      for (int i = 1; i < isize - 1; ++i) {
        double u, v, w, A, B, C;
        v = 0.25 * (AT(Q, i, j - 1, k - 1, IV) + AT(Q, i, j - 1, k, IV) + AT(Q, i, j, k - 1, IV) + AT(Q, i, j, k, IV));
        w = 0.25 * (AT(Q, i, j - 1, k - 1, IW) + AT(Q, i, j - 1, k, IW) + AT(Q, i, j, k - 1, IW) + AT(Q, i, j, k, IW));
        B = 0.5 * (AT(Uin, i, j, k - 1, IB) + AT(Uin, i, j, k, IB));
        C = 0.5 * (AT(Uin, i, j - 1, k, IC) + AT(Uin, i, j, k, IC));
        AT(Elec, i, j, k, 0) = v * C - w * B;
        u = 0.25 * (AT(Q, i - 1, j, k - 1, IU) + AT(Q, i - 1, j, k, IU) + AT(Q, i, j, k - 1, IU) + AT(Q, i, j, k, IU));
        w = 0.25 * (AT(Q, i - 1, j, k - 1, IW) + AT(Q, i - 1, j, k, IW) + AT(Q, i, j, k - 1, IW) + AT(Q, i, j, k, IW));
        A = 0.5 * (AT(Uin, i, j, k - 1, IA) + AT(Uin, i, j, k, IA));
        C = 0.5 * (AT(Uin, i - 1, j, k, IC) + AT(Uin, i, j, k, IC));
        AT(Elec, i, j, k, 1) = w * A - u * C;
        u = 0.25 * (AT(Q, i - 1, j - 1, k, IU) + AT(Q, i - 1, j, k, IU) + AT(Q, i, j - 1, k, IU) + AT(Q, i, j, k, IU));
        v = 0.25 * (AT(Q, i - 1, j - 1, k, IV) + AT(Q, i - 1, j, k, IV) + AT(Q, i, j - 1, k, IV) + AT(Q, i, j, k, IV));
        A = 0.5 * (AT(Uin, i, j - 1, k, IA) + AT(Uin, i, j, k, IA));
        B = 0.5 * (AT(Uin, i - 1, j, k, IB) + AT(Uin, i, j, k, IB));
        AT(Elec, i, j, k, 2) = u * B - v * A;
      }

  /* ComputeMagSlopesFunctor3D + slope_unsplit_mhd_3d, muscl/MHDRunFunctors3D.h:441-538,
   * muscl/MHDBaseFunctor3D.h:563-668 */
  {
    const double xst = fmin(p->slope_type, 2.0);
    for (int k = 1; k < ksize - 1; ++k)
      for (int j = 1; j < jsize - 1; ++j)
        for (int i = 1; i < isize - 1; ++i) {
          const double bfx = AT(Uin, i, j, k, IA), bfy = AT(Uin, i, j, k, IB), bfz = AT(Uin, i, j, k, IC);
          /* DeltaA(.,0) = dbfX[IX] = 0 (never set), DeltaA(.,1) = dbfY[IX], DeltaA(.,2) = dbfZ[IX] */
          AT(DeltaA, i, j, k, 0) = 0.0;
          AT(DeltaA, i, j, k, 1) = slope1(xst, bfx, AT(Uin, i, j + 1, k, IA), AT(Uin, i, j - 1, k, IA));
          AT(DeltaA, i, j, k, 2) = slope1(xst, bfx, AT(Uin, i, j, k + 1, IA), AT(Uin, i, j, k - 1, IA));
          AT(DeltaB, i, j, k, 0) = slope1(xst, bfy, AT(Uin, i + 1, j, k, IB), AT(Uin, i - 1, j, k, IB));
          AT(DeltaB, i, j, k, 1) = 0.0;
          AT(DeltaB, i, j, k, 2) = slope1(xst, bfy, AT(Uin, i, j, k + 1, IB), AT(Uin, i, j, k - 1, IB));
          AT(DeltaC, i, j, k, 0) = slope1(xst, bfz, AT(Uin, i + 1, j, k, IC), AT(Uin, i - 1, j, k, IC));
          AT(DeltaC, i, j, k, 1) = slope1(xst, bfz, AT(Uin, i, j + 1, k, IC), AT(Uin, i, j - 1, k, IC));
          AT(DeltaC, i, j, k, 2) = 0.0;
        }
  }

  /* ComputeTraceFunctor3D_MHD, muscl/MHDRunFunctors3D.h:543-856 */
  double* Qm[3];
  double* Qp[3];
  double* QE[4][3];
  for (int d = 0; d < 3; ++d) {
    Qm[d] = calloc(nstate, sizeof(double));
    Qp[d] = calloc(nstate, sizeof(double));
    for (int e = 0; e < 4; ++e)
      QE[e][d] = calloc(nstate, sizeof(double));
  }
  for (int k = gw - 2; k < ksize - gw + 1; ++k)
    for (int j = gw - 2; j < jsize - gw + 1; ++j)
      for (int i = gw - 2; i < isize - gw + 1; ++i) {
        double q[NV], qpx[NV], qmx[NV], qpy[NV], qmy[NV], qpz[NV], qmz[NV], dq[3][NV];
        double bfNb[6], dbf[12], ef[3][2][2], qm[3][NV], qp[3][NV], qEdge[4][3][NV];
        get_state(g, Q, i, j, k, q);
        get_state(g, Q, i + 1, j, k, qpx);
        get_state(g, Q, i - 1, j, k, qmx);
        get_state(g, Q, i, j + 1, k, qpy);
        get_state(g, Q, i, j - 1, k, qmy);
        get_state(g, Q, i, j, k + 1, qpz);
        get_state(g, Q, i, j, k - 1, qmz);
        slope_unsplit_hydro_3d(p, q, qpx, qmx, qpy, qmy, qpz, qmz, dq);
        bfNb[0] = AT(Uin, i, j, k, IA);
        bfNb[1] = AT(Uin, i + 1, j, k, IA);
        bfNb[2] = AT(Uin, i, j, k, IB);
        bfNb[3] = AT(Uin, i, j + 1, k, IB);
        bfNb[4] = AT(Uin, i, j, k, IC);
        bfNb[5] = AT(Uin, i, j, k + 1, IC);
        dbf[0] = AT(DeltaA, i, j, k, 1);
        dbf[1] = AT(DeltaA, i, j, k, 2);
        dbf[2] = AT(DeltaB, i, j, k, 0);
        dbf[3] = AT(DeltaB, i, j, k, 2);
        dbf[4] = AT(DeltaC, i, j, k, 0);
        dbf[5] = AT(DeltaC, i, j, k, 1);
        dbf[6] = AT(DeltaA, i + 1, j, k, 1);
        dbf[7] = AT(DeltaA, i + 1, j, k, 2);
        dbf[8] = AT(DeltaB, i, j + 1, k, 0);
        dbf[9] = AT(DeltaB, i, j + 1, k, 2);
        dbf[10] = AT(DeltaC, i, j, k + 1, 0);
        dbf[11] = AT(DeltaC, i, j, k + 1, 1);
        ef[0][0][0] = AT(Elec, i, j, k, 0);
        ef[0][0][1] = AT(Elec, i, j, k + 1, 0);
        ef[0][1][0] = AT(Elec, i, j + 1, k, 0);
        ef[0][1][1] = AT(Elec, i, j + 1, k + 1, 0);
        ef[1][0][0] = AT(Elec, i, j, k, 1);
        ef[1][0][1] = AT(Elec, i, j, k + 1, 1);
        ef[1][1][0] = AT(Elec, i + 1, j, k, 1);
        ef[1][1][1] = AT(Elec, i + 1, j, k + 1, 1);
        ef[2][0][0] = AT(Elec, i, j, k, 2);
        ef[2][0][1] = AT(Elec, i, j + 1, k, 2);
        ef[2][1][0] = AT(Elec, i + 1, j, k, 2);
        ef[2][1][1] = AT(Elec, i + 1, j + 1, k, 2);
        trace_unsplit_mhd_3d_simpler(p, q, dq, bfNb, dbf, (const double(*)[2][2])ef, dtdx, dtdy, dtdz, qm, qp, qEdge);
        for (int d = 0; d < 3; ++d) {
          set_state(g, Qm[d], i, j, k, qm[d]);
          set_state(g, Qp[d], i, j, k, qp[d]);
          for (int e = 0; e < 4; ++e)
            set_state(g, QE[e][d], i, j, k, qEdge[e][d]);
        }
      }

  /* ComputeFluxesAndStoreFunctor3D_MHD, muscl/MHDRunFunctors3D.h:1783-1909 (raw flux, swaps not undone) */
  double* F[3];
  for (int d = 0; d < 3; ++d)
    F[d] = calloc(nstate, sizeof(double));
  double* Emf = calloc(ncell * 3, sizeof(double)); /* I_EMFZ = 0, I_EMFY = 1, I_EMFX = 2 (enums.h:220-222) */
  for (int k = gw; k < ksize - gw + 1; ++k)
    for (int j = gw; j < jsize - gw + 1; ++j)
      for (int i = gw; i < isize - gw + 1; ++i) {
        double ql[NV], qr[NV], flux[NV];
        memset(flux, 0, sizeof flux);
        get_state(g, Qm[0], i - 1, j, k, ql);
        get_state(g, Qp[0], i, j, k, qr);
        riemann_mhd(p, ql, qr, flux);
        set_state(g, F[0], i, j, k, flux);

        get_state(g, Qm[1], i, j - 1, k, ql);
        swapv(&ql[IU], &ql[IV]);
        swapv(&ql[IA], &ql[IB]);
        get_state(g, Qp[1], i, j, k, qr);
        swapv(&qr[IU], &qr[IV]);
        swapv(&qr[IA], &qr[IB]);
        riemann_mhd(p, ql, qr, flux);
        set_state(g, F[1], i, j, k, flux);

        get_state(g, Qm[2], i, j, k - 1, ql);
        swapv(&ql[IU], &ql[IW]);
        swapv(&ql[IA], &ql[IC]);
        get_state(g, Qp[2], i, j, k, qr);
        swapv(&qr[IU], &qr[IW]);
        swapv(&qr[IA], &qr[IC]);
        riemann_mhd(p, ql, qr, flux);
        set_state(g, F[2], i, j, k, flux);

        /* ComputeEmfAndStoreFunctor3D, muscl/MHDRunFunctors3D.h:2100-2238 */
        double qe[4][NV];
        get_state(g, QE[IRT][2], i - 1, j - 1, k, qe[IRT]);
        get_state(g, QE[IRB][2], i - 1, j, k, qe[IRB]);
        get_state(g, QE[ILT][2], i, j - 1, k, qe[ILT]);
        get_state(g, QE[ILB][2], i, j, k, qe[ILB]);
        AT(Emf, i, j, k, 0) = compute_emf(p, EMFZ, qe);
        get_state(g, QE[IRT][1], i - 1, j, k - 1, qe[IRT]);
        get_state(g, QE[ILT][1], i, j, k - 1, qe[IRB]);
        get_state(g, QE[IRB][1], i - 1, j, k, qe[ILT]);
        get_state(g, QE[ILB][1], i, j, k, qe[ILB]);
        AT(Emf, i, j, k, 1) = compute_emf(p, EMFY, qe);
        get_state(g, QE[IRT][0], i, j - 1, k - 1, qe[IRT]);
        get_state(g, QE[IRB][0], i, j - 1, k, qe[IRB]);
        get_state(g, QE[ILT][0], i, j, k - 1, qe[ILT]);
        get_state(g, QE[ILB][0], i, j, k, qe[ILB]);
        AT(Emf, i, j, k, 2) = compute_emf(p, EMFX, qe);
      }

  /* UpdateFunctor3D_MHD (:2430-2544) then UpdateEmfFunctor3D (:2549-2628), interior cells */
  for (int k = gw; k < ksize - gw; ++k)
    for (int j = gw; j < jsize - gw; ++j)
      for (int i = gw; i < isize - gw; ++i) {
        double u[NV], f[NV];
        get_state(g, Uout, i, j, k, u);
        get_state(g, F[0], i, j, k, f);
        u[ID] += f[ID] * dtdx; u[IP] += f[IP] * dtdx; u[IU] += f[IU] * dtdx; u[IV] += f[IV] * dtdx; u[IW] += f[IW] * dtdx;
        get_state(g, F[0], i + 1, j, k, f);
        u[ID] -= f[ID] * dtdx; u[IP] -= f[IP] * dtdx; u[IU] -= f[IU] * dtdx; u[IV] -= f[IV] * dtdx; u[IW] -= f[IW] * dtdx;
        get_state(g, F[1], i, j, k, f);
        u[ID] += f[ID] * dtdy; u[IP] += f[IP] * dtdy; u[IU] += f[IV] * dtdy; u[IV] += f[IU] * dtdy; u[IW] += f[IW] * dtdy;
        get_state(g, F[1], i, j + 1, k, f);
        u[ID] -= f[ID] * dtdy; u[IP] -= f[IP] * dtdy; u[IU] -= f[IV] * dtdy; u[IV] -= f[IU] * dtdy; u[IW] -= f[IW] * dtdy;
        get_state(g, F[2], i, j, k, f);
        u[ID] += f[ID] * dtdz; u[IP] += f[IP] * dtdz; u[IU] += f[IW] * dtdz; u[IV] += f[IV] * dtdz; u[IW] += f[IU] * dtdz;
        get_state(g, F[2], i, j, k + 1, f);
        u[ID] -= f[ID] * dtdz; u[IP] -= f[IP] * dtdz; u[IU] -= f[IW] * dtdz; u[IV] -= f[IV] * dtdz; u[IW] -= f[IU] * dtdz;

        u[IA] += (AT(Emf, i, j + 1, k, 0) - AT(Emf, i, j, k, 0)) * dtdy;
        u[IB] -= (AT(Emf, i + 1, j, k, 0) - AT(Emf, i, j, k, 0)) * dtdx;
        u[IA] -= (AT(Emf, i, j, k + 1, 1) - AT(Emf, i, j, k, 1)) * dtdz;
        u[IB] += (AT(Emf, i, j, k + 1, 2) - AT(Emf, i, j, k, 2)) * dtdz;
        u[IC] += (AT(Emf, i + 1, j, k, 1) - AT(Emf, i, j, k, 1)) * dtdx;
        u[IC] -= (AT(Emf, i, j + 1, k, 2) - AT(Emf, i, j, k, 2)) * dtdy;
        set_state(g, Uout, i, j, k, u);
      }

  free(Q); free(Elec); free(DeltaA); free(DeltaB); free(DeltaC); free(Emf);
  for (int d = 0; d < 3; ++d) {
    free(Qm[d]); free(Qp[d]); free(F[d]);
    for (int e = 0; e < 4; ++e)
      free(QE[e][d]);
  }
}
