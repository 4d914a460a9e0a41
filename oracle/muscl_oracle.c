/*
 * TEST INFRASTRUCTURE ONLY (see muscl_oracle.h) -- hydro part of the CPU restatement.
 *
 * Every function follows the reference functor it cites, in the same floating-point order,
 * cell by cell ("gather" form, exactly like the reference's implementationVersion 0).
 * Citations are relative to /root/reference/src.
 */
#include "muscl_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { ID = 0, IP = 1, IU = 2, IV = 3, IW = 4 };
enum { RIEMANN_APPROX = 0, RIEMANN_LLF = 1, RIEMANN_HLL = 2, RIEMANN_HLLC = 3, RIEMANN_HLLD = 4 };
enum { BC_DIRICHLET = 1, BC_NEUMANN = 2, BC_PERIODIC = 3, BC_COPY = 4 };

typedef struct {
  int isize, jsize, ksize;
  size_t sj, sk, sv; /* strides of j, k, var */
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

/* ------------------------------------------------------------------------------------------- */
/* shared/BoundariesFunctors.h:17-523 (hydro) and :534-1050 (MHD: normal B component flips too) */
static void fill_face(const orc_params* p, double* U, int face)
{
  const geom_t g = geom(p);
  const int gw = p->ghost, nbvar = p->nbvar;
  const int n[3] = { p->nx, p->ny, p->nz };
  const int dir = face / 2, side = face % 2;
  const int bt = p->bc[face];
  if (bt == BC_COPY)
    return; /* internal face: filled by the halo exchange, not by a physical BC */
  const int size[3] = { g.isize, g.jsize, g.ksize };
  const int ivel = IU + dir;     /* normal momentum */
  const int imag = 5 + dir;      /* normal magnetic field (MHD) */
  int lo[3] = { 0, 0, 0 }, hi[3] = { size[0], size[1], size[2] };
  lo[dir] = side ? n[dir] + gw : 0;
  hi[dir] = side ? n[dir] + 2 * gw : gw;
  for (int k = lo[2]; k < hi[2]; ++k)
    for (int j = lo[1]; j < hi[1]; ++j)
      for (int i = lo[0]; i < hi[0]; ++i) {
        int c[3] = { i, j, k };
        int c0[3] = { i, j, k };
        const int x = c[dir];
        int x0;
        if (bt == BC_DIRICHLET)
          x0 = side ? 2 * n[dir] + 2 * gw - 1 - x : 2 * gw - 1 - x;
        else if (bt == BC_NEUMANN)
          x0 = side ? n[dir] + gw - 1 : gw;
        else /* periodic (and any other value, as in the reference) */
          x0 = side ? x - n[dir] : n[dir] + x;
        c0[dir] = x0;
        for (int v = 0; v < nbvar; ++v) {
          double sign = 1.0;
          if (bt == BC_DIRICHLET && (v == ivel || (p->mhd && v == imag)))
            sign = -1.0;
          AT(U, c[0], c[1], c[2], v) = AT(U, c0[0], c0[1], c0[2], v) * sign;
        }
      }
}

/* shared/SolverBase.cpp:464-512 make_boundaries_serial: XMIN,XMAX,YMIN,YMAX[,ZMIN,ZMAX] */
void orc_make_boundaries(const orc_params* p, double* U)
{
  const int nfaces = (p->dim == 3) ? 6 : 4;
  for (int f = 0; f < nfaces; ++f)
    fill_face(p, U, f);
}

/* ------------------------------------------------------------------------------------------- */
/* muscl/HydroBaseFunctor3D.h:81-107, HydroBaseFunctor2D.h:81-105 */
static void compute_primitives(const orc_params* p, const double* u, double* c, double* q)
{
  const double gamma0 = p->gamma0, smallr = p->smallr, smallp = p->smallp;
  double d = fmax(u[ID], smallr);
  double ux = u[IU] / d;
  double uy = u[IV] / d;
  double eken;
  double uz = 0.0;
  if (p->dim == 3) {
    uz = u[IW] / d;
    eken = 0.5 * (ux * ux + uy * uy + uz * uz);
  } else {
    eken = 0.5 * (ux * ux + uy * uy);
  }
  double e = u[IP] / d - eken;
  double pr = fmax((gamma0 - 1.0) * d * e, d * smallp);
  *c = sqrt(gamma0 * pr / d);
  q[ID] = d;
  q[IP] = pr;
  q[IU] = ux;
  q[IV] = uy;
  if (p->dim == 3)
    q[IW] = uz;
}

/* muscl/HydroRunFunctors3D.h:200-265 (whole box, ghosts included) */
void orc_hydro_prims(const orc_params* p, const double* U, double* Q)
{
  const geom_t g = geom(p);
  for (int k = 0; k < g.ksize; ++k)
    for (int j = 0; j < g.jsize; ++j)
      for (int i = 0; i < g.isize; ++i) {
        double u[5], q[5], c;
        for (int v = 0; v < p->nbvar; ++v)
          u[v] = AT(U, i, j, k, v);
        compute_primitives(p, u, &c, q);
        for (int v = 0; v < p->nbvar; ++v)
          AT(Q, i, j, k, v) = q[v];
      }
}

/* muscl/HydroRunFunctors3D.h:16-195, HydroRunFunctors2D.h:18-180 */
double orc_hydro_invdt(const orc_params* p, const double* U, const double* grav)
{
  const geom_t g = geom(p);
  const int gw = p->ghost;
  const int k0 = (p->dim == 3) ? gw : 0, k1 = (p->dim == 3) ? g.ksize - gw : 1;
  double invDt = 0.0;
  for (int k = k0; k < k1; ++k)
    for (int j = gw; j < g.jsize - gw; ++j)
      for (int i = gw; i < g.isize - gw; ++i) {
        double u[5], q[5], c;
        for (int v = 0; v < p->nbvar; ++v)
          u[v] = AT(U, i, j, k, v);
        compute_primitives(p, u, &c, q);
        if (!p->gravity_enabled) {
          double vx = c + fabs(q[IU]);
          double vy = c + fabs(q[IV]);
          if (p->dim == 3) {
            double vz = c + fabs(q[IW]);
            invDt = fmax(invDt, vx / p->dx + vy / p->dy + vz / p->dz);
          } else {
            invDt = fmax(invDt, vx / p->dx + vy / p->dy);
          }
        } else {
          double dx = fmin(p->dx, p->dy);
          if (p->dim == 3)
            dx = fmin(dx, p->dz);
          double velocity = 0.0;
          velocity += c + fabs(q[IU]);
          velocity += c + fabs(q[IV]);
          if (p->dim == 3)
            velocity += c + fabs(q[IW]);
          double kk = fabs(AT(grav, i, j, k, 0)) + fabs(AT(grav, i, j, k, 1));
          if (p->dim == 3)
            kk = fabs(AT(grav, i, j, k, 0)) + fabs(AT(grav, i, j, k, 1)) + fabs(AT(grav, i, j, k, 2));
          kk *= p->cfl * dx / (velocity * velocity);
          kk = fmax(kk, 1e-4);
          velocity *= kk / (sqrt(1.0 + 2.0 * kk) - 1.0);
          invDt = fmax(invDt, velocity / dx);
        }
      }
  return invDt;
}

/* ------------------------------------------------------------------------------------------- */
/* muscl/HydroBaseFunctor3D.h:413-461 (one direction of slope_unsplit_hydro_3d_scalar) */
static double slope_scalar(double slope_type, double q, double qPlus, double qMinus)
{
  double dlft = slope_type * (q - qMinus);
  double drgt = slope_type * (qPlus - q);
  double dcen = 0.5 * (qPlus - qMinus);
  double dsgn = (dcen >= 0.0) ? 1.0 : -1.0;
  double slop = fmin(fabs(dlft), fabs(drgt));
  double dlim = slop;
  if ((dlft * drgt) <= 0.0)
    dlim = 0.0;
  return dsgn * fmin(dlim, fabs(dcen));
}

/* slopes of all variables of cell (i,j,k) along the dim directions:
 * muscl/HydroBaseFunctor3D.h:482-578, HydroBaseFunctor2D.h:476-... */
static void cell_slopes(const orc_params* p, const double* Q, int i, int j, int k, double dq[3][5])
{
  const geom_t g = geom(p);
  const double st = p->slope_type;
  for (int d = 0; d < 3; ++d)
    for (int v = 0; v < 5; ++v)
      dq[d][v] = 0.0;
  if (st == 0)
    return;
  if (st == 1 || st == 2) {
    for (int v = 0; v < p->nbvar; ++v) {
      double q = AT(Q, i, j, k, v);
      dq[0][v] = slope_scalar(st, q, AT(Q, i + 1, j, k, v), AT(Q, i - 1, j, k, v));
      dq[1][v] = slope_scalar(st, q, AT(Q, i, j + 1, k, v), AT(Q, i, j - 1, k, v));
      if (p->dim == 3)
        dq[2][v] = slope_scalar(st, q, AT(Q, i, j, k + 1, v), AT(Q, i, j, k - 1, v));
    }
  }
  /* any other slope_type: slopes are left as initialised by the caller (the reference leaves them
   * uninitialised; not exercised by any shipped input) */
}

/* muscl/HydroBaseFunctor3D.h:274-389 trace_unsplit_3d_along_dir, HydroBaseFunctor2D.h:216-286.
 * face: 0 XMIN, 1 XMAX, 2 YMIN, 3 YMAX, 4 ZMIN, 5 ZMAX */
static void trace_along_dir(const orc_params* p, const double* q, double dq[3][5], double dtdx,
                            double dtdy, double dtdz, int face, double* qface)
{
  const double gamma0 = p->gamma0, smallr = p->smallr;
  const double r = q[ID], pp = q[IP], u = q[IU], v = q[IV];
  const double drx = dq[0][ID], dpx = dq[0][IP], dux = dq[0][IU], dvx = dq[0][IV];
  const double dry = dq[1][ID], dpy = dq[1][IP], duy = dq[1][IU], dvy = dq[1][IV];
  double sr0, sp0, su0, sv0, sw0 = 0.0;
  double w = 0.0, dwx = 0.0, dwy = 0.0, drz = 0.0, dpz = 0.0, duz = 0.0, dvz = 0.0, dwz = 0.0;
  if (p->dim == 3) {
    w = q[IW];
    dwx = dq[0][IW];
    dwy = dq[1][IW];
    drz = dq[2][ID];
    dpz = dq[2][IP];
    duz = dq[2][IU];
    dvz = dq[2][IV];
    dwz = dq[2][IW];
    sr0 = -u * drx - v * dry - w * drz - (dux + dvy + dwz) * r;
    sp0 = -u * dpx - v * dpy - w * dpz - (dux + dvy + dwz) * gamma0 * pp;
    su0 = -u * dux - v * duy - w * duz - (dpx) / r;
    sv0 = -u * dvx - v * dvy - w * dvz - (dpy) / r;
    sw0 = -u * dwx - v * dwy - w * dwz - (dpz) / r;
  } else {
    sr0 = -u * drx - v * dry - (dux + dvy) * r;
    sp0 = -u * dpx - v * dpy - (dux + dvy) * gamma0 * pp;
    su0 = -u * dux - v * duy - (dpx) / r;
    sv0 = -u * dvx - v * dvy - (dpy) / r;
  }
  const int d = face / 2;
  const double sgn = (face % 2) ? 0.5 : -0.5; /* XMAX: +HALF*dq, XMIN: -HALF*dq */
  const double dt_d = (d == 0) ? dtdx : (d == 1) ? dtdy : dtdz;
  /* the reference writes  q -/+ HALF_F*dq + s0*dtdx*HALF_F ;  (q - 0.5*dq) == (q + (-0.5)*dq) */
  qface[ID] = r + sgn * dq[d][ID] + sr0 * dt_d * 0.5;
  qface[IP] = pp + sgn * dq[d][IP] + sp0 * dt_d * 0.5;
  qface[IU] = u + sgn * dq[d][IU] + su0 * dt_d * 0.5;
  qface[IV] = v + sgn * dq[d][IV] + sv0 * dt_d * 0.5;
  if (p->dim == 3)
    qface[IW] = w + sgn * dq[d][IW] + sw0 * dt_d * 0.5;
  qface[ID] = fmax(smallr, qface[ID]);
}

/* ------------------------------------------------------------------------------------------- */
/* shared/RiemannSolvers.h:21-48 */
static void cmpflx(const orc_params* p, const double* qgdnv, double* flux)
{
  const double gamma0 = p->gamma0;
  flux[ID] = qgdnv[ID] * qgdnv[IU];
  flux[IU] = flux[ID] * qgdnv[IU] + qgdnv[IP];
  flux[IV] = flux[ID] * qgdnv[IV];
  if (p->dim == 3)
    flux[IW] = flux[ID] * qgdnv[IW];
  double entho = 1.0 / (gamma0 - 1.0);
  double ekin = 0.5 * qgdnv[ID] * (qgdnv[IU] * qgdnv[IU] + qgdnv[IV] * qgdnv[IV]);
  if (p->dim == 3)
    ekin += 0.5 * qgdnv[ID] * (qgdnv[IW] * qgdnv[IW]);
  double etot = qgdnv[IP] * entho + ekin;
  flux[IP] = qgdnv[IU] * (etot + qgdnv[IP]);
}

/* shared/RiemannSolvers.h:59-199 */
static void riemann_approx(const orc_params* p, const double* qleft, const double* qright, double* flux)
{
  const double gamma0 = p->gamma0, gamma6 = p->gamma6, smallr = p->smallr, smallc = p->smallc,
               smallp = p->smallp, smallpp = p->smallpp;
  double qgdnv[5];
  double rl = fmax(qleft[ID], smallr);
  double ul = qleft[IU];
  double pl = fmax(qleft[IP], rl * smallp);
  double rr = fmax(qright[ID], smallr);
  double ur = qright[IU];
  double pr = fmax(qright[IP], rr * smallp);
  double cl = gamma0 * pl * rl;
  double cr = gamma0 * pr * rr;
  double wl = sqrt(cl);
  double wr = sqrt(cr);
  double pstar = fmax(((wr * pl + wl * pr) + wl * wr * (ul - ur)) / (wl + wr), 0.0);
  double pold = pstar;
  double conv = 1.0;
  for (int iter = 0; (iter < 10) && (conv > 1e-6); ++iter) {
    double wwl = sqrt(cl * (1.0 + gamma6 * (pold - pl) / pl));
    double wwr = sqrt(cr * (1.0 + gamma6 * (pold - pr) / pr));
    double ql = 2.0 * wwl * wwl * wwl / (wwl * wwl + cl);
    double qr = 2.0 * wwr * wwr * wwr / (wwr * wwr + cr);
    double usl = ul - (pold - pl) / wwl;
    double usr = ur + (pold - pr) / wwr;
    double delp = fmax(qr * ql / (qr + ql) * (usl - usr), -pold);
    pold = pold + delp;
    conv = fabs(delp / (pold + smallpp));
  }
  pstar = pold;
  wl = sqrt(cl * (1.0 + gamma6 * (pstar - pl) / pl));
  wr = sqrt(cr * (1.0 + gamma6 * (pstar - pr) / pr));
  double ustar = 0.5 * (ul + (pl - pstar) / wl + ur - (pr - pstar) / wr);
  double sgnm = copysign(1.0, ustar);
  double ro, uo, po, wo;
  if (sgnm > 0.0) {
    ro = rl; uo = ul; po = pl; wo = wl;
  } else {
    ro = rr; uo = ur; po = pr; wo = wr;
  }
  double co = fmax(smallc, sqrt(fabs(gamma0 * po / ro)));
  double rstar = fmax(ro / (1.0 + ro * (po - pstar) / (wo * wo)), smallr);
  double cstar = fmax(smallc, sqrt(fabs(gamma0 * pstar / rstar)));
  double spout = co - sgnm * uo;
  double spin = cstar - sgnm * ustar;
  double ushock = wo / ro - sgnm * uo;
  if (pstar >= po) {
    spin = ushock;
    spout = ushock;
  }
  double scr = fmax(spout - spin, smallc + fabs(spout + spin));
  double frac = 0.5 * (1.0 + (spout + spin) / scr);
  if (frac != frac)
    frac = 0.0;
  else
    frac = frac >= 1.0 ? 1.0 : frac <= 0.0 ? 0.0 : frac;
  qgdnv[ID] = frac * rstar + (1.0 - frac) * ro;
  qgdnv[IU] = frac * ustar + (1.0 - frac) * uo;
  qgdnv[IP] = frac * pstar + (1.0 - frac) * po;
  if (spout < 0.0) {
    qgdnv[ID] = ro; qgdnv[IU] = uo; qgdnv[IP] = po;
  }
  if (spin > 0.0) {
    qgdnv[ID] = rstar; qgdnv[IU] = ustar; qgdnv[IP] = pstar;
  }
  qgdnv[IW] = 0.0;
  if (sgnm > 0.0) {
    qgdnv[IV] = qleft[IV];
    if (p->dim == 3) qgdnv[IW] = qleft[IW];
  } else {
    qgdnv[IV] = qright[IV];
    if (p->dim == 3) qgdnv[IW] = qright[IW];
  }
  cmpflx(p, qgdnv, flux);
}

/* left/right conservative states and physical fluxes shared by LLF and HLL,
 * shared/RiemannSolvers.h:257-313 (LLF) and :382-422 (HLL; fleft[ID] = uleft[IU] there) */
static void cons_and_flux(const orc_params* p, const double* q, double entho, int hll, double* u, double* f)
{
  u[ID] = q[ID];
  u[IP] = q[IP] * entho + 0.5 * q[ID] * q[IU] * q[IU];
  u[IP] += 0.5 * q[ID] * q[IV] * q[IV];
  if (p->dim == 3)
    u[IP] += 0.5 * q[ID] * q[IW] * q[IW];
  u[IU] = q[ID] * q[IU];
  u[IV] = q[ID] * q[IV];
  if (p->dim == 3)
    u[IW] = q[ID] * q[IW];
  f[ID] = hll ? u[IU] : u[ID] * q[IU];
  f[IP] = q[IU] * (u[IP] + q[IP]);
  f[IU] = q[IP] + u[IU] * q[IU];
  f[IV] = f[ID] * q[IV];
  if (p->dim == 3)
    f[IW] = f[ID] * q[IW];
}

/* shared/RiemannSolvers.h:217-327 */
static void riemann_llf(const orc_params* p, const double* qleft, const double* qright, double* flux)
{
  const double gamma0 = p->gamma0, smallr = p->smallr, smallp = p->smallp;
  const double entho = 1.0 / (gamma0 - 1.0);
  double rl = fmax(qleft[ID], smallr), ul = qleft[IU], pl = fmax(qleft[IP], rl * smallp);
  double rr = fmax(qright[ID], smallr), ur = qright[IU], pr = fmax(qright[IP], rr * smallp);
  double cl = sqrt(gamma0 * pl / rl);
  double cr = sqrt(gamma0 * pr / rr);
  double cmax = fmax(fabs(ul) + cl, fabs(ur) + cr);
  double uL[5], uR[5], fL[5], fR[5];
  cons_and_flux(p, qleft, entho, 0, uL, fL);
  cons_and_flux(p, qright, entho, 0, uR, fR);
  for (int v = 0; v < p->nbvar; ++v)
    flux[v] = 0.5 * (fL[v] + fR[v] - cmax * (uR[v] - uL[v]));
}

/* shared/RiemannSolvers.h:343-435 (only reachable from the 2-D dispatcher, :587-591) */
static void riemann_hll(const orc_params* p, const double* qleft, const double* qright, double* flux)
{
  const double gamma0 = p->gamma0, smallr = p->smallr, smallp = p->smallp;
  const double entho = 1.0 / (gamma0 - 1.0);
  double rl = fmax(qleft[ID], smallr), ul = qleft[IU], pl = fmax(qleft[IP], rl * smallp);
  double rr = fmax(qright[ID], smallr), ur = qright[IU], pr = fmax(qright[IP], rr * smallp);
  double cl = sqrt(gamma0 * pl / rl);
  double cr = sqrt(gamma0 * pr / rr);
  double SL = fmin(fmin(ul, ur) - fmax(cl, cr), 0.0);
  double SR = fmax(fmax(ul, ur) + fmax(cl, cr), 0.0);
  double uL[5], uR[5], fL[5], fR[5];
  cons_and_flux(p, qleft, entho, 1, uL, fL);
  cons_and_flux(p, qright, entho, 1, uR, fR);
  for (int v = 0; v < p->nbvar; ++v)
    flux[v] = (SR * fL[v] - SL * fR[v] + SR * SL * (uR[v] - uL[v])) / (SR - SL);
}

/* shared/RiemannSolvers.h:445-568 */
static void riemann_hllc(const orc_params* p, const double* qleft, const double* qright, double* flux)
{
  const double gamma0 = p->gamma0, smallr = p->smallr, smallp = p->smallp, smallc = p->smallc;
  const double entho = 1.0 / (gamma0 - 1.0);
  double rl = fmax(qleft[ID], smallr);
  double pl = fmax(qleft[IP], rl * smallp);
  double ul = qleft[IU];
  double ecinl = 0.5 * rl * ul * ul;
  ecinl += 0.5 * rl * qleft[IV] * qleft[IV];
  if (p->dim == 3)
    ecinl += 0.5 * rl * qleft[IW] * qleft[IW];
  double etotl = pl * entho + ecinl;
  double ptotl = pl;
  double rr = fmax(qright[ID], smallr);
  double pr = fmax(qright[IP], rr * smallp);
  double ur = qright[IU];
  double ecinr = 0.5 * rr * ur * ur;
  ecinr += 0.5 * rr * qright[IV] * qright[IV];
  if (p->dim == 3)
    ecinr += 0.5 * rr * qright[IW] * qright[IW];
  double etotr = pr * entho + ecinr;
  double ptotr = pr;
  double cfastl = sqrt(fmax(gamma0 * pl / rl, smallc * smallc));
  double cfastr = sqrt(fmax(gamma0 * pr / rr, smallc * smallc));
  double SL = fmin(ul, ur) - fmax(cfastl, cfastr);
  double SR = fmax(ul, ur) + fmax(cfastl, cfastr);
  double rcl = rl * (ul - SL);
  double rcr = rr * (SR - ur);
  double ustar = (rcr * ur + rcl * ul + (ptotl - ptotr)) / (rcr + rcl);
  double ptotstar = (rcr * ptotl + rcl * ptotr + rcl * rcr * (ul - ur)) / (rcr + rcl);
  double rstarl = rl * (SL - ul) / (SL - ustar);
  double etotstarl = ((SL - ul) * etotl - ptotl * ul + ptotstar * ustar) / (SL - ustar);
  double rstarr = rr * (SR - ur) / (SR - ustar);
  double etotstarr = ((SR - ur) * etotr - ptotr * ur + ptotstar * ustar) / (SR - ustar);
  double ro, uo, ptoto, etoto;
  if (SL > 0.0) {
    ro = rl; uo = ul; ptoto = ptotl; etoto = etotl;
  } else if (ustar > 0.0) {
    ro = rstarl; uo = ustar; ptoto = ptotstar; etoto = etotstarl;
  } else if (SR > 0.0) {
    ro = rstarr; uo = ustar; ptoto = ptotstar; etoto = etotstarr;
  } else {
    ro = rr; uo = ur; ptoto = ptotr; etoto = etotr;
  }
  flux[ID] = ro * uo;
  flux[IU] = ro * uo * uo + ptoto;
  flux[IP] = (etoto + ptoto) * uo;
  if (flux[ID] > 0.0)
    flux[IV] = flux[ID] * qleft[IV];
  else
    flux[IV] = flux[ID] * qright[IV];
  if (p->dim == 3) {
    if (flux[ID] > 0.0)
      flux[IW] = flux[ID] * qleft[IW];
    else
      flux[IW] = flux[ID] * qright[IW];
  }
}

/* shared/RiemannSolvers.h:573-633.  The 3-D dispatcher has no HLL branch (flux left
 * uninitialised in the reference); we return zeros there and the product rejects the combination. */
static void riemann_hydro(const orc_params* p, const double* ql, const double* qr, double* flux)
{
  for (int v = 0; v < 5; ++v)
    flux[v] = 0.0;
  if (p->riemann == RIEMANN_APPROX)
    riemann_approx(p, ql, qr, flux);
  else if (p->riemann == RIEMANN_HLL && p->dim == 2)
    riemann_hll(p, ql, qr, flux);
  else if (p->riemann == RIEMANN_HLLC)
    riemann_hllc(p, ql, qr, flux);
  else if (p->riemann == RIEMANN_LLF)
    riemann_llf(p, ql, qr, flux);
}

void orc_riemann_hydro(const orc_params* p, const double* ql, const double* qr, double* flux)
{
  riemann_hydro(p, ql, qr, flux);
}

/* ------------------------------------------------------------------------------------------- */
static void swapv(double* a, double* b)
{
  double t = *a;
  *a = *b;
  *b = t;
}

/* flux (times dt/dx) through the lower face of cell (i,j,k) along dir d:
 * muscl/HydroRunFunctors3D.h:327-707 (ComputeAndStoreFluxesFunctor3D), 2D: HydroRunFunctors2D.h:538-760 */
static void face_flux(const orc_params* p, const double* Q, const double* grav, int i, int j, int k,
                      int d, double dt, double* out)
{
  const geom_t g = geom(p);
  const double dtdx = dt / p->dx, dtdy = dt / p->dy, dtdz = dt / p->dz;
  const int in = i - (d == 0), jn = j - (d == 1), kn = k - (d == 2);
  double qLoc[5] = { 0 }, qNb[5] = { 0 }, dq[3][5], dqN[3][5], qleft[5] = { 0 }, qright[5] = { 0 }, flux[5];
  for (int v = 0; v < p->nbvar; ++v) {
    qLoc[v] = AT(Q, i, j, k, v);
    qNb[v] = AT(Q, in, jn, kn, v);
  }
  cell_slopes(p, Q, i, j, k, dq);
  cell_slopes(p, Q, in, jn, kn, dqN);
  trace_along_dir(p, qLoc, dq, dtdx, dtdy, dtdz, 2 * d, qright);    /* FACE_?MIN of the cell */
  trace_along_dir(p, qNb, dqN, dtdx, dtdy, dtdz, 2 * d + 1, qleft); /* FACE_?MAX of the neighbour */
  if (p->gravity_enabled && p->gravity_hancock) {
    const int nc = p->dim;
    for (int c = 0; c < nc; ++c)
      qleft[IU + c] += 0.5 * dt * AT(grav, in, jn, kn, c);
    /* implementationVersion 1 omits the predictor on qright along Z
     * (HydroRunFunctors3D.h:1351 vs :1277-1285) */
    if (!(p->impl_version == 1 && p->dim == 3 && d == 2))
      for (int c = 0; c < nc; ++c)
        qright[IU + c] += 0.5 * dt * AT(grav, i, j, k, c);
  }
  if (d == 1) {
    swapv(&qleft[IU], &qleft[IV]);
    swapv(&qright[IU], &qright[IV]);
  } else if (d == 2) {
    swapv(&qleft[IU], &qleft[IW]);
    swapv(&qright[IU], &qright[IW]);
  }
  riemann_hydro(p, qleft, qright, flux);
  if (d == 1)
    swapv(&flux[IU], &flux[IV]);
  else if (d == 2)
    swapv(&flux[IU], &flux[IW]);
  const double s = (d == 0) ? dtdx : (d == 1) ? dtdy : dtdz;
  for (int v = 0; v < p->nbvar; ++v)
    out[v] = flux[v] * s;
}

/* muscl/SolverHydroMuscl.cpp:216-297 (2D) and :306-390 (3D) */
void orc_hydro_step(const orc_params* p, double* Uin, double* Uout, double dt, const double* grav)
{
  const geom_t g = geom(p);
  const int gw = p->ghost, nbvar = p->nbvar, dim = p->dim;
  const size_t n = g.sv * nbvar;
  orc_make_boundaries(p, Uin);
  memcpy(Uout, Uin, n * sizeof(double));
  double* Q = (double*)malloc(n * sizeof(double));
  double* F[3] = { 0, 0, 0 };
  for (int d = 0; d < dim; ++d)
    F[d] = (double*)calloc(n, sizeof(double));
  orc_hydro_prims(p, Uin, Q);
  const int k0 = (dim == 3) ? gw : 0, k1 = (dim == 3) ? g.ksize - gw : 0; /* inclusive upper bounds */
  for (int k = k0; k <= k1; ++k)
    for (int j = gw; j <= g.jsize - gw; ++j)
      for (int i = gw; i <= g.isize - gw; ++i) {
        /* the reference evaluates all faces of cell (i,j,k) inside one guard that allows
         * i == isize-gw etc. (reads up to index isize-gw+1 = last ghost: in bounds) */
        for (int d = 0; d < dim; ++d) {
          double f[5];
          face_flux(p, Q, grav, i, j, k, d, dt, f);
          for (int v = 0; v < nbvar; ++v)
            AT(F[d], i, j, k, v) = f[v];
        }
      }
  /* UpdateFunctor{2D,3D}: HydroRunFunctors3D.h:762-813 */
  const int k1i = (dim == 3) ? g.ksize - gw : 1;
  for (int k = k0; k < k1i; ++k)
    for (int j = gw; j < g.jsize - gw; ++j)
      for (int i = gw; i < g.isize - gw; ++i)
        for (int v = 0; v < nbvar; ++v) {
          double u = AT(Uout, i, j, k, v);
          u += AT(F[0], i, j, k, v);
          u -= AT(F[0], i + 1, j, k, v);
          u += AT(F[1], i, j, k, v);
          u -= AT(F[1], i, j + 1, k, v);
          if (dim == 3) {
            u += AT(F[2], i, j, k, v);
            u -= AT(F[2], i, j, k + 1, v);
          }
          AT(Uout, i, j, k, v) = u;
        }
  /* GravitySourceTermFunctor: HydroRunFunctors3D.h:1939-1975, HydroRunFunctors2D.h:1594-1625 */
  if (p->gravity_enabled) {
    for (int k = k0; k < k1i; ++k)
      for (int j = gw; j < g.jsize - gw; ++j)
        for (int i = gw; i < g.isize - gw; ++i) {
          double rhoOld = AT(Uin, i, j, k, ID);
          double rhoNew = AT(Uout, i, j, k, ID);
          if (dim == 2)
            rhoNew = fmax(p->smallr, rhoNew);
          double m[3] = { AT(Uout, i, j, k, IU), AT(Uout, i, j, k, IV), 0.0 };
          if (dim == 3)
            m[2] = AT(Uout, i, j, k, IW);
          double ekin_old = (dim == 3) ? 0.5 * (m[0] * m[0] + m[1] * m[1] + m[2] * m[2]) / rhoNew
                                       : 0.5 * (m[0] * m[0] + m[1] * m[1]) / rhoNew;
          for (int c = 0; c < dim; ++c)
            m[c] += 0.5 * dt * AT(grav, i, j, k, c) * (rhoOld + rhoNew);
          AT(Uout, i, j, k, IU) = m[0];
          AT(Uout, i, j, k, IV) = m[1];
          if (dim == 3)
            AT(Uout, i, j, k, IW) = m[2];
          double ekin_new = (dim == 3) ? 0.5 * (m[0] * m[0] + m[1] * m[1] + m[2] * m[2]) / rhoNew
                                       : 0.5 * (m[0] * m[0] + m[1] * m[1]) / rhoNew;
          AT(Uout, i, j, k, IP) += (ekin_new - ekin_old);
        }
  }
  free(Q);
  for (int d = 0; d < dim; ++d)
    free(F[d]);
}
