// TEST INFRASTRUCTURE ONLY: the per-cell / per-face functions of the fused CUDA hydro step
// (euler_kokkos_b200/csrc/hydro_physics.cuh -- the very functions hydro_step_kernel calls) compiled for the HOST
// and driven by plain loops in the kernel's own order of operations (conversion -> limited slopes -> Hancock trace
// with FMA-formed face states -> the lower x / y / z Riemann problems of a cell solved together -> running update in
// the kernel's summation order -> gravity source term -> CFL contribution), so that the arithmetic the kernel
// performs can be checked against the oracle on a machine without a GPU (pytest -m "not gpu").  What this does NOT
// cover is the kernel's tiling, shared-memory ring and barriers: that is the GPU tier's job.
// Never loaded by the product library.
#include <cmath>
#include <vector>

#include "../../euler_kokkos_b200/csrc/ek_grid.h"
#include "../../euler_kokkos_b200/csrc/hydro_physics.cuh"

using namespace ek;

namespace {

template <int DIM, int RS, bool GRAV>
void step(const Grid& g, const double* Uin, double* Uout, const double* grav, double dt)
{
  constexpr int NV = HydroTraits<DIM>::NV;
  const long long n = g.sv, sj = g.sj, sk = g.sk;
  const double dtd[3] = { dt * g.invdx, dt * g.invdy, (DIM == 3) ? dt * g.invdz : 0.0 };
  const int gw = g.gw;
  const int k0 = (DIM == 3) ? 0 : 0, k1 = (DIM == 3) ? g.ksize : 1;
  auto at = [&](int i, int j, int k) { return (long long)i + sj * j + sk * k; };

  std::vector<double> Q((size_t)NV * n), LO((size_t)15 * n, 1.0), HI((size_t)15 * n, 1.0), F((size_t)15 * n, 0.0);
  // cons -> prim on every cell (prims_to_ring)
  for (long long c = 0; c < n; ++c) {
    double u[NV], q[5], rd;
    for (int v = 0; v < NV; ++v)
      u[v] = Uin[c + n * v];
    hydro_cons_to_prim<DIM>(g, u, q, rd);
    for (int v = 0; v < NV; ++v)
      Q[(size_t)v * n + c] = q[v];
  }
  // slopes + Hancock trace on the interior plus a one-cell ring (do_trace)
  const double st = g.slope_type;
  for (int k = (DIM == 3 ? 1 : 0); k < (DIM == 3 ? g.ksize - 1 : 1); ++k)
    for (int j = 1; j <= g.ny + gw; ++j)
      for (int i = 1; i <= g.nx + gw; ++i) {
        const long long c = at(i, j, k);
        double qc[NV], dq[3][5];
        for (int v = 0; v < NV; ++v) {
          const double* q = &Q[(size_t)v * n];
          qc[v] = q[c];
          dq[0][v] = limited_slope(st, g.slope_k, q[c - 1], q[c], q[c + 1]);
          dq[1][v] = limited_slope(st, g.slope_k, q[c - sj], q[c], q[c + sj]);
          if constexpr (DIM == 3)
            dq[2][v] = limited_slope(st, g.slope_k, q[c - sk], q[c], q[c + sk]);
        }
        if (st == 0.0)
          for (int d = 0; d < DIM; ++d)
            for (int v = 0; v < NV; ++v)
              dq[d][v] = 0.0;
        double s[NV];
        hydro_sources<DIM>(g, qc, dq, s);
        double gh[3] = { 0.0, 0.0, 0.0 };
        if constexpr (GRAV) {
          if (g.gravity_hancock)
            for (int d = 0; d < DIM; ++d)
              gh[d] = 0.5 * dt * grav[c + n * d];
        }
        const double dth[3] = { dtd[0] * 0.5, dtd[1] * 0.5, dtd[2] * 0.5 };
        for (int v = 0; v < NV; ++v) {
          const double add = (GRAV && v >= IU) ? gh[v - IU] : 0.0;
          double bx = fma(s[v], dth[0], qc[v]), by = fma(s[v], dth[1], qc[v]);
          if constexpr (GRAV) {
            bx += add;
            by += add;
          }
          double lox = fma(-0.5, dq[0][v], bx), loy = fma(-0.5, dq[1][v], by);
          double mxv = fma(0.5, dq[0][v], bx), myv = fma(0.5, dq[1][v], by);
          if (v == ID) {
            lox = dmax(g.smallr, lox);
            loy = dmax(g.smallr, loy);
            mxv = dmax(g.smallr, mxv);
            myv = dmax(g.smallr, myv);
          }
          LO[(size_t)(0 * 5 + v) * n + c] = lox;
          LO[(size_t)(1 * 5 + v) * n + c] = loy;
          HI[(size_t)(0 * 5 + v) * n + c] = mxv;
          HI[(size_t)(1 * 5 + v) * n + c] = myv;
          if constexpr (DIM == 3) {
            const double bz0 = fma(s[v], dth[2], qc[v]);
            double bzl = bz0, bzh = bz0;
            if constexpr (GRAV) {
              bzl += (g.impl_version == 1) ? 0.0 : add;
              bzh += add;
            }
            double loz = fma(-0.5, dq[2][v], bzl), mzv = fma(0.5, dq[2][v], bzh);
            if (v == ID) {
              loz = dmax(g.smallr, loz);
              mzv = dmax(g.smallr, mzv);
            }
            LO[(size_t)(2 * 5 + v) * n + c] = loz;
            HI[(size_t)(2 * 5 + v) * n + c] = mzv;
          }
        }
      }
  // the lower x / y / z faces of a cell, solved together (riemann_hydro_n), flux * dt/dx stored per face
  for (int k = (DIM == 3 ? gw : 0); k <= (DIM == 3 ? g.nz + gw : 0); ++k)
    for (int j = gw; j <= g.ny + gw; ++j)
      for (int i = gw; i <= g.nx + gw; ++i) {
        const long long c = at(i, j, k);
        const long long nb[3] = { c - 1, c - sj, c - sk };
        double ql[3][5], qr[3][5], f[3][5];
        for (int d = 0; d < 3; ++d)
          for (int v = 0; v < 5; ++v)
            ql[d][v] = qr[d][v] = 1.0;
        for (int d = 0; d < DIM; ++d)
          for (int v = 0; v < NV; ++v) {
            ql[d][v] = HI[(size_t)(d * 5 + v) * n + nb[d]];
            qr[d][v] = LO[(size_t)(d * 5 + v) * n + c];
          }
        std::swap(ql[1][IU], ql[1][IV]);
        std::swap(qr[1][IU], qr[1][IV]);
        if constexpr (DIM == 3) {
          std::swap(ql[2][IU], ql[2][IW]);
          std::swap(qr[2][IU], qr[2][IW]);
        }
        riemann_hydro_n<DIM, RS, DIM>(g, ql, qr, f);
        std::swap(f[1][IU], f[1][IV]);
        if constexpr (DIM == 3)
          std::swap(f[2][IU], f[2][IW]);
        for (int d = 0; d < DIM; ++d)
          for (int v = 0; v < NV; ++v)
            F[(size_t)(d * 5 + v) * n + c] = f[d][v] * dtd[d];
      }
  // running update in the kernel's order, gravity source term
  for (int k = (DIM == 3 ? gw : 0); k < (DIM == 3 ? g.nz + gw : 1); ++k)
    for (int j = gw; j < g.ny + gw; ++j)
      for (int i = gw; i < g.nx + gw; ++i) {
        const long long c = at(i, j, k);
        double un[NV];
        for (int v = 0; v < NV; ++v) {
          double u = Uin[c + n * v];
          u += F[(size_t)(0 * 5 + v) * n + c];
          u -= F[(size_t)(0 * 5 + v) * n + c + 1];
          u += F[(size_t)(1 * 5 + v) * n + c];
          u -= F[(size_t)(1 * 5 + v) * n + c + sj];
          if constexpr (DIM == 3) {
            u += F[(size_t)(2 * 5 + v) * n + c];
            u = u - F[(size_t)(2 * 5 + v) * n + c + sk];
          }
          un[v] = u;
        }
        if constexpr (GRAV) {
          const double rhoOld = Uin[c];
          const double rhoNew = (DIM == 3) ? un[ID] : dmax(g.smallr, un[ID]);
          const double ir = 1.0 / rhoNew;
          double ek0 = un[IU] * un[IU] + un[IV] * un[IV];
          if constexpr (DIM == 3)
            ek0 += un[IW] * un[IW];
          ek0 = 0.5 * ek0 * ir;
          for (int d = 0; d < DIM; ++d)
            un[IU + d] += 0.5 * dt * grav[c + n * d] * (rhoOld + rhoNew);
          double ek1 = un[IU] * un[IU] + un[IV] * un[IV];
          if constexpr (DIM == 3)
            ek1 += un[IW] * un[IW];
          ek1 = 0.5 * ek1 * ir;
          un[IP] += (ek1 - ek0);
        }
        for (int v = 0; v < NV; ++v)
          Uout[c + n * v] = un[v];
      }
  (void)k0;
  (void)k1;
}

template <int DIM, bool GRAV>
double invdt(const Grid& g, const double* U, const double* grav)
{
  constexpr int NV = HydroTraits<DIM>::NV;
  const long long n = g.sv;
  double m = 0.0;
  for (int k = (DIM == 3 ? g.gw : 0); k < (DIM == 3 ? g.nz + g.gw : 1); ++k)
    for (int j = g.gw; j < g.ny + g.gw; ++j)
      for (int i = g.gw; i < g.nx + g.gw; ++i) {
        const long long c = (long long)i + g.sj * j + g.sk * k;
        double u[NV], gv[3] = { 0.0, 0.0, 0.0 };
        for (int v = 0; v < NV; ++v)
          u[v] = U[c + n * v];
        if constexpr (GRAV)
          for (int d = 0; d < DIM; ++d)
            gv[d] = grav[c + n * d];
        m = dmax(m, hydro_cell_invdt<DIM, GRAV>(g, u, gv));
      }
  return m;
}

template <int DIM, int RS>
int step_grav(const Grid& g, const ek_params* p, const double* Uin, double* Uout, const double* grav, double dt)
{
  if (p->gravity_enabled)
    step<DIM, RS, true>(g, Uin, Uout, grav, dt);
  else
    step<DIM, RS, false>(g, Uin, Uout, grav, dt);
  return 0;
}

} // namespace

extern "C" {

// one step with a given dt on a ghost-filled Uin; Uout must be a copy of Uin (ghosts are not written)
int emul_hydro_step(const ek_params* p, const double* Uin, double* Uout, const double* grav, double dt)
{
  const Grid g = make_grid(*p);
  if (p->dim == 3) {
    switch (p->riemann) {
      case RIEMANN_HLLC: return step_grav<3, RIEMANN_HLLC>(g, p, Uin, Uout, grav, dt);
      case RIEMANN_APPROX: return step_grav<3, RIEMANN_APPROX>(g, p, Uin, Uout, grav, dt);
      case RIEMANN_LLF: return step_grav<3, RIEMANN_LLF>(g, p, Uin, Uout, grav, dt);
    }
    return 1;
  }
  switch (p->riemann) {
    case RIEMANN_HLLC: return step_grav<2, RIEMANN_HLLC>(g, p, Uin, Uout, grav, dt);
    case RIEMANN_APPROX: return step_grav<2, RIEMANN_APPROX>(g, p, Uin, Uout, grav, dt);
    case RIEMANN_LLF: return step_grav<2, RIEMANN_LLF>(g, p, Uin, Uout, grav, dt);
    case RIEMANN_HLL: return step_grav<2, RIEMANN_HLL>(g, p, Uin, Uout, grav, dt);
  }
  return 1;
}

// max inverse dt over the interior (hydro_cell_invdt, the reduction fused at the end of hydro_step_kernel)
double emul_hydro_invdt(const ek_params* p, const double* U, const double* grav)
{
  const Grid g = make_grid(*p);
  if (p->dim == 3)
    return p->gravity_enabled ? invdt<3, true>(g, U, grav) : invdt<3, false>(g, U, grav);
  return p->gravity_enabled ? invdt<2, true>(g, U, grav) : invdt<2, false>(g, U, grav);
}
}
