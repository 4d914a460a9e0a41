/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C, single thread) of the reference's
 * MUSCL-Hancock time step.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; the product library never does.
 *
 * Parity pin: the reference has no golden vectors; this restatement is pinned
 * against outputs of the reference binary itself (oracle/_ref/euler_kokkos_ref, built from
 * /root/reference by oracle/ref_build/Makefile) -- see tests/test_oracle_vs_reference.py and the
 * fixtures under tests/golden/ (made by tests/golden/make_golden.py).
 *
 * Layout: U[var][k][j][i], ghosts included, i fastest.
 */
#ifndef MUSCL_ORACLE_H_
#define MUSCL_ORACLE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_params {
  int32_t dim, mhd, nbvar, ghost, nx, ny, nz, riemann, impl_version;
  int32_t bc[6];
  int32_t gravity_enabled, gravity_hancock, reserved0;
  double dx, dy, dz;
  double gamma0, gamma6, smallr, smallc, smallp, smallpp, cfl, slope_type;
  double reserved1[4];
} orc_params;

/* ghost fill, serial order XMIN..ZMAX (SolverBase.cpp:502-512, BoundariesFunctors.h) */
void orc_make_boundaries(const orc_params* p, double* U);
/* max over interior of the inverse time step (ComputeDt*Functor*); hydro reads U, MHD reads Q */
double orc_hydro_invdt(const orc_params* p, const double* U, const double* grav);
/* U -> Q on the whole box (ConvertToPrimitivesFunctor{2D,3D}) */
void orc_hydro_prims(const orc_params* p, const double* U, double* Q);
/* one godunov_unsplit_impl (hydro): fills ghosts of Uin, writes Uout (all cells) */
void orc_hydro_step(const orc_params* p, double* Uin, double* Uout, double dt, const double* grav);

/* MHD (oracle/mhd_oracle.c) */
void orc_mhd_prims(const orc_params* p, const double* U, double* Q);
double orc_mhd_invdt(const orc_params* p, const double* Q);
/* one godunov_unsplit_impl (MHD, impl 0) with a given dt; dt<0 => compute it from Q as the
 * reference does (cfl/invDt, clamped to t_end - t) and return it through dt_used */
void orc_mhd_step(const orc_params* p, double* Uin, double* Uout, double dt, double t, double t_end,
                  double* dt_used);

/* single Riemann problems for known-answer tests: q = (rho, p, u, v, w[, A, B, C]) */
void orc_riemann_hydro(const orc_params* p, const double* ql, const double* qr, double* flux);
void orc_riemann_mhd(const orc_params* p, const double* ql, const double* qr, double* flux);

#ifdef __cplusplus
}
#endif
#endif
