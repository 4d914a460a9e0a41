// 3-D fused Euler MUSCL-Hancock step, ONE block barrier per plane ("1b" schedule).
//
// Same physics functions, same tiling and the same floating-point order as hydro_step_kernel<3,...>
// (hydro_kernels.cuh; reference: muscl/SolverHydroMuscl.cpp:306-390, muscl/HydroRunFunctors3D.h:16-820,1903-1985),
// but the three phases of a plane are software-pipelined across planes so that one __syncthreads() per plane is
// enough.  ncu on the two-barrier kernel (profiles/r01_hydro_step_v3_ncu.txt): 1 CTA of 15 warps per SM, barrier
// stalls are the largest stall reason (1.3 of 6.4 warp-cycles per issue) because each of the short phases ends
// with every warp waiting for the slowest one and only 12 of the 15 warps solve Riemann problems.
//
// Iteration p of the z-march does, in program order and without a block barrier in between:
//   B(p)    the three Riemann problems at the lower x / y / z faces of the cell in plane p (states traced in the
//           previous iteration), x-neighbour flux exchanged with a warp shuffle (a tile row is one warp);
//   C(p-1)  plane p-1 is finished: its y-neighbour flux (shared memory, written in the previous iteration) and the
//           z flux just computed are applied, U(n+1) is stored, the CFL maximum of the NEXT step is accumulated;
//           plane p starts its running update U + fx - fx' + fy;
//   A(p+1)  limited slopes + Hancock trace of plane p+1 -> face states (own lower faces stay in registers);
//   P(p+3)  the primitives of plane p+3 replace plane p in the ring.
// Cross-warp data (primitives ring, YMAX-face states, y fluxes) are double buffered by plane parity, so a buffer
// written in iteration p is read in iteration p+1 and not rewritten before iteration p+2.  Cross-lane data of one
// warp (XMAX-face states) need only a __syncwarp().
#pragma once
#include "hydro_kernels.cuh"

namespace ek {

template <int TY>
struct HydroSmem1b {
  static constexpr int NV = 5;
  static constexpr int ROWS = TY + 4;
  static constexpr int PLANE = ROWS * 32; // doubles per variable
  double q[3][NV * PLANE];  // primitives: ring of planes
  double mx[NV * PLANE];    // XMAX-face state of each cell, read by the next lane of the same warp
  double my[2][NV * PLANE]; // YMAX-face state, by plane parity
  double fy[2][NV * PLANE]; // flux*dt/dy through the lower y face, by plane parity
  double mz[2][NV * PLANE]; // ZMAX-face state, by plane parity (own slot only)
  double part[NV * PLANE];  // running update of the plane in flight (own slot only)
  double fz[NV * PLANE];    // z flux through the lower face of the plane in flight (own slot only)
  double red[ROWS];         // per-warp invDt maxima
};

template <int RS, int TY, bool GRAV>
__global__ void __launch_bounds__(32 * (TY + 4), 1) hydro_step1b_kernel(const HydroArgs a)
{
  constexpr int DIM = 3, NV = 5;
  using Smem = HydroSmem1b<TY>;
  constexpr int PLANE = Smem::PLANE;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

  const Grid& g = a.g;
  if (a.ts->finished)
    return;
  const double dt = a.ts->dt;
  const double dtd[3] = { dt * g.invdx, dt * g.invdy, dt * g.invdz };

  const int lane = threadIdx.x, row = threadIdx.y;
  const int xg = g.gw + blockIdx.x * HTX + lane - 2; // ghosted coordinates of this thread's column
  const int yg = g.gw + blockIdx.y * TY + row - 2;
  const int so = row * 32 + lane; // offset inside one smem variable plane

  const bool inb = (xg < g.isize) && (yg < g.jsize);
  const bool do_trace = inb && lane >= 1 && lane <= 30 && row >= 1 && row <= TY + 2 && xg <= g.nx + g.gw &&
                        yg <= g.ny + g.gw;
  const bool in_x = lane >= 2 && lane <= 29 && xg < g.nx + g.gw;
  const bool in_y = row >= 2 && row <= TY + 1 && yg < g.ny + g.gw;
  const bool do_fx = inb && lane >= 2 && lane <= 30 && xg <= g.nx + g.gw && in_y; // lower x faces
  const bool do_fy = inb && in_x && row >= 2 && row <= TY + 2 && yg <= g.ny + g.gw;
  const bool do_cell = in_x && in_y;
  const bool row_flux = row >= 2 && row <= TY + 2; // warp-uniform: this row solves Riemann problems

  const long long col = (long long)xg + g.sj * (long long)yg;
  const double* __restrict__ Uin = a.Uin;

  auto prims_to_ring = [&](const double* u, int slot) {
    double q[NV], rd;
    hydro_cons_to_prim<DIM>(g, u, q, rd);
#pragma unroll
    for (int v = 0; v < NV; ++v)
      sm.q[slot][v * PLANE + so] = q[v];
  };
  auto load_u = [&](int k, double* u) {
#pragma unroll
    for (int v = 0; v < NV; ++v)
      u[v] = inb ? __ldg(Uin + col + g.sk * k + g.sv * v) : 1.0;
  };

  const int kbeg = g.gw + (a.zfirst + (int)blockIdx.z * a.zstep) * a.zchunk;
  const int kend = min(kbeg + a.zchunk, g.nz + g.gw);
  double invdt = 0.0;

  // prologue: primitives of planes kbeg-2, kbeg-1, kbeg into ring slots 0, 1, 2
  {
    double u[NV];
    load_u(kbeg - 2, u);
    prims_to_ring(u, 0);
    load_u(kbeg - 1, u);
    prims_to_ring(u, 1);
    load_u(kbeg, u);
    prims_to_ring(u, 2);
  }
  __syncthreads();

  double lo[3][5]; // states on the XMIN, YMIN, ZMIN faces of the traced plane (right states of its lower faces)
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int v = 0; v < 5; ++v)
      lo[d][v] = 1.0; // halo lanes ride along in the lock-step solve; their fluxes are discarded

  int sp = 0; // ring slot of plane p (planes p+1, p+2 follow cyclically)
  for (int p = kbeg - 2; p <= kend; ++p) {
    const int s1 = (sp == 2) ? 0 : sp + 1, s2 = (s1 == 2) ? 0 : s1 + 1;
    const int zb = p & 1;

    // -- B(p): Riemann problems at the lower x / y / z faces of plane p, solved in lock step
    double fl[3][5], fxr[5];
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      fl[0][v] = 0.0;
      fl[1][v] = 0.0;
      fl[2][v] = 0.0;
      fxr[v] = 0.0;
    }
    if (row_flux && p >= kbeg) {
      double ql[3][5], f[3][5];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        ql[0][v] = sm.mx[v * PLANE + so - 1];
        ql[1][v] = sm.my[zb][v * PLANE + so - 32];
        ql[2][v] = sm.mz[zb ^ 1][v * PLANE + so];
      }
      // rotate into the face-normal frame: normal velocity first (HydroRunFunctors3D.h:597-608,692-703)
      { double t = ql[1][IU]; ql[1][IU] = ql[1][IV]; ql[1][IV] = t; }
      { double t = lo[1][IU]; lo[1][IU] = lo[1][IV]; lo[1][IV] = t; }
      { double t = ql[2][IU]; ql[2][IU] = ql[2][IW]; ql[2][IW] = t; }
      { double t = lo[2][IU]; lo[2][IU] = lo[2][IW]; lo[2][IW] = t; }
      riemann_hydro_n<DIM, RS, DIM>(g, ql, lo, f);
      { double t = f[1][IU]; f[1][IU] = f[1][IV]; f[1][IV] = t; }
      { double t = f[2][IU]; f[2][IU] = f[2][IW]; f[2][IW] = t; }
      if (do_fx) {
#pragma unroll
        for (int v = 0; v < NV; ++v)
          fl[0][v] = f[0][v] * dtd[0];
      }
      if (do_fy) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          fl[1][v] = f[1][v] * dtd[1];
          sm.fy[zb][v * PLANE + so] = fl[1][v];
        }
      }
      if (do_cell) {
#pragma unroll
        for (int v = 0; v < NV; ++v)
          fl[2][v] = f[2][v] * dtd[2];
      }
      // flux through the lower x face of the next cell of the row (= this cell's upper x face)
#pragma unroll
      for (int v = 0; v < NV; ++v)
        fxr[v] = __shfl_down_sync(0xffffffffu, fl[0][v], 1);
    }

    // -- C(p-1): finish plane p-1; start the running update of plane p
    if (do_cell && p >= kbeg) {
      double u0[NV];
      const bool start = p < kend;
      if (start) {
#pragma unroll
        for (int v = 0; v < NV; ++v)
          u0[v] = __ldg(Uin + col + g.sk * p + g.sv * v);
      }
      if (p - 1 >= kbeg) {
        double un[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          double t = sm.part[v * PLANE + so] - sm.fy[zb ^ 1][v * PLANE + so + 32];
          t += sm.fz[v * PLANE + so];
          un[v] = t - fl[2][v];
        }
        const long long o = col + g.sk * (p - 1);
        double gv[3] = { 0.0, 0.0, 0.0 };
        if constexpr (GRAV) {
          // GravitySourceTermFunctor3D (HydroRunFunctors3D.h:1939-1975)
          const double rhoOld = __ldg(Uin + o);
          const double rhoNew = un[ID];
          const double ir = 1.0 / rhoNew;
          const double ek0 = 0.5 * (un[IU] * un[IU] + un[IV] * un[IV] + un[IW] * un[IW]) * ir;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            gv[c] = __ldg(a.grav + o + g.sv * c);
            un[IU + c] += 0.5 * dt * gv[c] * (rhoOld + rhoNew);
          }
          const double ek1 = 0.5 * (un[IU] * un[IU] + un[IV] * un[IV] + un[IW] * un[IW]) * ir;
          un[IP] += (ek1 - ek0);
        }
#pragma unroll
        for (int v = 0; v < NV; ++v)
          a.Uout[o + g.sv * v] = un[v];
        invdt = dmax(invdt, hydro_cell_invdt<DIM, GRAV>(g, un, gv));
      }
      if (start) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          double u = u0[v];
          u += fl[0][v];
          u -= fxr[v];
          u += fl[1][v];
          sm.part[v * PLANE + so] = u;
          sm.fz[v * PLANE + so] = fl[2][v];
        }
      }
    }

    // the lanes of this warp have read their left neighbour's XMAX-face state of plane p; plane p+1 overwrites it
    __syncwarp();

    // -- prefetch the conserved state three planes ahead (converted at the bottom of the iteration)
    double unext[NV];
    const bool have_next = (p + 3 <= kend + 1);
    if (have_next)
      load_u(p + 3, unext);

    // -- A(p+1): slopes + Hancock trace of plane p+1
    if (do_trace && p + 1 >= kbeg - 1 && p + 1 <= kend) {
      double qc[NV], dq[3][5];
      const double st = g.slope_type;
      const double* __restrict__ q0 = sm.q[s1];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const double c = q0[v * PLANE + so];
        qc[v] = c;
        dq[0][v] = limited_slope(st, g.slope_k, q0[v * PLANE + so - 1], c, q0[v * PLANE + so + 1]);
        dq[1][v] = limited_slope(st, g.slope_k, q0[v * PLANE + so - 32], c, q0[v * PLANE + so + 32]);
        dq[2][v] = limited_slope(st, g.slope_k, sm.q[sp][v * PLANE + so], c, sm.q[s2][v * PLANE + so]);
      }
      if (st == 0.0) { // slope_type 0: first-order (HydroBaseFunctor3D.h:498-520)
#pragma unroll
        for (int d = 0; d < DIM; ++d)
#pragma unroll
          for (int v = 0; v < NV; ++v)
            dq[d][v] = 0.0;
      }
      double s[NV];
      hydro_sources<DIM>(g, qc, dq, s);
      double gh[3] = { 0.0, 0.0, 0.0 }; // gravity predictor 0.5*dt*g on the velocities
      if constexpr (GRAV) {
        if (g.gravity_hancock) {
#pragma unroll
          for (int c = 0; c < DIM; ++c)
            gh[c] = 0.5 * dt * __ldg(a.grav + col + g.sk * (p + 1) + g.sv * c);
        }
      }
      const double dth[3] = { dtd[0] * 0.5, dtd[1] * 0.5, dtd[2] * 0.5 };
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const double add = (GRAV && v >= IU) ? gh[v - IU] : 0.0;
        double bx = fma(s[v], dth[0], qc[v]), by = fma(s[v], dth[1], qc[v]);
        if constexpr (GRAV) {
          bx += add;
          by += add;
        }
        lo[0][v] = fma(-0.5, dq[0][v], bx);
        lo[1][v] = fma(-0.5, dq[1][v], by);
        double mxv = fma(0.5, dq[0][v], bx);
        double myv = fma(0.5, dq[1][v], by);
        if (v == ID) {
          lo[0][v] = dmax(g.smallr, lo[0][v]);
          lo[1][v] = dmax(g.smallr, lo[1][v]);
          mxv = dmax(g.smallr, mxv);
          myv = dmax(g.smallr, myv);
        }
        sm.mx[v * PLANE + so] = mxv;
        sm.my[zb ^ 1][v * PLANE + so] = myv;
        const double bz0 = fma(s[v], dth[2], qc[v]);
        double bzl = bz0, bzh = bz0;
        if constexpr (GRAV) {
          // implementationVersion 1 omits the gravity predictor on the ZMIN-face state
          // (HydroRunFunctors3D.h:1351 vs :1277-1285)
          bzl += (g.impl_version == 1) ? 0.0 : add;
          bzh += add;
        }
        lo[2][v] = fma(-0.5, dq[2][v], bzl);
        double mzv = fma(0.5, dq[2][v], bzh);
        if (v == ID) {
          lo[2][v] = dmax(g.smallr, lo[2][v]);
          mzv = dmax(g.smallr, mzv);
        }
        sm.mz[zb ^ 1][v * PLANE + so] = mzv;
      }
    }

    // -- P(p+3): primitives of plane p+3 replace plane p in the ring (own slot; the x/y neighbours read it as the
    //    centre plane two iterations from now)
    if (have_next)
      prims_to_ring(unext, sp);
    sp = s1;
    __syncthreads();
  }

  // -- CFL reduction for the NEXT step: warp shuffles, then one atomic per CTA (grid reduce)
  invdt = warp_max(invdt);
  if (lane == 0)
    sm.red[row] = invdt;
  __syncthreads();
  if (row == 0) {
    double m = (lane < TY + 4) ? sm.red[lane] : 0.0;
    m = warp_max(m);
    if (lane == 0)
      atomic_max_nonneg(&a.ts->invdt_bits, m);
  }
}

} // namespace ek
