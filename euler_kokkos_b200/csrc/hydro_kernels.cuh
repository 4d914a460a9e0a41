// Fused Euler MUSCL-Hancock step for sm_100a: cons->prim, limited slopes, Hancock trace, Riemann
// fluxes, conservative update, gravity source term and the next step's CFL reduction in ONE kernel.
//
// Replaces, per time step, the reference's ConvertToPrimitivesFunctor3D + ComputeAndStoreFluxesFunctor3D
// + UpdateFunctor3D (+ GravitySourceTermFunctor3D + ComputeDtFunctor3D) and the deep_copy(U2,U)
// (muscl/SolverHydroMuscl.cpp:306-390, muscl/HydroRunFunctors3D.h:16-820,1903-1985): the conserved
// state is read once and written once; Q, the three flux arrays and the slopes never exist in HBM.
//
// Tiling ("2.5-D z-march"): a CTA owns a TX x TY column of interior cells (TX = 28 so that the
// tile + 2-cell halo is exactly one 32-lane warp wide; one warp per tile row, TY+4 warps) and
// marches along z over a chunk of planes.  All plane-to-plane state lives in shared memory:
//   q[3]   ring of primitive planes p-1, p, p+1 (x/y neighbours AND the thread's own z neighbours)
//   mz[2]  ZMAX-face state of the previous / current plane (own slot only)
//   part   running update U + fx - fx' + fy - fy' + fz of the previous plane (own slot only)
// which keeps the register file free for the three Riemann problems of a cell (lower x, y and z face),
// solved in lock step so that their serial FP64 chains interleave (riemann_hllc_n).  Every Riemann
// problem is solved once per face, every slope and trace once per cell; only the 1-cell ring around
// the tile is computed redundantly.  Two barriers per plane.
#pragma once
#include "hydro_physics.cuh"

namespace ek {

constexpr int HTX = 28; // interior cells per tile row (tile + halo 2 = 32 lanes)

template <int DIM, int TY>
struct HydroSmem {
  static constexpr int NV = HydroTraits<DIM>::NV;
  static constexpr int ROWS = TY + 4;
  static constexpr int PLANE = ROWS * 32; // doubles per variable of the primitive ring (tile + 2-cell halo)
  // face states, fluxes and the running update only ever exist for rows 1 .. TY+2 (the cells that are traced):
  // their planes are two rows shorter and indexed with the row offset by one (so2 = so - 32)
  static constexpr int PLANE2 = (TY + 2) * 32;
  static constexpr int NRING = (DIM == 3) ? 3 : 1;
  double q[NRING][NV * PLANE];        // primitives: ring of planes
  double mx[NV * PLANE2];             // state on the XMAX face of each cell (left state of face i+1)
  double my[NV * PLANE2];             // state on the YMAX face of each cell
  double fx[NV * PLANE2];             // flux*dt/dx through the lower x face of each cell
  double fy[NV * PLANE2];             // flux*dt/dy through the lower y face of each cell
  double mz[(DIM == 3) ? 2 : 1][NV * PLANE2]; // ZMAX-face state, double buffered by plane parity (3-D)
  double part[1][NV * PLANE2];                // partial update of the previous plane (3-D)
  double red[ROWS];                   // per-warp invDt maxima
};

struct HydroArgs {
  Grid g;
  const double* __restrict__ Uin;
  double* __restrict__ Uout;
  const double* __restrict__ grav; // [DIM][k][j][i] or nullptr
  TimeState* ts;
  int zchunk; // planes per CTA along z (3-D)
  int zfirst; // z-chunk of blockIdx.z = 0 ...
  int zstep;  // ... and chunk stride per blockIdx.z (lets one launch take the two end chunks, another the middle ones)
};

// FOLD = 1 ("folded halo rows", 3-D): the CTA has one warp per row 1 .. TY+2 only.  Rows 0 and TY+3 of the tile
// are never traced nor solved -- only their primitives are needed, as y neighbours of rows 1 and TY+2 -- so in the
// FOLD = 0 layout their two warps (and the registers reserved for them) idle through every phase.  Here the warp of
// row 1, which traces but solves no Riemann problem, converts those two rows while the other warps solve theirs.
// With TY = 14 that is 16 warps of which 15 solve Riemann problems (4 per warp scheduler) on a 28 x 14 tile,
// against 12 of 15 warps on a 28 x 11 tile.
template <int DIM, int RS, int TY, bool GRAV, int FOLD = 0>
__global__ void __launch_bounds__(32 * (TY + 4 - 2 * FOLD), 1) hydro_step_kernel(const HydroArgs a)
{
  static_assert(FOLD == 0 || DIM == 3, "folded halo rows: 3-D only");
  constexpr int NV = HydroTraits<DIM>::NV;
  using Smem = HydroSmem<DIM, TY>;
  constexpr int PLANE = Smem::PLANE, PLANE2 = Smem::PLANE2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

  const Grid& g = a.g;
  if (a.ts->finished)
    return;
  const double dt = a.ts->dt;
  const double dtd[3] = { dt * g.invdx, dt * g.invdy, (DIM == 3) ? dt * g.invdz : 0.0 };

  const int lane = threadIdx.x, row = threadIdx.y + FOLD;
  const int xg = g.gw + blockIdx.x * HTX + lane - 2; // ghosted coordinates of this thread's column
  const int yg = g.gw + blockIdx.y * TY + row - 2;
  const int so = row * 32 + lane; // offset inside one variable plane of the primitive ring
  const int so2 = so - 32;        // ... of the face-state / flux / update planes (rows 1 .. TY+2 only)

  const bool inb = (xg < g.isize) && (yg < g.jsize);
  // cells whose slopes/trace are needed: the tile plus a 1-cell ring (inside the ghosted box)
  const bool do_trace = inb && lane >= 1 && lane <= 30 && row >= 1 && row <= TY + 2 && xg <= g.nx + g.gw &&
                        yg <= g.ny + g.gw;
  const bool in_x = lane >= 2 && lane <= 29 && xg < g.nx + g.gw; // interior column range of the tile
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
  // FOLD: rows 0 and TY+3 of plane k, converted by the warp of row 1 (same columns, other rows)
  const bool fold_warp = (FOLD != 0) && threadIdx.y == 0;
  auto fold_rows_to_ring = [&](int k, int slot) {
    const bool in0 = (xg < g.isize) && (yg - 1 < g.jsize), in3 = (xg < g.isize) && (yg + TY + 2 < g.jsize);
    double u0[NV], u3[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      u0[v] = in0 ? __ldg(Uin + col - g.sj + g.sk * k + g.sv * v) : 1.0;
      u3[v] = in3 ? __ldg(Uin + col + g.sj * (TY + 2) + g.sk * k + g.sv * v) : 1.0;
    }
    double q0[NV], q3[NV], rd;
    hydro_cons_to_prim<DIM>(g, u0, q0, rd);
    hydro_cons_to_prim<DIM>(g, u3, q3, rd);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      sm.q[slot][v * PLANE + so - 32] = q0[v];
      sm.q[slot][v * PLANE + so + 32 * (TY + 2)] = q3[v];
    }
  };

  int kbeg = 0, kend = 1;
  if constexpr (DIM == 3) {
    kbeg = g.gw + (a.zfirst + (int)blockIdx.z * a.zstep) * a.zchunk;
    kend = min(kbeg + a.zchunk, g.nz + g.gw);
  }
  double invdt = 0.0;

  // prologue: primitives of the first planes into the ring (slot of plane k = (k - kbeg + 2) % 3)
  {
    double u[NV];
    if constexpr (DIM == 3) {
      load_u(kbeg - 2, u);
      prims_to_ring(u, 0);
      load_u(kbeg - 1, u);
      prims_to_ring(u, 1);
      load_u(kbeg, u);
      prims_to_ring(u, 2);
      if (fold_warp) {
        fold_rows_to_ring(kbeg - 2, 0);
        fold_rows_to_ring(kbeg - 1, 1);
        fold_rows_to_ring(kbeg, 2);
      }
    } else {
      load_u(0, u);
      prims_to_ring(u, 0);
    }
  }
  __syncthreads();

  const int pbeg = (DIM == 3) ? kbeg - 1 : 0;
  const int pend = (DIM == 3) ? kend : 0; // inclusive
  int s0 = (DIM == 3) ? 1 : 0;            // ring slot of plane p

  for (int p = pbeg; p <= pend; ++p) {
    const int sm1 = (s0 == 0) ? 2 : s0 - 1, sp1 = (s0 == 2) ? 0 : s0 + 1;
    const int zb = p & 1; // mz buffer written for plane p; plane p-1 sits in zb^1
    // -- prefetch the conserved state two planes ahead (converted at the bottom of the loop)
    double unext[NV];
    const bool have_next = (DIM == 3) && (p + 2 <= kend + 1);
    if (have_next)
      load_u(p + 2, unext);

    // -- slopes + Hancock trace of this cell
    double lo[3][5]; // states on the XMIN, YMIN, ZMIN faces (right states of the lower faces)
    if (do_trace) {
      double qc[NV], dq[3][5];
      const double st = g.slope_type;
      const double* __restrict__ q0 = sm.q[(DIM == 3) ? s0 : 0];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const double c = q0[v * PLANE + so];
        qc[v] = c;
        dq[0][v] = limited_slope(st, g.slope_k, q0[v * PLANE + so - 1], c, q0[v * PLANE + so + 1]);
        dq[1][v] = limited_slope(st, g.slope_k, q0[v * PLANE + so - 32], c, q0[v * PLANE + so + 32]);
        if constexpr (DIM == 3)
          dq[2][v] = limited_slope(st, g.slope_k, sm.q[sm1][v * PLANE + so], c, sm.q[sp1][v * PLANE + so]);
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
            gh[c] = 0.5 * dt * __ldg(a.grav + col + g.sk * p + g.sv * c);
        }
      }
      // face states q -/+ dq/2 + s*dt/(2 dx) [+ gravity predictor]: the cell-centred half-step value is formed
      // once per direction (one FMA), the two faces are one FMA each
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
        sm.mx[v * PLANE2 + so2] = mxv;
        sm.my[v * PLANE2 + so2] = myv;
        if constexpr (DIM == 3) {
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
          sm.mz[zb][v * PLANE2 + so2] = mzv;
        }
      }
    } else {
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int v = 0; v < 5; ++v)
          lo[d][v] = 1.0; // halo lanes ride along in the lock-step solve; their fluxes are discarded
    }
    __syncthreads();

    // -- Riemann problems at the lower x / y / z faces of this cell, solved in lock step
    double fl[3][5];
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int v = 0; v < 5; ++v)
        fl[d][v] = 0.0;
    if (row_flux) {
      double ql[3][5], qr[3][5], f[3][5];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        ql[0][v] = sm.mx[v * PLANE2 + so2 - 1];
        ql[1][v] = sm.my[v * PLANE2 + so2 - 32];
        qr[0][v] = lo[0][v];
        qr[1][v] = lo[1][v];
        if constexpr (DIM == 3) {
          ql[2][v] = sm.mz[zb ^ 1][v * PLANE2 + so2];
          qr[2][v] = lo[2][v];
        }
      }
      // rotate into the face-normal frame: normal velocity first (HydroRunFunctors3D.h:597-608,692-703)
      { double t = ql[1][IU]; ql[1][IU] = ql[1][IV]; ql[1][IV] = t; }
      { double t = qr[1][IU]; qr[1][IU] = qr[1][IV]; qr[1][IV] = t; }
      if constexpr (DIM == 3) {
        { double t = ql[2][IU]; ql[2][IU] = ql[2][IW]; ql[2][IW] = t; }
        { double t = qr[2][IU]; qr[2][IU] = qr[2][IW]; qr[2][IW] = t; }
      }
      riemann_hydro_n<DIM, RS, DIM>(g, ql, qr, f);
      { double t = f[1][IU]; f[1][IU] = f[1][IV]; f[1][IV] = t; }
      if constexpr (DIM == 3) {
        double t = f[2][IU]; f[2][IU] = f[2][IW]; f[2][IW] = t;
      }
      if (do_fx) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          fl[0][v] = f[0][v] * dtd[0];
          sm.fx[v * PLANE2 + so2] = fl[0][v];
        }
      }
      if (do_fy) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          fl[1][v] = f[1][v] * dtd[1];
          sm.fy[v * PLANE2 + so2] = fl[1][v];
        }
      }
      if constexpr (DIM == 3) {
        // face between plane p-1 and p (plane p-1 has been traced except at p == pbeg)
        if (do_cell && p > pbeg) {
#pragma unroll
          for (int v = 0; v < NV; ++v)
            fl[2][v] = f[2][v] * dtd[2];
        }
      }
    } else if (fold_warp && have_next) {
      // rows 0 / TY+3 of plane p+2 into the ring slot of plane p-1: nobody reads those rows of that slot before
      // plane p+2 is the centre plane, two iterations (four barriers) from now
      fold_rows_to_ring(p + 2, sm1);
    }
    __syncthreads();

    // -- finish plane p-1 (3-D) / plane p (2-D), start the accumulation of plane p
    if (do_cell) {
      if constexpr (DIM == 3) {
        if (p - 1 >= kbeg) {
          double un[NV];
#pragma unroll
          for (int v = 0; v < NV; ++v)
            un[v] = sm.part[0][v * PLANE2 + so2] - fl[2][v];
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
        if (p < kend) {
          const long long o = col + g.sk * p;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            double u = __ldg(Uin + o + g.sv * v);
            u += fl[0][v];
            u -= sm.fx[v * PLANE2 + so2 + 1];
            u += fl[1][v];
            u -= sm.fy[v * PLANE2 + so2 + 32];
            u += fl[2][v];
            sm.part[0][v * PLANE2 + so2] = u;
          }
        }
      } else {
        double un[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          double u = __ldg(Uin + col + g.sv * v);
          u += fl[0][v];
          u -= sm.fx[v * PLANE2 + so2 + 1];
          u += fl[1][v];
          u -= sm.fy[v * PLANE2 + so2 + 32];
          un[v] = u;
        }
        double gv[3] = { 0.0, 0.0, 0.0 };
        if constexpr (GRAV) {
          // GravitySourceTermFunctor2D (HydroRunFunctors2D.h:1594-1625): rhoNew floored at smallr
          const double rhoOld = __ldg(Uin + col);
          const double rhoNew = dmax(g.smallr, un[ID]);
          const double ir = 1.0 / rhoNew;
          const double ek0 = 0.5 * (un[IU] * un[IU] + un[IV] * un[IV]) * ir;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            gv[c] = __ldg(a.grav + col + g.sv * c);
            un[IU + c] += 0.5 * dt * gv[c] * (rhoOld + rhoNew);
          }
          const double ek1 = 0.5 * (un[IU] * un[IU] + un[IV] * un[IV]) * ir;
          un[IP] += (ek1 - ek0);
        }
#pragma unroll
        for (int v = 0; v < NV; ++v)
          a.Uout[col + g.sv * v] = un[v];
        invdt = dmax(invdt, hydro_cell_invdt<DIM, GRAV>(g, un, gv));
      }
    }

    // -- primitives of plane p+2 replace plane p-1 in the ring (own slot: this thread is the only
    //    reader of its plane p-1 entry, and it is done with it); visible to the neighbours two
    //    barriers before they read it as plane "p" of iteration p+2
    if constexpr (DIM == 3) {
      if (have_next)
        prims_to_ring(unext, sm1);
      s0 = sp1;
    }
  }

  // -- CFL reduction for the NEXT step: warp shuffles, then one atomic per CTA (grid reduce)
  invdt = warp_max(invdt);
  if (lane == 0)
    sm.red[threadIdx.y] = invdt;
  __syncthreads();
  if (threadIdx.y == 0) {
    double m = (lane < TY + 4 - 2 * FOLD) ? sm.red[lane] : 0.0;
    m = warp_max(m);
    if (lane == 0)
      atomic_max_nonneg(&a.ts->invdt_bits, m);
  }
}

// Stand-alone CFL reduction (first step, ek_compute_dt): ComputeDtFunctor{2D,3D} / Gravity variant.
template <int DIM, bool GRAV>
__global__ void __launch_bounds__(256) hydro_invdt_kernel(const Grid g, const double* __restrict__ U, const double* __restrict__ grav,
                                                          unsigned long long* __restrict__ out_bits)
{
  constexpr int NV = HydroTraits<DIM>::NV;
  const long long ncell = (long long)g.nx * g.ny * g.nz;
  double m = 0.0;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(c % g.nx) + g.gw;
    const int j = (int)((c / g.nx) % g.ny) + g.gw;
    const int k = (DIM == 3) ? (int)(c / ((long long)g.nx * g.ny)) + g.gw : 0;
    const long long o = i + g.sj * j + g.sk * k;
    double u[NV], gv[3] = { 0.0, 0.0, 0.0 };
#pragma unroll
    for (int v = 0; v < NV; ++v)
      u[v] = U[o + g.sv * v];
    if constexpr (GRAV) {
#pragma unroll
      for (int cc = 0; cc < DIM; ++cc)
        gv[cc] = grav[o + g.sv * cc];
    }
    m = dmax(m, hydro_cell_invdt<DIM, GRAV>(g, u, gv));
  }
  __shared__ double red[8];
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0)
    red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    double r = (threadIdx.x < 8) ? red[threadIdx.x] : 0.0;
    r = warp_max(r);
    if (threadIdx.x == 0)
      atomic_max_nonneg(out_bits, r);
  }
}

} // namespace ek
