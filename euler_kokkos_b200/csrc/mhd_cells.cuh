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

// Work arrays are private to the library, so they use a tiled "AoSoA" layout instead of the ABI's
// SoA: [tile = c/32][variable][lane = c%32] over the flattened ghosted cell index c.  A warp of 32
// consecutive cells reads one 256-byte segment per variable (coalesced), and the variable offset is
// a compile-time constant: one base address per neighbour cell, immediate offsets for everything else.
template <int NV>
EK_HD size_t tile_off(unsigned c)
{
  return (size_t)(c >> 5) * (unsigned)(NV * 32) + (c & 31u);
}
constexpr int TL = 32; // stride between variables inside a tile

// work-array layout, in units of one padded cell-array (mhd_padded_cells doubles).  The product's 3-D kernels keep the
// primitives and the edge field in shared memory (mhd_pet*.cuh), so the library allocates W3_DEVICE arrays only; the
// staged per-cell functions (CPU emulation of the tests) also use Q and E, placed after them.
enum : int {
  WS = 0,         // compact half-step state, 38
  WF = 38,        // face fluxes Fx,Fy,Fz x (mass, energy, mom_x, mom_y, mom_z), 15
  WEMF = 53,      // EMFs at the lower edges: Ex, Ey, Ez, 3
  W3_DEVICE = 56, // what libek_b200 allocates for MHD_Muscl_3D (512^3: 62.3 GB)
  WQ = 56,        // primitives, 8                                   (emulation only)
  WE = 64,        // edge electric field Ex,Ey,Ez at the lower edges, 3 (emulation only)
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

EK_HD long long mhd_padded_cells(long long sv) { return ((sv + 31) / 32) * 32; }

EK_HD MhdWork mhd_work_split(double* base, long long sv)
{
  const long long T = mhd_padded_cells(sv);
  MhdWork w;
  w.Q = base + (long long)WQ * T;
  w.E = base + (long long)WE * T;
  w.S = base + (long long)WS * T;
  w.F = base + (long long)WF * T;
  w.EMF = base + (long long)WEMF * T;
  return w;
}

// Cell addressing: c = i + isize*(j + jsize*k) as unsigned (the ghosted box of one B200 always has
// < 2^31 cells); neighbour strides in c: 1, sj = isize, sk = isize*jsize.
struct MhdIdx {
  unsigned c, sj, sk;
  int i, j, k;
};

// Thread -> cell map of the stencil stages ("y-chunked"): the ghosted box is cut into slabs of `yc`
// rows; inside a slab the order is k (slowest), then j, then i.  Two consecutive planes of a slab are
// then only yc*isize cells apart in launch order instead of a full plane, so the k-1 / k+1 neighbour
// lines a stage reads are still in L2 when they are needed again (at 518^2 cells x 38 doubles a full
// plane of S is 82 MB; ncu showed S being read 1.7x and Q 2.7x with the flat order).  Each slab-plane
// is a contiguous range of the flat cell index; it is widened to 32-cell boundaries so that warps stay
// aligned with the AoSoA tiles (lanes that fall into the neighbouring slab are masked).
//   t: linear thread index;  returns false for masked threads.   yc <= 0: flat order (t == c).
struct MhdChunkMap {
  unsigned yc;     // rows per slab
  unsigned nchunk; // slabs
  unsigned span;   // threads per slab-plane: yc*isize rounded up to tiles + 1 tile
};
EK_HD MhdChunkMap mhd_chunk_map(const Grid& g, int yc)
{
  MhdChunkMap m;
  m.yc = (yc <= 0) ? 0u : (unsigned)(yc < g.jsize ? yc : g.jsize);
  if (m.yc == 0) {
    m.nchunk = 1;
    m.span = 0;
    return m;
  }
  m.nchunk = ((unsigned)g.jsize + m.yc - 1) / m.yc;
  m.span = ((m.yc * (unsigned)g.isize + 31u) & ~31u) + 32u;
  return m;
}
EK_HD unsigned long long mhd_chunk_threads(const Grid& g, const MhdChunkMap& m)
{
  if (m.yc == 0)
    return (unsigned long long)g.sv;
  return (unsigned long long)m.nchunk * (unsigned)g.ksize * m.span;
}
EK_HD bool mhd_chunk_cell(const Grid& g, const MhdChunkMap& m, unsigned t, MhdIdx& x)
{
  const unsigned sj = (unsigned)g.isize, sk = (unsigned)g.sk;
  x.sj = sj;
  x.sk = sk;
  unsigned c;
  if (m.yc == 0) {
    c = t;
    if (c >= (unsigned)g.sv)
      return false;
  } else {
    const unsigned per_chunk = m.span * (unsigned)g.ksize;
    const unsigned ch = t / per_chunk;
    if (ch >= m.nchunk)
      return false;
    const unsigned r = t - ch * per_chunk;
    const unsigned k = r / m.span; // inside a slab: plane (slowest), then the cells of the slab-plane
    const unsigned o = r - k * m.span;
    const unsigned j0 = ch * m.yc;
    const unsigned j1 = (j0 + m.yc < (unsigned)g.jsize) ? j0 + m.yc : (unsigned)g.jsize;
    const unsigned c0 = sj * j0 + sk * k;       // first cell of the slab-plane
    const unsigned c1 = sj * j1 + sk * k;       // one past its last cell
    c = (c0 & ~31u) + o;
    if (c < c0 || c >= c1)
      return false;
  }
  x.c = c;
  const unsigned k = c / sk;
  const unsigned r = c - k * sk;
  const unsigned j = r / sj;
  x.k = (int)k;
  x.j = (int)j;
  x.i = (int)(r - j * sj);
  return true;
}

// ---- stage P ------------------------------------------------------------------------------------
// Q(i,j,k) for 0 <= i < isize-1 (the last layer is never written, MHDRunFunctors3D.h:120-122);
// returns the cell's inverse-dt contribution (0 outside the interior).
// primitives of one cell (0 <= i < isize-1 etc., checked by the caller) into q[8]
// conserved state of one cell + the field on its three upper faces: u[0..7], u[8..10] (1.0 when !active)
EK_HD void mhd_cons_load(const Grid& g, const double* __restrict__ U, const MhdIdx& x, bool active, double* __restrict__ u)
{
  const double* __restrict__ Uc = U + x.c;
  const long long sv = g.sv;
#pragma unroll
  for (int v = 0; v < 8; ++v)
    u[v] = active ? Uc[sv * v] : 1.0;
  u[8] = active ? Uc[sv * IA + 1] : 1.0;
  u[9] = active ? Uc[sv * IB + x.sj] : 1.0;
  u[10] = active ? Uc[sv * IC + x.sk] : 1.0;
}

EK_HD void mhd_prims_at(const Grid& g, const double* __restrict__ U, const MhdIdx& x, double* __restrict__ q)
{
  double u[11];
  mhd_cons_load(g, U, x, true, u);
  mhd_cons_to_prim(g, u, u[8], u[9], u[10], q);
}

EK_HD double mhd_stage_prims(const Grid& g, const double* __restrict__ U, double* __restrict__ Q, const MhdIdx& x)
{
  if (x.i >= g.isize - 1 || x.j >= g.jsize - 1 || x.k >= g.ksize - 1)
    return 0.0;
  double q[8];
  mhd_prims_at(g, U, x, q);
  double* __restrict__ Qc = Q + tile_off<8>(x.c);
#pragma unroll
  for (int v = 0; v < 8; ++v)
    Qc[TL * v] = q[v];
  const int gw = g.gw;
  if (x.i >= gw && x.i < g.isize - gw && x.j >= gw && x.j < g.jsize - gw && x.k >= gw && x.k < g.ksize - gw)
    return mhd_cell_invdt(g, q);
  return 0.0;
}

// ---- stage E ------------------------------------------------------------------------------------
// E at the three lower edges of cell (i,j,k), 1 <= i <= isize-2 (MHDRunFunctors3D.h:303-352)
// Tile pointers (variable v at p[TL*v]) of the primitives of the 7 cells touching the three lower edges of a cell:
// qXYZ = cell shifted by -X, -Y, -Z.  They may point into the global Q array or into a staged copy in shared memory.
struct MhdElecQ {
  const double *q000, *q100, *q010, *q001, *q110, *q101, *q011;
};

EK_HD bool mhd_elec_active(const Grid& g, const MhdIdx& x)
{
  return !(x.i < 1 || x.j < 1 || x.k < 1 || x.i > g.isize - 2 || x.j > g.jsize - 2 || x.k > g.ksize - 2);
}

// face-centred field values the three lower edges of a cell touch (conserved state U; loaded apart from the
// arithmetic so that a kernel can issue the loads early)
struct MhdElecB {
  double A0, B0, C0, Aym, Azm, Bxm, Bzm, Cxm, Cym;
};

EK_HD void mhd_elec_load(const Grid& g, const double* __restrict__ U, const MhdIdx& x, bool active, MhdElecB& f)
{
  const long long sv = g.sv, sj = x.sj, sk = x.sk;
  const double* __restrict__ UA = U + sv * IA + x.c;
  const double* __restrict__ UB = U + sv * IB + x.c;
  const double* __restrict__ UC = U + sv * IC + x.c;
  f.A0 = active ? UA[0] : 0.0;
  f.B0 = active ? UB[0] : 0.0;
  f.C0 = active ? UC[0] : 0.0;
  f.Aym = active ? UA[-sj] : 0.0;
  f.Azm = active ? UA[-sk] : 0.0;
  f.Bxm = active ? UB[-1] : 0.0;
  f.Bzm = active ? UB[-sk] : 0.0;
  f.Cxm = active ? UC[-1] : 0.0;
  f.Cym = active ? UC[-sj] : 0.0;
}

// Ec: tile pointer of the output (Ex, Ey, Ez at Ec[0], Ec[TL], Ec[2*TL])
EK_HD void mhd_elec_core(const MhdElecB& f, const MhdElecQ& P, double* __restrict__ Ec)
{
  const double *q000 = P.q000, *q100 = P.q100, *q010 = P.q010, *q001 = P.q001, *q110 = P.q110, *q101 = P.q101, *q011 = P.q011;
  const double A0 = f.A0, B0 = f.B0, C0 = f.C0;
  double u, v, w, A, B, C;
  v = 0.25 * (q011[TL * IV] + q010[TL * IV] + q001[TL * IV] + q000[TL * IV]);
  w = 0.25 * (q011[TL * IW] + q010[TL * IW] + q001[TL * IW] + q000[TL * IW]);
  B = 0.5 * (f.Bzm + B0);
  C = 0.5 * (f.Cym + C0);
  Ec[0] = v * C - w * B;
  u = 0.25 * (q101[TL * IU] + q100[TL * IU] + q001[TL * IU] + q000[TL * IU]);
  w = 0.25 * (q101[TL * IW] + q100[TL * IW] + q001[TL * IW] + q000[TL * IW]);
  A = 0.5 * (f.Azm + A0);
  C = 0.5 * (f.Cxm + C0);
  Ec[TL] = w * A - u * C;
  u = 0.25 * (q110[TL * IU] + q100[TL * IU] + q010[TL * IU] + q000[TL * IU]);
  v = 0.25 * (q110[TL * IV] + q100[TL * IV] + q010[TL * IV] + q000[TL * IV]);
  A = 0.5 * (f.Aym + A0);
  B = 0.5 * (f.Bxm + B0);
  Ec[2 * TL] = u * B - v * A;
}

EK_HD void mhd_elec_core(const Grid& g, const double* __restrict__ U, const MhdIdx& x, const MhdElecQ& P, double* __restrict__ Ec)
{
  MhdElecB f;
  mhd_elec_load(g, U, x, true, f);
  mhd_elec_core(f, P, Ec);
}

EK_HD void mhd_stage_elec(const Grid& g, const double* __restrict__ U, const double* __restrict__ Q, double* __restrict__ E,
                          const MhdIdx& x)
{
  if (!mhd_elec_active(g, x))
    return;
  const unsigned c = x.c, sj = x.sj, sk = x.sk;
  MhdElecQ P;
  P.q000 = Q + tile_off<8>(c);
  P.q100 = Q + tile_off<8>(c - 1);
  P.q010 = Q + tile_off<8>(c - sj);
  P.q001 = Q + tile_off<8>(c - sk);
  P.q110 = Q + tile_off<8>(c - 1 - sj);
  P.q101 = Q + tile_off<8>(c - 1 - sk);
  P.q011 = Q + tile_off<8>(c - sj - sk);
  mhd_elec_core(g, U, x, P, E + tile_off<3>(c));
}

// ---- stage T ------------------------------------------------------------------------------------
// Hancock predictor of cell (i,j,k) (trace_unsplit_mhd_3d_simpler, MHDBaseFunctor3D.h:690-914) and of
// its three lower faces.  Cells in [2, isize-3], faces in [2, isize-2] (what stage R consumes).
// Tile pointers of what the Hancock predictor of one cell reads from the work arrays: primitives of the cell and of
// its six neighbours, edge fields of the cell and of its +x, +y, +z neighbours (global arrays or staged copies).
struct MhdTraceIn {
  const double *Q0, *Qxm, *Qxp, *Qym, *Qyp, *Qzm, *Qzp;
  const double *E0, *Ei, *Ej, *Ek;
  // the access interface mhd_trace_core is written against (mhd_pet_tma.cuh has a second implementation that forms the
  // cell-centred field from face values held in shared memory): primitive variable v of the cell / of a neighbour,
  // component c of the edge field of the cell / of its +x, +y, +z neighbour
  EK_HD double q0(int v) const { return Q0[TL * v]; }
  EK_HD double qxm(int v) const { return Qxm[TL * v]; }
  EK_HD double qxp(int v) const { return Qxp[TL * v]; }
  EK_HD double qym(int v) const { return Qym[TL * v]; }
  EK_HD double qyp(int v) const { return Qyp[TL * v]; }
  EK_HD double qzm(int v) const { return Qzm[TL * v]; }
  EK_HD double qzp(int v) const { return Qzp[TL * v]; }
  EK_HD double e0(int c) const { return E0[TL * c]; }
  EK_HD double ei(int c) const { return Ei[TL * c]; }
  EK_HD double ej(int c) const { return Ej[TL * c]; }
  EK_HD double ek(int c) const { return Ek[TL * c]; }
};

EK_HD bool mhd_trace_active(const Grid& g, const MhdIdx& x)
{
  return !(x.i < 2 || x.j < 2 || x.k < 2 || x.i > g.isize - 2 || x.j > g.jsize - 2 || x.k > g.ksize - 2);
}

// face-centred field values the predictor of one cell reads from the conserved state U: the cell's own lower
// faces and their neighbours along the two transverse directions and along the normal (loaded apart from the
// arithmetic so that a kernel can issue the loads early); all of them exist for an active cell (2 <= i <= isize-2)
struct MhdTraceB {
  double A0, Axp, Aym, Ayp, Azm, Azp;
  double B0, Bxm, Bxp, Byp, Bzm, Bzp;
  double C0, Cxm, Cxp, Cym, Cyp, Czp;
};

EK_HD void mhd_trace_load(const Grid& g, const double* __restrict__ U, const MhdIdx& x, bool active, MhdTraceB& f)
{
  const long long sj = x.sj, sk = x.sk, sv = g.sv;
  const double* __restrict__ UA = U + sv * IA + x.c;
  const double* __restrict__ UB = U + sv * IB + x.c;
  const double* __restrict__ UC = U + sv * IC + x.c;
  f.A0 = active ? UA[0] : 0.0;
  f.Axp = active ? UA[1] : 0.0;
  f.Aym = active ? UA[-sj] : 0.0;
  f.Ayp = active ? UA[sj] : 0.0;
  f.Azm = active ? UA[-sk] : 0.0;
  f.Azp = active ? UA[sk] : 0.0;
  f.B0 = active ? UB[0] : 0.0;
  f.Bxm = active ? UB[-1] : 0.0;
  f.Bxp = active ? UB[1] : 0.0;
  f.Byp = active ? UB[sj] : 0.0;
  f.Bzm = active ? UB[-sk] : 0.0;
  f.Bzp = active ? UB[sk] : 0.0;
  f.C0 = active ? UC[0] : 0.0;
  f.Cxm = active ? UC[-1] : 0.0;
  f.Cxp = active ? UC[1] : 0.0;
  f.Cym = active ? UC[-sj] : 0.0;
  f.Cyp = active ? UC[sj] : 0.0;
  f.Czp = active ? UC[sk] : 0.0;
}

// store of one value of the compact state S.  (Measured: a streaming / evict-first store, __stcs, does not keep the
// planes of U in L2 for the neighbouring CTAs -- DRAM reads unchanged at 16 GB -- and costs 0.7 ms per step.)
EK_HD void st_s(double* p, double v) { *p = v; }

// Sc: tile pointer of the cell's compact state (output)
template <class IN>
EK_HD void mhd_trace_core(const Grid& g, const MhdTraceB& fb, const IN& P, double* __restrict__ Sc, double dtdx,
                          double dtdy, double dtdz, const MhdIdx& x)
{
  const bool in_i = x.i <= g.isize - 3, in_j = x.j <= g.jsize - 3, in_k = x.k <= g.ksize - 3;
  const double xst = dmin(g.slope_type, 2.0); // slope_unsplit_mhd_3d, MHDBaseFunctor3D.h:600

  const double AL = fb.A0, BL = fb.B0, CL = fb.C0;
  const double ELL = P.e0(0), FLL = P.e0(1), GLL = P.e0(2);
  // face-centred half step: AL + sAL0 etc. (:872-877, :889-894) and halved transverse slopes (:810-822)
  if (in_j && in_k) {
    const double GLR = P.ej(2), FLR = P.ek(1);
    st_s(Sc + TL * SFA, AL + (+(GLR - GLL) * dtdy * 0.5 - (FLR - FLL) * dtdz * 0.5));
    st_s(Sc + TL * SDAY, mhd_half_slope(0.5 * xst, g.slope_k, fb.Aym, AL, fb.Ayp));
    st_s(Sc + TL * SDAZ, mhd_half_slope(0.5 * xst, g.slope_k, fb.Azm, AL, fb.Azp));
  }
  if (in_i && in_k) {
    const double GRL = P.ei(2), ELR = P.ek(0);
    st_s(Sc + TL * SFB, BL + (-(GRL - GLL) * dtdx * 0.5 + (ELR - ELL) * dtdz * 0.5));
    st_s(Sc + TL * SDBX, mhd_half_slope(0.5 * xst, g.slope_k, fb.Bxm, BL, fb.Bxp));
    st_s(Sc + TL * SDBZ, mhd_half_slope(0.5 * xst, g.slope_k, fb.Bzm, BL, fb.Bzp));
  }
  if (in_i && in_j) {
    const double FRL = P.ei(1), ERL = P.ej(0);
    st_s(Sc + TL * SFC, CL + (+(FRL - FLL) * dtdx * 0.5 - (ERL - ELL) * dtdy * 0.5));
    st_s(Sc + TL * SDCX, mhd_half_slope(0.5 * xst, g.slope_k, fb.Cxm, CL, fb.Cxp));
    st_s(Sc + TL * SDCY, mhd_half_slope(0.5 * xst, g.slope_k, fb.Cym, CL, fb.Cyp));
  }
  if (!(in_i && in_j && in_k))
    return;

  // cell-centred half step, accumulated direction by direction (x, then y, then z: the reference's
  // own summation order) so that only one direction's slopes are live at a time
  const double st = g.slope_type;
  const double sth = 0.5 * st;
  const bool use_slopes = (st == 1.0 || st == 2.0); // slope_unsplit_hydro_3d, :379,:470
  double q[8];
#pragma unroll
  for (int v = 0; v < 8; ++v)
    q[v] = P.q0(v);
  const double r = q[ID], p = q[IP], u = q[IU], v = q[IV], w = q[IW], A = q[IA], B = q[IB], C = q[IC];
  const double ir = rcp_fast(r);
  const double gp = g.gamma0 * p;
  double sr0, sp0, su0, sv0, sw0, sA0, sB0, sC0;
  {
    double d[8];
#pragma unroll
    for (int n = 0; n < 8; ++n)
      d[n] = (use_slopes && n != IA) ? mhd_half_slope(sth, g.slope_k, P.qxm(n), q[n], P.qxp(n)) : 0.0;
    const double drx = d[ID], dpx = d[IP], dux = d[IU], dvx = d[IV], dwx = d[IW], dBx = d[IB], dCx = d[IC];
    const double dAx = 0.5 * (fb.Axp - AL);
    sr0 = (-u * drx - dux * r) * dtdx;
    su0 = (-u * dux - (dpx + B * dBx + C * dCx) * ir) * dtdx;
    sv0 = (-u * dvx + A * dBx * ir) * dtdx;
    sw0 = (-u * dwx + A * dCx * ir) * dtdx;
    sp0 = (-u * dpx - dux * gp) * dtdx;
    sB0 = (v * dAx + A * dvx - u * dBx - B * dux) * dtdx;
    sC0 = (w * dAx + A * dwx - u * dCx - C * dux) * dtdx;
    st_s(Sc + TL * (SX + 0), drx); st_s(Sc + TL * (SX + 1), dpx); st_s(Sc + TL * (SX + 2), dux); st_s(Sc + TL * (SX + 3), dvx); st_s(Sc + TL * (SX + 4), dwx);
    st_s(Sc + TL * (SX + 5), dBx); st_s(Sc + TL * (SX + 6), dCx);
  }
  {
    double d[8];
#pragma unroll
    for (int n = 0; n < 8; ++n)
      d[n] = (use_slopes && n != IB) ? mhd_half_slope(sth, g.slope_k, P.qym(n), q[n], P.qyp(n)) : 0.0;
    const double dry = d[ID], dpy = d[IP], duy = d[IU], dvy = d[IV], dwy = d[IW], dAy = d[IA], dCy = d[IC];
    const double dBy = 0.5 * (fb.Byp - BL);
    sr0 += (-v * dry - dvy * r) * dtdy;
    su0 += (-v * duy + B * dAy * ir) * dtdy;
    sv0 += (-v * dvy - (dpy + A * dAy + C * dCy) * ir) * dtdy;
    sw0 += (-v * dwy + B * dCy * ir) * dtdy;
    sp0 += (-v * dpy - dvy * gp) * dtdy;
    sA0 = (u * dBy + B * duy - v * dAy - A * dvy) * dtdy;
    sC0 += (w * dBy + B * dwy - v * dCy - C * dvy) * dtdy;
    st_s(Sc + TL * (SY + 0), dry); st_s(Sc + TL * (SY + 1), dpy); st_s(Sc + TL * (SY + 2), duy); st_s(Sc + TL * (SY + 3), dvy); st_s(Sc + TL * (SY + 4), dwy);
    st_s(Sc + TL * (SY + 5), dAy); st_s(Sc + TL * (SY + 6), dCy);
  }
  {
    double d[8];
#pragma unroll
    for (int n = 0; n < 8; ++n)
      d[n] = (use_slopes && n != IC) ? mhd_half_slope(sth, g.slope_k, P.qzm(n), q[n], P.qzp(n)) : 0.0;
    const double drz = d[ID], dpz = d[IP], duz = d[IU], dvz = d[IV], dwz = d[IW], dAz = d[IA], dBz = d[IB];
    const double dCz = 0.5 * (fb.Czp - CL);
    sr0 += (-w * drz - dwz * r) * dtdz;
    su0 += (-w * duz + C * dAz * ir) * dtdz;
    sv0 += (-w * dvz + C * dBz * ir) * dtdz;
    sw0 += (-w * dwz - (dpz + A * dAz + B * dBz) * ir) * dtdz;
    sp0 += (-w * dpz - dwz * gp) * dtdz;
    sA0 += (u * dCz + C * duz - w * dAz - A * dwz) * dtdz;
    sB0 += (v * dCz + C * dvz - w * dBz - B * dwz) * dtdz;
    st_s(Sc + TL * (SZ + 0), drz); st_s(Sc + TL * (SZ + 1), dpz); st_s(Sc + TL * (SZ + 2), duz); st_s(Sc + TL * (SZ + 3), dvz); st_s(Sc + TL * (SZ + 4), dwz);
    st_s(Sc + TL * (SZ + 5), dAz); st_s(Sc + TL * (SZ + 6), dBz);
  }
  st_s(Sc + TL * (SC + ID), r + sr0);
  st_s(Sc + TL * (SC + IP), p + sp0);
  st_s(Sc + TL * (SC + IU), u + su0);
  st_s(Sc + TL * (SC + IV), v + sv0);
  st_s(Sc + TL * (SC + IW), w + sw0);
  st_s(Sc + TL * (SC + IA), A + sA0);
  st_s(Sc + TL * (SC + IB), B + sB0);
  st_s(Sc + TL * (SC + IC), C + sC0);
}

EK_HD void mhd_trace_core(const Grid& g, const double* __restrict__ U, const MhdTraceIn& P, double* __restrict__ Sc, double dtdx,
                          double dtdy, double dtdz, const MhdIdx& x)
{
  MhdTraceB fb;
  mhd_trace_load(g, U, x, true, fb);
  mhd_trace_core(g, fb, P, Sc, dtdx, dtdy, dtdz, x);
}

EK_HD void mhd_stage_trace(const Grid& g, const double* __restrict__ U, const double* __restrict__ Q, const double* __restrict__ E,
                           double* __restrict__ S, double dtdx, double dtdy, double dtdz, const MhdIdx& x)
{
  if (!mhd_trace_active(g, x))
    return;
  const unsigned c = x.c;
  MhdTraceIn P;
  P.Q0 = Q + tile_off<8>(c);
  P.Qxm = Q + tile_off<8>(c - 1);
  P.Qxp = Q + tile_off<8>(c + 1);
  P.Qym = Q + tile_off<8>(c - x.sj);
  P.Qyp = Q + tile_off<8>(c + x.sj);
  P.Qzm = Q + tile_off<8>(c - x.sk);
  P.Qzp = Q + tile_off<8>(c + x.sk);
  P.E0 = E + tile_off<3>(c);
  P.Ei = E + tile_off<3>(c + 1);
  P.Ej = E + tile_off<3>(c + x.sj);
  P.Ek = E + tile_off<3>(c + x.sk);
  mhd_trace_core(g, U, P, S + tile_off<S_TOTAL>(c), dtdx, dtdy, dtdz, x);
}

// ---- stage R ------------------------------------------------------------------------------------
// one side of a face: cell state +/- halved slope along the face normal, floors as :896-975
// D: normal direction; sgn = +1 (upper face of the cell = left state of the face), -1 (lower face)
// Sc: tile pointer of the cell (variable v at Sc[TL*v])
// PRHO: pressure floor smallp*rho (2-D trace, MHDBaseFunctor2D.h:985-1100) instead of smallp (3-D, :908)
// The value source is a "view" v with v(var) = variable var of the compact state S of one cell: a tile pointer
// (TileRef, below); any other source of the 38 values of a cell can stand in.
struct TileRef {
  const double* __restrict__ p;
  EK_HD double operator()(int var) const { return p[TL * var]; }
};

template <int D, bool PRHO = false, class V>
EK_HD MhdSide mhd_face_side_v(const Grid& g, const V& Sc, double sgn)
{
  constexpr int SD = (D == 0) ? SX : (D == 1) ? SY : SZ;
  // velocity / field components in the face-normal frame (the reference swaps IU<->IV, IA<->IB for y
  // and IU<->IW, IA<->IC for z, MHDRunFunctors3D.h:1857-1890)
  constexpr int VN = (D == 0) ? IU : (D == 1) ? IV : IW;
  constexpr int VT1 = (D == 0) ? IV : (D == 1) ? IU : IV;
  constexpr int VT2 = (D == 0) ? IW : (D == 1) ? IW : IU;
  constexpr int BT1 = (D == 0) ? IB : (D == 1) ? IA : IB;
  constexpr int BT2 = (D == 0) ? IC : (D == 1) ? IC : IA;
  // slot of the two transverse-field slopes inside the 7-slope group of direction D:
  //   x: (.., dBx, dCx)   y: (.., dAy, dCy)   z: (.., dAz, dBz)
  constexpr int KT1 = (D == 2) ? 6 : 5; // z: bt1 = B (slot 6);  x: bt1 = B (slot 5);  y: bt1 = A (slot 5)
  constexpr int KT2 = (D == 2) ? 5 : 6; // z: bt2 = A (slot 5);  x: bt2 = C (slot 6);  y: bt2 = C (slot 6)
  MhdSide s;
  s.r = dmax(g.smallr, Sc(SC + ID) + sgn * Sc(SD + 0));
  s.p = dmax(PRHO ? g.smallp * s.r : g.smallp, Sc(SC + IP) + sgn * Sc(SD + 1));
  s.un = Sc(SC + VN) + sgn * Sc(SD + 2 + (VN - IU));
  s.ut1 = Sc(SC + VT1) + sgn * Sc(SD + 2 + (VT1 - IU));
  s.ut2 = Sc(SC + VT2) + sgn * Sc(SD + 2 + (VT2 - IU));
  s.bt1 = Sc(SC + BT1) + sgn * Sc(SD + KT1);
  s.bt2 = Sc(SC + BT2) + sgn * Sc(SD + KT2);
  return s;
}

template <int D, bool PRHO = false>
EK_HD MhdSide mhd_face_side(const Grid& g, const double* __restrict__ Sc, double sgn)
{
  return mhd_face_side_v<D, PRHO>(g, TileRef{ Sc }, sgn);
}

// flux through the lower D-face of cell c: f = (mass, energy, mom_x, mom_y, mom_z)
template <int D, int RS>
EK_HD void mhd_face_flux(const Grid& g, const double* __restrict__ S, const MhdIdx& x, double* __restrict__ fout)
{
  const unsigned sd = (D == 0) ? 1u : (D == 1) ? x.sj : x.sk;
  const double* __restrict__ Sc = S + tile_off<S_TOTAL>(x.c);
  const double* __restrict__ Sm = S + tile_off<S_TOTAL>(x.c - sd);
  const MhdSide L = mhd_face_side<D>(g, Sm, +1.0);
  const MhdSide R = mhd_face_side<D>(g, Sc, -1.0);
  const double a = Sc[TL * (SFA + D)]; // both sides carry the same half-step face field
  double f[5];
  mhd_riemann<RS>(g, L, R, a, f);
  fout[0] = f[0];
  fout[1] = f[1];
  if (D == 0) {
    fout[2] = f[2]; fout[3] = f[3]; fout[4] = f[4];
  } else if (D == 1) {
    fout[2] = f[3]; fout[3] = f[2]; fout[4] = f[4];
  } else {
    fout[2] = f[4]; fout[3] = f[3]; fout[4] = f[2];
  }
}

// one corner state of the edge along D3: cell tile pointer Sc, slopes applied with signs s1 (along D1), s2 (along D2)
template <int D3, bool PRHO = false, class V>
EK_HD MhdCorner mhd_edge_corner_v(const Grid& g, const V& Sc, double s1, double s2, double a, double b)
{
  constexpr int D1 = (D3 + 1) % 3, D2 = (D3 + 2) % 3;
  constexpr int S1 = (D1 == 0) ? SX : (D1 == 1) ? SY : SZ;
  constexpr int S2 = (D2 == 0) ? SX : (D2 == 1) ? SY : SZ;
  // slot of the slope of the orthogonal field B_D3 in the groups of D1 and D2:
  //   group x = (B,C) at (5,6), group y = (A,C) at (5,6), group z = (A,B) at (5,6)
  constexpr int K1 = (D1 == 0) ? (D3 == 1 ? 5 : 6) : (D1 == 1) ? (D3 == 0 ? 5 : 6) : (D3 == 0 ? 5 : 6);
  constexpr int K2 = (D2 == 0) ? (D3 == 1 ? 5 : 6) : (D2 == 1) ? (D3 == 0 ? 5 : 6) : (D3 == 0 ? 5 : 6);
  MhdCorner c;
  c.r = dmax(g.smallr, Sc(SC + ID) + (s1 * Sc(S1 + 0) + s2 * Sc(S2 + 0)));
  c.p = dmax(PRHO ? g.smallp * c.r : g.smallp, Sc(SC + IP) + (s1 * Sc(S1 + 1) + s2 * Sc(S2 + 1)));
  c.u = Sc(SC + IU + D1) + (s1 * Sc(S1 + 2 + D1) + s2 * Sc(S2 + 2 + D1));
  c.v = Sc(SC + IU + D2) + (s1 * Sc(S1 + 2 + D2) + s2 * Sc(S2 + 2 + D2));
  c.c = Sc(SC + IA + D3) + (s1 * Sc(S1 + K1) + s2 * Sc(S2 + K2));
  c.a = a;
  c.b = b;
  return c;
}

template <int D3, bool PRHO = false>
EK_HD MhdCorner mhd_edge_corner(const Grid& g, const double* __restrict__ Sc, double s1, double s2, double a, double b)
{
  return mhd_edge_corner_v<D3, PRHO>(g, TileRef{ Sc }, s1, s2, a, b);
}

// EMF at the lower edge of cell c that runs along D3 (compute_emf<EMFX|EMFY|EMFZ> with the corner
// assignment of ComputeEmfAndStoreFunctor3D, MHDRunFunctors3D.h:2199-2226)
template <int D3, bool PRHO = false>
EK_HD double mhd_edge_emf(const Grid& g, const double* __restrict__ S, const MhdIdx& x)
{
  constexpr int D1 = (D3 + 1) % 3, D2 = (D3 + 2) % 3;
  const unsigned s1 = (D1 == 0) ? 1u : (D1 == 1) ? x.sj : x.sk;
  const unsigned s2 = (D2 == 0) ? 1u : (D2 == 1) ? x.sj : x.sk;
  // face field B_D1 lives on D1-faces and varies along D2 (and vice versa); halved transverse slopes
  constexpr int F1 = SFA + D1, F2 = SFA + D2;
  constexpr int DF1 = (D1 == 0) ? (D2 == 1 ? SDAY : SDAZ) : (D1 == 1) ? (D2 == 0 ? SDBX : SDBZ) : (D2 == 0 ? SDCX : SDCY);
  constexpr int DF2 = (D2 == 0) ? (D1 == 1 ? SDAY : SDAZ) : (D2 == 1) ? (D1 == 0 ? SDBX : SDBZ) : (D1 == 0 ? SDCX : SDCY);
  const double* __restrict__ S00 = S + tile_off<S_TOTAL>(x.c);
  const double* __restrict__ S10 = S + tile_off<S_TOTAL>(x.c - s1);
  const double* __restrict__ S01 = S + tile_off<S_TOTAL>(x.c - s2);
  const double* __restrict__ S11 = S + tile_off<S_TOTAL>(x.c - s1 - s2);
  const double aLow = S01[TL * F1] + S01[TL * DF1];  // aLL = aRL
  const double aHigh = S00[TL * F1] - S00[TL * DF1]; // aLR = aRR
  const double bLow = S10[TL * F2] + S10[TL * DF2];  // bLL = bLR
  const double bHigh = S00[TL * F2] - S00[TL * DF2]; // bRL = bRR
  const MhdCorner LL = mhd_edge_corner<D3, PRHO>(g, S11, +1.0, +1.0, aLow, bLow);
  const MhdCorner RL = mhd_edge_corner<D3, PRHO>(g, S01, -1.0, +1.0, aLow, bHigh);
  const MhdCorner LR = mhd_edge_corner<D3, PRHO>(g, S10, +1.0, -1.0, aHigh, bLow);
  const MhdCorner RR = mhd_edge_corner<D3, PRHO>(g, S00, -1.0, -1.0, aHigh, bHigh);
  return mhd_emf2d_hlld(g, LL, RL, LR, RR);
}

EK_HD bool mhd_stage_r_active(const Grid& g, const MhdIdx& x)
{
  const int gw = g.gw;
  return x.i >= gw && x.j >= gw && x.k >= gw && x.i <= g.isize - gw && x.j <= g.jsize - gw && x.k <= g.ksize - gw;
}

template <int RS>
EK_HD void mhd_stage_fluxes(const Grid& g, const double* __restrict__ S, double* __restrict__ F, const MhdIdx& x)
{
  if (!mhd_stage_r_active(g, x))
    return;
  double* __restrict__ Fc = F + tile_off<15>(x.c);
  double f[5];
  mhd_face_flux<0, RS>(g, S, x, f);
#pragma unroll
  for (int v = 0; v < 5; ++v)
    Fc[TL * v] = f[v];
  mhd_face_flux<1, RS>(g, S, x, f);
#pragma unroll
  for (int v = 0; v < 5; ++v)
    Fc[TL * (5 + v)] = f[v];
  mhd_face_flux<2, RS>(g, S, x, f);
#pragma unroll
  for (int v = 0; v < 5; ++v)
    Fc[TL * (10 + v)] = f[v];
}

EK_HD void mhd_stage_emfs(const Grid& g, const double* __restrict__ S, double* __restrict__ EMF, const MhdIdx& x)
{
  if (!mhd_stage_r_active(g, x))
    return;
  double* __restrict__ Ec = EMF + tile_off<3>(x.c);
  Ec[0] = mhd_edge_emf<0>(g, S, x);
  Ec[TL] = mhd_edge_emf<1>(g, S, x);
  Ec[2 * TL] = mhd_edge_emf<2>(g, S, x);
}

// ---- stage U ------------------------------------------------------------------------------------
// UpdateFunctor3D_MHD + UpdateEmfFunctor3D on one interior cell, same accumulation order
// WANT_INVDT: also return the cell's inverse-dt contribution for the NEXT step (ComputeDtFunctor3D_MHD on the new
// state, MHDRunFunctors3D.h:16-86) when it can be formed here: the cell-centred field needs the NEW field on the
// cell's three upper faces, which the +x/+y/+z neighbours own; for neighbours that are interior cells the same CT
// update is repeated here (same expression, same operand order: the value is the one the neighbour stores).  Cells of
// the last interior layer along x, y or z return 0: their upper faces are ghosts, set by the boundary fill / halo
// exchange that follows the step (mhd_invdt_layers_kernel takes those cells afterwards).
//
// Where the fluxes and EMFs come from is a template parameter SRC: flux component v (0..4) of the cell's own lower
// x / y / z face and of the lower face of its +x / +y / +z neighbour, EMF component c (0 = x, 1 = y, 2 = z) at the
// lower edges of the cell and of its +x, +y, +z, +x+y, +x+z, +y+z neighbours.  MhdUpdGlobal reads the F / EMF work
// arrays (stage U as a kernel of its own); the fused R+U kernel reads its shared-memory planes (mhd_ru.cuh).
struct MhdUpdGlobal {
  const double *F0, *Fi, *Fj, *Fk;           // tile pointers of the flux array at c, c+1, c+sj, c+sk
  const double *E0, *Ei, *Ej, *Ek, *Eij, *Eik, *Ejk; // ... of the EMF array
  EK_HD double fx0(int v) const { return F0[TL * v]; }
  EK_HD double fy0(int v) const { return F0[TL * (5 + v)]; }
  EK_HD double fz0(int v) const { return F0[TL * (10 + v)]; }
  EK_HD double fxi(int v) const { return Fi[TL * v]; }
  EK_HD double fyj(int v) const { return Fj[TL * (5 + v)]; }
  EK_HD double fzk(int v) const { return Fk[TL * (10 + v)]; }
  EK_HD double e0(int c) const { return E0[TL * c]; }
  EK_HD double ei(int c) const { return Ei[TL * c]; }
  EK_HD double ej(int c) const { return Ej[TL * c]; }
  EK_HD double ek(int c) const { return Ek[TL * c]; }
  EK_HD double eij(int c) const { return Eij[TL * c]; }
  EK_HD double eik(int c) const { return Eik[TL * c]; }
  EK_HD double ejk(int c) const { return Ejk[TL * c]; }
};

// x: an INTERIOR cell (checked by the caller)
template <bool WANT_INVDT, class SRC>
EK_HD double mhd_update_core(const Grid& g, const double* __restrict__ Uin, double* __restrict__ Uout, const SRC& P, double dtdx,
                             double dtdy, double dtdz, const MhdIdx& x)
{
  const int gw = g.gw;
  const unsigned c = x.c;
  const long long sv = g.sv;
  const double* __restrict__ Uc = Uin + c;
  double* __restrict__ Oc = Uout + c;
  double un[8];
#pragma unroll
  for (int v = 0; v < 5; ++v) {
    double u = Uc[sv * v];
    u += P.fx0(v) * dtdx;
    u -= P.fxi(v) * dtdx;
    u += P.fy0(v) * dtdy;
    u -= P.fyj(v) * dtdy;
    u += P.fz0(v) * dtdz;
    u -= P.fzk(v) * dtdz;
    Oc[sv * v] = u;
    un[v] = u;
  }
  double A = Uc[sv * IA], B = Uc[sv * IB], C = Uc[sv * IC];
  const double ex = P.e0(0), ey = P.e0(1), ez = P.e0(2);
  A += (P.ej(2) - ez) * dtdy;
  B -= (P.ei(2) - ez) * dtdx;
  A -= (P.ek(1) - ey) * dtdz;
  B += (P.ek(0) - ex) * dtdz;
  C += (P.ei(1) - ey) * dtdx;
  C -= (P.ej(0) - ex) * dtdy;
  Oc[sv * IA] = A;
  Oc[sv * IB] = B;
  Oc[sv * IC] = C;
  if constexpr (!WANT_INVDT)
    return 0.0;
  if (x.i >= g.isize - gw - 1 || x.j >= g.jsize - gw - 1 || x.k >= g.ksize - gw - 1)
    return 0.0;
  // the new field on the three upper faces = lower faces of the +x / +y / +z neighbours
  const unsigned sj = x.sj, sk = x.sk;
  double A1 = Uc[sv * IA + 1], B1 = Uc[sv * IB + sj], C1 = Uc[sv * IC + sk];
  A1 += (P.eij(2) - P.ei(2)) * dtdy;
  A1 -= (P.eik(1) - P.ei(1)) * dtdz;
  B1 -= (P.eij(2) - P.ej(2)) * dtdx;
  B1 += (P.ejk(0) - P.ej(0)) * dtdz;
  C1 += (P.eik(1) - P.ek(1)) * dtdx;
  C1 -= (P.ejk(0) - P.ek(0)) * dtdy;
  un[IA] = A;
  un[IB] = B;
  un[IC] = C;
  double q[8];
  mhd_cons_to_prim(g, un, A1, B1, C1, q);
  return mhd_cell_invdt(g, q);
}

template <bool WANT_INVDT>
EK_HD double mhd_stage_update_t(const Grid& g, const double* __restrict__ Uin, const double* __restrict__ F,
                                const double* __restrict__ EMF, double* __restrict__ Uout, double dtdx, double dtdy, double dtdz,
                                const MhdIdx& x)
{
  const int gw = g.gw;
  if (x.i < gw || x.j < gw || x.k < gw || x.i >= g.isize - gw || x.j >= g.jsize - gw || x.k >= g.ksize - gw)
    return 0.0;
  const unsigned c = x.c;
  MhdUpdGlobal P;
  P.F0 = F + tile_off<15>(c);
  P.Fi = F + tile_off<15>(c + 1);
  P.Fj = F + tile_off<15>(c + x.sj);
  P.Fk = F + tile_off<15>(c + x.sk);
  P.E0 = EMF + tile_off<3>(c);
  P.Ei = EMF + tile_off<3>(c + 1);
  P.Ej = EMF + tile_off<3>(c + x.sj);
  P.Ek = EMF + tile_off<3>(c + x.sk);
  P.Eij = EMF + tile_off<3>(c + 1 + x.sj);
  P.Eik = EMF + tile_off<3>(c + 1 + x.sk);
  P.Ejk = EMF + tile_off<3>(c + x.sj + x.sk);
  return mhd_update_core<WANT_INVDT>(g, Uin, Uout, P, dtdx, dtdy, dtdz, x);
}

EK_HD void mhd_stage_update(const Grid& g, const double* __restrict__ Uin, const double* __restrict__ F, const double* __restrict__ EMF,
                            double* __restrict__ Uout, double dtdx, double dtdy, double dtdz, const MhdIdx& x)
{
  mhd_stage_update_t<false>(g, Uin, F, EMF, Uout, dtdx, dtdy, dtdz, x);
}

} // namespace ek
