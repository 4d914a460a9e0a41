// [Fallback of mhd_pet_tma.cuh for grids whose row length is odd (TMA needs 16-byte row strides), and EK_MHD_PET_TMA=0.]
// Stages P + E + T of the CT-MHD step (primitives, edge electric field, Hancock predictor) fused in ONE kernel
// that marches along z with the primitives and the edge field of the planes around the current one staged in
// shared memory.  In the staged version these were three kernels talking through the Q (8 doubles/cell) and E (3)
// work arrays; ncu: the predictor alone moved 150 GB through L2 per launch at 512^3 (7 neighbour cells x 8
// primitives per cell, L1 hit rate 18 %) and was bound by L2 throughput (11 of ~12 TB/s), and the three kernels
// together took 19 ms of a 59 ms step.  Here a CTA owns a 32 x 8 column of cells (+ 1-cell ring = 34 x 10 box):
//   q[4]  ring of primitive planes k-1 .. k+2, computed from the conserved state U on the fly (U is read ~1.3x)
//   e[2]  edge electric field of planes k and k+1, computed from the staged primitives
// so Q and E never exist in HBM, and the predictor reads its 7-point stencil from shared memory.  The per-cell
// arithmetic is the code of mhd_cells.cuh (mhd_prims_at / mhd_elec_core / mhd_trace_core), unchanged.
#pragma once
#include "mhd_cells.cuh"

namespace ek {

constexpr int PET_TX = 32, PET_TY = 8;
constexpr int PET_BX = PET_TX + 2, PET_BY = PET_TY + 2, PET_BOX = PET_BX * PET_BY; // 34 x 10 cells
constexpr int PET_TILES = (PET_BOX + 31) / 32;                                      // AoSoA tiles per staged plane
constexpr int PET_QPLANE = PET_TILES * 8 * 32, PET_EPLANE = PET_TILES * 3 * 32;     // doubles

struct PetSmem {
  double q[4][PET_QPLANE];
  double e[2][PET_EPLANE];
};

template <int DUMMY = 0>
__global__ void __launch_bounds__(PET_TX* PET_TY, 2) mhd_pet_kernel(const MhdArgs a, const int zchunk)
{
  extern __shared__ __align__(16) unsigned char pet_raw[];
  PetSmem& sm = *reinterpret_cast<PetSmem*>(pet_raw);
  if (a.ts->finished)
    return;
  const Grid& g = a.g;
  const MhdWork w = mhd_work_split(a.work, g.sv);
  const double dt = a.ts->dt;
  const double dtdx = dt / g.dx, dtdy = dt / g.dy, dtdz = dt / g.dz;
  const int t = threadIdx.x;
  const int i0 = 2 + PET_TX * blockIdx.x, j0 = 2 + PET_TY * (int)blockIdx.y; // first cell of the tile (ghosted coordinates)
  const int kbeg = 2 + (int)blockIdx.z * zchunk;
  const int kend = min(kbeg + zchunk, g.ksize - 1); // the predictor runs on planes 2 .. ksize-2
  if (kbeg >= kend)
    return;
  const unsigned sj = (unsigned)g.isize, sk = (unsigned)g.sk;

  auto cell_of_box = [&](int b, int k, MhdIdx& x) {
    const int bj = b / PET_BX, bi = b - bj * PET_BX;
    x.i = i0 - 1 + bi;
    x.j = j0 - 1 + bj;
    x.k = k;
    x.sj = sj;
    x.sk = sk;
    x.c = (unsigned)x.i + sj * (unsigned)x.j + sk * (unsigned)k;
  };
  constexpr int NT = PET_TX * PET_TY;
  static_assert(PET_BOX > NT && PET_BOX <= 2 * NT, "the staging passes below are written for two passes over the box");
  // this thread's two box cells in the staging passes (the second one exists for the first PET_BOX - NT threads)
  const int bA = t, bB = t + NT;
  const bool hasB = bB < PET_BOX;

  // primitives of the box cells of plane k -> ring slot k & 3 (cells the reference never converts are skipped);
  // the loads of both passes are issued before the first conversion
  auto prims_plane = [&](int k) {
    if (k >= g.ksize - 1)
      return;
    double* __restrict__ qs = sm.q[k & 3];
    MhdIdx xA, xB;
    cell_of_box(bA, k, xA);
    cell_of_box(hasB ? bB : bA, k, xB);
    const bool okA = xA.i < g.isize - 1 && xA.j < g.jsize - 1;
    const bool okB = hasB && xB.i < g.isize - 1 && xB.j < g.jsize - 1;
    double uA[11], uB[11];
    mhd_cons_load(g, a.Uin, xA, okA, uA);
    mhd_cons_load(g, a.Uin, xB, okB, uB);
    if (okA) {
      double q[8];
      mhd_cons_to_prim(g, uA, uA[8], uA[9], uA[10], q);
      double* __restrict__ qc = qs + tile_off<8>((unsigned)bA);
#pragma unroll
      for (int v = 0; v < 8; ++v)
        qc[TL * v] = q[v];
    }
    if (okB) {
      double q[8];
      mhd_cons_to_prim(g, uB, uB[8], uB[9], uB[10], q);
      double* __restrict__ qc = qs + tile_off<8>((unsigned)bB);
#pragma unroll
      for (int v = 0; v < 8; ++v)
        qc[TL * v] = q[v];
    }
  };
  // edge field of plane k (needs the primitives of planes k-1 and k) -> slot k & 1.  The face-centred field values
  // come from the conserved state: elec_load issues those loads (before the barrier that publishes the primitives),
  // elec_plane does the arithmetic after it.
  auto elec_cell = [&](int b, int k, MhdIdx& x) {
    const int bj = b / PET_BX, bi = b - bj * PET_BX;
    cell_of_box(b, k, x);
    return k <= g.ksize - 2 && bi >= 1 && bj >= 1 && mhd_elec_active(g, x);
  };
  auto elec_load = [&](int k, MhdElecB& fA, MhdElecB& fB) {
    MhdIdx x;
    const bool okA = elec_cell(bA, k, x);
    mhd_elec_load(g, a.Uin, x, okA, fA);
    const bool okB = hasB && elec_cell(hasB ? bB : bA, k, x);
    mhd_elec_load(g, a.Uin, x, okB, fB);
  };
  auto elec_one = [&](int b, int k, const MhdElecB& f) {
    MhdIdx x;
    if (!elec_cell(b, k, x))
      return;
    const double* __restrict__ q0 = sm.q[k & 3];
    const double* __restrict__ qm = sm.q[(k - 1) & 3];
    MhdElecQ P;
    P.q000 = q0 + tile_off<8>((unsigned)b);
    P.q100 = q0 + tile_off<8>((unsigned)(b - 1));
    P.q010 = q0 + tile_off<8>((unsigned)(b - PET_BX));
    P.q110 = q0 + tile_off<8>((unsigned)(b - 1 - PET_BX));
    P.q001 = qm + tile_off<8>((unsigned)b);
    P.q101 = qm + tile_off<8>((unsigned)(b - 1));
    P.q011 = qm + tile_off<8>((unsigned)(b - PET_BX));
    mhd_elec_core(f, P, sm.e[k & 1] + tile_off<3>((unsigned)b));
  };
  auto elec_plane = [&](int k, const MhdElecB& fA, const MhdElecB& fB) {
    elec_one(bA, k, fA);
    if (hasB)
      elec_one(bB, k, fB);
  };

  const int lx = t & (PET_TX - 1), ly = t / PET_TX;
  const int bt = (1 + lx) + PET_BX * (1 + ly); // this thread's cell inside the box
  // face-centred field values of the predictor of plane k (conserved state, 18 loads)
  auto trace_cell = [&](int k, MhdIdx& x) {
    cell_of_box(bt, k, x);
    return x.i <= g.isize - 2 && x.j <= g.jsize - 2 && mhd_trace_active(g, x);
  };

  MhdElecB eA, eB;
  MhdTraceB fb;
  prims_plane(kbeg - 1);
  prims_plane(kbeg);
  prims_plane(kbeg + 1);
  elec_load(kbeg, eA, eB);
  __syncthreads();
  elec_plane(kbeg, eA, eB);
  elec_load(kbeg + 1, eA, eB);
  elec_plane(kbeg + 1, eA, eB);
  {
    MhdIdx x;
    const bool act = trace_cell(kbeg, x);
    mhd_trace_load(g, a.Uin, x, act, fb);
  }
  __syncthreads();

  for (int p = kbeg; p < kend; ++p) {
    {
      MhdIdx x;
      if (trace_cell(p, x)) {
        const double* __restrict__ q0 = sm.q[p & 3];
        const double* __restrict__ e0 = sm.e[p & 1];
        MhdTraceIn P;
        P.Q0 = q0 + tile_off<8>((unsigned)bt);
        P.Qxm = q0 + tile_off<8>((unsigned)(bt - 1));
        P.Qxp = q0 + tile_off<8>((unsigned)(bt + 1));
        P.Qym = q0 + tile_off<8>((unsigned)(bt - PET_BX));
        P.Qyp = q0 + tile_off<8>((unsigned)(bt + PET_BX));
        P.Qzm = sm.q[(p - 1) & 3] + tile_off<8>((unsigned)bt);
        P.Qzp = sm.q[(p + 1) & 3] + tile_off<8>((unsigned)bt);
        P.E0 = e0 + tile_off<3>((unsigned)bt);
        P.Ei = e0 + tile_off<3>((unsigned)(bt + 1));
        P.Ej = e0 + tile_off<3>((unsigned)(bt + PET_BX));
        P.Ek = sm.e[(p + 1) & 1] + tile_off<3>((unsigned)bt);
        mhd_trace_core(g, fb, P, w.S + tile_off<S_TOTAL>(x.c), dtdx, dtdy, dtdz, x);
      }
    }
    if (p + 1 < kend) {
      // the next plane's field values are requested now and consumed after the two barriers below
      {
        MhdIdx x;
        const bool act = trace_cell(p + 1, x);
        mhd_trace_load(g, a.Uin, x, act, fb);
      }
      prims_plane(p + 2); // slot (p+2)&3 held plane p-2: free
      elec_load(p + 2, eA, eB);
      __syncthreads();
      elec_plane(p + 2, eA, eB); // slot p&1 held plane p: its readers passed the barrier above
      __syncthreads();
    }
  }
}

} // namespace ek
