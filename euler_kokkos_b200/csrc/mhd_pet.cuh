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
  const int i0 = 2 + PET_TX * blockIdx.x, j0 = 2 + PET_TY * blockIdx.y; // first cell of the tile (ghosted coordinates)
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
  // primitives of the box cells of plane k -> ring slot k & 3 (cells the reference never converts are skipped)
  auto prims_plane = [&](int k) {
    if (k >= g.ksize - 1)
      return;
    double* __restrict__ qs = sm.q[k & 3];
    for (int b = t; b < PET_BOX; b += PET_TX * PET_TY) {
      MhdIdx x;
      cell_of_box(b, k, x);
      if (x.i >= g.isize - 1 || x.j >= g.jsize - 1)
        continue;
      double q[8];
      mhd_prims_at(g, a.Uin, x, q);
      double* __restrict__ qc = qs + tile_off<8>((unsigned)b);
#pragma unroll
      for (int v = 0; v < 8; ++v)
        qc[TL * v] = q[v];
    }
  };
  // edge field of plane k (needs the primitives of planes k-1 and k) -> slot k & 1
  auto elec_plane = [&](int k) {
    if (k > g.ksize - 2)
      return;
    const double* __restrict__ q0 = sm.q[k & 3];
    const double* __restrict__ qm = sm.q[(k - 1) & 3];
    double* __restrict__ es = sm.e[k & 1];
    for (int b = t; b < PET_BOX; b += PET_TX * PET_TY) {
      const int bj = b / PET_BX, bi = b - bj * PET_BX;
      if (bi < 1 || bj < 1)
        continue;
      MhdIdx x;
      cell_of_box(b, k, x);
      if (!mhd_elec_active(g, x))
        continue;
      MhdElecQ P;
      P.q000 = q0 + tile_off<8>((unsigned)b);
      P.q100 = q0 + tile_off<8>((unsigned)(b - 1));
      P.q010 = q0 + tile_off<8>((unsigned)(b - PET_BX));
      P.q110 = q0 + tile_off<8>((unsigned)(b - 1 - PET_BX));
      P.q001 = qm + tile_off<8>((unsigned)b);
      P.q101 = qm + tile_off<8>((unsigned)(b - 1));
      P.q011 = qm + tile_off<8>((unsigned)(b - PET_BX));
      mhd_elec_core(g, a.Uin, x, P, es + tile_off<3>((unsigned)b));
    }
  };

  prims_plane(kbeg - 1);
  prims_plane(kbeg);
  prims_plane(kbeg + 1);
  __syncthreads();
  elec_plane(kbeg);
  elec_plane(kbeg + 1);
  __syncthreads();

  const int lx = t & (PET_TX - 1), ly = t / PET_TX;
  const int bt = (1 + lx) + PET_BX * (1 + ly); // this thread's cell inside the box
  for (int p = kbeg; p < kend; ++p) {
    {
      MhdIdx x;
      cell_of_box(bt, p, x);
      if (x.i <= g.isize - 2 && x.j <= g.jsize - 2 && mhd_trace_active(g, x)) {
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
        mhd_trace_core(g, a.Uin, P, w.S + tile_off<S_TOTAL>(x.c), dtdx, dtdy, dtdz, x);
      }
    }
    if (p + 1 < kend) {
      prims_plane(p + 2); // slot (p+2)&3 held plane p-2: free
      __syncthreads();
      elec_plane(p + 2);  // slot p&1 held plane p: its readers passed the barrier above
      __syncthreads();
    }
  }
}

} // namespace ek
