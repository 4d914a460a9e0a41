// TEST INFRASTRUCTURE ONLY: the per-cell stages of the CUDA MHD path (euler_kokkos_b200/csrc/
// mhd_cells.cuh, mhd_physics.cuh -- the very functions the kernels call) compiled for the HOST and
// driven by plain loops, so that the kernel logic can be checked against the oracle on a machine
// without a GPU (pytest -m "not gpu").  Never loaded by the product library.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../euler_kokkos_b200/csrc/ek_grid.h"
#include "../../euler_kokkos_b200/csrc/mhd_cells2d.cuh"

using namespace ek;

static MhdIdx idx(const Grid& g, int i, int j, int k)
{
  MhdIdx x;
  x.sj = (unsigned)g.isize;
  x.sk = (unsigned)g.sk;
  x.c = (unsigned)i + x.sj * (unsigned)j + x.sk * (unsigned)k;
  x.i = i; x.j = j; x.k = k;
  return x;
}

template <int RS>
static void step(const Grid& g, const double* Uin, double* Uout, double* work, double dt)
{
  const MhdWork w = mhd_work_split(work, g.sv);
  const double dtdx = dt / g.dx, dtdy = dt / g.dy, dtdz = dt / g.dz;
#define LOOP for (int k = 0; k < g.ksize; ++k) for (int j = 0; j < g.jsize; ++j) for (int i = 0; i < g.isize; ++i) if (MhdIdx x = idx(g, i, j, k); true)
  LOOP mhd_stage_elec(g, Uin, w.Q, w.E, x);
  LOOP mhd_stage_trace(g, Uin, w.Q, w.E, w.S, dtdx, dtdy, dtdz, x);
  LOOP mhd_stage_fluxes<RS>(g, w.S, w.F, x);
  LOOP mhd_stage_emfs(g, w.S, w.EMF, x);
  LOOP mhd_stage_update(g, Uin, w.F, w.EMF, Uout, dtdx, dtdy, dtdz, x);
}

// the update stage as mhd_update_kernel<true> runs it: new state + the part of the next step's CFL maximum that can
// be formed before the boundary fill (work must hold the fluxes and EMFs of a preceding step())
static double update_with_invdt(const Grid& g, const double* Uin, double* Uout, double* work, double dt)
{
  const MhdWork w = mhd_work_split(work, g.sv);
  const double dtdx = dt / g.dx, dtdy = dt / g.dy, dtdz = dt / g.dz;
  double m = 0.0;
  LOOP m = fmax(m, mhd_stage_update_t<true>(g, Uin, w.F, w.EMF, Uout, dtdx, dtdy, dtdz, x));
  return m;
}

template <int RS>
static void step2d(const Grid& g, const double* Uin, double* Uout, double* work, double dt)
{
  const MhdWork2D w = mhd2d_work_split(work, g.sv);
  const double dtdx = dt / g.dx, dtdy = dt / g.dy;
  LOOP mhd2d_stage_elec(g, Uin, w.Q, w.E, x);
  LOOP mhd2d_stage_trace(g, Uin, w.Q, w.E, w.S, dtdx, dtdy, x);
  LOOP mhd2d_stage_faces_edge<RS>(g, w.S, w.F, w.EMF, x);
  LOOP mhd2d_stage_update(g, Uin, w.F, w.EMF, Uout, dtdx, dtdy, x);
}

extern "C" {

// max inverse dt of a ghost-filled state (stage P); Q is left in `work`
double emul_mhd_invdt(const ek_params* p, const double* U, double* work)
{
  const Grid g = make_grid(*p);
  double m = 0.0;
  if (g.ksize == 1) {
    const MhdWork2D w2 = mhd2d_work_split(work, g.sv);
    LOOP m = fmax(m, mhd2d_stage_prims(g, U, w2.Q, x));
    return m;
  }
  const MhdWork w = mhd_work_split(work, g.sv);
  LOOP m = fmax(m, mhd_stage_prims(g, U, w.Q, x));
  return m;
}

// 3-D: the folded CFL reduction.  After emul_mhd_step (fluxes and EMFs still in `work`): redo the update stage the way
// mhd_update_kernel<true> does and return its partial maximum ...
double emul_mhd_update_invdt(const ek_params* p, const double* Uin, double* Uout, double* work, double dt)
{
  return update_with_invdt(make_grid(*p), Uin, Uout, work, dt);
}

// ... and the contribution of the last interior layer along x, y, z of a ghost-filled state (mhd_invdt_layers_kernel)
double emul_mhd_invdt_layers(const ek_params* p, const double* U)
{
  const Grid g = make_grid(*p);
  const int gw = g.gw;
  double m = 0.0;
  for (int k = gw; k < g.ksize - gw; ++k)
    for (int j = gw; j < g.jsize - gw; ++j)
      for (int i = gw; i < g.isize - gw; ++i) {
        if (i != g.isize - gw - 1 && j != g.jsize - gw - 1 && k != g.ksize - gw - 1)
          continue;
        const MhdIdx x = idx(g, i, j, k);
        double q[8];
        mhd_prims_at(g, U, x, q);
        m = fmax(m, mhd_cell_invdt(g, q));
      }
  return m;
}

long long emul_mhd_work_len(const ek_params* p) { return (long long)W_TOTAL * mhd_padded_cells(make_grid(*p).sv); }

// one step with a given dt on a ghost-filled Uin; Uout must be a copy of Uin (ghosts are not written)
int emul_mhd_step(const ek_params* p, const double* Uin, double* Uout, double* work, double dt)
{
  const Grid g = make_grid(*p);
  emul_mhd_invdt(p, Uin, work);
  if (p->dim == 2) {
    switch (p->riemann) {
      case RIEMANN_HLLD: step2d<RIEMANN_HLLD>(g, Uin, Uout, work, dt); return 0;
      case RIEMANN_HLL: step2d<RIEMANN_HLL>(g, Uin, Uout, work, dt); return 0;
      case RIEMANN_LLF: step2d<RIEMANN_LLF>(g, Uin, Uout, work, dt); return 0;
    }
    return 1;
  }
  switch (p->riemann) {
    case RIEMANN_HLLD: step<RIEMANN_HLLD>(g, Uin, Uout, work, dt); return 0;
    case RIEMANN_HLL: step<RIEMANN_HLL>(g, Uin, Uout, work, dt); return 0;
    case RIEMANN_LLF: step<RIEMANN_LLF>(g, Uin, Uout, work, dt); return 0;
  }
  return 1;
}

// the y-chunked launch order of the 3-D stencil kernels (mhd_chunk_cell): every cell of the ghosted box
// must be produced by exactly one thread, with thread%32 == cell%32 (warps aligned to the AoSoA tiles).
// Returns the number of violations.
long long emul_chunk_map_check(int isize, int jsize, int ksize, int yc)
{
  Grid g{};
  g.isize = isize; g.jsize = jsize; g.ksize = ksize;
  g.sj = isize; g.sk = (long long)isize * jsize; g.sv = g.sk * ksize;
  const MhdChunkMap m = mhd_chunk_map(g, yc);
  const unsigned long long nt = mhd_chunk_threads(g, m);
  std::vector<unsigned char> seen((size_t)g.sv, 0);
  long long bad = 0;
  for (unsigned long long t = 0; t < nt + 64; ++t) {
    MhdIdx x;
    if (!mhd_chunk_cell(g, m, (unsigned)t, x))
      continue;
    if (x.c >= (unsigned)g.sv || (x.c & 31u) != ((unsigned)t & 31u) ||
        (unsigned)x.i + (unsigned)isize * ((unsigned)x.j + (unsigned)jsize * (unsigned)x.k) != x.c || seen[x.c]++)
      ++bad;
  }
  for (size_t c = 0; c < seen.size(); ++c)
    bad += (seen[c] != 1);
  return bad;
}
}
