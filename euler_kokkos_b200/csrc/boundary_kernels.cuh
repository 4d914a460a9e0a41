// Ghost-cell fill for physical boundaries (reflective / outflow / periodic).
// Replaces MakeBoundariesFunctor{2D,3D}{,_MHD}<face> (shared/BoundariesFunctors.h:17-1050) and the
// six sequential launches of SolverBase::make_boundaries_serial (shared/SolverBase.cpp:464-512):
// the two faces of one direction write disjoint ghost slabs and read only interior layers of that
// direction, so one launch per direction (X, then Y, then Z) reproduces the reference's result,
// including edge/corner ghosts (each pass covers the full transverse extent, ghosts included).
#pragma once
#include "ek_common.cuh"

namespace ek {

// DIR: 0 x, 1 y, 2 z.  Thread index space: [var][k][j][i] restricted to the 2*gw ghost layers of DIR
// (i fastest for y/z passes => coalesced).  Faces with BC_COPY are left to the halo exchange.
template <int DIR>
__global__ void __launch_bounds__(256) fill_ghosts_kernel(const Grid g, double* __restrict__ U, const int mhd,
                                                          const TimeState* __restrict__ ts)
{
  if (ts && ts->finished)
    return;
  const int gw = g.gw;
  const int n[3] = { g.nx, g.ny, g.nz };
  const int ext0 = (DIR == 0) ? 2 * gw : g.isize;
  const int ext1 = (DIR == 1) ? 2 * gw : g.jsize;
  const int ext2 = (DIR == 2) ? 2 * gw : g.ksize;
  const long long total = (long long)ext0 * ext1 * ext2;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c[3];
    c[0] = (int)(idx % ext0);
    c[1] = (int)((idx / ext0) % ext1);
    c[2] = (int)(idx / ((long long)ext0 * ext1));
    const int l = c[DIR];            // 0..2gw-1 : ghost layer id
    const int side = (l >= gw);      // 0: min face, 1: max face
    const int x = side ? n[DIR] + l : l; // ghosted coordinate along DIR  (n+gw+(l-gw))
    const int bt = g.bc[2 * DIR + side];
    if (bt == BC_COPY)
      continue;
    int x0;
    if (bt == BC_DIRICHLET)
      x0 = side ? 2 * n[DIR] + 2 * gw - 1 - x : 2 * gw - 1 - x;
    else if (bt == BC_NEUMANN)
      x0 = side ? n[DIR] + gw - 1 : gw;
    else // periodic, and any other value (as the reference's else-branch)
      x0 = side ? x - n[DIR] : n[DIR] + x;
    c[DIR] = x;
    const long long dst = c[0] + g.sj * c[1] + g.sk * c[2];
    c[DIR] = x0;
    const long long src = c[0] + g.sj * c[1] + g.sk * c[2];
    for (int v = 0; v < g.nbvar; ++v) {
      double val = U[src + g.sv * v];
      if (bt == BC_DIRICHLET && (v == IU + DIR || (mhd && v == IA + DIR)))
        val = -val;
      U[dst + g.sv * v] = val;
    }
  }
}

} // namespace ek
