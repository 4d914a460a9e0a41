// ek_params -> Grid (geometry + physics constants handed to every kernel).  Host-only, shared by
// the C ABI (ek_api.cu) and the CPU emulation of the per-cell stages used by the tests (tests/emul).
#pragma once
#include "../../include/ek_b200.h"
#include "ek_common.cuh"

namespace ek {

inline Grid make_grid(const ek_params& p)
{
  Grid g{};
  g.nx = p.nx; g.ny = p.ny; g.nz = (p.dim == 3) ? p.nz : 1;
  g.gw = p.ghost;
  g.isize = p.nx + 2 * p.ghost;
  g.jsize = p.ny + 2 * p.ghost;
  g.ksize = (p.dim == 3) ? p.nz + 2 * p.ghost : 1;
  g.nbvar = p.nbvar;
  g.sj = g.isize;
  g.sk = (long long)g.isize * g.jsize;
  g.sv = g.sk * g.ksize;
  g.dx = p.dx; g.dy = p.dy; g.dz = p.dz;
  g.invdx = 1.0 / p.dx; g.invdy = 1.0 / p.dy; g.invdz = (p.dim == 3) ? 1.0 / p.dz : 0.0;
  g.entho = 1.0 / (p.gamma0 - 1.0);
  g.slope_k = (p.slope_type > 0.0) ? 0.5 / p.slope_type : 0.0;
  g.gamma0 = p.gamma0; g.gamma6 = p.gamma6; g.smallr = p.smallr; g.smallc = p.smallc;
  g.smallp = p.smallp; g.smallpp = p.smallpp; g.cfl = p.cfl; g.slope_type = p.slope_type;
  for (int f = 0; f < 6; ++f) g.bc[f] = p.bc[f];
  g.impl_version = p.impl_version;
  g.gravity_hancock = p.gravity_enabled && p.gravity_hancock;
  return g;
}

} // namespace ek
