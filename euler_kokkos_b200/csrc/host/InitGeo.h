// Cell geometry shared by the initial-condition files (cell-centre coordinates exactly as the reference's init
// functors compute them: xmin + dx/2 + (i + nx*i_mpi - ghostWidth)*dx).
#pragma once
#include <cstddef>

#include "ConfigMap.h"
#include "HydroParams.h"
#include "enums.h"

namespace euler_b200 {

struct Geo {
  int dim, gw, nx, ny, nz, isize, jsize, ksize;
  int i_mpi, j_mpi, k_mpi;
  size_t sj, sk, sv;
  double xmin, xmax, ymin, ymax, zmin, zmax, dx, dy, dz;
  explicit Geo(const HydroParams& p)
  {
    dim = p.dimType; gw = p.ghostWidth; nx = p.nx; ny = p.ny; nz = p.nz;
    isize = p.isize; jsize = p.jsize; ksize = (dim == 3) ? p.ksize : 1;
    i_mpi = p.myMpiPos[0]; j_mpi = p.myMpiPos[1]; k_mpi = p.myMpiPos[2];
    sj = isize; sk = (size_t)isize * jsize; sv = sk * ksize;
    xmin = p.xmin; xmax = p.xmax; ymin = p.ymin; ymax = p.ymax; zmin = p.zmin; zmax = p.zmax;
    dx = p.dx; dy = p.dy; dz = p.dz;
  }
  double x(int i) const { return xmin + dx / 2 + (i + nx * i_mpi - gw) * dx; }
  double y(int j) const { return ymin + dy / 2 + (j + ny * j_mpi - gw) * dy; }
  double z(int k) const { return zmin + dz / 2 + (k + nz * k_mpi - gw) * dz; }
  size_t at(int i, int j, int k) const { return (size_t)i + sj * j + sk * k; }
};

template <class F>
void for_each_cell(const Geo& g, F f)
{
#pragma omp parallel for schedule(static)
  for (int k = 0; k < g.ksize; ++k)
    for (int j = 0; j < g.jsize; ++j)
      for (int i = 0; i < g.isize; ++i)
        f(i, j, k);
}

// InitConditionsMore.cpp: the remaining problems of the reference.  Return false when the
// requested variant does not exist in the reference either.
bool init_kelvin_helmholtz(const HydroParams& p, const ConfigMap& cm, double* U);
bool init_gresho_vortex(const HydroParams& p, const ConfigMap& cm, double* U);
bool init_four_quadrant(const HydroParams& p, const ConfigMap& cm, double* U);
bool init_isentropic_vortex(const HydroParams& p, const ConfigMap& cm, double* U);
bool init_mhd_blast(const HydroParams& p, const ConfigMap& cm, double* U);
bool init_mhd_implode(const HydroParams& p, const ConfigMap& cm, double* U);
bool init_mhd_brio_wu(const HydroParams& p, const ConfigMap& cm, double* U);
bool init_mhd_kelvin_helmholtz(const HydroParams& p, const ConfigMap& cm, double* U);
bool init_mhd_rotor(const HydroParams& p, const ConfigMap& cm, double* U);
bool init_mhd_field_loop(const HydroParams& p, const ConfigMap& cm, double* U);

} // namespace euler_b200
