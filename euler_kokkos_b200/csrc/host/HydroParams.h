// Run / mesh / scheme parameters built from the ini file, with the reference's defaults, float
// parsing and derived quantities (src/shared/HydroParams.{h,cpp}, src/shared/GravityParams.h).
// Kokkos types are gone; the sub-domain description that the reference takes from its MPI Cartesian
// communicator (HydroParams::setup_mpi, HydroParams.cpp:186-320) comes from ParallelEnv here.
#pragma once
#include <array>
#include <string>

#include "../../../include/ek_b200.h"
#include "ConfigMap.h"
#include "enums.h"

namespace euler_b200 {

//! rank / size of the process group (one process per GPU); stands in for utils/mpi/ParallelEnv
struct ParallelEnv {
  int rank = 0;
  int nranks = 1;
  int local_rank = 0; // CUDA device index
  //! from the torchrun-style environment (RANK, WORLD_SIZE, LOCAL_RANK); defaults to a single rank
  static ParallelEnv fromEnvironment();
};

struct HydroSettings { // src/shared/HydroParams.h:32-69 (same defaults)
  real_t gamma0 = 1.4, gamma6 = 1.0, cfl = 1.0, slope_type = 2.0;
  int iorder = 1;
  real_t smallr = 1e-8, smallc = 1e-8, smallp = 1e-6, smallpp = 1e-6;
  real_t cIso = 0, Omega0 = 0.0, cp = 0.0, mu = 0.0, kappa = 0.0;
};

struct GravityParams { // src/shared/GravityParams.h:21-39
  bool enabled = false, hancock_predictor_enabled = false, well_balanced_reconstruction_enabled = false;
  GravityParams() = default;
  explicit GravityParams(const ConfigMap& configMap);
};

DimensionType get_dim(const ConfigMap& configMap);

class HydroParams {
public:
  // run parameters
  int nStepmax;
  real_t tEnd;
  int nOutput;
  bool enableOutput;
  bool mhdEnabled;
  int nlog;
  // geometry (LOCAL sizes, as in the reference's MPI build)
  int nx, ny, nz;
  int ghostWidth, nbvar;
  DimensionType dimType;
  int imin, imax, jmin, jmax, kmin, kmax;
  int isize, jsize, ksize;
  real_t xmin, xmax, ymin, ymax, zmin, zmax;
  real_t dx, dy, dz;
  BoundaryConditionType boundary_type_xmin, boundary_type_xmax, boundary_type_ymin, boundary_type_ymax,
    boundary_type_zmin, boundary_type_zmax;
  bool ioVTK, ioHDF5;
  HydroSettings settings;
  int niter_riemann;
  int riemannSolverType;
  int implementationVersion = 0;
  GravityParams gravity;
  // domain decomposition (reference: only in MPI builds)
  int mx = 1, my = 1, mz = 1;
  int myRank = 0, nProcs = 1;
  std::array<int, 3> myMpiPos{ 0, 0, 0 };
  std::array<int, 6> neighborsRank{ 0, 0, 0, 0, 0, 0 };
  std::array<BoundaryConditionType, 6> neighborsBC{};

  HydroParams(const ConfigMap& configMap, const ParallelEnv& par_env);

  void print() const;
  //! the C-ABI view of these parameters (internal faces become EK_BC_COPY)
  ek_params to_ek_params() const;

private:
  void init();
  void setup_decomposition(const ConfigMap& configMap, const ParallelEnv& par_env);
};

} // namespace euler_b200
