#include "HydroParams.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

namespace euler_b200 {

ParallelEnv ParallelEnv::fromEnvironment()
{
  ParallelEnv e;
  if (const char* s = std::getenv("RANK")) e.rank = std::atoi(s);
  if (const char* s = std::getenv("WORLD_SIZE")) e.nranks = std::atoi(s);
  if (const char* s = std::getenv("LOCAL_RANK")) e.local_rank = std::atoi(s);
  if (e.nranks < 1) e.nranks = 1;
  return e;
}

GravityParams::GravityParams(const ConfigMap& configMap)
  : enabled(configMap.getBool("gravity", "enabled", false))
  , hancock_predictor_enabled(configMap.getBool("gravity", "hancock_predictor_enabled", false))
  , well_balanced_reconstruction_enabled(configMap.getBool("gravity", "well_balanced_reconstruction_enabled", false))
{}

// src/shared/HydroParams.cpp:16-33
DimensionType get_dim(const ConfigMap& configMap)
{
  const std::string solver_name = configMap.getString("run", "solver_name", "unknown");
  if (solver_name == "Hydro_Muscl_2D" || solver_name == "MHD_Muscl_2D")
    return TWO_D;
  if (solver_name == "Hydro_Muscl_3D" || solver_name == "MHD_Muscl_3D")
    return THREE_D;
  std::cerr << "Solver name not valid : " << solver_name << "\n";
  return TWO_D;
}

// src/shared/HydroParams.cpp:37-180: same keys, same defaults, same (float) parsing
HydroParams::HydroParams(const ConfigMap& configMap, const ParallelEnv& par_env)
  : nStepmax(configMap.getInteger("run", "nstepmax", 1000))
  , tEnd(configMap.getFloat("run", "tend", 0.0))
  , nOutput(configMap.getInteger("run", "noutput", 100))
  , enableOutput(nOutput == 0 ? false : true)
  , mhdEnabled(false)
  , nlog(configMap.getInteger("run", "nlog", 10))
  , nx(configMap.getInteger("mesh", "nx", 1))
  , ny(configMap.getInteger("mesh", "ny", 1))
  , nz(configMap.getInteger("mesh", "nz", 1))
  , ghostWidth(2)
  , nbvar(4)
  , dimType(TWO_D)
  , imin(0), imax(0), jmin(0), jmax(0), kmin(0), kmax(0)
  , isize(0), jsize(0), ksize(0)
  , xmin(configMap.getFloat("mesh", "xmin", 0.0))
  , xmax(configMap.getFloat("mesh", "xmax", 1.0))
  , ymin(configMap.getFloat("mesh", "ymin", 0.0))
  , ymax(configMap.getFloat("mesh", "ymax", 1.0))
  , zmin(configMap.getFloat("mesh", "zmin", 0.0))
  , zmax(configMap.getFloat("mesh", "zmax", 1.0))
  , dx(0.0), dy(0.0), dz(0.0)
  , boundary_type_xmin(static_cast<BoundaryConditionType>(configMap.getInteger("mesh", "boundary_type_xmin", BC_DIRICHLET)))
  , boundary_type_xmax(static_cast<BoundaryConditionType>(configMap.getInteger("mesh", "boundary_type_xmax", BC_DIRICHLET)))
  , boundary_type_ymin(static_cast<BoundaryConditionType>(configMap.getInteger("mesh", "boundary_type_ymin", BC_DIRICHLET)))
  , boundary_type_ymax(static_cast<BoundaryConditionType>(configMap.getInteger("mesh", "boundary_type_ymax", BC_DIRICHLET)))
  , boundary_type_zmin(static_cast<BoundaryConditionType>(configMap.getInteger("mesh", "boundary_type_zmin", BC_DIRICHLET)))
  , boundary_type_zmax(static_cast<BoundaryConditionType>(configMap.getInteger("mesh", "boundary_type_zmax", BC_DIRICHLET)))
  , ioVTK(true)
  , ioHDF5(false)
  , settings()
  , niter_riemann(configMap.getInteger("hydro", "niter_riemann", 10))
  , riemannSolverType()
  , implementationVersion(0)
  , gravity(configMap)
{
  dimType = get_dim(configMap);
  const std::string solver_name = configMap.getString("run", "solver_name", "unknown");
  if (solver_name == "Hydro_Muscl_2D") {
    nbvar = 4; ghostWidth = 2;
  } else if (solver_name == "Hydro_Muscl_3D") {
    nbvar = 5; ghostWidth = 2;
  } else if (solver_name == "MHD_Muscl_2D" || solver_name == "MHD_Muscl_3D") {
    nbvar = 8; ghostWidth = 3; mhdEnabled = true;
  } else {
    std::cerr << "Solver name not valid : " << solver_name << "\n";
  }

  const std::string riemannSolverStr = configMap.getString("hydro", "riemann", "approx");
  if (riemannSolverStr == "approx") riemannSolverType = RIEMANN_APPROX;
  else if (riemannSolverStr == "llf") riemannSolverType = RIEMANN_LLF;
  else if (riemannSolverStr == "hll") riemannSolverType = RIEMANN_HLL;
  else if (riemannSolverStr == "hllc") riemannSolverType = RIEMANN_HLLC;
  else if (riemannSolverStr == "hlld") riemannSolverType = RIEMANN_HLLD;
  else {
    std::cout << "Riemann Solver specified in parameter file is invalid\n";
    std::cout << "Use the default one : approx\n";
    riemannSolverType = RIEMANN_APPROX;
  }

  implementationVersion = static_cast<int>(configMap.getFloat("OTHER", "implementationVersion", 0));
  if (implementationVersion != 0 && implementationVersion != 1 && implementationVersion != 2) {
    std::cout << "Implementation version is invalid (must be 0, 1 or 2)\n";
    std::cout << "Use the default : 0\n";
    implementationVersion = 0;
  }

  settings.gamma0 = configMap.getFloat("hydro", "gamma0", 1.4);
  settings.cfl = configMap.getFloat("hydro", "cfl", 0.5);
  settings.iorder = configMap.getInteger("hydro", "iorder", 2);
  settings.slope_type = configMap.getFloat("hydro", "slope_type", 1.0);
  settings.smallc = configMap.getFloat("hydro", "smallc", 1e-10);
  settings.smallr = configMap.getFloat("hydro", "smallr", 1e-10);
  settings.cp = configMap.getFloat("hydro", "cp", 0.0);
  settings.mu = configMap.getFloat("hydro", "mu", 0.0);
  settings.kappa = configMap.getFloat("hydro", "kappa", 0.0);

  init();
  setup_decomposition(configMap, par_env);
}

// src/shared/HydroParams.cpp:327-359
void HydroParams::init()
{
  imin = 0; jmin = 0; kmin = 0;
  imax = nx - 1 + 2 * ghostWidth;
  jmax = ny - 1 + 2 * ghostWidth;
  kmax = nz - 1 + 2 * ghostWidth;
  isize = imax - imin + 1;
  jsize = jmax - jmin + 1;
  ksize = kmax - kmin + 1;
  dx = (xmax - xmin) / nx;
  dy = (ymax - ymin) / ny;
  dz = (zmax - zmin) / nz;
  settings.smallp = settings.smallc * settings.smallc / settings.gamma0;
  settings.smallpp = settings.smallr * settings.smallp;
  settings.gamma6 = (settings.gamma0 + 1.0) / (2.0 * settings.gamma0);
}

// src/shared/HydroParams.cpp:186-320 (setup_mpi) -- Cartesian ranks as MPI_Cart_create lays them out
// (last dimension fastest), topology always periodic (src/main.cpp:74-77), outer faces overridden
// by the physical boundary types.
void HydroParams::setup_decomposition(const ConfigMap& configMap, const ParallelEnv& par_env)
{
  mx = configMap.getInteger("mpi", "mx", 1);
  my = configMap.getInteger("mpi", "my", 1);
  mz = configMap.getInteger("mpi", "mz", 1);
  if (dimType == TWO_D) mz = 1;
  nProcs = par_env.nranks;
  myRank = par_env.rank;
  if (nProcs == 1) {
    // serial build of the reference: [mpi] is ignored
    mx = my = mz = 1;
  } else if (nProcs != mx * my * mz) {
    // convenience beyond the reference: a default slab decomposition along the slowest axis
    if (mx * my * mz == 1) {
      if (dimType == THREE_D) mz = nProcs; else my = nProcs;
    } else {
      std::cerr << "Inconsistent cartesian virtual topology geometry; \n mx*my*mz must match the number of ranks !!!\n";
      std::exit(EXIT_FAILURE);
    }
  }
  if (dimType == TWO_D) {
    myMpiPos = { myRank / my, myRank % my, 0 };
  } else {
    myMpiPos = { myRank / (my * mz), (myRank / mz) % my, myRank % mz };
  }
  auto rank_of = [&](int cx, int cy, int cz) {
    cx = (cx + mx) % mx; cy = (cy + my) % my; cz = (cz + mz) % mz;
    return (dimType == TWO_D) ? cx * my + cy : (cx * my + cy) * mz + cz;
  };
  const int cx = myMpiPos[0], cy = myMpiPos[1], cz = myMpiPos[2];
  neighborsRank = { rank_of(cx - 1, cy, cz), rank_of(cx + 1, cy, cz), rank_of(cx, cy - 1, cz),
                    rank_of(cx, cy + 1, cz), rank_of(cx, cy, cz - 1), rank_of(cx, cy, cz + 1) };
  neighborsBC = { BC_COPY, BC_COPY, BC_COPY, BC_COPY, BC_COPY, BC_COPY };
  if (dimType == TWO_D) {
    neighborsBC[FACE_ZMIN] = neighborsBC[FACE_ZMAX] = BC_UNDEFINED;
    neighborsRank[FACE_ZMIN] = neighborsRank[FACE_ZMAX] = 0;
  }
  if (cx == 0) neighborsBC[FACE_XMIN] = boundary_type_xmin;
  if (cx == mx - 1) neighborsBC[FACE_XMAX] = boundary_type_xmax;
  if (cy == 0) neighborsBC[FACE_YMIN] = boundary_type_ymin;
  if (cy == my - 1) neighborsBC[FACE_YMAX] = boundary_type_ymax;
  if (dimType == THREE_D) {
    if (cz == 0) neighborsBC[FACE_ZMIN] = boundary_type_zmin;
    if (cz == mz - 1) neighborsBC[FACE_ZMAX] = boundary_type_zmax;
  }
  // resolution accounts for the number of sub-domains per direction
  dx = (xmax - xmin) / (nx * mx);
  dy = (ymax - ymin) / (ny * my);
  dz = (zmax - zmin) / (nz * mz);
  // the driver's banner (HydroParams.cpp:305-318 of the reference); library users (bench, tests) silence it
  if (myRank == 0 && nProcs > 1 && !std::getenv("EKH_QUIET")) {
    std::cout << "We are about to start simulation with the following characteristics\n";
    std::cout << "Global resolution : " << nx * mx << " x " << ny * my << " x " << nz * mz << "\n";
    std::cout << "Local  resolution : " << nx << " x " << ny << " x " << nz << "\n";
    std::cout << "Cartesian topology (one GPU per sub-domain) : " << mx << "x" << my << "x" << mz << std::endl;
  }
}

// src/shared/HydroParams.cpp:365-412 (same lines, same formats)
void HydroParams::print() const
{
  printf("##########################\n");
  printf("Simulation run parameters:\n");
  printf("##########################\n");
  printf("nx         : %d\n", nx);
  printf("ny         : %d\n", ny);
  printf("nz         : %d\n", nz);
  printf("dx         : %f\n", dx);
  printf("dy         : %f\n", dy);
  printf("dz         : %f\n", dz);
  printf("imin       : %d\n", imin);
  printf("imax       : %d\n", imax);
  printf("jmin       : %d\n", jmin);
  printf("jmax       : %d\n", jmax);
  printf("kmin       : %d\n", kmin);
  printf("kmax       : %d\n", kmax);
  printf("ghostWidth : %d\n", ghostWidth);
  printf("nbvar      : %d\n", nbvar);
  printf("nStepmax   : %d\n", nStepmax);
  printf("tEnd       : %f\n", tEnd);
  printf("nOutput    : %d\n", nOutput);
  printf("gamma0     : %f\n", settings.gamma0);
  printf("gamma6     : %f\n", settings.gamma6);
  printf("cfl        : %f\n", settings.cfl);
  printf("smallr     : %12.10f\n", settings.smallr);
  printf("smallc     : %12.10f\n", settings.smallc);
  printf("smallp     : %12.10f\n", settings.smallp);
  printf("smallpp    : %g\n", settings.smallpp);
  printf("cp (specific heat)          : %g\n", settings.cp);
  printf("mu (dynamic visosity)       : %g\n", settings.mu);
  printf("kappa (thermal diffusivity) : %g\n", settings.kappa);
  printf("iorder     : %d\n", settings.iorder);
  printf("slope_type : %f\n", settings.slope_type);
  printf("riemann    : %d\n", riemannSolverType);
  printf("implementation version : %d\n", implementationVersion);
  printf("##########################\n");
}

ek_params HydroParams::to_ek_params() const
{
  ek_params p;
  std::memset(&p, 0, sizeof(p));
  p.dim = dimType;
  p.mhd = mhdEnabled ? 1 : 0;
  p.nbvar = nbvar;
  p.ghost = ghostWidth;
  p.nx = nx; p.ny = ny; p.nz = (dimType == THREE_D) ? nz : 1;
  p.riemann = riemannSolverType;
  p.impl_version = implementationVersion;
  for (int f = 0; f < 6; ++f) {
    BoundaryConditionType b = neighborsBC[f];
    // a periodic outer face of a decomposed direction is an exchange with the wrap-around neighbour
    // (SolverBase.cpp:583-664: BC_COPY and BC_PERIODIC both unpack the received buffer)
    const int m = (f / 2 == 0) ? mx : (f / 2 == 1) ? my : mz;
    if (b == BC_PERIODIC && m > 1)
      b = BC_COPY;
    p.bc[f] = (dimType == TWO_D && f >= 4) ? EK_BC_UNDEFINED : static_cast<int>(b);
  }
  p.gravity_enabled = gravity.enabled ? 1 : 0;
  p.gravity_hancock = gravity.hancock_predictor_enabled ? 1 : 0;
  p.dx = dx; p.dy = dy; p.dz = (dimType == THREE_D) ? dz : 1.0;
  p.gamma0 = settings.gamma0; p.gamma6 = settings.gamma6;
  p.smallr = settings.smallr; p.smallc = settings.smallc; p.smallp = settings.smallp; p.smallpp = settings.smallpp;
  p.cfl = settings.cfl; p.slope_type = settings.slope_type;
  return p;
}

} // namespace euler_b200
