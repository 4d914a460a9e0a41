#include "SolverFactory.h"

#include <cstdio>
#include <cstdlib>

#include "SolverMuscl.h"

namespace euler_b200 {

// src/shared/SolverFactory.cpp:18-22: the same four keys
SolverFactory::SolverFactory()
{
  registerSolver("Hydro_Muscl_2D", &SolverHydroMuscl<2>::create);
  registerSolver("Hydro_Muscl_3D", &SolverHydroMuscl<3>::create);
  registerSolver("MHD_Muscl_2D", &SolverMHDMuscl<2>::create);
  registerSolver("MHD_Muscl_3D", &SolverMHDMuscl<3>::create);
}

SolverBase* SolverFactory::create(const std::string& solver_name, HydroParams& params, ConfigMap& configMap)
{
  auto it = m_solverCreateMap.find(solver_name);
  if (it != m_solverCreateMap.end()) {
    SolverBase* solver = it->second(params, configMap);
    solver->init_io();
    return solver;
  }
  printf("############ WARNING: ############\n");
  printf("%s: is not recognized as a valid application name key.\n", solver_name.c_str());
  printf("Valid solver names are:\n");
  for (const auto& kv : m_solverCreateMap)
    printf("%s\n", kv.first.c_str());
  printf("############ WARNING: ############\n");
  printf("Solver application name not found\n");
  std::abort();
  return nullptr;
}

} // namespace euler_b200
