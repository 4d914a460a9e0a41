// euler_b200 <file.ini> -- same command line and the same driver loop as the reference's
// src/main.cpp:54-159: read the ini, build HydroParams, ask the factory for run/solver_name, loop
// next_iteration() until finished(), print the monitoring summary.
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "ConfigMap.h"
#include "HydroParams.h"
#include "SolverFactory.h"
#include "solver_utils.h"

namespace euler_b200 {

int run_ini(const std::string& input_file, const std::string& overrides)
{
  ParallelEnv par_env = ParallelEnv::fromEnvironment();
  ConfigMap configMap = broadcast_parameters(input_file);
  if (configMap.ParseError() < 0)
    return EXIT_FAILURE;
  configMap.applyOverrides(overrides);
  HydroParams params(configMap, par_env);
  const std::string solver_name = configMap.getString("run", "solver_name", "Unknown");
  SolverBase* solver = SolverFactory::Instance().create(solver_name, params, configMap);

  if (params.nOutput != 0)
    solver->save_solution();
  if (par_env.rank == 0)
    std::cout << "Start computation....\n";
  solver->timers[TIMER_TOTAL]->start();
  // B200 fast path: without output there is nothing the host needs per step, so whole stretches of
  // nlog iterations run device-resident ([run] sync_free=no restores the per-step host loop)
  const bool sync_free = !params.enableOutput && configMap.getBool("run", "sync_free", true);
  while (!solver->finished()) {
    if (sync_free) {
      const int nlog = solver->m_nlog > 0 ? solver->m_nlog : 10;
      solver->advance(nlog - (solver->m_iteration % nlog));
    } else {
      solver->next_iteration();
    }
  }
  solver->timers[TIMER_TOTAL]->stop();
  if (params.nOutput != 0)
    solver->save_solution();
  if (par_env.rank == 0)
    printf("final time is %f\n", solver->m_t);
  print_solver_monitoring_info(solver);
  delete solver;
  return EXIT_SUCCESS;
}

} // namespace euler_b200

#ifndef EKH_NO_MAIN
int main(int argc, char* argv[])
{
  if (argc < 2) {
    fprintf(stderr, "usage: %s <parameter_file.ini>\n", argv[0]);
    return EXIT_FAILURE;
  }
  return euler_b200::run_ini(argv[1], "");
}
#endif
