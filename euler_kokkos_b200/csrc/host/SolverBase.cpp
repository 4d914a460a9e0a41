#include "SolverBase.h"

#include <cmath>

namespace euler_b200 {

// src/shared/SolverBase.cpp:22-67
SolverBase::SolverBase(HydroParams& params_, ConfigMap& configMap_)
  : params(params_)
  , configMap(configMap_)
  , solver_type(SOLVER_UNDEFINED)
  , m_restart_run_enabled(configMap_.getInteger("run", "restart_enabled", 0))
  , m_restart_run_filename(configMap_.getString("run", "restart_filename", ""))
  , m_t(configMap_.getFloat("run", "tCurrent", 0.0))
  , m_dt(0.0)
  , m_iteration(0)
  , m_tEnd(configMap_.getFloat("run", "tEnd", 0.0))
  , m_cfl(configMap_.getFloat("hydro", "cfl", 1.0))
  , m_nlog(configMap_.getInteger("run", "nlog", 10))
  , m_nCells(-1)
  , m_nDofsPerCell(-1)
  , m_problem_name(configMap_.getString("hydro", "problem", "unknown"))
  , m_solver_name(configMap_.getString("run", "solver_name", "unknown"))
  , m_static_gravity_enabled(configMap_.getBool("gravity", "static", false))
  , m_gravity(configMap_)
  , m_times_saved(0)
{
  for (int id : { TIMER_TOTAL, TIMER_IO, TIMER_DT, TIMER_COMPUTE_PRIMITIVES, TIMER_COMPUTE_FLUXES, TIMER_HYDRO_UPDATE,
                  TIMER_BOUNDARIES, TIMER_NUM_SCHEME })
    timers[id] = std::make_shared<Timer>();
  m_variables_names[ID] = "rho";
  m_variables_names[IP] = "energy";
  m_variables_names[IU] = "rho_vx";
  m_variables_names[IV] = "rho_vy";
  m_variables_names[IW] = "rho_vz";
  m_variables_names[IA] = "bx";
  m_variables_names[IB] = "by";
  m_variables_names[IC] = "bz";
}

SolverBase::~SolverBase() = default;

void SolverBase::init_io() {}

void SolverBase::compute_dt()
{
  m_dt = compute_dt_local(); // the cross-rank minimum is taken inside the C ABI (NCCL)
  if (m_t + m_dt > m_tEnd)
    m_dt = m_tEnd - m_t;
}

double SolverBase::compute_dt_local() { return m_tEnd; }

int SolverBase::finished() { return m_t >= (m_tEnd - 1e-14) || m_iteration >= params.nStepmax; }

void SolverBase::next_iteration()
{
  next_iteration_impl();
  ++m_iteration;
  m_t += m_dt;
}

void SolverBase::next_iteration_impl() {}

void SolverBase::save_solution()
{
  save_solution_impl();
  ++m_times_saved;
}

void SolverBase::save_solution_impl() {}

int SolverBase::should_save_solution()
{
  const double interval = m_tEnd / params.nOutput;
  if (params.nOutput < 0)
    return 1;
  if ((m_t - (m_times_saved - 1) * interval) > interval)
    return 1;
  if (std::fabs(m_t - m_tEnd) < 1e-12) // ISFUZZYNULL, THRESHOLD 1e-12 (src/shared/utils.h:20-22)
    return 1;
  return 0;
}

void SolverBase::make_boundaries() {}

int SolverBase::advance(int n)
{
  int done = 0;
  while (done < n && !finished()) {
    next_iteration();
    ++done;
  }
  return done;
}

} // namespace euler_b200
