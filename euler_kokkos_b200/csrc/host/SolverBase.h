// Abstract solver with the reference's public surface (src/shared/SolverBase.h:53-259): what
// main() and solver_utils touch -- save_solution(), finished(), next_iteration(), timers[], m_t,
// m_dt, m_iteration, m_nCells, m_times_saved, m_variables_names, params -- exists with the same
// meaning.  Ghost fill, halo exchange and all field storage live behind the C ABI (ek_b200.h).
#pragma once
#include <map>
#include <memory>
#include <string>

#include "ConfigMap.h"
#include "HydroParams.h"
#include "Timer.h"
#include "enums.h"

namespace euler_b200 {

enum solver_type_t { SOLVER_UNDEFINED = 0, SOLVER_MUSCL_HANCOCK = 1 };

class SolverBase {
public:
  SolverBase(HydroParams& params, ConfigMap& configMap);
  virtual ~SolverBase();

  HydroParams& params;
  ConfigMap& configMap;
  solver_type_t solver_type;

  bool m_restart_run_enabled;
  std::string m_restart_run_filename;
  real_t m_t;      //!< current time
  real_t m_dt;     //!< current time step
  int m_iteration; //!< current iteration
  real_t m_tEnd;
  real_t m_cfl;
  int m_nlog;
  long long int m_nCells; //!< number of cells INCLUDING ghosts, as in the reference
  long long int m_nDofsPerCell;
  std::string m_problem_name;
  std::string m_solver_name;
  bool m_static_gravity_enabled;
  GravityParams m_gravity;
  int m_times_saved;
  std::map<int, std::string> m_variables_names;
  using TimerMap = std::map<int, std::shared_ptr<Timer>>;
  TimerMap timers;

  virtual void init_io();
  virtual void compute_dt();           // SolverBase.cpp:125-154
  virtual double compute_dt_local();
  virtual int finished();              // SolverBase.cpp:171-176
  virtual void next_iteration();       // SolverBase.cpp:181-195
  virtual void next_iteration_impl();
  virtual void save_solution();
  virtual void save_solution_impl();
  virtual int should_save_solution();  // SolverBase.cpp:240-265
  virtual void make_boundaries();      // ghost fill of the current time level (device)

  //! B200 addition: run up to n iterations without any per-step host synchronisation (device
  //! resident dt / t); returns the number of iterations done.  Default: falls back to next_iteration.
  virtual int advance(int n);
};

} // namespace euler_b200
