// B200-native MUSCL-Hancock solvers behind the reference's solver names.
//   SolverHydroMuscl<dim>  <->  src/muscl/SolverHydroMuscl.{h,cpp}
//   SolverMHDMuscl<dim>    <->  src/muscl/SolverMHDMuscl.{h,cpp}
// Same time-loop semantics (log cadence, output cadence, where dt is computed, tEnd clamp, U/U2
// ping-pong by iteration parity -- kept inside the C ABI), but all fields live on the GPU behind
// ek_b200.h and one step is one fused launch sequence; there are no Q / flux / slope / trace arrays.
#pragma once
#include <string>
#include <thread>
#include <vector>

#include "../../../include/ek_b200.h"
#include "SolverBase.h"

namespace euler_b200 {

class SolverMusclB200 : public SolverBase {
public:
  SolverMusclB200(HydroParams& params, ConfigMap& configMap);
  ~SolverMusclB200() override;

  double compute_dt_local() override;
  void next_iteration_impl() override;
  void save_solution_impl() override;
  void make_boundaries() override;
  int advance(int n) override;

  //! one godunov_unsplit with the current m_dt
  void godunov_unsplit(real_t dt);
  //! current state, ghosts included, SoA [var][k][j][i] (downloads from the device)
  const std::vector<double>& host_state();
  ek_ctx* context() { return m_ctx; }
  //! block until the dump started by the last save_solution() is on disk (asynchronous output)
  void finish_pending_output();

protected:
  void construct(); // allocation + init + first boundaries + first dt, as the reference's ctors
  void check(int rc, const char* what);
  ek_ctx* m_ctx = nullptr;
  std::vector<double> m_Uhost;
  // asynchronous output ([output] async_io, default yes): the dump is a device snapshot that travels to a page-locked
  // buffer on its own stream and is written by a helper thread while the next iterations run
  double* m_io_buf = nullptr;
  std::thread m_io_thread;
  bool m_async_io = false;
  int isize, jsize, ksize;
};

template <int dim>
class SolverHydroMuscl : public SolverMusclB200 {
public:
  SolverHydroMuscl(HydroParams& params, ConfigMap& configMap)
    : SolverMusclB200(params, configMap)
  {
    construct();
  }
  static SolverBase* create(HydroParams& params, ConfigMap& configMap) { return new SolverHydroMuscl<dim>(params, configMap); }
};

template <int dim>
class SolverMHDMuscl : public SolverMusclB200 {
public:
  SolverMHDMuscl(HydroParams& params, ConfigMap& configMap)
    : SolverMusclB200(params, configMap)
  {
    construct();
  }
  static SolverBase* create(HydroParams& params, ConfigMap& configMap) { return new SolverMHDMuscl<dim>(params, configMap); }
};

} // namespace euler_b200
