#include "SolverMuscl.h"

#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "IO_VTK.h"
#include "InitConditions.h"
#include "comm_rendezvous.h"

namespace euler_b200 {

SolverMusclB200::SolverMusclB200(HydroParams& params_, ConfigMap& configMap_)
  : SolverBase(params_, configMap_)
  , isize(params_.isize)
  , jsize(params_.jsize)
  , ksize(params_.dimType == THREE_D ? params_.ksize : 1)
{
  solver_type = SOLVER_MUSCL_HANCOCK;
  // the reference counts ghost cells here (SolverHydroMuscl.h:182); kept for its Perf line
  m_nCells = (long long)isize * jsize * ksize;
  m_nDofsPerCell = 1;
}

SolverMusclB200::~SolverMusclB200()
{
  finish_pending_output();
  if (m_io_buf)
    ek_free_host(m_io_buf);
  if (m_ctx)
    ek_destroy(m_ctx);
}

void SolverMusclB200::finish_pending_output()
{
  if (m_io_thread.joinable())
    m_io_thread.join();
}

void SolverMusclB200::check(int rc, const char* what)
{
  if (rc != EK_OK) {
    // no error returns in the reference's solver interface: report and stop, as it does
    // (HydroParams.cpp:352-357, SolverFactory.h:105-113)
    fprintf(stderr, "[euler_b200] %s failed (%d): %s\n", what, rc, ek_last_error(m_ctx));
    std::exit(EXIT_FAILURE);
  }
}

// SolverHydroMuscl ctor (SolverHydroMuscl.h:164-307) / SolverMHDMuscl ctor (SolverMHDMuscl.h:265-420)
void SolverMusclB200::construct()
{
  const ek_params p = params.to_ek_params();
  const ParallelEnv env = ParallelEnv::fromEnvironment();
  check(ek_create(&p, env.local_rank, &m_ctx), "ek_create");
  if (params.nProcs > 1) {
    unsigned char id[EK_COMM_ID_BYTES];
    rendezvous_unique_id(params.myRank, params.nProcs, id);
    int32_t nb[6];
    for (int f = 0; f < 6; ++f)
      nb[f] = params.neighborsRank[f];
    check(ek_comm_attach(m_ctx, id, params.myRank, params.nProcs, nb), "ek_comm_attach");
    rendezvous_done(params.myRank);
  }
  const size_t n = (size_t)ek_state_len(m_ctx);
  m_Uhost.assign(n, 0.0);
  m_async_io = configMap.getBool("output", "async_io", true);
  std::vector<double> grav;
  if (params.gravity.enabled)
    grav.assign((size_t)isize * jsize * ksize * params.dimType, 0.0);
  // initial condition (host; cold path), then upload: sets U and U2 (deep_copy(U2,U))
  if (!init_problem(m_problem_name, params, configMap, m_Uhost.data(), grav.empty() ? nullptr : grav.data())) {
    // restart_enabled=1 (HDF5/PnetCDF loaders are not part of this build) or an initial condition that does not
    // exist for this solver: stop, as the reference does for configurations it cannot run (HydroParams.cpp:352-357)
    fprintf(stderr, "[euler_b200] cannot initialise problem '%s' for solver %s\n", m_problem_name.c_str(), m_solver_name.c_str());
    std::exit(EXIT_FAILURE);
  }
  check(ek_upload_state(m_ctx, m_Uhost.data()), "ek_upload_state");
  if (!grav.empty())
    check(ek_set_gravity(m_ctx, grav.data()), "ek_set_gravity");
  // initialize boundaries, then the first time step
  check(ek_fill_boundaries(m_ctx), "ek_fill_boundaries");
  compute_dt();

  if (params.myRank == 0) {
    std::cout << "##########################" << "\n";
    std::cout << "Solver is " << m_solver_name << "\n";
    std::cout << "Problem (init condition) is " << m_problem_name << "\n";
    std::cout << "##########################" << "\n";
    params.print();
    std::cout << "##########################" << "\n";
    std::cout << "Memory requested : " << (2.0 * n * sizeof(double) / 1e6) << " MBytes (device: U + U2; no Q/flux/slope arrays)\n";
    std::cout << "##########################" << "\n";
  }
}

// SolverHydroMuscl::compute_dt_local (SolverHydroMuscl.h:591-631), SolverMHDMuscl (:718-737)
double SolverMusclB200::compute_dt_local()
{
  double dt = 0.0;
  timers[TIMER_DT]->start();
  check(ek_compute_dt(m_ctx, &dt), "ek_compute_dt");
  timers[TIMER_DT]->stop();
  return dt;
}

void SolverMusclB200::make_boundaries()
{
  timers[TIMER_BOUNDARIES]->start();
  check(ek_fill_boundaries(m_ctx), "ek_fill_boundaries");
  timers[TIMER_BOUNDARIES]->stop();
}

// SolverHydroMuscl::next_iteration_impl (SolverHydroMuscl.h:637-679);
// SolverMHDMuscl::next_iteration_impl (SolverMHDMuscl.h:743-783) computes dt inside the step, after
// the ghost fill -- ek_compute_dt does that ghost fill for MHD, so the order of operations on the
// data is the same.
void SolverMusclB200::next_iteration_impl()
{
  if (m_iteration % m_nlog == 0 && params.myRank == 0)
    printf("time step=%7d (dt=% 10.8f t=% 10.8f)\n", m_iteration, m_dt, m_t);
  if (params.enableOutput && should_save_solution()) {
    if (params.myRank == 0)
      std::cout << "Output results at time t=" << m_t << " step " << m_iteration << " dt=" << m_dt << std::endl;
    save_solution();
  }
  compute_dt();
  godunov_unsplit(m_dt);
}

void SolverMusclB200::godunov_unsplit(real_t dt)
{
  timers[TIMER_NUM_SCHEME]->start();
  check(ek_step(m_ctx, dt), "ek_step");
  timers[TIMER_NUM_SCHEME]->stop();
}

// Sync-free variant of the while(!finished()) next_iteration() loop for stretches without output:
// the log line is printed at the start of the stretch only.
int SolverMusclB200::advance(int n)
{
  if (n <= 0 || finished())
    return 0;
  const int left = params.nStepmax - m_iteration;
  if (n > left)
    n = left;
  if (m_iteration % m_nlog == 0 && params.myRank == 0)
    printf("time step=%7d (dt=% 10.8f t=% 10.8f)\n", m_iteration, m_dt, m_t);
  double t_out = m_t, dt_last = m_dt;
  int done = 0;
  timers[TIMER_NUM_SCHEME]->start();
  check(ek_run(m_ctx, n, m_t, m_tEnd, &t_out, &dt_last, &done), "ek_run");
  timers[TIMER_NUM_SCHEME]->stop();
  m_t = t_out;
  m_dt = dt_last;
  m_iteration += done;
  return done;
}

const std::vector<double>& SolverMusclB200::host_state()
{
  check(ek_download_state(m_ctx, m_Uhost.data()), "ek_download_state");
  return m_Uhost;
}

// SolverHydroMuscl::save_solution_impl (SolverHydroMuscl.h:751-762) -> save_VTK_{2D,3D}
void SolverMusclB200::save_solution_impl()
{
  // [output] vtk_enabled=no: IO_ReadWrite::save_data writes nothing then (IO_ReadWrite.cpp:37, 85-92; the HDF5 / PnetCDF
  // writers it would call instead are not part of this build); the "Output results" line and the dump counter of
  // SolverBase::save_solution stay
  if (!configMap.getBool("output", "vtk_enabled", true))
    return;
  timers[TIMER_IO]->start();
  finish_pending_output(); // at most one dump in flight: its buffer and the device snapshot are reused
  if (m_async_io && !m_io_buf) {
    void* p = nullptr;
    if (ek_alloc_host(m_Uhost.size() * sizeof(double), &p) == EK_OK)
      m_io_buf = static_cast<double*>(p);
    else
      m_async_io = false; // no page-locked memory: synchronous dumps
  }
  if (m_async_io) {
    // SolverHydroMuscl::save_solution_impl copies, then writes; here the copy is a device snapshot taken NOW (ordered
    // on the stream, so it is the state of this iteration) and the write happens beside the next iterations
    check(ek_download_state_async(m_ctx, m_io_buf), "ek_download_state_async");
    const int step = m_times_saved;
    m_io_thread = std::thread([this, step] {
      if (ek_download_wait(m_ctx) == EK_OK)
        save_vti(m_io_buf, params, configMap, params.nbvar, m_variables_names, step);
    });
  } else {
    host_state();
    save_vti(m_Uhost.data(), params, configMap, params.nbvar, m_variables_names, m_times_saved);
  }
  timers[TIMER_IO]->stop();
}

template class SolverHydroMuscl<2>;
template class SolverHydroMuscl<3>;
template class SolverMHDMuscl<2>;
template class SolverMHDMuscl<3>;

} // namespace euler_b200
