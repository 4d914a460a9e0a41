// Small extern "C" surface of the C++ host layer, so that bench.py and the Python tests go through
// the SAME ini parsing, parameter derivation, decomposition and initial conditions as euler_b200.
#include <cstring>
#include <fstream>
#include <string>

#include "ConfigMap.h"
#include "HydroParams.h"
#include "IO_VTK.h"
#include "InitConditions.h"
#include "SolverFactory.h"
#include "SolverMuscl.h"

namespace euler_b200 {
int run_ini(const std::string& input_file, const std::string& overrides);

// a solver obtained from the factory together with the objects it refers to (what main() keeps on its stack)
struct HostSolver {
  ConfigMap configMap;
  ParallelEnv env;
  HydroParams* params = nullptr;
  SolverBase* solver = nullptr;
  ~HostSolver()
  {
    delete solver;
    delete params;
  }
};

struct HostCase {
  ConfigMap configMap;
  ParallelEnv env;
  HydroParams* params = nullptr;
  ~HostCase() { delete params; }
};
} // namespace euler_b200

using namespace euler_b200;

extern "C" {

// overrides: "section.name=value" lines.  Returns nullptr if the file cannot be read.
void* ekh_case_create(const char* ini_path, const char* overrides, int rank, int nranks)
{
  HostCase* c = new HostCase();
  c->configMap = ConfigMap(ini_path);
  if (c->configMap.ParseError() < 0) {
    delete c;
    return nullptr;
  }
  if (overrides)
    c->configMap.applyOverrides(overrides);
  c->env.rank = rank;
  c->env.nranks = nranks;
  c->params = new HydroParams(c->configMap, c->env);
  return c;
}
void ekh_case_destroy(void* h) { delete static_cast<HostCase*>(h); }

int ekh_case_params(void* h, ek_params* out)
{
  if (!h || !out) return 1;
  *out = static_cast<HostCase*>(h)->params->to_ek_params();
  return 0;
}
int ekh_case_neighbors(void* h, int32_t* nb6)
{
  if (!h || !nb6) return 1;
  for (int f = 0; f < 6; ++f) nb6[f] = static_cast<HostCase*>(h)->params->neighborsRank[f];
  return 0;
}
// run scalars: out[0]=tEnd, out[1]=nStepmax, out[2]=nOutput, out[3]=nlog, out[4]=xmin, out[5]=ymin, out[6]=zmin,
// out[7..9]=mpi pos, out[10..12]=mx,my,mz
int ekh_case_run_info(void* h, double* out)
{
  if (!h || !out) return 1;
  const HydroParams& p = *static_cast<HostCase*>(h)->params;
  out[0] = p.tEnd; out[1] = p.nStepmax; out[2] = p.nOutput; out[3] = p.nlog; out[4] = p.xmin; out[5] = p.ymin; out[6] = p.zmin;
  out[7] = p.myMpiPos[0]; out[8] = p.myMpiPos[1]; out[9] = p.myMpiPos[2]; out[10] = p.mx; out[11] = p.my; out[12] = p.mz;
  return 0;
}
int ekh_case_get_string(void* h, const char* section, const char* name, char* out, int cap)
{
  if (!h || !out || cap <= 0) return 1;
  const std::string v = static_cast<HostCase*>(h)->configMap.getString(section, name, "");
  std::strncpy(out, v.c_str(), cap - 1);
  out[cap - 1] = 0;
  return 0;
}
double ekh_case_get_float(void* h, const char* section, const char* name, float dflt)
{
  return static_cast<HostCase*>(h)->configMap.getFloat(section, name, dflt);
}
// U (and grav, may be null) must be zero-filled, sized for the ghosted local box
int ekh_case_init_state(void* h, double* U, double* grav)
{
  if (!h || !U) return 1;
  HostCase* c = static_cast<HostCase*>(h);
  const std::string problem = c->configMap.getString("hydro", "problem", "unknown");
  return init_problem(problem, *c->params, c->configMap, U, grav) ? 0 : 2;
}
int ekh_case_write_ini(void* h, const char* path)
{
  if (!h || !path) return 1;
  std::ofstream f(path);
  if (!f) return 2;
  f << static_cast<HostCase*>(h)->configMap.toIni();
  return 0;
}
// write the ghosted state U as dump number `step` exactly as SolverBase::save_solution would (piece + .pvti when decomposed)
int ekh_case_save_vti(void* h, const double* U, int step, char* path_out, int cap)
{
  if (!h || !U) return 1;
  HostCase* c = static_cast<HostCase*>(h);
  const HydroParams& p = *c->params;
  std::map<int, std::string> names;
  const char* hydro2[] = { "rho", "energy", "rho_vx", "rho_vy" };
  const char* all8[] = { "rho", "energy", "rho_vx", "rho_vy", "rho_vz", "bx", "by", "bz" };
  for (int v = 0; v < p.nbvar; ++v)
    names[v] = (p.nbvar == 4) ? hydro2[v] : all8[v]; // SolverBase.cpp:60-67
  const std::string f = save_vti(U, p, c->configMap, p.nbvar, names, step);
  if (f.empty()) return 2;
  if (path_out && cap > 0) {
    strncpy(path_out, f.c_str(), (size_t)cap - 1);
    path_out[cap - 1] = 0;
  }
  return 0;
}
int ekh_solver_registered(const char* name) { return SolverFactory::Instance().has(name) ? 1 : 0; }
// ---- the SolverBase surface main() and solver_utils use (src/main.cpp:112-153), one call per member -------------
// SolverFactory::Instance().create(run/solver_name, params, configMap) on an ini file (+ overrides)
void* ekh_solver_create(const char* ini_path, const char* overrides)
{
  HostSolver* s = new HostSolver();
  s->configMap = ConfigMap(ini_path);
  if (s->configMap.ParseError() < 0) {
    delete s;
    return nullptr;
  }
  if (overrides)
    s->configMap.applyOverrides(overrides);
  s->env = ParallelEnv::fromEnvironment();
  s->params = new HydroParams(s->configMap, s->env);
  const std::string name = s->configMap.getString("run", "solver_name", "Unknown");
  if (!SolverFactory::Instance().has(name)) { // create() would abort the process (SolverFactory.h:105-113)
    delete s;
    return nullptr;
  }
  s->solver = SolverFactory::Instance().create(name, *s->params, s->configMap);
  return s;
}
void ekh_solver_destroy(void* h) { delete static_cast<HostSolver*>(h); }
int ekh_solver_finished(void* h) { return static_cast<HostSolver*>(h)->solver->finished(); }
void ekh_solver_next_iteration(void* h) { static_cast<HostSolver*>(h)->solver->next_iteration(); }
int ekh_solver_advance(void* h, int n) { return static_cast<HostSolver*>(h)->solver->advance(n); }
void ekh_solver_save_solution(void* h)
{
  SolverBase* b = static_cast<HostSolver*>(h)->solver;
  b->save_solution();
  if (SolverMusclB200* m = dynamic_cast<SolverMusclB200*>(b))
    m->finish_pending_output(); // a caller of this C function expects the file to exist when it returns
}
// out[0]=m_t, out[1]=m_dt, out[2]=m_iteration, out[3]=m_times_saved, out[4]=m_nCells, out[5]=state length (doubles)
int ekh_solver_info(void* h, double* out)
{
  if (!h || !out) return 1;
  SolverBase* b = static_cast<HostSolver*>(h)->solver;
  out[0] = b->m_t; out[1] = b->m_dt; out[2] = b->m_iteration; out[3] = b->m_times_saved; out[4] = (double)b->m_nCells;
  SolverMusclB200* m = dynamic_cast<SolverMusclB200*>(b);
  out[5] = m ? (double)ek_state_len(m->context()) : 0.0;
  return 0;
}
// current state, ghosts included, SoA [var][k][j][i]
int ekh_solver_state(void* h, double* U, long long cap)
{
  if (!h || !U) return 1;
  SolverMusclB200* m = dynamic_cast<SolverMusclB200*>(static_cast<HostSolver*>(h)->solver);
  if (!m) return 2;
  const std::vector<double>& u = m->host_state();
  if ((long long)u.size() > cap) return 3;
  std::memcpy(U, u.data(), u.size() * sizeof(double));
  return 0;
}
// the whole driver (main.cpp) as a function
int ekh_run_ini(const char* ini_path, const char* overrides) { return run_ini(ini_path, overrides ? overrides : ""); }

} // extern "C"
