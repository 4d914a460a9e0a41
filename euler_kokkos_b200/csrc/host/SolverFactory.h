// String registry of solvers, same surface as the reference (src/shared/SolverFactory.h:24-118):
// Instance(), registerSolver(key, fn) -- the runtime plug-in hook --, create(name, params, configMap)
// which calls init_io() on the new solver and aborts after listing the valid names when the key is
// unknown.  The four reference names are pre-registered with the B200-native solvers
// (SolverFactory.cpp).
#pragma once
#include <map>
#include <string>

#include "SolverBase.h"

namespace euler_b200 {

class SolverFactory {
public:
  using SolverCreateFn = SolverBase* (*)(HydroParams& params, ConfigMap& configMap);

  static SolverFactory& Instance()
  {
    static SolverFactory instance;
    return instance;
  }
  void registerSolver(const std::string& key, SolverCreateFn cfn) { m_solverCreateMap[key] = cfn; }
  bool has(const std::string& key) const { return m_solverCreateMap.count(key) != 0; }
  SolverBase* create(const std::string& solver_name, HydroParams& params, ConfigMap& configMap);

private:
  SolverFactory();
  SolverFactory(const SolverFactory&) = delete;
  SolverFactory& operator=(const SolverFactory&) = delete;
  std::map<std::string, SolverCreateFn> m_solverCreateMap;
};

} // namespace euler_b200
