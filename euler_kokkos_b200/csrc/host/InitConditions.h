// Initial conditions of the problems the benchmark configurations need, evaluated on the host into
// the ghosted SoA state (cold path; uploaded once).  Same formulas, same evaluation order and the
// same float-parsed parameters as the reference's init functors:
//   implode          src/muscl/HydroInitFunctors2D.h:33-125, HydroInitFunctors3D.h:68-168
//   blast            src/muscl/HydroInitFunctors2D.h:127-212, HydroInitFunctors3D.h:173-265
//   rayleigh_taylor  src/muscl/HydroInitFunctors2D.h:665-779, HydroInitFunctors3D.h:728-852
//   orszag_tang      src/muscl/MHDInitFunctors2D.h:345-482, MHDInitFunctors3D.h:376-555
// plus the parameter structs in src/shared/problems/*.h.  Problem dispatch follows
// SolverHydroMuscl<dim>::init (src/muscl/SolverHydroMuscl.cpp:80-207) and SolverMHDMuscl<dim>::init
// (src/muscl/SolverMHDMuscl.h:640-716).  The other problems of the reference (Kelvin-Helmholtz, Gresho,
// four-quadrant, isentropic vortex, Brio-Wu, rotor, field loop, MHD blast/implode/KH) live in
// InitConditionsMore.cpp.
#pragma once
#include <string>

#include "ConfigMap.h"
#include "HydroParams.h"

namespace euler_b200 {

//! U: [nbvar][ksize][jsize][isize] (zero-filled on entry); grav: [dim][ksize][jsize][isize] or null.
//! Returns false when the problem is not available.
bool init_problem(const std::string& problem, const HydroParams& params, const ConfigMap& configMap, double* U, double* grav);

} // namespace euler_b200
