// VTK ImageData (.vti) writer, byte-compatible with the reference's default (raw appended) output
// (src/utils/io/IO_VTK.cpp:30-410): interior cells only, one Float64 array per variable.
#pragma once
#include <map>
#include <string>

#include "ConfigMap.h"
#include "HydroParams.h"

namespace euler_b200 {

//! U: ghosted SoA state [var][k][j][i]; writes <outputDir>/<outputPrefix>_%07d.vti
//! (multi-rank runs append _rank%05d to the prefix; the .pvti collection file is not written)
std::string save_vti(const double* U, const HydroParams& params, const ConfigMap& configMap, int nbvar,
                     const std::map<int, std::string>& variables_names, int iStep);

} // namespace euler_b200
