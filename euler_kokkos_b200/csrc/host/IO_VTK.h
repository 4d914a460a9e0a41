// VTK ImageData (.vti) writer, byte-compatible with the reference's default (raw appended) output
// (src/utils/io/IO_VTK.cpp:30-410): interior cells only, one Float64 array per variable.
#pragma once
#include <map>
#include <string>

#include "ConfigMap.h"
#include "HydroParams.h"

namespace euler_b200 {

//! U: ghosted SoA state [var][k][j][i]; writes <outputDir>/<outputPrefix>_%07d.vti.  Decomposed runs write one piece
//! per rank, <outputPrefix>_time%07d_mpi%05d.vti with global cell extents, and rank 0 adds the
//! <outputPrefix>_time%07d.pvti header (the reference's save_VTK_*_mpi / write_pvti_header, IO_VTK.cpp:415-1011).
void write_pvti_header(const std::string& headerFilename, const std::string& outputPrefix, const HydroParams& params, int nbvar,
                       const std::map<int, std::string>& varNames, int iStep);
std::string save_vti(const double* U, const HydroParams& params, const ConfigMap& configMap, int nbvar,
                     const std::map<int, std::string>& variables_names, int iStep);

} // namespace euler_b200
