#include "IO_VTK.h"

#include <cstdint>
#include <fstream>
#include <sstream>
#include <vector>

namespace euler_b200 {

std::string save_vti(const double* U, const HydroParams& params, const ConfigMap& configMap, int nbvar,
                     const std::map<int, std::string>& variables_names, int iStep)
{
  const bool three_d = params.dimType == THREE_D;
  const int nx = params.nx, ny = params.ny, nz = three_d ? params.nz : 1;
  const int gw = params.ghostWidth;
  const size_t isize = params.isize, jsize = params.jsize, ksize = three_d ? params.ksize : 1;
  const size_t sv = isize * jsize * ksize;

  std::string outputDir = configMap.getString("output", "outputDir", "./");
  std::string outputPrefix = configMap.getString("output", "outputPrefix", "output");
  if (params.nProcs > 1) {
    char buf[32];
    snprintf(buf, sizeof(buf), "_rank%05d", params.myRank);
    outputPrefix += buf;
  }
  std::ostringstream stepNum;
  stepNum.width(7);
  stepNum.fill('0');
  stepNum << iStep;
  const std::string filename = outputDir + "/" + outputPrefix + "_" + stepNum.str() + ".vti";

  std::ofstream out(filename.c_str(), std::ios_base::out | std::ios_base::binary);
  if (!out)
    return std::string();
  // sub-domain origin (the reference writes the global xmin and relies on .pvti extents)
  const double ox = params.xmin + params.myMpiPos[0] * nx * params.dx;
  const double oy = params.ymin + params.myMpiPos[1] * ny * params.dy;
  const double oz = params.zmin + params.myMpiPos[2] * nz * params.dz;
  out << "<VTKFile type=\"ImageData\" version=\"0.1\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n";
  if (three_d) {
    out << "  <ImageData WholeExtent=\"" << 0 << " " << nx << " " << 0 << " " << ny << " " << 0 << " " << nz << "\" "
        << "Origin=\"" << ox << " " << oy << " " << oz << "\" "
        << "Spacing=\"" << params.dx << " " << params.dy << " " << params.dz << "\">\n";
    out << "  <Piece Extent=\"" << 0 << " " << nx << " " << 0 << " " << ny << " " << 0 << " " << nz << " " << "\">\n";
  } else {
    out << "  <ImageData WholeExtent=\"" << 0 << " " << nx << " " << 0 << " " << ny << " " << 0 << " " << 0 << "\" "
        << "Origin=\"" << ox << " " << oy << " " << 0 << "\" "
        << "Spacing=\"" << params.dx << " " << params.dy << " " << 0 << "\">\n";
    out << "  <Piece Extent=\"" << 0 << " " << nx << " " << 0 << " " << ny << " " << 0 << " " << 0 << " " << "\">\n";
  }
  out << "    <PointData>\n";
  out << "    </PointData>\n";
  out << "    <CellData>\n";
  const uint64_t nbytes = (uint64_t)nx * ny * nz * sizeof(double);
  for (int iVar = 0; iVar < nbvar; ++iVar)
    out << "     <DataArray type=\"Float64\" Name=\"" << variables_names.at(iVar) << "\" format=\"appended\" offset=\""
        << iVar * nbytes + iVar * sizeof(uint64_t) << "\" />\n";
  out << "    </CellData>\n";
  out << "  </Piece>\n";
  out << "  </ImageData>\n";
  out << "  <AppendedData encoding=\"raw\">\n";
  out << "_";
  std::vector<double> row((size_t)nx);
  const int k0 = three_d ? gw : 0;
  for (int iVar = 0; iVar < nbvar; ++iVar) {
    out.write(reinterpret_cast<const char*>(&nbytes), sizeof(uint64_t));
    for (int k = k0; k < k0 + nz; ++k)
      for (int j = gw; j < gw + ny; ++j) {
        const double* src = U + sv * iVar + (size_t)k * isize * jsize + (size_t)j * isize + gw;
        out.write(reinterpret_cast<const char*>(src), (std::streamsize)(nx * sizeof(double)));
      }
  }
  out << "  </AppendedData>\n";
  out << "</VTKFile>\n";
  return filename;
}

} // namespace euler_b200
