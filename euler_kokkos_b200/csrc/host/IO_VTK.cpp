#include "IO_VTK.h"

#include <cstdint>
#include <fstream>
#include <sstream>
#include <vector>

namespace euler_b200 {

// .pvti header of a decomposed run: whole extent, one <Piece> per rank (rank -> Cartesian coordinates as
// HydroParams::setup_mpi, row-major like MPI_Cart_create)
void write_pvti_header(const std::string& headerFilename, const std::string& outputPrefix, const HydroParams& params, int nbvar,
                       const std::map<int, std::string>& varNames, int iStep)
{
  const bool three_d = params.dimType == THREE_D;
  const int nx = params.nx, ny = params.ny, nz = three_d ? params.nz : 0;
  const int mx = params.mx, my = params.my, mz = three_d ? params.mz : 0;
  std::ostringstream stepNum;
  stepNum.width(7);
  stepNum.fill('0');
  stepNum << iStep;
  std::ofstream out(headerFilename.c_str(), std::ios_base::out);
  if (!out)
    return;
  out << "<?xml version=\"1.0\"?>\n";
  out << "<VTKFile type=\"PImageData\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n";
  out << "  <PImageData WholeExtent=\"" << 0 << " " << mx * nx << " " << 0 << " " << my * ny << " " << 0 << " " << mz * nz
      << "\" GhostLevel=\"0\" Origin=\"" << params.xmin << " " << params.ymin << " " << params.zmin << "\" Spacing=\"" << params.dx << " "
      << params.dy << " " << (three_d ? params.dz : 0.0) << "\">\n";
  out << "    <PCellData Scalars=\"Scalars_\">\n";
  for (int iVar = 0; iVar < nbvar; ++iVar)
    out << "      <PDataArray type=\"Float64\" Name=\"" << varNames.at(iVar) << "\"/>\n";
  out << "    </PCellData>\n";
  for (int r = 0; r < params.nProcs; ++r) {
    std::ostringstream rankNum;
    rankNum.width(5);
    rankNum.fill('0');
    rankNum << r;
    const std::string piece = outputPrefix + "_time" + stepNum.str() + "_mpi" + rankNum.str() + ".vti";
    int cx, cy, cz = 0;
    if (three_d) {
      cx = r / (my * params.mz);
      cy = (r / params.mz) % my;
      cz = r % params.mz;
    } else {
      cx = r / my;
      cy = r % my;
    }
    // (the reference indents the pieces of a 3-D file by one blank and those of a 2-D file by four, IO_VTK.cpp:951, :986)
    out << (three_d ? " " : "    ") << "<Piece Extent=\"" << cx * nx << " " << cx * nx + nx << " " << cy * ny << " " << cy * ny + ny << " ";
    if (three_d)
      out << cz * nz << " " << cz * nz + nz << " ";
    else
      out << 0 << " " << 1;
    out << "\" Source=\"" << piece << "\"/>\n";
  }
  out << "</PImageData>\n";
  out << "</VTKFile>\n";
}

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
  std::ostringstream stepNum;
  stepNum.width(7);
  stepNum.fill('0');
  stepNum << iStep;
  // decomposed runs: one piece per rank + a .pvti header written by rank 0, with the reference's file names and
  // global cell extents (save_VTK_{2D,3D}_mpi + write_pvti_header, src/utils/io/IO_VTK.cpp:415-1011)
  const bool pieces = params.nProcs > 1;
  std::string filename = outputDir + "/" + outputPrefix + "_" + stepNum.str() + ".vti";
  if (pieces) {
    std::ostringstream rankNum;
    rankNum.width(5);
    rankNum.fill('0');
    rankNum << params.myRank;
    filename = outputDir + "/" + outputPrefix + "_time" + stepNum.str() + "_mpi" + rankNum.str() + ".vti";
    if (params.myRank == 0)
      write_pvti_header(outputDir + "/" + outputPrefix + "_time" + stepNum.str() + ".pvti", outputPrefix, params, nbvar, variables_names, iStep);
  }

  std::ofstream out(filename.c_str(), std::ios_base::out | std::ios_base::binary);
  if (!out)
    return std::string();
  // [output] outputVtkAscii=true: the data arrays as text (operator<< of a double: 6 significant digits), the file then
  // is a well-formed XML document and starts with the XML declaration (IO_VTK.cpp:91-94, 118-152, 289-292, 316-358)
  const bool ascii = configMap.getBool("output", "outputVtkAscii", false);
  if (ascii)
    out << "<?xml version=\"1.0\"?>\n";
  // global cell extents of this piece; the origin is the global one (the pieces are placed by their extents)
  const int x0 = params.myMpiPos[0] * nx, y0 = params.myMpiPos[1] * ny, z0 = three_d ? params.myMpiPos[2] * nz : 0;
  const double ox = params.xmin, oy = params.ymin, oz = params.zmin;
  // the reference writes version 0.1 only from its serial 3-D writer (IO_VTK.cpp:298-302; 2-D :100-104 and the MPI pieces :513-517, :737-741 say 1.0)
  out << "<VTKFile type=\"ImageData\" version=\"" << ((pieces || !three_d) ? "1.0" : "0.1") << "\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n";
  if (three_d) {
    out << "  <ImageData WholeExtent=\"" << x0 << " " << x0 + nx << " " << y0 << " " << y0 + ny << " " << z0 << " " << z0 + nz << "\" "
        << "Origin=\"" << ox << " " << oy << " " << oz << "\" "
        << "Spacing=\"" << params.dx << " " << params.dy << " " << params.dz << "\">\n";
    // (the serial writers leave a blank before the closing quote, the MPI piece writers do not: IO_VTK.cpp:311-312 vs :753-754)
    out << "  <Piece Extent=\"" << x0 << " " << x0 + nx << " " << y0 << " " << y0 + ny << " " << z0 << " " << z0 + nz << (pieces ? "" : " ") << "\">\n";
  } else {
    out << "  <ImageData WholeExtent=\"" << x0 << " " << x0 + nx << " " << y0 << " " << y0 + ny << " " << 0 << " " << 0 << "\" "
        << "Origin=\"" << ox << " " << oy << " " << 0 << "\" "
        << "Spacing=\"" << params.dx << " " << params.dy << " " << 0 << "\">\n";
    out << "  <Piece Extent=\"" << x0 << " " << x0 + nx << " " << y0 << " " << y0 + ny << " " << 0 << " " << 0 << (pieces ? "" : " ") << "\">\n";
  }
  out << "    <PointData>\n";
  out << "    </PointData>\n";
  const int k0 = three_d ? gw : 0;
  if (ascii) {
    out << "    <CellData>\n";
    for (int iVar = 0; iVar < nbvar; ++iVar) {
      out << "    <DataArray type=\"Float64\" Name=\"" << variables_names.at(iVar) << "\" format=\"ascii\" >\n";
      for (int k = k0; k < k0 + nz; ++k)
        for (int j = gw; j < gw + ny; ++j) {
          const double* src = U + sv * iVar + (size_t)k * isize * jsize + (size_t)j * isize + gw;
          for (int i = 0; i < nx; ++i)
            out << src[i] << " ";
        }
      out << "\n    </DataArray>\n";
    }
    out << "    </CellData>\n";
    out << "  </Piece>\n";
    out << "  </ImageData>\n";
    out << "</VTKFile>\n";
    return filename;
  }
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
