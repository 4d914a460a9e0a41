#include "InitConditions.h"
#include "InitGeo.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>

namespace euler_b200 {

namespace {

// ---- implode: ImplodeParams.h, InitImplodeFunctor{2D,3D} ----------------------------------------
void init_implode(const HydroParams& p, const ConfigMap& cm, double* U)
{
  const Geo g(p);
  const double gamma0 = p.settings.gamma0;
  const double rho_out = cm.getFloat("implode", "density_outer", 1.0);
  const double p_out = cm.getFloat("implode", "pressure_outer", 1.0);
  const double u_out = cm.getFloat("implode", "vx_outer", 0.0);
  const double v_out = cm.getFloat("implode", "vy_outer", 0.0);
  const double w_out = cm.getFloat("implode", "vz_outer", 0.0);
  const double rho_in = cm.getFloat("implode", "density_inner", 0.125);
  const double p_in = cm.getFloat("implode", "pressure_inner", 0.14);
  const double u_in = cm.getFloat("implode", "vx_inner", 0.0);
  const double v_in = cm.getFloat("implode", "vy_inner", 0.0);
  const double w_in = cm.getFloat("implode", "vz_inner", 0.0);
  const int shape = cm.getInteger("implode", "shape_region", 0);
  for_each_cell(g, [&](int i, int j, int k) {
    const double x = g.x(i), y = g.y(j);
    const size_t o = g.at(i, j, k);
    bool outer;
    if (g.dim == 2) {
      if (shape == 1)
        outer = x + y * y > 0.5 && x + y * y < 1.5;
      else
        outer = x + y > (g.xmin + g.xmax) / 2. + g.ymin;
      // note: the reference stores the VELOCITY in the momentum slots here (IU = u_out), kept as is
      if (outer) {
        U[o + g.sv * ID] = rho_out;
        U[o + g.sv * IP] = p_out / (gamma0 - 1.0) + 0.5 * rho_out * (u_out * u_out + v_out * v_out);
        U[o + g.sv * IU] = u_out;
        U[o + g.sv * IV] = v_out;
      } else {
        U[o + g.sv * ID] = rho_in;
        U[o + g.sv * IP] = p_in / (gamma0 - 1.0) + 0.5 * rho_in * (u_in * u_in + v_in * v_in);
        U[o + g.sv * IU] = u_in;
        U[o + g.sv * IV] = v_in;
      }
    } else {
      const double z = g.z(k);
      if (shape == 1)
        outer = x + y + z > 0.5 && x + y + z < 2.5;
      else
        outer = x + y + z > (g.xmin + g.xmax) / 2. + g.ymin + g.zmin;
      if (outer) {
        U[o + g.sv * ID] = rho_out;
        U[o + g.sv * IP] = p_out / (gamma0 - 1.0) + 0.5 * rho_out * (u_out * u_out + v_out * v_out + w_out * w_out);
        U[o + g.sv * IU] = u_out;
        U[o + g.sv * IV] = v_out;
        U[o + g.sv * IW] = w_out;
      } else {
        U[o + g.sv * ID] = rho_in;
        U[o + g.sv * IP] = p_in / (gamma0 - 1.0) + 0.5 * rho_in * (u_in * u_in + v_in * v_in + w_in * w_in);
        U[o + g.sv * IU] = u_in;
        U[o + g.sv * IV] = v_in;
        U[o + g.sv * IW] = w_in;
      }
    }
  });
}

// ---- blast: BlastParams.h:44-57, InitBlastFunctor{2D,3D} ------------------------------------------
void init_blast(const HydroParams& p, const ConfigMap& cm, double* U)
{
  const Geo g(p);
  const double gamma0 = p.settings.gamma0;
  // BlastParams re-reads the mesh bounds itself (defaults are doubles narrowed to float)
  const double xmin = cm.getFloat("mesh", "xmin", 0.0), ymin = cm.getFloat("mesh", "ymin", 0.0), zmin = cm.getFloat("mesh", "zmin", 0.0);
  const double xmax = cm.getFloat("mesh", "xmax", 1.0), ymax = cm.getFloat("mesh", "ymax", 1.0), zmax = cm.getFloat("mesh", "zmax", 1.0);
  const double radius = cm.getFloat("blast", "radius", (float)((xmin + xmax) / 2.0 / 10));
  const double cx = cm.getFloat("blast", "center_x", (float)((xmin + xmax) / 2));
  const double cy = cm.getFloat("blast", "center_y", (float)((ymin + ymax) / 2));
  const double cz = cm.getFloat("blast", "center_z", (float)((zmin + zmax) / 2));
  const double d_in = cm.getFloat("blast", "density_in", 1.0);
  const double d_out = cm.getFloat("blast", "density_out", 1.2);
  const double p_in = cm.getFloat("blast", "pressure_in", 10.0);
  const double p_out = cm.getFloat("blast", "pressure_out", 0.1);
  const double radius2 = radius * radius;
  for_each_cell(g, [&](int i, int j, int k) {
    const double x = g.x(i), y = g.y(j);
    double d2;
    if (g.dim == 2) {
      d2 = (x - cx) * (x - cx) + (y - cy) * (y - cy);
    } else {
      const double z = g.z(k);
      d2 = (x - cx) * (x - cx) + (y - cy) * (y - cy) + (z - cz) * (z - cz);
    }
    const size_t o = g.at(i, j, k);
    const bool in = d2 < radius2;
    U[o + g.sv * ID] = in ? d_in : d_out;
    U[o + g.sv * IP] = (in ? p_in : p_out) / (gamma0 - 1.0);
    U[o + g.sv * IU] = 0.0;
    U[o + g.sv * IV] = 0.0;
    if (g.dim == 3)
      U[o + g.sv * IW] = 0.0;
  });
}

// ---- Rayleigh-Taylor: RayleighTaylorInstabilityParams.h, RayleighTaylorInstabilityFunctor{2D,3D} ---
void init_rayleigh_taylor(const HydroParams& p, const ConfigMap& cm, double* U, double* grav)
{
  const Geo g(p);
  const double gamma0 = p.settings.gamma0;
  const double amplitude = cm.getFloat("rayleigh_taylor", "amplitude", 0.01);
  const double d0 = cm.getFloat("rayleigh_taylor", "d0", 1.0);
  const double d1 = cm.getFloat("rayleigh_taylor", "d1", 2.0);
  const double gx = cm.getFloat("rayleigh_taylor", "gx", 0.0);
  const double gy = cm.getFloat("rayleigh_taylor", "gy", -0.1);
  const double gz = cm.getFloat("rayleigh_taylor", "gz", 0.0);
  const double Lx = g.xmax - g.xmin, Ly = g.ymax - g.ymin;
  const double P0 = 1.0;
  for_each_cell(g, [&](int i, int j, int k) {
    const double x = g.x(i), y = g.y(j);
    const size_t o = g.at(i, j, k);
    if (g.dim == 2) {
      const double d = (y > (g.ymin + g.ymax) / 2) ? d1 : d0;
      U[o + g.sv * ID] = d;
      U[o + g.sv * IU] = 0.0;
      U[o + g.sv * IV] = amplitude * d * (1 + cos(2 * M_PI * x / Lx)) * (1 + cos(2 * M_PI * y / Ly)) / 4;
      U[o + g.sv * IE] = (P0 + U[o + g.sv * ID] * (gx * x + gy * y)) / (gamma0 - 1.0);
      if (grav) {
        grav[o + g.sv * IX] = gx;
        grav[o + g.sv * IY] = gy;
      }
    } else {
      const double z = g.z(k);
      const double d = (z > (g.zmin + g.zmax) / 2) ? d1 : d0;
      U[o + g.sv * ID] = d;
      U[o + g.sv * IU] = 0.0;
      U[o + g.sv * IV] = 0.0;
      U[o + g.sv * IW] = amplitude * d * (1 + cos(2 * M_PI * x / Lx)) * (1 + cos(2 * M_PI * y / Ly)) / 4;
      U[o + g.sv * IE] = (P0 + U[o + g.sv * ID] * (gx * x + gy * y + gz * z)) / (gamma0 - 1);
      if (grav) {
        grav[o + g.sv * IX] = gx;
        grav[o + g.sv * IY] = gy;
        grav[o + g.sv * IZ] = gz;
      }
    }
  });
}

// ---- Orszag-Tang: OrszagTangParams.h, InitOrszagTangFunctor{2D,3D} (two passes) --------------------
void init_orszag_tang(const HydroParams& p, const ConfigMap& cm, double* U)
{
  const Geo g(p);
  const double gamma0 = p.settings.gamma0;
  const double PI = 3.141592653589793238462643383279502884L; // PI_v<double> (real_type.h)
  const double TWOPI = 2 * PI;                                // TWOPI_F
  const double kt = cm.getFloat("OrszagTang", "kt", 0.0);
  const int vortex_dir = cm.getInteger("OrszagTang", "vortex_dir", 2); // VortexDir::Z
  const double v0 = 1.0;
  if (g.dim == 3) {
    const double B0 = 1.0 / sqrt(4 * PI);
    const double p0 = gamma0 / (4 * PI);
    const double d0 = gamma0 * p0;
    for_each_cell(g, [&](int i, int j, int k) {
      const double xPos = g.x(i), yPos = g.y(j), zPos = g.z(k);
      const size_t o = g.at(i, j, k);
      U[o + g.sv * ID] = d0;
      if (vortex_dir == 2) {
        U[o + g.sv * IU] = -d0 * v0 * sin(yPos * TWOPI);
        U[o + g.sv * IV] = d0 * v0 * sin(xPos * TWOPI);
        U[o + g.sv * IW] = 0.0;
        U[o + g.sv * IA] = -B0 * cos(2 * TWOPI * kt * (zPos - g.zmin) / (g.zmax - g.zmin)) * sin(yPos * TWOPI);
        U[o + g.sv * IB] = B0 * cos(2 * TWOPI * kt * (zPos - g.zmin) / (g.zmax - g.zmin)) * sin(2.0 * xPos * TWOPI);
        U[o + g.sv * IC] = 0.0;
      } else if (vortex_dir == 0) {
        U[o + g.sv * IV] = -d0 * v0 * sin(zPos * TWOPI);
        U[o + g.sv * IW] = d0 * v0 * sin(yPos * TWOPI);
        U[o + g.sv * IU] = 0.0;
        U[o + g.sv * IB] = -B0 * cos(2 * TWOPI * kt * (xPos - g.xmin) / (g.xmax - g.xmin)) * sin(zPos * TWOPI);
        U[o + g.sv * IC] = B0 * cos(2 * TWOPI * kt * (xPos - g.xmin) / (g.xmax - g.xmin)) * sin(2.0 * yPos * TWOPI);
        U[o + g.sv * IA] = 0.0;
      } else if (vortex_dir == 1) {
        U[o + g.sv * IW] = -d0 * v0 * sin(xPos * TWOPI);
        U[o + g.sv * IU] = d0 * v0 * sin(zPos * TWOPI);
        U[o + g.sv * IV] = 0.0;
        U[o + g.sv * IC] = -B0 * cos(2 * TWOPI * kt * (yPos - g.ymin) / (g.ymax - g.ymin)) * sin(xPos * TWOPI);
        U[o + g.sv * IA] = B0 * cos(2 * TWOPI * kt * (yPos - g.ymin) / (g.ymax - g.ymin)) * sin(2.0 * zPos * TWOPI);
        U[o + g.sv * IB] = 0.0;
      }
    });
    for_each_cell(g, [&](int i, int j, int k) {
      if (i < g.isize - 1 && j < g.jsize - 1 && k < g.ksize - 1) {
        const size_t o = g.at(i, j, k);
        auto sqr = [](double a) { return a * a; };
        const double d = U[o + g.sv * ID];
        U[o + g.sv * IP] = p0 / (gamma0 - 1.0) +
                           0.5 * (sqr(U[o + g.sv * IU]) / d + sqr(U[o + g.sv * IV]) / d + sqr(U[o + g.sv * IW]) / d) +
                           0.5 * (sqr((U[o + g.sv * IA] + U[o + 1 + g.sv * IA]) / 2) + sqr((U[o + g.sv * IB] + U[o + g.sj + g.sv * IB]) / 2) +
                                  sqr((U[o + g.sv * IC] + U[o + g.sk + g.sv * IC]) / 2));
      }
    });
  } else {
    const double B0 = 1.0 / sqrt(2 * TWOPI);
    const double p0 = gamma0 / (2 * TWOPI);
    const double d0 = gamma0 * p0;
    for_each_cell(g, [&](int i, int j, int k) {
      const double xPos = g.x(i), yPos = g.y(j);
      const size_t o = g.at(i, j, k);
      U[o + g.sv * ID] = d0;
      U[o + g.sv * IU] = -d0 * v0 * sin(yPos * TWOPI);
      U[o + g.sv * IV] = d0 * v0 * sin(xPos * TWOPI);
      U[o + g.sv * IW] = 0.0;
      U[o + g.sv * IA] = -B0 * sin(yPos * TWOPI);
      U[o + g.sv * IB] = B0 * sin(2.0 * xPos * TWOPI);
      U[o + g.sv * IC] = 0.0;
    });
    const double TwoPi = 4.0 * asin(1.0);
    const double p0e = gamma0 / (2.0 * TwoPi);
    for_each_cell(g, [&](int i, int j, int k) {
      if (i < g.isize - 1 && j < g.jsize - 1) {
        const size_t o = g.at(i, j, k);
        auto sqr = [](double a) { return a * a; };
        const double d = U[o + g.sv * ID];
        U[o + g.sv * IP] = p0e / (gamma0 - 1.0) + 0.5 * (sqr(U[o + g.sv * IU]) / d + sqr(U[o + g.sv * IV]) / d) +
                           0.5 * (sqr((U[o + g.sv * IA] + U[o + 1 + g.sv * IA]) / 2) + sqr((U[o + g.sv * IB] + U[o + g.sj + g.sv * IB]) / 2));
      }
    });
  }
}

} // namespace

bool init_problem(const std::string& problem, const HydroParams& params, const ConfigMap& configMap, double* U, double* grav)
{
  if (configMap.getBool("run", "restart_enabled", false)) {
    std::cerr << "[euler_b200] restart runs need the HDF5/PnetCDF loaders, which are outside the hot-path scope\n";
    return false;
  }
  if (!params.mhdEnabled) {
    if (problem == "implode") { init_implode(params, configMap, U); return true; }
    if (problem == "blast") { init_blast(params, configMap, U); return true; }
    if (problem == "rayleigh_taylor") { init_rayleigh_taylor(params, configMap, U, grav); return true; }
    if (problem == "kelvin_helmholtz") return init_kelvin_helmholtz(params, configMap, U);
    if (problem == "gresho_vortex") return init_gresho_vortex(params, configMap, U);
    if (problem == "four_quadrant") return init_four_quadrant(params, configMap, U);
    if (problem == "isentropic_vortex") return init_isentropic_vortex(params, configMap, U);
    // SolverHydroMuscl.cpp:131-137: unknown problem -> implode
    std::cout << "Problem : " << problem << " is not recognized / implemented." << std::endl;
    std::cout << "Use default - implode" << std::endl;
    init_implode(params, configMap, U);
    return true;
  }
  if (problem == "orszag_tang") { init_orszag_tang(params, configMap, U); return true; }
  if (problem == "blast") return init_mhd_blast(params, configMap, U);
  if (problem == "implode") return init_mhd_implode(params, configMap, U);
  if (problem == "brio-wu") return init_mhd_brio_wu(params, configMap, U);
  if (problem == "kelvin_helmholtz") return init_mhd_kelvin_helmholtz(params, configMap, U);
  if (problem == "rotor") return init_mhd_rotor(params, configMap, U);
  if (problem == "field_loop" || problem == "field loop") return init_mhd_field_loop(params, configMap, U);
  // SolverMHDMuscl.h:706-713: unknown problem -> Orszag-Tang
  std::cout << "Problem : " << problem << " is not recognized / implemented." << std::endl;
  std::cout << "Use default - Orszag-Tang vortex" << std::endl;
  init_orszag_tang(params, configMap, U);
  return true;
}

} // namespace euler_b200
