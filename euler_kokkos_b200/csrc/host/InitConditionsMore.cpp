// The remaining initial conditions of the reference, evaluated on the host with the same
// formulas, evaluation order and float-parsed parameters, so that dump 0 of the reference binary is reproduced
// bit for bit (tests/test_oracle_golden.py::test_initial_conditions_*):
//   hydro  kelvin_helmholtz   src/muscl/HydroInitFunctors2D.h:214-348, HydroInitFunctors3D.h:270-426, shared/problems/KHParams.h
//          gresho_vortex      HydroInitFunctors2D.h:350-449, HydroInitFunctors3D.h:428-539, GreshoParams.h
//          four_quadrant      HydroInitFunctors2D.h:451-565, HydroInitFunctors3D.h:541-726, initRiemannConfig2d.{h,cpp}
//          isentropic_vortex  HydroInitFunctors2D.h:567-663 (2-D only), IsentropicVortexParams.h
//   MHD    implode            src/muscl/MHDInitFunctors2D.h:37-143, MHDInitFunctors3D.h:34-153, ImplodeParams.h
//          brio-wu            MHDInitFunctors2D.h:145-246, MHDInitFunctors3D.h:155-265, BrioWuParams.h
//          blast              MHDInitFunctors2D.h:248-343, MHDInitFunctors3D.h:267-374, BlastParams.h
//          kelvin_helmholtz   MHDInitFunctors2D.h:484-676, MHDInitFunctors3D.h:560-765
//          rotor              MHDInitFunctors2D.h:678-803, MHDInitFunctors3D.h:767-897, RotorParams.h
//          field_loop         MHDInitFunctors2D.h:805-1036, MHDInitFunctors3D.h:899-1164, FieldLoopParams.h
// KH "perturbation_rand": the reference draws from Kokkos::Random_XorShift64_Pool inside an MDRangePolicy loop, so its
// stream depends on the back-end, the thread count and the iteration-to-thread map.  The stream of its OpenMP back-end
// is deterministic for a given thread count and is reproduced here bit for bit (KokkosPoolStream below: the pool's
// seeding, the xorshift64* generator, the host tiling of MDRangePolicy and the schedule(static,1) tile-to-thread map;
// tests/test_oracle_golden.py::test_kh_random_perturbation_matches_the_reference_stream against dumps of the reference
// binary run with 1, 3 and 4 threads).  Thread count: $EK_KH_RAND_THREADS, else $OMP_NUM_THREADS (what the reference
// binary obeys), else 1; per-rank seed = rand_seed * (rank + 1) as KHParams.h:60-70.  "[KH] rand_generator=hash" selects
// the counter-based generator instead (keyed on (seed, global cell index): identical for every decomposition and
// thread count, which the reference's is not).  The stream of the reference's Cuda back-end (state picked by GPU
// thread / block index behind atomic locks) is not reproducible and is not attempted.
// Known, documented departure:
//   * MHD blast: the reference never assigns blast_bz_in / blast_bz_out (BlastParams.h:56-62 assigns by_* twice and
//     leaves bz_* uninitialised); Bz = 0 here.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "InitGeo.h"
#include "RiemannConfigTable.h"

namespace euler_b200 {

namespace {

inline double sqr(double a) { return a * a; }

// uniform double in [0,1) from a 64-bit key (splitmix64 finaliser)
inline double hashed_uniform(uint64_t key)
{
  uint64_t z = key + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// global (decomposition-independent) index of a ghosted cell; ghosts of neighbouring sub-domains alias interior cells
inline uint64_t global_cell(const Geo& g, const HydroParams& p, int i, int j, int k)
{
  const int64_t gx = (int64_t)p.nx * p.mx, gy = (int64_t)p.ny * p.my, gz = (g.dim == 3) ? (int64_t)p.nz * p.mz : 1;
  auto wrap = [](int64_t a, int64_t n) { return ((a % n) + n) % n; };
  const int64_t I = wrap((int64_t)i + (int64_t)g.nx * g.i_mpi - g.gw, gx);
  const int64_t J = wrap((int64_t)j + (int64_t)g.ny * g.j_mpi - g.gw, gy);
  const int64_t K = (g.dim == 3) ? wrap((int64_t)k + (int64_t)g.nz * g.k_mpi - g.gw, gz) : 0;
  return (uint64_t)(I + gx * (J + gy * K));
}

// ---- Kokkos::Random_XorShift64_Pool as consumed by a host MDRangePolicy loop ------------------------------------
// external/kokkos/algorithms/src/Kokkos_Random.hpp:750-1000 (generator :769-785, :859; pool seeding :958-973; state of a
// thread = pool entry omp_get_thread_num(), :621-630), external/kokkos/core/src/KokkosExp_MDRangePolicy.hpp:140-175,
// 383-450 (host tiles: 2 in every rank but the innermost of the iteration direction, which takes its full range;
// Iterate::Right on host back-ends), impl/KokkosExp_Host_IterateTile.hpp:1567-1592 (tile index -> offsets, last rank
// fastest), OpenMP/Kokkos_OpenMP_Parallel_For.hpp:181-186 (schedule(static, 1) over the tiles).
struct KokkosXorShift64 {
  uint64_t s;
  explicit KokkosXorShift64(uint64_t state) : s(state == 0 ? 1318319ull : state) {}
  void step()
  {
    s ^= s >> 12;
    s ^= s << 25;
    s ^= s >> 27;
  }
  uint32_t urand()
  {
    step();
    return (uint32_t)(((s * 2685821657736338717ull) >> 16) & 0xffffffffull);
  }
  int rand() { return (int)(urand() / 2); }
  uint64_t urand64()
  {
    step();
    return s * 2685821657736338717ull - 1;
  }
  double drand() { return urand64() / (double)UINT64_MAX; }
};

// ndraw uniform numbers per cell of the ghosted box, in the order the reference's OpenMP run with `threads` threads
// draws them: value d of cell (i,j,k) at r[ndraw * g.at(i,j,k) + d]
struct KokkosPoolStream {
  std::vector<double> r;
  int ndraw;
  KokkosPoolStream(const Geo& g, uint64_t seed, int threads, int ndraw_) : r((size_t)ndraw_ * g.sv), ndraw(ndraw_)
  {
    if (seed == 0)
      seed = 1318319ull;
    if (threads < 1)
      threads = 1;
    std::vector<KokkosXorShift64> state;
    {
      KokkosXorShift64 gen(seed);
      for (int i = 0; i < 17; ++i)
        gen.rand();
      for (int t = 0; t < threads; ++t) {
        const uint64_t n1 = (uint64_t)gen.rand(), n2 = (uint64_t)gen.rand(), n3 = (uint64_t)gen.rand(), n4 = (uint64_t)gen.rand();
        state.emplace_back(((n1 & 0xffff) << 0) | ((n2 & 0xffff) << 16) | ((n3 & 0xffff) << 32) | ((n4 & 0xffff) << 48));
      }
    }
    // tiles: 2 x jsize (2-D), 2 x 2 x ksize (3-D); tile index = ti (2-D), ti * ntj + tj (3-D); tile w -> thread w % T,
    // every thread takes its tiles in increasing order; inside a tile the last index runs fastest
    const int nti = (g.isize + 1) / 2;
    const int ntj = (g.dim == 3) ? (g.jsize + 1) / 2 : 1;
    const int tj = (g.dim == 3) ? 2 : g.jsize;
    for (int w = 0; w < nti * ntj; ++w) {
      KokkosXorShift64& gen = state[(size_t)(w % threads)];
      const int i0 = (w / ntj) * 2, j0 = (w % ntj) * tj;
      for (int i = i0; i < i0 + 2 && i < g.isize; ++i)
        for (int j = j0; j < j0 + tj && j < g.jsize; ++j)
          for (int k = 0; k < g.ksize; ++k) {
            double* o = r.data() + (size_t)ndraw * g.at(i, j, k);
            for (int d = 0; d < ndraw; ++d)
              o[d] = gen.drand(); // get_state / free_state around every cell hand the same state back
          }
    }
  }
  double operator()(const Geo& g, int i, int j, int k, int d) const { return r[(size_t)ndraw * g.at(i, j, k) + d]; }
};

inline int kh_rand_threads()
{
  for (const char* name : { "EK_KH_RAND_THREADS", "OMP_NUM_THREADS" }) {
    const char* v = std::getenv(name);
    if (v && std::atoi(v) > 0)
      return std::atoi(v);
  }
  return 1;
}

struct KH {
  double d_in, d_out, pressure, vflow_in, vflow_out, amplitude, w0 = 0.0, delta = 0.0;
  bool p_sine, p_sine_rob, p_rand;
  bool hash_rng = false; // [KH] rand_generator=hash: the counter-based generator instead of the reference's stream
  long seed = 0;
  int mode = 0;
  explicit KH(const ConfigMap& cm)
  {
    d_in = cm.getFloat("KH", "d_in", 1.0);
    d_out = cm.getFloat("KH", "d_out", 1.0);
    pressure = cm.getFloat("KH", "pressure", 10.0);
    p_sine = cm.getBool("KH", "perturbation_sine", false);
    p_sine_rob = cm.getBool("KH", "perturbation_sine_robertson", true);
    p_rand = cm.getBool("KH", "perturbation_rand", false);
    vflow_in = cm.getFloat("KH", "vflow_in", -0.5);
    vflow_out = cm.getFloat("KH", "vflow_out", 0.5);
    if (p_rand) {
      seed = cm.getInteger("KH", "rand_seed", 12);
      hash_rng = cm.getString("KH", "rand_generator", "kokkos") == "hash";
    }
    amplitude = cm.getFloat("KH", "amplitude", 0.1);
    if (p_sine_rob || p_sine) {
      mode = (int)cm.getInteger("KH", "mode", 2);
      w0 = cm.getFloat("KH", "w0", 0.1);
      delta = cm.getFloat("KH", "delta", 0.03);
    }
  }
};

// the ndraw uniform numbers of a cell: the reference's OpenMP stream (default) or the counter-based generator
struct KHRandom {
  const Geo& g;
  const HydroParams& p;
  long seed;
  std::vector<KokkosPoolStream> ks; // empty: hashed generator
  KHRandom(const Geo& g_, const HydroParams& p_, const KH& kh, int ndraw) : g(g_), p(p_), seed(kh.seed)
  {
    if (kh.p_rand && !kh.hash_rng)
      ks.emplace_back(g, (uint64_t)(kh.seed * (long)(p.myRank + 1)), kh_rand_threads(), ndraw);
  }
  double operator()(int i, int j, int k, int d) const
  {
    if (!ks.empty())
      return ks[0](g, i, j, k, d);
    const uint64_t key = ((uint64_t)seed << 40) ^ (global_cell(g, p, i, j, k) * 4);
    return hashed_uniform(key + (uint64_t)d);
  }
};

} // namespace

// ---- hydro ---------------------------------------------------------------------------------------
bool init_kelvin_helmholtz(const HydroParams& p, const ConfigMap& cm, double* U)
{
  const Geo g(p);
  const KH kh(cm);
  const double gamma0 = p.settings.gamma0;
  const double ampl = kh.amplitude, pressure = kh.pressure;
  const KHRandom rnd(g, p, kh, g.dim);
  for_each_cell(g, [&](int i, int j, int k) {
    const double x = g.x(i), y = g.y(j);
    const size_t o = g.at(i, j, k);
    if (g.dim == 2) {
      const double yn = (y - g.ymin) / (g.ymax - g.ymin);
      if (kh.p_rand) {
        double d, u, v;
        if (yn < 0.25 || yn > 0.75) { d = kh.d_out; u = kh.vflow_out; v = 0.0; } else { d = kh.d_in; u = kh.vflow_in; v = 0.0; }
        u += ampl * (rnd(i, j, k, 0) - 0.5);
        v += ampl * (rnd(i, j, k, 1) - 0.5);
        U[o + g.sv * ID] = d;
        U[o + g.sv * IU] = d * u;
        U[o + g.sv * IV] = d * v;
        U[o + g.sv * IE] = pressure / (gamma0 - 1.0) + 0.5 * d * (u * u + v * v);
      } else if (kh.p_sine_rob) {
        const int n = kh.mode;
        const double y1 = 0.25, y2 = 0.75, rho1 = kh.d_in, rho2 = kh.d_out, v1 = kh.vflow_in, v2 = kh.vflow_out;
        const double ramp = 1.0 / (1.0 + exp(2 * (y - y1) / kh.delta)) + 1.0 / (1.0 + exp(2 * (y2 - y) / kh.delta));
        const double d = rho1 + ramp * (rho2 - rho1);
        const double u = v1 + ramp * (v2 - v1);
        const double v = kh.w0 * sin(n * M_PI * x);
        U[o + g.sv * ID] = d;
        U[o + g.sv * IU] = d * u;
        U[o + g.sv * IV] = d * v;
        U[o + g.sv * IP] = pressure / (gamma0 - 1.0) + 0.5 * d * (u * u + v * v);
      } // perturbation_sine alone: the hydro functor writes nothing (the state stays zero), HydroInitFunctors2D.h:266-338
    } else {
      const double z = g.z(k);
      const double zn = (z - g.zmin) / (g.zmax - g.zmin);
      if (kh.p_rand) {
        double d, u, v = 0.0, w = 0.0;
        if (zn < 0.25 || zn > 0.75) { d = kh.d_out; u = kh.vflow_out; } else { d = kh.d_in; u = kh.vflow_in; }
        u += ampl * (rnd(i, j, k, 0) - 0.5);
        v += ampl * (rnd(i, j, k, 1) - 0.5);
        w += ampl * (rnd(i, j, k, 2) - 0.5);
        U[o + g.sv * ID] = d;
        U[o + g.sv * IU] = d * u;
        U[o + g.sv * IV] = d * v;
        U[o + g.sv * IW] = d * w;
        U[o + g.sv * IP] = pressure / (gamma0 - 1.0) + 0.5 * d * (u * u + v * v + w * w);
      } else if (kh.p_sine_rob) {
        const int n = kh.mode;
        const double z1 = 0.25, z2 = 0.75, rho1 = kh.d_in, rho2 = kh.d_out;
        const double v1x = kh.vflow_in, v2x = kh.vflow_out, v1y = kh.vflow_in / 2, v2y = kh.vflow_out / 2;
        const double ramp = 1.0 / (1.0 + exp(2 * (z - z1) / kh.delta)) + 1.0 / (1.0 + exp(2 * (z2 - z) / kh.delta));
        const double d = rho1 + ramp * (rho2 - rho1);
        const double u = v1x + ramp * (v2x - v1x);
        const double v = v1y + ramp * (v2y - v1y);
        const double w = kh.w0 * sin(n * M_PI * x) * sin(n * M_PI * y);
        U[o + g.sv * ID] = d;
        U[o + g.sv * IU] = d * u;
        U[o + g.sv * IV] = d * v;
        U[o + g.sv * IW] = d * w;
        U[o + g.sv * IP] = pressure / (gamma0 - 1.0) + 0.5 * d * (u * u + v * v + w * w);
      }
    }
  });
  return true;
}

bool init_gresho_vortex(const HydroParams& p, const ConfigMap& cm, double* U)
{
  const Geo g(p);
  const double gamma0 = p.settings.gamma0;
  const double rho0 = cm.getFloat("Gresho", "rho0", 1.0);
  const double Ma = cm.getFloat("Gresho", "Ma", 0.1);
  const double p0 = rho0 / (gamma0 * Ma * Ma);
  for_each_cell(g, [&](int i, int j, int k) {
    const double x = g.x(i), y = g.y(j);
    const size_t o = g.at(i, j, k);
    const double r = sqrt(x * x + y * y);
    const double theta = atan2(y, x);
    const double cosT = cos(theta), sinT = sin(theta);
    double uphi, pr;
    if (r < 0.2) {
      uphi = 5 * r;
      pr = p0 + 25 / 2.0 * r * r;
    } else if (r < 0.4) {
      uphi = 2 - 5 * r;
      pr = p0 + 25 / 2.0 * r * r + 4 * (1 - 5 * r - log(0.2) + log(r));
    } else {
      uphi = 0;
      pr = p0 - 2 + 4 * log(2.0);
    }
    double* d = U + o + g.sv * ID;
    double* mu = U + o + g.sv * IU;
    double* mv = U + o + g.sv * IV;
    *d = rho0;
    *mu = rho0 * (-sinT * uphi);
    *mv = rho0 * (cosT * uphi);
    if (g.dim == 2) {
      U[o + g.sv * IP] = pr / (gamma0 - 1.0) + 0.5 * (*mu * *mu + *mv * *mv) / *d;
    } else {
      double* mw = U + o + g.sv * IW;
      *mw = rho0 * (cosT * uphi); // as the reference (HydroInitFunctors3D.h:524)
      U[o + g.sv * IP] = pr / (gamma0 - 1.0) + 0.5 * (*mu * *mu + *mv * *mv + *mw * *mw) / *d;
    }
  });
  return true;
}

bool init_four_quadrant(const HydroParams& p, const ConfigMap& cm, double* U)
{
  const Geo g(p);
  const double gamma0 = p.settings.gamma0;
  const int cfg = (int)cm.getInteger("riemann2d", "config_number", 0);
  const double xt = cm.getFloat("riemann2d", "x", 0.8);
  const double yt = cm.getFloat("riemann2d", "y", 0.8);
  const double zt = cm.getFloat("riemann2d", "z", 0.8);
  double Us[8][5] = {}; // conservative states, index = variable (ID, IP, IU, IV[, IW])
  if (g.dim == 2) {
    if (cfg < 0 || cfg >= kRiemann2dConfigs) {
      std::cerr << "[euler_b200] riemann2d/config_number " << cfg << " is not one of the tabulated configurations\n";
      return false;
    }
    for (int q = 0; q < 4; ++q) {
      const double rho = kRiemann2d[cfg][q][0], u = kRiemann2d[cfg][q][1], v = kRiemann2d[cfg][q][2], pr = kRiemann2d[cfg][q][3];
      Us[q][ID] = rho;
      Us[q][IU] = u * rho;
      Us[q][IV] = v * rho;
      Us[q][IP] = pr / (gamma0 - 1.0) + rho * (u * u + v * v) * 0.5;
    }
  } else {
    int idx = -1;
    for (size_t n = 0; n < sizeof(kRiemann3dIds) / sizeof(int); ++n)
      if (kRiemann3dIds[n] == cfg)
        idx = (int)n;
    if (idx < 0) {
      std::cerr << "[euler_b200] the reference defines no 3-D four-quadrant configuration " << cfg << " (initRiemannConfig2d.cpp:490-655)\n";
      return false;
    }
    for (int q = 0; q < 8; ++q) {
      const double* s = kRiemann3d[idx][q];
      const double rho = s[0], u = s[1], v = s[2], w = s[3], pr = s[4];
      Us[q][ID] = rho;
      Us[q][IU] = u * rho;
      Us[q][IV] = v * rho;
      Us[q][IW] = w * rho;
      Us[q][IP] = pr / (gamma0 - 1.0) + rho * (u * u + v * v + w * w) * 0.5;
    }
  }
  const int nv = (g.dim == 2) ? 4 : 5;
  for_each_cell(g, [&](int i, int j, int k) {
    const double x = g.x(i), y = g.y(j);
    int q = (x < xt) ? ((y < yt) ? 2 : 1) : ((y < yt) ? 3 : 0);
    if (g.dim == 3 && !(g.z(k) < zt))
      q += 4;
    const size_t o = g.at(i, j, k);
    for (int v = 0; v < nv; ++v)
      U[o + g.sv * v] = Us[q][v];
  });
  return true;
}

bool init_isentropic_vortex(const HydroParams& p, const ConfigMap& cm, double* U)
{
  const Geo g(p);
  if (g.dim != 2) {
    std::cerr << "You shouldn't be here: isentropic vortex is not implemented in 3D !\n"; // SolverHydroMuscl.h:471-477
    return false;
  }
  const double gamma0 = p.settings.gamma0;
  const double xmin = cm.getFloat("mesh", "xmin", 0.0), ymin = cm.getFloat("mesh", "ymin", 0.0);
  const double xmax = cm.getFloat("mesh", "xmax", 1.0), ymax = cm.getFloat("mesh", "ymax", 1.0);
  const double rho_a = cm.getFloat("isentropic_vortex", "density_ambient", 1.0);
  const double T_a = cm.getFloat("isentropic_vortex", "temperature_ambient", 1.0);
  const double u_a = cm.getFloat("isentropic_vortex", "vx_ambient", 1.0);
  const double v_a = cm.getFloat("isentropic_vortex", "vy_ambient", 1.0);
  const double vortex_x = cm.getFloat("isentropic_vortex", "center_x", (float)((xmin + xmax) / 2));
  const double vortex_y = cm.getFloat("isentropic_vortex", "center_y", (float)((ymin + ymax) / 2));
  const double beta = cm.getFloat("isentropic_vortex", "strength", 5.0);
  for_each_cell(g, [&](int i, int j, int k) {
    const double x = g.x(i), y = g.y(j);
    const size_t o = g.at(i, j, k);
    const double xp = x - vortex_x, yp = y - vortex_y;
    const double r = sqrt(xp * xp + yp * yp);
    const double du = -yp * beta / (2 * M_PI) * exp(0.5 * (1.0 - r * r));
    const double dv = xp * beta / (2 * M_PI) * exp(0.5 * (1.0 - r * r));
    const double T = T_a - (gamma0 - 1) * beta * beta / (8 * gamma0 * M_PI * M_PI) * exp(1.0 - r * r);
    const double rho = rho_a * pow(T / T_a, 1.0 / (gamma0 - 1));
    U[o + g.sv * ID] = rho;
    U[o + g.sv * IU] = rho * (u_a + du);
    U[o + g.sv * IV] = rho * (v_a + dv);
    U[o + g.sv * IP] = rho * T / (gamma0 - 1.0) + 0.5 * rho * (u_a + du) * (u_a + du) + 0.5 * rho * (v_a + dv) * (v_a + dv);
  });
  return true;
}

// ---- MHD -------------------------------------------------------------------------------------------
bool init_mhd_implode(const HydroParams& p, const ConfigMap& cm, double* U)
{
  const Geo g(p);
  const double gamma0 = p.settings.gamma0;
  const double rho_out = cm.getFloat("implode", "density_outer", 1.0), p_out = cm.getFloat("implode", "pressure_outer", 1.0);
  const double u_out = cm.getFloat("implode", "vx_outer", 0.0), v_out = cm.getFloat("implode", "vy_outer", 0.0),
               w_out = cm.getFloat("implode", "vz_outer", 0.0);
  const double Bx_out = cm.getFloat("implode", "Bx_outer", 0.0), By_out = cm.getFloat("implode", "By_outer", 0.0),
               Bz_out = cm.getFloat("implode", "Bz_outer", 0.0);
  const double rho_in = cm.getFloat("implode", "density_inner", 0.125), p_in = cm.getFloat("implode", "pressure_inner", 0.14);
  const double u_in = cm.getFloat("implode", "vx_inner", 0.0), v_in = cm.getFloat("implode", "vy_inner", 0.0),
               w_in = cm.getFloat("implode", "vz_inner", 0.0);
  const double Bx_in = cm.getFloat("implode", "Bx_inner", 0.0), By_in = cm.getFloat("implode", "By_inner", 0.0),
               Bz_in = cm.getFloat("implode", "Bz_inner", 0.0);
  const int shape = (int)cm.getInteger("implode", "shape_region", 0);
  for_each_cell(g, [&](int i, int j, int k) {
    const double x = g.x(i), y = g.y(j);
    const size_t o = g.at(i, j, k);
    if (g.dim == 2) {
      const bool outer = (shape == 1) ? (x + y > 0.5 && x + y < 2.5) : (x + y > (g.xmin + g.xmax) / 2. + g.ymin);
      if (outer) {
        U[o + g.sv * ID] = rho_out;
        U[o + g.sv * IP] = p_out / (gamma0 - 1.0) + 0.5 * rho_out * (u_out * u_out + v_out * v_out) + 0.5 * (Bx_out * Bx_out + By_out * By_out);
        U[o + g.sv * IU] = u_out; U[o + g.sv * IV] = v_out; U[o + g.sv * IW] = 0.0;
        U[o + g.sv * IA] = Bx_out; U[o + g.sv * IB] = By_out; U[o + g.sv * IC] = 0.0;
      } else {
        U[o + g.sv * ID] = rho_in;
        U[o + g.sv * IP] = p_in / (gamma0 - 1.0) + 0.5 * rho_in * (u_in * u_in + v_in * v_in) + 0.5 * (Bx_in * Bx_in + By_in * By_in);
        U[o + g.sv * IU] = u_in; U[o + g.sv * IV] = v_in; U[o + g.sv * IW] = 0.0;
        U[o + g.sv * IA] = Bx_in; U[o + g.sv * IB] = By_in; U[o + g.sv * IC] = 0.0;
      }
    } else {
      const double z = g.z(k);
      const bool outer = (shape == 1) ? (x + y + z > 0.5 && x + y + z < 2.5) : (x + y + z > (g.xmin + g.xmax) / 2. + g.ymin + g.zmin);
      if (outer) {
        U[o + g.sv * ID] = rho_out;
        U[o + g.sv * IP] = p_out / (gamma0 - 1.0) + 0.5 * rho_out * (u_out * u_out + v_out * v_out + w_out * w_out) +
                           0.5 * (Bx_out * Bx_out + By_out * By_out + Bz_out * Bz_out);
        U[o + g.sv * IU] = u_out; U[o + g.sv * IV] = v_out; U[o + g.sv * IW] = w_out;
        U[o + g.sv * IA] = Bx_out; U[o + g.sv * IB] = By_out; U[o + g.sv * IC] = Bz_out;
      } else {
        U[o + g.sv * ID] = rho_in;
        U[o + g.sv * IP] = p_in / (gamma0 - 1.0) + 0.5 * rho_in * (u_in * u_in + v_in * v_in + w_in * w_in) +
                           0.5 * (Bx_in * Bx_in + By_in * By_in + Bz_in * Bz_in);
        U[o + g.sv * IU] = u_in; U[o + g.sv * IV] = v_in; U[o + g.sv * IW] = w_in;
        U[o + g.sv * IA] = Bx_in; U[o + g.sv * IB] = By_in; U[o + g.sv * IC] = Bz_in;
      }
    }
  });
  return true;
}

bool init_mhd_brio_wu(const HydroParams& p, const ConfigMap& cm, double* U)
{
  const Geo g(p);
  const double gamma0 = p.settings.gamma0;
  const double rhoL = cm.getFloat("brio-wu", "rhoL", 1.0), pL = cm.getFloat("brio-wu", "pL", 1.0), uL = cm.getFloat("brio-wu", "uL", 0.0),
               ByL = cm.getFloat("brio-wu", "ByL", 1.0);
  const double rhoR = cm.getFloat("brio-wu", "rhoR", 0.125), pR = cm.getFloat("brio-wu", "pR", 0.1), uR = cm.getFloat("brio-wu", "uR", 0.0),
               ByR = cm.getFloat("brio-wu", "ByR", -1.0);
  const double Bx = cm.getFloat("brio-wu", "Bx", 0.75);
  const int dir0 = (int)cm.getInteger("brio-wu", "direction", 0);
  if (dir0 < 0 || dir0 >= g.dim) {
    std::cerr << "[euler_b200] brio-wu/direction must be 0.." << g.dim - 1 << "\n";
    return false;
  }
  const char* lo[3] = { "xmin", "ymin", "zmin" };
  const char* hi[3] = { "xmax", "ymax", "zmax" };
  const float lof = cm.getFloat("mesh", lo[dir0], 0.0), hif = cm.getFloat("mesh", hi[dir0], 1.0); // float + float, then / 2.0 in double
  const double xd = cm.getFloat("brio-wu", "xd", (float)((lof + hif) / 2.0));
  const int dir1 = (dir0 + 1) % g.dim, dir2 = (dir0 + 2) % 3;
  for_each_cell(g, [&](int i, int j, int k) {
    const double pos = (dir0 == 0) ? g.x(i) : (dir0 == 1) ? g.y(j) : g.z(k);
    const size_t o = g.at(i, j, k);
    const bool left = pos <= xd;
    const double rho = left ? rhoL : rhoR, pr = left ? pL : pR, u = left ? uL : uR, By = left ? ByL : ByR;
    U[o + g.sv * ID] = rho;
    U[o + g.sv * IP] = pr / (gamma0 - 1.0) + 0.5 * rho * u * u + 0.5 * (Bx * Bx + By * By);
    U[o + g.sv * (IU + dir0)] = rho * u;
    U[o + g.sv * (IU + dir1)] = 0.0;
    U[o + g.sv * (IA + dir0)] = Bx;
    U[o + g.sv * (IA + dir1)] = By;
    if (g.dim == 2) {
      U[o + g.sv * IW] = 0.0;
      U[o + g.sv * IC] = 0.0;
    } else {
      U[o + g.sv * (IU + dir2)] = 0.0;
      U[o + g.sv * (IA + dir2)] = 0.0;
    }
  });
  return true;
}

bool init_mhd_blast(const HydroParams& p, const ConfigMap& cm, double* U)
{
  const Geo g(p);
  const double gamma0 = p.settings.gamma0;
  const double xmin = cm.getFloat("mesh", "xmin", 0.0), ymin = cm.getFloat("mesh", "ymin", 0.0), zmin = cm.getFloat("mesh", "zmin", 0.0);
  const double xmax = cm.getFloat("mesh", "xmax", 1.0), ymax = cm.getFloat("mesh", "ymax", 1.0), zmax = cm.getFloat("mesh", "zmax", 1.0);
  const double radius = cm.getFloat("blast", "radius", (float)((xmin + xmax) / 2.0 / 10));
  const double cx = cm.getFloat("blast", "center_x", (float)((xmin + xmax) / 2));
  const double cy = cm.getFloat("blast", "center_y", (float)((ymin + ymax) / 2));
  const double cz = cm.getFloat("blast", "center_z", (float)((zmin + zmax) / 2));
  const double d_in = cm.getFloat("blast", "density_in", 1.0), d_out = cm.getFloat("blast", "density_out", 1.2);
  const double p_in = cm.getFloat("blast", "pressure_in", 10.0), p_out = cm.getFloat("blast", "pressure_out", 0.1);
  const double bx_in = cm.getFloat("blast", "bx_in", 0.5), bx_out = cm.getFloat("blast", "bx_out", 0.5);
  // BlastParams.h:56-62: by_* is read from "by_*" and then overwritten from "bz_*"
  cm.getFloat("blast", "by_in", 0.5);
  const double by_in = cm.getFloat("blast", "bz_in", 0.5);
  cm.getFloat("blast", "by_out", 0.5);
  const double by_out = cm.getFloat("blast", "bz_out", 0.5);
  const double bz_in = 0.0, bz_out = 0.0; // never assigned by the reference (see the header of this file)
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
    const double A = in ? bx_in : bx_out, B = in ? by_in : by_out, C = in ? bz_in : bz_out;
    U[o + g.sv * ID] = in ? d_in : d_out;
    U[o + g.sv * IU] = 0.0;
    U[o + g.sv * IV] = 0.0;
    U[o + g.sv * IW] = 0.0;
    U[o + g.sv * IA] = A;
    U[o + g.sv * IB] = B;
    U[o + g.sv * IC] = C;
    U[o + g.sv * IP] = (in ? p_in : p_out) / (gamma0 - 1.0) + 0.5 * (sqr(A) + sqr(B) + sqr(C));
  });
  return true;
}

bool init_mhd_kelvin_helmholtz(const HydroParams& p, const ConfigMap& cm, double* U)
{
  const Geo g(p);
  const KH kh(cm);
  const double gamma0 = p.settings.gamma0;
  const double ampl = kh.amplitude, pressure = kh.pressure;
  const double PI = 3.141592653589793238462643383279502884L; // PI_F (real_type.h)
  const double bx = 0.5, by = 0.0, bz = 0.0;
  const KHRandom rnd(g, p, kh, g.dim);
  for_each_cell(g, [&](int i, int j, int k) {
    const double x = g.x(i), y = g.y(j);
    const size_t o = g.at(i, j, k);
    if (g.dim == 2) {
      const double yn = (y - g.ymin) / (g.ymax - g.ymin);
      double d, u, v;
      if (kh.p_rand) {
        if (yn < 0.25 || yn > 0.75) { d = kh.d_out; u = kh.vflow_out; v = 0.0; } else { d = kh.d_in; u = kh.vflow_in; v = 0.0; }
        u += ampl * (rnd(i, j, k, 0) - 0.5);
        v += ampl * (rnd(i, j, k, 1) - 0.5);
      } else if (kh.p_sine_rob) {
        const int n = kh.mode;
        const double y1 = 0.25, y2 = 0.75, rho1 = kh.d_in, rho2 = kh.d_out, v1 = kh.vflow_in, v2 = kh.vflow_out;
        const double ramp = 1.0 / (1.0 + exp(2 * (y - y1) / kh.delta)) + 1.0 / (1.0 + exp(2 * (y2 - y) / kh.delta));
        d = rho1 + ramp * (rho2 - rho1);
        u = v1 + ramp * (v2 - v1);
        v = kh.w0 * sin(n * PI * x);
      } else if (kh.p_sine) {
        const int n = kh.mode;
        const double y1 = 0.25, y2 = 0.75;
        d = (y >= y1 && y <= y2) ? kh.d_in : kh.d_out;
        u = (y >= y1 && y <= y2) ? kh.vflow_in : kh.vflow_out;
        v = kh.w0 * sin(n * PI * x);
      } else {
        return;
      }
      U[o + g.sv * ID] = d;
      U[o + g.sv * IU] = d * u;
      U[o + g.sv * IV] = d * v;
      U[o + g.sv * IA] = bx;
      U[o + g.sv * IB] = by;
      U[o + g.sv * IC] = bz;
      U[o + g.sv * IP] = pressure / (gamma0 - 1.0) + 0.5 * d * (u * u + v * v) + 0.5 * (bx * bx + by * by + bz * bz);
    } else {
      const double z = g.z(k);
      const double zn = (z - g.zmin) / (g.zmax - g.zmin);
      if (kh.p_rand) {
        // the 3-D MHD functor leaves the magnetic field at zero in this branch (MHDInitFunctors3D.h:623-660)
        double d, u, v = 0.0, w = 0.0;
        if (zn < 0.25 || zn > 0.75) { d = kh.d_out; u = kh.vflow_out; } else { d = kh.d_in; u = kh.vflow_in; }
        u += ampl * (rnd(i, j, k, 0) - 0.5);
        v += ampl * (rnd(i, j, k, 1) - 0.5);
        w += ampl * (rnd(i, j, k, 2) - 0.5);
        U[o + g.sv * ID] = d;
        U[o + g.sv * IU] = d * u;
        U[o + g.sv * IV] = d * v;
        U[o + g.sv * IW] = d * w;
        U[o + g.sv * IP] = pressure / (gamma0 - 1.0) + 0.5 * d * (u * u + v * v + w * w);
        return;
      }
      double d, u, v, w;
      if (kh.p_sine_rob) {
        const int n = kh.mode;
        const double z1 = 0.25, z2 = 0.75, rho1 = kh.d_in, rho2 = kh.d_out;
        const double v1x = kh.vflow_in, v2x = kh.vflow_out, v1y = kh.vflow_in / 2, v2y = kh.vflow_out / 2;
        const double ramp = 1.0 / (1.0 + exp(2 * (z - z1) / kh.delta)) + 1.0 / (1.0 + exp(2 * (z2 - z) / kh.delta));
        d = rho1 + ramp * (rho2 - rho1);
        u = v1x + ramp * (v2x - v1x);
        v = v1y + ramp * (v2y - v1y);
        w = kh.w0 * sin(n * M_PI * x) * sin(n * M_PI * y);
      } else if (kh.p_sine) {
        const int n = kh.mode;
        const double z1 = 0.25, z2 = 0.75;
        d = (z >= z1 && z <= z2) ? kh.d_in : kh.d_out;
        u = (z >= z1 && z <= z2) ? kh.vflow_in : kh.vflow_out;
        v = 0;
        w = kh.w0 * sin(n * M_PI * x);
      } else {
        return;
      }
      U[o + g.sv * ID] = d;
      U[o + g.sv * IU] = d * u;
      U[o + g.sv * IV] = d * v;
      U[o + g.sv * IW] = d * w;
      U[o + g.sv * IA] = bx;
      U[o + g.sv * IB] = by;
      U[o + g.sv * IC] = bz;
      U[o + g.sv * IP] = pressure / (gamma0 - 1.0) + 0.5 * d * (u * u + v * v + w * w) + 0.5 * (bx * bx + by * by + bz * bz);
    }
  });
  return true;
}

bool init_mhd_rotor(const HydroParams& p, const ConfigMap& cm, double* U)
{
  const Geo g(p);
  const double gamma0 = p.settings.gamma0;
  const double r0 = cm.getFloat("rotor", "r0", 0.1), r1 = cm.getFloat("rotor", "r1", 0.115), u0 = cm.getFloat("rotor", "u0", 2.0),
               p0 = cm.getFloat("rotor", "p0", 1.0), b0 = cm.getFloat("rotor", "b0", (float)(5.0 / sqrt(4 * M_PI)));
  const double xCenter = (g.xmax + g.xmin) / 2, yCenter = (g.ymax + g.ymin) / 2;
  for_each_cell(g, [&](int i, int j, int k) {
    const double x = g.x(i), y = g.y(j);
    const size_t o = g.at(i, j, k);
    const double r = sqrt((x - xCenter) * (x - xCenter) + (y - yCenter) * (y - yCenter));
    const double f_r = (r1 - r) / (r1 - r0);
    double& d = U[o + g.sv * ID];
    double& mu = U[o + g.sv * IU];
    double& mv = U[o + g.sv * IV];
    double& mw = U[o + g.sv * IW];
    double& A = U[o + g.sv * IA];
    if (g.dim == 2) {
      // the 2-D functor scales the momenta by the density and (also inside r0) by f_r, MHDInitFunctors2D.h:741-760
      if (r <= r0) {
        d = 10.0;
        mu = -d * f_r * u0 * (y - yCenter) / r0;
        mv = d * f_r * u0 * (x - xCenter) / r0;
      } else if (r <= r1) {
        d = 1 + 9 * f_r;
        mu = -d * f_r * u0 * (y - yCenter) / r;
        mv = d * f_r * u0 * (x - xCenter) / r;
      } else {
        d = 1.0;
        mu = 0.0;
        mv = 0.0;
      }
    } else {
      // the 3-D functor stores velocities in the momentum slots, MHDInitFunctors3D.h:838-857
      if (r <= r0) {
        d = 10.0;
        mu = -u0 * (y - yCenter) / r0;
        mv = u0 * (x - xCenter) / r0;
      } else if (r <= r1) {
        d = 1 + 9 * f_r;
        mu = -f_r * u0 * (y - yCenter) / r;
        mv = f_r * u0 * (x - xCenter) / r;
      } else {
        d = 1.0;
        mu = 0.0;
        mv = 0.0;
      }
    }
    mw = 0.0;
    A = b0;
    U[o + g.sv * IB] = 0.0;
    U[o + g.sv * IC] = 0.0;
    if (g.dim == 2)
      U[o + g.sv * IP] = p0 / (gamma0 - 1.0) + 0.5 * (mu * mu + mv * mv + mw * mw) / d + 0.5 * (A * A);
    else
      U[o + g.sv * IP] = p0 / (gamma0 - 1.0) + (mu * mu + mv * mv + mw * mw) / 2 / d + (A * A) / 2;
  });
  return true;
}

bool init_mhd_field_loop(const HydroParams& p, const ConfigMap& cm, double* U)
{
  const Geo g(p);
  const double gamma0 = p.settings.gamma0;
  const double radius = cm.getFloat("FieldLoop", "radius", 1.0), density_in = cm.getFloat("FieldLoop", "density_in", 1.0),
               amplitude = cm.getFloat("FieldLoop", "amplitude", 1.0), vflow = cm.getFloat("FieldLoop", "vflow", 1.0);
  // z component of the vector potential at every (ghosted) cell; the x and y components are zero
  std::vector<double> Az(g.sv, 0.0);
  for_each_cell(g, [&](int i, int j, int k) {
    const double x = g.x(i), y = g.y(j);
    const double r = sqrt(x * x + y * y);
    Az[g.at(i, j, k)] = (r < radius) ? amplitude * (radius - r) : 0.0;
  });
  const double cos_theta = 2.0 / sqrt(5.0);
  const double sin_theta = sqrt(1 - cos_theta * cos_theta);
  const int gw = g.gw;
  auto interior = [&](int i, int j, int k) {
    return i >= gw && i < g.isize - gw && j >= gw && j < g.jsize - gw && (g.dim == 2 || (k >= gw && k < g.ksize - gw));
  };
  // ghost cells are left at zero, as in the reference (they are filled by the first make_boundaries)
  for_each_cell(g, [&](int i, int j, int k) {
    if (!interior(i, j, k))
      return;
    const double x = g.x(i), y = g.y(j);
    const size_t o = g.at(i, j, k);
    const double r = sqrt(x * x + y * y);
    double& d = U[o + g.sv * ID];
    d = (r < radius) ? density_in : 1.0;
    U[o + g.sv * IU] = d * vflow * cos_theta;
    U[o + g.sv * IV] = d * vflow * sin_theta;
    if (g.dim == 2) {
      const double diag = sqrt(1.0 * (g.nx * g.nx + g.ny * g.ny + g.nz * g.nz));
      U[o + g.sv * IW] = d * vflow * g.nz / diag;
      U[o + g.sv * IA] = (Az[o + g.sj] - Az[o]) / g.dy;
      U[o + g.sv * IB] = -(Az[o + 1] - Az[o]) / g.dx;
      U[o + g.sv * IC] = 0.0;
    } else {
      U[o + g.sv * IW] = d * vflow;
      // curl of (0, 0, Az): the A_x, A_y differences of the reference are differences of zeros
      U[o + g.sv * IA] = (Az[o + g.sj] - Az[o]) / g.dy - (0.0 - 0.0) / g.dz;
      U[o + g.sv * IB] = (0.0 - 0.0) / g.dz - (Az[o + 1] - Az[o]) / g.dx;
      U[o + g.sv * IC] = (0.0 - 0.0) / g.dx - (0.0 - 0.0) / g.dy;
    }
  });
  for_each_cell(g, [&](int i, int j, int k) {
    if (!interior(i, j, k))
      return;
    const size_t o = g.at(i, j, k);
    const double d = U[o + g.sv * ID], mu = U[o + g.sv * IU], mv = U[o + g.sv * IV], mw = U[o + g.sv * IW];
    if (g.dim == 2)
      U[o + g.sv * IP] = 1.0f / (gamma0 - 1.0) +
                         0.5 * (0.25 * sqr(U[o + g.sv * IA] + U[o + 1 + g.sv * IA]) + 0.25 * sqr(U[o + g.sv * IB] + U[o + g.sj + g.sv * IB])) +
                         0.5 * (mu * mu + mv * mv) / d;
    else
      U[o + g.sv * IP] = 1.0f / (gamma0 - 1.0) +
                         0.5 * (0.25 * sqr(U[o + g.sv * IA] + U[o + 1 + g.sv * IA]) + 0.25 * sqr(U[o + g.sv * IB] + U[o + g.sj + g.sv * IB]) +
                                0.25 * sqr(U[o + g.sv * IC] + U[o + g.sk + g.sv * IC])) +
                         0.5 * (mu * mu + mv * mv + mw * mw) / d;
  });
  return true;
}

} // namespace euler_b200
