// C ABI of libek_b200.so (include/ek_b200.h): context management, state transfer, ghost fill,
// CFL time step, one fused step, and the sync-free time loop.  No CPU fallback: every entry point
// that computes fails with EK_ERR_CUDA when there is no usable CUDA device.
#include <cstdio>
#include <cstring>
#include <new>
#include "ek_internal.h"
#include "ek_grid.h"
#include "boundary_kernels.cuh"

using namespace ek;

static thread_local std::string g_create_err;

#define EK_CUDA(ctx, call)                                                                     \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                         \
      return EK_ERR_CUDA;                                                                      \
    }                                                                                          \
  } while (0)

// ---- tiny kernels of the device-resident time loop ---------------------------------------------
// dt = cfl / invDt, clamped to t_end - t  (SolverHydroMuscl::compute_dt_local, SolverHydroMuscl.h:627;
// SolverBase::compute_dt, shared/SolverBase.cpp:149-152).  Consumes and resets the running max.
__global__ void pre_step_kernel(TimeState* ts, double cfl, int use_host_dt, double host_dt)
{
  if (ts->finished)
    return;
  double dt;
  if (use_host_dt) {
    dt = host_dt;
  } else {
    const double invdt = __longlong_as_double((long long)ts->invdt_bits);
    dt = cfl / invdt;
    if (ts->t + dt > ts->t_end)
      dt = ts->t_end - ts->t;
  }
  ts->dt = dt;
  ts->invdt_bits = 0ull;
}

// ++iteration; t += dt; finished? (SolverBase::next_iteration / finished, SolverBase.cpp:171-195)
__global__ void post_step_kernel(TimeState* ts)
{
  if (ts->finished)
    return;
  ts->t += ts->dt;
  ts->steps_done += 1;
  if (ts->t >= ts->t_end - 1e-14 || ts->steps_done >= ts->max_steps)
    ts->finished = 1;
}

__global__ void reset_time_kernel(TimeState* ts, double t, double t_end, int max_steps)
{
  ts->t = t;
  ts->t_end = t_end;
  ts->steps_done = 0;
  ts->finished = 0;
  ts->max_steps = max_steps;
}

// ---- diagnostics ---------------------------------------------------------------------------------
// Sums of the conserved variables over the interior, accurate enough to SHOW conservation to round-off at 512^3:
// compensated (Neumaier) accumulation per thread, fixed-shape tree per block, block partials combined by a second
// compensated pass.  Deterministic for a given grid (a plain 256-thread sum loses ~1e-11 of the total on a uniform
// field of 1.3e8 cells, which is larger than the drift it is meant to detect).
constexpr int SUMS_BLOCKS = 512; // partial sums per variable

__device__ __forceinline__ void neumaier_add(double& s, double& comp, double v)
{
  const double t = s + v;
  comp += (fabs(s) >= fabs(v)) ? ((s - t) + v) : ((v - t) + s);
  s = t;
}

__global__ void __launch_bounds__(256) conserved_sums_kernel(const Grid g, const double* __restrict__ U, double* __restrict__ partial)
{
  const int v = blockIdx.y;
  const long long ncell = (long long)g.nx * g.ny * g.nz;
  const long long per_block = (ncell + gridDim.x - 1) / gridDim.x;
  const long long c0 = per_block * blockIdx.x;
  const long long c1 = (c0 + per_block < ncell) ? c0 + per_block : ncell;
  double s = 0.0, comp = 0.0;
  for (long long c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
    const int i = (int)(c % g.nx) + g.gw;
    const int j = (int)((c / g.nx) % g.ny) + g.gw;
    const int k = (g.ksize > 1) ? (int)(c / ((long long)g.nx * g.ny)) + g.gw : 0;
    neumaier_add(s, comp, U[i + g.sj * j + g.sk * k + g.sv * v]);
  }
  __shared__ double red[256], redc[256];
  red[threadIdx.x] = s;
  redc[threadIdx.x] = comp;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      double a = red[threadIdx.x], ca = redc[threadIdx.x] + redc[threadIdx.x + o];
      neumaier_add(a, ca, red[threadIdx.x + o]);
      red[threadIdx.x] = a;
      redc[threadIdx.x] = ca;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[(size_t)v * 2 * gridDim.x + 2 * blockIdx.x] = red[0];
    partial[(size_t)v * 2 * gridDim.x + 2 * blockIdx.x + 1] = redc[0];
  }
}

__global__ void conserved_sums_final_kernel(const double* __restrict__ partial, int nblk, double* __restrict__ out)
{
  const int v = blockIdx.x;
  if (threadIdx.x != 0)
    return;
  double s = 0.0, comp = 0.0;
  for (int b = 0; b < nblk; ++b) {
    neumaier_add(s, comp, partial[(size_t)v * 2 * nblk + 2 * b]);
    comp += partial[(size_t)v * 2 * nblk + 2 * b + 1];
  }
  out[v] = s + comp;
}

__global__ void __launch_bounds__(256) max_divb_kernel(const Grid g, const double* __restrict__ U, unsigned long long* __restrict__ out_bits)
{
  const long long ncell = (long long)g.nx * g.ny * g.nz;
  double m = 0.0;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(c % g.nx) + g.gw;
    const int j = (int)((c / g.nx) % g.ny) + g.gw;
    const int k = (g.ksize > 1) ? (int)(c / ((long long)g.nx * g.ny)) + g.gw : 0;
    const long long o = i + g.sj * j + g.sk * k;
    double d = (U[o + 1 + g.sv * IA] - U[o + g.sv * IA]) / g.dx + (U[o + g.sj + g.sv * IB] - U[o + g.sv * IB]) / g.dy;
    if (g.ksize > 1)
      d += (U[o + g.sk + g.sv * IC] - U[o + g.sv * IC]) / g.dz;
    m = fmax(m, fabs(d));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0)
    atomic_max_nonneg(out_bits, m);
}

// device-side arithmetic primitives on caller-supplied operands (ek_selftest_math): the inline-PTX select forms and
// the MUFU-seeded division / square roots of ek_common.cuh next to their IEEE counterparts
__global__ void selftest_math_kernel(const double* __restrict__ x, int n, double* __restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const double a = x[i];
  out[i] = dmax0(a);
  out[n + i] = dmin0(a);
  out[2 * n + i] = rcp_fast(a);
  out[3 * n + i] = sqrt_fast(a);
  out[4 * n + i] = rsqrt_fast(a);
  out[5 * n + i] = div_fast(1.0 + a, a);
}

// ---- helpers -------------------------------------------------------------------------------------
static int validate(const ek_params& p, std::string& err)
{
  if (p.dim != 2 && p.dim != 3) { err = "dim must be 2 or 3"; return EK_ERR_INVALID; }
  const int nbvar = p.mhd ? 8 : (p.dim == 3 ? 5 : 4);
  const int gw = p.mhd ? 3 : 2;
  if (p.nbvar != nbvar || p.ghost != gw) { err = "nbvar/ghost inconsistent with solver (HydroParams.cpp:94-115)"; return EK_ERR_INVALID; }
  if (p.nx < gw || p.ny < gw || (p.dim == 3 && p.nz < gw)) { err = "interior sizes must be >= ghost width"; return EK_ERR_INVALID; }
  if (p.dim == 2 && p.nz != 1) { err = "nz must be 1 in 2-D"; return EK_ERR_INVALID; }
  if (!(p.dx > 0) || !(p.dy > 0) || (p.dim == 3 && !(p.dz > 0))) { err = "dx,dy,dz must be > 0"; return EK_ERR_INVALID; }
  if (p.impl_version < 0 || p.impl_version > 2) { err = "implementationVersion must be 0,1,2"; return EK_ERR_INVALID; }
  if (!(p.cfl > 0.0)) { err = "cfl must be > 0"; return EK_ERR_INVALID; }
  if (!(p.gamma0 > 1.0)) { err = "gamma0 must be > 1 (entho = 1/(gamma0-1))"; return EK_ERR_INVALID; }
  // slope_unsplit_hydro_{2d,3d} know slope_type 0 (no slopes), 1 and 2 only (HydroBaseFunctor3D.h:413-578,
  // MHDBaseFunctor3D.h:364-495); anything else would give a limiter that is neither the reference's nor exact
  if (!(p.slope_type == 0.0 || p.slope_type == 1.0 || p.slope_type == 2.0)) { err = "slope_type must be 0, 1 or 2"; return EK_ERR_INVALID; }
  if (!p.mhd) {
    const bool ok = p.riemann == EK_RIEMANN_APPROX || p.riemann == EK_RIEMANN_LLF || p.riemann == EK_RIEMANN_HLLC ||
                    (p.riemann == EK_RIEMANN_HLL && p.dim == 2);
    // the reference's 3-D dispatcher has no HLL branch and leaves the flux uninitialised
    // (shared/RiemannSolvers.h:610-633); hlld is MHD-only
    if (!ok) { err = "riemann solver not available for this hydro solver (RiemannSolvers.h:575-633)"; return EK_ERR_INVALID; }
  } else {
    const bool ok = p.riemann == EK_RIEMANN_HLLD || p.riemann == EK_RIEMANN_HLL || p.riemann == EK_RIEMANN_LLF;
    if (!ok) { err = "riemann solver not available for MHD (RiemannSolvers_MHD.h:371-389)"; return EK_ERR_INVALID; }
    if (p.gravity_enabled) { err = "gravity is hydro-only in the reference"; return EK_ERR_INVALID; }
    // implementationVersion 1 (3-D and 2-D) and 2 (2-D) are other schedules of the same arithmetic: the reference's
    // own builds agree with its version 0 to <= 2e-15 (atomic-add order) and are served by the version-0 kernels.
    // Version 2 in 3-D is a different scheme in the reference (MHDRunFunctors3D.h:865-1778): its results differ from
    // versions 0/1 by up to 1e-1 on z-dependent problems (test_mhd_brio_wu_3d_z.ini).  It is not
    // provided, and is refused rather than silently replaced.
    if (p.dim == 3 && p.impl_version == 2) {
      err = "MHD_Muscl_3D: implementationVersion=2 is not provided (the reference's version 2 differs from its versions 0/1 "
            "on z-dependent problems, DESIGN.md section 7); use other/implementationVersion=0 or 1";
      return EK_ERR_INVALID;
    }
  }
  for (int f = 0; f < 6; ++f)
    if (p.bc[f] < 0 || p.bc[f] > EK_BC_COPY) { err = "bad boundary type"; return EK_ERR_INVALID; }
  return EK_OK;
}

static int blocks_for(long long n, int sm_count)
{
  long long b = (n + 255) / 256;
  const long long cap = (long long)sm_count * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// physical BCs + halo exchange, X then Y then Z (SolverBase.cpp:464-512 serial, :583-664 MPI order)
static int fill_boundaries(Ctx& c, double* U)
{
  const Grid& g = c.g;
  const int ndir = (c.p.dim == 3) ? 3 : 2;
  if (c.halo_ready_for == U)
    EK_CUDA(&c, cudaStreamWaitEvent(c.stream, c.ev_halo, 0));
  for (int d = 0; d < ndir; ++d) {
    const bool any_copy = c.p.bc[2 * d] == EK_BC_COPY || c.p.bc[2 * d + 1] == EK_BC_COPY;
    if (any_copy && d == 2 && c.halo_ready_for == U) {
      // exchanged right after the end chunks of the previous step, while its middle chunks were still running
      // (waited for above).  The x and y fills then ran over ghost planes that already held the neighbour's
      // interior cells, which gives the same corner values as the reference order (fill x and y on both sides,
      // then exchange): a physical fill of one (j,k) line only reads interior cells of that line.
    } else if (any_copy) {
      int rc = comm_exchange_dir(c, U, d);
      if (rc != EK_OK)
        return rc;
    }
    if (c.p.bc[2 * d] == EK_BC_COPY && c.p.bc[2 * d + 1] == EK_BC_COPY)
      continue;
    const long long ext0 = (d == 0) ? 2 * g.gw : g.isize, ext1 = (d == 1) ? 2 * g.gw : g.jsize,
                    ext2 = (d == 2) ? 2 * g.gw : g.ksize;
    const int nb = blocks_for(ext0 * ext1 * ext2, c.sm_count);
    if (d == 0)
      fill_ghosts_kernel<0><<<nb, 256, 0, c.stream>>>(g, U, c.p.mhd, c.ts);
    else if (d == 1)
      fill_ghosts_kernel<1><<<nb, 256, 0, c.stream>>>(g, U, c.p.mhd, c.ts);
    else
      fill_ghosts_kernel<2><<<nb, 256, 0, c.stream>>>(g, U, c.p.mhd, c.ts);
    c.launches++;
    EK_CUDA(&c, cudaGetLastError());
  }
  if (c.halo_ready_for == U)
    c.halo_ready_for = nullptr; // consumed
  c.ghosts_valid_for = U;
  return EK_OK;
}

// MHD: max invDt of the ghost-filled level U into ts->invdt_bits.  When the update kernel that produced U has already
// reduced every interior cell but the last layers (mhd_dt_partial_for == U), only those layers are left.
static int mhd_local_invdt(Ctx& c, const double* U)
{
  if (c.mhd_dt_partial_for == U && mhd_dt_folded(c)) {
    EK_CUDA(&c, launch_mhd_invdt_layers(c, U));
  } else {
    EK_CUDA(&c, cudaMemsetAsync(&c.ts->invdt_bits, 0, sizeof(unsigned long long), c.stream));
    EK_CUDA(&c, launch_mhd_invdt(c, U));
  }
  // the all-reduce that follows leaves the GLOBAL maximum in invdt_bits: still a valid starting value for a repeated
  // call on the same level (a maximum is idempotent), so mhd_dt_partial_for stays
  return EK_OK;
}

// max invDt of the current level into ts->invdt_bits (+ max over ranks)
static int prime_invdt(Ctx& c)
{
  if (c.p.mhd) {
    int rc = fill_boundaries(c, c.U[c.cur]); // MHD dt reads primitives that need the upper ghost faces
    if (rc != EK_OK)
      return rc;
    rc = mhd_local_invdt(c, c.U[c.cur]);
    if (rc != EK_OK)
      return rc;
  } else {
    EK_CUDA(&c, cudaMemsetAsync(&c.ts->invdt_bits, 0, sizeof(unsigned long long), c.stream));
    EK_CUDA(&c, launch_hydro_invdt(c, c.U[c.cur]));
  }
  int rc = comm_allreduce_invdt(c);
  if (rc != EK_OK)
    return rc;
  c.dt_primed = true;
  return EK_OK;
}

// one godunov_unsplit on the stream; dt comes from the device (use_host_dt = 0) or from the host
static int enqueue_step(Ctx& c, int use_host_dt, double host_dt)
{
  double* Uin = c.U[c.cur];
  double* Uout = c.U[1 - c.cur];
  if (c.p.mhd) {
    // MHD: boundaries -> (prims) -> dt -> step   (SolverMHDMuscl.cpp:463-513).  The host-driven loop calls
    // ek_compute_dt first, which has just filled and exchanged the ghosts of this level: not repeated (idempotent anyway)
    int rc = (c.ghosts_valid_for == Uin) ? EK_OK : fill_boundaries(c, Uin);
    if (rc != EK_OK)
      return rc;
    if (!use_host_dt) {
      rc = mhd_local_invdt(c, Uin);
      if (rc != EK_OK)
        return rc;
      rc = comm_allreduce_invdt(c);
      if (rc != EK_OK)
        return rc;
    }
    pre_step_kernel<<<1, 1, 0, c.stream>>>(c.ts, c.p.cfl, use_host_dt, host_dt); // zeroes invdt_bits
    c.launches++;
    c.mhd_dt_partial_for = mhd_dt_folded(c) ? Uout : nullptr; // the update kernel reduces the new level
    ktime_mark(c);
    if (c.overlap_ok) {
      // every stage, the update of the END planes first; they travel to the z neighbours beside the update of the rest
      EK_CUDA(&c, launch_mhd_step(c, Uin, Uout, 1));
      EK_CUDA(&c, cudaEventRecord(c.ev_edge, c.stream));
      EK_CUDA(&c, cudaStreamWaitEvent(c.comm_stream, c.ev_edge, 0));
      rc = comm_exchange_dir(c, Uout, 2, c.comm_stream);
      if (rc != EK_OK)
        return rc;
      EK_CUDA(&c, cudaEventRecord(c.ev_halo, c.comm_stream));
      c.halo_ready_for = Uout;
      EK_CUDA(&c, launch_mhd_step(c, Uin, Uout, 2));
    } else {
      EK_CUDA(&c, launch_mhd_step(c, Uin, Uout));
    }
    ktime_mark(c);
  } else {
    // hydro: dt (from the previous update's fused reduction) -> boundaries -> fused step
    // (SolverHydroMuscl.h:637-679, SolverHydroMuscl.cpp:306-390)
    if (!use_host_dt && !c.dt_primed) {
      int rc = prime_invdt(c);
      if (rc != EK_OK)
        return rc;
    }
    pre_step_kernel<<<1, 1, 0, c.stream>>>(c.ts, c.p.cfl, use_host_dt, host_dt);
    c.launches++;
    int rc = fill_boundaries(c, Uin);
    if (rc != EK_OK)
      return rc;
    ktime_mark(c);
    if (c.overlap_ok) {
      // end chunks first; their planes travel to the z neighbours (comm stream) while the middle chunks run
      EK_CUDA(&c, cudaEventRecord(c.ev_pre, c.stream));
      EK_CUDA(&c, launch_hydro_step(c, Uin, Uout, 1, c.stream));
      EK_CUDA(&c, cudaEventRecord(c.ev_edge, c.stream));
      EK_CUDA(&c, cudaStreamWaitEvent(c.aux_stream, c.ev_pre, 0));
      EK_CUDA(&c, launch_hydro_step(c, Uin, Uout, 2, c.aux_stream));
      EK_CUDA(&c, cudaEventRecord(c.ev_mid, c.aux_stream));
      EK_CUDA(&c, cudaStreamWaitEvent(c.comm_stream, c.ev_edge, 0));
      rc = comm_exchange_dir(c, Uout, 2, c.comm_stream);
      if (rc != EK_OK)
        return rc;
      EK_CUDA(&c, cudaEventRecord(c.ev_halo, c.comm_stream));
      c.halo_ready_for = Uout;
      EK_CUDA(&c, cudaStreamWaitEvent(c.stream, c.ev_mid, 0));
    } else {
      EK_CUDA(&c, launch_hydro_step(c, Uin, Uout));
    }
    ktime_mark(c);
    rc = comm_allreduce_invdt(c); // the fused reduction produced this rank's max for the next step
    if (rc != EK_OK)
      return rc;
    c.dt_primed = true;
  }
  c.ghosts_valid_for = nullptr; // the new level's ghosts are not filled yet
  post_step_kernel<<<1, 1, 0, c.stream>>>(c.ts);
  c.launches++;
  EK_CUDA(&c, cudaGetLastError());
  return EK_OK;
}

// ---- ABI -----------------------------------------------------------------------------------------
extern "C" {

int ek_abi_version(void) { return EK_ABI_VERSION; }

const char* ek_last_error(const ek_ctx* ctx)
{
  if (!ctx)
    return g_create_err.c_str();
  return reinterpret_cast<const Ctx*>(ctx)->err.c_str();
}

int ek_create(const ek_params* params, int device, ek_ctx** out)
{
  if (!params || !out) { g_create_err = "null argument"; return EK_ERR_INVALID; }
  *out = nullptr;
  int rc = validate(*params, g_create_err);
  if (rc != EK_OK)
    return rc;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_err = std::string("no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback";
    return EK_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { g_create_err = "device index out of range"; return EK_ERR_INVALID; }
  int prev_device = -1;
  cudaGetDevice(&prev_device); // restored before returning: creating a context must not move the caller's current device
  Ctx* c = new (std::nothrow) Ctx();
  if (!c) { g_create_err = "out of host memory"; return EK_ERR_INVALID; }
  c->p = *params;
  c->g = make_grid(*params);
  c->device = device;
  c->state_len = c->g.sv * c->g.nbvar;
#define CK(call)                                                                   \
  do {                                                                             \
    cudaError_t e2 = (call);                                                       \
    if (e2 != cudaSuccess) {                                                       \
      g_create_err = std::string(#call) + ": " + cudaGetErrorString(e2);           \
      ek_destroy(reinterpret_cast<ek_ctx*>(c));                                    \
      if (prev_device >= 0) cudaSetDevice(prev_device);                            \
      return EK_ERR_CUDA;                                                          \
    }                                                                              \
  } while (0)
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  {
    // the halo exchange must not queue behind the thousands of pending CTAs of the step kernel
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CK(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, prio_hi));
  }
  CK(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  CK(cudaEventCreateWithFlags(&c->ev_a, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_b, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_pre, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_edge, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_mid, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming));
  const size_t bytes = (size_t)c->state_len * sizeof(double);
  CK(cudaMalloc(&c->U[0], bytes));
  CK(cudaMalloc(&c->U[1], bytes));
  CK(cudaMemset(c->U[0], 0, bytes));
  CK(cudaMemset(c->U[1], 0, bytes));
  if (params->gravity_enabled) {
    CK(cudaMalloc(&c->grav, (size_t)c->g.sv * params->dim * sizeof(double)));
    CK(cudaMemset(c->grav, 0, (size_t)c->g.sv * params->dim * sizeof(double)));
  }
  if (params->mhd) {
    c->work_bytes = mhd_work_bytes(c->g);
    if (c->work_bytes)
      CK(cudaMalloc(&c->work, c->work_bytes));
    mhd_make_tensor_maps(*c); // optional: without them the plain-load P+E+T kernel is used
  }
  CK(cudaMalloc(&c->ts, sizeof(TimeState)));
  CK(cudaMemset(c->ts, 0, sizeof(TimeState)));
  CK(cudaMallocHost(&c->ts_host, sizeof(TimeState)));
  CK(cudaMalloc(&c->scratch, (64 + 2 * 8 * SUMS_BLOCKS) * sizeof(double)));
#undef CK
  if (prev_device >= 0)
    cudaSetDevice(prev_device);
  *out = reinterpret_cast<ek_ctx*>(c);
  return EK_OK;
}

int ek_destroy(ek_ctx* ctx)
{
  if (!ctx)
    return EK_OK;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  cudaSetDevice(c->device);
  ek_comm_detach(ctx);
  for (cudaEvent_t e : c->kev) cudaEventDestroy(e);
  if (c->own_stream) cudaStreamSynchronize(c->own_stream);
  if (c->io_stream) {
    cudaStreamSynchronize(c->io_stream);
    cudaStreamDestroy(c->io_stream);
  }
  if (c->ev_snap) cudaEventDestroy(c->ev_snap);
  cudaFree(c->snap);
  for (cudaGraphExec_t& ge : c->step_graph) {
    if (ge) cudaGraphExecDestroy(ge);
    ge = nullptr;
  }
  if (c->aux_stream) cudaStreamSynchronize(c->aux_stream);
  if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
  cudaFree(c->U[0]); cudaFree(c->U[1]); cudaFree(c->grav); cudaFree(c->work); cudaFree(c->ts); cudaFree(c->scratch);
  if (c->ts_host) cudaFreeHost(c->ts_host);
  if (c->ev_a) cudaEventDestroy(c->ev_a);
  if (c->ev_b) cudaEventDestroy(c->ev_b);
  if (c->ev_pre) cudaEventDestroy(c->ev_pre);
  if (c->ev_edge) cudaEventDestroy(c->ev_edge);
  if (c->ev_mid) cudaEventDestroy(c->ev_mid);
  if (c->ev_halo) cudaEventDestroy(c->ev_halo);
  if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  delete c;
  return EK_OK;
}

int ek_set_stream(ek_ctx* ctx, void* cuda_stream)
{
  if (!ctx) return EK_ERR_INVALID;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  cudaStream_t next = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : c->own_stream;
  if (next == c->stream)
    return EK_OK;
  // order the new stream after everything queued so far: the step stream and the helper streams whose results
  // (middle chunks, halo planes) the next call reads
  EK_CUDA(c, cudaSetDevice(c->device));
  cudaStream_t prev[3] = { c->stream, c->aux_stream, c->comm_stream };
  for (cudaStream_t st : prev) {
    if (!st)
      continue;
    EK_CUDA(c, cudaEventRecord(c->ev_a, st));
    EK_CUDA(c, cudaStreamWaitEvent(next, c->ev_a, 0));
  }
  c->stream = next;
  return EK_OK;
}

int ek_sync(ek_ctx* ctx)
{
  if (!ctx) return EK_ERR_INVALID;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  EK_CUDA(c, cudaSetDevice(c->device));
  EK_CUDA(c, cudaStreamSynchronize(c->stream));
  return EK_OK;
}

int64_t ek_state_len(const ek_ctx* ctx) { return ctx ? reinterpret_cast<const Ctx*>(ctx)->state_len : 0; }
int64_t ek_launch_count(const ek_ctx* ctx) { return ctx ? reinterpret_cast<const Ctx*>(ctx)->launches : 0; }

int ek_upload_state(ek_ctx* ctx, const double* U_host)
{
  if (!ctx || !U_host) return EK_ERR_INVALID;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  EK_CUDA(c, cudaSetDevice(c->device));
  const size_t bytes = (size_t)c->state_len * sizeof(double);
  c->cur = 0;
  EK_CUDA(c, cudaMemcpyAsync(c->U[0], U_host, bytes, cudaMemcpyHostToDevice, c->stream));
  EK_CUDA(c, cudaMemcpyAsync(c->U[1], c->U[0], bytes, cudaMemcpyDeviceToDevice, c->stream));
  c->dt_primed = false;
  c->mhd_prims_of = nullptr;
  c->mhd_dt_partial_for = nullptr;
  c->halo_ready_for = nullptr;
  c->ghosts_valid_for = nullptr;
  EK_CUDA(c, cudaStreamSynchronize(c->stream)); // the host buffer may be pageable / reused
  return EK_OK;
}

int ek_download_state(ek_ctx* ctx, double* U_host)
{
  if (!ctx || !U_host) return EK_ERR_INVALID;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  EK_CUDA(c, cudaSetDevice(c->device));
  EK_CUDA(c, cudaMemcpyAsync(U_host, c->U[c->cur], (size_t)c->state_len * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  EK_CUDA(c, cudaStreamSynchronize(c->stream));
  return EK_OK;
}

int ek_alloc_host(size_t bytes, void** out)
{
  if (!out) return EK_ERR_INVALID;
  *out = nullptr;
  if (cudaMallocHost(out, bytes) != cudaSuccess) {
    cudaGetLastError();
    g_create_err = "cudaMallocHost failed";
    return EK_ERR_CUDA;
  }
  return EK_OK;
}

int ek_free_host(void* p)
{
  if (p) cudaFreeHost(p);
  return EK_OK;
}

int ek_download_wait(ek_ctx* ctx)
{
  if (!ctx) return EK_ERR_INVALID;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c->io_pending)
    return EK_OK;
  EK_CUDA(c, cudaSetDevice(c->device));
  EK_CUDA(c, cudaStreamSynchronize(c->io_stream));
  c->io_pending = false;
  return EK_OK;
}

int ek_download_state_async(ek_ctx* ctx, double* U_host)
{
  if (!ctx || !U_host) return EK_ERR_INVALID;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  EK_CUDA(c, cudaSetDevice(c->device));
  int rc = ek_download_wait(ctx); // one in flight: the snapshot buffer is about to be overwritten
  if (rc != EK_OK)
    return rc;
  const size_t bytes = (size_t)c->state_len * sizeof(double);
  if (!c->snap) {
    if (cudaMalloc(&c->snap, bytes) != cudaSuccess) {
      cudaGetLastError();
      c->snap = nullptr;
      return ek_download_state(ctx, U_host); // no room for a snapshot: the synchronous copy
    }
    EK_CUDA(c, cudaStreamCreateWithFlags(&c->io_stream, cudaStreamNonBlocking));
    EK_CUDA(c, cudaEventCreateWithFlags(&c->ev_snap, cudaEventDisableTiming));
  }
  EK_CUDA(c, cudaMemcpyAsync(c->snap, c->U[c->cur], bytes, cudaMemcpyDeviceToDevice, c->stream));
  EK_CUDA(c, cudaEventRecord(c->ev_snap, c->stream));
  EK_CUDA(c, cudaStreamWaitEvent(c->io_stream, c->ev_snap, 0));
  EK_CUDA(c, cudaMemcpyAsync(U_host, c->snap, bytes, cudaMemcpyDeviceToHost, c->io_stream));
  c->io_pending = true;
  return EK_OK;
}

int ek_device_state(ek_ctx* ctx, double** dptr)
{
  if (!ctx || !dptr) return EK_ERR_INVALID;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  *dptr = c->U[c->cur];
  c->dt_primed = false; // the caller may modify the state through this pointer
  c->ghosts_valid_for = nullptr;
  c->mhd_prims_of = nullptr;
  c->mhd_dt_partial_for = nullptr;
  c->halo_ready_for = nullptr;
  return EK_OK;
}

int ek_set_gravity(ek_ctx* ctx, const double* g_host)
{
  if (!ctx || !g_host) return EK_ERR_INVALID;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c->grav) { c->err = "gravity not enabled in ek_params"; return EK_ERR_STATE; }
  EK_CUDA(c, cudaSetDevice(c->device));
  EK_CUDA(c, cudaMemcpyAsync(c->grav, g_host, (size_t)c->g.sv * c->p.dim * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  EK_CUDA(c, cudaStreamSynchronize(c->stream));
  c->dt_primed = false;
  return EK_OK;
}

int ek_fill_boundaries(ek_ctx* ctx)
{
  if (!ctx) return EK_ERR_INVALID;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  EK_CUDA(c, cudaSetDevice(c->device));
  // not part of a device-resident loop: never skipped by the 'finished' flag
  EK_CUDA(c, cudaMemsetAsync(&c->ts->finished, 0, sizeof(int), c->stream));
  return fill_boundaries(*c, c->U[c->cur]);
}

int ek_compute_dt(ek_ctx* ctx, double* dt_out)
{
  if (!ctx || !dt_out) return EK_ERR_INVALID;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  EK_CUDA(c, cudaSetDevice(c->device));
  EK_CUDA(c, cudaMemsetAsync(&c->ts->finished, 0, sizeof(int), c->stream));
  if (!c->dt_primed || c->p.mhd) {
    int rc = prime_invdt(*c);
    if (rc != EK_OK)
      return rc;
  }
  EK_CUDA(c, cudaMemcpyAsync(c->ts_host, c->ts, sizeof(TimeState), cudaMemcpyDeviceToHost, c->stream));
  EK_CUDA(c, cudaStreamSynchronize(c->stream));
  double invdt;
  memcpy(&invdt, &c->ts_host->invdt_bits, sizeof(double));
  *dt_out = c->p.cfl / invdt;
  return EK_OK;
}

int ek_step(ek_ctx* ctx, double dt)
{
  if (!ctx) return EK_ERR_INVALID;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  EK_CUDA(c, cudaSetDevice(c->device));
  reset_time_kernel<<<1, 1, 0, c->stream>>>(c->ts, 0.0, 1e300, 1 << 30);
  c->launches++;
  int rc = enqueue_step(*c, 1, dt);
  if (rc != EK_OK)
    return rc;
  c->cur = 1 - c->cur;
  if (c->halo_ready_for)
    EK_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_halo, 0)); // later work on the stream sees the exchanged ghosts
  return EK_OK;
}

// Two consecutive steps of the device-resident loop captured as a CUDA graph (the level ping-pong makes a pair periodic).
// For launch-bound grids (configs[0]: 1024 x 512 cells, 49 us of kernels in a 74 us step) the five launches of a step
// cost as much as a third of the work; a replay submits ten of them at once.  Single GPU only (no NCCL calls, one
// stream); not while per-kernel event timing is on.  EK_GRAPH=1 / 0 forces / forbids it; default: grids below 4 M cells.
static bool graph_wanted(const Ctx& c, int nsteps)
{
  if (c.nranks != 1 || c.nccl_comm || c.ktiming || nsteps < 3)
    return false;
  const char* e = getenv("EK_GRAPH");
  if (e)
    return atoi(e) != 0;
  return (long long)c.g.nx * c.g.ny * c.g.nz < (4LL << 20);
}

static int capture_step_pair(Ctx& c)
{
  const int par = c.cur;
  cudaGraph_t graph = nullptr;
  const int64_t l0 = c.launches;
  EK_CUDA(&c, cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal));
  int rc = EK_OK;
  for (int s = 0; s < 2 && rc == EK_OK; ++s) {
    rc = enqueue_step(c, 0, 0.0);
    c.cur = 1 - c.cur;
  }
  cudaError_t e = cudaStreamEndCapture(c.stream, &graph);
  c.launches = l0; // nothing ran: the replays are counted
  if (rc != EK_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess || !graph) {
    c.err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e);
    return EK_ERR_CUDA;
  }
  // kernels of one replay (memset nodes etc. are not launches of ours)
  size_t nodes = 0, kernels = 0;
  cudaGraphGetNodes(graph, nullptr, &nodes);
  std::vector<cudaGraphNode_t> nv(nodes);
  if (nodes && cudaGraphGetNodes(graph, nv.data(), &nodes) == cudaSuccess) {
    for (cudaGraphNode_t nd : nv) {
      cudaGraphNodeType ty;
      if (cudaGraphNodeGetType(nd, &ty) == cudaSuccess && ty == cudaGraphNodeTypeKernel)
        ++kernels;
    }
  }
  e = cudaGraphInstantiate(&c.step_graph[par], graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) {
    c.step_graph[par] = nullptr;
    c.err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e);
    return EK_ERR_CUDA;
  }
  c.graph_kernels[par] = (int64_t)kernels;
  return EK_OK;
}

int ek_run(ek_ctx* ctx, int nsteps, double t, double t_end, double* t_out, double* dt_last_out, int* steps_done_out)
{
  if (!ctx || nsteps < 0) return EK_ERR_INVALID;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  EK_CUDA(c, cudaSetDevice(c->device));
  reset_time_kernel<<<1, 1, 0, c->stream>>>(c->ts, t, t_end, nsteps);
  c->launches++;
  // SolverBase::finished() is evaluated before the first iteration too
  if (t >= t_end - 1e-14)
    nsteps = 0;
  int s = 0;
  if (graph_wanted(*c, nsteps)) {
    // the first step runs as plain launches: it primes the CFL reduction and leaves the flags in their steady state
    int rc = enqueue_step(*c, 0, 0.0);
    if (rc != EK_OK)
      return rc;
    c->cur = 1 - c->cur;
    s = 1;
    const int par = c->cur;
    if (!c->step_graph[par] && capture_step_pair(*c) != EK_OK) {
      c->step_graph[par] = nullptr; // not capturable here: the plain loop below takes over
      cudaGetLastError();
    }
    if (c->step_graph[par]) {
      for (; nsteps - s >= 2; s += 2) {
        EK_CUDA(c, cudaGraphLaunch(c->step_graph[par], c->stream));
        c->launches += c->graph_kernels[par];
      }
    }
  }
  for (; s < nsteps; ++s) {
    int rc = enqueue_step(*c, 0, 0.0);
    if (rc != EK_OK)
      return rc;
    c->cur = 1 - c->cur; // provisional; corrected below if the loop finished early on the device
  }
  if (c->halo_ready_for)
    EK_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
  EK_CUDA(c, cudaMemcpyAsync(c->ts_host, c->ts, sizeof(TimeState), cudaMemcpyDeviceToHost, c->stream));
  EK_CUDA(c, cudaStreamSynchronize(c->stream));
  const int done = c->ts_host->steps_done;
  if (((nsteps - done) & 1) != 0)
    c->cur = 1 - c->cur; // skipped steps did not flip the levels
  if (done < nsteps) {
    c->dt_primed = false; // skipped steps consumed nothing, but be conservative
    c->halo_ready_for = nullptr;
    c->mhd_dt_partial_for = nullptr;
  }
  if (t_out) *t_out = c->ts_host->t;
  if (dt_last_out) *dt_last_out = c->ts_host->dt;
  if (steps_done_out) *steps_done_out = done;
  return EK_OK;
}

int ek_kernel_timing(ek_ctx* ctx, int enable)
{
  if (!ctx) return EK_ERR_INVALID;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  for (cudaEvent_t e : c->kev) cudaEventDestroy(e);
  c->kev.clear();
  c->ktiming = enable != 0;
  return EK_OK;
}

int ek_kernel_time(ek_ctx* ctx, double* total_ms_out, int* launches_out)
{
  if (!ctx || !total_ms_out || !launches_out) return EK_ERR_INVALID;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  EK_CUDA(c, cudaSetDevice(c->device));
  EK_CUDA(c, cudaStreamSynchronize(c->stream));
  double tot = 0.0;
  int n = 0;
  for (size_t i = 0; i + 1 < c->kev.size(); i += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->kev[i], c->kev[i + 1]) == cudaSuccess) {
      tot += ms;
      ++n;
    }
  }
  for (cudaEvent_t e : c->kev) cudaEventDestroy(e);
  c->kev.clear();
  *total_ms_out = tot;
  *launches_out = n;
  return EK_OK;
}

int ek_conserved_sums(ek_ctx* ctx, double* sums_out)
{
  if (!ctx || !sums_out) return EK_ERR_INVALID;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  EK_CUDA(c, cudaSetDevice(c->device));
  conserved_sums_kernel<<<dim3(SUMS_BLOCKS, c->g.nbvar), 256, 0, c->stream>>>(c->g, c->U[c->cur], c->scratch + 64);
  conserved_sums_final_kernel<<<c->g.nbvar, 32, 0, c->stream>>>(c->scratch + 64, SUMS_BLOCKS, c->scratch);
  c->launches += 2;
  int rc = comm_allreduce_sum(*c, c->scratch, c->g.nbvar);
  if (rc != EK_OK) return rc;
  EK_CUDA(c, cudaMemcpyAsync(sums_out, c->scratch, c->g.nbvar * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  EK_CUDA(c, cudaStreamSynchronize(c->stream));
  return EK_OK;
}

int ek_selftest_math(int device, const double* x_host, int n, double* out_host)
{
  if (!x_host || !out_host || n <= 0) return EK_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    g_create_err = "no CUDA device; this library has no CPU fallback";
    return EK_ERR_CUDA;
  }
  int prev = -1;
  cudaGetDevice(&prev);
  double *dx = nullptr, *dout = nullptr;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaMalloc(&dx, (size_t)n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dout, (size_t)6 * n * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(dx, x_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    selftest_math_kernel<<<(n + 127) / 128, 128>>>(dx, n, dout);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out_host, dout, (size_t)6 * n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(dx);
  cudaFree(dout);
  if (prev >= 0) cudaSetDevice(prev);
  if (e != cudaSuccess) {
    g_create_err = std::string("ek_selftest_math: ") + cudaGetErrorString(e);
    return EK_ERR_CUDA;
  }
  return EK_OK;
}

int ek_max_divb(ek_ctx* ctx, double* out)
{
  if (!ctx || !out) return EK_ERR_INVALID;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c->p.mhd) { c->err = "ek_max_divb: MHD only"; return EK_ERR_STATE; }
  EK_CUDA(c, cudaSetDevice(c->device));
  EK_CUDA(c, cudaMemsetAsync(c->scratch, 0, sizeof(double), c->stream));
  const long long ncell = (long long)c->g.nx * c->g.ny * c->g.nz;
  max_divb_kernel<<<blocks_for(ncell, c->sm_count), 256, 0, c->stream>>>(c->g, c->U[c->cur], reinterpret_cast<unsigned long long*>(c->scratch));
  c->launches++;
  int rc = comm_allreduce_max(*c, c->scratch, 1);
  if (rc != EK_OK) return rc;
  EK_CUDA(c, cudaMemcpyAsync(out, c->scratch, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  EK_CUDA(c, cudaStreamSynchronize(c->stream));
  return EK_OK;
}

} // extern "C"
