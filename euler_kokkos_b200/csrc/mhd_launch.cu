// Host-side dispatch of the CT-MHD step (MHD_Muscl_3D).
#include <cstdlib>
#include "ek_internal.h"
#include "mhd_kernels.cuh"

namespace ek {

size_t mhd_work_bytes(const Grid& g)
{
  return (size_t)(g.ksize > 1 ? W_TOTAL : W2_TOTAL) * (size_t)mhd_padded_cells(g.sv) * sizeof(double);
}

static MhdArgs make_args(Ctx& c, const double* Uin, double* Uout)
{
  MhdArgs a;
  a.g = c.g;
  a.Uin = Uin;
  a.Uout = Uout;
  a.work = c.work;
  a.ts = c.ts;
  // rows per slab of the y-chunked launch order (mhd_cells.cuh); EK_MHD_YCHUNK=0 restores the flat order
  const int ychunk = getenv("EK_MHD_YCHUNK") ? atoi(getenv("EK_MHD_YCHUNK")) : 16;
  a.cm = mhd_chunk_map(c.g, c.g.ksize > 1 ? ychunk : 0, 1);
  return a;
}

// the same arguments with consecutive warps of a CTA of `threads` threads mapped to consecutive planes
static MhdArgs zgrouped(const Ctx& c, MhdArgs a, int threads)
{
  const int zg = getenv("EK_MHD_ZGROUP") ? atoi(getenv("EK_MHD_ZGROUP")) : 0; // 0: one plane per CTA (measured: no gain from L1 sharing between planes)
  if (zg && a.cm.yc)
    a.cm = mhd_chunk_map(c.g, (int)a.cm.yc, threads / 32);
  return a;
}

// fused P+E+T kernel (mhd_pet.cuh); EK_MHD_PET=0 selects the three staged kernels
static bool use_pet(const Ctx& c)
{
  const int pet = getenv("EK_MHD_PET") ? atoi(getenv("EK_MHD_PET")) : 1;
  return pet != 0 && c.p.dim == 3;
}

// CFL reduction of the next step inside the update kernel (3-D); EK_MHD_DTFOLD=0 restores the stand-alone pass
bool mhd_dt_folded(const Ctx& c)
{
  const int v = getenv("EK_MHD_DTFOLD") ? atoi(getenv("EK_MHD_DTFOLD")) : 1;
  return v != 0 && c.p.dim == 3;
}

static void launch_update(const Ctx& c, dim3 grid, const MhdArgs& a)
{
  if (mhd_dt_folded(c))
    mhd_update_kernel<true><<<grid, 256, 0, c.stream>>>(a);
  else
    mhd_update_kernel<false><<<grid, 256, 0, c.stream>>>(a);
}

static unsigned blocks(const Grid& g, int threads) { return (unsigned)((g.sv + threads - 1) / threads); }
static unsigned blocks(const MhdArgs& a, int threads)
{
  return (unsigned)((mhd_chunk_threads(a.g, a.cm) + (unsigned)threads - 1) / (unsigned)threads);
}

// primitives of the (ghost-filled) level + max inverse dt into ts->invdt_bits
cudaError_t launch_mhd_invdt(Ctx& c, const double* U)
{
  const MhdArgs a = make_args(c, U, nullptr);
  if (c.p.dim == 2)
    mhd2d_prims_kernel<<<blocks(c.g, 256), 256, 0, c.stream>>>(a, 1);
  else
    mhd_prims_kernel<<<blocks(c.g, 256), 256, 0, c.stream>>>(a, 1, use_pet(c) ? 0 : 1);
  c.launches++;
  c.mhd_prims_of = U;
  return cudaGetLastError();
}

// CFL contribution of the last interior layers of the (ghost-filled) level U, on top of what the update kernel that
// produced U has accumulated in ts->invdt_bits
cudaError_t launch_mhd_invdt_layers(Ctx& c, const double* U)
{
  const MhdArgs a = make_args(c, U, nullptr);
  const Grid& g = c.g;
  const unsigned n = (unsigned)g.ny * g.nz + (unsigned)g.nx * g.nz + (unsigned)g.nx * g.ny;
  mhd_invdt_layers_kernel<<<(n + 255) / 256, 256, 0, c.stream>>>(a);
  c.launches++;
  return cudaGetLastError();
}

// stage U restricted to planes: the END planes (first and last gw interior planes, what the z neighbours need as
// ghosts) as two small flat launches, or the MIDDLE planes with the regular launch order
static void launch_update_ends(Ctx& c, MhdArgs a)
{
  const Grid& g = c.g;
  a.cm = mhd_chunk_map(g, 0, 1); // flat: thread = cell
  const unsigned per_plane = (unsigned)g.sk;
  const int k0[2] = { g.gw, g.nz };
  for (int e = 0; e < 2; ++e) {
    a.t0 = per_plane * (unsigned)k0[e];
    a.upd_klo = k0[e];
    a.upd_khi = k0[e] + g.gw;
    const unsigned nthreads = per_plane * (unsigned)g.gw;
    launch_update(c, dim3((nthreads + 255) / 256), a);
    c.launches++;
  }
}

// part: 0 = the whole step; 1 = every stage, but the update only on the END planes; 2 = the update of the MIDDLE planes
template <int RS>
static cudaError_t launch_step(Ctx& c, const double* Uin, double* Uout, int part)
{
  const MhdArgs a = make_args(c, Uin, Uout);
  if (part == 2) {
    MhdArgs m = zgrouped(c, a, 256);
    m.upd_klo = 2 * c.g.gw;
    m.upd_khi = c.g.nz;
    launch_update(c, dim3(blocks(m, 256)), m);
    c.launches++;
    return cudaGetLastError();
  }
  const MhdArgs a8 = zgrouped(c, a, 256), a4 = zgrouped(c, a, 128);
  if (use_pet(c)) {
    static bool pet_attr[16] = { false };
    if (!pet_attr[c.device & 15]) {
      cudaError_t e = cudaFuncSetAttribute(mhd_pet_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PetSmem));
      if (e != cudaSuccess)
        return e;
      pet_attr[c.device & 15] = true;
    }
    const Grid& g = c.g;
    dim3 grid((g.isize - 3 + PET_TX - 1) / PET_TX, (g.jsize - 3 + PET_TY - 1) / PET_TY, 1); // cells 2 .. isize-2
    const int nplanes = g.ksize - 3;                                                          // planes 2 .. ksize-2
    // enough CTAs for whole waves (2 resident CTAs per SM), chunks long enough to amortise the 3-plane warm-up
    const long long slots = 2LL * c.sm_count, xy = (long long)grid.x * grid.y;
    int nchunk = 1;
    double best = 1e300;
    for (int n = 1; n <= nplanes; ++n) {
      const int zc = (nplanes + n - 1) / n;
      if (zc < 16 && n > 1)
        break;
      const long long ctas = xy * ((nplanes + zc - 1) / zc);
      const double cost = (double)(((ctas + slots - 1) / slots) * slots) / (double)ctas * (zc + 3.0) / zc;
      if (cost < best - 1e-12) {
        best = cost;
        nchunk = n;
      }
    }
    const int zchunk = (nplanes + nchunk - 1) / nchunk;
    grid.z = (nplanes + zchunk - 1) / zchunk;
    static const int pet_pf = getenv("EK_MHD_PET_PF") ? atoi(getenv("EK_MHD_PET_PF")) : 0; // L2 prefetch of the plane converted next
    mhd_pet_kernel<0><<<grid, PET_TX * PET_TY, sizeof(PetSmem), c.stream>>>(a, zchunk, pet_pf);
    c.launches++;
    c.mhd_prims_of = nullptr;
  } else {
    if (c.mhd_prims_of != Uin) { // dt given by the host: the primitives have not been computed yet
      mhd_prims_kernel<<<blocks(c.g, 256), 256, 0, c.stream>>>(a, 0);
      c.launches++;
    }
    c.mhd_prims_of = nullptr;
    mhd_elec_kernel<<<blocks(a8, 256), 256, 0, c.stream>>>(a8);
    mhd_trace_kernel<<<blocks(a8, 256), 256, 0, c.stream>>>(a8);
    c.launches += 2;
  }
  const int fuse_r = getenv("EK_MHD_FUSE_R") ? atoi(getenv("EK_MHD_FUSE_R")) : 1; // tuning switch (default: fused)
  // stage R variant: 2 (default) = plain loads + Riemann solvers as real functions; 1 = cp.async software pipeline;
  // 0 = fully inlined (mhd_staged.cuh has the measurements)
  const int staged = getenv("EK_MHD_STAGED") ? atoi(getenv("EK_MHD_STAGED")) : 2;
  if (staged == 2) {
    MhdArgs ar = a4;
    ar.r_pf = getenv("EK_MHD_R_PF") ? atoi(getenv("EK_MHD_R_PF")) : 0; // planes ahead (y-chunked order only)
    if (!ar.cm.yc)
      ar.r_pf = 0;
    mhd_faces_edges_call_kernel<RS><<<blocks(ar, 128), 128, 0, c.stream>>>(ar);
    c.launches++;
  } else if (staged) {
    static bool attr_done[16] = { false };
    if (!attr_done[c.device & 15]) {
      cudaError_t e = cudaFuncSetAttribute(mhd_faces_edges_staged_kernel<RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STG_SMEM_BYTES);
      if (e != cudaSuccess)
        return e;
      attr_done[c.device & 15] = true;
    }
    mhd_faces_edges_staged_kernel<RS><<<blocks(a4, STG_THREADS), STG_THREADS, STG_SMEM_BYTES, c.stream>>>(a4);
    c.launches++;
  } else if (fuse_r) {
    mhd_faces_edges_kernel<RS><<<blocks(a4, 128), 128, 0, c.stream>>>(a4);
    c.launches++;
  } else {
    mhd_flux_kernel<RS><<<blocks(a4, 128), 128, 0, c.stream>>>(a4);
    mhd_emf_kernel<<<blocks(a4, 128), 128, 0, c.stream>>>(a4);
    c.launches += 2;
  }
  if (part == 1) {
    launch_update_ends(c, a);
  } else {
    launch_update(c, dim3(blocks(a8, 256)), a8);
    c.launches++;
  }
  return cudaGetLastError();
}

template <int RS>
static cudaError_t launch_step_2d(Ctx& c, const double* Uin, double* Uout)
{
  const MhdArgs a = make_args(c, Uin, Uout);
  if (c.mhd_prims_of != Uin) {
    mhd2d_prims_kernel<<<blocks(c.g, 256), 256, 0, c.stream>>>(a, 0);
    c.launches++;
  }
  c.mhd_prims_of = nullptr;
  mhd2d_elec_kernel<<<blocks(c.g, 256), 256, 0, c.stream>>>(a);
  mhd2d_trace_kernel<<<blocks(c.g, 256), 256, 0, c.stream>>>(a);
  mhd2d_faces_edge_kernel<RS><<<blocks(c.g, 128), 128, 0, c.stream>>>(a);
  mhd2d_update_kernel<<<blocks(c.g, 256), 256, 0, c.stream>>>(a);
  c.launches += 4;
  return cudaGetLastError();
}

cudaError_t launch_mhd_step(Ctx& c, const double* Uin, double* Uout, int part)
{
  if (c.p.dim == 2) {
    switch (c.p.riemann) {
      case RIEMANN_HLLD: return launch_step_2d<RIEMANN_HLLD>(c, Uin, Uout);
      case RIEMANN_HLL: return launch_step_2d<RIEMANN_HLL>(c, Uin, Uout);
      case RIEMANN_LLF: return launch_step_2d<RIEMANN_LLF>(c, Uin, Uout);
      default: return cudaErrorInvalidValue;
    }
  }
  switch (c.p.riemann) {
    case RIEMANN_HLLD: return launch_step<RIEMANN_HLLD>(c, Uin, Uout, part);
    case RIEMANN_HLL: return launch_step<RIEMANN_HLL>(c, Uin, Uout, part);
    case RIEMANN_LLF: return launch_step<RIEMANN_LLF>(c, Uin, Uout, part);
    default: return cudaErrorInvalidValue;
  }
}

} // namespace ek
