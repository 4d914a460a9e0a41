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
  return a;
}

static unsigned blocks(const Grid& g, int threads) { return (unsigned)((g.sv + threads - 1) / threads); }

// primitives of the (ghost-filled) level + max inverse dt into ts->invdt_bits
cudaError_t launch_mhd_invdt(Ctx& c, const double* U)
{
  const MhdArgs a = make_args(c, U, nullptr);
  if (c.p.dim == 2)
    mhd2d_prims_kernel<<<blocks(c.g, 256), 256, 0, c.stream>>>(a, 1);
  else
    mhd_prims_kernel<<<blocks(c.g, 256), 256, 0, c.stream>>>(a, 1);
  c.launches++;
  c.mhd_prims_of = U;
  return cudaGetLastError();
}

template <int RS>
static cudaError_t launch_step(Ctx& c, const double* Uin, double* Uout)
{
  const MhdArgs a = make_args(c, Uin, Uout);
  if (c.mhd_prims_of != Uin) { // dt given by the host: the primitives have not been computed yet
    mhd_prims_kernel<<<blocks(c.g, 256), 256, 0, c.stream>>>(a, 0);
    c.launches++;
  }
  c.mhd_prims_of = nullptr;
  mhd_elec_kernel<<<blocks(c.g, 256), 256, 0, c.stream>>>(a);
  mhd_trace_kernel<<<blocks(c.g, 256), 256, 0, c.stream>>>(a);
  static const int fuse_r = getenv("EK_MHD_FUSE_R") ? atoi(getenv("EK_MHD_FUSE_R")) : 1; // tuning switch (default: fused)
  if (fuse_r) {
    mhd_faces_edges_kernel<RS><<<blocks(c.g, 128), 128, 0, c.stream>>>(a);
    c.launches += 4;
  } else {
    mhd_flux_kernel<RS><<<blocks(c.g, 128), 128, 0, c.stream>>>(a);
    mhd_emf_kernel<<<blocks(c.g, 128), 128, 0, c.stream>>>(a);
    c.launches += 5;
  }
  mhd_update_kernel<<<blocks(c.g, 256), 256, 0, c.stream>>>(a);
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

cudaError_t launch_mhd_step(Ctx& c, const double* Uin, double* Uout)
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
    case RIEMANN_HLLD: return launch_step<RIEMANN_HLLD>(c, Uin, Uout);
    case RIEMANN_HLL: return launch_step<RIEMANN_HLL>(c, Uin, Uout);
    case RIEMANN_LLF: return launch_step<RIEMANN_LLF>(c, Uin, Uout);
    default: return cudaErrorInvalidValue;
  }
}

} // namespace ek
