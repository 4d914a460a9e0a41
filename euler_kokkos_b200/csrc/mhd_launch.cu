// Host-side dispatch of the CT-MHD step (MHD_Muscl_3D).
#include <cstdio>
#include <cstdlib>
#include "ek_internal.h"
#include "mhd_kernels.cuh"

#ifndef EK_MHD_R_PFL2_DEFAULT
#define EK_MHD_R_PFL2_DEFAULT 2 // stage R: lead (in planes) of the bulk L2 prefetch of S; -1 = off
#endif

namespace ek {

size_t mhd_work_bytes(const Grid& g)
{
  return (size_t)(g.ksize > 1 ? W3_DEVICE : W2_TOTAL) * (size_t)mhd_padded_cells(g.sv) * sizeof(double);
}

// 4-D tensor maps (x, y, z, variable) over both time levels, box = 36 x (tile rows + 3) x 1 x 1 doubles, no swizzle, zero fill
// outside the array.  The encoder is a driver entry point: looked up at run time, libcuda is not linked.
bool mhd_make_tensor_maps(Ctx& c)
{
  c.umap_ok = false;
  const Grid& g = c.g;
  if (c.p.dim != 3 || !c.p.mhd || (g.isize & 1))
    return false; // global strides must be multiples of 16 bytes
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
      qres != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return false;
  }
  const cuuint64_t dims[4] = { (cuuint64_t)g.isize, (cuuint64_t)g.jsize, (cuuint64_t)g.ksize, 8 };
  const cuuint64_t strides[3] = { (cuuint64_t)g.isize * 8, (cuuint64_t)g.sk * 8, (cuuint64_t)g.sv * 8 };
  const cuuint32_t estr[4] = { 1, 1, 1, 1 };
  // the box shape belongs to the map: one variable of one plane per instruction (9 instructions per staged plane)
  const cuuint32_t box[4] = { (cuuint32_t)PT_RX, (cuuint32_t)PT_RY, 1, 1 };
  for (int l = 0; l < 2; ++l) {
    CUresult r = ((EncodeFn)fn)(&c.umap[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, c.U[l], dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return false;
  }
  c.umap_ok = true;
  return true;
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
  a.cm = mhd_chunk_map(c.g, c.g.ksize > 1 ? ychunk : 0);
  return a;
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

// max inverse dt of the (ghost-filled) level into ts->invdt_bits
cudaError_t launch_mhd_invdt(Ctx& c, const double* U)
{
  const MhdArgs a = make_args(c, U, nullptr);
  if (c.p.dim == 2)
    mhd2d_prims_kernel<<<blocks(c.g, 256), 256, 0, c.stream>>>(a, 1);
  else
    mhd_invdt_kernel<<<blocks(c.g, 256), 256, 0, c.stream>>>(a);
  c.launches++;
  c.mhd_prims_of = U; // 2-D: the primitives now sit in the work area
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
  a.cm = mhd_chunk_map(g, 0); // flat: thread = cell
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

// z-chunking of a z-marching kernel: enough CTAs for whole waves over `slots` resident CTAs, chunks long enough to
// amortise the 3-plane warm-up
static int pick_pet_zchunk(int nplanes, long long xy, long long slots)
{
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
  return (nplanes + nchunk - 1) / nchunk;
}

// TMA-fed P+E+T kernel (mhd_pet_tma.cuh) when the grid allows it; EK_MHD_PET_TMA=0 forces the plain-load kernel
static bool use_pet_tma(const Ctx& c)
{
  const int v = getenv("EK_MHD_PET_TMA") ? atoi(getenv("EK_MHD_PET_TMA")) : 1;
  return v != 0 && c.umap_ok;
}

// stages P+E+T over the whole box
static cudaError_t launch_pet(Ctx& c, const MhdArgs& a)
{
  const Grid& g = c.g;
  const int nplanes = g.ksize - 3; // planes 2 .. ksize-2
  if (use_pet_tma(c)) {
    if (!c.mhd_pet_tma_attr) {
      cudaError_t e = cudaFuncSetAttribute(mhd_pet_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PetTmaSmem));
      if (e != cudaSuccess)
        return e;
      c.mhd_pet_tma_attr = true;
    }
    dim3 grid((g.isize - 3 + PT_TX - 1) / PT_TX, (g.jsize - 3 + PT_TY - 1) / PT_TY, 1); // cells 2 .. isize-2
    const int zchunk = pick_pet_zchunk(nplanes, (long long)grid.x * grid.y, (long long)PT_MINB * c.sm_count);
    grid.z = (nplanes + zchunk - 1) / zchunk;
    const int level = (a.Uin == c.U[0]) ? 0 : 1;
    mhd_pet_tma_kernel<<<grid, PT_TX * PT_TY, sizeof(PetTmaSmem), c.stream>>>(a, zchunk, c.umap[level]);
    c.launches++;
    return cudaGetLastError();
  }
  if (!c.mhd_pet_attr) {
    cudaError_t e = cudaFuncSetAttribute(mhd_pet_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PetSmem));
    if (e != cudaSuccess)
      return e;
    c.mhd_pet_attr = true;
  }
  dim3 grid((g.isize - 3 + PET_TX - 1) / PET_TX, (g.jsize - 3 + PET_TY - 1) / PET_TY, 1);
  const int zchunk = pick_pet_zchunk(nplanes, (long long)grid.x * grid.y, 2LL * c.sm_count);
  grid.z = (nplanes + zchunk - 1) / zchunk;
  mhd_pet_kernel<0><<<grid, PET_TX * PET_TY, sizeof(PetSmem), c.stream>>>(a, zchunk);
  c.launches++;
  return cudaGetLastError();
}

// ---- fused R+U kernel (mhd_ru.cuh): z-chunking ----------------------------------------------------------------------
// one CTA per SM; a chunk of zc planes solves zc+1 (the Riemann problems of the plane above its last one)
struct RuPlan {
  int zchunk, nchunk;
  bool splittable; // >= 3 chunks whose end chunks hold the gw planes the z neighbours need
};
static int ru_pick_zchunk(int nz, long long xy, int sm_count, int min_chunks, int min_len)
{
  int best = nz;
  double best_cost = 1e300;
  for (int n = min_chunks; n <= nz; ++n) {
    const int zc = (nz + n - 1) / n;
    if (zc < min_len && n > min_chunks)
      break;
    const long long ctas = xy * ((nz + zc - 1) / zc);
    const long long waves = (ctas + sm_count - 1) / sm_count;
    const double cost = (double)(waves * sm_count) / (double)ctas * (zc + 1.0) / zc;
    if (cost < best_cost - 1e-12) {
      best_cost = cost;
      best = zc;
    }
  }
  return best;
}
static RuPlan ru_plan(const Ctx& c, bool want_split)
{
  const Grid& g = c.g;
  const long long xy = (long long)((g.nx + RU_TX - 2) / (RU_TX - 1)) * ((g.ny + RU_TY - 2) / (RU_TY - 1));
  RuPlan z;
  z.zchunk = ru_pick_zchunk(g.nz, xy, c.sm_count, 1, 8);
  z.nchunk = (g.nz + z.zchunk - 1) / z.zchunk;
  z.splittable = false;
  if (!want_split)
    return z;
  const int zc = (z.nchunk >= 3) ? z.zchunk : ru_pick_zchunk(g.nz, xy, c.sm_count, 3, g.gw);
  const int nc = (g.nz + zc - 1) / zc;
  const int last = g.nz - (nc - 1) * zc;
  if (nc >= 3 && zc >= g.gw && last >= g.gw) {
    z.zchunk = zc;
    z.nchunk = nc;
    z.splittable = true;
  }
  return z;
}

// EK_MHD_RU=1 selects the fused R+U kernel.  Default: the two-kernel form (stage R -> F / EMF arrays -> stage U).
// Measured at 512^3 (profiles/r02_mhd_ru_ncu.txt): fused 50.1 ms against 28.6 + 7.8 ms.  Its 147 KB of shared-memory
// planes leave stage R's ~290 loads per thread a third of the L1 they have in the flat kernel (L1 hit rate 41 % -> 22 %,
// long-scoreboard stalls 1.8 -> 3.9 per issue, DRAM reads 44 -> 59 GB): what the fusion saves in F / EMF traffic (48 GB)
// it loses several times over in latency of the reconstruction loads.
static bool use_ru(const Ctx& c)
{
  const int v = getenv("EK_MHD_RU") ? atoi(getenv("EK_MHD_RU")) : 0;
  return v != 0 && c.p.dim == 3;
}
bool mhd_z_splittable(const Ctx& c) { return use_ru(c) ? ru_plan(c, true).splittable : (c.p.nz >= 2 * c.p.ghost); }

// part: 0 = every chunk, 1 = the two end chunks, 2 = the middle chunks
template <int RS>
static cudaError_t launch_ru(Ctx& c, const MhdArgs& a, int part)
{
  const bool fold = mhd_dt_folded(c);
  void (*kern)(const MhdRuArgs) = fold ? mhd_ru_kernel<RS, true> : mhd_ru_kernel<RS, false>;
  if (c.mhd_ru_attr_for != (const void*)kern) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RuSmem));
    if (e != cudaSuccess)
      return e;
    c.mhd_ru_attr_for = (const void*)kern;
  }
  const Grid& g = c.g;
  const RuPlan z = ru_plan(c, part != 0);
  if (part != 0 && !z.splittable)
    return cudaErrorInvalidValue;
  MhdRuArgs ra;
  ra.a = a;
  ra.zchunk = z.zchunk;
  ra.zfirst = 0;
  ra.zstep = 1;
  dim3 grid((g.nx + RU_TX - 2) / (RU_TX - 1), (g.ny + RU_TY - 2) / (RU_TY - 1), z.nchunk);
  if (part == 1) {
    ra.zstep = z.nchunk - 1;
    grid.z = 2;
  } else if (part == 2) {
    ra.zfirst = 1;
    grid.z = z.nchunk - 2;
  }
  kern<<<grid, dim3(RU_TX, RU_TY, 1), sizeof(RuSmem), c.stream>>>(ra);
  c.launches++;
  return cudaGetLastError();
}

// One 3-D step = P+E+T -> S, then R+U fused (or R -> F / EMF arrays, then U).
// part: 0 = the whole step; 1 = every stage, but the update only on the END planes (chunks); 2 = the update of the rest
template <int RS>
static cudaError_t launch_step(Ctx& c, const double* Uin, double* Uout, int part)
{
  const MhdArgs a = make_args(c, Uin, Uout);
  if (use_ru(c)) {
    if (part != 2) {
      cudaError_t e = launch_pet(c, a);
      if (e != cudaSuccess)
        return e;
    }
    return launch_ru<RS>(c, a, part);
  }
  if (part == 2) {
    MhdArgs m = a;
    m.upd_klo = 2 * c.g.gw;
    m.upd_khi = c.g.nz;
    launch_update(c, dim3(blocks(m, 256)), m);
    c.launches++;
    return cudaGetLastError();
  }
  cudaError_t e = launch_pet(c, a);
  if (e != cudaSuccess)
    return e;
  {
    // EK_MHD_R_PFL2: every warp of stage R asks L2 for the S tile of the warp this many planes further along z
    // (mhd_stage_r.cuh); meaningful with the y-chunked launch order only (flat: the next plane is a whole plane of CTAs away)
    MhdArgs ar = a;
    ar.pf_planes = getenv("EK_MHD_R_PFL2") ? atoi(getenv("EK_MHD_R_PFL2")) : EK_MHD_R_PFL2_DEFAULT;
    if (ar.cm.yc == 0)
      ar.pf_planes = -1;
    mhd_faces_edges_call_kernel<RS><<<blocks(a, 128), 128, 0, c.stream>>>(ar);
  }
  c.launches++;
  if (part == 1) {
    launch_update_ends(c, a);
  } else {
    launch_update(c, dim3(blocks(a, 256)), a);
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
