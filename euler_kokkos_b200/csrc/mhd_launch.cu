// Host-side dispatch of the CT-MHD step (MHD_Muscl_3D).
#include <cstdio>
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

// P+E+T kernel on the tile rows [ty0, ty1) (a tile row = PET_TY rows of cells starting at row 2 + PET_TY*ty)
static void launch_pet(Ctx& c, const MhdArgs& a, int ty0, int ty1, cudaStream_t stream)
{
  const Grid& g = c.g;
  dim3 grid((g.isize - 3 + PET_TX - 1) / PET_TX, ty1 - ty0, 1); // cells 2 .. isize-2
  const int nplanes = g.ksize - 3;                               // planes 2 .. ksize-2
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
  const int pet_pf = getenv("EK_MHD_PET_PF") ? atoi(getenv("EK_MHD_PET_PF")) : 0; // L2 prefetch of the plane converted next
  mhd_pet_kernel<0><<<grid, PET_TX * PET_TY, sizeof(PetSmem), stream>>>(a, zchunk, pet_pf, ty0);
  c.launches++;
}

// ---- slab-pipelined step (EK_MHD_PIPE = m > 0: m y-slabs per pipeline stage; opt-in) ---------------------------
// The three kernels of a step have complementary bottlenecks (ncu, 512^3): P+E+T is latency bound (FP64 pipe 20 %
// active, DRAM 38 %), R is FP64-issue bound (59 %, DRAM 26 %), U is DRAM bound.  Launched back to back each has the
// GPU for itself.  Here the box is cut in stages of m y-slabs (the unit of the y-chunked launch order) and every
// kernel is launched once per stage on its own stream, ordered by events only where data flows:
//   R of stage g   reads S of rows up to the last row of the stage  -> after P+E+T of the tile rows covering them
//   U of stage g   reads F / EMF of rows up to the first row of stage g+1 -> after R of stage g+1
// so P+E+T of stage g+2, R of stage g+1 and U of stage g can be resident on the SMs at the same time.
bool mhd_pipelined(const Ctx& c)
{
  const int m = getenv("EK_MHD_PIPE") ? atoi(getenv("EK_MHD_PIPE")) : 0;
  const int ychunk = getenv("EK_MHD_YCHUNK") ? atoi(getenv("EK_MHD_YCHUNK")) : 16;
  const int staged = getenv("EK_MHD_STAGED") ? atoi(getenv("EK_MHD_STAGED")) : 2;
  const int zg = getenv("EK_MHD_ZGROUP") ? atoi(getenv("EK_MHD_ZGROUP")) : 0;
  return m > 0 && c.p.dim == 3 && use_pet(c) && staged == 2 && zg == 0 && ychunk > 0 && ychunk % PET_TY == 0 &&
         ychunk < c.g.jsize;
}

void mhd_pipe_destroy(Ctx& c)
{
  for (cudaStream_t& st : c.pipe_stream) {
    if (st) {
      cudaStreamSynchronize(st);
      cudaStreamDestroy(st);
      st = nullptr;
    }
  }
  for (cudaEvent_t e : c.pipe_ev)
    cudaEventDestroy(e);
  c.pipe_ev.clear();
}

template <int RS>
static cudaError_t launch_step_pipelined(Ctx& c, const MhdArgs& a)
{
  const Grid& g = c.g;
  const int m = atoi(getenv("EK_MHD_PIPE"));
  const unsigned yc = a.cm.yc, nslab = a.cm.nchunk;
  const int nstage = (int)((nslab + (unsigned)m - 1) / (unsigned)m);
  const int ntr = (g.jsize - 3 + PET_TY - 1) / PET_TY; // tile rows of the P+E+T kernel
  // streams and events (created once per context)
  if (!c.pipe_stream[0]) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    int pr[3] = { 0, 0, 0 }; // EK_MHD_PIPE_PRIO="p,r,u": 0 = default priority, k > 0 = k levels above it
    if (const char* e = getenv("EK_MHD_PIPE_PRIO"))
      sscanf(e, "%d,%d,%d", &pr[0], &pr[1], &pr[2]);
    const int which[4] = { 0, 0, 1, 2 };
    for (int i = 0; i < 4; ++i) {
      int p = lo - pr[which[i]];
      if (p < hi)
        p = hi;
      cudaError_t e = cudaStreamCreateWithPriority(&c.pipe_stream[i], cudaStreamNonBlocking, p);
      if (e != cudaSuccess)
        return e;
    }
  }
  const size_t nev = 2 * (size_t)nstage + 4;
  while (c.pipe_ev.size() < nev) {
    cudaEvent_t e;
    cudaError_t rc = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    if (rc != cudaSuccess)
      return rc;
    c.pipe_ev.push_back(e);
  }
  cudaEvent_t ev_start = c.pipe_ev[0], ev_endp = c.pipe_ev[1], ev_endr = c.pipe_ev[2], ev_endu = c.pipe_ev[3];
  cudaEvent_t* evP = c.pipe_ev.data() + 4;
  cudaEvent_t* evR = evP + nstage;
  cudaStream_t sP[2] = { c.pipe_stream[0], c.pipe_stream[1] }, sR = c.pipe_stream[2], sU = c.pipe_stream[3];

  cudaEventRecord(ev_start, c.stream);
  for (cudaStream_t st : c.pipe_stream)
    cudaStreamWaitEvent(st, ev_start, 0);

  MhdArgs ar = a, au = a; // same y-chunked map (kg = 1) for R and U
  const unsigned per_slab = a.cm.span * a.cm.kpad;
  int ty_done = 0;
  // P+E+T stages alternate between two streams (they do not depend on each other); R and U stages run in order
  for (int gI = 0; gI < nstage; ++gI) {
    const unsigned s1 = ((unsigned)(gI + 1) * (unsigned)m < nslab) ? (unsigned)(gI + 1) * (unsigned)m : nslab;
    // R of the slabs [s0, s1) reads S of rows <= yc*s1 - 1 (and >= yc*s0 - 1, produced by earlier stages)
    int ty_end = (gI == nstage - 1) ? ntr : (int)((yc * s1 - 3u) / (unsigned)PET_TY) + 1;
    if (ty_end > ntr)
      ty_end = ntr;
    cudaStream_t sp = sP[gI & 1];
    if (ty_end > ty_done) {
      launch_pet(c, a, ty_done, ty_end, sp);
      ty_done = ty_end;
    }
    cudaEventRecord(evP[gI], sp);
  }
  for (int gI = 0; gI < nstage; ++gI) {
    const unsigned s0 = (unsigned)gI * (unsigned)m;
    const unsigned s1 = (s0 + (unsigned)m < nslab) ? s0 + (unsigned)m : nslab;
    cudaStreamWaitEvent(sR, evP[gI], 0);
    if (gI > 0)
      cudaStreamWaitEvent(sR, evP[gI - 1], 0); // the other P+E+T stream: every earlier stage of the other parity
    ar.t0 = s0 * per_slab;
    ar.t1 = s1 * per_slab;
    const unsigned nthreads = ar.t1 - ar.t0;
    mhd_faces_edges_call_kernel<RS><<<(nthreads + 127) / 128, 128, 0, sR>>>(ar);
    c.launches++;
    cudaEventRecord(evR[gI], sR);
  }
  for (int gI = 0; gI < nstage; ++gI) {
    const unsigned s0 = (unsigned)gI * (unsigned)m;
    const unsigned s1 = (s0 + (unsigned)m < nslab) ? s0 + (unsigned)m : nslab;
    cudaStreamWaitEvent(sU, evR[(gI + 1 < nstage) ? gI + 1 : gI], 0);
    au.t0 = s0 * per_slab;
    au.t1 = s1 * per_slab;
    const unsigned nthreads = au.t1 - au.t0;
    if (mhd_dt_folded(c))
      mhd_update_kernel<true><<<(nthreads + 255) / 256, 256, 0, sU>>>(au);
    else
      mhd_update_kernel<false><<<(nthreads + 255) / 256, 256, 0, sU>>>(au);
    c.launches++;
  }
  cudaEventRecord(ev_endp, sP[0]);
  cudaEventRecord(ev_endr, sP[1]);
  cudaEventRecord(ev_endu, sU); // U is last in its stream and waited on every R stage
  cudaStreamWaitEvent(c.stream, ev_endp, 0);
  cudaStreamWaitEvent(c.stream, ev_endr, 0);
  cudaStreamWaitEvent(c.stream, ev_endu, 0);
  c.mhd_prims_of = nullptr;
  return cudaGetLastError();
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
    if (part == 0 && mhd_pipelined(c))
      return launch_step_pipelined<RS>(c, a);
    launch_pet(c, a, 0, (c.g.jsize - 3 + PET_TY - 1) / PET_TY, c.stream);
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
