// Multi-GPU plumbing: ghost-layer exchange between neighbouring sub-domains and the cross-rank CFL
// reduction, over NCCL (NVLink 5 / NVSwitch).  Replaces the reference's MPI path:
//   CopyDataArray_To_BorderBuf / CopyBorderBuf_To_DataArray   shared/mpiBorderUtils.h:36-330
//   SolverBase::transfer_boundaries_{2d,3d} (MPI_Sendrecv)    shared/SolverBase.cpp:756-845
//   MPI_Allreduce<MIN>(dt)                                     shared/SolverBase.cpp:134-138
// NCCL is loaded lazily with dlopen so that single-GPU use has no NCCL dependency.
//
// Along the slowest dimension (z in 3-D, y in 2-D -- the default slab decomposition) a ghost slab of
// one variable is one contiguous block of the SoA state, so it is sent straight from / received
// straight into the state array: no pack/unpack kernels, no staging buffers.  Other directions go
// through small packed buffers.
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include "ek_internal.h"

namespace ek {

// minimal NCCL surface (types as in nccl.h)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclDouble = 8 };
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };

struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string err;
};

static NcclApi* nccl_api_load(NcclApi& api)
{
  const char* names[] = { "libnccl.so.2", "libnccl.so" };
  for (const char* n : names) {
    api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.h)
      break;
  }
  if (!api.h) {
    api.err = std::string("dlopen(libnccl.so.2): ") + dlerror();
    return nullptr;
  }
#define LD(field, sym)                                                      \
  do {                                                                      \
    *(void**)(&api.field) = dlsym(api.h, sym);                              \
    if (!api.field) {                                                       \
      api.err = std::string("dlsym ") + sym + " failed";                   \
      dlclose(api.h);                                                       \
      api.h = nullptr;                                                      \
      return nullptr;                                                       \
    }                                                                       \
  } while (0)
  LD(GetUniqueId, "ncclGetUniqueId");
  LD(CommInitRank, "ncclCommInitRank");
  LD(CommDestroy, "ncclCommDestroy");
  LD(Send, "ncclSend");
  LD(Recv, "ncclRecv");
  LD(GroupStart, "ncclGroupStart");
  LD(GroupEnd, "ncclGroupEnd");
  LD(AllReduce, "ncclAllReduce");
  LD(GetErrorString, "ncclGetErrorString");
#undef LD
  return &api;
}

// loaded once per process, also when several host threads each drive their own context
static NcclApi* nccl_api()
{
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] { nccl_api_load(api); });
  return api.h ? &api : nullptr;
}

#define EK_NCCL(c, call)                                                              \
  do {                                                                                \
    int r_ = (call);                                                                  \
    if (r_ != ncclSuccess) {                                                          \
      (c).err = std::string(#call) + ": " + api->GetErrorString(r_);                  \
      return EK_ERR_NCCL;                                                             \
    }                                                                                 \
  } while (0)

// pack gw layers [lo, lo+gw) along DIR (full transverse extent, all variables) into buf, or unpack
template <bool PACK>
__global__ void __launch_bounds__(256) halo_copy_kernel(const Grid g, double* __restrict__ U, double* __restrict__ buf, int dir, int lo)
{
  const int ext0 = (dir == 0) ? g.gw : g.isize;
  const int ext1 = (dir == 1) ? g.gw : g.jsize;
  const int ext2 = (dir == 2) ? g.gw : g.ksize;
  const long long per_var = (long long)ext0 * ext1 * ext2;
  const long long total = per_var * g.nbvar;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(idx / per_var);
    const long long r = idx - (long long)v * per_var;
    int c0 = (int)(r % ext0), c1 = (int)((r / ext0) % ext1), c2 = (int)(r / ((long long)ext0 * ext1));
    if (dir == 0) c0 += lo;
    if (dir == 1) c1 += lo;
    if (dir == 2) c2 += lo;
    const long long o = c0 + g.sj * c1 + g.sk * c2 + g.sv * v;
    if (PACK)
      buf[idx] = U[o];
    else
      U[o] = buf[idx];
  }
}

int comm_exchange_dir(Ctx& c, double* U, int dir, cudaStream_t stream)
{
  if (!stream)
    stream = c.stream;
  NcclApi* api = nccl_api();
  if (!api || !c.nccl_comm) {
    c.err = "boundary type EK_BC_COPY needs a communicator (ek_comm_attach)";
    return EK_ERR_STATE;
  }
  ncclComm_t comm = (ncclComm_t)c.nccl_comm;
  const Grid& g = c.g;
  const int gw = g.gw;
  const int n[3] = { g.nx, g.ny, g.nz };
  const int slow = c.p.dim - 1;
  const bool has_min = c.p.bc[2 * dir] == EK_BC_COPY, has_max = c.p.bc[2 * dir + 1] == EK_BC_COPY;
  const int nb_min = c.neighbor[2 * dir], nb_max = c.neighbor[2 * dir + 1];
  if (dir == slow) {
    const long long layer = (dir == 2) ? g.sk : g.sj; // doubles per layer per variable
    const size_t cnt = (size_t)(layer * gw);
    EK_NCCL(c, api->GroupStart());
    for (int v = 0; v < g.nbvar; ++v) {
      double* base = U + g.sv * v;
      // towards the min neighbour: my layers [gw,2gw) -> its max ghosts; I receive my max ghosts
      if (has_min) EK_NCCL(c, api->Send(base + layer * gw, cnt, ncclDouble, nb_min, comm, stream));
      if (has_max) EK_NCCL(c, api->Recv(base + layer * (n[dir] + gw), cnt, ncclDouble, nb_max, comm, stream));
      // towards the max neighbour: my layers [n,n+gw) -> its min ghosts; I receive my min ghosts
      if (has_max) EK_NCCL(c, api->Send(base + layer * n[dir], cnt, ncclDouble, nb_max, comm, stream));
      if (has_min) EK_NCCL(c, api->Recv(base, cnt, ncclDouble, nb_min, comm, stream));
    }
    EK_NCCL(c, api->GroupEnd());
    return EK_OK;
  }
  // packed path (x, or y in 3-D)
  const long long ext0 = (dir == 0) ? gw : g.isize, ext1 = (dir == 1) ? gw : g.jsize, ext2 = (dir == 2) ? gw : g.ksize;
  const size_t cnt = (size_t)(ext0 * ext1 * ext2 * g.nbvar);
  for (int f = 2 * dir; f < 2 * dir + 2; ++f) {
    if (!c.halo_send[f]) {
      if (cudaMalloc(&c.halo_send[f], cnt * sizeof(double)) != cudaSuccess || cudaMalloc(&c.halo_recv[f], cnt * sizeof(double)) != cudaSuccess) {
        c.err = "cudaMalloc(halo buffers) failed";
        return EK_ERR_CUDA;
      }
    }
  }
  long long nbl = (long long)((cnt + 255) / 256);
  if (nbl > (long long)c.sm_count * 8) nbl = (long long)c.sm_count * 8;
  if (has_min) { halo_copy_kernel<true><<<(int)nbl, 256, 0, stream>>>(g, U, c.halo_send[2 * dir], dir, gw); c.launches++; }
  if (has_max) { halo_copy_kernel<true><<<(int)nbl, 256, 0, stream>>>(g, U, c.halo_send[2 * dir + 1], dir, n[dir]); c.launches++; }
  EK_NCCL(c, api->GroupStart());
  if (has_min) EK_NCCL(c, api->Send(c.halo_send[2 * dir], cnt, ncclDouble, nb_min, comm, stream));
  if (has_max) EK_NCCL(c, api->Recv(c.halo_recv[2 * dir + 1], cnt, ncclDouble, nb_max, comm, stream));
  if (has_max) EK_NCCL(c, api->Send(c.halo_send[2 * dir + 1], cnt, ncclDouble, nb_max, comm, stream));
  if (has_min) EK_NCCL(c, api->Recv(c.halo_recv[2 * dir], cnt, ncclDouble, nb_min, comm, stream));
  EK_NCCL(c, api->GroupEnd());
  if (has_min) { halo_copy_kernel<false><<<(int)nbl, 256, 0, stream>>>(g, U, c.halo_recv[2 * dir], dir, 0); c.launches++; }
  if (has_max) { halo_copy_kernel<false><<<(int)nbl, 256, 0, stream>>>(g, U, c.halo_recv[2 * dir + 1], dir, n[dir] + gw); c.launches++; }
  if (cudaGetLastError() != cudaSuccess) {
    c.err = "halo pack/unpack launch failed";
    return EK_ERR_CUDA;
  }
  return EK_OK;
}

static int allreduce(Ctx& c, void* p, int n, int op)
{
  if (!c.nccl_comm || c.nranks == 1)
    return EK_OK;
  NcclApi* api = nccl_api();
  if (!api) { c.err = "NCCL not available"; return EK_ERR_NCCL; }
  EK_NCCL(c, api->AllReduce(p, p, (size_t)n, ncclDouble, op, (ncclComm_t)c.nccl_comm, c.stream));
  return EK_OK;
}

// the bit pattern of a non-negative double orders like the double itself, so the running max can be
// reduced across ranks as a double
int comm_allreduce_invdt(Ctx& c) { return allreduce(c, &c.ts->invdt_bits, 1, ncclMax); }
int comm_allreduce_sum(Ctx& c, double* v, int n) { return allreduce(c, v, n, ncclSum); }
int comm_allreduce_max(Ctx& c, double* v, int n) { return allreduce(c, v, n, ncclMax); }

} // namespace ek

using namespace ek;

extern "C" {

int ek_comm_unique_id(void* id_out)
{
  if (!id_out) return EK_ERR_INVALID;
  NcclApi* api = nccl_api();
  if (!api) return EK_ERR_NCCL;
  static_assert(sizeof(ncclUniqueId) == EK_COMM_ID_BYTES, "id size");
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != ncclSuccess) return EK_ERR_NCCL;
  memcpy(id_out, &id, sizeof(id));
  return EK_OK;
}

int ek_comm_attach(ek_ctx* ctx, const void* id, int rank, int nranks, const int32_t neighbor_rank[6])
{
  if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) return EK_ERR_INVALID;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  NcclApi* api = nccl_api();
  if (!api) { c->err = "NCCL not loadable"; return EK_ERR_NCCL; }
  if (cudaSetDevice(c->device) != cudaSuccess) { c->err = "cudaSetDevice failed"; return EK_ERR_CUDA; }
  ek_comm_detach(ctx);
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  ncclComm_t comm = nullptr;
  EK_NCCL(*c, api->CommInitRank(&comm, nranks, uid, rank));
  c->nccl_comm = comm;
  c->rank = rank;
  c->nranks = nranks;
  for (int f = 0; f < 6; ++f) {
    c->neighbor[f] = neighbor_rank ? neighbor_rank[f] : -1;
    if (c->p.bc[f] == EK_BC_COPY && (c->neighbor[f] < 0 || c->neighbor[f] >= nranks)) {
      c->err = "EK_BC_COPY face without a valid neighbour rank";
      return EK_ERR_INVALID;
    }
  }
  c->dt_primed = false;
  c->halo_ready_for = nullptr;
  // Overlap of the z-halo exchange with the interior of the step: 3-D, slabs along z only, and a z extent that can
  // be cut into end + middle parts (hydro: z-chunks of the fused kernel; MHD: planes of the update stage).  Every rank must take the same decision (it changes the order of
  // the NCCL calls), hence the min-reduction.
  {
    const bool want = !(getenv("EK_HALO_OVERLAP") && atoi(getenv("EK_HALO_OVERLAP")) == 0);
    const bool zcopy = c->p.bc[4] == EK_BC_COPY || c->p.bc[5] == EK_BC_COPY;
    const bool xycopy = c->p.bc[0] == EK_BC_COPY || c->p.bc[1] == EK_BC_COPY || c->p.bc[2] == EK_BC_COPY || c->p.bc[3] == EK_BC_COPY;
    const bool splittable = c->p.mhd ? mhd_z_splittable(*c) : hydro_z_splittable(*c);
    double ok = (want && c->p.dim == 3 && zcopy && !xycopy && splittable) ? 1.0 : 0.0;
    if (nranks > 1) {
      if (cudaMemcpyAsync(c->scratch, &ok, sizeof(double), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { c->err = "overlap probe: H2D failed"; return EK_ERR_CUDA; }
      EK_NCCL(*c, api->AllReduce(c->scratch, c->scratch, 1, ncclDouble, ncclMin, comm, c->stream));
      if (cudaMemcpyAsync(&ok, c->scratch, sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
          cudaStreamSynchronize(c->stream) != cudaSuccess) { c->err = "overlap probe: D2H failed"; return EK_ERR_CUDA; }
    }
    c->overlap_ok = ok > 0.5;
  }
  return EK_OK;
}

int ek_comm_detach(ek_ctx* ctx)
{
  if (!ctx) return EK_ERR_INVALID;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (c->nccl_comm) {
    NcclApi* api = nccl_api();
    if (api) api->CommDestroy((ncclComm_t)c->nccl_comm);
    c->nccl_comm = nullptr;
  }
  for (int f = 0; f < 6; ++f) {
    cudaFree(c->halo_send[f]); cudaFree(c->halo_recv[f]);
    c->halo_send[f] = c->halo_recv[f] = nullptr;
  }
  c->nranks = 1;
  c->rank = 0;
  c->overlap_ok = false;
  c->halo_ready_for = nullptr;
  return EK_OK;
}

} // extern "C"
