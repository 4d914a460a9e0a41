// Internal (non-ABI) declarations shared by the translation units of libek_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "../../include/ek_b200.h"
#include "ek_common.cuh"

namespace ek {

struct NcclApi; // ek_comm.cu

struct Ctx {
  ek_params p{};
  Grid g{};
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;      // own_stream or caller-provided
  cudaStream_t comm_stream = nullptr; // halo exchange (high priority)
  cudaStream_t aux_stream = nullptr;  // middle z-chunks of a split step (runs beside the halo exchange)
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  cudaEvent_t ev_pre = nullptr, ev_edge = nullptr, ev_mid = nullptr, ev_halo = nullptr;
  // overlap of the z-halo exchange with the interior of the step (hydro 3-D, z slabs): the ghost planes of
  // level `halo_ready_for` have already been exchanged (on comm_stream, completion = ev_halo)
  bool overlap_ok = false;              // decided at ek_comm_attach, identical on every rank
  const double* halo_ready_for = nullptr;
  const double* ghosts_valid_for = nullptr; // level whose ghost cells are filled and exchanged (nothing has written it since)
  double* U[2] = { nullptr, nullptr }; // ping-pong levels
  double* snap = nullptr;              // device snapshot of a level for ek_download_state_async (allocated on first use)
  cudaStream_t io_stream = nullptr;    // its device -> host copy runs here, beside the stepping
  cudaEvent_t ev_snap = nullptr;
  bool io_pending = false;
  int cur = 0;                         // index of the current level
  double* grav = nullptr;
  double* work = nullptr;   // solver scratch (MHD: staged arrays)
  size_t work_bytes = 0;
  TimeState* ts = nullptr;  // device
  TimeState* ts_host = nullptr; // pinned
  double* scratch = nullptr; // small device scratch for diagnostics (>= 64 doubles)
  int64_t state_len = 0;
  int64_t launches = 0;
  bool dt_primed = false; // ts->invdt_bits holds max invDt of the current level
  const double* mhd_prims_of = nullptr; // level whose primitives currently sit in the MHD work area
  // level whose CFL maximum (all interior cells but the last layers) the update kernel has left in ts->invdt_bits
  const double* mhd_dt_partial_for = nullptr;
  int sm_count = 148;
  // device-resident time loop as a CUDA graph (single GPU, launch-bound grids): two consecutive steps captured once per
  // starting level, replayed by ek_run; the kernels launched by one replay are counted in graph_kernels
  cudaGraphExec_t step_graph[2] = { nullptr, nullptr };
  int64_t graph_kernels[2] = { 0, 0 };
  const void* smem_attr_for = nullptr; // hydro step kernel whose dynamic shared-memory limit has been raised on this device
  bool mhd_pet_attr = false; // dynamic shared-memory limit of mhd_pet_kernel raised on this device
  // TMA descriptors of the two time levels as 4-D tensors (x, y, z, variable) for the TMA-fed P+E+T kernel
  // (mhd_pet_tma.cuh); valid when the row stride is a multiple of 16 bytes and the driver provides the encoder
  alignas(64) CUtensorMap umap[2];
  bool umap_ok = false;
  bool mhd_pet_tma_attr = false;
  const void* mhd_ru_attr_for = nullptr; // fused R+U kernel instance whose dynamic shared-memory limit has been raised
  // communicator (z-slab / cartesian neighbours)
  void* nccl_comm = nullptr;
  int rank = 0, nranks = 1;
  int neighbor[6] = { -1, -1, -1, -1, -1, -1 };
  double* halo_send[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
  double* halo_recv[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
  // optional event pairs around the dominant kernel (ek_kernel_timing)
  bool ktiming = false;
  std::vector<cudaEvent_t> kev; // start/stop pairs recorded so far
  std::string err;
};

// record an event on the context's stream when kernel timing is enabled
inline void ktime_mark(Ctx& c)
{
  if (!c.ktiming)
    return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess)
    return;
  cudaEventRecord(e, c.stream);
  c.kev.push_back(e);
}

// hydro_launch.cu
cudaError_t launch_hydro_step(Ctx& c, const double* Uin, double* Uout, int part = 0, cudaStream_t stream = nullptr);
bool hydro_z_splittable(const Ctx& c);
cudaError_t launch_hydro_invdt(Ctx& c, const double* U);
// mhd_launch.cu
cudaError_t launch_mhd_step(Ctx& c, const double* Uin, double* Uout, int part = 0);
cudaError_t launch_mhd_invdt(Ctx& c, const double* U);
cudaError_t launch_mhd_invdt_layers(Ctx& c, const double* U);
bool mhd_dt_folded(const Ctx& c);
bool mhd_z_splittable(const Ctx& c); // the update can be cut into end planes / chunks first, the rest after
size_t mhd_work_bytes(const Grid& g);
bool mhd_make_tensor_maps(Ctx& c); // false: TMA path unavailable for this grid / driver (the plain-load kernel serves it)
// ek_comm.cu
int comm_exchange_dir(Ctx& c, double* U, int dir, cudaStream_t stream = nullptr); // halo exchange along one direction
int comm_allreduce_invdt(Ctx& c);                        // max over ranks of ts->invdt_bits
int comm_allreduce_sum(Ctx& c, double* dev_vals, int n); // diagnostics
int comm_allreduce_max(Ctx& c, double* dev_vals, int n);

} // namespace ek
