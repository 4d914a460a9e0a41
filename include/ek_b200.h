/*
 * ek_b200.h -- C ABI of the B200-native MUSCL-Hancock time step (Euler + CT-MHD).
 *
 * This is the drop-in boundary for the hot path of pkestene/euler_kokkos.  The reference has no
 * FFI: its boundary is the C++ virtual interface SolverBase + the string registry SolverFactory
 * (src/shared/SolverBase.h:53-259, src/shared/SolverFactory.h:24-118).  Our C++ host classes of the
 * same names (euler_kokkos_b200/csrc/host) are what is registered under the four solver names; they
 * call ONLY the entry points below.  Each entry point cites the reference code it replaces.
 *
 * Contract
 *  - opaque handle, one per GPU / sub-domain; the library owns all device memory;
 *  - every call returns an int status (EK_OK == 0); nothing aborts or throws across the ABI;
 *    ek_last_error() returns a description of the last failure on that handle;
 *  - one host thread per handle; all work is enqueued on the handle's stream; the only host
 *    synchronisations are inside calls that return a host scalar or host array;
 *  - data layout: SoA  U[var][k][j][i], ghosts included, i fastest  (== the reference's
 *    Kokkos::LayoutLeft device layout of DataArray3d(i,j,k,ivar), src/shared/kokkos_shared.h:24-27);
 *    2-D arrays are U[var][j][i] (ksize == 1);
 *  - all arithmetic is IEEE FP64 (reference: real_t = double, src/shared/real_type.h:26-30);
 *  - there is NO CPU fallback: without a CUDA device every compute call fails with EK_ERR_CUDA.
 */
#ifndef EK_B200_H_
#define EK_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EK_ABI_VERSION 1

/* status codes */
enum {
  EK_OK = 0,
  EK_ERR_INVALID = 1,     /* bad argument / unsupported parameter combination */
  EK_ERR_CUDA = 2,        /* CUDA runtime error (no device, OOM, launch failure) */
  EK_ERR_NCCL = 3,        /* NCCL error or NCCL not loadable */
  EK_ERR_STATE = 4        /* call not legal in the current state */
};

/* variable indices, src/shared/enums.h:17-40 */
enum { EK_ID = 0, EK_IP = 1, EK_IU = 2, EK_IV = 3, EK_IW = 4, EK_IA = 5, EK_IB = 6, EK_IC = 7 };

/* Riemann solvers, src/shared/enums.h:142-149 (same numeric values) */
enum { EK_RIEMANN_APPROX = 0, EK_RIEMANN_LLF = 1, EK_RIEMANN_HLL = 2, EK_RIEMANN_HLLC = 3,
       EK_RIEMANN_HLLD = 4 };

/* boundary conditions, src/shared/enums.h:153-160 (same numeric values) */
enum { EK_BC_UNDEFINED = 0, EK_BC_DIRICHLET = 1, EK_BC_NEUMANN = 2, EK_BC_PERIODIC = 3,
       EK_BC_COPY = 4 /* face is an internal sub-domain interface: filled by the halo exchange */ };

/* faces, src/shared/enums.h:120-130 */
enum { EK_XMIN = 0, EK_XMAX = 1, EK_YMIN = 2, EK_YMAX = 3, EK_ZMIN = 4, EK_ZMAX = 5 };

/*
 * Everything the step needs to know; mirrors HydroParams + HydroSettings + GravityParams
 * (src/shared/HydroParams.h:32-197, src/shared/GravityParams.h:21-39).  The caller is responsible
 * for the reference's float-parsing semantics (ConfigMap::getFloat, src/utils/config/ConfigMap.cpp:
 * 38-46) and for the derived quantities of HydroParams::init (src/shared/HydroParams.cpp:327-359).
 */
typedef struct ek_params {
  int32_t dim;              /* 2 or 3 */
  int32_t mhd;              /* 0: Hydro_Muscl_*, 1: MHD_Muscl_* */
  int32_t nbvar;            /* 4 (hydro 2D), 5 (hydro 3D), 8 (MHD) */
  int32_t ghost;            /* 2 (hydro), 3 (MHD) */
  int32_t nx, ny, nz;       /* LOCAL interior sizes (nz = 1 in 2-D) */
  int32_t riemann;          /* EK_RIEMANN_* */
  int32_t impl_version;     /* [other] implementationVersion: 0,1,2 */
  int32_t bc[6];            /* EK_BC_* per face; EK_BC_COPY for internal faces */
  int32_t gravity_enabled;
  int32_t gravity_hancock;  /* [gravity] hancock_predictor_enabled */
  int32_t reserved0;
  double dx, dy, dz;
  double gamma0, gamma6, smallr, smallc, smallp, smallpp, cfl, slope_type;
  double reserved1[4];
} ek_params;

typedef struct ek_ctx ek_ctx;

/* ---- life cycle ------------------------------------------------------------------------------ */
/* Allocates U, U2 (+ solver work arrays) on `device`.  Replaces the allocations of
 * SolverHydroMuscl ctor (src/muscl/SolverHydroMuscl.h:164-274) / SolverMHDMuscl ctor
 * (src/muscl/SolverMHDMuscl.h:265-382). */
int ek_create(const ek_params* params, int device, ek_ctx** out);
int ek_destroy(ek_ctx* ctx);
const char* ek_last_error(const ek_ctx* ctx);   /* ctx may be NULL: last create() error */
int ek_abi_version(void);

/* Use an externally owned cudaStream_t for all work of this handle (NULL: back to the own one). */
int ek_set_stream(ek_ctx* ctx, void* cuda_stream);
/* cudaStreamSynchronize on the handle's stream. */
int ek_sync(ek_ctx* ctx);

/* ---- state ----------------------------------------------------------------------------------- */
/* number of doubles of one ghosted state: nbvar*ksize*jsize*isize */
int64_t ek_state_len(const ek_ctx* ctx);
/* Host -> current time level (ghosts included).  Replaces Init*Functor + deep_copy(U2,U)
 * (src/muscl/SolverHydroMuscl.h:277-283): both levels are set. Resets iteration parity to 0. */
int ek_upload_state(ek_ctx* ctx, const double* U_host);
/* Current time level -> host (ghosted SoA).  Replaces Kokkos::deep_copy(Uhost,Udata) in
 * save_solution_impl (src/muscl/SolverHydroMuscl.h:751-762). Synchronises. */
int ek_download_state(ek_ctx* ctx, double* U_host);
/* Asynchronous variant for output that overlaps the stepping: the current level is snapshot on the
 * device (D2D, ordered on the handle's stream, a few ms), then copied to U_host -- pinned memory for a real overlap --
 * on a separate stream while later steps run.  ek_download_wait blocks until that copy has landed (and is implied by the
 * next ek_download_state_async and by ek_destroy).  One download in flight per handle. */
int ek_download_state_async(ek_ctx* ctx, double* U_host);
/* page-locked host memory for it (cudaMallocHost / cudaFreeHost: the host layer does not link the CUDA runtime itself) */
int ek_alloc_host(size_t bytes, void** out);
int ek_free_host(void* p);
int ek_download_wait(ek_ctx* ctx);
/* Device pointer to the current time level (for zero-copy callers holding device memory). */
int ek_device_state(ek_ctx* ctx, double** dptr);
/* Static gravity field g[3 or 2][k][j][i] (src/muscl/SolverHydroMuscl.h:268-272). */
int ek_set_gravity(ek_ctx* ctx, const double* g_host);

/* ---- the hot path ---------------------------------------------------------------------------- */
/* Ghost fill of the current level: physical BCs in the reference's order XMIN,XMAX,YMIN,YMAX,
 * ZMIN,ZMAX with full transverse extent, internal faces (EK_BC_COPY) by halo exchange.
 * Replaces SolverBase::make_boundaries_serial / _mpi (src/shared/SolverBase.cpp:502-512,583-664)
 * and MakeBoundariesFunctor{2D,3D}{,_MHD} (src/shared/BoundariesFunctors.h:17-1050). */
int ek_fill_boundaries(ek_ctx* ctx);

/* CFL time step of the current level: cfl / max_cells(invDt), min over ranks when a communicator
 * is attached; NOT clamped to tEnd (the caller does, SolverBase.cpp:149-152).  Synchronises.
 * Hydro: ComputeDtFunctor{2D,3D} / ComputeDtGravityFunctor (src/muscl/HydroRunFunctors3D.h:16-195)
 * from U.  MHD: ConvertToPrimitives + ComputeDtFunctor3D_MHD (src/muscl/MHDRunFunctors3D.h:16-163)
 * from the primitive variables; ghosts must be filled (ek_fill_boundaries) before. */
int ek_compute_dt(ek_ctx* ctx, double* dt_out);

/* One godunov_unsplit with the given dt: ghost fill of the current level, fused
 * primitives/slopes/trace/Riemann/update into the other level, flip.
 * Replaces SolverHydroMuscl<dim>::godunov_unsplit_impl (src/muscl/SolverHydroMuscl.cpp:216-390)
 * and SolverMHDMuscl<dim>::godunov_unsplit_impl minus its compute_dt
 * (src/muscl/SolverMHDMuscl.cpp:371-568).  Asynchronous. */
int ek_step(ek_ctx* ctx, double dt);

/* Sync-free time loop: up to `nsteps` iterations of {dt = cfl/invDt (device resident, min over
 * ranks), clamp to tEnd - t, step, t += dt}; stops early once t >= tEnd - 1e-14
 * (SolverBase::finished, src/shared/SolverBase.cpp:171-176).  Replaces the while loop of
 * src/main.cpp:126-131 with nOutput == 0.  One host synchronisation at the end. */
int ek_run(ek_ctx* ctx, int nsteps, double t, double t_end,
           double* t_out, double* dt_last_out, int* steps_done_out);

/* ---- diagnostics (off the hot path) ------------------------------------------------------------ */
/* Sum over interior cells of every conserved variable (nbvar doubles). Synchronises. */
int ek_conserved_sums(ek_ctx* ctx, double* sums_out);
/* MHD only: max over interior cells of |div B| (face-centred differences). Synchronises. */
int ek_max_divb(ek_ctx* ctx, double* out);
/* number of kernels launched by this handle since creation */
int64_t ek_launch_count(const ek_ctx* ctx);
/* Device timing of the dominant (fused step) kernel: when enabled, a CUDA event pair is recorded on
 * the handle's stream around every launch of that kernel (no host synchronisation is added);
 * ek_kernel_time synchronises, returns the summed duration and launch count, and resets both. */
int ek_kernel_timing(ek_ctx* ctx, int enable);
int ek_kernel_time(ek_ctx* ctx, double* total_ms_out, int* launches_out);

/* Arithmetic self-test of the device primitives the kernels are built from (csrc/ek_common.cuh): for each of the n
 * operands x[i] writes max(x,0), min(x,0) (the setp/selp forms), 1/x, sqrt(x), 1/sqrt(x) and (1+x)/x (MUFU seed +
 * Newton steps) to out[0..6n).  The tests compare them with IEEE arithmetic on +-0, denormals, +-inf, NaN and random
 * operands; nothing in the reference corresponds to it. */
int ek_selftest_math(int device, const double* x_host, int n, double* out_host);

/* ---- multi-GPU: z-slab (or mx*my*mz) decomposition, halos over NCCL/NVLink -------------------- */
/* Replaces src/shared/mpiBorderUtils.h + SolverBase::transfer_boundaries_* (MPI_Sendrecv,
 * src/shared/SolverBase.cpp:756-845) and MPI_Allreduce<MIN> (SolverBase.cpp:134-138). */
#define EK_COMM_ID_BYTES 128
int ek_comm_unique_id(void* id_out /* EK_COMM_ID_BYTES */);
/* neighbor_rank[f]: rank owning the sub-domain across face f (used only where bc[f]==EK_BC_COPY) */
int ek_comm_attach(ek_ctx* ctx, const void* id, int rank, int nranks, const int32_t neighbor_rank[6]);
int ek_comm_detach(ek_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* EK_B200_H_ */
