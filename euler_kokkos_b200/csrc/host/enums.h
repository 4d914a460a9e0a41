// Enumerations shared with the reference's ini/solver vocabulary (src/shared/enums.h:11-243).
#pragma once

namespace euler_b200 {

using real_t = double; // EULER_KOKKOS_USE_DOUBLE (src/shared/real_type.h:26-30)

enum DimensionType { TWO_D = 2, THREE_D = 3 };

enum VarIndex { ID = 0, IP = 1, IE = 1, IU = 2, IV = 3, IW = 4, IA = 5, IB = 6, IC = 7 };

enum ComponentIndex3D { IX = 0, IY = 1, IZ = 2 };

enum FaceIdType { FACE_XMIN = 0, FACE_XMAX = 1, FACE_YMIN = 2, FACE_YMAX = 3, FACE_ZMIN = 4, FACE_ZMAX = 5 };

enum RiemannSolverType { RIEMANN_APPROX, RIEMANN_LLF, RIEMANN_HLL, RIEMANN_HLLC, RIEMANN_HLLD };

enum BoundaryConditionType { BC_UNDEFINED, BC_DIRICHLET, BC_NEUMANN, BC_PERIODIC, BC_COPY };

//! identifiers of the accumulating timers (src/shared/SolverBase.h:35-45)
enum TimerIds {
  TIMER_TOTAL = 0,
  TIMER_IO = 1,
  TIMER_DT = 2,
  TIMER_COMPUTE_PRIMITIVES = 3,
  TIMER_COMPUTE_FLUXES = 4,
  TIMER_HYDRO_UPDATE = 5,
  TIMER_BOUNDARIES = 6,
  TIMER_NUM_SCHEME = 7
};

} // namespace euler_b200
