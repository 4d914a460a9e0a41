// Host-side dispatch of the CT-MHD step (placeholder until the MHD kernels land).
#include "ek_internal.h"

namespace ek {
size_t mhd_work_bytes(const Grid&) { return 0; }
cudaError_t launch_mhd_step(Ctx&, const double*, double*) { return cudaErrorNotSupported; }
cudaError_t launch_mhd_invdt(Ctx&, const double*) { return cudaErrorNotSupported; }
} // namespace ek
