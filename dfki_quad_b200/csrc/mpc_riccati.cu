// Kernel K2b placeholder: batched Riccati interior-point solve of the sparse OCP-QP (not yet implemented).
#include "common.cuh"
namespace b200qp {
cudaError_t launch_mpc_riccati(const MpcParams &, int, const double *, const uint8_t *, double *, double *,
                               b200qp_solver_info *, double *, double *, cudaStream_t) {
  return cudaErrorNotSupported;
}
}  // namespace b200qp
