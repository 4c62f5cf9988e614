// msda_det.cu -- deterministic two-pass backward (placeholder until the bucketed kernels land).
#include "msda_common.cuh"
#include "msda_launch.h"

namespace msda {

size_t deterministic_workspace_bytes(int, int, int, int, int, int, int, int) { return 0; }

int launch_backward_deterministic(const Problem &, int, void *, size_t, cudaStream_t) {
  return MSDA_E_UNSUPPORTED;
}

}  // namespace msda
