// Tensor-core adjoint of the squared-distance cost (placeholder until the tcgen05 kernel lands:
// reports "unsupported" so that cost_abi.cu takes the CUDA-core kernels).
#include "cost.cuh"

namespace kccot {

bool tc_grad_supported(const float*, const float*, int, int, long long, const float*, const float*) { return false; }

int launch_grad_tc(const float*, const float*, const float*, int, int, int, long long, float, float*, float*, int,
                   cudaStream_t) {
  set_error("tcgen05 gradient kernel not built");
  return KCCOT_EUNSUPPORTED;
}

}  // namespace kccot
