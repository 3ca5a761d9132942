// ukf_batch.cu -- batched UKF-SLAM step (placeholder until the kernel lands; fails loudly).
#include "common.cuh"
namespace slam {
size_t ukf_step_smem_bytes(const BatchState&) { return 0; }
cudaError_t ukf_step_configure(const BatchState&) { return cudaSuccess; }
cudaError_t launch_ukf_step(const BatchState&, const FilterConst&, const StepInputs&, cudaStream_t) { return cudaErrorNotSupported; }
}
