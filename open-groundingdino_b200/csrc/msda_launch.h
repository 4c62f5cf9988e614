// msda_launch.h -- host-side entry points of the individual kernel families (internal).
#pragma once

#include <cuda_runtime.h>

#include "msda_common.cuh"

namespace msda {

// Launch with programmatic dependent launch (sm_90+): the kernel may be scheduled while the previous kernel of the
// stream drains and waits at its own griddepcontrol.wait (pdl_wait(), the first statement of every kernel launched this
// way) until that kernel has completed and its writes are visible.  Stream-order semantics are unchanged; the launch
// latency and the ramp-up overlap the predecessor's tail.  Same-box A/B of two builds: cfg-3 step 1.018 -> 1.011 ms,
// cfg-2 step 6.477 -> 6.468 ms.
template <typename... KArgs, typename... Args>
inline int launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// vector kernels (fp32 / bf16 value, D in {16,32,64}, P in {4,8})
bool fwd_vec_supported(int dtype, int D, int P, int L);
int launch_forward_vec(const Problem &pr, int dtype, int grid, int min_blocks, cudaStream_t stream);
int launch_backward_vec(const Problem &pr, int dtype, int grid, int min_blocks, cudaStream_t stream);

// query-tile kernels (msda_tile.cu): encoder self-attention, D = 32, P = 4, L <= 4; `pre` as in msda_pre.cuh
bool tile_supported(int dtype, int N, int S, int M, int D, int L, int P);
int launch_forward_tile(const Problem &pr, int dtype, int sms, int pre, cudaStream_t stream);
int launch_backward_tile(const Problem &pr, int dtype, int sms, int pre, cudaStream_t stream);
int launch_backward_warp(const Problem &pr, int dtype, int sms, int pre, cudaStream_t stream);

// generic per-channel kernels (any D, P, L; fp32 / fp64 / bf16 value)
int launch_forward_generic(const Problem &pr, int dtype, cudaStream_t stream);
int launch_backward_generic(const Problem &pr, int dtype, cudaStream_t stream);

// fp32 accumulator -> bf16 grad_value
int launch_f32_to_bf16(const float *src, void *dst, long long n, int D, bool permuted, bool half, cudaStream_t stream);

// deterministic backward (two-pass)
size_t deterministic_workspace_bytes(int N, int S, int M, int D, int L, int Lq, int P, int dtype);
int launch_backward_deterministic(const Problem &pr, int dtype, void *workspace, size_t workspace_bytes,
                                  cudaStream_t stream);

int launch_sample_index(const Problem &pr, int32_t *rec, cudaStream_t stream);

}  // namespace msda
