/*
 * msda_b200.h -- C ABI of libmsda_b200.so: multi-scale deformable attention for NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary.  Each entry point replaces one function the reference exposes
 * through its pybind module `MultiScaleDeformableAttention` (paths relative to
 * /root/reference/models/GroundingDINO/ops/):
 *
 *   msda_forward            <- ms_deform_attn_forward   src/ms_deform_attn.h:20-39,  src/vision.cpp:14
 *                              (= ms_deform_attn_cuda_forward  src/cuda/ms_deform_attn_cuda.cu:20-80)
 *   msda_backward           <- ms_deform_attn_backward  src/ms_deform_attn.h:41-61,  src/vision.cpp:15
 *                              (= ms_deform_attn_cuda_backward src/cuda/ms_deform_attn_cuda.cu:83-153)
 *   msda_backward_workspace_bytes  (new: the reference allocates inside the op; here the caller owns
 *                              every byte, so the library never allocates and holds no state)
 *   msda_sample_index       (new, test hook: the integer side of the sampler,
 *                              src/cuda/ms_deform_im2col_cuda.cuh:285-288,38-41,274-278, for the
 *                              bit-exactness tests)
 *   msda_host_*             (new: the same forward+backward driven from HOST buffers, used for the
 *                              end-to-end measurement and by non-torch callers)
 *
 * Conventions
 *   - plain pointers and ints only; no torch / ATen types; device pointers unless the name says host
 *   - tensors are contiguous, row-major, exactly the reference layouts:
 *       value             (N, S, M, D)          dtype = value_dtype
 *       spatial_shapes    (L, 2) int64, (H, W)  device memory, read by the kernels (no host sync)
 *       level_start_index (L)    int64          device memory
 *       sampling_loc      (N, Lq, M, L, P, 2)   (x, y) normalised to [0,1]; fp32 (fp64 for MSDA_F64), also for
 *                                               16-bit values: a 16-bit pixel coordinate is useless
 *       attn_weight       (N, Lq, M, L, P)      fp32 (fp64 for MSDA_F64)
 *       output            (N, Lq, M*D)          dtype = value_dtype
 *       grad_output/grad_value like output/value; grad_sampling_loc/grad_attn_weight like their inputs
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it
 *   - return value: 0 on success, a negative MSDA_E_* code for a rejected call, or a positive
 *     cudaError_t if the CUDA runtime reported one.  Nothing is thrown, printed or swallowed.
 *   - im2col_step (reference .cu:50-52) does not appear: results never depended on it; the Python
 *     mirror keeps the argument and its `batch % step == 0` check.
 */
#ifndef MSDA_B200_H_
#define MSDA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSDA_ABI_VERSION 1

/* dtype of value / output / grad_output / grad_value */
enum { MSDA_F32 = 0, MSDA_BF16 = 1, MSDA_F64 = 2, MSDA_F16 = 3 };

/* backward accumulation mode for grad_value */
enum {
  MSDA_BWD_ATOMIC = 0,        /* vector red.global.add into fp32 (order not reproducible)        */
  MSDA_BWD_DETERMINISTIC = 1  /* two-pass: max|grad| -> scale, then 64-bit fixed-point integer atomics */
};

enum {
  MSDA_OK = 0,
  MSDA_E_INVALID = -1,      /* null pointer, non-positive dimension, bad enum                       */
  MSDA_E_UNSUPPORTED = -2,  /* combination this build has no kernel for                            */
  MSDA_E_WORKSPACE = -3,    /* workspace missing or smaller than msda_backward_workspace_bytes()   */
  MSDA_E_TOO_LARGE = -4     /* a per-batch slab exceeds 32-bit element indexing                    */
};

int msda_abi_version(void);
const char *msda_error_string(int code);

/* Variant selector for experiments and A/B profiling; 0 = the library's default choice.
 * bit 0..3  : kernel family (1 = generic per-channel kernels, 2 = vector kernels, linear order,
 *             3 = vector kernels, query-tile order)
 * bit 4..5  : CTAs/SM the vector kernels are register-limited for (0: default, 1: 2, 2: 3, 3: 4); non-default values need a
 *             library built with -DMSDA_BUILD_VARIANTS (MSDA_NVCC_EXTRA), otherwise MSDA_E_UNSUPPORTED
 * bit 8      : (backward, atomic mode) MSDA_VARIANT_NO_ZERO   -- the caller has already zeroed the
 *             accumulator (grad_value, or the workspace for bf16); skip the memset
 * bit 9      : (backward, bf16)         MSDA_VARIANT_NO_NARROW -- leave the fp32 accumulator in the
 *             workspace and do not write grad_value; finish with msda_backward_finish_bf16()
 *             (bits 8/9 let a profiler bracket the scatter kernel alone; same kernels, same order)
 * bit 10     : MSDA_VARIANT_OLD_TILED -- in query-tile order use the row-group vector kernels instead of the phase-
 *             structured tile kernels (A/B measurements; both are complete implementations)
 * bit 11     : MSDA_VARIANT_TILE_BWD -- (backward, query-tile order, D = 32, P = 4, L <= 4) use the tile backward (per-tile
 *             counting sort by value row + row-stationary accumulation) instead of the default warp-scoped kernel
 * bit 12     : MSDA_VARIANT_FORCE_TILE -- use the tile kernels also with 1 or 2 levels, where the dispatcher prefers the
 *             row-group kernels (tests)
 * bit 13     : MSDA_VARIANT_ROWGROUP_BWD -- (backward, query-tile order) keep the tile forward's counterpart out and use the
 *             row-group vector kernel of msda_bwd.cu (the default until round 2; A/B measurements, also on the fused entry points)
 * other bits: reserved, must be 0. */
#define MSDA_VARIANT_NO_ZERO 0x100
#define MSDA_VARIANT_NO_NARROW 0x200
#define MSDA_VARIANT_OLD_TILED 0x400
#define MSDA_VARIANT_TILE_BWD 0x800
#define MSDA_VARIANT_FORCE_TILE 0x1000
#define MSDA_VARIANT_ROWGROUP_BWD 0x2000
int msda_forward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                 const void *sampling_loc, const void *attn_weight, void *output,
                 int batch, int spatial_size, int num_heads, int channels, int num_levels,
                 int num_query, int num_point, int value_dtype, int variant, void *stream);

size_t msda_backward_workspace_bytes(int batch, int spatial_size, int num_heads, int channels,
                                     int num_levels, int num_query, int num_point, int value_dtype,
                                     int mode);

/* grad_value, grad_sampling_loc and grad_attn_weight are fully overwritten (no pre-zeroing needed). */
int msda_backward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                  const void *sampling_loc, const void *attn_weight, const void *grad_output,
                  void *grad_value, void *grad_sampling_loc, void *grad_attn_weight,
                  int batch, int spatial_size, int num_heads, int channels, int num_levels,
                  int num_query, int num_point, int value_dtype, int mode, int variant,
                  void *workspace, size_t workspace_bytes, void *stream);

/* Last stage of a bf16 backward that was launched with MSDA_VARIANT_NO_NARROW: rounds the fp32
 * accumulator held in `workspace` (library-private row layout) to bf16 grad_value (N, S, M, D). */
int msda_backward_finish_bf16(const void *workspace, void *grad_value, int batch, int spatial_size,
                              int num_heads, int channels, int num_levels, void *stream);

/* ---- fused pre-op (next row f1 of the scope table) --------------------------------------------------
 * MSDeformAttn.forward computes attention_weights = softmax(logits over L*P) and sampling_locations =
 * ref + offsets / (W,H)  (or  ref_xy + offsets / P * ref_wh * 0.5 for 4-d references) with elementwise
 * PyTorch kernels before calling the op (ops/modules/ms_deform_attn.py:98-108).  These entry points take the
 * raw linear-layer outputs instead and do that arithmetic in the kernel prologue (and its transpose in the
 * backward epilogue), so neither tensor round-trips through HBM:
 *   sampling_offsets  (N, Lq, M, L, P, 2)  q_dtype (MSDA_F32, MSDA_BF16 or MSDA_F16)
 *   attn_logits       (N, Lq, M, L*P)      q_dtype
 *   reference_points  (N, Lq, L, ref_dim)  fp32, ref_dim = 2 or 4 (no gradient is produced for them)
 * grad_offsets / grad_logits have the shapes and dtype of their inputs.  Atomic backward mode only.
 * msda_fused_supported() tells whether this build has a fused kernel for a head width / level count /
 * point count / dtype pair (D = 32, P = 4); otherwise callers use msda_forward/msda_backward. */
int msda_fused_supported(int channels, int num_levels, int num_point, int value_dtype, int q_dtype);
int msda_fused_forward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                       const void *sampling_offsets, const void *attn_logits, const float *reference_points,
                       void *output, int batch, int spatial_size, int num_heads, int channels, int num_levels,
                       int num_query, int num_point, int ref_dim, int value_dtype, int q_dtype, int variant,
                       void *stream);
int msda_fused_backward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                        const void *sampling_offsets, const void *attn_logits, const float *reference_points,
                        const void *grad_output, void *grad_value, void *grad_offsets, void *grad_logits,
                        int batch, int spatial_size, int num_heads, int channels, int num_levels, int num_query,
                        int num_point, int ref_dim, int value_dtype, int q_dtype, int variant, void *workspace,
                        size_t workspace_bytes, void *stream);

/* The same with MSDeformAttn's input_padding_mask folded in (scope row f4, ops/modules/ms_deform_attn.py:95-96):
 * padding_mask is (N, S) bytes, non-zero = padding (a torch.bool tensor as is), or NULL.  Value rows flagged as padding
 * gather as zero and their grad_value rows are zero -- exactly what `value.masked_fill(mask[..., None], 0)` and its
 * backward give -- so the caller feeds the value_proj output directly and the masked_fill pass over the whole value
 * tensor (and its backward) disappears. */
/* whether the masked entry points accept a non-NULL mask for this geometry (the model's own: 8 heads, 4 levels) */
int msda_fused_mask_supported(int num_heads, int num_levels);
int msda_fused_forward_masked(const void *value, const unsigned char *padding_mask, const int64_t *spatial_shapes,
                              const int64_t *level_start_index, const void *sampling_offsets, const void *attn_logits,
                              const float *reference_points, void *output, int batch, int spatial_size, int num_heads,
                              int channels, int num_levels, int num_query, int num_point, int ref_dim, int value_dtype,
                              int q_dtype, int variant, void *stream);
int msda_fused_backward_masked(const void *value, const unsigned char *padding_mask, const int64_t *spatial_shapes,
                               const int64_t *level_start_index, const void *sampling_offsets, const void *attn_logits,
                               const float *reference_points, const void *grad_output, void *grad_value, void *grad_offsets,
                               void *grad_logits, int batch, int spatial_size, int num_heads, int channels, int num_levels,
                               int num_query, int num_point, int ref_dim, int value_dtype, int q_dtype, int variant,
                               void *workspace, size_t workspace_bytes, void *stream);

/* rec: int32 (N, Lq, M, L, P, 4) = [h_low, w_low, flags, slab_row]; flags bit0 = sample accepted,
 * bits1..4 = which of the 4 neighbours exist; slab_row = b*S + level_start_index[l].  fp32 path. */
int msda_sample_index(const int64_t *spatial_shapes, const int64_t *level_start_index,
                      const float *sampling_loc, int batch, int spatial_size, int num_heads,
                      int num_levels, int num_query, int num_point, int32_t *rec, void *stream);

/* ---- host-buffer path ---------------------------------------------------------------------- */
typedef struct msda_host_ctx msda_host_ctx;

/* Allocates device buffers (and nothing else) for one problem size. */
int msda_host_create(msda_host_ctx **ctx, int batch, int spatial_size, int num_heads, int channels,
                     int num_levels, int num_query, int num_point, int value_dtype, int mode);
void msda_host_destroy(msda_host_ctx *ctx);
/* bytes moved per msda_host_fwd_bwd call */
size_t msda_host_h2d_bytes(const msda_host_ctx *ctx);
size_t msda_host_d2h_bytes(const msda_host_ctx *ctx);
/* All pointers are HOST memory (pinned for full speed).  Copies inputs to the device, runs forward
 * and backward, copies the four results back, and returns after the stream has drained. */
int msda_host_fwd_bwd(msda_host_ctx *ctx, const void *value, const int64_t *spatial_shapes,
                      const int64_t *level_start_index, const void *sampling_loc,
                      const void *attn_weight, const void *grad_output, void *output,
                      void *grad_value, void *grad_sampling_loc, void *grad_attn_weight);
/* The same, split: msda_host_submit() enqueues the H2D copies, both kernels and the D2H copies on the
 * context's own three streams and returns at once; msda_host_wait() blocks until the results are in the
 * host buffers.  One submission per context at a time; with several contexts in flight the copies of one
 * overlap the kernels and the opposite-direction copies of the others.  Host buffers must be left alone
 * between submit and wait. */
int msda_host_wait(msda_host_ctx *ctx);
int msda_host_submit(msda_host_ctx *ctx, const void *value, const int64_t *spatial_shapes,
                      const int64_t *level_start_index, const void *sampling_loc,
                      const void *attn_weight, const void *grad_output, void *output,
                      void *grad_value, void *grad_sampling_loc, void *grad_attn_weight);

#ifdef __cplusplus
}
#endif
#endif /* MSDA_B200_H_ */
