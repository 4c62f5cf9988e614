// msda_api.cu -- the extern "C" surface declared in include/msda_b200.h: argument checks, kernel
// family selection, accumulator handling, and the host-buffer path.  No allocation and no state
// outside msda_host_ctx; every device call is enqueued on the caller's stream.
#include <cstdlib>
#include <new>

#include "msda_common.cuh"
#include "msda_launch.h"

using namespace msda;

namespace {

struct DeviceInfo {
  int sms = 0;
};

int sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return 148;
  return sms;
}

bool is16(int dtype) { return dtype == MSDA_BF16 || dtype == MSDA_F16; }
size_t elem_size(int dtype) { return dtype == MSDA_F64 ? 8 : is16(dtype) ? 2 : 4; }
size_t coord_size(int dtype) { return dtype == MSDA_F64 ? 8 : 4; }

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int check_dims(int N, int S, int M, int D, int L, int Lq, int P, int dtype) {
  if (N <= 0 || S <= 0 || M <= 0 || D <= 0 || L <= 0 || Lq <= 0 || P <= 0) return MSDA_E_INVALID;
  if (dtype != MSDA_F32 && dtype != MSDA_BF16 && dtype != MSDA_F64 && dtype != MSDA_F16) return MSDA_E_INVALID;
  // in-slab offsets are 32-bit element indices (like the reference, cuh:47-53)
  if ((long long)S * M * D >= (1LL << 31)) return MSDA_E_TOO_LARGE;
  if ((long long)N * Lq * M >= (1LL << 31) - 64 || (long long)N * S >= (1LL << 31)) return MSDA_E_TOO_LARGE;
  return MSDA_OK;
}

// family: 1 generic, 2 vector/linear order, 3 vector/query-tile order, 0 auto
int pick_family(int variant, int dtype, int D, int P, int L, int Lq, int S, bool all_aligned) {
  const int want = variant & 15;
  const bool vec_ok = fwd_vec_supported(dtype, D, P, L) && all_aligned;
  if (want == 1 || !vec_ok) return 1;
  if (want == 2) return 2;
  if (want == 3) return 3;
  return (Lq == S) ? 3 : 2;
}

// The bf16 backward accumulates fp32 rows in the permuted channel order of acc_pos() whenever the head
// width allows it; a function of (dtype, D) only, so every kernel family and the narrowing agree.
// Exception: D = 32 keeps the natural channel order -- its kernels (pair-cooperative whole-row reductions of msda_bwd.cu,
// row-list flushes of msda_tile.cu) write whole 128-byte rows.
bool acc_is_permuted(int dtype, int D, int L) { (void)L; return is16(dtype) && D % 8 == 0 && D != 32; }

// variant bits 4..5: CTAs per SM the vector kernels are compiled for (1 -> 2, 2 -> 3, 3 -> 4); 0 picks the
// measured best: 3 for the forward (gather-latency bound, profits from 24 resident warps), 2 for the
// backward (bound by the L2 atomic units; more warps in flight only add contention) -- tools/sweep.py.
int min_blocks(int variant, int dflt) {
  const int sel = (variant >> 4) & 3;
  return sel == 0 ? dflt : sel + 1;
}

// query-tile order (family 3) runs on the tile kernels of msda_tile.cu where they exist; MSDA_VARIANT_OLD_TILED keeps the
// round-1 vector kernels in that order (A/B measurements)
bool use_tile_kernels(int variant, int family, int dtype, const Problem &pr) {
  // With one or two levels half or more of the tile kernels' (query, level) threads and lane groups idle and the row-group
  // kernels are faster (forward 0.130 vs 0.188 ms at L = 1, 0.247 vs 0.277 ms at L = 2, bs 4 fp32; profiles/r02_shape_sweep.md):
  // there the tile kernels run only when MSDA_VARIANT_FORCE_TILE asks for them.
  if (pr.L < 3 && !(variant & MSDA_VARIANT_FORCE_TILE)) return false;
  return family == 3 && !(variant & MSDA_VARIANT_OLD_TILED) && tile_supported(dtype, pr.N, pr.S, pr.M, pr.D, pr.L, pr.P);
}

int window_margin() {
  static const int m = [] {
    const char *e = getenv("MSDA_WIN_MARGIN");
    const int v = e ? atoi(e) : 5;
    return v < 0 ? 0 : (v > 64 ? 64 : v);
  }();
  return m;
}

int vec_grid(const Problem &pr, int resident_per_sm) {
  // Persistent CTAs: exactly as many as stay resident (the kernels are register-limited to
  // `resident_per_sm` CTAs of 256 threads), each walking units with a grid-stride loop, so the table
  // prologue runs once per CTA slot.  Capped by an upper bound of the unit count that needs no host
  // read of the (device-resident) shape table.
  const long long items = (long long)pr.N * pr.Lq * pr.M;
  const long long upper = (items + 7) / 8 + (long long)pr.N * pr.M * 4 * pr.L + 64;   // >= units: the smallest unit is 8 items
  const long long want = (long long)sm_count() * resident_per_sm;
  return (int)(upper < want ? upper : want);
}

}  // namespace

extern "C" {

int msda_abi_version(void) { return MSDA_ABI_VERSION; }

const char *msda_error_string(int code) {
  switch (code) {
    case MSDA_OK: return "ok";
    case MSDA_E_INVALID: return "invalid argument (null pointer, non-positive dimension or bad enum)";
    case MSDA_E_UNSUPPORTED: return "unsupported dtype/shape/mode combination";
    case MSDA_E_WORKSPACE: return "workspace missing or too small";
    case MSDA_E_TOO_LARGE: return "tensor too large for 32-bit in-slab indexing";
    default: return code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "unknown msda error";
  }
}

int msda_forward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                 const void *sampling_loc, const void *attn_weight, void *output, int N, int S, int M, int D,
                 int L, int Lq, int P, int dtype, int variant, void *stream_) {
  if (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !output)
    return MSDA_E_INVALID;
  if (int rc = check_dims(N, S, M, D, L, Lq, P, dtype)) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Problem pr{};
  pr.value = value; pr.shapes = spatial_shapes; pr.starts = level_start_index;
  pr.loc = sampling_loc; pr.attn = attn_weight; pr.out = output;
  pr.N = N; pr.S = S; pr.M = M; pr.D = D; pr.L = L; pr.Lq = Lq; pr.P = P;
  pr.tile_w_log2 = 3;
  const bool al = aligned16(value) && aligned16(sampling_loc) && aligned16(attn_weight) && aligned16(output);
  const int family = pick_family(variant, dtype, D, P, L, Lq, S, al);
  if (family == 1) return launch_forward_generic(pr, dtype, stream);
  pr.tiled = (family == 3);
  if (use_tile_kernels(variant, family, dtype, pr)) return launch_forward_tile(pr, dtype, sm_count(), 0, stream);
  return launch_forward_vec(pr, dtype, vec_grid(pr, min_blocks(variant, 3)), min_blocks(variant, 3), stream);
}

size_t msda_backward_workspace_bytes(int N, int S, int M, int D, int L, int Lq, int P, int dtype, int mode) {
  if (check_dims(N, S, M, D, L, Lq, P, dtype) != MSDA_OK) return 0;
  if (mode == MSDA_BWD_DETERMINISTIC) return deterministic_workspace_bytes(N, S, M, D, L, Lq, P, dtype);
  return is16(dtype) ? (size_t)N * S * M * D * sizeof(float) : 0;   // fp32 accumulator for bf16 / fp16
}

int msda_backward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                  const void *sampling_loc, const void *attn_weight, const void *grad_output, void *grad_value,
                  void *grad_sampling_loc, void *grad_attn_weight, int N, int S, int M, int D, int L, int Lq,
                  int P, int dtype, int mode, int variant, void *workspace, size_t workspace_bytes,
                  void *stream_) {
  if (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !grad_output ||
      !grad_value || !grad_sampling_loc || !grad_attn_weight)
    return MSDA_E_INVALID;
  if (int rc = check_dims(N, S, M, D, L, Lq, P, dtype)) return rc;
  if (mode != MSDA_BWD_ATOMIC && mode != MSDA_BWD_DETERMINISTIC) return MSDA_E_INVALID;
  const size_t need = msda_backward_workspace_bytes(N, S, M, D, L, Lq, P, dtype, mode);
  if (need > 0 && (!workspace || workspace_bytes < need)) return MSDA_E_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);

  Problem pr{};
  pr.value = value; pr.shapes = spatial_shapes; pr.starts = level_start_index;
  pr.loc = sampling_loc; pr.attn = attn_weight; pr.grad_out = grad_output;
  pr.grad_loc = grad_sampling_loc; pr.grad_attn = grad_attn_weight;
  pr.N = N; pr.S = S; pr.M = M; pr.D = D; pr.L = L; pr.Lq = Lq; pr.P = P;
  pr.tile_w_log2 = 3;

  const size_t n_value = (size_t)N * S * M * D;
  if (mode == MSDA_BWD_DETERMINISTIC) {
    pr.grad_value_acc = grad_value;
    return launch_backward_deterministic(pr, dtype, workspace, workspace_bytes, stream);
  }

  // atomic mode: accumulate into fp32 (fp64) -- grad_value itself, or the workspace for bf16
  void *acc = is16(dtype) ? workspace : grad_value;
  pr.acc_permuted = acc_is_permuted(dtype, D, L);
  const size_t acc_bytes = n_value * (dtype == MSDA_F64 ? 8 : 4);
  if (!(variant & MSDA_VARIANT_NO_ZERO))
    if (cudaError_t e = cudaMemsetAsync(acc, 0, acc_bytes, stream)) return (int)e;
  pr.grad_value_acc = acc;

  const bool al = aligned16(value) && aligned16(sampling_loc) && aligned16(attn_weight) &&
                  aligned16(grad_output) && aligned16(acc) && aligned16(grad_sampling_loc) &&
                  aligned16(grad_attn_weight);
  const int family = pick_family(variant, dtype, D, P, L, Lq, S, al);
  int rc;
  if (family == 1) {
    rc = launch_backward_generic(pr, dtype, stream);
  } else {
    pr.tiled = (family == 3);
    pr.win_margin = window_margin();
    const int abl = (variant >> 6) & 3;   // profiling ablations, see msda_bwd.cu
    // Query-tile order, D = 32, P = 4, L <= 4: the warp-scoped backward of msda_tile.cu (0.68 ms at cfg-2 against 0.73 ms
    // for the row-group kernel, which MSDA_VARIANT_ROWGROUP_BWD selects).  The tile backward with aggregation across
    // queries (counting sort + row-stationary pass) is correct and sends 3x fewer reductions to the L2, but measured slower
    // (0.93 ms, profiles/r02_tile_backward.md): opt-in through MSDA_VARIANT_TILE_BWD.
    const bool tile_ok = use_tile_kernels(variant, family, dtype, pr) && !abl;
    if (tile_ok && (variant & MSDA_VARIANT_TILE_BWD))
      rc = launch_backward_tile(pr, dtype, sm_count(), 0, stream);
    else if (tile_ok && !(variant & MSDA_VARIANT_ROWGROUP_BWD))
      rc = launch_backward_warp(pr, dtype, sm_count(), 0, stream);
    else rc = launch_backward_vec(pr, dtype, vec_grid(pr, min_blocks(variant, 2)), abl ? 100 + abl : min_blocks(variant, 2), stream);
  }
  if (rc != MSDA_OK) return rc;
  if (is16(dtype) && !(variant & MSDA_VARIANT_NO_NARROW))
    rc = launch_f32_to_bf16(static_cast<const float *>(acc), grad_value, (long long)n_value, D, pr.acc_permuted != 0,
                            dtype == MSDA_F16, stream);
  return rc;
}

// ---- fused pre-op entry points --------------------------------------------
int msda_fused_supported(int D, int L, int P, int value_dtype, int q_dtype) {
  if (D != 32 || P != 4 || L < 1) return 0;
  if (value_dtype == MSDA_F32) return (q_dtype == MSDA_F32 && L <= 4) ? 1 : 0;       // 4 level groups per warp
  if (value_dtype == MSDA_BF16) return ((q_dtype == MSDA_F32 || q_dtype == MSDA_BF16) && L <= 8) ? 1 : 0;
  if (value_dtype == MSDA_F16) return ((q_dtype == MSDA_F32 || q_dtype == MSDA_F16) && L <= 8) ? 1 : 0;
  return 0;
}

int msda_fused_mask_supported(int M, int L) { return (M == 8 && L == 4) ? 1 : 0; }

int msda_fused_forward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                       const void *sampling_offsets, const void *attn_logits, const float *reference_points,
                       void *output, int N, int S, int M, int D, int L, int Lq, int P, int ref_dim, int dtype,
                       int q_dtype, int variant, void *stream_) {
  return msda_fused_forward_masked(value, nullptr, spatial_shapes, level_start_index, sampling_offsets, attn_logits,
                                   reference_points, output, N, S, M, D, L, Lq, P, ref_dim, dtype, q_dtype, variant, stream_);
}

int msda_fused_forward_masked(const void *value, const unsigned char *padding_mask, const int64_t *spatial_shapes,
                              const int64_t *level_start_index, const void *sampling_offsets, const void *attn_logits,
                              const float *reference_points, void *output, int N, int S, int M, int D, int L, int Lq, int P,
                              int ref_dim, int dtype, int q_dtype, int variant, void *stream_) {
  if (!value || !spatial_shapes || !level_start_index || !sampling_offsets || !attn_logits || !reference_points || !output)
    return MSDA_E_INVALID;
  if (int rc = check_dims(N, S, M, D, L, Lq, P, dtype)) return rc;
  if (ref_dim != 2 && ref_dim != 4) return MSDA_E_INVALID;
  if (!msda_fused_supported(D, L, P, dtype, q_dtype)) return MSDA_E_UNSUPPORTED;
  if (padding_mask && !msda_fused_mask_supported(M, L)) return MSDA_E_UNSUPPORTED;
  if (!(aligned16(value) && aligned16(sampling_offsets) && aligned16(attn_logits) && aligned16(reference_points) &&
        aligned16(output)))
    return MSDA_E_UNSUPPORTED;
  Problem pr{};
  pr.value = value; pr.shapes = spatial_shapes; pr.starts = level_start_index;
  pr.loc = sampling_offsets; pr.attn = attn_logits; pr.ref = reference_points; pr.ref_dim = ref_dim; pr.out = output;
  pr.N = N; pr.S = S; pr.M = M; pr.D = D; pr.L = L; pr.Lq = Lq; pr.P = P;
  pr.pad_mask = padding_mask;
  pr.tile_w_log2 = 3;
  pr.tiled = ((variant & 15) == 2) ? 0 : (Lq == S);
  const int pre = q_dtype == MSDA_BF16 ? 2 : q_dtype == MSDA_F16 ? 3 : 1;
  if (use_tile_kernels(variant, pr.tiled ? 3 : 2, dtype, pr))
    return launch_forward_tile(pr, dtype, sm_count(), pre, static_cast<cudaStream_t>(stream_));
  return launch_forward_vec(pr, dtype, vec_grid(pr, 3), 200 + pre, static_cast<cudaStream_t>(stream_));
}

int msda_fused_backward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                        const void *sampling_offsets, const void *attn_logits, const float *reference_points,
                        const void *grad_output, void *grad_value, void *grad_offsets, void *grad_logits, int N, int S,
                        int M, int D, int L, int Lq, int P, int ref_dim, int dtype, int q_dtype, int variant,
                        void *workspace, size_t workspace_bytes, void *stream_) {
  return msda_fused_backward_masked(value, nullptr, spatial_shapes, level_start_index, sampling_offsets, attn_logits,
                                    reference_points, grad_output, grad_value, grad_offsets, grad_logits, N, S, M, D, L, Lq, P,
                                    ref_dim, dtype, q_dtype, variant, workspace, workspace_bytes, stream_);
}

int msda_fused_backward_masked(const void *value, const unsigned char *padding_mask, const int64_t *spatial_shapes,
                               const int64_t *level_start_index, const void *sampling_offsets, const void *attn_logits,
                               const float *reference_points, const void *grad_output, void *grad_value, void *grad_offsets,
                               void *grad_logits, int N, int S, int M, int D, int L, int Lq, int P, int ref_dim, int dtype,
                               int q_dtype, int variant, void *workspace, size_t workspace_bytes, void *stream_) {
  if (!value || !spatial_shapes || !level_start_index || !sampling_offsets || !attn_logits || !reference_points ||
      !grad_output || !grad_value || !grad_offsets || !grad_logits)
    return MSDA_E_INVALID;
  if (int rc = check_dims(N, S, M, D, L, Lq, P, dtype)) return rc;
  if (ref_dim != 2 && ref_dim != 4) return MSDA_E_INVALID;
  if (!msda_fused_supported(D, L, P, dtype, q_dtype)) return MSDA_E_UNSUPPORTED;
  if (padding_mask && !msda_fused_mask_supported(M, L)) return MSDA_E_UNSUPPORTED;
  const size_t need = msda_backward_workspace_bytes(N, S, M, D, L, Lq, P, dtype, MSDA_BWD_ATOMIC);
  if (need > 0 && (!workspace || workspace_bytes < need)) return MSDA_E_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Problem pr{};
  pr.value = value; pr.shapes = spatial_shapes; pr.starts = level_start_index;
  pr.loc = sampling_offsets; pr.attn = attn_logits; pr.ref = reference_points; pr.ref_dim = ref_dim;
  pr.grad_out = grad_output; pr.grad_loc = grad_offsets; pr.grad_attn = grad_logits;
  pr.N = N; pr.S = S; pr.M = M; pr.D = D; pr.L = L; pr.Lq = Lq; pr.P = P;
  pr.pad_mask = padding_mask;
  pr.tile_w_log2 = 3;
  pr.tiled = ((variant & 15) == 2) ? 0 : (Lq == S);
  void *acc = is16(dtype) ? workspace : grad_value;
  pr.acc_permuted = acc_is_permuted(dtype, D, L);
  pr.grad_value_acc = acc;
  if (!(aligned16(value) && aligned16(sampling_offsets) && aligned16(attn_logits) && aligned16(reference_points) &&
        aligned16(grad_output) && aligned16(acc)))
    return MSDA_E_UNSUPPORTED;
  const size_t n_value = (size_t)N * S * M * D;
  if (cudaError_t e = cudaMemsetAsync(acc, 0, n_value * 4, stream)) return (int)e;
  const int pre = q_dtype == MSDA_BF16 ? 2 : q_dtype == MSDA_F16 ? 3 : 1;
  pr.win_margin = window_margin();
  int rc;
  if (use_tile_kernels(variant, pr.tiled ? 3 : 2, dtype, pr) && !(variant & MSDA_VARIANT_ROWGROUP_BWD))
    rc = launch_backward_warp(pr, dtype, sm_count(), pre, stream);
  else rc = launch_backward_vec(pr, dtype, vec_grid(pr, 2), 200 + pre, stream);
  if (rc != MSDA_OK) return rc;
  if (is16(dtype))
    rc = launch_f32_to_bf16(static_cast<const float *>(acc), grad_value, (long long)n_value, D, pr.acc_permuted != 0,
                            dtype == MSDA_F16, stream);
  return rc;
}

int msda_backward_finish_bf16(const void *workspace, void *grad_value, int N, int S, int M, int D, int L, void *stream_) {
  if (!workspace || !grad_value || N <= 0 || S <= 0 || M <= 0 || D <= 0 || L <= 0) return MSDA_E_INVALID;
  return launch_f32_to_bf16(static_cast<const float *>(workspace), grad_value, (long long)N * S * M * D, D,
                            acc_is_permuted(MSDA_BF16, D, L), false, static_cast<cudaStream_t>(stream_));
}

int msda_sample_index(const int64_t *spatial_shapes, const int64_t *level_start_index, const float *sampling_loc,
                      int N, int S, int M, int L, int Lq, int P, int32_t *rec, void *stream_) {
  if (!spatial_shapes || !level_start_index || !sampling_loc || !rec) return MSDA_E_INVALID;
  if (int rc = check_dims(N, S, M, 1, L, Lq, P, MSDA_F32)) return rc;
  if (!aligned16(rec)) return MSDA_E_INVALID;
  Problem pr{};
  pr.shapes = spatial_shapes; pr.starts = level_start_index; pr.loc = sampling_loc;
  pr.N = N; pr.S = S; pr.M = M; pr.D = 1; pr.L = L; pr.Lq = Lq; pr.P = P;
  return launch_sample_index(pr, rec, static_cast<cudaStream_t>(stream_));
}

// ---- host-buffer path ---------------------------------------------------------------------------
// One context = device buffers for one problem size + three streams (H2D, kernels, D2H) chained by events,
// so several contexts in flight overlap their copies in both PCIe directions with each other's kernels.
struct msda_host_ctx {
  int N, S, M, D, L, Lq, P, dtype, mode;
  size_t b_value, b_loc, b_attn, b_out, b_shapes, b_starts, b_ws;
  char *d_value, *d_loc, *d_attn, *d_gout, *d_out, *d_gvalue, *d_gloc, *d_gattn, *d_ws;
  int64_t *d_shapes, *d_starts;
  cudaStream_t s_in, s_k, s_out;
  cudaEvent_t e_in, e_fwd, e_bwd, e_done, e_free;
  bool in_flight;
};

void msda_host_destroy(msda_host_ctx *c) {
  if (!c) return;
  if (c->in_flight && c->e_done) cudaEventSynchronize(c->e_done);
  void *ptrs[] = {c->d_value, c->d_loc, c->d_attn, c->d_gout, c->d_out, c->d_gvalue,
                  c->d_gloc, c->d_gattn, c->d_ws, c->d_shapes, c->d_starts};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  cudaStream_t streams[] = {c->s_in, c->s_k, c->s_out};
  for (cudaStream_t st : streams)
    if (st) cudaStreamDestroy(st);
  cudaEvent_t events[] = {c->e_in, c->e_fwd, c->e_bwd, c->e_done, c->e_free};
  for (cudaEvent_t ev : events)
    if (ev) cudaEventDestroy(ev);
  delete c;
}

int msda_host_create(msda_host_ctx **out, int N, int S, int M, int D, int L, int Lq, int P, int dtype, int mode) {
  if (!out) return MSDA_E_INVALID;
  if (int rc = check_dims(N, S, M, D, L, Lq, P, dtype)) return rc;
  if (mode != MSDA_BWD_ATOMIC && mode != MSDA_BWD_DETERMINISTIC) return MSDA_E_INVALID;
  msda_host_ctx *c = new (std::nothrow) msda_host_ctx();
  if (!c) return (int)cudaErrorMemoryAllocation;
  c->N = N; c->S = S; c->M = M; c->D = D; c->L = L; c->Lq = Lq; c->P = P; c->dtype = dtype; c->mode = mode;
  c->b_value = (size_t)N * S * M * D * elem_size(dtype);
  c->b_out = (size_t)N * Lq * M * D * elem_size(dtype);
  c->b_loc = (size_t)N * Lq * M * L * P * 2 * coord_size(dtype);
  c->b_attn = c->b_loc / 2;
  c->b_shapes = (size_t)L * 2 * sizeof(int64_t);
  c->b_starts = (size_t)L * sizeof(int64_t);
  c->b_ws = msda_backward_workspace_bytes(N, S, M, D, L, Lq, P, dtype, mode);
  cudaError_t e = cudaSuccess;
  cudaStream_t *streams[] = {&c->s_in, &c->s_k, &c->s_out};
  for (cudaStream_t *st : streams)
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(st, cudaStreamNonBlocking);
  cudaEvent_t *events[] = {&c->e_in, &c->e_fwd, &c->e_bwd, &c->e_done, &c->e_free};
  for (cudaEvent_t *ev : events)
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
  auto alloc = [&](char **p, size_t bytes) {
    if (e == cudaSuccess && bytes > 0) e = cudaMalloc(reinterpret_cast<void **>(p), bytes);
  };
  alloc(&c->d_value, c->b_value); alloc(&c->d_loc, c->b_loc); alloc(&c->d_attn, c->b_attn);
  alloc(&c->d_gout, c->b_out); alloc(&c->d_out, c->b_out); alloc(&c->d_gvalue, c->b_value);
  alloc(&c->d_gloc, c->b_loc); alloc(&c->d_gattn, c->b_attn); alloc(&c->d_ws, c->b_ws);
  alloc(reinterpret_cast<char **>(&c->d_shapes), c->b_shapes);
  alloc(reinterpret_cast<char **>(&c->d_starts), c->b_starts);
  if (e != cudaSuccess) {
    msda_host_destroy(c);
    return (int)e;
  }
  *out = c;
  return MSDA_OK;
}

size_t msda_host_h2d_bytes(const msda_host_ctx *c) {
  return c ? c->b_value + c->b_loc + c->b_attn + c->b_out + c->b_shapes + c->b_starts : 0;
}
size_t msda_host_d2h_bytes(const msda_host_ctx *c) { return c ? c->b_out + c->b_value + c->b_loc + c->b_attn : 0; }

int msda_host_submit(msda_host_ctx *c, const void *value, const int64_t *spatial_shapes,
                     const int64_t *level_start_index, const void *sampling_loc, const void *attn_weight,
                     const void *grad_output, void *output, void *grad_value, void *grad_sampling_loc,
                     void *grad_attn_weight) {
  if (!c || !value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !grad_output ||
      !output || !grad_value || !grad_sampling_loc || !grad_attn_weight)
    return MSDA_E_INVALID;
  if (c->in_flight) return MSDA_E_INVALID;   // one submission per context at a time: wait first
  cudaError_t e;
  // A failure after the first enqueue must not leave copies in flight that still reference the caller's host
  // buffers (the caller is free to release them once this returns): drain the three streams first.
  auto fail = [&](int rc) {
    cudaStreamSynchronize(c->s_in);
    cudaStreamSynchronize(c->s_k);
    cudaStreamSynchronize(c->s_out);
    return rc;
  };
#define MSDA_TRY(x) if ((e = (x)) != cudaSuccess) return fail((int)e)
  // inputs (the previous use of the device input buffers ended with the previous wait)
  MSDA_TRY(cudaMemcpyAsync(c->d_shapes, spatial_shapes, c->b_shapes, cudaMemcpyHostToDevice, c->s_in));
  MSDA_TRY(cudaMemcpyAsync(c->d_starts, level_start_index, c->b_starts, cudaMemcpyHostToDevice, c->s_in));
  MSDA_TRY(cudaMemcpyAsync(c->d_value, value, c->b_value, cudaMemcpyHostToDevice, c->s_in));
  MSDA_TRY(cudaMemcpyAsync(c->d_loc, sampling_loc, c->b_loc, cudaMemcpyHostToDevice, c->s_in));
  MSDA_TRY(cudaMemcpyAsync(c->d_attn, attn_weight, c->b_attn, cudaMemcpyHostToDevice, c->s_in));
  MSDA_TRY(cudaMemcpyAsync(c->d_gout, grad_output, c->b_out, cudaMemcpyHostToDevice, c->s_in));
  MSDA_TRY(cudaEventRecord(c->e_in, c->s_in));
  // kernels
  MSDA_TRY(cudaStreamWaitEvent(c->s_k, c->e_in, 0));
  int rc = msda_forward(c->d_value, c->d_shapes, c->d_starts, c->d_loc, c->d_attn, c->d_out, c->N, c->S, c->M,
                        c->D, c->L, c->Lq, c->P, c->dtype, 0, c->s_k);
  if (rc != MSDA_OK) return fail(rc);
  MSDA_TRY(cudaEventRecord(c->e_fwd, c->s_k));
  rc = msda_backward(c->d_value, c->d_shapes, c->d_starts, c->d_loc, c->d_attn, c->d_gout, c->d_gvalue,
                     c->d_gloc, c->d_gattn, c->N, c->S, c->M, c->D, c->L, c->Lq, c->P, c->dtype, c->mode, 0,
                     c->d_ws, c->b_ws, c->s_k);
  if (rc != MSDA_OK) return fail(rc);
  MSDA_TRY(cudaEventRecord(c->e_bwd, c->s_k));
  // results: the output leaves while the backward runs
  MSDA_TRY(cudaStreamWaitEvent(c->s_out, c->e_fwd, 0));
  MSDA_TRY(cudaMemcpyAsync(output, c->d_out, c->b_out, cudaMemcpyDeviceToHost, c->s_out));
  MSDA_TRY(cudaStreamWaitEvent(c->s_out, c->e_bwd, 0));
  MSDA_TRY(cudaMemcpyAsync(grad_value, c->d_gvalue, c->b_value, cudaMemcpyDeviceToHost, c->s_out));
  MSDA_TRY(cudaMemcpyAsync(grad_sampling_loc, c->d_gloc, c->b_loc, cudaMemcpyDeviceToHost, c->s_out));
  MSDA_TRY(cudaMemcpyAsync(grad_attn_weight, c->d_gattn, c->b_attn, cudaMemcpyDeviceToHost, c->s_out));
  MSDA_TRY(cudaEventRecord(c->e_done, c->s_out));
#undef MSDA_TRY
  c->in_flight = true;
  return MSDA_OK;
}

int msda_host_wait(msda_host_ctx *c) {
  if (!c) return MSDA_E_INVALID;
  if (!c->in_flight) return MSDA_OK;
  c->in_flight = false;
  return (int)cudaEventSynchronize(c->e_done);
}

int msda_host_fwd_bwd(msda_host_ctx *c, const void *value, const int64_t *spatial_shapes,
                      const int64_t *level_start_index, const void *sampling_loc, const void *attn_weight,
                      const void *grad_output, void *output, void *grad_value, void *grad_sampling_loc,
                      void *grad_attn_weight) {
  const int rc = msda_host_submit(c, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                                  output, grad_value, grad_sampling_loc, grad_attn_weight);
  return rc != MSDA_OK ? rc : msda_host_wait(c);
}

}  // extern "C"
