// msda_det.cu -- deterministic backward (MSDA_BWD_DETERMINISTIC): bit-reproducible grad_value.
//
// fp32 atomic adds commute but do not associate, so the atomic mode's grad_value depends on the order
// in which the L2 atomic units see the reductions.  This mode makes the accumulation exactly
// associative instead of ordering it: contributions are converted to 64-bit fixed point and summed
// with integer atomics (order-independent by construction), in two passes over grad_output:
//   pass 1  max|grad_output| over the call  ->  a power-of-two scale 2^k such that no accumulator can
//           overflow: every contribution is w_k * attn * g with w_k <= 1, and one element of grad_value
//           receives at most Lq * L * P of them, so |sum| <= max|g| * max(1, max|attn|) * Lq * L * P < 2^(62-k).
//           (The op accepts any attention weights, not only softmax outputs, so max|attn| is measured too.)
//   pass 2  the scatter of msda_generic.cu's backward with llrint(c * 2^k) added by RED.E.ADD.64;
//           grad_sampling_loc / grad_attn_weight are written exactly as in the atomic mode (their
//           channel sums are fixed-order shuffles there too, so they were already reproducible).
//   pass 3  accumulator * 2^-k, rounded once to the value dtype.
// Resolution is max|g| * Lq*L*P * 2^-62 (~1e-13 relative to the largest gradient), far below fp32
// rounding, so results agree with the atomic mode to fp32 accuracy.  Cost: 8-byte instead of 4-byte
// accumulators (twice the reduction sectors) plus the two small extra passes.  Workspace:
// 8 * N*S*M*D + 256 bytes.  Non-finite grad_output makes the scale meaningless; the result is then NaN.
#include "msda_common.cuh"
#include "msda_launch.h"

namespace msda {

namespace {

struct DetHeader {
  unsigned max_bits;        // bit pattern of max|grad_output| (non-negative floats order like unsigned ints)
  unsigned max_attn_bits;   // bit pattern of max|attn_weight|
  unsigned pad[62];
};

template <typename VT>
__device__ __forceinline__ float to_float(VT v) { return (float)v; }
template <>
__device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_float<__half>(__half v) { return __half2float(v); }
template <typename CT, typename VT>
__device__ __forceinline__ CT widen(VT v) { return (CT)v; }
template <>
__device__ __forceinline__ float widen<float, __half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float widen<float, __nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename VT>
__global__ void __launch_bounds__(256) det_max_kernel(const VT *__restrict__ g, long long n, unsigned *slot) {
  float m = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
  {
    float a = fabsf(to_float(g[i]));
    if (!(a <= 3.4028234e38f)) a = __int_as_float(0x7f800000);   // NaN or Inf poisons the call (see header)
    m = fmaxf(m, a);
  }
  for (int d = 16; d >= 1; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0) atomicMax(slot, __float_as_uint(m));
}

// 2^k with k chosen from the header; identical in every thread of pass 2 and pass 3.
__device__ __forceinline__ double det_scale(const DetHeader *hdr, const Problem &pr) {
  // 1.000001: the maxima were rounded to fp32 (fp64 inputs may sit half an ulp above them)
  const double amax = fmax(1.0, (double)__uint_as_float(hdr->max_attn_bits));
  const double bound = (double)__uint_as_float(hdr->max_bits) * amax * (double)pr.Lq * (double)pr.L * (double)pr.P * 1.000001;
  if (!(bound > 0.0) || !isfinite(bound)) return 1.0;
  int ex;
  frexp(bound, &ex);                 // bound < 2^ex
  return ldexp(1.0, 61 - ex);        // |sum| * scale < 2^61
}

template <typename VT, typename CT>
__global__ void __launch_bounds__(128) det_scatter_kernel(const Problem pr, const DetHeader *hdr,
                                                          unsigned long long *__restrict__ acc) {
  const VT *__restrict__ value = static_cast<const VT *>(pr.value);
  const VT *__restrict__ grad_out = static_cast<const VT *>(pr.grad_out);
  const CT *__restrict__ loc = static_cast<const CT *>(pr.loc);
  const CT *__restrict__ attn = static_cast<const CT *>(pr.attn);
  CT *__restrict__ grad_loc = static_cast<CT *>(pr.grad_loc);
  CT *__restrict__ grad_attn = static_cast<CT *>(pr.grad_attn);
  const double scale = det_scale(hdr, pr);
  const int lane = threadIdx.x & 31;
  const long long items = (long long)pr.N * pr.Lq * pr.M;
  const long long MD = (long long)pr.M * pr.D;
  const int LP = pr.L * pr.P;

  for (long long item = (long long)blockIdx.x * 4 + (threadIdx.x >> 5); item < items; item += (long long)gridDim.x * 4) {
    const int m = (int)(item % pr.M);
    const int b = (int)(item / ((long long)pr.M * pr.Lq));
    for (int s = 0; s < LP; ++s) {
      const int l = s / pr.P;
      const int H = (int)pr.shapes[2 * l], W = (int)pr.shapes[2 * l + 1];
      const CT x = loc[(item * LP + s) * 2], y = loc[(item * LP + s) * 2 + 1];
      const CT a = attn[item * LP + s];
      const Cell<CT> g = locate(x, y, H, W);
      CT d1 = 0, d2 = 0, d3 = 0, d4 = 0, hh = 1, hw = 1, lh = 0, lw = 0;
      if (g.inside) {
        const int mask = corner_mask(g.h_low, g.w_low, H, W);
        lh = g.lh; lw = g.lw; hh = 1 - lh; hw = 1 - lw;
        const long long o00 = (((long long)b * pr.S + (int)pr.starts[l]) * pr.M + m) * pr.D +
                              ((long long)g.h_low * W + g.w_low) * MD;
        const long long offs[4] = {o00, o00 + MD, o00 + W * MD, o00 + W * MD + MD};
        const CT wk[4] = {hh * hw * a, hh * lw * a, lh * hw * a, lh * lw * a};
        CT dk[4] = {0, 0, 0, 0};
        for (int c = lane; c < pr.D; c += 32) {
          const CT top = widen<CT, VT>(grad_out[item * pr.D + c]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if ((mask >> k) & 1) {
              dk[k] += top * widen<CT, VT>(value[offs[k] + c]);
              const long long q = __double2ll_rn((double)(wk[k] * top) * scale);
              atomicAdd(acc + offs[k] + c, (unsigned long long)q);
            }
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          for (int dd = 16; dd >= 1; dd >>= 1) dk[k] += __shfl_xor_sync(0xffffffffu, dk[k], dd);
        d1 = dk[0]; d2 = dk[1]; d3 = dk[2]; d4 = dk[3];
      }
      if (lane == 0) {
        const bool in = g.inside;
        grad_loc[(item * LP + s) * 2] = in ? (CT)W * a * (hh * (d2 - d1) + lh * (d4 - d3)) : (CT)0;
        grad_loc[(item * LP + s) * 2 + 1] = in ? (CT)H * a * (hw * (d3 - d1) + lw * (d4 - d2)) : (CT)0;
        grad_attn[item * LP + s] = in ? hh * hw * d1 + hh * lw * d2 + lh * hw * d3 + lh * lw * d4 : (CT)0;
      }
    }
  }
}

template <typename VT>
__device__ __forceinline__ VT from_double(double v) { return (VT)v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_double<__nv_bfloat16>(double v) { return __float2bfloat16_rn((float)v); }

template <>
__device__ __forceinline__ __half from_double<__half>(double v) { return __float2half_rn((float)v); }

template <typename VT>
__global__ void __launch_bounds__(256) det_finish_kernel(const Problem pr, const DetHeader *hdr,
                                                         const long long *__restrict__ acc, VT *__restrict__ out, long long n) {
  const float mx = __uint_as_float(hdr->max_bits), ma = __uint_as_float(hdr->max_attn_bits);
  const double inv = 1.0 / det_scale(hdr, pr);
  const bool bad = !isfinite(mx) || !isfinite(ma);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = bad ? from_double<VT>(nan("")) : from_double<VT>((double)acc[i] * inv);
}

template <typename VT, typename CT>
int run_det(const Problem &pr, void *workspace, cudaStream_t stream) {
  DetHeader *hdr = static_cast<DetHeader *>(workspace);
  unsigned long long *acc = reinterpret_cast<unsigned long long *>(static_cast<char *>(workspace) + sizeof(DetHeader));
  const long long n_value = (long long)pr.N * pr.S * pr.M * pr.D;
  const long long n_go = (long long)pr.N * pr.Lq * pr.M * pr.D;
  if (cudaError_t e = cudaMemsetAsync(workspace, 0, sizeof(DetHeader) + (size_t)n_value * 8, stream)) return (int)e;
  const int cap = 148 * 16;
  auto grid_for = [&](long long n, int per) { long long b = (n + per - 1) / per; return (int)(b < cap ? b : cap); };
  det_max_kernel<VT><<<grid_for(n_go, 256), 256, 0, stream>>>(static_cast<const VT *>(pr.grad_out), n_go, &hdr->max_bits);
  const long long n_attn = (long long)pr.N * pr.Lq * pr.M * pr.L * pr.P;
  det_max_kernel<CT><<<grid_for(n_attn, 256), 256, 0, stream>>>(static_cast<const CT *>(pr.attn), n_attn, &hdr->max_attn_bits);
  const long long items = (long long)pr.N * pr.Lq * pr.M;
  const long long blocks = (items + 3) / 4;
  det_scatter_kernel<VT, CT><<<(int)(blocks < 148LL * 64 ? blocks : 148LL * 64), 128, 0, stream>>>(pr, hdr, acc);
  det_finish_kernel<VT><<<grid_for(n_value, 256), 256, 0, stream>>>(pr, hdr, reinterpret_cast<const long long *>(acc),
                                                                    static_cast<VT *>(pr.grad_value_acc), n_value);
  return (int)cudaGetLastError();
}

}  // namespace

size_t deterministic_workspace_bytes(int N, int S, int M, int D, int, int, int, int) {
  return sizeof(DetHeader) + (size_t)N * S * M * D * 8;
}

// pr.grad_value_acc holds the caller's grad_value tensor (value dtype) in this mode.
int launch_backward_deterministic(const Problem &pr, int dtype, void *workspace, size_t workspace_bytes,
                                  cudaStream_t stream) {
  if (!workspace || workspace_bytes < deterministic_workspace_bytes(pr.N, pr.S, pr.M, pr.D, 0, 0, 0, dtype))
    return MSDA_E_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(workspace) & 7u) != 0) return MSDA_E_INVALID;
  if (dtype == MSDA_F32) return run_det<float, float>(pr, workspace, stream);
  if (dtype == MSDA_BF16) return run_det<__nv_bfloat16, float>(pr, workspace, stream);
  if (dtype == MSDA_F16) return run_det<__half, float>(pr, workspace, stream);
  if (dtype == MSDA_F64) return run_det<double, double>(pr, workspace, stream);
  return MSDA_E_UNSUPPORTED;
}

}  // namespace msda
