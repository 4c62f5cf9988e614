// msda_generic.cu -- shape-agnostic kernels: any channel count, any number of levels/points, fp32,
// fp64 and bf16 values.  They cover what the vector kernels do not (the reference's odd head widths
// of OPS/test.py:85 -- D = 30, 71 -- and its fp64 path, ms_deform_attn_cuda.cu:64,134) with the same
// arithmetic.  One warp owns one (query, head) item; lanes stride over channels, so a warp reads
// whole contiguous head rows; the three channel sums of the backward are warp-shuffle reductions.
// Also: the integer-only "sample index" kernel behind msda_sample_index().
#include "msda_common.cuh"
#include "msda_launch.h"

namespace msda {

template <typename VT, typename CT>
__device__ __forceinline__ CT load_as(const VT *p, long long i) {
  return (CT)p[i];
}
template <>
__device__ __forceinline__ float load_as<__nv_bfloat16, float>(const __nv_bfloat16 *p, long long i) {
  return __bfloat162float(p[i]);
}
template <>
__device__ __forceinline__ float load_as<__half, float>(const __half *p, long long i) {
  return __half2float(p[i]);
}
template <typename VT, typename CT>
__device__ __forceinline__ void store_as(VT *p, long long i, CT v) {
  p[i] = (VT)v;
}
template <>
__device__ __forceinline__ void store_as<__nv_bfloat16, float>(__nv_bfloat16 *p, long long i, float v) {
  p[i] = __float2bfloat16_rn(v);
}

template <>
__device__ __forceinline__ void store_as<__half, float>(__half *p, long long i, float v) {
  p[i] = __float2half_rn(v);
}

constexpr int kGenericThreads = 128;

template <typename VT, typename CT>
__global__ void __launch_bounds__(kGenericThreads) fwd_generic_kernel(const Problem pr) {
  const VT *__restrict__ value = static_cast<const VT *>(pr.value);
  const CT *__restrict__ loc = static_cast<const CT *>(pr.loc);
  const CT *__restrict__ attn = static_cast<const CT *>(pr.attn);
  VT *__restrict__ out = static_cast<VT *>(pr.out);
  const int lane = threadIdx.x & 31;
  const long long items = (long long)pr.N * pr.Lq * pr.M;
  const long long MD = (long long)pr.M * pr.D;
  const int LP = pr.L * pr.P;

  for (long long item = (long long)blockIdx.x * (kGenericThreads / 32) + (threadIdx.x >> 5); item < items;
       item += (long long)gridDim.x * (kGenericThreads / 32)) {
    const int m = (int)(item % pr.M);
    const int b = (int)(item / ((long long)pr.M * pr.Lq));
    for (int c0 = 0; c0 < pr.D; c0 += 32) {
      const int c = c0 + lane;
      const bool chan = c < pr.D;
      CT acc = 0;
      for (int s = 0; s < LP; ++s) {
        const int l = s / pr.P;
        const int H = (int)pr.shapes[2 * l], W = (int)pr.shapes[2 * l + 1];
        const CT x = loc[(item * LP + s) * 2], y = loc[(item * LP + s) * 2 + 1];
        const CT a = attn[item * LP + s];
        const Cell<CT> g = locate(x, y, H, W);
        if (!g.inside || !chan) continue;
        const int mask = corner_mask(g.h_low, g.w_low, H, W);
        const CT hh = 1 - g.lh, hw = 1 - g.lw;
        const long long base = (((long long)b * pr.S + (int)pr.starts[l]) * pr.M + m) * pr.D + c;
        const long long o00 = base + ((long long)g.h_low * W + g.w_low) * MD;
        const CT v1 = (mask & 1) ? load_as<VT, CT>(value, o00) : (CT)0;
        const CT v2 = (mask & 2) ? load_as<VT, CT>(value, o00 + MD) : (CT)0;
        const CT v3 = (mask & 4) ? load_as<VT, CT>(value, o00 + W * MD) : (CT)0;
        const CT v4 = (mask & 8) ? load_as<VT, CT>(value, o00 + W * MD + MD) : (CT)0;
        acc += (hh * hw * v1 + hh * g.lw * v2 + g.lh * hw * v3 + g.lh * g.lw * v4) * a;
      }
      if (chan) store_as<VT, CT>(out, item * pr.D + c, acc);
    }
  }
}

template <typename CT>
__device__ __forceinline__ CT warp_sum(CT v) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// AT = accumulator type of grad_value (float for fp32/bf16 values, double for fp64)
template <typename VT, typename CT>
__global__ void __launch_bounds__(kGenericThreads) bwd_generic_kernel(const Problem pr) {
  const VT *__restrict__ value = static_cast<const VT *>(pr.value);
  const VT *__restrict__ grad_out = static_cast<const VT *>(pr.grad_out);
  const CT *__restrict__ loc = static_cast<const CT *>(pr.loc);
  const CT *__restrict__ attn = static_cast<const CT *>(pr.attn);
  CT *__restrict__ gacc = static_cast<CT *>(pr.grad_value_acc);
  CT *__restrict__ grad_loc = static_cast<CT *>(pr.grad_loc);
  CT *__restrict__ grad_attn = static_cast<CT *>(pr.grad_attn);
  const int lane = threadIdx.x & 31;
  const long long items = (long long)pr.N * pr.Lq * pr.M;
  const long long MD = (long long)pr.M * pr.D;
  const int LP = pr.L * pr.P;

  for (long long item = (long long)blockIdx.x * (kGenericThreads / 32) + (threadIdx.x >> 5); item < items;
       item += (long long)gridDim.x * (kGenericThreads / 32)) {
    const int m = (int)(item % pr.M);
    const int b = (int)(item / ((long long)pr.M * pr.Lq));
    for (int s = 0; s < LP; ++s) {
      const int l = s / pr.P;
      const int H = (int)pr.shapes[2 * l], W = (int)pr.shapes[2 * l + 1];
      const CT x = loc[(item * LP + s) * 2], y = loc[(item * LP + s) * 2 + 1];
      const CT a = attn[item * LP + s];
      const Cell<CT> g = locate(x, y, H, W);
      CT d1 = 0, d2 = 0, d3 = 0, d4 = 0;
      CT hh = 1, hw = 1, lh = 0, lw = 0;
      if (g.inside) {     // warp-uniform: every lane reads the same coordinates
        const int mask = corner_mask(g.h_low, g.w_low, H, W);
        lh = g.lh; lw = g.lw; hh = 1 - lh; hw = 1 - lw;
        const long long base = (((long long)b * pr.S + (int)pr.starts[l]) * pr.M + m) * pr.D;
        const long long o00 = base + ((long long)g.h_low * W + g.w_low) * MD;
        for (int c = lane; c < pr.D; c += 32) {
          const CT top = load_as<VT, CT>(grad_out, item * pr.D + c);
          const CT tgv = top * a;
          const int ca = pr.acc_permuted ? acc_pos(c, pr.D) : c;   // accumulator column (see acc_pos)
          if (mask & 1) { d1 += top * load_as<VT, CT>(value, o00 + c); atomicAdd(gacc + o00 + ca, hh * hw * tgv); }
          if (mask & 2) { d2 += top * load_as<VT, CT>(value, o00 + MD + c); atomicAdd(gacc + o00 + MD + ca, hh * lw * tgv); }
          if (mask & 4) { d3 += top * load_as<VT, CT>(value, o00 + W * MD + c); atomicAdd(gacc + o00 + W * MD + ca, lh * hw * tgv); }
          if (mask & 8) { d4 += top * load_as<VT, CT>(value, o00 + W * MD + MD + c); atomicAdd(gacc + o00 + W * MD + MD + ca, lh * lw * tgv); }
        }
        d1 = warp_sum(d1); d2 = warp_sum(d2); d3 = warp_sum(d3); d4 = warp_sum(d4);
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

static int generic_grid(const Problem &pr) {
  const long long items = (long long)pr.N * pr.Lq * pr.M;
  const long long blocks = (items + (kGenericThreads / 32) - 1) / (kGenericThreads / 32);
  const long long cap = 148LL * 64;
  return (int)(blocks < cap ? blocks : cap);
}

int launch_forward_generic(const Problem &pr, int dtype, cudaStream_t stream) {
  const int grid = generic_grid(pr);
  if (dtype == MSDA_F32) fwd_generic_kernel<float, float><<<grid, kGenericThreads, 0, stream>>>(pr);
  else if (dtype == MSDA_F64) fwd_generic_kernel<double, double><<<grid, kGenericThreads, 0, stream>>>(pr);
  else if (dtype == MSDA_BF16) fwd_generic_kernel<__nv_bfloat16, float><<<grid, kGenericThreads, 0, stream>>>(pr);
  else if (dtype == MSDA_F16) fwd_generic_kernel<__half, float><<<grid, kGenericThreads, 0, stream>>>(pr);
  else return MSDA_E_UNSUPPORTED;
  return (int)cudaGetLastError();
}

int launch_backward_generic(const Problem &pr, int dtype, cudaStream_t stream) {
  const int grid = generic_grid(pr);
  if (dtype == MSDA_F32) bwd_generic_kernel<float, float><<<grid, kGenericThreads, 0, stream>>>(pr);
  else if (dtype == MSDA_F64) bwd_generic_kernel<double, double><<<grid, kGenericThreads, 0, stream>>>(pr);
  else if (dtype == MSDA_BF16) bwd_generic_kernel<__nv_bfloat16, float><<<grid, kGenericThreads, 0, stream>>>(pr);
  else if (dtype == MSDA_F16) bwd_generic_kernel<__half, float><<<grid, kGenericThreads, 0, stream>>>(pr);
  else return MSDA_E_UNSUPPORTED;
  return (int)cudaGetLastError();
}

// ---- integer side of the sampler, one thread per sample ------------------------------------------
__global__ void __launch_bounds__(256) sample_index_kernel(const Problem pr, int32_t *__restrict__ rec) {
  const float *__restrict__ loc = static_cast<const float *>(pr.loc);
  const long long LP = (long long)pr.L * pr.P;
  const long long total = (long long)pr.N * pr.Lq * pr.M * LP;
  for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < total;
       s += (long long)gridDim.x * blockDim.x) {
    const int l = (int)((s % LP) / pr.P);
    const int b = (int)(s / (LP * pr.M * pr.Lq));
    const int H = (int)pr.shapes[2 * l], W = (int)pr.shapes[2 * l + 1];
    const Cell<float> g = locate(loc[2 * s], loc[2 * s + 1], H, W);
    int4 r;
    r.x = g.inside ? g.h_low : 0;
    r.y = g.inside ? g.w_low : 0;
    r.z = g.inside ? (1 | (corner_mask(g.h_low, g.w_low, H, W) << 1)) : 0;
    r.w = b * pr.S + (int)pr.starts[l];
    reinterpret_cast<int4 *>(rec)[s] = r;
  }
}

int launch_sample_index(const Problem &pr, int32_t *rec, cudaStream_t stream) {
  const long long total = (long long)pr.N * pr.Lq * pr.M * pr.L * pr.P;
  const long long blocks = (total + 255) / 256;
  const int grid = (int)(blocks < 148LL * 32 ? blocks : 148LL * 32);
  sample_index_kernel<<<grid, 256, 0, stream>>>(pr, rec);
  return (int)cudaGetLastError();
}

}  // namespace msda
