// msda_bwd.cu -- backward kernels, atomic mode (replaces the six
// ms_deformable_col2im_gpu_kernel_* variants, reference ms_deform_im2col_cuda.cuh:301-920, the
// per-sample routine :87-159 and the launcher :956-1327).
//
// Same item/level/lane mapping as the forward (msda_fwd.cu).  Per sample a lane holds 16 bytes of
// each of the 4 neighbour rows and of grad_output, and produces
//   * the grad_value contribution  w_k * attn * grad_out  as ONE 16-byte vector reduction per
//     neighbour (REDG.E.ADD.F32x4, executed by the L2 atomic units) instead of the reference's
//     scalar atomicAdd per channel -- 4x fewer atomic requests, each a full 128-byte row per 8 lanes;
//   * four partial dot products d_k = <grad_out, v_k> over its channels.  Everything the two small
//     gradients need is linear in the d_k:
//         grad_attn  = w1 d1 + w2 d2 + w3 d3 + w4 d4
//         grad_loc.x = W * attn * ( hh (d2 - d1) + lh (d4 - d3) )
//         grad_loc.y = H * attn * ( hw (d3 - d1) + lw (d4 - d2) )
//     so the channel reduction is one reduce-scatter of the 4*P dot products over the LPG lanes of
//     the group (xor-shuffles; 14 + 2 shuffles for fp32 D=32 P=4) replacing the reference's
//     shared-memory staging, 2 barriers and lane-0 serial sums per sample.  The lanes that end up
//     owning a point write its three numbers with plain coalesced stores: every element of
//     grad_sampling_loc / grad_attn_weight is written exactly once, so neither needs a memset
//     (the reference zero-fills both, ms_deform_attn_cuda.cu:122-123).
// grad_value is accumulated in fp32 whatever the value dtype (bf16 atomics would round at every
// add); for bf16 the accumulator is a workspace that msda_api.cu narrows once at the end.
#include "msda_common.cuh"
#include "msda_launch.h"

namespace msda {

// After the call lane j (0 <= j < LPG) holds the group sums of d[j*c .. j*c+c) in d[0..c), c = NV/LPG.
// When NV < LPG the tail rounds degrade to plain all-reduces.
template <int N, int DIST>
__device__ __forceinline__ void reduce_scatter(float *d, int j) {
  if constexpr (DIST >= 1) {
    if constexpr (N >= 2) {
      constexpr int HALF = N / 2;
      const bool upper = (j & DIST) != 0;
#pragma unroll
      for (int i = 0; i < HALF; ++i) {
        const float send = upper ? d[i] : d[i + HALF];
        const float keep = upper ? d[i + HALF] : d[i];
        d[i] = keep + __shfl_xor_sync(0xffffffffu, send, DIST);
      }
      reduce_scatter<HALF, DIST / 2>(d, j);
    } else {
      d[0] += __shfl_xor_sync(0xffffffffu, d[0], DIST);
      reduce_scatter<1, DIST / 2>(d, j);
    }
  }
}

template <typename VT, int D, int P, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) bwd_vec_kernel(const Problem pr) {
  constexpr int CH = Chan<VT>::kPerLane;
  constexpr int LPG = D / CH;
  constexpr int G = 32 / LPG;
  constexpr int NV = 4 * P;                       // dot products per (item, level)
  constexpr int C = NV >= LPG ? NV / LPG : 1;     // of which a lane owns C after the reduce-scatter

  __shared__ Tables tb;
  load_tables(pr, tb);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / LPG, j = lane % LPG;
  const int GI = min(G, pow2_ceil(pr.L));
  const int IPW = G / GI;
  const int rounds = (pr.L + GI - 1) / GI;
  const int sub = g / GI, gi = g % GI;
  const int MD = pr.M * D;

  const VT *__restrict__ value = static_cast<const VT *>(pr.value);
  const VT *__restrict__ grad_out = static_cast<const VT *>(pr.grad_out);
  const float *__restrict__ loc = static_cast<const float *>(pr.loc);
  const float *__restrict__ attn = static_cast<const float *>(pr.attn);
  float *__restrict__ gacc = static_cast<float *>(pr.grad_value_acc);
  float *__restrict__ grad_loc = static_cast<float *>(pr.grad_loc);
  float *__restrict__ grad_attn = static_cast<float *>(pr.grad_attn);

  for (long long unit = blockIdx.x; unit < tb.n_units; unit += gridDim.x) {
    for (int k0 = warp * IPW; k0 < kUnitItems; k0 += (kThreads / 32) * IPW) {
      const Item it = decode_item(pr, tb, unit, k0 + sub);
      const long long item = ((long long)it.b * pr.Lq + it.q) * pr.M + it.m;

      float go[CH];
      Chan<VT>::unpack(gather16(grad_out + item * D + j * CH, it.valid), go);

      for (int r = 0; r < rounds; ++r) {
        const int l = r * GI + gi;
        const bool act = it.valid && l < pr.L;
        const int lc = act ? l : 0;
        const int H = tb.H[lc], W = tb.W[lc];

        float xy[2 * P], a[P];
        {
          const uint4 *lp = reinterpret_cast<const uint4 *>(loc + (item * pr.L + lc) * (2 * P));
          const uint4 *ap = reinterpret_cast<const uint4 *>(attn + (item * pr.L + lc) * P);
#pragma unroll
          for (int v = 0; v < P / 2; ++v) {
            const uint4 u = gather16(lp + v, act);
            xy[4 * v + 0] = __uint_as_float(u.x); xy[4 * v + 1] = __uint_as_float(u.y);
            xy[4 * v + 2] = __uint_as_float(u.z); xy[4 * v + 3] = __uint_as_float(u.w);
          }
#pragma unroll
          for (int v = 0; v < P / 4; ++v) {
            const uint4 u = gather16(ap + v, act);
            a[4 * v + 0] = __uint_as_float(u.x); a[4 * v + 1] = __uint_as_float(u.y);
            a[4 * v + 2] = __uint_as_float(u.z); a[4 * v + 3] = __uint_as_float(u.w);
          }
        }
        const long long slab_off = (((long long)it.b * pr.S + tb.start[lc]) * pr.M + it.m) * D + j * CH;
        const VT *slab = value + slab_off;
        float *gslab = gacc + slab_off;

        float d[NV];
        float g_lh[P], g_lw[P], g_aw[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
          const Cell<float> c = locate(xy[2 * p], xy[2 * p + 1], H, W);
          const bool in = act && c.inside;
          const int mask = in ? corner_mask(c.h_low, c.w_low, H, W) : 0;
          g_lh[p] = in ? c.lh : 0.f;
          g_lw[p] = in ? c.lw : 0.f;
          g_aw[p] = in ? a[p] : 0.f;
          const float hh = 1.f - g_lh[p], hw = 1.f - g_lw[p];
          const float cw[4] = {hh * hw * g_aw[p], hh * g_lw[p] * g_aw[p], g_lh[p] * hw * g_aw[p],
                               g_lh[p] * g_lw[p] * g_aw[p]};
          const int o00 = (c.h_low * W + c.w_low) * MD;
          const int off[4] = {o00, o00 + MD, o00 + W * MD, o00 + W * MD + MD};
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const bool on = (mask >> kk) & 1;
            float v[CH];
            Chan<VT>::unpack(gather16(slab + off[kk], on), v);
            float dot = 0.f;
#pragma unroll
            for (int cidx = 0; cidx < CH; ++cidx) dot = fmaf(go[cidx], v[cidx], dot);
            d[4 * p + kk] = dot;
            if (on) {
#pragma unroll
              for (int h4 = 0; h4 < CH / 4; ++h4)
                red_add_f32x4(gslab + off[kk] + 4 * h4, cw[kk] * go[4 * h4], cw[kk] * go[4 * h4 + 1],
                              cw[kk] * go[4 * h4 + 2], cw[kk] * go[4 * h4 + 3]);
            }
          }
        }

        reduce_scatter<NV, LPG / 2>(d, j);

        // Which point(s) this lane now owns, and with which of the four dot products.
        float *gl = grad_loc + (item * pr.L + lc) * (2 * P);
        float *ga = grad_attn + (item * pr.L + lc) * P;
        if constexpr (C >= 4) {
#pragma unroll
          for (int t = 0; t < C / 4; ++t) {
            const int p = j * (C / 4) + t;
            float lh = 0.f, lw = 0.f, aw = 0.f;
#pragma unroll
            for (int pp = 0; pp < P; ++pp)
              if (pp == p) { lh = g_lh[pp]; lw = g_lw[pp]; aw = g_aw[pp]; }
            const float hh = 1.f - lh, hw = 1.f - lw;
            const float d1 = d[4 * t], d2 = d[4 * t + 1], d3 = d[4 * t + 2], d4 = d[4 * t + 3];
            const float gx = (float)W * aw * (hh * (d2 - d1) + lh * (d4 - d3));
            const float gy = (float)H * aw * (hw * (d3 - d1) + lw * (d4 - d2));
            const float gw = hh * hw * d1 + hh * lw * d2 + lh * hw * d3 + lh * lw * d4;
            if (act) {
              *reinterpret_cast<float2 *>(gl + 2 * p) = make_float2(gx, gy);
              ga[p] = gw;
            }
          }
        } else {
          // C == 2: lanes (2p, 2p+1) hold (d1,d2) / (d3,d4);  C == 1: lanes 4p..4p+3 hold d1..d4.
          constexpr int LPP = 4 / C;               // lanes per point
          const int p = j / LPP, part = j % LPP;
          float d1, d2, d3, d4;
          if constexpr (C == 2) {
            const float o0 = __shfl_xor_sync(0xffffffffu, d[0], 1);
            const float o1 = __shfl_xor_sync(0xffffffffu, d[1], 1);
            d1 = part ? o0 : d[0]; d2 = part ? o1 : d[1];
            d3 = part ? d[0] : o0; d4 = part ? d[1] : o1;
          } else {
            const float o1 = __shfl_xor_sync(0xffffffffu, d[0], 1);
            const float lo = (part & 1) ? o1 : d[0], hi = (part & 1) ? d[0] : o1;   // (even, odd) of my pair
            const float plo = __shfl_xor_sync(0xffffffffu, lo, 2);
            const float phi = __shfl_xor_sync(0xffffffffu, hi, 2);
            d1 = (part & 2) ? plo : lo; d2 = (part & 2) ? phi : hi;
            d3 = (part & 2) ? lo : plo; d4 = (part & 2) ? hi : phi;
          }
          float lh = 0.f, lw = 0.f, aw = 0.f;
#pragma unroll
          for (int pp = 0; pp < P; ++pp)
            if (pp == p) { lh = g_lh[pp]; lw = g_lw[pp]; aw = g_aw[pp]; }
          const float hh = 1.f - lh, hw = 1.f - lw;
          const float gx = (float)W * aw * (hh * (d2 - d1) + lh * (d4 - d3));
          const float gy = (float)H * aw * (hw * (d3 - d1) + lw * (d4 - d2));
          const float gw = hh * hw * d1 + hh * lw * d2 + lh * hw * d3 + lh * lw * d4;
          if (act && p < P) {
            if (part == 0) gl[2 * p] = gx;
            if (part == 1) gl[2 * p + 1] = gy;
            if (part == LPP - 1) ga[p] = gw;
          }
        }
      }
    }
  }
}

template <typename VT, int D, int P>
static cudaError_t launch_bwd_vec(const Problem &pr, int grid, cudaStream_t stream) {
  bwd_vec_kernel<VT, D, P, 2><<<grid, kThreads, 0, stream>>>(pr);
  return cudaGetLastError();
}

template <typename VT>
static int dispatch_bwd(const Problem &pr, int grid, cudaStream_t stream) {
#define MSDA_CASE(DD, PP) \
  if (pr.D == DD && pr.P == PP) return (int)launch_bwd_vec<VT, DD, PP>(pr, grid, stream);
  MSDA_CASE(32, 4) MSDA_CASE(32, 8) MSDA_CASE(16, 4) MSDA_CASE(16, 8) MSDA_CASE(64, 4) MSDA_CASE(64, 8)
#undef MSDA_CASE
  return MSDA_E_UNSUPPORTED;
}

int launch_backward_vec(const Problem &pr, int dtype, int grid, cudaStream_t stream) {
  if (dtype == MSDA_F32) return dispatch_bwd<float>(pr, grid, stream);
  if (dtype == MSDA_BF16) return dispatch_bwd<__nv_bfloat16>(pr, grid, stream);
  return MSDA_E_UNSUPPORTED;
}

// ---- fp32 accumulator -> bf16 ------------------------------------------------------------------
__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float4 *__restrict__ src, uint2 *__restrict__ dst,
                                                         long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 v = src[i];
    dst[i] = make_uint2(Chan<__nv_bfloat16>::pack2(v.x, v.y), Chan<__nv_bfloat16>::pack2(v.z, v.w));
  }
}

__global__ void f32_to_bf16_tail_kernel(const float *__restrict__ src, __nv_bfloat16 *__restrict__ dst,
                                        long long first, long long n) {
  const long long i = first + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}

int launch_f32_to_bf16(const float *src, void *dst, long long n, cudaStream_t stream) {
  const long long n4 = n / 4;
  if (n4 > 0) {
    const long long blocks = (n4 + 255) / 256;
    const int grid = (int)(blocks < 148LL * 16 ? blocks : 148LL * 16);
    f32_to_bf16_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4 *>(src),
                                                 reinterpret_cast<uint2 *>(dst), n4);
  }
  if (n4 * 4 < n)
    f32_to_bf16_tail_kernel<<<1, 32, 0, stream>>>(src, static_cast<__nv_bfloat16 *>(dst), n4 * 4, n);
  return (int)cudaGetLastError();
}

}  // namespace msda
