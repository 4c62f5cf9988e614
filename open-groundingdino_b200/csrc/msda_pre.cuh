// msda_pre.cuh -- per-(item, level) inputs of the vector kernels.
//
// PRE = 0: the reference operator boundary -- sampling locations and attention weights are given
//          (MSDeformAttnFunction.apply, ops/functions/ms_deform_attn_func.py:23).
// PRE = 1/2: the "next" row f1 of SURVEY.md section 8 -- the softmax over L*P and the location arithmetic of
//          MSDeformAttn.forward (ops/modules/ms_deform_attn.py:98-108) are folded into the kernel prologue, so
//          the kernel reads the raw sampling_offsets / attention logits (fp32: PRE 1, bf16: PRE 2, fp16: PRE 3) and the
//          reference points, and neither `sampling_locations` nor `attention_weights` ever exists in HBM.
//          Arithmetic mirrors the PyTorch expressions operation by operation so the pixel coordinates stay
//          bit-identical to the unfused module:
//            2-d refs:  loc = ref + off / (W_l, H_l)
//            4-d refs:  loc = ref_xy + off / P * ref_wh * 0.5        (evaluated left to right)
//          and under autocast (bf16 offsets) the first quotient is rounded to bf16, because that is the dtype
//          PyTorch gives `bf16_tensor / int`.  The softmax is exp(x - max) / sum in fp32.
#pragma once

#include "msda_common.cuh"

namespace msda {

// what the backward needs to turn d(loss)/d(loc) into d(loss)/d(offset): g * mul / div
struct OffsetScale {
  float mulx, muly, divx, divy;
};

template <typename QT>
__device__ __forceinline__ float round_like(float v);
template <>
__device__ __forceinline__ float round_like<float>(float v) { return v; }
template <>
__device__ __forceinline__ float round_like<__nv_bfloat16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

template <>
__device__ __forceinline__ float round_like<__half>(float v) { return __half2float(__float2half_rn(v)); }

// P floats
template <int P>
__device__ __forceinline__ void load_q_f32(const float *p, bool act, float *out) {
#pragma unroll
  for (int v = 0; v < P / 4; ++v) {
    const uint4 u = gather16(reinterpret_cast<const uint4 *>(p) + v, act);
    out[4 * v + 0] = __uint_as_float(u.x); out[4 * v + 1] = __uint_as_float(u.y);
    out[4 * v + 2] = __uint_as_float(u.z); out[4 * v + 3] = __uint_as_float(u.w);
  }
}
// P bf16 values (P multiple of 4)
template <int P>
__device__ __forceinline__ void load_q_bf16(const __nv_bfloat16 *p, bool act, float *out) {
#pragma unroll
  for (int v = 0; v < P / 4; ++v) {
    uint2 u = make_uint2(0u, 0u);
    if (act) u = __ldg(reinterpret_cast<const uint2 *>(p) + v);
    out[4 * v + 0] = __uint_as_float(u.x << 16); out[4 * v + 1] = __uint_as_float(u.x & 0xffff0000u);
    out[4 * v + 2] = __uint_as_float(u.y << 16); out[4 * v + 3] = __uint_as_float(u.y & 0xffff0000u);
  }
}

// P fp16 values (P multiple of 4)
template <int P>
__device__ __forceinline__ void load_q_f16(const __half *p, bool act, float *out) {
#pragma unroll
  for (int v = 0; v < P / 4; ++v) {
    uint2 u = make_uint2(0u, 0u);
    if (act) u = __ldg(reinterpret_cast<const uint2 *>(p) + v);
    const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&u.y));
    out[4 * v + 0] = a.x; out[4 * v + 1] = a.y; out[4 * v + 2] = b.x; out[4 * v + 3] = b.y;
  }
}

// 16-bit quotient rounding of the PyTorch expression `half_or_bf16_tensor / int`
template <int PRE>
__device__ __forceinline__ float round_q(float v) {
  if constexpr (PRE == 2) return round_like<__nv_bfloat16>(v);
  else if constexpr (PRE == 3) return round_like<__half>(v);
  else return v;
}

// PRE = 0 inputs of level `lc` from a staged item (kStageItemFloats floats: L*P (x,y) pairs, then L*P weights), P = 4.
// Slots of invalid items hold stale data: every use downstream is gated by `act && inside` (NaN / inf fail the inside
// test, the gathers and reductions are predicated off), so nothing is zeroed here.
__device__ __forceinline__ void level_inputs_staged(const float *sl, int lc, float (&xy)[8], float (&a)[4]) {
  const float4 p0 = *reinterpret_cast<const float4 *>(sl + lc * 8);
  const float4 p1 = *reinterpret_cast<const float4 *>(sl + lc * 8 + 4);
  const float4 w = *reinterpret_cast<const float4 *>(sl + 32 + lc * 4);
  xy[0] = p0.x; xy[1] = p0.y; xy[2] = p0.z; xy[3] = p0.w;
  xy[4] = p1.x; xy[5] = p1.y; xy[6] = p1.z; xy[7] = p1.w;
  a[0] = w.x; a[1] = w.y; a[2] = w.z; a[3] = w.w;
}

// Fills xy[2P] (normalised sampling positions) and a[P] (attention weights) of level `lc` of item `it`.
// `span` = lanes that belong to the item (LPG * GI); all of them must call this together (shuffles).
template <int P, int PRE>
__device__ __forceinline__ void level_inputs(const Problem &pr, const Item &it, int lc, bool act, int H, int W,
                                             int LPG, int span, int j, float (&xy)[2 * P], float (&a)[P], OffsetScale &sc) {
  const long long unit = (long long)it.item * pr.L + lc;       // (item, level) index
  if constexpr (PRE == 0) {
    load_q_f32<2 * P>(static_cast<const float *>(pr.loc) + unit * (2 * P), act, xy);
    load_q_f32<P>(static_cast<const float *>(pr.attn) + unit * P, act, a);
    sc.mulx = sc.muly = sc.divx = sc.divy = 1.f;
  } else {
    float off[2 * P], lg[P];
    if constexpr (PRE == 1) {
      load_q_f32<2 * P>(static_cast<const float *>(pr.loc) + unit * (2 * P), act, off);
      load_q_f32<P>(static_cast<const float *>(pr.attn) + unit * P, act, lg);
    } else if constexpr (PRE == 2) {
      load_q_bf16<2 * P>(static_cast<const __nv_bfloat16 *>(pr.loc) + unit * (2 * P), act, off);
      load_q_bf16<P>(static_cast<const __nv_bfloat16 *>(pr.attn) + unit * P, act, lg);
    } else {
      load_q_f16<2 * P>(static_cast<const __half *>(pr.loc) + unit * (2 * P), act, off);
      load_q_f16<P>(static_cast<const __half *>(pr.attn) + unit * P, act, lg);
    }
    // reference point of (b, q, level): 2 or 4 floats
    const float *rp = pr.ref + ((long long)(it.b * pr.Lq + it.q) * pr.L + lc) * pr.ref_dim;
    float rx = 0.f, ry = 0.f, rw = 0.f, rh = 0.f;
    if (act) {
      if (pr.ref_dim == 2) {
        const float2 r = __ldg(reinterpret_cast<const float2 *>(rp));
        rx = r.x; ry = r.y;
      } else {
        const float4 r = __ldg(reinterpret_cast<const float4 *>(rp));
        rx = r.x; ry = r.y; rw = r.z; rh = r.w;
      }
    }
    if (P == 4 && LPG == 4) {
      // The four lanes of a group hold the same 4 points: lane j does point j's exp, its three IEEE divisions and its
      // rounding, and the results are exchanged with 12 width-4 shuffles -- a quarter of the divisions per lane
      // (the redundant prologue cost the fused kernels +37 % instructions over the plain ones).  Same operations on the
      // same operands as the replicated form below, so coordinates stay bit-identical to the PyTorch expressions.
      float m = -INFINITY;
      if (act) m = fmaxf(fmaxf(lg[0], lg[1]), fmaxf(lg[2], lg[3]));
      for (int d = LPG; d < span; d <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
      const float lgj = j == 0 ? lg[0] : (j == 1 ? lg[1] : (j == 2 ? lg[2] : lg[3]));
      const float ox = j == 0 ? off[0] : (j == 1 ? off[2] : (j == 2 ? off[4] : off[6]));
      const float oy = j == 0 ? off[1] : (j == 1 ? off[3] : (j == 2 ? off[5] : off[7]));
      const float e = act ? expf(lgj - m) : 0.f;
      float s = e;
      for (int d = 1; d < span; d <<= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
      const float my_a = act ? __fdiv_rn(e, s) : 0.f;
      float my_x, my_y;
      if (pr.ref_dim == 2) {
        const float fw = (float)W, fh = (float)H;
        my_x = __fadd_rn(rx, round_q<PRE>(__fdiv_rn(ox, fw)));
        my_y = __fadd_rn(ry, round_q<PRE>(__fdiv_rn(oy, fh)));
        sc.mulx = sc.muly = 1.f; sc.divx = fw; sc.divy = fh;
      } else {
        const float fp = (float)P;
        const float qx = round_q<PRE>(__fdiv_rn(ox, fp)), qy = round_q<PRE>(__fdiv_rn(oy, fp));
        my_x = __fadd_rn(rx, __fmul_rn(__fmul_rn(qx, rw), 0.5f));
        my_y = __fadd_rn(ry, __fmul_rn(__fmul_rn(qy, rh), 0.5f));
        sc.mulx = 0.5f * rw; sc.muly = 0.5f * rh; sc.divx = sc.divy = fp;
      }
#pragma unroll
      for (int p = 0; p < P; ++p) {
        a[p] = __shfl_sync(0xffffffffu, my_a, p, 4);
        xy[2 * p] = __shfl_sync(0xffffffffu, my_x, p, 4);
        xy[2 * p + 1] = __shfl_sync(0xffffffffu, my_y, p, 4);
      }
      return;
    }
    // softmax over the item's L*P logits: the item's groups hold one level each
    float m = -INFINITY;
    if (act) {
#pragma unroll
      for (int p = 0; p < P; ++p) m = fmaxf(m, lg[p]);
    }
    for (int d = LPG; d < span; d <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      a[p] = act ? expf(lg[p] - m) : 0.f;
      s += a[p];
    }
    for (int d = LPG; d < span; d <<= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
#pragma unroll
    for (int p = 0; p < P; ++p) a[p] = act ? __fdiv_rn(a[p], s) : 0.f;

    if (pr.ref_dim == 2) {
      const float fw = (float)W, fh = (float)H;
#pragma unroll
      for (int p = 0; p < P; ++p) {
        xy[2 * p] = __fadd_rn(rx, round_q<PRE>(__fdiv_rn(off[2 * p], fw)));
        xy[2 * p + 1] = __fadd_rn(ry, round_q<PRE>(__fdiv_rn(off[2 * p + 1], fh)));
      }
      sc.mulx = sc.muly = 1.f; sc.divx = fw; sc.divy = fh;
    } else {
      const float fp = (float)P;
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const float qx = round_q<PRE>(__fdiv_rn(off[2 * p], fp)), qy = round_q<PRE>(__fdiv_rn(off[2 * p + 1], fp));
        xy[2 * p] = __fadd_rn(rx, __fmul_rn(__fmul_rn(qx, rw), 0.5f));
        xy[2 * p + 1] = __fadd_rn(ry, __fmul_rn(__fmul_rn(qy, rh), 0.5f));
      }
      sc.mulx = 0.5f * rw; sc.muly = 0.5f * rh; sc.divx = sc.divy = fp;
    }
  }
}

}  // namespace msda
