// msda_bwd.cu -- backward kernels, atomic mode (replaces the six
// ms_deformable_col2im_gpu_kernel_* variants, reference ms_deform_im2col_cuda.cuh:301-920, the
// per-sample routine :87-159 and the launcher :956-1327).
//
// Same item/level/lane mapping as the forward (msda_fwd.cu).  Per sample a lane holds 16 or 32 bytes of
// each of the 4 neighbour rows and of grad_output, and produces
//   * the grad_value contribution  w_k * attn * grad_out  as 16-byte vector reductions
//     (REDG.E.ADD.F32x4, executed by the L2 atomic units) instead of the reference's scalar atomicAdd per
//     channel.  With 8 lanes per row (narrow fp32) one instruction covers 4 whole 128-byte rows; with 4 lanes
//     per row (bf16, wide fp32) two level groups of the same item pair up (COOP below) so that every
//     instruction still covers whole rows -- the L2 atomic units are the bound of this kernel and take
//     whole rows ~20 % faster than the same bytes as half rows.  Zero coefficients send nothing;
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
#include "msda_pre.cuh"

namespace msda {

// ABL: ablation for profiling only (0 = the real kernel, 1 = reductions suppressed, 2 = value gathers suppressed)
// SPEC: the model's own geometry (L = 4 levels, M = 8 heads) as compile-time constants
// MASK: input_padding_mask folded in (fused entry points with a mask only; section 5f of DESIGN.md).  A template parameter
// and not a run-time test: even the never-taken `pad_mask != nullptr` branch per point cost this kernel 6 % (0.750 -> 0.794 ms).
template <typename VT, int D, int P, int MINB, int ABL = 0, int PRE = 0, bool SPEC = false, bool MASK = false>
__global__ void __launch_bounds__(kThreads, MINB) bwd_vec_kernel(const Problem pr_in) {
  pdl_wait();
  Problem pr = pr_in;
  if constexpr (SPEC) { pr.L = 4; pr.M = 8; pr.tile_w_log2 = 3; pr.acc_permuted = 0; }
  constexpr int CH = Chan<VT>::kPerLane;
  constexpr int NP = Chan<VT>::kPairs;
  constexpr int LPG = D / CH;
  constexpr int G = 32 / LPG;
  constexpr int NV = 4 * P;                       // dot products per (item, level)
  constexpr int C = NV >= LPG ? NV / LPG : 1;     // of which a lane owns C after the reduce-scatter
  constexpr bool COOP = (CH == 8 && LPG == 4);    // pairs of level groups send whole accumulator rows
  // sliding-window aggregation of neighbour rows across the points of an (item, level): compiled into the kernels of the
  // model's own geometry only -- in the generic variants its registers spill and cost 10-20 % on uniform inputs
  constexpr bool WIN = SPEC;

  __shared__ Tables tb;
  load_tables(pr, tb, (kThreads / 32) * (G / min(G, pow2_ceil(pr.L))));

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / LPG, j = lane % LPG;
  const int GI = min(G, pow2_ceil(pr.L));
  const int IPW = G / GI;
  const int rounds = (pr.L + GI - 1) / GI;
  const int sub = g / GI, gi = g % GI;
  const int MD = pr.M * D;

  const VT *__restrict__ value = static_cast<const VT *>(pr.value);
  const VT *__restrict__ grad_out = static_cast<const VT *>(pr.grad_out);
  float *__restrict__ gacc = static_cast<float *>(pr.grad_value_acc);

  for (int unit = blockIdx.x; unit < tb.n_units; unit += gridDim.x) {
    const UnitCtx uc = decode_unit(pr, tb, unit);
    for (int k0 = warp * IPW; k0 < tb.unit_items; k0 += (kThreads / 32) * IPW) {
      const Item it = decode_item(pr, tb, uc, k0 + sub);
      const long long item = it.item;

      unsigned long long go[NP];
      Chan<VT>::unpack2(Chan<VT>::gather(grad_out + item * D + j * CH, it.valid), go);

      for (int r = 0; r < rounds; ++r) {
        const int l = r * GI + gi;
        const bool act = it.valid && l < pr.L;
        const int lc = act ? l : 0;
        const int H = tb.H[lc], W = tb.W[lc];

        float xy[2 * P], a[P];
        OffsetScale sc;
        level_inputs<P, PRE>(pr, it, lc, act, H, W, LPG, LPG * GI, j, xy, a, sc);
        const long long row_off = (long long)(it.b * pr.S + tb.start[lc]) * MD + it.m * D;
        const VT *slab = as_given(value + row_off + j * CH);
        // fp32 accumulator row of this head.  fp32 value: natural channel order (it IS grad_value).
        // bf16 value: the row is stored as [first halves of every lane's 8 channels | second halves]
        // (acc_pos() in msda_common.cuh) so each 16-byte reduction of a 4-lane group lands in ONE
        // fully-used 64-byte span instead of half of every 32-byte sector.
        float *gslab = as_given(gacc + row_off + (pr.acc_permuted ? j * 4 : j * CH));
        const int chunk_stride = pr.acc_permuted ? D / 2 : 4;     // second 4-channel chunk of a lane's 8
        const int WMD = W * MD;

        float d[NV];
        float g_lh[P], g_lw[P], g_aw[P];
        // COOP (bf16, 4 lanes per row): two level groups of the same item pair up so that every reduction
        // instruction covers WHOLE 128-byte accumulator rows -- lanes of the first group send channels
        // [8j, 8j+4), lanes of the second [8j+4, 8j+8), first of the first group's row, then of the second's.
        // Both groups hold the same grad_output (same item, same head), so only the row offset and the
        // coefficient travel (two xor-shuffles per neighbour).  The L2 atomic units take whole rows ~20 %
        // faster than the same bytes as two half rows (profiles/r01_microbench_scatter.txt).
        const bool coop = COOP && !pr.acc_permuted && GI >= 2;   // the partner group must belong to the same item
        const bool first = (lane & LPG) == 0;
        const unsigned long long gsel0 = (COOP && !first) ? go[NP / 2] : go[0];
        const unsigned long long gsel1 = (COOP && !first) ? go[NP / 2 + 1] : go[1];
        float *gcoop = as_given(gacc + (long long)(it.b * pr.S) * MD + it.m * D + j * CH + (first ? 0 : 4));
        const int lvl_off = tb.start[lc] * MD;
        // one round of reductions of this group alone: coefficient e[kk] for the row at off + corner(kk); gslab-relative
        auto plain_emit = [&](const float (&e)[4], int off, int wmd) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const int o = off + ((kk & 1) ? MD : 0) + ((kk & 2) ? wmd : 0);
            if (ABL != 1 && e[kk] != 0.f) {
              const unsigned long long c2 = pack_f2(e[kk], e[kk]);
#pragma unroll
              for (int h4 = 0; h4 < NP / 2; ++h4) {
                float r0, r1, r2, r3;
                unpack_f2(fmul2(c2, go[2 * h4]), r0, r1);
                unpack_f2(fmul2(c2, go[2 * h4 + 1]), r2, r3);
                red_add_f32x4(gslab + o + h4 * chunk_stride, r0, r1, r2, r3);
              }
            }
          }
        };
        // pending 2x2 window of neighbour-row coefficients (see the point loop)
        float win_c[4] = {0.f, 0.f, 0.f, 0.f};
        int win_h = 0, win_w = 0, win_off = 0, win_mask = 0;
        bool win_valid = false;
        // one round of pair-cooperative whole-row reductions: coefficient e[kk] for the row at off + corner(kk)
        auto coop_emit = [&](const float (&e)[4], int off, int wmd) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const float my_c = e[kk];
            const int my_off = off + ((kk & 1) ? MD : 0) + ((kk & 2) ? wmd : 0);
            const float o_c = __shfl_xor_sync(0xffffffffu, my_c, LPG);
            const int o_off = __shfl_xor_sync(0xffffffffu, my_off, LPG);
            const float c_a = first ? my_c : o_c, c_b = first ? o_c : my_c;
            const int off_a = first ? my_off : o_off, off_b = first ? o_off : my_off;
            if (ABL != 1 && c_a != 0.f) {
              float r0, r1, r2, r3;
              const unsigned long long c2 = pack_f2(c_a, c_a);
              unpack_f2(fmul2(c2, gsel0), r0, r1);
              unpack_f2(fmul2(c2, gsel1), r2, r3);
              red_add_f32x4(gcoop + off_a, r0, r1, r2, r3);
            }
            if (ABL != 1 && c_b != 0.f) {
              float r0, r1, r2, r3;
              const unsigned long long c2 = pack_f2(c_b, c_b);
              unpack_f2(fmul2(c2, gsel0), r0, r1);
              unpack_f2(fmul2(c2, gsel1), r2, r3);
              red_add_f32x4(gcoop + off_b, r0, r1, r2, r3);
            }
          }
        };
        // SHARE: the 4 lanes of a row group hold the same 4 points; lane j locates point j only (cell, neighbour mask,
        // padding test) and the results travel by width-4 shuffles inside the point loop -- a quarter of the scalar
        // geometry per lane (it was 12 % of this kernel's instructions, all of it computed four times).  Measured: fp32 rows
        // -1 % (cfg-1 2.02 -> 2.00 ms, cfg-3 1.034 -> 1.025 ms per step); 16-bit rows +0.8 % at cfg-2 (752.7 -> 759.0 us)
        // and +3 % at the decoder shape (87.3 -> 90.2 us) in a same-box A/B of two builds -- the six shuffles per point add
        // latency to a kernel that is latency-sensitive at 16 warps per SM -- so it is compiled for fp32 rows only (the
        // forward kernel uses it for every dtype).
        constexpr bool SHARE = SPEC && P == 4 && LPG == 4 && sizeof(VT) == 4;
        float my_lh = 0.f, my_lw = 0.f, my_aw = 0.f;
        int my_h = 0, my_w = 0, my_mask = 0;              // bit 4 of my_mask: the sample is accepted (cuh:288)
        if constexpr (SHARE) {
          const float px = j == 0 ? xy[0] : (j == 1 ? xy[2] : (j == 2 ? xy[4] : xy[6]));
          const float py = j == 0 ? xy[1] : (j == 1 ? xy[3] : (j == 2 ? xy[5] : xy[7]));
          const float pa = j == 0 ? a[0] : (j == 1 ? a[1] : (j == 2 ? a[2] : a[3]));
          const Cell<float> c = locate(px, py, H, W);
          const bool in = act && c.inside;
          int mask = in ? corner_mask(c.h_low, c.w_low, H, W) : 0;
          if constexpr (MASK) {
            if (mask != 0) mask = drop_padded(pr.pad_mask, mask, it.b * pr.S + tb.start[lc] + c.h_low * W + c.w_low, W);
          }
          my_mask = mask | (in ? 16 : 0);
          my_lh = in ? c.lh : 0.f; my_lw = in ? c.lw : 0.f; my_aw = in ? pa : 0.f;
          my_h = c.h_low; my_w = c.w_low;
        }
#pragma unroll
        for (int p = 0; p < P; ++p) {
          Cell<float> c;
          bool in;
          int mask;
          if constexpr (SHARE) {
            const int m5 = __shfl_sync(0xffffffffu, my_mask, p, 4);
            c.h_low = __shfl_sync(0xffffffffu, my_h, p, 4);
            c.w_low = __shfl_sync(0xffffffffu, my_w, p, 4);
            g_lh[p] = __shfl_sync(0xffffffffu, my_lh, p, 4);
            g_lw[p] = __shfl_sync(0xffffffffu, my_lw, p, 4);
            g_aw[p] = __shfl_sync(0xffffffffu, my_aw, p, 4);
            in = (m5 & 16) != 0;
            mask = m5 & 15;
          } else {
            c = locate(xy[2 * p], xy[2 * p + 1], H, W);
            in = act && c.inside;
            mask = in ? corner_mask(c.h_low, c.w_low, H, W) : 0;
            if constexpr (MASK) {
              if (mask != 0) mask = drop_padded(pr.pad_mask, mask, it.b * pr.S + tb.start[lc] + c.h_low * W + c.w_low, W);
            }
            g_lh[p] = in ? c.lh : 0.f;
            g_lw[p] = in ? c.lw : 0.f;
            g_aw[p] = in ? a[p] : 0.f;
          }
          const float hh = 1.f - g_lh[p], hw = 1.f - g_lw[p];
          const float ah = hh * g_aw[p], al = g_lh[p] * g_aw[p];
          const float cw[4] = {ah * hw, ah * g_lw[p], al * hw, al * g_lw[p]};
          const int o00 = in ? c.h_low * WMD + c.w_low * MD : 0;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const bool on = (mask >> kk) & 1;
            const int off = o00 + ((kk & 1) ? MD : 0) + ((kk & 2) ? WMD : 0);
            unsigned long long v[NP];
            Chan<VT>::unpack2(Chan<VT>::gather(slab + off, on && ABL != 2), v);
            unsigned long long dot2 = 0ull;
#pragma unroll
            for (int cidx = 0; cidx < NP; ++cidx) dot2 = ffma2(go[cidx], v[cidx], dot2);
            float dlo, dhi;
            unpack_f2(dot2, dlo, dhi);
            d[4 * p + kk] = dlo + dhi;
          }
          if constexpr (WIN) {
            // Sliding 2x2 window over the points of this (item, level): neighbour rows that the next point's cell
            // shares with the current one (cells at most one pixel apart) are carried in registers and reduced
            // once -- warp-level aggregation before the L2 atomic units.  Points of a head usually walk along a
            // ray, a pixel or less apart on the coarser levels (default-init offsets: 26 % fewer reductions;
            // uniform random locations: none, at a cost of ~40 ALU instructions per point).
            const float ncw[4] = {(mask & 1) ? cw[0] : 0.f, (mask & 2) ? cw[1] : 0.f, (mask & 4) ? cw[2] : 0.f,
                                  (mask & 8) ? cw[3] : 0.f};
            const int dh = c.h_low - win_h, dw = c.w_low - win_w;
            const bool near = in && win_valid && (unsigned)(dh + 1) <= 2u && (unsigned)(dw + 1) <= 2u;
            float carry[4] = {0.f, 0.f, 0.f, 0.f};
            if (p > 0) {
              // coefficient of the old window that lands on corner (a', b') of the new one: old corner (a'+dh, b'+dw)
              const float t0 = dh == 0 ? win_c[0] : (dh == 1 ? win_c[2] : 0.f);
              const float t1 = dh == 0 ? win_c[1] : (dh == 1 ? win_c[3] : 0.f);
              const float t2 = dh == 0 ? win_c[2] : (dh == -1 ? win_c[0] : 0.f);
              const float t3 = dh == 0 ? win_c[3] : (dh == -1 ? win_c[1] : 0.f);
              if (near) {
                carry[0] = dw == 0 ? t0 : (dw == 1 ? t1 : 0.f);
                carry[1] = dw == 0 ? t1 : (dw == -1 ? t0 : 0.f);
                carry[2] = dw == 0 ? t2 : (dw == 1 ? t3 : 0.f);
                carry[3] = dw == 0 ? t3 : (dw == -1 ? t2 : 0.f);
              }
              // old corner (a, b) stays in the window iff a - dh and b - dw are both in {0, 1}
              const bool r0 = near && dh <= 0, r1 = near && dh >= 0, c0 = dw <= 0, c1 = dw >= 0;
              const float e[4] = {(r0 && c0) ? 0.f : win_c[0], (r0 && c1) ? 0.f : win_c[1],
                                  (r1 && c0) ? 0.f : win_c[2], (r1 && c1) ? 0.f : win_c[3]};
              if (coop) coop_emit(e, lvl_off + win_off, WMD);
              else plain_emit(e, win_off, WMD);
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) win_c[kk] = ncw[kk] + carry[kk];
            win_h = c.h_low; win_w = c.w_low; win_valid = in; win_off = o00;
          } else if (coop) {
            const float e[4] = {(mask & 1) ? cw[0] : 0.f, (mask & 2) ? cw[1] : 0.f, (mask & 4) ? cw[2] : 0.f,
                                (mask & 8) ? cw[3] : 0.f};
            coop_emit(e, lvl_off + o00, WMD);
          } else if (p > 0 && in && mask == win_mask && o00 == win_off) {
            // same cell <=> same offset AND same neighbour mask: (h, W-1) and (h+1, -1) share the
            // row-major offset but never the mask.  Consecutive points in one cell are summed in registers.
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) win_c[kk] += cw[kk];
          } else {
            if (p > 0) {
              const float e[4] = {(win_mask & 1) ? win_c[0] : 0.f, (win_mask & 2) ? win_c[1] : 0.f,
                                  (win_mask & 4) ? win_c[2] : 0.f, (win_mask & 8) ? win_c[3] : 0.f};
              plain_emit(e, win_off, WMD);
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) win_c[kk] = cw[kk];
            win_off = o00;
            win_mask = mask;
          }
        }
        if constexpr (WIN) {
          if (coop) coop_emit(win_c, lvl_off + win_off, WMD);
          else plain_emit(win_c, win_off, WMD);
        } else if (!coop) {
          const float e[4] = {(win_mask & 1) ? win_c[0] : 0.f, (win_mask & 2) ? win_c[1] : 0.f,
                              (win_mask & 4) ? win_c[2] : 0.f, (win_mask & 8) ? win_c[3] : 0.f};
          plain_emit(e, win_off, WMD);
        }

        reduce_scatter<NV, LPG / 2>(d, j);

        // Which point(s) this lane now owns, and with which of the four dot products.
        constexpr int NOWN = C >= 4 ? C / 4 : 1;      // points owned by this lane
        constexpr int LPP = C >= 4 ? 1 : 4 / C;       // lanes sharing one point
        const int part = j % LPP;
        float o_gx[NOWN], o_gy[NOWN], o_gw[NOWN], o_a[NOWN];
        int o_p[NOWN];
        if constexpr (C >= 4) {
#pragma unroll
          for (int t = 0; t < NOWN; ++t) {
            const int p = j * NOWN + t;
            float lh = 0.f, lw = 0.f, aw = 0.f, araw = 0.f;
#pragma unroll
            for (int pp = 0; pp < P; ++pp)
              if (pp == p) { lh = g_lh[pp]; lw = g_lw[pp]; aw = g_aw[pp]; araw = a[pp]; }
            const float hh = 1.f - lh, hw = 1.f - lw;
            const float d1 = d[4 * t], d2 = d[4 * t + 1], d3 = d[4 * t + 2], d4 = d[4 * t + 3];
            o_gx[t] = (float)W * aw * (hh * (d2 - d1) + lh * (d4 - d3));
            o_gy[t] = (float)H * aw * (hw * (d3 - d1) + lw * (d4 - d2));
            o_gw[t] = hh * hw * d1 + hh * lw * d2 + lh * hw * d3 + lh * lw * d4;
            o_a[t] = araw;
            o_p[t] = p;
          }
        } else {
          // C == 2: lanes (2p, 2p+1) hold (d1,d2) / (d3,d4);  C == 1: lanes 4p..4p+3 hold d1..d4.
          const int p = j / LPP;
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
          float lh = 0.f, lw = 0.f, aw = 0.f, araw = 0.f;
#pragma unroll
          for (int pp = 0; pp < P; ++pp)
            if (pp == p) { lh = g_lh[pp]; lw = g_lw[pp]; aw = g_aw[pp]; araw = a[pp]; }
          const float hh = 1.f - lh, hw = 1.f - lw;
          o_gx[0] = (float)W * aw * (hh * (d2 - d1) + lh * (d4 - d3));
          o_gy[0] = (float)H * aw * (hw * (d3 - d1) + lw * (d4 - d2));
          o_gw[0] = hh * hw * d1 + hh * lw * d2 + lh * hw * d3 + lh * lw * d4;
          o_a[0] = araw;
          o_p[0] = p;
        }

        const long long lvl = item * pr.L + lc;
        if constexpr (PRE == 0) {
          float *gl = static_cast<float *>(pr.grad_loc) + lvl * (2 * P);
          float *ga = static_cast<float *>(pr.grad_attn) + lvl * P;
          if (act) {
#pragma unroll
            for (int t = 0; t < NOWN; ++t) {
              if constexpr (C >= 4) {
                *reinterpret_cast<float2 *>(gl + 2 * o_p[t]) = make_float2(o_gx[t], o_gy[t]);
                ga[o_p[t]] = o_gw[t];
              } else {
                if (part == 0) gl[2 * o_p[t]] = o_gx[t];
                if (part == 1) gl[2 * o_p[t] + 1] = o_gy[t];
                if (part == LPP - 1) ga[o_p[t]] = o_gw[t];
              }
            }
          }
        } else {
          // softmax backward needs S = sum over the item's L*P samples of a * d(loss)/da; each point is
          // counted by one lane, then summed over the item's lanes
          float part_s = 0.f;
#pragma unroll
          for (int t = 0; t < NOWN; ++t)
            if (act && part == 0) part_s += o_a[t] * o_gw[t];
          for (int dd = 1; dd < LPG * GI; dd <<= 1) part_s += __shfl_xor_sync(0xffffffffu, part_s, dd);
          if (act) {
#pragma unroll
            for (int t = 0; t < NOWN; ++t) {
              const float gox = __fdiv_rn(o_gx[t] * sc.mulx, sc.divx), goy = __fdiv_rn(o_gy[t] * sc.muly, sc.divy);
              const float glg = o_a[t] * (o_gw[t] - part_s);
              if constexpr (PRE == 1) {
                float *go_ = static_cast<float *>(pr.grad_loc) + lvl * (2 * P);
                float *gg_ = static_cast<float *>(pr.grad_attn) + lvl * P;
                if (part == 0) go_[2 * o_p[t]] = gox;
                if (part == (LPP > 1 ? 1 : 0)) go_[2 * o_p[t] + 1] = goy;
                if (part == LPP - 1) gg_[o_p[t]] = glg;
              } else if constexpr (PRE == 2) {
                __nv_bfloat16 *go_ = static_cast<__nv_bfloat16 *>(pr.grad_loc) + lvl * (2 * P);
                __nv_bfloat16 *gg_ = static_cast<__nv_bfloat16 *>(pr.grad_attn) + lvl * P;
                if (part == 0) go_[2 * o_p[t]] = __float2bfloat16_rn(gox);
                if (part == (LPP > 1 ? 1 : 0)) go_[2 * o_p[t] + 1] = __float2bfloat16_rn(goy);
                if (part == LPP - 1) gg_[o_p[t]] = __float2bfloat16_rn(glg);
              } else {
                __half *go_ = static_cast<__half *>(pr.grad_loc) + lvl * (2 * P);
                __half *gg_ = static_cast<__half *>(pr.grad_attn) + lvl * P;
                if (part == 0) go_[2 * o_p[t]] = __float2half_rn(gox);
                if (part == (LPP > 1 ? 1 : 0)) go_[2 * o_p[t] + 1] = __float2half_rn(goy);
                if (part == LPP - 1) gg_[o_p[t]] = __float2half_rn(glg);
              }
            }
          }
        }
      }
    }
  }
}

template <typename VT, int D, int P>
static int launch_bwd_vec(const Problem &pr, int grid, int minb, cudaStream_t stream) {
  if constexpr (D == 32 && P == 4) {
    const bool spec = pr.L == 4 && pr.M == 8;
#define MSDA_FUSED(CODE, PREV)                                                                          \
    if (minb == CODE) {                                                                                   \
      if (spec && pr.pad_mask != nullptr) launch_pdl(bwd_vec_kernel<VT, D, P, 2, 0, PREV, true, true>, grid, kThreads, 0, stream, pr);   \
      else if (pr.pad_mask != nullptr) return MSDA_E_UNSUPPORTED;                                        \
      else if (spec) launch_pdl(bwd_vec_kernel<VT, D, P, 2, 0, PREV, true>, grid, kThreads, 0, stream, pr);      \
      else launch_pdl(bwd_vec_kernel<VT, D, P, 2, 0, PREV>, grid, kThreads, 0, stream, pr);                      \
      return (int)cudaGetLastError();                                                                     \
    }
    MSDA_FUSED(201, 1) MSDA_FUSED(202, 2) MSDA_FUSED(203, 3)
#undef MSDA_FUSED
#ifdef MSDA_BUILD_VARIANTS      // profiling ablations (variant bits 6..7)
    if (minb == 101) { launch_pdl(bwd_vec_kernel<VT, D, P, 2, 1>, grid, kThreads, 0, stream, pr); return (int)cudaGetLastError(); }
    if (minb == 102) { launch_pdl(bwd_vec_kernel<VT, D, P, 2, 2>, grid, kThreads, 0, stream, pr); return (int)cudaGetLastError(); }
#endif
  }
  if (minb >= 100) return MSDA_E_UNSUPPORTED;
  if constexpr (D == 32 && P == 4) {
    if (minb == 2 && pr.L == 4 && pr.M == 8) {
      launch_pdl(bwd_vec_kernel<VT, D, P, 2, 0, 0, true>, grid, kThreads, 0, stream, pr);
      return (int)cudaGetLastError();
    }
  }
#ifdef MSDA_BUILD_VARIANTS      // occupancy experiments (variant bits 4..5): 3 and 4 CTAs per SM spill and lose (tools/sweep.py)
  if (minb == 4) { launch_pdl(bwd_vec_kernel<VT, D, P, 4>, grid, kThreads, 0, stream, pr); return (int)cudaGetLastError(); }
  if (minb == 3) { launch_pdl(bwd_vec_kernel<VT, D, P, 3>, grid, kThreads, 0, stream, pr); return (int)cudaGetLastError(); }
#endif
  if (minb != 2) return MSDA_E_UNSUPPORTED;
  launch_pdl(bwd_vec_kernel<VT, D, P, 2>, grid, kThreads, 0, stream, pr);
  return (int)cudaGetLastError();
}

template <typename VT>
static int dispatch_bwd(const Problem &pr, int grid, int minb, cudaStream_t stream) {
#define MSDA_CASE(DD, PP) \
  if (pr.D == DD && pr.P == PP) return launch_bwd_vec<VT, DD, PP>(pr, grid, minb, stream);
  MSDA_CASE(32, 4) MSDA_CASE(32, 8) MSDA_CASE(16, 4) MSDA_CASE(16, 8) MSDA_CASE(64, 4) MSDA_CASE(64, 8)
#undef MSDA_CASE
  return MSDA_E_UNSUPPORTED;
}

int launch_backward_vec(const Problem &pr, int dtype, int grid, int minb, cudaStream_t stream) {
  if (dtype == MSDA_F32 && pr.D >= 32) return dispatch_bwd<wide_f32>(pr, grid, minb, stream);   // 8 channels per lane
  if (dtype == MSDA_F32) return dispatch_bwd<float>(pr, grid, minb, stream);
  if (dtype == MSDA_BF16) return dispatch_bwd<__nv_bfloat16>(pr, grid, minb, stream);
  if (dtype == MSDA_F16) return dispatch_bwd<__half>(pr, grid, minb, stream);
  return MSDA_E_UNSUPPORTED;
}

// ---- fp32 accumulator -> bf16 ------------------------------------------------------------------
// PERM: the accumulator rows use the permuted channel order of acc_pos() (bf16 value, D % 8 == 0).
// One thread produces 8 consecutive bf16 channels (16 bytes) from two 16-byte accumulator chunks.
template <bool PERM, typename HT>
__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float *__restrict__ src, uint4 *__restrict__ dst,
                                                         long long n8, int D) {
  const int lanes = D / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    float4 a, b;
    if (PERM) {
      const long long row = i / lanes;
      const int j = (int)(i - row * lanes);
      const float *r = src + row * D;
      a = *reinterpret_cast<const float4 *>(r + 4 * j);
      b = *reinterpret_cast<const float4 *>(r + D / 2 + 4 * j);
    } else {
      a = *reinterpret_cast<const float4 *>(src + 8 * i);
      b = *reinterpret_cast<const float4 *>(src + 8 * i + 4);
    }
    dst[i] = make_uint4(Chan<HT>::pack2(a.x, a.y), Chan<HT>::pack2(a.z, a.w), Chan<HT>::pack2(b.x, b.y),
                        Chan<HT>::pack2(b.z, b.w));
  }
}

template <typename HT>
__global__ void f32_to_bf16_scalar_kernel(const float *__restrict__ src, HT *__restrict__ dst, long long first,
                                          long long n) {
  for (long long i = first + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    if constexpr (sizeof(HT) == 2 && !__is_same(HT, __half)) dst[i] = __float2bfloat16_rn(src[i]);
    else dst[i] = __float2half_rn(src[i]);
  }
}

template <typename HT>
static int launch_narrow(const float *src, void *dst, long long n, int D, bool permuted, cudaStream_t stream) {
  const bool vec = (D % 8 == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) &&
                   ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
  if (vec) {
    const long long n8 = n / 8;
    const long long blocks = (n8 + 255) / 256;
    const int grid = (int)(blocks < 148LL * 16 ? blocks : 148LL * 16);
    if (permuted) f32_to_bf16_kernel<true, HT><<<grid, 256, 0, stream>>>(src, static_cast<uint4 *>(dst), n8, D);
    else f32_to_bf16_kernel<false, HT><<<grid, 256, 0, stream>>>(src, static_cast<uint4 *>(dst), n8, D);
  } else {
    if (permuted) return MSDA_E_UNSUPPORTED;
    const long long blocks = (n + 255) / 256;
    const int grid = (int)(blocks < 148LL * 16 ? blocks : 148LL * 16);
    f32_to_bf16_scalar_kernel<HT><<<grid, 256, 0, stream>>>(src, static_cast<HT *>(dst), 0, n);
  }
  return (int)cudaGetLastError();
}

// fp32 accumulator -> 16-bit grad_value (`half` selects fp16, otherwise bf16)
int launch_f32_to_bf16(const float *src, void *dst, long long n, int D, bool permuted, bool half, cudaStream_t stream) {
  return half ? launch_narrow<__half>(src, dst, n, D, permuted, stream)
              : launch_narrow<__nv_bfloat16>(src, dst, n, D, permuted, stream);
}

}  // namespace msda
