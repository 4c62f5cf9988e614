// msda_fwd.cu -- forward gather kernels (replaces ms_deformable_im2col_gpu_kernel,
// reference ms_deform_im2col_cuda.cuh:237-299, and its launcher :923-954).
//
// Mapping.  A (query, head) "item" owns D output channels.  A lane owns 16 bytes of them (4 fp32 or
// 8 bf16), so LPG = D*sizeof(VT)/16 lanes cover one head and one warp instruction gathers 32/LPG
// whole 128-byte (fp32) or 64-byte (bf16) head rows.  The groups of a warp are spread over the
// pyramid levels of an item: each group preloads the P (x,y) pairs and P attention weights of its
// level with two or three 16-byte loads, computes the bilinear geometry in registers and issues the
// 4*P corner gathers back to back; the accumulation is packed fp32x2 (FFMA2); the per-level partial
// sums are combined with xor-shuffles and one lane group stores the 16-byte chunks of the output
// row.  Nothing is written twice, so the output needs no memset (the reference zero-fills it first,
// ms_deform_attn_cuda.cu:54).
//
// Work order.  CTAs walk "units" of 64 items with a grid-stride loop.  For encoder self-attention
// (one query per pixel) a unit is an 8x8 tile of queries of one level and one head, so the gathers
// of a CTA overlap in L1; otherwise units are consecutive items.  The choice is made on the device
// from the int64 shape table (no host read of device memory).
#include "msda_common.cuh"
#include "msda_launch.h"
#include "msda_pre.cuh"

namespace msda {

// PB = points whose 4 gathers are in flight together (register staging: PB*16 registers)
// SPEC: the model's own geometry (L = 4 levels, M = 8 heads) as compile-time constants
// ST: stage the next unit's positions / weights with cp.async (model-geometry kernels in linear query order)
// MASK: input_padding_mask folded in (fused entry points with a mask only); a template parameter, see msda_bwd.cu
template <typename VT, int D, int P, int MINB, int PB, int PRE = 0, bool SPEC = false, bool ST = false, bool MASK = false>
__global__ void __launch_bounds__(kThreads, MINB) fwd_vec_kernel(const Problem pr_in) {
  pdl_wait();
  Problem pr = pr_in;
  if constexpr (SPEC) { pr.L = 4; pr.M = 8; pr.tile_w_log2 = 3; }
  constexpr int CH = Chan<VT>::kPerLane;   // channels per lane
  constexpr int NP = Chan<VT>::kPairs;     // fp32x2 accumulators per lane
  constexpr int LPG = D / CH;              // lanes per (item, level) group
  constexpr int G = 32 / LPG;              // groups per warp
  static_assert(D % CH == 0 && LPG >= 1 && LPG <= 32 && (LPG & (LPG - 1)) == 0, "unsupported head width");

  __shared__ Tables tb;
  load_tables(pr, tb, (kThreads / 32) * (G / min(G, pow2_ceil(pr.L))));

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / LPG, j = lane % LPG;
  const int GI = min(G, pow2_ceil(pr.L));  // groups cooperating on one item
  const int IPW = G / GI;                  // items per warp pass
  const int rounds = (pr.L + GI - 1) / GI;
  const int sub = g / GI, gi = g % GI;
  const int MD = pr.M * D;

  const VT *__restrict__ value = static_cast<const VT *>(pr.value);
  VT *__restrict__ out = static_cast<VT *>(pr.out);

  // positions / weights of the next unit are staged by cp.async (msda_common.cuh).  Launched for the linear query order
  // only (decoder-like shapes, latency-bound: -9 %); in query-tile order the kernel is issue-bound and the ~6 % extra
  // instructions cancel the shorter stalls (measured +0.5 %).
  constexpr bool STAGE = ST && SPEC && PRE == 0 && P == 4 && G == 8;
  __shared__ __align__(16) float stage[STAGE ? 2 * (kThreads / 32) * kStageSlots * kStageItemFloats : 4];
  auto stage_slot = [&](int buf, int slot) {
    return stage + ((buf * (kThreads / 32) + warp) * kStageSlots + slot) * kStageItemFloats;
  };
  auto stage_issue = [&](int u, int buf) {
    const UnitCtx nu = decode_unit(pr, tb, u);
    const int s = lane >> 2;                                             // slot = pass * IPW + sub
    const int k = (s / IPW) * ((kThreads / 32) * IPW) + warp * IPW + (s % IPW);
    const Item ni = decode_item(pr, tb, nu, k);
    if (ni.valid && k < tb.unit_items) {
      const float *gl = static_cast<const float *>(pr.loc) + (long long)ni.item * 32;
      const float *ga = static_cast<const float *>(pr.attn) + (long long)ni.item * 16;
      float *dst = stage_slot(buf, s);
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const int c = (lane & 3) + 4 * t;                                // 12 chunks of 16 bytes per item
        cp_async16(dst + c * 4, c < 8 ? gl + c * 4 : ga + (c - 8) * 4);
      }
    }
    cp_async_commit();
  };
  int buf = 0;
  if constexpr (STAGE) {
    if ((int)blockIdx.x < tb.n_units) stage_issue(blockIdx.x, 0);
  }

  // passes per unit (1 in linear order).  Rounded up: with 16 items per warp pass (16-bit rows, D = 16, L = 1) one pass
  // of the 8 warps covers 128 > 64 items, and the items beyond the unit are masked below.
  const int ipp = (kThreads / 32) * IPW;
  const int ppu = (tb.unit_items + ipp - 1) / ipp;
  const int n_work = work_count(tb, ppu);
  for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
    const WorkItem wi = work_item(tb, ppu, w);
    const int unit = wi.unit;
    const UnitCtx uc = decode_unit(pr, tb, unit);
    if constexpr (STAGE) {                                     // linear order: one pass per unit, so w == unit
      if (unit + (int)gridDim.x < tb.n_units) {
        stage_issue(unit + gridDim.x, buf ^ 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncwarp();
    }
    for (int pass = wi.pass_begin; pass < wi.pass_end; ++pass) {
      const int k0 = warp * IPW + pass * (kThreads / 32) * IPW;
      Item it = decode_item(pr, tb, uc, k0 + sub);
      it.valid = it.valid && (k0 + sub < tb.unit_items);

      unsigned long long acc[NP];
#pragma unroll
      for (int c = 0; c < NP; ++c) acc[c] = 0ull;

      for (int r = 0; r < rounds; ++r) {
        const int l = r * GI + gi;
        const bool act = it.valid && l < pr.L;
        const int lc = act ? l : 0;
        const int H = tb.H[lc], W = tb.W[lc];

        // (x,y) pairs and weights of this level (given, or derived from raw offsets/logits: msda_pre.cuh)
        float xy[2 * P], a[P];
        OffsetScale sc;
        if constexpr (STAGE) {
          level_inputs_staged(stage_slot(buf, (k0 / ((kThreads / 32) * IPW)) * IPW + sub), lc, xy, a);
          sc.mulx = sc.muly = sc.divx = sc.divy = 1.f;
        } else {
          level_inputs<P, PRE>(pr, it, lc, act, H, W, LPG, LPG * GI, j, xy, a, sc);
        }
        const VT *slab = value + (long long)(it.b * pr.S + tb.start[lc]) * MD + (it.m * D + j * CH);
        const int WMD = W * MD;

        // SHARE: the 4 lanes of a row group hold the same 4 points; lane j locates point j only and computes its four
        // weights, the results travel by width-4 shuffles.  Same-box A/B (two builds, tools/sweep.py): decoder forward
        // bf16 33.3 -> 31.4 us, fp32 46.2 -> 45.9 us per layer.
        constexpr bool SHARE = SPEC && P == 4 && LPG == 4;
        float my_w[4] = {0.f, 0.f, 0.f, 0.f};
        int my_off = 0, my_mask = 0;
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
          const float aw = in ? pa : 0.f;
          const float lh = in ? c.lh : 0.f, lw = in ? c.lw : 0.f;      // keeps NaN coordinates out
          const float hh = 1.f - lh, hw = 1.f - lw;
          const float ah = hh * aw, al = lh * aw;
          my_w[0] = ah * hw; my_w[1] = ah * lw; my_w[2] = al * hw; my_w[3] = al * lw;
          my_off = c.h_low * WMD + c.w_low * MD;
          my_mask = mask;
        }
#pragma unroll
        for (int p0 = 0; p0 < P; p0 += PB) {
          typename Chan<VT>::Raw raw[PB][4];
          float wt[PB][4];
#pragma unroll
          for (int pp = 0; pp < PB; ++pp) {
            const int p = p0 + pp;
            int mask;
            const VT *p00;
            if constexpr (SHARE) {
              mask = __shfl_sync(0xffffffffu, my_mask, p, 4);
              p00 = slab + __shfl_sync(0xffffffffu, my_off, p, 4);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) wt[pp][kk] = __shfl_sync(0xffffffffu, my_w[kk], p, 4);
            } else {
              const Cell<float> c = locate(xy[2 * p], xy[2 * p + 1], H, W);
              const bool in = act && c.inside;
              mask = in ? corner_mask(c.h_low, c.w_low, H, W) : 0;
              if constexpr (MASK) {
                if (mask != 0) mask = drop_padded(pr.pad_mask, mask, it.b * pr.S + tb.start[lc] + c.h_low * W + c.w_low, W);
              }
              const float aw = in ? a[p] : 0.f;
              const float lh = in ? c.lh : 0.f, lw = in ? c.lw : 0.f;    // keeps NaN coordinates out
              const float hh = 1.f - lh, hw = 1.f - lw;
              const float ah = hh * aw, al = lh * aw;
              wt[pp][0] = ah * hw; wt[pp][1] = ah * lw; wt[pp][2] = al * hw; wt[pp][3] = al * lw;
              p00 = slab + (c.h_low * WMD + c.w_low * MD);
            }
            // fp32 only: measured +6-14 % there, -5-10 % for bf16 (the branch costs more than the zero fill)
            if (sizeof(VT) == 4 && mask == 15) {   // interior cell (the common case): plain loads, no predicates
              raw[pp][0] = Chan<VT>::gather_plain(p00);
              raw[pp][1] = Chan<VT>::gather_plain(p00 + MD);
              raw[pp][2] = Chan<VT>::gather_plain(p00 + WMD);
              raw[pp][3] = Chan<VT>::gather_plain(p00 + WMD + MD);
            } else {
              raw[pp][0] = Chan<VT>::gather(p00, mask & 1);
              raw[pp][1] = Chan<VT>::gather(p00 + MD, mask & 2);
              raw[pp][2] = Chan<VT>::gather(p00 + WMD, mask & 4);
              raw[pp][3] = Chan<VT>::gather(p00 + WMD + MD, mask & 8);
            }
          }
#pragma unroll
          for (int pp = 0; pp < PB; ++pp) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              unsigned long long v[NP];
              Chan<VT>::unpack2(raw[pp][kk], v);
              const unsigned long long w2 = pack_f2(wt[pp][kk], wt[pp][kk]);
#pragma unroll
              for (int c = 0; c < NP; ++c) acc[c] = ffma2(w2, v[c], acc[c]);
            }
          }
        }
      }

      float res[CH];
#pragma unroll
      for (int c = 0; c < NP; ++c) unpack_f2(acc[c], res[2 * c], res[2 * c + 1]);
      // combine the level groups of each item
      for (int d = LPG; d < LPG * GI; d <<= 1) {
#pragma unroll
        for (int c = 0; c < CH; ++c) res[c] += __shfl_xor_sync(0xffffffffu, res[c], d);
      }
      if (it.valid && gi == 0) {
        Chan<VT>::store(out + (long long)it.item * D + j * CH, Chan<VT>::pack(res));
      }
    }
    if constexpr (STAGE) {
      __syncwarp();          // every lane is done with this buffer before the next unit's copies land in it
      buf ^= 1;
    }
  }
}

template <typename VT, int D, int P>
static int launch_fwd_vec(const Problem &pr, int grid, int minb, cudaStream_t stream) {
  // points staged together: PB * 4 gathers * (16 or 32 bytes) of registers per lane
  constexpr int WR = sizeof(typename Chan<VT>::Raw) / 16;
  constexpr int PB2 = 4 / WR, PB3 = 2 / WR;
  if constexpr (D == 32 && P == 4) {
    const bool spec = pr.L == 4 && pr.M == 8;
#define MSDA_FUSED(CODE, PREV)                                                                          \
    if (minb == CODE) {                                                                                   \
      if (spec && pr.pad_mask != nullptr) launch_pdl(fwd_vec_kernel<VT, D, P, 3, PB3, PREV, true, false, true>, grid, kThreads, 0, stream, pr);   \
      else if (pr.pad_mask != nullptr) return MSDA_E_UNSUPPORTED;                                        \
      else if (spec) launch_pdl(fwd_vec_kernel<VT, D, P, 3, PB3, PREV, true>, grid, kThreads, 0, stream, pr);    \
      else launch_pdl(fwd_vec_kernel<VT, D, P, 3, PB3, PREV>, grid, kThreads, 0, stream, pr);                    \
      return (int)cudaGetLastError();                                                                     \
    }
    MSDA_FUSED(201, 1) MSDA_FUSED(202, 2) MSDA_FUSED(203, 3)
#undef MSDA_FUSED
  }
  if (minb > 200) return MSDA_E_UNSUPPORTED;
  if constexpr (D == 32 && P == 4) {
    if (minb == 3 && pr.L == 4 && pr.M == 8) {
      if (pr.tiled) launch_pdl(fwd_vec_kernel<VT, D, P, 3, PB3, 0, true, false>, grid, kThreads, 0, stream, pr);
      else launch_pdl(fwd_vec_kernel<VT, D, P, 3, PB3, 0, true, true>, grid, kThreads, 0, stream, pr);
      return (int)cudaGetLastError();
    }
  }
#ifdef MSDA_BUILD_VARIANTS      // occupancy experiments (variant bits 4..5); the product build carries the measured best only
  if (minb == 4) { launch_pdl(fwd_vec_kernel<VT, D, P, 4, 1>, grid, kThreads, 0, stream, pr); return (int)cudaGetLastError(); }
  if (minb == 2) { launch_pdl(fwd_vec_kernel<VT, D, P, 2, PB2>, grid, kThreads, 0, stream, pr); return (int)cudaGetLastError(); }
#else
  (void)PB2;
#endif
  if (minb != 3) return MSDA_E_UNSUPPORTED;
  launch_pdl(fwd_vec_kernel<VT, D, P, 3, PB3>, grid, kThreads, 0, stream, pr);
  return (int)cudaGetLastError();
}

template <typename VT>
static int dispatch_fwd(const Problem &pr, int grid, int minb, cudaStream_t stream) {
#define MSDA_CASE(DD, PP) \
  if (pr.D == DD && pr.P == PP) return launch_fwd_vec<VT, DD, PP>(pr, grid, minb, stream);
  MSDA_CASE(32, 4) MSDA_CASE(32, 8) MSDA_CASE(16, 4) MSDA_CASE(16, 8) MSDA_CASE(64, 4) MSDA_CASE(64, 8)
#undef MSDA_CASE
  return MSDA_E_UNSUPPORTED;
}

bool fwd_vec_supported(int dtype, int D, int P, int L) {
  if (dtype != MSDA_F32 && dtype != MSDA_BF16 && dtype != MSDA_F16) return false;
  if (!(D == 16 || D == 32 || D == 64)) return false;
  if (!(P == 4 || P == 8)) return false;
  return L >= 1 && L <= kMaxLevels;
}

int launch_forward_vec(const Problem &pr, int dtype, int grid, int minb, cudaStream_t stream) {
  if (dtype == MSDA_F32 && pr.D >= 32) return dispatch_fwd<wide_f32>(pr, grid, minb, stream);   // 8 channels per lane
  if (dtype == MSDA_F32) return dispatch_fwd<float>(pr, grid, minb, stream);
  if (dtype == MSDA_BF16) return dispatch_fwd<__nv_bfloat16>(pr, grid, minb, stream);
  if (dtype == MSDA_F16) return dispatch_fwd<__half>(pr, grid, minb, stream);
  return MSDA_E_UNSUPPORTED;
}

}  // namespace msda
