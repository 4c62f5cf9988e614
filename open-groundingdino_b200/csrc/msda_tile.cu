// msda_tile.cu -- query-tile kernels for encoder self-attention (one query per pyramid pixel), D = 32, P = 4, L <= 4.
// Replace ms_deformable_im2col_gpu_kernel (reference ms_deform_im2col_cuda.cuh:237-299) and the six
// ms_deformable_col2im_gpu_kernel_* variants (:301-920) on the shapes the model's encoder runs.
//
// Why a second kernel family.  The vector kernels of msda_fwd.cu / msda_bwd.cu give every lane of a 4-lane row group
// the whole per-sample scalar work (coordinates, neighbour masks, bilinear weights, addresses): computed four times,
// it was two thirds of their instructions (ncu: 655 warp instructions per 16 row gathers forward, 1426 backward), and
// the backward sent 25-35 M row reductions per launch to the L2 atomic units (67 % of their measured ceiling).  Here a
// work unit -- an 8x8 tile of queries of one level and one head -- is processed in phases by one CTA:
//
//   A  geometry, one thread per (query, level): 4 samples x 4 neighbours -> 16 records {row offset | flags, weight} in
//      shared memory.  Every scalar is computed once.
//   1  row loop: a lane owns 8 channels (16 bytes of a 16-bit row, 32 of an fp32 row), 4 lanes a row, so one warp
//      instruction gathers 8 rows; a record read (one LDS) replaces the geometry.  Forward: weighted accumulation,
//      two queries per warp so the neighbour groups combine with two shuffle stages.  Backward: the four dot products
//      <grad_out, row> per sample go to shared memory for phase C.
//   2  (backward) grad_value: during phase A every neighbour with a non-zero coefficient that falls into the tile's
//      AGGREGATION WINDOW (the tile's footprint on each level plus a halo) was pushed on a per-row linked list
//      (one ATOMS.EXCH on the row's head word).  Now groups of 8 lanes walk the lists, sum coefficient x grad_out in
//      registers (grad_out rows of the 64 queries are staged in shared memory) and send ONE red.global.add.v4.f32 row
//      per touched window row -- adjacent queries of a tile share most of their neighbours, so this is 3-6x fewer
//      L2 reductions than per-sample scatter, with no shared-memory accumulators and no shared-memory atomics on
//      floats (they are CAS loops on sm_100: profiles/r01_microbench_scatter.txt).  Neighbours outside the window
//      (uniform random locations, large learned offsets) are reduced directly from the row loop, whole 128-byte rows
//      per instruction (group pairs cooperate as in msda_bwd.cu).
//   C  (backward) grad_sampling_loc / grad_attn_weight, one thread per (query, level), plain 16-byte stores.
//
// Three kernels live here: fwd_tile_kernel (phases A + 1, the default forward of these shapes), bwd_warp_kernel (the default
// backward: phases A + 1 + C at warp scope, per-query reduction lists instead of phase 2, see its own header below) and
// bwd_tile_kernel (all four phases with the per-tile aggregation of phase 2; opt-in, measured slower).
//
// The window placement assumes what encoder self-attention does -- query i is pixel i of its level and samples near
// its own position on every level (models/GroundingDINO/transformer.py:470-485) -- but only performance depends on it.
#include "msda_common.cuh"
#include "msda_launch.h"
#include "msda_pre.cuh"

namespace msda {

namespace {

constexpr int kTileItems = kUnitItems;        // 64 queries (x one head) per unit
constexpr int kRecStride = 17;                // 16 records per (query, level) thread + 1 pad word: conflict-free writes and reads
constexpr int kRecs = kThreads * kRecStride;  // 4352
constexpr int kTablesBytes = 512;
constexpr unsigned kNoLoad = 0x40000000u;     // record flag: no row behind this record (outside the image, padding, ...)
constexpr unsigned kOffMask = 0x3fffffffu;

static_assert(sizeof(Tables) <= kTablesBytes, "Tables outgrew its shared-memory slot");

// Per-thread inputs of one (query, level): normalised positions and attention weights of its P = 4 samples.
// PRE = 0: given (the op boundary).  PRE = 1/2/3: derived from raw offsets / logits (fp32 / bf16 / fp16) and the reference
// points like MSDeformAttn.forward does (msda_pre.cuh states the arithmetic contract).  The four threads of a query
// (consecutive lanes) hold its levels; the softmax statistics cross them with two xor-shuffles.  Every thread of the
// warp must call this (full-mask shuffles).
template <int PRE>
__device__ __forceinline__ void thread_level_inputs(const Problem &pr, const Item &it, int lc, bool act, int H, int W,
                                                    float (&xy)[8], float (&a)[4], OffsetScale &sc) {
  const long long u = (long long)it.item * pr.L + lc;
  if constexpr (PRE == 0) {
    load_q_f32<8>(static_cast<const float *>(pr.loc) + u * 8, act, xy);
    load_q_f32<4>(static_cast<const float *>(pr.attn) + u * 4, act, a);
    sc.mulx = sc.muly = sc.divx = sc.divy = 1.f;
  } else {
    float off[8], lg[4];
    if constexpr (PRE == 1) {
      load_q_f32<8>(static_cast<const float *>(pr.loc) + u * 8, act, off);
      load_q_f32<4>(static_cast<const float *>(pr.attn) + u * 4, act, lg);
    } else if constexpr (PRE == 2) {
      load_q_bf16<8>(static_cast<const __nv_bfloat16 *>(pr.loc) + u * 8, act, off);
      load_q_bf16<4>(static_cast<const __nv_bfloat16 *>(pr.attn) + u * 4, act, lg);
    } else {
      load_q_f16<8>(static_cast<const __half *>(pr.loc) + u * 8, act, off);
      load_q_f16<4>(static_cast<const __half *>(pr.attn) + u * 4, act, lg);
    }
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
    float m = act ? fmaxf(fmaxf(lg[0], lg[1]), fmaxf(lg[2], lg[3])) : -INFINITY;
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    float e[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) e[p] = act ? expf(lg[p] - m) : 0.f;
    float s = (e[0] + e[1]) + (e[2] + e[3]);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
#pragma unroll
    for (int p = 0; p < 4; ++p) a[p] = act ? __fdiv_rn(e[p], s) : 0.f;
    if (pr.ref_dim == 2) {
      const float fw = (float)W, fh = (float)H;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        xy[2 * p] = __fadd_rn(rx, round_q<PRE>(__fdiv_rn(off[2 * p], fw)));
        xy[2 * p + 1] = __fadd_rn(ry, round_q<PRE>(__fdiv_rn(off[2 * p + 1], fh)));
      }
      sc.mulx = sc.muly = 1.f; sc.divx = fw; sc.divy = fh;
    } else {
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float qx = round_q<PRE>(__fdiv_rn(off[2 * p], 4.f)), qy = round_q<PRE>(__fdiv_rn(off[2 * p + 1], 4.f));
        xy[2 * p] = __fadd_rn(rx, __fmul_rn(__fmul_rn(qx, rw), 0.5f));
        xy[2 * p + 1] = __fadd_rn(ry, __fmul_rn(__fmul_rn(qy, rh), 0.5f));
      }
      sc.mulx = 0.5f * rw; sc.muly = 0.5f * rh; sc.divx = sc.divy = 4.f;
    }
  }
}

// grad_sampling_loc / grad_attn_weight of one (query, level), or with the fused pre-op the gradients of the raw offsets and
// logits (softmax backward across the query's four level threads: every thread of the warp must call this).
template <int PRE>
__device__ __forceinline__ void store_small_grads(const Problem &pr, const Item &it, int lc, bool act, const float (&a)[4],
                                                  const OffsetScale &sc, const float (&gx)[4], const float (&gy)[4],
                                                  const float (&gw)[4]) {
  const long long u = (long long)it.item * pr.L + lc;
  if constexpr (PRE == 0) {
    if (act) {
      float4 *gl = reinterpret_cast<float4 *>(static_cast<float *>(pr.grad_loc) + u * 8);
      gl[0] = make_float4(gx[0], gy[0], gx[1], gy[1]);
      gl[1] = make_float4(gx[2], gy[2], gx[3], gy[3]);
      *reinterpret_cast<float4 *>(static_cast<float *>(pr.grad_attn) + u * 4) = make_float4(gw[0], gw[1], gw[2], gw[3]);
    }
  } else {
    // softmax backward: S = sum over the query's L*P samples of a * d(loss)/da
    float part = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) part += act ? a[p] * gw[p] : 0.f;
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    float ox[4], oy[4], og[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      ox[p] = __fdiv_rn(gx[p] * sc.mulx, sc.divx);
      oy[p] = __fdiv_rn(gy[p] * sc.muly, sc.divy);
      og[p] = a[p] * (gw[p] - part);
    }
    if (act) {
      if constexpr (PRE == 1) {
        float4 *gl = reinterpret_cast<float4 *>(static_cast<float *>(pr.grad_loc) + u * 8);
        gl[0] = make_float4(ox[0], oy[0], ox[1], oy[1]);
        gl[1] = make_float4(ox[2], oy[2], ox[3], oy[3]);
        *reinterpret_cast<float4 *>(static_cast<float *>(pr.grad_attn) + u * 4) = make_float4(og[0], og[1], og[2], og[3]);
      } else if constexpr (PRE == 2) {
        uint4 q;
        q.x = Chan<__nv_bfloat16>::pack2(ox[0], oy[0]); q.y = Chan<__nv_bfloat16>::pack2(ox[1], oy[1]);
        q.z = Chan<__nv_bfloat16>::pack2(ox[2], oy[2]); q.w = Chan<__nv_bfloat16>::pack2(ox[3], oy[3]);
        *reinterpret_cast<uint4 *>(static_cast<__nv_bfloat16 *>(pr.grad_loc) + u * 8) = q;
        *reinterpret_cast<uint2 *>(static_cast<__nv_bfloat16 *>(pr.grad_attn) + u * 4) =
            make_uint2(Chan<__nv_bfloat16>::pack2(og[0], og[1]), Chan<__nv_bfloat16>::pack2(og[2], og[3]));
      } else {
        uint4 q;
        q.x = Chan<__half>::pack2(ox[0], oy[0]); q.y = Chan<__half>::pack2(ox[1], oy[1]);
        q.z = Chan<__half>::pack2(ox[2], oy[2]); q.w = Chan<__half>::pack2(ox[3], oy[3]);
        *reinterpret_cast<uint4 *>(static_cast<__half *>(pr.grad_loc) + u * 8) = q;
        *reinterpret_cast<uint2 *>(static_cast<__half *>(pr.grad_attn) + u * 4) =
            make_uint2(Chan<__half>::pack2(og[0], og[1]), Chan<__half>::pack2(og[2], og[3]));
      }
    }
  }
}

// Geometry of one sample as phase A needs it.
struct Sample {
  float lh, lw, aw;   // fractions and attention weight, zeroed when the sample is rejected (cuh:288)
  int mask;           // which of the 4 neighbours exist (cuh:56,62,68,74), 0 when rejected
  int r00;            // value row (b*S + level start + h_low*W + w_low) of the (low, low) neighbour
  int h_low, w_low;
};

__device__ __forceinline__ Sample make_sample(float x, float y, float a, bool act, int H, int W, int rowbase) {
  const Cell<float> c = locate(x, y, H, W);
  const bool in = act && c.inside;
  Sample s;
  s.mask = in ? corner_mask(c.h_low, c.w_low, H, W) : 0;
  s.lh = in ? c.lh : 0.f;
  s.lw = in ? c.lw : 0.f;
  s.aw = in ? a : 0.f;
  s.h_low = in ? c.h_low : 0;
  s.w_low = in ? c.w_low : 0;
  s.r00 = rowbase + s.h_low * W + s.w_low;
  return s;
}

// Occupancy and gather batching of the tile forward, from a same-box A/B of builds (cfg-2 bf16 / cfg-1 fp32, us per layer):
// 16-bit rows: 2 CTAs/SM x 8 gathers in flight 299.8, 3 x 4 271.1, 4 x 4 257.2, 4 x 2 257.0, 4 x 8 (spills) 271.5, 5 x 4 259.3,
// 6 x 2 262.2; software-pipelined batches no better.  fp32 rows (8 registers per gather): 3 x 2 103.4, 3 x 4 100.7, 4 x 2 106.6.
template <typename VT>
struct TileFwdCfg {
  static constexpr bool wide = sizeof(typename Chan<VT>::Raw) != 16;
#if defined(MSDA_TILE_MINB16) && defined(MSDA_TILE_RB)     // A/B builds only (MSDA_NVCC_EXTRA)
  static constexpr int minb = wide ? 3 : MSDA_TILE_MINB16;
  static constexpr int rb = wide ? 4 : MSDA_TILE_RB;
#else
  static constexpr int minb = wide ? 3 : 4;      // resident CTAs per SM the kernel is register-limited for
  static constexpr int rb = 4;                   // row gathers kept in flight per lane
#endif
};
template <typename VT, int MINB>
struct RowBatch {   // backward: row gathers kept in flight per lane in the direct path (registers: 4 or 8 per gather)
  static constexpr int value = (sizeof(typename Chan<VT>::Raw) == 16 ? 8 : 4) / (MINB >= 3 ? 2 : 1);
};

// ================================================================================================ forward
template <typename VT, int PRE, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) fwd_tile_kernel(const Problem pr) {
  pdl_wait();
  constexpr int D = 32, CH = 8, NP = 4;
  static_assert(Chan<VT>::kPerLane == CH && Chan<VT>::kPairs == NP, "8 channels per lane");
  extern __shared__ __align__(16) unsigned char smem[];
  Tables &tb = *reinterpret_cast<Tables *>(smem);
  int *itab = reinterpret_cast<int *>(smem + kTablesBytes + (size_t)kRecs * sizeof(int2));   // [64] output row of every query, -1 = none
  int2 *rb = reinterpret_cast<int2 *>(smem + kTablesBytes);        // [kRecs] {offset | flags, weight}; single copy: with
  // two copies (one barrier per unit) three resident CTAs leave the SM 28 KB of L1 for the gathers
  load_tables(pr, tb, kTileItems);

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const VT *__restrict__ value = static_cast<const VT *>(pr.value);
  VT *__restrict__ out = static_cast<VT *>(pr.out);

  for (int unit = blockIdx.x; unit < tb.n_units; unit += gridDim.x) {
    const UnitCtx uc = decode_unit(pr, tb, unit);
    {  // ---- phase A: thread = (query k, level l)
      const int k = t >> 2, l = t & 3;
      const Item it = decode_item(pr, tb, uc, k);
      const bool act = it.valid && l < pr.L;
      const int lc = act ? l : 0;
      const int H = tb.H[lc], W = tb.W[lc];
      float xy[8], a[4];
      OffsetScale sc;
      thread_level_inputs<PRE>(pr, it, lc, act, H, W, xy, a, sc);
      const int rowbase = it.b * pr.S + tb.start[lc];
      int2 *my = rb + t * kRecStride;
      if (l == 0) itab[k] = it.valid ? it.item : -1;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const Sample s = make_sample(xy[2 * p], xy[2 * p + 1], a[p], act, H, W, rowbase);
        const float hh = 1.f - s.lh, hw = 1.f - s.lw;
        const float ah = hh * s.aw, al = s.lh * s.aw;
        const float cw[4] = {ah * hw, ah * s.lw, al * hw, al * s.lw};
        int mask = s.mask;
        if (pr.pad_mask != nullptr && mask != 0) mask = drop_padded(pr.pad_mask, mask, s.r00, W);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int r = s.r00 + (kk & 1) + ((kk & 2) ? W : 0);
          const bool on = (mask >> kk) & 1;
          const float w = on ? cw[kk] : 0.f;
          // a zero weight needs no row: 0 * row is 0 for every finite row
          const int o = (on && w != 0.f) ? (r * pr.M + it.m) * D : -1;           // negative: no row behind this record
          my[p * 4 + kk] = make_int2(o, __float_as_int(w));
        }
      }
    }
    __syncwarp();        // a warp's row loop reads only what its own 32 (query, level) threads wrote: no block barrier
    {  // ---- row loop: two queries per warp (half-warps), group = neighbour, lane = 8 channels
      const int half = lane >> 4, kk = (lane >> 2) & 3, j = lane & 3;
      const VT *vbase = as_given(value + j * CH);   // address = one IMAD.WIDE (msda_common.cuh)
      constexpr int RB = TileFwdCfg<VT>::rb;
#pragma unroll 1
      for (int pair = 0; pair < 4; ++pair) {
        const int k = warp * 8 + pair * 2 + half;
        const int out_item = itab[k];
        const int2 *rp = rb + (k * 4) * kRecStride + kk;
        unsigned long long acc[NP] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
        for (int s0 = 0; s0 < 16; s0 += RB) {
          typename Chan<VT>::Raw raw[RB];
          float w[RB];
#pragma unroll
          for (int u = 0; u < RB; ++u) {
            const int s = s0 + u;
            const int2 r = rp[(s >> 2) * kRecStride + (s & 3) * 4];
            w[u] = __int_as_float(r.y);
            raw[u] = Chan<VT>::gather(vbase + r.x, r.x >= 0);
          }
#pragma unroll
          for (int u = 0; u < RB; ++u) {
            unsigned long long v[NP];
            Chan<VT>::unpack2(raw[u], v);
            const unsigned long long w2 = pack_f2(w[u], w[u]);
#pragma unroll
            for (int c = 0; c < NP; ++c) acc[c] = ffma2(w2, v[c], acc[c]);
          }
        }
        float res[CH];
#pragma unroll
        for (int c = 0; c < NP; ++c) unpack_f2(acc[c], res[2 * c], res[2 * c + 1]);
        {   // reduce-scatter over the 4 neighbour groups (6 shuffles instead of 16): every group ends with 2 of the lane's 8
            // channels and stores them (same-box A/B with the query table above: 257.4 -> 255.5 us, U 277.5 -> 272.3 us)
          const bool up1 = (kk & 1) != 0;
          float h[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float send = up1 ? res[c] : res[c + 4];
            const float keep = up1 ? res[c + 4] : res[c];
            h[c] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
          const bool up2 = (kk & 2) != 0;
          float q[2];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const float send = up2 ? h[c] : h[c + 2];
            const float keep = up2 ? h[c + 2] : h[c];
            q[c] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
          // channels held: base = (up1 ? 4 : 0) + (up2 ? 2 : 0)
          const int cbase = (up1 ? 4 : 0) + (up2 ? 2 : 0);
          if (out_item >= 0) {
            if constexpr (sizeof(VT) == 2 && !TileFwdCfg<VT>::wide) {
              unsigned word;
              if constexpr (__is_same(VT, __half)) word = Chan<__half>::pack2(q[0], q[1]);
              else word = Chan<__nv_bfloat16>::pack2(q[0], q[1]);
              *reinterpret_cast<unsigned *>(out + (long long)out_item * D + j * CH + cbase) = word;
            } else {
              *reinterpret_cast<float2 *>(reinterpret_cast<float *>(out) + (long long)out_item * D + j * CH + cbase) = make_float2(q[0], q[1]);
            }
          }
        }
      }
    }
    __syncwarp();        // the next unit's phase A overwrites the warp's records
  }
}

// ================================================================================================ backward
// Aggregation window of one level for the current tile (per thread: the level it owns in phase A).
struct Window {
  int y0, x0;      // first pixel row / column
  int h;           // rows (0 = no window: every sample is handled directly)
  int lgw;         // log2 of the row pitch (columns rounded up to a power of two)
  int base;        // first head word
};

__device__ __forceinline__ int ceil_log2(int v) {
  return v <= 1 ? 0 : 32 - __clz(v - 1);
}

// Window of level `l` for the tile `uc`: the tile's footprint on that level (query i sits at the centre of pixel i of its
// own level) plus `margin` pixels on every side, clipped to the image and to `cap` head words.
__device__ __forceinline__ Window make_window(const UnitCtx &uc, const Tables &tb, int l, bool tiled, int margin, int cap) {
  Window w;
  w.y0 = w.x0 = w.h = w.lgw = w.base = 0;
  if (!tiled) return w;
  const int H = tb.H[l], W = tb.W[l];
  const float sy = (float)H / (float)uc.Hq, sx = (float)W / (float)uc.Wq;
  const int th = kUnitItems >> 3;       // tile is 8 x 8 (tile_w_log2 = 3)
  int ya = (int)floorf(((float)uc.y0 + 0.5f) * sy - 0.5f) - margin;
  int yb = (int)floorf(((float)(uc.y0 + th - 1) + 0.5f) * sy - 0.5f) + 1 + margin;
  int xa = (int)floorf(((float)uc.x0 + 0.5f) * sx - 0.5f) - margin;
  int xb = (int)floorf(((float)(uc.x0 + 7) + 0.5f) * sx - 0.5f) + 1 + margin;
  ya = max(ya, 0); xa = max(xa, 0); yb = min(yb, H - 1); xb = min(xb, W - 1);
  if (yb < ya || xb < xa) return w;
  w.lgw = ceil_log2(xb - xa + 1);
  w.y0 = ya; w.x0 = xa;
  w.h = min(yb - ya + 1, cap >> w.lgw);
  return w;
}

// Shared-memory plan of the backward (bytes).  tmp is a union over the life of a record: {window row, slot} while the
// sort runs, the row offset for directly handled samples, and finally the record's dot product <grad_out, row>.
__device__ __forceinline__ int head_cap(int l) { return l == 0 ? 1024 : (l == 1 ? 768 : 384); }   // head words per level
constexpr int kHeads = 2560;
constexpr int kMaxEntries = kThreads * 16;               // 4096 (sample, neighbour) records per unit
constexpr int kMaxDirect = kThreads * 4;                 // 1024 samples
constexpr size_t kOffCoef = kTablesBytes;
constexpr size_t kOffTmp = kOffCoef + (size_t)kRecs * 4;
constexpr size_t kOffSent = kOffTmp + (size_t)kRecs * 4;
constexpr size_t kOffCnt = kOffSent + (size_t)kMaxEntries * 8;
constexpr size_t kOffStart = kOffCnt + (size_t)kHeads * 4;
constexpr size_t kOffRowoff = kOffStart + (size_t)(kHeads + 16) * 4;
constexpr size_t kOffGo = kOffRowoff + (size_t)kHeads * 4;
constexpr size_t kOffDlist = kOffGo + (size_t)kTileItems * 32 * 4;
constexpr size_t kOffMisc = kOffDlist + (size_t)kMaxDirect * 2;
constexpr size_t kBwdSmem = kOffMisc + 64 * 4;
static_assert(kOffSent % 16 == 0 && kOffGo % 16 == 0 && kOffCnt % 16 == 0, "vector accesses need 16-byte alignment");

// grad_value rows as 4 lanes x 8 channels: whole-row reductions by pairing two row groups that share grad_out (the two
// x-neighbours of a sample): the pair sends the first group's row, then the second's; lanes of the first group carry
// channels [8j, 8j+4), lanes of the second [8j+4, 8j+8) -- every instruction covers whole 128-byte rows.
__device__ __forceinline__ void coop_row_reduce(float *gbase, bool first, float my_c, int my_o, const unsigned long long (&gp)[4]) {
  const float o_c = __shfl_xor_sync(0xffffffffu, my_c, 4);
  const int o_o = __shfl_xor_sync(0xffffffffu, my_o, 4);
  const float c_a = first ? my_c : o_c, c_b = first ? o_c : my_c;
  const int off_a = first ? my_o : o_o, off_b = first ? o_o : my_o;
  const unsigned long long g0 = first ? gp[0] : gp[2], g1 = first ? gp[1] : gp[3];
  float *dst = gbase + (first ? 0 : 4);
  if (c_a != 0.f) {
    float r0, r1, r2, r3;
    const unsigned long long c2 = pack_f2(c_a, c_a);
    unpack_f2(fmul2(c2, g0), r0, r1);
    unpack_f2(fmul2(c2, g1), r2, r3);
    red_add_f32x4(dst + off_a, r0, r1, r2, r3);
  }
  if (c_b != 0.f) {
    float r0, r1, r2, r3;
    const unsigned long long c2 = pack_f2(c_b, c_b);
    unpack_f2(fmul2(c2, g0), r0, r1);
    unpack_f2(fmul2(c2, g1), r2, r3);
    red_add_f32x4(dst + off_b, r0, r1, r2, r3);
  }
}

// Backward of one tile in six steps (barriers between them):
//   A   geometry per (query, level) thread.  A sample whose existing neighbours all fall into the level's window becomes
//       4 ENTRIES: each takes a slot in its window row's counter (ATOMS.ADD) -- the count pass of a counting sort by row.
//       Other samples (uniform random locations, large offsets) are queued for step D.
//   S   exclusive scan of the row counters (all 256 threads), then every thread scatters its entries to their sorted
//       positions: {coefficient, row | query | record}.
//   2   row-stationary pass over the sorted entries, 64 lane groups x an equal share of entries each: a group (4 lanes x
//       8 channels) loads a value row ONCE, and for each entry of that row adds coefficient x grad_out[query] to a
//       register accumulator and forms <grad_out[query], row>; the row's accumulator leaves as one whole-row reduction.
//       Value rows and grad_value rows are touched once per tile and row instead of once per sample.
//   D   queued samples: gather, dot products and direct reductions, two samples x four neighbours per warp step.
//   C   grad_sampling_loc / grad_attn_weight per (query, level) thread from the 16 dot products, 16-byte stores.
template <typename VT, int PRE>
__global__ void __launch_bounds__(kThreads, 2) bwd_tile_kernel(const Problem pr) {
  pdl_wait();
  constexpr int D = 32, CH = 8, NP = 4;
  static_assert(Chan<VT>::kPerLane == CH && Chan<VT>::kPairs == NP, "8 channels per lane");
  extern __shared__ __align__(16) unsigned char smem[];
  Tables &tb = *reinterpret_cast<Tables *>(smem);
  float *coef = reinterpret_cast<float *>(smem + kOffCoef);
  unsigned *tmp = reinterpret_cast<unsigned *>(smem + kOffTmp);
  int2 *sent = reinterpret_cast<int2 *>(smem + kOffSent);
  int *cnt = reinterpret_cast<int *>(smem + kOffCnt);
  int *start = reinterpret_cast<int *>(smem + kOffStart);
  unsigned *rowoff = reinterpret_cast<unsigned *>(smem + kOffRowoff);
  float *go32 = reinterpret_cast<float *>(smem + kOffGo);
  unsigned short *dlist = reinterpret_cast<unsigned short *>(smem + kOffDlist);
  int *misc = reinterpret_cast<int *>(smem + kOffMisc);      // [0..7] warp totals of the scan, [8], [9] direct-sample counters
  load_tables(pr, tb, kTileItems);

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int i = t; i < kHeads; i += kThreads) cnt[i] = 0;
  if (t < 2) misc[8 + t] = 0;
  const bool tiled = tb.n_tiles > 0;
  const VT *__restrict__ value = static_cast<const VT *>(pr.value);
  const VT *__restrict__ grad_out = static_cast<const VT *>(pr.grad_out);
  float *__restrict__ gacc = static_cast<float *>(pr.grad_value_acc);
  __syncthreads();

  int parity = 0;
  for (int unit = blockIdx.x; unit < tb.n_units; unit += gridDim.x, parity ^= 1) {
    const UnitCtx uc = decode_unit(pr, tb, unit);
    // ---- A: thread = (query k, level l)
    const int k = t >> 2, l = t & 3;
    const Item it = decode_item(pr, tb, uc, k);
    const bool act = it.valid && l < pr.L;
    const int lc = act ? l : 0;
    const int H = tb.H[lc], W = tb.W[lc];
    float xy[8], a[4];
    OffsetScale sc;
    thread_level_inputs<PRE>(pr, it, lc, act, H, W, xy, a, sc);
    {  // this query's grad_out row -> fp32 in shared memory; the query's four threads take 8 channels each
      unsigned long long g[NP];
      Chan<VT>::unpack2(Chan<VT>::gather(grad_out + (long long)it.item * D + l * CH, it.valid), g);
      float4 lo, hi;
      unpack_f2(g[0], lo.x, lo.y); unpack_f2(g[1], lo.z, lo.w);
      unpack_f2(g[2], hi.x, hi.y); unpack_f2(g[3], hi.z, hi.w);
      *reinterpret_cast<float4 *>(go32 + k * D + l * CH) = lo;
      *reinterpret_cast<float4 *>(go32 + k * D + l * CH + 4) = hi;
    }
    // window of my level (a function of the unit only, so every quad of the CTA derives the same layout)
    Window win = make_window(uc, tb, min(l, pr.L - 1), tiled && l < pr.L, pr.win_margin, head_cap(l));
    {
      const int sz = win.h << win.lgw;
      const int q = lane & ~3;
      const int s0 = __shfl_sync(0xffffffffu, sz, q), s1 = __shfl_sync(0xffffffffu, sz, q + 1),
                s2 = __shfl_sync(0xffffffffu, sz, q + 2);
      win.base = (l > 0 ? s0 : 0) + (l > 1 ? s1 : 0) + (l > 2 ? s2 : 0);
    }
    const int rowbase = it.b * pr.S + tb.start[lc];
    float g_lh[4], g_lw[4], g_aw[4];
    const int rec0 = t * kRecStride;
    unsigned winmask = 0u;                       // which of my 16 records are window entries
    if (t == 0) misc[8 + (parity ^ 1)] = 0;      // the next unit's direct-sample counter (last read before the previous barrier)
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const Sample s = make_sample(xy[2 * p], xy[2 * p + 1], a[p], act, H, W, rowbase);
      g_lh[p] = s.lh; g_lw[p] = s.lw; g_aw[p] = s.aw;
      const float hh = 1.f - s.lh, hw = 1.f - s.lw;
      const float ah = hh * s.aw, al = s.lh * s.aw;
      const float cw[4] = {ah * hw, ah * s.lw, al * hw, al * s.lw};
      int mask = s.mask;
      if (pr.pad_mask != nullptr && mask != 0) mask = drop_padded(pr.pad_mask, mask, s.r00, W);
      // the cell's existing rows / columns must lie inside the window
      const int wr = s.h_low - win.y0, wc = s.w_low - win.x0;
      const int r_lo = wr + ((mask & 3) ? 0 : 1), r_hi = wr + ((mask & 12) ? 1 : 0);
      const int c_lo = wc + ((mask & 5) ? 0 : 1), c_hi = wc + ((mask & 10) ? 1 : 0);
      const bool inwin = mask != 0 && r_lo >= 0 && r_hi < win.h && c_lo >= 0 && c_hi < (1 << win.lgw);
      const int hidx00 = win.base + (wr << win.lgw) + wc;
      const bool direct = mask != 0 && !inwin;
      {  // queue directly handled samples: one counter update per warp and point
        const unsigned dm = __ballot_sync(0xffffffffu, direct);
        if (dm != 0u) {
          int basepos = 0;
          if (lane == 0) basepos = atomicAdd(misc + 8 + parity, __popc(dm));
          basepos = __shfl_sync(0xffffffffu, basepos, 0);
          if (direct) dlist[basepos + __popc(dm & ((1u << lane) - 1u))] = (unsigned short)(t * 4 + p);
        }
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int rec = rec0 + p * 4 + kk;
        const bool on = (mask >> kk) & 1;
        const unsigned off = (unsigned)(((s.r00 + (kk & 1) + ((kk & 2) ? W : 0)) * pr.M + it.m) * D);
        coef[rec] = on ? cw[kk] : 0.f;
        if (inwin && on) {
          const int hidx = hidx00 + (kk & 1) + ((kk & 2) ? (1 << win.lgw) : 0);
          const int slot = atomicAdd(cnt + hidx, 1);
          if (slot == 0) rowoff[hidx] = off;
          tmp[rec] = ((unsigned)hidx << 16) | (unsigned)slot;
          winmask |= 1u << (p * 4 + kk);
        } else {
          tmp[rec] = direct ? (on ? off : kNoLoad) : 0u;       // 0u doubles as the dot product 0.0f of a missing row
        }
      }
    }
    __syncthreads();

    // ---- S: exclusive scan of the row counters -> start[], counters back to zero, entries to their sorted places
    int n_entries = 0;
    if (tiled) {
      constexpr int PER = kHeads / kThreads;                 // 10 consecutive counters per thread
      int c[PER];
#pragma unroll
      for (int i = 0; i < PER; i += 2) {
        const int2 v = *reinterpret_cast<const int2 *>(cnt + t * PER + i);
        c[i] = v.x; c[i + 1] = v.y;
        *reinterpret_cast<int2 *>(cnt + t * PER + i) = make_int2(0, 0);
      }
      int sum = 0;
#pragma unroll
      for (int i = 0; i < PER; ++i) { const int v = c[i]; c[i] = sum; sum += v; }
      int incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      if (lane == 31) misc[warp] = incl;
      __syncthreads();
      int woff = 0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) {
        const int v = misc[w];
        if (w < warp) woff += v;
        n_entries += v;
      }
      const int excl = woff + incl - sum;
#pragma unroll
      for (int i = 0; i < PER; i += 2)
        *reinterpret_cast<int2 *>(start + t * PER + i) = make_int2(excl + c[i], excl + c[i + 1]);
      if (t == kThreads - 1) start[kHeads] = excl + sum;        // == n_entries: end of the last row's entries
      __syncthreads();
      for (unsigned m = winmask; m != 0u; m &= m - 1u) {
        const int ci = __ffs(m) - 1;
        const int rec = rec0 + ci;
        const unsigned v = tmp[rec];
        const int hidx = (int)(v >> 16);
        const int pos = start[hidx] + (int)(v & 0xffffu);
        sent[pos] = make_int2(__float_as_int(coef[rec]), (hidx << 19) | (k << 13) | rec);
      }
    }
    __syncthreads();

    const int g = lane >> 2, j = lane & 3;
    if (n_entries > 0) {  // ---- 2: row-stationary pass; group = 4 lanes x 8 channels, 64 groups share the entries evenly
      const int gid = warp * 8 + g;
      const int share = (n_entries + 63) >> 6;
      int pos = gid * share;
      const int end = min(n_entries, pos + share);
      int cur = -1;
      unsigned cur_off = 0u, nxt_off = 0u;
      unsigned long long vr[NP] = {0ull, 0ull, 0ull, 0ull}, acc[NP] = {0ull, 0ull, 0ull, 0ull};
      // The value row of the NEXT row segment is requested when the current one starts, so its L2 latency is covered by
      // the current segment's entries (a row load exposed at every segment change, with the 8 groups of a warp in
      // lockstep, was the largest stall of this step: ncu, 16 % of all samples; 1.11 -> 0.93 ms with the 4 entries of a
      // batch also fetched by independent loads).  Starting the segments of the 8 groups together in rounds instead
      // (one straight-line block per round) was measured and lost: 1.33 ms -- the segments of a warp's groups differ
      // too much in length.
      typename Chan<VT>::Raw raw_next{};
      if (pos < end) {
        nxt_off = rowoff[sent[pos].y >> 19];
        raw_next = Chan<VT>::gather_plain(value + nxt_off + j * CH);
      }
      auto flush = [&]() {
        float r[CH];
#pragma unroll
        for (int c = 0; c < NP; ++c) unpack_f2(acc[c], r[2 * c], r[2 * c + 1]);
        float *dst = gacc + cur_off + j * CH;
        red_add_f32x4(dst, r[0], r[1], r[2], r[3]);
        red_add_f32x4(dst + 4, r[4], r[5], r[6], r[7]);
      };
      while (__any_sync(0xffffffffu, pos < end)) {
        const bool live = pos < end;
        // the next 4 entries of my share, fetched together (independent loads; indices clamped into the share)
        int2 e[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) e[i] = sent[live ? min(pos + i, end - 1) : 0];
        const int h = live ? (e[0].y >> 19) : -1;
        if (live && h != cur) {                 // next row of this group
          if (cur >= 0) flush();
          cur = h;
          cur_off = nxt_off;
          Chan<VT>::unpack2(raw_next, vr);
#pragma unroll
          for (int c = 0; c < NP; ++c) acc[c] = 0ull;
          const int seg_end = start[h + 1];
          if (seg_end < end) {
            nxt_off = rowoff[sent[seg_end].y >> 19];
            raw_next = Chan<VT>::gather_plain(value + nxt_off + j * CH);
          }
        }
        // how many of them belong to the current row (entries are sorted by row)
        const bool ok1 = live && pos + 1 < end && (e[1].y >> 19) == h;
        const bool ok2 = ok1 && pos + 2 < end && (e[2].y >> 19) == h;
        const bool ok3 = ok2 && pos + 3 < end && (e[3].y >> 19) == h;
        const int n = live ? (ok3 ? 4 : (ok2 ? 3 : (ok1 ? 2 : 1))) : 0;
        float d[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool ok = i < n;
          const float cf = ok ? __int_as_float(e[i].x) : 0.f;
          const int kq = ok ? ((e[i].y >> 13) & 63) : 0;
          const float4 lo = *reinterpret_cast<const float4 *>(go32 + kq * D + j * CH);
          const float4 hi = *reinterpret_cast<const float4 *>(go32 + kq * D + j * CH + 4);
          const unsigned long long g0 = pack_f2(lo.x, lo.y), g1 = pack_f2(lo.z, lo.w), g2 = pack_f2(hi.x, hi.y),
                                   g3 = pack_f2(hi.z, hi.w);
          const unsigned long long c2 = pack_f2(cf, cf);
          acc[0] = ffma2(c2, g0, acc[0]); acc[1] = ffma2(c2, g1, acc[1]);
          acc[2] = ffma2(c2, g2, acc[2]); acc[3] = ffma2(c2, g3, acc[3]);
          unsigned long long d2 = ffma2(vr[0], g0, 0ull);
          d2 = ffma2(vr[1], g1, d2); d2 = ffma2(vr[2], g2, d2); d2 = ffma2(vr[3], g3, d2);
          float dlo, dhi;
          unpack_f2(d2, dlo, dhi);
          d[i] = dlo + dhi;
        }
        // the 4 lanes' partial dot products of 4 entries -> lane j holds entry j's sum (reduce-scatter)
        {
          const bool up = (j & 2) != 0;
          const float s0 = up ? d[0] : d[2], s1 = up ? d[1] : d[3];
          const float k0 = up ? d[2] : d[0], k1 = up ? d[3] : d[1];
          const float t0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 2);     // entries (0|2)
          const float t1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 2);     // entries (1|3)
          const bool odd = (j & 1) != 0;
          const float snd = odd ? t0 : t1, kp = odd ? t1 : t0;
          const float mine = kp + __shfl_xor_sync(0xffffffffu, snd, 1);  // entry j
          const int my = j == 0 ? e[0].y : (j == 1 ? e[1].y : (j == 2 ? e[2].y : e[3].y));
          if (j < n) tmp[my & 8191] = __float_as_uint(mine);
        }
        pos += n;
      }
      if (cur >= 0) flush();
    }

    {  // ---- D: directly handled samples; warp step = 2 samples x 4 neighbours, lane = 8 channels of one row
      const int n_direct = misc[8 + parity];
      const bool first = (g & 1) == 0;
      const VT *vbase = as_given(value + j * CH);   // address = one IMAD.WIDE (msda_common.cuh)
      float *gbase = gacc + j * CH;
      for (int i0 = warp * 2; i0 < n_direct; i0 += (kThreads / 32) * 2) {
        const int si = i0 + (g >> 2);
        const bool have = si < n_direct;
        const int sid = have ? (int)dlist[si] : 0;               // (thread << 2) | point
        const int rec = (sid >> 2) * kRecStride + (sid & 3) * 4 + (g & 3);
        const int kq = sid >> 4;
        const unsigned o = have ? tmp[rec] : kNoLoad;
        const bool ld = (o & kNoLoad) == 0u;
        const typename Chan<VT>::Raw raw = Chan<VT>::gather(vbase + (o & kOffMask), ld);
        unsigned long long gp[NP];
        {
          const float4 lo = *reinterpret_cast<const float4 *>(go32 + kq * D + j * CH);
          const float4 hi = *reinterpret_cast<const float4 *>(go32 + kq * D + j * CH + 4);
          gp[0] = pack_f2(lo.x, lo.y); gp[1] = pack_f2(lo.z, lo.w);
          gp[2] = pack_f2(hi.x, hi.y); gp[3] = pack_f2(hi.z, hi.w);
        }
        unsigned long long v[NP];
        Chan<VT>::unpack2(raw, v);
        unsigned long long d2 = 0ull;
#pragma unroll
        for (int c = 0; c < NP; ++c) d2 = ffma2(gp[c], v[c], d2);
        float dlo, dhi;
        unpack_f2(d2, dlo, dhi);
        float d = dlo + dhi;
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        const float my_c = ld ? coef[rec] : 0.f;
        __syncwarp();                                            // every lane has read tmp[rec] before it becomes the dot
        if (have && j == 0) tmp[rec] = __float_as_uint(d);
        coop_row_reduce(gbase, first, my_c, (int)(o & kOffMask), gp);
      }
    }
    __syncthreads();

    {  // ---- C: the two small gradients of my (query, level)
      const float *dq = reinterpret_cast<const float *>(tmp) + rec0;
      float gx[4], gy[4], gw[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float d1 = dq[p * 4], d2 = dq[p * 4 + 1], d3 = dq[p * 4 + 2], d4 = dq[p * 4 + 3];
        const float lh = g_lh[p], lw = g_lw[p], aw = g_aw[p];
        const float hh = 1.f - lh, hw = 1.f - lw;
        gx[p] = (float)W * aw * (hh * (d2 - d1) + lh * (d4 - d3));
        gy[p] = (float)H * aw * (hw * (d3 - d1) + lw * (d4 - d2));
        gw[p] = hh * hw * d1 + hh * lw * d2 + lh * hw * d3 + lh * lw * d4;
      }
      store_small_grads<PRE>(pr, it, lc, act, a, sc, gx, gy, gw);
    }
    // The next unit's step A may start at once: everything it overwrites (coef, tmp, row counters, go32, the other
    // direct-sample queue) was last read before the barrier above, except tmp, which this thread has just consumed
    // itself -- but other threads' step C still reads THEIR tmp words, which step A of a faster thread never touches.
  }
}

// ================================================================================================ backward, warp-scoped
#ifndef MSDA_WARP_BWD_QB
#define MSDA_WARP_BWD_QB 2
#endif
// The row-group backward (msda_bwd.cu) restructured like the tile forward, WITHOUT aggregation across queries: per (query,
// level) the scalar work -- cell, neighbour mask, bilinear coefficients, the sliding-window merge of neighbour rows shared
// by consecutive points -- is done once by one thread (phase A) instead of four times by the lanes of a row group, and the
// row loop reads lists from shared memory:
//   A  thread = (query, level): 16 gather offsets (4 points x 4 neighbours, -1 = no row), a COMPACTED list of the
//      (row offset, coefficient) reductions the window merge emits (at most 16), and the count.
//   1  lane group = (query, level), two queries per warp pass: 16 row gathers x <grad_out, row> -> reduce-scatter over the
//      4 lanes -> the 16 dot products overwrite the gather offsets; then the reductions: the two groups of a level pair
//      walk BOTH lists, the first sending channels [8j, 8j+4) and the second [8j+4, 8j+8) of every row, so each
//      red.global.add.v4.f32 instruction still covers whole 128-byte rows (as COOP in msda_bwd.cu, without the shuffles).
//   C  thread = (query, level): grad_sampling_loc / grad_attn_weight from the dot products, 16-byte stores.
// A warp reads only what its own 32 threads wrote: no block barrier inside the unit loop.
constexpr int kGlStride = 20;   // ints per (query, level): 16 offsets / dot products, the count, pad.  16-byte aligned and
                                // 20 * seg mod 32 is a permutation of the 4-word bank groups over 8 segments: LDS.128 conflict-free
constexpr size_t kWbOffGl = kTablesBytes;
constexpr size_t kWbOffRl = kWbOffGl + (size_t)kThreads * kGlStride * 4;
constexpr size_t kWbOffItab = kWbOffRl + (size_t)kRecs * sizeof(int2);
constexpr size_t kWbSmem = kWbOffItab + 256;

template <typename VT, int PRE, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) bwd_warp_kernel(const Problem pr) {
  pdl_wait();
  constexpr int D = 32, CH = 8, NP = 4;
  constexpr int QB = MSDA_WARP_BWD_QB;
  static_assert(Chan<VT>::kPerLane == CH && Chan<VT>::kPairs == NP, "8 channels per lane");
  extern __shared__ __align__(16) unsigned char smem[];
  Tables &tb = *reinterpret_cast<Tables *>(smem);
  int *GL = reinterpret_cast<int *>(smem + kWbOffGl);
  int2 *RL = reinterpret_cast<int2 *>(smem + kWbOffRl);
  int *itab = reinterpret_cast<int *>(smem + kWbOffItab);
  load_tables(pr, tb, kTileItems);

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const VT *__restrict__ value = static_cast<const VT *>(pr.value);
  const VT *__restrict__ grad_out = static_cast<const VT *>(pr.grad_out);
  float *__restrict__ gacc = static_cast<float *>(pr.grad_value_acc);
  const int MD = pr.M * D;

  for (int unit = blockIdx.x; unit < tb.n_units; unit += gridDim.x) {
    const UnitCtx uc = decode_unit(pr, tb, unit);
    // ---- phase A: thread = (query k, level l)
    const int k = t >> 2, l = t & 3;
    const Item it = decode_item(pr, tb, uc, k);
    const bool act = it.valid && l < pr.L;
    const int lc = act ? l : 0;
    const int H = tb.H[lc], W = tb.W[lc];
    float xy[8], a[4];
    OffsetScale sc;
    thread_level_inputs<PRE>(pr, it, lc, act, H, W, xy, a, sc);
    float g_lh[4], g_lw[4], g_aw[4];
    {
      const int rowbase = it.b * pr.S + tb.start[lc];
      const int WMD = W * MD;
      const int base_off = rowbase * MD + it.m * D;      // element offset of (first row of the level, head m)
      int *gl = GL + t * kGlStride;
      int2 *rl = RL + t * kRecStride;
      if (l == 0) itab[k] = it.valid ? it.item : -1;
      int nred = 0;
      // sliding 2x2 window over the points of this (query, level): see msda_bwd.cu (WIN)
      float win_c[4] = {0.f, 0.f, 0.f, 0.f};
      int win_h = 0, win_w = 0, win_off = 0;
      bool win_valid = false;
      auto emit = [&](const float (&e)[4], int off) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          if (e[kk] != 0.f) {
            rl[nred] = make_int2(off + ((kk & 1) ? MD : 0) + ((kk & 2) ? WMD : 0), __float_as_int(e[kk]));
            ++nred;
          }
      };
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const Cell<float> c = locate(xy[2 * p], xy[2 * p + 1], H, W);
        const bool in = act && c.inside;
        int mask = in ? corner_mask(c.h_low, c.w_low, H, W) : 0;
        if (pr.pad_mask != nullptr && mask != 0) mask = drop_padded(pr.pad_mask, mask, rowbase + c.h_low * W + c.w_low, W);
        g_lh[p] = in ? c.lh : 0.f;
        g_lw[p] = in ? c.lw : 0.f;
        g_aw[p] = in ? a[p] : 0.f;
        const float hh = 1.f - g_lh[p], hw = 1.f - g_lw[p];
        const float ah = hh * g_aw[p], al = g_lh[p] * g_aw[p];
        const float ncw[4] = {(mask & 1) ? ah * hw : 0.f, (mask & 2) ? ah * g_lw[p] : 0.f, (mask & 4) ? al * hw : 0.f,
                              (mask & 8) ? al * g_lw[p] : 0.f};
        const int o00 = in ? base_off + c.h_low * WMD + c.w_low * MD : 0;
        *reinterpret_cast<int4 *>(gl + 4 * p) = make_int4((mask & 1) ? o00 : -1, (mask & 2) ? o00 + MD : -1,
                                                          (mask & 4) ? o00 + WMD : -1, (mask & 8) ? o00 + WMD + MD : -1);
        const int dh = c.h_low - win_h, dw = c.w_low - win_w;
        const bool near = in && win_valid && (unsigned)(dh + 1) <= 2u && (unsigned)(dw + 1) <= 2u;
        float carry[4] = {0.f, 0.f, 0.f, 0.f};
        if (p > 0) {
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
          const bool r0 = near && dh <= 0, r1 = near && dh >= 0, c0 = dw <= 0, c1 = dw >= 0;
          const float e[4] = {(r0 && c0) ? 0.f : win_c[0], (r0 && c1) ? 0.f : win_c[1],
                              (r1 && c0) ? 0.f : win_c[2], (r1 && c1) ? 0.f : win_c[3]};
          emit(e, win_off);
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) win_c[kk] = ncw[kk] + carry[kk];
        win_h = c.h_low; win_w = c.w_low; win_valid = in; win_off = o00;
      }
      emit(win_c, win_off);
      gl[16] = nred;
    }
    __syncwarp();
    {  // ---- row loop: two queries per warp pass (half-warps), group = level, lane = 8 channels
      const int half = lane >> 4, g = (lane >> 2) & 3, j = lane & 3;
      const bool first = (g & 1) == 0;
      const VT *vbase = as_given(value + j * CH);
      float *gbase = as_given(gacc + j * CH + (first ? 0 : 4));
#pragma unroll 1
      for (int pair = 0; pair < 4; ++pair) {
        const int kq = warp * 8 + pair * 2 + half;
        const int seg = kq * 4 + g;
        const int item = itab[kq];
        unsigned long long go[NP];
        Chan<VT>::unpack2(Chan<VT>::gather(grad_out + (long long)(item < 0 ? 0 : item) * D + j * CH, item >= 0), go);
        int *glp = GL + seg * kGlStride;
        float d[16];
#pragma unroll
        for (int q0 = 0; q0 < 4; q0 += QB) {                 // QB points = 4 * QB row gathers in flight per lane
          typename Chan<VT>::Raw raw[4 * QB];
#pragma unroll
          for (int q = 0; q < QB; ++q) {
            const int4 o4 = *reinterpret_cast<const int4 *>(glp + 4 * (q0 + q));
            const int o[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) raw[4 * q + kk] = Chan<VT>::gather(vbase + o[kk], o[kk] >= 0);
          }
#pragma unroll
          for (int i = 0; i < 4 * QB; ++i) {
            unsigned long long v[NP];
            Chan<VT>::unpack2(raw[i], v);
            unsigned long long dot2 = 0ull;
#pragma unroll
            for (int c = 0; c < NP; ++c) dot2 = ffma2(go[c], v[c], dot2);
            float lo, hi;
            unpack_f2(dot2, lo, hi);
            d[4 * q0 + i] = lo + hi;
          }
        }
        reduce_scatter<16, 2>(d, j);                         // lane j: the four dot products of point j
        __syncwarp();                                        // every lane has read the offsets the dot products replace
        *reinterpret_cast<float4 *>(glp + 4 * j) = make_float4(d[0], d[1], d[2], d[3]);
        // reductions of this level pair: list of the even level first, then of the odd one
        const int segA = seg & ~1;
        const int nA = GL[segA * kGlStride + 16], nB = GL[(segA + 1) * kGlStride + 16];
        const unsigned long long gsel0 = first ? go[0] : go[2], gsel1 = first ? go[1] : go[3];
        auto send = [&](const int2 &r) {
          const float c = __int_as_float(r.y);
          const unsigned long long c2 = pack_f2(c, c);
          float r0, r1, r2, r3;
          unpack_f2(fmul2(c2, gsel0), r0, r1);
          unpack_f2(fmul2(c2, gsel1), r2, r3);
          red_add_f32x4(gbase + r.x, r0, r1, r2, r3);
        };
        auto walk = [&](const int2 *list, int n) {          // every entry of a list has a non-zero coefficient
          const int most = __reduce_max_sync(0xffffffffu, n);
#pragma unroll 1
          for (int e = 0; e < most; e += 2) {
            const bool p0 = e < n, p1 = e + 1 < n;
            int2 e0 = make_int2(0, 0), e1 = make_int2(0, 0);
            if (p0) e0 = list[e];
            if (p1) e1 = list[e + 1];
            if (p0) send(e0);
            if (p1) send(e1);
          }
        };
        walk(RL + segA * kRecStride, nA);
        walk(RL + (segA + 1) * kRecStride, nB);
      }
    }
    __syncwarp();
    {  // ---- phase C: the two small gradients of my (query, level)
      const float *dq = reinterpret_cast<const float *>(GL + t * kGlStride);
      float gx[4], gy[4], gw[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float4 dd = *reinterpret_cast<const float4 *>(dq + 4 * p);
        const float lh = g_lh[p], lw = g_lw[p], aw = g_aw[p];
        const float hh = 1.f - lh, hw = 1.f - lw;
        gx[p] = (float)W * aw * (hh * (dd.y - dd.x) + lh * (dd.w - dd.z));
        gy[p] = (float)H * aw * (hw * (dd.z - dd.x) + lw * (dd.w - dd.y));
        gw[p] = hh * hw * dd.x + hh * lw * dd.y + lh * hw * dd.z + lh * lw * dd.w;
      }
      store_small_grads<PRE>(pr, it, lc, act, a, sc, gx, gy, gw);
    }
    __syncwarp();        // the next unit's phase A overwrites the warp's lists
  }
}

constexpr size_t kFwdSmem = kTablesBytes + (size_t)kRecs * sizeof(int2) + 256;

template <typename K>
int set_smem(K kernel, size_t bytes) {
  return (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <typename VT>
int launch_fwd(const Problem &pr, int sms, int pre, cudaStream_t stream) {
  constexpr int MB = TileFwdCfg<VT>::minb;
#define MSDA_TILE_FWD(PREV)                                                              \
  if (pre == PREV) {                                                                     \
    if (int rc = set_smem(fwd_tile_kernel<VT, PREV, MB>, kFwdSmem)) return rc;           \
    return launch_pdl(fwd_tile_kernel<VT, PREV, MB>, sms * MB, kThreads, kFwdSmem, stream, pr);  \
  }
  MSDA_TILE_FWD(0) MSDA_TILE_FWD(1) MSDA_TILE_FWD(2) MSDA_TILE_FWD(3)
#undef MSDA_TILE_FWD
  return MSDA_E_UNSUPPORTED;
}

template <typename VT>
int launch_bwd(const Problem &pr, int grid, int pre, cudaStream_t stream) {
#define MSDA_TILE_BWD(PREV)                                                          \
  if (pre == PREV) {                                                                 \
    if (int rc = set_smem(bwd_tile_kernel<VT, PREV>, kBwdSmem)) return rc;           \
    launch_pdl(bwd_tile_kernel<VT, PREV>, grid, kThreads, kBwdSmem, stream, pr);             \
    return (int)cudaGetLastError();                                                  \
  }
  MSDA_TILE_BWD(0)       // the op boundary only: this kernel is the measured alternative, not the default (see msda_api.cu)
#undef MSDA_TILE_BWD
  return MSDA_E_UNSUPPORTED;
}

}  // namespace

bool tile_supported(int dtype, int N, int S, int M, int D, int L, int P) {
  if (dtype != MSDA_F32 && dtype != MSDA_BF16 && dtype != MSDA_F16) return false;
  if (D != 32 || P != 4 || L < 1 || L > 4) return false;
  return (long long)N * S * M * D < (1LL << 30);       // record offsets keep two flag bits
}

// pre: 0 = sampling locations / attention weights given; 1 / 2 / 3 = raw fp32 / bf16 / fp16 offsets and logits (msda_pre.cuh)
int launch_forward_tile(const Problem &pr, int dtype, int sms, int pre, cudaStream_t stream) {
  if (dtype == MSDA_F32) return launch_fwd<wide_f32>(pr, sms, pre, stream);
  if (dtype == MSDA_BF16) return launch_fwd<__nv_bfloat16>(pr, sms, pre, stream);
  if (dtype == MSDA_F16) return launch_fwd<__half>(pr, sms, pre, stream);
  return MSDA_E_UNSUPPORTED;
}

#ifndef MSDA_WARP_BWD_MINB
#define MSDA_WARP_BWD_MINB 3
#endif
template <typename VT>
int launch_bwd_warp(const Problem &pr, int sms, int pre, cudaStream_t stream) {
  constexpr int MB = MSDA_WARP_BWD_MINB;
#define MSDA_WARP_BWD(PREV)                                                              \
  if (pre == PREV) {                                                                     \
    if (int rc = set_smem(bwd_warp_kernel<VT, PREV, MB>, kWbSmem)) return rc;            \
    return launch_pdl(bwd_warp_kernel<VT, PREV, MB>, sms * MB, kThreads, kWbSmem, stream, pr);  \
  }
  MSDA_WARP_BWD(0) MSDA_WARP_BWD(1) MSDA_WARP_BWD(2) MSDA_WARP_BWD(3)
#undef MSDA_WARP_BWD
  return MSDA_E_UNSUPPORTED;
}

int launch_backward_warp(const Problem &pr, int dtype, int sms, int pre, cudaStream_t stream) {
  if (dtype == MSDA_F32) return launch_bwd_warp<wide_f32>(pr, sms, pre, stream);
  if (dtype == MSDA_BF16) return launch_bwd_warp<__nv_bfloat16>(pr, sms, pre, stream);
  if (dtype == MSDA_F16) return launch_bwd_warp<__half>(pr, sms, pre, stream);
  return MSDA_E_UNSUPPORTED;
}

int launch_backward_tile(const Problem &pr, int dtype, int sms, int pre, cudaStream_t stream) {
  const int grid = sms * 2;
  if (dtype == MSDA_F32) return launch_bwd<wide_f32>(pr, grid, pre, stream);
  if (dtype == MSDA_BF16) return launch_bwd<__nv_bfloat16>(pr, grid, pre, stream);
  if (dtype == MSDA_F16) return launch_bwd<__half>(pr, grid, pre, stream);
  return MSDA_E_UNSUPPORTED;
}

}  // namespace msda
