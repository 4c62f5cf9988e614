// msda_common.cuh -- shared device code of libmsda_b200 (sm_100a only).
//
// Arithmetic contract (what "bit-exact sampling index" means here): the pixel coordinate of a
// sample is the single-rounded fl32(loc * size - 0.5), i.e. fmaf(loc, (float)size, -0.5f); the sample is
// used iff h > -1 && w > -1 && h < H && w < W, and low = (int)floorf(coord).  That is what the
// reference kernels compute AS COMPILED for this GPU: the source reads `loc_h * spatial_h - 0.5`
// (ms_deform_im2col_cuda.cuh:285-286, inside test :288, floor :38-41) and nvcc 12.9 -O3 folds the
// double literal away and contracts the pair into `FFMA R, loc, size, -0.5` in the forward and in
// every backward kernel (cuobjdump of oracle/_ref/msda_ref_cuda.so; excerpt in
// profiles/reference_sass_coordinate.txt).  An unfused multiply-then-subtract differs in 1-10 % of
// coordinates and moves samples that sit on integer pixel coordinates into the neighbouring cell,
// which changes grad_sampling_loc by O(1) there (tests/test_gpu_vs_reference_cuda.py).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "msda_b200.h"

namespace msda {

constexpr int kMaxLevels = 16;   // pyramid levels the vector kernels keep in shared memory
constexpr int kUnitItems = 64;   // (query, head) items per work unit == queries per query tile
constexpr int kThreads = 256;    // threads per CTA of the vector kernels

// Kernel parameter block (passed by value; lives in constant bank 0).
struct Problem {
  const void *value;
  const int64_t *shapes;   // (L,2) int64 (H,W), device
  const int64_t *starts;   // (L) int64, device
  const void *loc;
  const void *attn;
  void *out;               // forward only
  const void *grad_out;    // backward only
  void *grad_value_acc;    // backward: fp32 (fp64 for F64) accumulator, zeroed by the caller
  void *grad_loc;
  void *grad_attn;
  int N, S, M, D, L, Lq, P;
  int tiled;               // 1: enumerate queries as 2-D tiles of their own pyramid level
  int tile_w_log2;         // tile is (kUnitItems >> tile_w_log2) rows x (1 << tile_w_log2) columns
  int acc_permuted;        // backward, bf16 value: accumulator rows use acc_pos() order
  // fused pre-op (msda_pre.cuh): loc/attn then hold raw offsets/logits, grad_loc/grad_attn their gradients
  const float *ref;        // reference points (N, Lq, L, ref_dim) fp32
  int ref_dim;             // 2 or 4
  // value rows to treat as zero (MSDeformAttn's input_padding_mask, ops/modules/ms_deform_attn.py:95-96): (N, S) bytes,
  // non-zero = padding; such rows gather as zero and receive no gradient.  nullptr = no mask.
  const unsigned char *pad_mask;
  int win_margin;          // tile kernels: halo (pixels) of the per-tile aggregation windows (msda_tile.cu)
};

// Position of channel c inside a head's fp32 accumulator row when the row is permuted (bf16 value,
// D % 8 == 0): lane j = c / 8 owns channels 8j..8j+7; its first four go to [4j, 4j+4), its last four
// to [D/2 + 4j, D/2 + 4j + 4).
__host__ __device__ inline int acc_pos(int c, int D) {
  const int j = c >> 3, t = c & 7;
  return t < 4 ? 4 * j + t : D / 2 + 4 * j + (t - 4);
}

// Per-CTA copy of the pyramid tables, narrowed to int like the reference does (cuh:274-277).
struct Tables {
  int H[kMaxLevels];
  int W[kMaxLevels];
  int start[kMaxLevels];       // level_start_index (rows of the value tensor)
  int qstart[kMaxLevels];      // first query of the level when queries == pixels (prefix of H*W)
  int tiles_x[kMaxLevels];
  int tile_first[kMaxLevels + 1];
  int n_tiles;                 // tiles per batch element (tiled order) -- 0 if tiling is off
  int n_items;                 // N * Lq * M
  int n_units;                 // work units in the whole launch
  int unit_items;              // items per unit: kUnitItems for query tiles, one CTA pass in linear order
  // reciprocals for the divisions of decode_unit (fast_divmod): by the head count, the tiles per image, the tiles per row
  unsigned magic_M, magic_T, magic_tx[kMaxLevels];
};

// q = x / d, r = x % d for x < 2^31 without an integer division: q' = umulhi(x, m) with m = 0xffffffff / d is floor(x / d) or
// one less (x * (2^32 / d - m) / 2^32 < 1), so one conditional correction makes it exact (also for d = 1, where q' = x - 1).
// Same-box A/B of two builds: tile forward 278.1 -> 271.5 us at cfg-2, backward 750.0 -> 747.6 us.
__host__ __device__ inline unsigned magic_of(unsigned d) { return 0xffffffffu / (d ? d : 1u); }
__device__ __forceinline__ void fast_divmod(unsigned x, unsigned d, unsigned m, unsigned &q, unsigned &r) {
  q = __umulhi(x, m);
  r = x - q * d;
  if (r >= d) { ++q; r -= d; }
}

// linear_unit: items per unit when the queries are walked in their natural order -- one pass of the CTA's warps, so that
// small launches (decoder: 57,600 items on 296-444 persistent CTAs) split into enough units to balance: with 64-item
// units the last of 3 rounds ran 12 of 444 CTAs.
__device__ __forceinline__ void load_tables(const Problem &pr, Tables &tb, int linear_unit) {
  // one thread per level fetches its three int64 entries (independent loads, one round trip) ...
  if (threadIdx.x < pr.L) {
    const int l = threadIdx.x;
    tb.H[l] = (int)pr.shapes[2 * l];
    tb.W[l] = (int)pr.shapes[2 * l + 1];
    tb.start[l] = (int)pr.starts[l];
  }
  __syncthreads();
  // ... then one thread derives the prefix tables from shared memory
  if (threadIdx.x == 0) {
    const int tw = 1 << pr.tile_w_log2, th = kUnitItems >> pr.tile_w_log2;
    long long area = 0;
    int tiles = 0;
    for (int l = 0; l < pr.L; ++l) {
      const int H = tb.H[l], W = tb.W[l];
      tb.qstart[l] = (int)area;
      tb.tiles_x[l] = (W + tw - 1) / tw;
      tb.magic_tx[l] = magic_of((unsigned)tb.tiles_x[l]);
      tb.tile_first[l] = tiles;
      tiles += tb.tiles_x[l] * ((H + th - 1) / th);
      area += (long long)H * W;
    }
    tb.tile_first[pr.L] = tiles;
    // The tile order is only a permutation of the work; it is valid iff the pyramid has exactly
    // one pixel per query.  Otherwise fall back to the natural (b, q, m) order.
    const bool tiled = pr.tiled && area == (long long)pr.Lq;
    tb.n_tiles = tiled ? tiles : 0;
    tb.magic_M = magic_of((unsigned)pr.M);
    tb.magic_T = magic_of((unsigned)(tiles > 0 ? tiles : 1));
    tb.n_items = pr.N * pr.Lq * pr.M;      // < 2^31, checked by the host
    tb.unit_items = tiled ? kUnitItems : linear_unit;
    tb.n_units = tiled ? pr.N * tiles * pr.M : (tb.n_items + linear_unit - 1) / linear_unit;
  }
  __syncthreads();
}

// One (query, head) item of a work unit.
struct Item {
  int b, q, m;
  int item;      // ((b * Lq) + q) * M + m
  bool valid;
};

// Everything about a work unit that does not depend on the item inside it.  Decoded once per unit
// (the divisions live here), so the per-item cost is a handful of integer ops.
struct UnitCtx {
  int b, m;              // tiled order: fixed per unit
  int y0, x0, Hq, Wq;    // tiled order: tile origin and the query level's extent
  int q0;                // tiled order: first query of that level
  int lq;                // tiled order: that level
  int idx0;              // linear order: first item of the unit
};

// unit -> (b, tile, m).  Consecutive units share the tile and differ in head, so CTAs running at
// the same time touch the same pixels' other heads (adjacent 64/128-byte rows in HBM).
__device__ __forceinline__ UnitCtx decode_unit(const Problem &pr, const Tables &tb, int unit) {
  UnitCtx uc;
  if (tb.n_tiles > 0) {
    unsigned rest, mu, bu, tu;
    fast_divmod((unsigned)unit, (unsigned)pr.M, tb.magic_M, rest, mu);
    fast_divmod(rest, (unsigned)tb.n_tiles, tb.magic_T, bu, tu);
    uc.m = (int)mu;
    uc.b = (int)bu;
    const int t = (int)tu;
    int l = 0;
    while (l + 1 < pr.L && t >= tb.tile_first[l + 1]) ++l;
    const int tl = t - tb.tile_first[l];
    unsigned tyu, txu;
    fast_divmod((unsigned)tl, (unsigned)tb.tiles_x[l], tb.magic_tx[l], tyu, txu);
    const int ty = (int)tyu, tx = (int)txu;
    uc.y0 = ty * (kUnitItems >> pr.tile_w_log2);
    uc.x0 = tx << pr.tile_w_log2;
    uc.Hq = tb.H[l];
    uc.Wq = tb.W[l];
    uc.q0 = tb.qstart[l];
    uc.lq = l;
    uc.idx0 = 0;
  } else {
    uc.b = uc.m = uc.y0 = uc.x0 = uc.Hq = uc.Wq = uc.q0 = uc.lq = 0;
    uc.idx0 = unit * tb.unit_items;
  }
  return uc;
}

// k in [0, tb.unit_items) -> item of the unit.
__device__ __forceinline__ Item decode_item(const Problem &pr, const Tables &tb, const UnitCtx &uc, int k) {
  Item it;
  if (tb.n_tiles > 0) {
    const int y = uc.y0 + (k >> pr.tile_w_log2);
    const int x = uc.x0 + (k & ((1 << pr.tile_w_log2) - 1));
    it.valid = (y < uc.Hq) && (x < uc.Wq);
    it.b = uc.b;
    it.m = uc.m;
    it.q = uc.q0 + y * uc.Wq + x;
    it.item = (it.b * pr.Lq + it.q) * pr.M + it.m;
  } else {
    const unsigned idx = (unsigned)(uc.idx0 + k);
    it.valid = idx < (unsigned)tb.n_items;
    it.item = (int)idx;
    it.m = (int)(idx % (unsigned)pr.M);
    const unsigned r = idx / (unsigned)pr.M;
    it.b = (int)(r / (unsigned)pr.Lq);
    it.q = (int)(r - (unsigned)it.b * (unsigned)pr.Lq);
  }
  return it;
}

// Work list of a persistent CTA.  Whole units are dealt round-robin; the units of the last, partial round are dealt
// PASS by pass (a pass = the CTA's warps once over 8 * IPW items) so that the tail costs one pass instead of one unit:
// 11,104 tile units on 444 CTAs are 25 full rounds plus 4 units -- without the split 440 CTAs idle for a whole unit.
struct WorkItem {
  int unit, pass_begin, pass_end;
};
__device__ __forceinline__ int work_count(const Tables &tb, int passes_per_unit) {
  const int full = (tb.n_units / (int)gridDim.x) * (int)gridDim.x;
  return full + (tb.n_units - full) * passes_per_unit;
}
__device__ __forceinline__ WorkItem work_item(const Tables &tb, int passes_per_unit, int w) {
  const int full = (tb.n_units / (int)gridDim.x) * (int)gridDim.x;
  WorkItem wi;
  if (w < full) {
    wi.unit = w; wi.pass_begin = 0; wi.pass_end = passes_per_unit;
  } else {
    const int t = w - full;
    wi.unit = full + t / passes_per_unit;
    wi.pass_begin = t % passes_per_unit;
    wi.pass_end = wi.pass_begin + 1;
  }
  return wi;
}

// Programmatic dependent launch (launch_pdl in msda_launch.h): block until the previous kernel of the stream has completed
// and flushed, then let the next kernel of the stream be scheduled as this one's CTAs retire.  A no-op in a kernel that
// was launched without the attribute.
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---- sampler geometry --------------------------------------------------------------------------
template <typename T>
struct Cell {
  int h_low, w_low;
  T lh, lw;
  bool inside;
};

__device__ __forceinline__ Cell<float> locate(float x, float y, int H, int W) {
  Cell<float> c;
  const float h = __fmaf_rn(y, (float)H, -0.5f);
  const float w = __fmaf_rn(x, (float)W, -0.5f);
  c.inside = (h > -1.0f) && (w > -1.0f) && (h < (float)H) && (w < (float)W);
  const float hf = floorf(h), wf = floorf(w);
  c.h_low = (int)hf;
  c.w_low = (int)wf;
  c.lh = __fsub_rn(h, hf);
  c.lw = __fsub_rn(w, wf);
  return c;
}

// fp64 twin.  The reference expression `loc * size - 0.5` is all-double there, which nvcc contracts.
__device__ __forceinline__ Cell<double> locate(double x, double y, int H, int W) {
  Cell<double> c;
  const double h = fma(y, (double)H, -0.5);
  const double w = fma(x, (double)W, -0.5);
  c.inside = (h > -1.0) && (w > -1.0) && (h < (double)H) && (w < (double)W);
  const double hf = floor(h), wf = floor(w);
  c.h_low = (int)hf;
  c.w_low = (int)wf;
  c.lh = h - hf;
  c.lw = w - wf;
  return c;
}

// bit0 = (low,low) bit1 = (low,high) bit2 = (high,low) bit3 = (high,high); cuh:56,62,68,74
__device__ __forceinline__ int corner_mask(int h_low, int w_low, int H, int W) {
  const bool top = h_low >= 0, left = w_low >= 0;
  const bool bottom = h_low + 1 <= H - 1, right = w_low + 1 <= W - 1;
  return (int)(top && left) | ((int)(top && right) << 1) | ((int)(bottom && left) << 2) |
         ((int)(bottom && right) << 3);
}

// input_padding_mask inside the kernels (scope row f4): neighbours whose value row is padding are treated like neighbours
// outside the image -- they gather as zero and receive no gradient, which is what MSDeformAttn's
// value.masked_fill(mask, 0) (ops/modules/ms_deform_attn.py:95-96) and its backward do.  `row00` = b*S + level start +
// h_low*W + w_low.
__device__ __forceinline__ int drop_padded(const unsigned char *__restrict__ pad, int mask, int row00, int W) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
    if ((mask >> kk) & 1)
      if (__ldg(pad + row00 + (kk & 1) + ((kk & 2) ? W : 0)) != 0) mask &= ~(1 << kk);
  return mask;
}

// Packed fp32x2 math (sm_100+: one FFMA2 issue slot for two lanes of the accumulator).
__device__ __forceinline__ unsigned long long pack_f2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f2(unsigned long long v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b,
                                                    unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// Read-only 16-byte gather; a false predicate yields zeros without touching memory.
__device__ __forceinline__ uint4 gather16(const void *p, bool pred) {
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (pred) v = __ldg(reinterpret_cast<const uint4 *>(p));
  return v;
}

// A pointer the compiler must take as given.  Left to itself it splits `base + row_off + lane part` back into the uniform
// kernel-parameter base and a 64-bit index, and rebuilds EVERY address from them (sign extension + two 64-bit adds: four to five
// ALU instructions per gather or reduction address); behind this fence an address is the pointer plus a 32-bit offset (two).
template <typename T>
__device__ __forceinline__ T *as_given(T *p) {
  asm("" : "+l"(p));
  return p;
}

// fp32 rows handled 8 channels (32 bytes) per lane: half the lanes per row, so a warp pass covers two items
// and the per-item geometry / bookkeeping instructions halve (the fp32 kernels are issue-bound).
struct wide_f32 {
  float x;
};
struct Raw32 {
  uint4 lo, hi;
};

// ---- 16-byte channel vectors ---------------------------------------------------------------------
template <typename VT>
struct Chan;

template <>
struct Chan<float> {
  using Raw = uint4;
  static __device__ __forceinline__ Raw gather(const float *p, bool pred) { return gather16(p, pred); }
  static __device__ __forceinline__ Raw gather_plain(const float *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
  static __device__ __forceinline__ void store(float *p, const Raw &r) { *reinterpret_cast<uint4 *>(p) = r; }
  static constexpr int kPerLane = 4;
  static __device__ __forceinline__ void unpack(const uint4 &u, float (&f)[4]) {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y);
    f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  }
  static __device__ __forceinline__ uint4 pack(const float (&f)[4]) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
  static constexpr int kPairs = 2;
  static __device__ __forceinline__ void unpack2(const uint4 &u, unsigned long long (&p)[2]) {
    p[0] = pack_f2(__uint_as_float(u.x), __uint_as_float(u.y));
    p[1] = pack_f2(__uint_as_float(u.z), __uint_as_float(u.w));
  }
};

template <>
struct Chan<wide_f32> {
  using Raw = Raw32;
  static constexpr int kPerLane = 8;
  static constexpr int kPairs = 4;
  static __device__ __forceinline__ Raw gather(const wide_f32 *p, bool pred) {
    Raw r;
    r.lo = gather16(p, pred);
    r.hi = gather16(p + 4, pred);
    return r;
  }
  static __device__ __forceinline__ Raw gather_plain(const wide_f32 *p) {
    Raw r;
    r.lo = __ldg(reinterpret_cast<const uint4 *>(p));
    r.hi = __ldg(reinterpret_cast<const uint4 *>(p) + 1);
    return r;
  }
  static __device__ __forceinline__ void store(wide_f32 *p, const Raw &r) {
    reinterpret_cast<uint4 *>(p)[0] = r.lo;
    reinterpret_cast<uint4 *>(p)[1] = r.hi;
  }
  static __device__ __forceinline__ void unpack2(const Raw &u, unsigned long long (&q)[4]) {
    q[0] = pack_f2(__uint_as_float(u.lo.x), __uint_as_float(u.lo.y));
    q[1] = pack_f2(__uint_as_float(u.lo.z), __uint_as_float(u.lo.w));
    q[2] = pack_f2(__uint_as_float(u.hi.x), __uint_as_float(u.hi.y));
    q[3] = pack_f2(__uint_as_float(u.hi.z), __uint_as_float(u.hi.w));
  }
  static __device__ __forceinline__ Raw pack(const float (&f)[8]) {
    Raw r;
    r.lo = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    r.hi = make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7]));
    return r;
  }
};

template <>
struct Chan<__nv_bfloat16> {
  using Raw = uint4;
  static __device__ __forceinline__ Raw gather(const __nv_bfloat16 *p, bool pred) { return gather16(p, pred); }
  static __device__ __forceinline__ Raw gather_plain(const __nv_bfloat16 *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
  static __device__ __forceinline__ void store(__nv_bfloat16 *p, const Raw &r) { *reinterpret_cast<uint4 *>(p) = r; }
  static constexpr int kPerLane = 8;
  // bf16 -> fp32 is a 16-bit shift: exact, one ALU op per channel.
  static __device__ __forceinline__ void unpack(const uint4 &u, float (&f)[8]) {
    f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
    f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
    f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
    f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
  }
  static constexpr int kPairs = 4;
  static __device__ __forceinline__ void unpack2(const uint4 &u, unsigned long long (&p)[4]) {
    p[0] = pack_f2(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u));
    p[1] = pack_f2(__uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
    p[2] = pack_f2(__uint_as_float(u.z << 16), __uint_as_float(u.z & 0xffff0000u));
    p[3] = pack_f2(__uint_as_float(u.w << 16), __uint_as_float(u.w & 0xffff0000u));
  }
  static __device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&h);
  }
  static __device__ __forceinline__ uint4 pack(const float (&f)[8]) {
    return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
  }
};

template <>
struct Chan<__half> {
  using Raw = uint4;
  static __device__ __forceinline__ Raw gather(const __half *p, bool pred) { return gather16(p, pred); }
  static __device__ __forceinline__ Raw gather_plain(const __half *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
  static __device__ __forceinline__ void store(__half *p, const Raw &r) { *reinterpret_cast<uint4 *>(p) = r; }
  static constexpr int kPerLane = 8;
  static constexpr int kPairs = 4;
  static __device__ __forceinline__ unsigned long long widen2(uint32_t w) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w));
    return pack_f2(f.x, f.y);
  }
  static __device__ __forceinline__ void unpack2(const uint4 &u, unsigned long long (&p)[4]) {
    p[0] = widen2(u.x); p[1] = widen2(u.y); p[2] = widen2(u.z); p[3] = widen2(u.w);
  }
  static __device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&h);
  }
  static __device__ __forceinline__ uint4 pack(const float (&f)[8]) {
    return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
  }
};

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

// Vector reduction: one 16-byte fp32x4 add performed by the L2 atomic unit (REDG.E.ADD.F32x4).
__device__ __forceinline__ void red_add_f32x4(float *p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// ---- cp.async staging of the per-item sampling positions / attention weights (model-geometry kernels) ----------------
// Each warp copies the inputs of ITS items of the NEXT work unit into a private shared-memory slot while it works on the
// current unit: the first use of the sampling positions was the longest single stall of both vector kernels (17 % of
// the forward's stall samples, 10 % of the backward's).  Warp-private, so no block barrier enters the unit loop.
constexpr int kStageItemFloats = 52;   // 32 position floats + 16 weights (+4 pad: the two items of a pass hit disjoint banks)
constexpr int kStageSlots = 8;         // items of one unit handled by one warp (64 items / 8 warps)
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ int pow2_ceil(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

}  // namespace msda
