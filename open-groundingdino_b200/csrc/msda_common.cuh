// msda_common.cuh -- shared device code of libmsda_b200 (sm_100a only).
//
// Arithmetic contract (what "bit-exact sampling index" means here): the pixel coordinate of a
// sample is fl32(fl32(loc * size) - 0.5f), the sample is used iff
// h > -1 && w > -1 && h < H && w < W, and low = (int)floorf(coord).  That is the value set the
// reference kernels produce (ms_deform_im2col_cuda.cuh:285-288 and :38-41; the reference subtracts a
// double 0.5 from the rounded float product, which rounds to the same float -- see
// tests/test_oracle_cpu.py::test_fp32_coordinate_is_mul_then_sub).  __fmul_rn/__fsub_rn keep ptxas
// from contracting the pair into an FMA, which would change 1-10 % of coordinates.
#pragma once

#include <cuda_bf16.h>
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
};

// Per-CTA copy of the pyramid tables, narrowed to int like the reference does (cuh:274-277).
struct Tables {
  int H[kMaxLevels];
  int W[kMaxLevels];
  int start[kMaxLevels];       // level_start_index (rows of the value tensor)
  int qstart[kMaxLevels];      // first query of the level when queries == pixels (prefix of H*W)
  int tiles_x[kMaxLevels];
  int tile_first[kMaxLevels + 1];
  int n_tiles;                 // tiles per batch element (tiled order) -- 0 if tiling is off
  long long n_units;           // work units in the whole launch
};

__device__ __forceinline__ void load_tables(const Problem &pr, Tables &tb) {
  if (threadIdx.x == 0) {
    const int tw = 1 << pr.tile_w_log2, th = kUnitItems >> pr.tile_w_log2;
    long long area = 0;
    int tiles = 0;
    for (int l = 0; l < pr.L; ++l) {
      const int H = (int)pr.shapes[2 * l], W = (int)pr.shapes[2 * l + 1];
      tb.H[l] = H;
      tb.W[l] = W;
      tb.start[l] = (int)pr.starts[l];
      tb.qstart[l] = (int)area;
      tb.tiles_x[l] = (W + tw - 1) / tw;
      tb.tile_first[l] = tiles;
      tiles += tb.tiles_x[l] * ((H + th - 1) / th);
      area += (long long)H * W;
    }
    tb.tile_first[pr.L] = tiles;
    // The tile order is only a permutation of the work; it is valid iff the pyramid has exactly
    // one pixel per query.  Otherwise fall back to the natural (b, q, m) order.
    const bool tiled = pr.tiled && area == (long long)pr.Lq;
    tb.n_tiles = tiled ? tiles : 0;
    const long long items = (long long)pr.N * pr.Lq * pr.M;
    tb.n_units = tiled ? (long long)pr.N * tiles * pr.M : (items + kUnitItems - 1) / kUnitItems;
  }
  __syncthreads();
}

// One (query, head) item of a work unit.
struct Item {
  int b, q, m;
  bool valid;
};

// unit -> (b, tile, m); k in [0, kUnitItems) -> query inside the tile.  Consecutive units share the
// tile and differ in head, so CTAs running at the same time touch the same pixels' other heads.
__device__ __forceinline__ Item decode_item(const Problem &pr, const Tables &tb, long long unit, int k) {
  Item it;
  if (tb.n_tiles > 0) {
    it.m = (int)(unit % pr.M);
    const long long rest = unit / pr.M;
    const int t = (int)(rest % tb.n_tiles);
    it.b = (int)(rest / tb.n_tiles);
    int l = 0;
    while (l + 1 < pr.L && t >= tb.tile_first[l + 1]) ++l;
    const int tl = t - tb.tile_first[l];
    const int ty = tl / tb.tiles_x[l], tx = tl - ty * tb.tiles_x[l];
    const int y = (ty * (kUnitItems >> pr.tile_w_log2)) + (k >> pr.tile_w_log2);
    const int x = (tx << pr.tile_w_log2) + (k & ((1 << pr.tile_w_log2) - 1));
    it.valid = (y < tb.H[l]) && (x < tb.W[l]);
    it.q = tb.qstart[l] + y * tb.W[l] + x;
  } else {
    const long long idx = unit * kUnitItems + k;
    it.valid = idx < (long long)pr.N * pr.Lq * pr.M;
    it.m = (int)(idx % pr.M);
    const long long r = idx / pr.M;
    it.q = (int)(r % pr.Lq);
    it.b = (int)(r / pr.Lq);
  }
  return it;
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
  const float h = __fsub_rn(__fmul_rn(y, (float)H), 0.5f);
  const float w = __fsub_rn(__fmul_rn(x, (float)W), 0.5f);
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

// ---- 16-byte channel vectors ---------------------------------------------------------------------
template <typename VT>
struct Chan;

template <>
struct Chan<float> {
  static constexpr int kPerLane = 4;
  static __device__ __forceinline__ void unpack(const uint4 &u, float (&f)[4]) {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y);
    f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  }
  static __device__ __forceinline__ uint4 pack(const float (&f)[4]) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
};

template <>
struct Chan<__nv_bfloat16> {
  static constexpr int kPerLane = 8;
  // bf16 -> fp32 is a 16-bit shift: exact, one ALU op per channel.
  static __device__ __forceinline__ void unpack(const uint4 &u, float (&f)[8]) {
    f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
    f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
    f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
    f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
  }
  static __device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&h);
  }
  static __device__ __forceinline__ uint4 pack(const float (&f)[8]) {
    return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
  }
};

// Read-only 16-byte gather; a false predicate yields zeros without touching memory.
__device__ __forceinline__ uint4 gather16(const void *p, bool pred) {
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (pred) v = __ldg(reinterpret_cast<const uint4 *>(p));
  return v;
}

// Vector reduction: one 16-byte fp32x4 add performed by the L2 atomic unit (REDG.E.ADD.F32x4).
__device__ __forceinline__ void red_add_f32x4(float *p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

__device__ __forceinline__ int pow2_ceil(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

}  // namespace msda
