/*
 * msda_oracle.c -- CPU restatement of the reference's multi-scale deformable attention op.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under open-groundingdino_b200/ may import, link or call
 * this file; it exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * have something independent of the CUDA path to compare against.
 *
 * What it follows (paths relative to /root/reference/models/GroundingDINO/ops/src/cuda/):
 *   forward   : ms_deform_im2col_cuda.cuh:237-299 (per-output accumulation, levels outer /
 *               points inner) and :33-84 (bilinear sampler, zero padding, 4 range checks)
 *   backward  : ms_deform_im2col_cuda.cuh:87-159 (per-sample, per-channel gradient terms) and
 *               :301-403 (channel sums for grad_sampling_loc / grad_attn_weight are taken in
 *               ascending channel order, like the reference's lane-0 serial loop)
 *   geometry  : ms_deform_im2col_cuda.cuh:285-288 (coordinate + inside test) and :38-41 (floor /
 *               low / high indices).  The coordinate `loc * size - 0.5` is evaluated with ONE
 *               rounding (fmaf / fma): that is what the reference computes as compiled by nvcc 12.9
 *               for sm_100a -- every kernel of oracle/_ref/msda_ref_cuda.so holds
 *               `FFMA R, loc, size, -0.5` (profiles/reference_sass_coordinate.txt) -- and it is
 *               confirmed on the GPU by grad_sampling_loc parity with the reference op on samples
 *               that sit on integer pixel coordinates (tests/test_gpu_vs_reference_cuda.py).
 *   layouts   : ms_deform_attn_cuda.cu:40-60 (value N,S,M,D; loc N,Lq,M,L,P,2 as (x,y);
 *               attn N,Lq,M,L,P; spatial_shapes (H,W) int64; level_start_index int64)
 *
 * Parity status: PINNED.  tests/test_oracle_cpu.py checks these functions against golden vectors
 * generated in the authoring container by the reference's own ms_deform_attn_core_pytorch
 * (functions/ms_deform_attn_func.py:41-61) + torch autograd in fp64 (tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off  (see oracle/Makefile)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* One sample's cell geometry.  Mirrors cuh:38-46 for either precision. */
typedef struct {
  int h_low, w_low;
  int inside;      /* cuh:288 */
  double lh, lw;   /* fractional parts, held in the working precision's value set */
} cell_t;

/* fp32 path: single-rounded loc*size - 0.5 (see header). */
static inline cell_t locate_f32(float loc_x, float loc_y, int H, int W) {
  cell_t c;
  const float h_im = fmaf(loc_y, (float)H, -0.5f);
  const float w_im = fmaf(loc_x, (float)W, -0.5f);
  c.inside = (h_im > -1.0f) && (w_im > -1.0f) && (h_im < (float)H) && (w_im < (float)W);
  c.h_low = 0; c.w_low = 0; c.lh = 0; c.lw = 0;
  if (c.inside) {
    c.h_low = (int)floorf(h_im);
    c.w_low = (int)floorf(w_im);
    c.lh = (double)(float)(h_im - (float)c.h_low);
    c.lw = (double)(float)(w_im - (float)c.w_low);
  }
  return c;
}

static inline cell_t locate_f64(double loc_x, double loc_y, int H, int W) {
  cell_t c;
  const double h_im = fma(loc_y, (double)H, -0.5);
  const double w_im = fma(loc_x, (double)W, -0.5);
  c.inside = (h_im > -1.0) && (w_im > -1.0) && (h_im < (double)H) && (w_im < (double)W);
  c.h_low = 0; c.w_low = 0; c.lh = 0; c.lw = 0;
  if (c.inside) {
    c.h_low = (int)floor(h_im);
    c.w_low = (int)floor(w_im);
    c.lh = h_im - (double)c.h_low;
    c.lw = w_im - (double)c.w_low;
  }
  return c;
}

/* Which of the four neighbours exist (cuh:56,62,68,74). bit0=v1(low,low) bit1=v2(low,high)
 * bit2=v3(high,low) bit3=v4(high,high). */
static inline int corner_mask(const cell_t *c, int H, int W) {
  const int top = c->h_low >= 0, left = c->w_low >= 0;
  const int bottom = c->h_low + 1 <= H - 1, right = c->w_low + 1 <= W - 1;
  return (top && left) | ((top && right) << 1) | ((bottom && left) << 2) | ((bottom && right) << 3);
}

#define MSDA_ORACLE_IMPL(T, SUF, LOCATE)                                                          \
  int msda_oracle_forward_##SUF(const T *value, const int64_t *shapes, const int64_t *starts,     \
                                const T *loc, const T *attn, int N, int S, int M, int D, int L,   \
                                int Lq, int P, T *out) {                                          \
    const long row = (long)M * D;                                                                 \
    _Pragma("omp parallel for collapse(2) schedule(static)")                                      \
    for (int b = 0; b < N; ++b) {                                                                 \
      for (int q = 0; q < Lq; ++q) {                                                              \
        for (int m = 0; m < M; ++m) {                                                             \
          const long item = ((long)b * Lq + q) * M + m;                                           \
          const T *loc_i = loc + item * L * P * 2;                                                \
          const T *attn_i = attn + item * L * P;                                                  \
          T *out_i = out + item * D;                                                              \
          for (int c = 0; c < D; ++c) out_i[c] = (T)0;                                            \
          for (int l = 0; l < L; ++l) {                                                           \
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                         \
            const T *slab = value + ((long)b * S + (int)starts[l]) * row + (long)m * D;           \
            for (int p = 0; p < P; ++p) {                                                         \
              const T lx = loc_i[(l * P + p) * 2], ly = loc_i[(l * P + p) * 2 + 1];               \
              const T a = attn_i[l * P + p];                                                      \
              const cell_t g = LOCATE(lx, ly, H, W);                                              \
              if (!g.inside) continue;                                                            \
              const int mask = corner_mask(&g, H, W);                                             \
              const T lh = (T)g.lh, lw = (T)g.lw, hh = (T)1 - lh, hw = (T)1 - lw;                 \
              const T w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;                     \
              const T *r0 = slab + ((long)g.h_low * W + g.w_low) * row;                           \
              const T *r1 = r0 + (long)W * row;                                                   \
              for (int c = 0; c < D; ++c) {                                                       \
                const T v1 = (mask & 1) ? r0[c] : (T)0, v2 = (mask & 2) ? r0[row + c] : (T)0;     \
                const T v3 = (mask & 4) ? r1[c] : (T)0, v4 = (mask & 8) ? r1[row + c] : (T)0;     \
                out_i[c] += (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4) * a;                          \
              }                                                                                   \
            }                                                                                     \
          }                                                                                       \
        }                                                                                         \
      }                                                                                           \
    }                                                                                             \
    return 0;                                                                                     \
  }                                                                                               \
                                                                                                  \
  int msda_oracle_backward_##SUF(const T *value, const int64_t *shapes, const int64_t *starts,    \
                                 const T *loc, const T *attn, const T *grad_out, int N, int S,    \
                                 int M, int D, int L, int Lq, int P, T *grad_value, T *grad_loc,  \
                                 T *grad_attn) {                                                  \
    const long row = (long)M * D;                                                                 \
    memset(grad_value, 0, sizeof(T) * (size_t)N * S * row);                                       \
    memset(grad_loc, 0, sizeof(T) * (size_t)N * Lq * M * L * P * 2);                              \
    memset(grad_attn, 0, sizeof(T) * (size_t)N * Lq * M * L * P);                                 \
    /* (b,m) slabs of grad_value are disjoint, so this is race-free and run-to-run stable. */     \
    _Pragma("omp parallel for collapse(2) schedule(static)")                                      \
    for (int b = 0; b < N; ++b) {                                                                 \
      for (int m = 0; m < M; ++m) {                                                               \
        for (int q = 0; q < Lq; ++q) {                                                            \
          const long item = ((long)b * Lq + q) * M + m;                                           \
          const T *g_i = grad_out + item * D;                                                     \
          for (int l = 0; l < L; ++l) {                                                           \
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                         \
            const long base = ((long)b * S + (int)starts[l]) * row + (long)m * D;                 \
            for (int p = 0; p < P; ++p) {                                                         \
              const long s = item * L * P + l * P + p;                                            \
              const T lx = loc[2 * s], ly = loc[2 * s + 1], a = attn[s];                          \
              const cell_t g = LOCATE(lx, ly, H, W);                                              \
              if (!g.inside) continue;                                                            \
              const int mask = corner_mask(&g, H, W);                                             \
              const T lh = (T)g.lh, lw = (T)g.lw, hh = (T)1 - lh, hw = (T)1 - lw;                 \
              const T w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;                     \
              const long o0 = base + ((long)g.h_low * W + g.w_low) * row;                         \
              const long o1 = o0 + (long)W * row;                                                 \
              T sum_w = (T)0, sum_h = (T)0, sum_a = (T)0;                                         \
              for (int c = 0; c < D; ++c) {                                                       \
                const T top = g_i[c], tgv = top * a;                                              \
                T gh = (T)0, gw = (T)0, v1 = (T)0, v2 = (T)0, v3 = (T)0, v4 = (T)0;               \
                if (mask & 1) { v1 = value[o0 + c]; gh -= hw * v1; gw -= hh * v1;                 \
                                grad_value[o0 + c] += w1 * tgv; }                                 \
                if (mask & 2) { v2 = value[o0 + row + c]; gh -= lw * v2; gw += hh * v2;           \
                                grad_value[o0 + row + c] += w2 * tgv; }                           \
                if (mask & 4) { v3 = value[o1 + c]; gh += hw * v3; gw -= lh * v3;                 \
                                grad_value[o1 + c] += w3 * tgv; }                                 \
                if (mask & 8) { v4 = value[o1 + row + c]; gh += lw * v4; gw += lh * v4;           \
                                grad_value[o1 + row + c] += w4 * tgv; }                           \
                sum_a += top * (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);                           \
                sum_w += (T)W * gw * tgv;                                                         \
                sum_h += (T)H * gh * tgv;                                                         \
              }                                                                                   \
              grad_loc[2 * s] = sum_w;                                                            \
              grad_loc[2 * s + 1] = sum_h;                                                        \
              grad_attn[s] = sum_a;                                                               \
            }                                                                                     \
          }                                                                                       \
        }                                                                                         \
      }                                                                                           \
    }                                                                                             \
    return 0;                                                                                     \
  }

MSDA_ORACLE_IMPL(float, f32, locate_f32)
MSDA_ORACLE_IMPL(double, f64, locate_f64)

/*
 * Integer side of the fp32 path, one record of 4 int32 per sample (N,Lq,M,L,P,4):
 *   [0] h_low  [1] w_low  (0 when the sample is rejected)
 *   [2] flags: bit0 = inside test passed, bits1..4 = corner_mask
 *   [3] row offset of the level's slab for this batch element = b*S + level_start[l]
 * This is what tests/test_index_exact.py compares bit-for-bit with the CUDA path.
 */
int msda_oracle_sample_index_f32(const int64_t *shapes, const int64_t *starts, const float *loc,
                                 int N, int S, int M, int L, int Lq, int P, int32_t *rec) {
  const long per_item = (long)L * P;
#pragma omp parallel for schedule(static)
  for (long item = 0; item < (long)N * Lq * M; ++item) {
    const int b = (int)(item / ((long)Lq * M));
    for (int l = 0; l < L; ++l) {
      const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
      for (int p = 0; p < P; ++p) {
        const long s = item * per_item + l * P + p;
        const cell_t g = locate_f32(loc[2 * s], loc[2 * s + 1], H, W);
        int32_t *r = rec + 4 * s;
        r[0] = g.h_low;
        r[1] = g.w_low;
        r[2] = g.inside ? (1 | (corner_mask(&g, H, W) << 1)) : 0;
        r[3] = (int32_t)((long)b * S + (int)starts[l]);
      }
    }
  }
  return 0;
}

int msda_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
