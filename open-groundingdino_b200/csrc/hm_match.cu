// hm_match.cu -- on-device Hungarian matcher (include/hungarian_b200.h).
//
// Replaces, for a whole batch of (output set, image) problems in two launches and without any host round trip,
//   * the ~40 elementwise / cdist / gemv launches per image that build the matching cost
//     (reference models/GroundingDINO/matcher.py:66-103, util/box_ops.py:9-14,25-62)   -> cost_kernel
//   * `C.cpu()` + scipy.optimize.linear_sum_assignment per image (matcher.py:101,104-107) -> lsap_kernel
//
// cost_kernel: a CTA owns 64 prediction rows of one problem.  The targets' positive-map rows are normalised once per CTA
// into shared memory (32 targets at a time) and their non-zero tokens compacted; a positive map marks the few tokens of
// one phrase, so a (prediction, target) pair normally costs one thread a handful of exp/log evaluations plus the L1 and
// GIoU terms (sparse path); wide rows fall back to one warp per prediction row with lane-strided dot products (dense
// path).  The arithmetic follows the PyTorch expressions operation by operation in fp32 (no contraction across the
// reference's ops).
//
// lsap_kernel: one CTA per problem runs the rectangular shortest-augmenting-path solver scipy uses (Crouse 2016), with
// the reduced costs and dual variables in fp64 exactly as scipy computes them, so the result -- including the choice among
// tied columns, which depends on scipy's reverse-filled / swap-compacted `remaining` array -- is the one scipy returns.
// All per-column state lives in shared memory; a Dijkstra step is a strided sweep over the remaining columns plus one
// block-wide argmin.  The step count (sum of the augmenting-path lengths) makes the kernel latency-bound, not
// bandwidth-bound: what it buys is the removal of the device->host synchronisation, not throughput.
#include <cuda_runtime.h>
#include <limits.h>
#include <stdlib.h>
#include <math_constants.h>
#include <stdint.h>

#include "hungarian_b200.h"
#include "msda_b200.h"

namespace hm {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kRowsPerCta = 64;      // prediction rows per CTA of the cost kernel
constexpr int kChunk = 32;           // targets normalised into shared memory at a time
constexpr int kMaxColumns = 4096;    // max(Q, T) of the solver: 40 bytes of shared memory per column / row

__device__ __forceinline__ float nan_max(float a, float b) { return (a != a || b != b) ? CUDART_NAN_F : fmaxf(a, b); }
__device__ __forceinline__ float nan_min(float a, float b) { return (a != a || b != b) ? CUDART_NAN_F : fminf(a, b); }
__device__ __forceinline__ float clamp0(float x) { return x < 0.f ? 0.f : x; }   // torch.clamp(min=0): NaN stays NaN

// matcher.py:98 with box_ops.py:9-14 (cxcywh -> xyxy) and :25-62 (IoU, hull, GIoU), and :82 (L1 distance)
__device__ __forceinline__ void box_costs(const float4 p, const float4 t, float &l1, float &giou) {
  l1 = __fadd_rn(__fadd_rn(__fadd_rn(fabsf(__fsub_rn(p.x, t.x)), fabsf(__fsub_rn(p.y, t.y))), fabsf(__fsub_rn(p.z, t.z))),
                 fabsf(__fsub_rn(p.w, t.w)));
  const float px0 = __fsub_rn(p.x, __fmul_rn(0.5f, p.z)), py0 = __fsub_rn(p.y, __fmul_rn(0.5f, p.w));
  const float px1 = __fadd_rn(p.x, __fmul_rn(0.5f, p.z)), py1 = __fadd_rn(p.y, __fmul_rn(0.5f, p.w));
  const float tx0 = __fsub_rn(t.x, __fmul_rn(0.5f, t.z)), ty0 = __fsub_rn(t.y, __fmul_rn(0.5f, t.w));
  const float tx1 = __fadd_rn(t.x, __fmul_rn(0.5f, t.z)), ty1 = __fadd_rn(t.y, __fmul_rn(0.5f, t.w));
  const float area_p = __fmul_rn(__fsub_rn(px1, px0), __fsub_rn(py1, py0));
  const float area_t = __fmul_rn(__fsub_rn(tx1, tx0), __fsub_rn(ty1, ty0));
  const float iw = clamp0(__fsub_rn(nan_min(px1, tx1), nan_max(px0, tx0)));
  const float ih = clamp0(__fsub_rn(nan_min(py1, ty1), nan_max(py0, ty0)));
  const float inter = __fmul_rn(iw, ih);
  const float uni = __fsub_rn(__fadd_rn(area_p, area_t), inter);
  const float iou = __fdiv_rn(inter, __fadd_rn(uni, 1e-6f));
  const float hw = clamp0(__fsub_rn(nan_max(px1, tx1), nan_min(px0, tx0)));
  const float hh = clamp0(__fsub_rn(nan_max(py1, ty1), nan_min(py0, ty0)));
  const float hull = __fmul_rn(hw, hh);
  giou = __fsub_rn(iou, __fdiv_rn(__fsub_rn(hull, uni), __fadd_rn(hull, 1e-6f)));
}

// Focal class-cost terms of one logit (matcher.py:68,74-81), fp32 like the reference.
__device__ __forceinline__ void focal_terms(float x, float alpha, float one_m_alpha, float &pos, float &neg) {
  const float prob = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x)));
  const float one_m = __fsub_rn(1.f, prob);
  neg = __fmul_rn(__fmul_rn(one_m_alpha, __fmul_rn(prob, prob)), -logf(__fadd_rn(one_m, 1e-8f)));
  pos = __fmul_rn(__fmul_rn(alpha, __fmul_rn(one_m, one_m)), -logf(__fadd_rn(prob, 1e-8f)));
}

__device__ __forceinline__ float combine(float l1, float cls, float giou, float w_class, float w_bbox, float w_giou) {
  float c = __fadd_rn(__fadd_rn(__fmul_rn(w_bbox, l1), __fmul_rn(w_class, cls)), __fmul_rn(w_giou, -giou));   // :100
  if (c != c || fabsf(c) == CUDART_INF_F) c = 0.f;                                                             // :102-103
  return c;
}

constexpr int kSparseCap = 16;   // non-zero tokens per positive-map row the sparse path keeps (a phrase is a few tokens)

// KI = ceil(K / 32): tokens per lane.
// Per chunk of 32 targets the CTA normalises the positive-map rows into shared memory and, while doing so, compacts
// their non-zero tokens.  A positive map is a handful of ones per row (create_positive_map marks the tokens of one
// phrase), so when every row of the chunk has at most kSparseCap non-zeros the class cost of a (prediction, target)
// pair is evaluated by ONE thread from just those tokens -- exp/log are spent on ~T*nnz logits per prediction instead of
// all K -- and the only pass over the full logits row is the NaN scan (a NaN logit anywhere in the row makes the
// reference's dense product NaN, hence cost 0, for every target: NaN * 0 = NaN).  Otherwise the chunk takes the dense
// path: one warp per prediction row, focal terms in registers, two lane-strided dot products per target.
template <int KI>
__global__ void __launch_bounds__(kThreads, KI <= 8 ? 4 : 2) cost_kernel(const float *__restrict__ logits, const float *__restrict__ pboxes,
                                                        const float *__restrict__ tboxes, const float *__restrict__ lmap,
                                                        const int4 *__restrict__ problems, float *__restrict__ cost, int Q,
                                                        int K, float w_class, float w_bbox, float w_giou, float alpha,
                                                        float one_m_alpha, int force_dense) {
  extern __shared__ float smem[];          // kChunk x KP normalised positive-map rows (KP = 32 * KI, zero padded)
  constexpr int KP = 32 * KI;
  __shared__ int sp_k[kChunk][kSparseCap];
  __shared__ float sp_m[kChunk][kSparseCap];
  __shared__ int sp_n[kChunk];             // non-zeros of the row; -1: the row holds NaN / inf (e.g. an all-zero map: 0/0)
  __shared__ int row_nan[kRowsPerCta];
  __shared__ int s_dense;
  const int4 pr = problems[blockIdx.y];
  const int t_first = pr.x, T = pr.y;
  if (T <= 0) return;
  float *__restrict__ block = cost + pr.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q0 = blockIdx.x * kRowsPerCta;

  // NaN scan of this CTA's prediction rows (the one full pass over the logits); all loads of a warp's rows are issued
  // before the first vote so that they overlap
  {
    constexpr int RPW = kRowsPerCta / kWarps;
    bool bad[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int q = q0 + warp + r * kWarps;
      bad[r] = false;
      if (q < Q) {
        const float *row = logits + ((long long)blockIdx.y * Q + q) * K;
#pragma unroll
        for (int i = 0; i < KI; ++i) {
          const int k = lane + 32 * i;
          const float x = k < K ? __ldg(row + k) : 0.f;
          bad[r] |= (x != x);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const unsigned any = __ballot_sync(0xffffffffu, bad[r]);
      if (lane == 0) row_nan[warp + r * kWarps] = any != 0u;
    }
  }

  for (int c0 = 0; c0 < T; c0 += kChunk) {
    const int nt = min(kChunk, T - c0);
    __syncthreads();
    if (threadIdx.x == 0) s_dense = force_dense;
    __syncthreads();
    // normalise this chunk's positive-map rows: row / row.sum()   (matcher.py:86-87), and compact their non-zeros
    for (int tt = warp; tt < nt; tt += kWarps) {
      float v[KI];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < KI; ++i) {
        const int k = lane + 32 * i;
        v[i] = (tt < nt && k < K) ? lmap[(long long)(t_first + c0 + tt) * K + k] : 0.f;
        s += v[i];
      }
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
      int count = 0;
      bool weird = false;
#pragma unroll
      for (int i = 0; i < KI; ++i) {
        const int k = lane + 32 * i;
        const float m = (tt < nt && k < K) ? __fdiv_rn(v[i], s) : 0.f;
        smem[tt * KP + k] = m;
        weird |= (m != m) || fabsf(m) == CUDART_INF_F;
        const bool nz = m != 0.f;                       // NaN counts as non-zero, and marks the row below
        const unsigned mask = __ballot_sync(0xffffffffu, nz);
        const int at = count + __popc(mask & ((1u << lane) - 1u));
        if (nz && at < kSparseCap) { sp_k[tt][at] = k; sp_m[tt][at] = m; }
        count += __popc(mask);
      }
      const unsigned any_weird = __ballot_sync(0xffffffffu, weird);
      if (lane == 0) {
        sp_n[tt] = any_weird ? -1 : count;
        if (!any_weird && count > kSparseCap) s_dense = 1;   // benign race: every writer stores 1
      }
    }
    __syncthreads();

    if (!s_dense) {
      // ---- sparse path: one thread per (prediction, target) pair, targets fastest (coalesced stores) ----
      for (int e = threadIdx.x; e < kRowsPerCta * nt; e += kThreads) {
        const int qq = e / nt, tt = e - qq * nt;
        const int q = q0 + qq;
        if (q >= Q) break;
        const long long row = (long long)blockIdx.y * Q + q;
        const int n = sp_n[tt];
        float cls;
        if (n < 0 || row_nan[qq]) {
          cls = CUDART_NAN_F;
        } else {
          float ap = 0.f, an = 0.f;
          for (int z = 0; z < n; ++z) {
            float pos, neg;
            focal_terms(logits[row * K + sp_k[tt][z]], alpha, one_m_alpha, pos, neg);
            const float m = sp_m[tt][z];
            ap = fmaf(pos, m, ap);
            an = fmaf(neg, m, an);
          }
          cls = __fsub_rn(ap, an);                                                 // pos @ m - neg @ m  (:88)
        }
        float l1, giou;
        box_costs(*reinterpret_cast<const float4 *>(pboxes + row * 4),
                  *reinterpret_cast<const float4 *>(tboxes + (long long)(t_first + c0 + tt) * 4), l1, giou);
        block[(long long)q * T + c0 + tt] = combine(l1, cls, giou, w_class, w_bbox, w_giou);
      }
      continue;
    }

    // ---- dense path ----
    for (int qq = warp; qq < kRowsPerCta; qq += kWarps) {
      const int q = q0 + qq;
      if (q >= Q) break;                                   // warp-uniform
      const long long row = (long long)blockIdx.y * Q + q;
      float pos[KI], neg[KI];
#pragma unroll
      for (int i = 0; i < KI; ++i) {
        const int k = lane + 32 * i;
        focal_terms(k < K ? logits[row * K + k] : 0.f, alpha, one_m_alpha, pos[i], neg[i]);
      }
      // one target at a time: two lane-strided dot products, each finished by a butterfly (this path is the rare one;
      // keeping it free of per-target register arrays lets the kernel run 4 CTAs per SM)
      float sp = 0.f, sn = 0.f;
      for (int tt = 0; tt < nt; ++tt) {
        float ap = 0.f, an = 0.f;
#pragma unroll
        for (int i = 0; i < KI; ++i) {
          const float m = smem[tt * KP + lane + 32 * i];
          ap = fmaf(pos[i], m, ap);
          an = fmaf(neg[i], m, an);
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
          ap += __shfl_xor_sync(0xffffffffu, ap, d);
          an += __shfl_xor_sync(0xffffffffu, an, d);
        }
        if (lane == tt) { sp = ap; sn = an; }
      }
      if (lane < nt) {
        const int t = c0 + lane;
        float l1, giou;
        box_costs(*reinterpret_cast<const float4 *>(pboxes + row * 4),
                  *reinterpret_cast<const float4 *>(tboxes + (long long)(t_first + t) * 4), l1, giou);
        block[(long long)q * T + t] = combine(l1, __fsub_rn(sp, sn), giou, w_class, w_bbox, w_giou);
      }
    }
  }
}

// ---- assignment -------------------------------------------------------------------------------------------------
struct Best {
  double val;
  int code;   // unassigned column at array position it: it + 1 (later wins); assigned: -it (earlier wins)
};
__device__ __forceinline__ bool better(const Best &a, const Best &b) {
  return a.val < b.val || (a.val == b.val && a.code > b.code);
}
__device__ __forceinline__ Best shfl_xor(const Best &b, int d) {
  Best r;
  r.val = __shfl_xor_sync(0xffffffffu, b.val, d);
  r.code = __shfl_xor_sync(0xffffffffu, b.code, d);
  return r;
}

__global__ void __launch_bounds__(kThreads) lsap_kernel(const float *__restrict__ cost, const int4 *__restrict__ problems,
                                                        long long *__restrict__ index_i, long long *__restrict__ index_j,
                                                        int *__restrict__ status, int Q, int stage_elems) {
  const int4 pr = problems[blockIdx.x];
  const int T = pr.y;
  const float *C = cost + pr.z;
  long long *__restrict__ out_i = index_i + pr.w;
  long long *__restrict__ out_j = index_j + pr.w;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (T <= 0 || Q <= 0) {
    if (tid == 0) status[blockIdx.x] = 0;
    return;
  }
  // scipy transposes when there are more rows than columns, so that rows <= columns
  const bool transposed = T < Q;
  const int nr = transposed ? T : Q, nc = transposed ? Q : T;

  extern __shared__ double smem_d[];
  double *spc = smem_d;                      // shortest path cost per column
  double *v = spc + nc;                      // column duals
  double *u = v + nc;                        // row duals (nr)
  int *path = reinterpret_cast<int *>(u + nr);
  int *row4col = path + nc;
  int *remaining = row4col + nc;
  int *col4row = remaining + nc;             // nr
  float *staged = reinterpret_cast<float *>(col4row + nr);   // the problem's cost block, when the host sized room for it
  __shared__ Best warp_best[kWarps];
  __shared__ int s_bad, s_i, s_sink, s_nrem;
  __shared__ double s_min;

  // scipy rejects NaN and -inf entries before solving
  if (tid == 0) s_bad = 0;
  for (int j = tid; j < nc; j += kThreads) { v[j] = 0.0; row4col[j] = -1; path[j] = -1; }
  for (int r = tid; r < nr; r += kThreads) { u[r] = 0.0; col4row[r] = -1; }
  __syncthreads();
  {
    // one pass over the block: the validity scan, and -- when it fits -- the copy into shared memory that takes the
    // L2 round trip out of every Dijkstra step
    const bool stage = (long long)Q * T <= (long long)stage_elems;
    int bad = 0;
    for (long long k = tid; k < (long long)Q * T; k += kThreads) {
      const float c = C[k];
      bad |= (c != c) || (c == -CUDART_INF_F);
      if (stage) staged[k] = c;
    }
    if (bad) s_bad = 1;
    if (stage) C = staged;
  }
  __syncthreads();
  int rc = s_bad ? -2 : 0;

  for (int cur = 0; cur < nr && rc == 0; ++cur) {
    // ---- shortest augmenting path from row `cur` ----
    for (int it = tid; it < nc; it += kThreads) { remaining[it] = nc - it - 1; spc[nc - it - 1] = CUDART_INF; }
    if (tid == 0) { s_i = cur; s_sink = -1; s_nrem = nc; s_min = 0.0; }
    __syncthreads();
    for (;;) {
      const int i = s_i, n_rem = s_nrem;
      const double min_val = s_min, u_i = u[i];
      Best best;
      best.val = CUDART_INF;
      best.code = INT_MIN;
      for (int it = tid; it < n_rem; it += kThreads) {
        const int j = remaining[it];
        const double c = (double)(transposed ? C[(long long)j * T + i] : C[(long long)i * T + j]);
        const double r = __dsub_rn(__dsub_rn(__dadd_rn(min_val, c), u_i), v[j]);
        double s = spc[j];
        if (r < s) { path[j] = i; spc[j] = r; s = r; }
        Best cand;
        cand.val = s;
        cand.code = row4col[j] == -1 ? it + 1 : -it;
        if (better(cand, best)) best = cand;
      }
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) {
        const Best o = shfl_xor(best, d);
        if (better(o, best)) best = o;
      }
      if (lane == 0) warp_best[warp] = best;
      __syncthreads();
      if (warp == 0) {
        Best b = lane < kWarps ? warp_best[lane] : best;
        if (lane >= kWarps) { b.val = CUDART_INF; b.code = INT_MIN; }
#pragma unroll
        for (int d = kWarps / 2; d >= 1; d >>= 1) {
          const Best o = shfl_xor(b, d);
          if (better(o, b)) b = o;
        }
        if (lane == 0) {
          s_min = b.val;
          if (b.val == CUDART_INF) {
            s_sink = -2;                               // infeasible
          } else {
            const int index = b.code > 0 ? b.code - 1 : -b.code;
            const int j = remaining[index];
            if (row4col[j] == -1) s_sink = j; else s_i = row4col[j];
            // compact; the freed tail slot keeps the scanned column so that slots [n_rem, nc) list the set SC
            remaining[index] = remaining[n_rem - 1];
            remaining[n_rem - 1] = j;
            s_nrem = n_rem - 1;
          }
        }
      }
      __syncthreads();
      if (s_sink != -1) break;
    }
    const int sink = s_sink;
    if (sink == -2) { rc = -1; break; }
    // ---- dual update: every scanned column j (and the row matched to it) moves by min_val - spc[j] ----
    const double min_val = s_min;
    const int n_rem = s_nrem;
    if (tid == 0) u[cur] = __dadd_rn(u[cur], min_val);
    for (int it = n_rem + tid; it < nc; it += kThreads) {
      const int j = remaining[it];
      const double delta = __dsub_rn(min_val, spc[j]);
      const int r = row4col[j];
      if (r != -1) u[r] = __dadd_rn(u[r], delta);     // r != cur: cur is unassigned until the augmentation below
      v[j] = __dsub_rn(v[j], delta);
    }
    __syncthreads();
    // ---- augment along the path ----
    if (tid == 0) {
      int j = sink;
      for (;;) {
        const int r = path[j];
        row4col[j] = r;
        const int t = col4row[r];
        col4row[r] = j;
        j = t;
        if (r == cur) break;
      }
    }
    __syncthreads();
  }

  if (tid == 0) status[blockIdx.x] = rc;
  if (rc != 0) {
    for (int r = tid; r < nr; r += kThreads) { out_i[r] = -1; out_j[r] = -1; }
    return;
  }
  if (transposed) {
    // rows of the caller are our columns: report the pairs by ascending prediction index
    for (int r = tid; r < nr; r += kThreads) {
      const int q = col4row[r];
      int rank = 0;
      for (int r2 = 0; r2 < nr; ++r2) rank += col4row[r2] < q;
      out_i[rank] = q;
      out_j[rank] = r;
    }
  } else {
    for (int r = tid; r < nr; r += kThreads) { out_i[r] = r; out_j[r] = col4row[r]; }
  }
}

static size_t lsap_state_bytes(int Q, int Tmax) {
  const size_t nc = (size_t)(Q > Tmax ? Q : Tmax), nr = (size_t)(Q > Tmax ? Q : Tmax);   // either side may be the rows
  return nc * (8 + 8 + 4 + 4 + 4) + nr * (8 + 4) + 16;
}
constexpr size_t kLsapSmemBudget = 200 * 1024;

}  // namespace hm

using namespace hm;

extern "C" {

int hm_max_columns(void) { return kMaxColumns; }

int hm_cost_matrix(const float *pred_logits, const float *pred_boxes, const float *tgt_boxes, const float *tgt_label_map,
                   const int32_t *problems, float *cost, int P, int Q, int K, float w_class, float w_bbox, float w_giou,
                   float alpha, void *stream) {
  if (P < 0 || Q <= 0 || K <= 0) return MSDA_E_INVALID;
  if (P == 0) return MSDA_OK;
  if (!pred_logits || !pred_boxes || !problems || !cost) return MSDA_E_INVALID;
  if (K > 1024) return MSDA_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(pred_boxes) & 15u) || (tgt_boxes && (reinterpret_cast<uintptr_t>(tgt_boxes) & 15u)))
    return MSDA_E_INVALID;
  if (P > 65535) return MSDA_E_TOO_LARGE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid((Q + kRowsPerCta - 1) / kRowsPerCta, P);
  const int4 *pr = reinterpret_cast<const int4 *>(problems);
  const int force_dense = getenv("HM_COST_DENSE") != nullptr;   // A/B switch for tests and profiling
  const float one_m_alpha = (float)(1.0 - (double)alpha);   // Python evaluates `1 - alpha` in double (matcher.py:80)
  if (K <= 256) {
    cost_kernel<8><<<grid, kThreads, kChunk * 256 * sizeof(float), st>>>(pred_logits, pred_boxes, tgt_boxes, tgt_label_map,
                                                                          pr, cost, Q, K, w_class, w_bbox, w_giou, alpha, one_m_alpha, force_dense);
  } else {
    const size_t bytes = kChunk * 1024 * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(cost_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    cost_kernel<32><<<grid, kThreads, bytes, st>>>(pred_logits, pred_boxes, tgt_boxes, tgt_label_map, pr, cost, Q, K,
                                                   w_class, w_bbox, w_giou, alpha, one_m_alpha, force_dense);
  }
  return (int)cudaGetLastError();
}

int hm_assign(const float *cost, const int32_t *problems, int64_t *index_i, int64_t *index_j, int32_t *status, int P, int Q,
              int max_targets, void *stream) {
  if (P < 0 || Q <= 0 || max_targets < 0) return MSDA_E_INVALID;
  if (P == 0) return MSDA_OK;
  if (!problems || !status || (max_targets > 0 && (!cost || !index_i || !index_j))) return MSDA_E_INVALID;
  if (Q > kMaxColumns || max_targets > kMaxColumns) return MSDA_E_TOO_LARGE;
  const size_t state = lsap_state_bytes(Q, max_targets);
  // room for the largest cost block of the call next to the solver state, if the CTA's shared memory allows
  const size_t block = (size_t)Q * (size_t)max_targets * sizeof(float);
  const int stage_elems = state + block <= kLsapSmemBudget ? Q * max_targets : 0;
  const size_t bytes = state + (stage_elems ? block : 0);
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(lsap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
  }
  lsap_kernel<<<P, kThreads, bytes, static_cast<cudaStream_t>(stream)>>>(
      cost, reinterpret_cast<const int4 *>(problems), reinterpret_cast<long long *>(index_i),
      reinterpret_cast<long long *>(index_j), status, Q, stage_elems);
  return (int)cudaGetLastError();
}

int hm_match(const float *pred_logits, const float *pred_boxes, const float *tgt_boxes, const float *tgt_label_map,
             const int32_t *problems, float *cost, int64_t *index_i, int64_t *index_j, int32_t *status, int P, int Q, int K,
             int max_targets, float w_class, float w_bbox, float w_giou, float alpha, void *stream) {
  int rc = MSDA_OK;
  if (max_targets > 0 && (!tgt_boxes || !tgt_label_map)) return MSDA_E_INVALID;   // some problem has targets
  if (max_targets > 0)
    rc = hm_cost_matrix(pred_logits, pred_boxes, tgt_boxes, tgt_label_map, problems, cost, P, Q, K, w_class, w_bbox, w_giou,
                        alpha, stream);
  if (rc != MSDA_OK) return rc;
  return hm_assign(cost, problems, index_i, index_j, status, P, Q, max_targets, stream);
}

}  // extern "C"
