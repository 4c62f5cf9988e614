/*
 * hungarian_b200.h -- C ABI of the on-device Hungarian matcher in libmsda_b200.so (NVIDIA B200, sm_100a).
 * Scope row f2 of SURVEY.md section 8: the matcher is the largest host synchronisation of the reference's training
 * step (7 output sets x batch images, each a device->host copy of the cost matrix plus scipy on the CPU).
 *
 * What each entry point replaces (paths under /root/reference/models/GroundingDINO/):
 *
 *   hm_cost_matrix   <- HungarianMatcher.forward, cost part      matcher.py:66-103
 *                       (token-sigmoid focal class cost :74-91, torch.cdist(p=1) :82, -generalized_box_iou :98 with
 *                        /root/reference/util/box_ops.py:9-14,25-62, weighted sum :100, NaN/inf -> 0 :102-103)
 *   hm_assign        <- `C.cpu()` + scipy.optimize.linear_sum_assignment per image   matcher.py:101,104-107
 *                       (scipy is not vendored by the reference; the solver follows scipy's rectangular
 *                        shortest-augmenting-path algorithm including its tie-breaking, see oracle/lsap_oracle.c)
 *   hm_match         <- the whole forward for a batch of problems (both steps, two kernel launches)
 *
 * A "problem" is one (output set, image) pair: Q predictions against the T_p targets of that image.  All problems
 * of a training step (main output, auxiliary decoder outputs, interm output: groundingdino.py:548-554,587-593,
 * 613-619) can go into ONE call.
 *
 * Conventions (same as msda_b200.h): plain pointers and ints, device pointers, contiguous row-major tensors, work is
 * only enqueued on `stream` (a cudaStream_t passed as void*), nothing is allocated, no host synchronisation, no global
 * state; return 0, a negative MSDA_E_* code (msda_b200.h) or a positive cudaError_t.
 *
 *   pred_logits    (P, Q, K) fp32   classification logits over the K text tokens
 *   pred_boxes     (P, Q, 4) fp32   cx, cy, w, h
 *   tgt_boxes      (Ttot, 4) fp32   all target boxes of the call, problem after problem may share rows
 *   tgt_label_map  (Ttot, K) fp32   row t = label_map[tgt_ids[t]] (matcher.py:78), NOT normalised (the kernel divides
 *                                   by the row sum like :87)
 *   problems       (P, 4) int32     per problem: { first target row, T_p, element offset of its block in `cost`,
 *                                   offset of its pairs in index_i / index_j }
 *   cost           sum_p Q*T_p fp32 block p is the (Q, T_p) row-major matrix the reference calls c[i] (:105)
 *   index_i/_j     sum_p min(Q,T_p) int64: predictions (ascending) and their targets, scipy's (row_ind, col_ind)
 *   status         (P) int32        0 = solved, -1 = infeasible (all remaining costs +inf), -2 = NaN or -inf entry
 *                                   (scipy raises ValueError for both; pairs are then filled with -1)
 *
 * Environment: HM_COST_DENSE, if set, forces the warp-per-row dense path of the cost kernel -- an A/B switch for tests
 * and profiling; results agree with the default sparse path to 2e-6 of the largest cost.
 */
#ifndef HUNGARIAN_B200_H_
#define HUNGARIAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* largest max(Q, T_p) hm_assign accepts (solver state lives in one CTA's shared memory) */
int hm_max_columns(void);

int hm_cost_matrix(const float *pred_logits, const float *pred_boxes, const float *tgt_boxes,
                   const float *tgt_label_map, const int32_t *problems, float *cost, int num_problems,
                   int num_queries, int num_tokens, float cost_class, float cost_bbox, float cost_giou,
                   float focal_alpha, void *stream);

/* max_targets: an upper bound of every T_p (sizes the CTA's shared memory; the host knows it from the target lists) */
int hm_assign(const float *cost, const int32_t *problems, int64_t *index_i, int64_t *index_j, int32_t *status,
              int num_problems, int num_queries, int max_targets, void *stream);

int hm_match(const float *pred_logits, const float *pred_boxes, const float *tgt_boxes, const float *tgt_label_map,
             const int32_t *problems, float *cost, int64_t *index_i, int64_t *index_j, int32_t *status,
             int num_problems, int num_queries, int num_tokens, int max_targets, float cost_class, float cost_bbox,
             float cost_giou, float focal_alpha, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HUNGARIAN_B200_H_ */
