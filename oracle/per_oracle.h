/* per_oracle.h -- CPU restatement of the reference's prioritised replay (TEST INFRASTRUCTURE ONLY, see uav_oracle.h).
 * Follows BaseClass/replay_buffer.py:57-223: SumTree (array heap of 2*capacity-1 nodes, sequential change propagation)
 * and ReplayTree (alpha 0.6, beta 0.4 -> 1 by 0.001 per sampling call, epsilon 0.01, error clip 1). */
#ifndef PER_ORACLE_H
#define PER_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct ora_per ora_per;
ora_per *ora_per_create(int32_t capacity);
void ora_per_destroy(ora_per *p);
/* ReplayTree.push (:152-154): priority (|err| + eps)^alpha in float32 (the error arrives as a float32 tensor) */
void ora_per_push(ora_per *p, float abs_err);
/* SumTree.add with an explicit priority (tests that need exactly representable sums) */
void ora_per_add(ora_per *p, double priority);
/* ReplayTree.sample2 (:186-213) with the uniform draws supplied: s_i = a_i + (b_i - a_i) * u[i].
 * tree_idx[B] (leaf index in the heap array), weights[B] (importance weights / max), returns beta after the call. */
double ora_per_sample(ora_per *p, int32_t B, const double *u, int64_t *tree_idx, double *weights);
/* ReplayTree.batch_update (:216-223): sequential leaf updates, float32 arithmetic for the new priorities */
void ora_per_batch_update(ora_per *p, int32_t B, const int64_t *tree_idx, const float *abs_err);
void ora_per_leaves(const ora_per *p, double *out_capacity);
double ora_per_total(const ora_per *p);
int32_t ora_per_n_entries(const ora_per *p);
#ifdef __cplusplus
}
#endif
#endif
