/* sac_oracle.h -- CPU restatement of the reference's SAC continuous update (TEST INFRASTRUCTURE, ORACLE ONLY).
 *   nets    BaseClass/BaseCNN.py:459-483 (PolicyNetContinuous_SAC), :486-500 (QValueNetContinuous_SAC)
 *   update  Trainer/SAC_Trainer.py:325-379, calc_target :122-131, soft_update :145-147, get_action :444-448
 * Reference quirks kept: critics output `action_dim` (=2) values and the TD target / losses are [B,2]-shaped;
 * the tanh correction applies tanh twice (`log(1 - tanh(action)^2 + 1e-7)` with action already tanh'ed, :481);
 * log_alpha starts at ln 0.01.  Pinned against tests/golden/sac_golden.npz (reference run with injected noise). */
#ifndef SAC_ORACLE_H
#define SAC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t obs_dim, hidden, act_dim;     /* 100, 64, 2 */
    float action_bound;
    float actor_lr, critic_lr, alpha_lr, target_entropy, gamma, tau;
} ora_sac_cfg;

typedef struct {
    float *actor, *c1, *c2, *t1, *t2;      /* flat state_dict-ordered parameters */
    float *actor_m, *actor_v, *c1_m, *c1_v, *c2_m, *c2_v;   /* Adam moments */
    float log_alpha, la_m, la_v;
    int64_t step;                           /* optimiser steps taken (all four optimisers advance together) */
} ora_sac_state;

int64_t ora_sac_actor_params(const ora_sac_cfg *c);
int64_t ora_sac_critic_params(const ora_sac_cfg *c);

/* actor(state) with injected noise: action [B][act], log_prob [B][act] */
void ora_sac_actor_forward(const ora_sac_cfg *c, const float *actor, const float *s, const float *eps, int32_t B,
                           float *action, float *log_prob);
void ora_sac_critic_forward(const ora_sac_cfg *c, const float *critic, const float *s, const float *a, int32_t B, float *q);

/* one SAC_Trainer.update (continuous): returns the actor loss; critic losses via out pointers (may be NULL) */
float ora_sac_update(const ora_sac_cfg *c, ora_sac_state *st, const float *s, const float *a, const float *r,
                     const float *s2, const float *d, const float *eps_next, const float *eps_cur, int32_t B,
                     float *critic1_loss, float *critic2_loss, float *alpha_loss);
#ifdef __cplusplus
}
#endif
#endif
