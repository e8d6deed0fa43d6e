/* dqn_oracle.h -- CPU restatement of the reference's DQN-family learner math (TEST INFRASTRUCTURE).
 *
 * ORACLE ONLY (same rules as uav_oracle.h).  Restates, for MLP Q-networks:
 *   nets      BaseClass/BaseCNN.py:93-102 (Qnet2), :329-343 (QValueNet_SAC), :120-139 (VAnet2),
 *             :142-163 (VAnet3), :166-217 (VAnet4/5)
 *   act       Trainer/DuelingDQN_Trainer.py:86-97 (get_action)
 *   update    Trainer/DQN_Trainer.py:107-126, Trainer/DDQN_Trainer.py:93-107,
 *             Trainer/DuelingDQN_Trainer.py:159-184  (MSE loss, Adam, hard update every Update_loop)
 *   optimiser torch.optim.Adam defaults (betas 0.9/0.999, eps 1e-8, no weight decay / amsgrad)
 * The arithmetic of nn.Linear / autograd / Adam lives in torch (third-party, not in the reference
 * tree): parity is pinned against torch 2.11.0 CPU fp32 executed through the reference's own
 * trainer classes (tests/golden/make_golden.py -> tests/golden/dqn_*.npz).
 */
#ifndef DQN_ORACLE_H
#define DQN_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
void ora_set_loss_kind(int32_t k);   /* 0 MSE (reference), 1 Huber / SmoothL1(beta=1) */

#endif

#define ORA_MAX_HIDDEN 4
enum { ORA_ALGO_DQN = 0, ORA_ALGO_DDQN = 1, ORA_ALGO_DUELING = 2 /* dueling net + double-DQN target */ };

typedef struct {
    int32_t in_dim;                       /* w */
    int32_t n_hidden;                     /* 1..4 trunk layers, ReLU after each */
    int32_t hidden[ORA_MAX_HIDDEN];
    int32_t n_actions;                    /* output */
    int32_t dueling;                      /* 0: plain head; 1: fc_A + fc_V, Q = V + A - mean(A) */
} ora_net;

int64_t ora_net_param_count(const ora_net *n);

/* Q = net(x).  params flat in state_dict order (weight [out][in] row-major, then bias, per layer;
 * dueling: ..., fc_A.weight, fc_A.bias, fc_V.weight, fc_V.bias). */
void ora_net_forward(const ora_net *n, const float *params, const float *x, int32_t B, float *q);

/* eps-greedy action (DuelingDQN_Trainer.py:86-97) on injected random tapes:
 * u[b] > eps (or !is_train) -> argmax_a Q(x_b) (first maximum), else rand_action[b]. */
void ora_act(const ora_net *n, const float *params, const float *x, int32_t B, float eps,
             int32_t is_train, const float *u, const int32_t *rand_action, int32_t *action, float *q_out);

/* One optimiser step.  Updates `local`, `m`, `v` in place; *t is the Adam step counter (incremented).
 * grads_out (optional) receives dLoss/dparams.  Returns the loss (mean squared TD error). */
float ora_dqn_update(const ora_net *n, int32_t algo, float *local, const float *target,
                     float *m, float *v, int64_t *t,
                     const float *s, const int32_t *a, const float *r, const float *s2,
                     const float *d, int32_t B, float gamma, float lr, float *grads_out);

/* prioritised-replay form: per-sample importance weights in the loss, |TD error| out (see dqn_oracle.c) */
float ora_dqn_update_per(const ora_net *n, int32_t algo, float *local, const float *target, float *m, float *v, int64_t *t,
                         const float *s, const int32_t *a, const float *r, const float *s2, const float *d, int32_t B,
                         float gamma, float lr, float *grads_out, const float *is_w, float *abs_err_out);
#ifdef __cplusplus
}
#endif
#endif
