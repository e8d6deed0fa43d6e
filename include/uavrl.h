/* uavrl.h -- C ABI of the B200-native UAV path-planning hot path (libuavrl_b200.so).
 *
 * The reference (young-how/DQN-based-UAV-3D_path_planer, "RLGF") is pure Python with no FFI; its
 * hot path sits behind a duck-typed plug-in API resolved by name from XML.
 * This header is the boundary a reference-side plug-in binds with ctypes: plain
 * pointers and sizes only, no torch types.  Each entry point cites the reference interface it
 * replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - every function returns 0 on success or a negative uavrl_status; uavrl_last_error() gives the
 *     thread-local message.  Nothing falls back to a CPU path: without a CUDA device create() fails.
 *   - "_dev" pointers are device memory owned by the CALLER (e.g. torch.Tensor.data_ptr());
 *     "_host" pointers are host memory.  The library owns only its handles and their internal state.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  All work is
 *     enqueued on it; there are no hidden synchronisations except where a _host output is written.
 *   - one host thread per handle.
 */
#ifndef UAVRL_H
#define UAVRL_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UAVRL_OBS_DIM 100            /* Agents/UAV.py:517 state_map (1,1,1,100) */
#define UAVRL_MAX_HIDDEN 4

typedef enum {
    UAVRL_OK = 0,
    UAVRL_ERR_INVALID = -1,          /* bad argument / configuration */
    UAVRL_ERR_CUDA = -2,             /* CUDA runtime error (no device, launch failure, ...) */
    UAVRL_ERR_STATE = -3,            /* call order (e.g. step before reset) */
    UAVRL_ERR_NOMEM = -4
} uavrl_status;

typedef enum { UAVRL_INFO_NORMAL = 0, UAVRL_INFO_SUCCESS = 1, UAVRL_INFO_LOSE = 2 } uavrl_info;

/* action encodings accepted by uavrl_env_step */
typedef enum {
    UAVRL_ACT_CONT_F32 = 0,          /* steering fraction a0 = action[0] in [-1,1]  (UAV.py:407,414) */
    UAVRL_ACT_CONT_F64 = 1,          /* same, double (exact replay of reference tapes) */
    UAVRL_ACT_DISCRETE27 = 2,        /* int32 k in 0..26: documented extension, see DESIGN.md */
    UAVRL_ACT_CONT_F32X2 = 3         /* float [n][2] as SAC's get_action returns it; only action[0] steers (UAV.py:414) */
} uavrl_action_kind;

typedef enum { UAVRL_ALGO_DQN = 0, UAVRL_ALGO_DDQN = 1, UAVRL_ALGO_DUELING = 2 } uavrl_algo;

/* ------------------------------------------------------------------ environment batch */
typedef struct uavrl_env uavrl_env;

typedef struct {
    int32_t n_envs;                  /* UAV instances stepped in lockstep */
    int32_t max_subgoals;            /* K: capacity of each sub-goal queue (RRT path length bound) */
    double len, width, h;            /* BaseClass/BaseEnv.py:19-21 (config/PathPlan_City.xml:4-6) */
    double max_v, min_v;             /* config/UAV.xml:12-13 (Agents/UAV.py:25); 0 < max_v < 7 (the sub-goal radius: see uavrl_env_create) */
    double steering_angle;           /* radians: Steering_angle/180*pi (UAV.py:26) */
    int32_t max_step;                /* UAV.py:32 */
    double climb_rate;               /* discrete-27 extension only */
    int32_t n_buildings;
    const double *buildings_host;    /* [n][5] = cx, cy, cz, _R, _H (Obstacles/building.py:8-11) */
    int32_t device;                  /* CUDA device ordinal */
    int32_t auto_reset;              /* 1: an env whose episode ended (UAV.done) restarts from the
                                        scenario pool inside the same step call */
} uavrl_env_config;

int uavrl_env_create(const uavrl_env_config *cfg, uavrl_env **out);
int uavrl_env_destroy(uavrl_env *env);

/* Scenario pool = pre-generated outcomes of UAV.reset() (UAV.py:335-366): start ~U(10,210)x U(1,10),
 * goal ~U(330,490)x U(420,490), heading ~U(0,2pi), sub-goal queue from RRT (PathPlan/RRT.py:63-105;
 * queue[0] is the start point -- the very Loc object of the UAV, `alias0`=1 -- and the last entry
 * is the goal).  Host arrays: start[P][3], goal[P][3], heading[P], subgoals[P][K][3], n_sub[P],
 * alias0[P] (NULL = all 1). */
int uavrl_env_set_pool(uavrl_env *env, int32_t n_scenarios, const double *start_host,
                       const double *goal_host, const double *heading_host,
                       const double *subgoals_host, const int32_t *n_sub_host,
                       const uint8_t *alias0_host);

/* UAV.reset() for every env: env e takes scenario (first_scenario + e) mod P of the pool. */
int uavrl_env_reset(uavrl_env *env, int32_t first_scenario, void *stream);

/* Host-side scenario generator (reset draws + RRT) for synthetic pools: statistical restatement of
 * UAV.py:344-360 + RRT.py:26-105 with a counter-based RNG (the reference's Python MT19937 stream is
 * not reproduced).  Writes the arrays uavrl_env_set_pool takes. */
int uavrl_make_scenarios(const uavrl_env_config *cfg, uint64_t seed, int32_t n_scenarios,
                         int32_t rrt_step, double *start_host, double *goal_host,
                         double *heading_host, double *subgoals_host, int32_t *n_sub_host);

/* The same generator as a device kernel (one thread per scenario): fills the env's device pool
 * directly -- replaces uavrl_make_scenarios + uavrl_env_set_pool, bit-identical scenarios for the same
 * (seed, index); replaces UAV.reset (UAV.py:335-366) + RRTPlanner.getPath (PathPlan/RRT.py:63-105) for the pool.
 * Synchronises `stream`.  uavrl_env_get_pool copies the current pool back (any pointer may be NULL):
 * start[P][3], goal[P][3], v0[P][3] = (Vx, Vy, |V|), subgoals[P][K][3], n_sub[P]. */
int uavrl_env_generate_pool(uavrl_env *env, int32_t n_scenarios, uint64_t seed, int32_t rrt_step, void *stream);
int uavrl_env_get_pool(uavrl_env *env, double *start_host, double *goal_host, double *v0_host,
                       double *subgoals_host, int32_t *n_sub_host);

/* UAV.state() -> state_PathPlan (UAV.py:515-567): obs_dev [n_envs][100] float32. */
int uavrl_env_observe(uavrl_env *env, float *obs_dev, void *stream);

/* BaseEnv.Move_Agent (BaseEnv.py:123-137) = UAV.update_PathPlan (UAV.py:397-513) followed by
 * UAV.state(), for all envs.  Outputs (any may be NULL): next_obs [n][100] f32, reward [n] f32,
 * done [n] u8 (the RETURNED flag, True at every sub-goal), info [n] u8 (uavrl_info),
 * collision [n] u8 (predicate at UAV.py:425), ended [n] u8 (UAV.done after the step, before any
 * auto-reset). */
int uavrl_env_step(uavrl_env *env, int32_t action_kind, const void *actions_dev,
                   float *next_obs_dev, float *reward_dev, uint8_t *done_dev, uint8_t *info_dev,
                   uint8_t *collision_dev, uint8_t *ended_dev, void *stream);

/* Same call with HOST buffers: copies actions in, runs the step, copies results out, synchronises.
 * This is the call a reference-side plug-in makes per lockstep iteration. */
int uavrl_env_step_host(uavrl_env *env, int32_t action_kind, const void *actions_host,
                        float *next_obs_host, float *reward_host, uint8_t *done_host,
                        uint8_t *info_host, uint8_t *collision_host, uint8_t *ended_host);

/* fp64 state read-back for parity tests / checkpoints; any pointer may be NULL.  Each [n_envs]. */
typedef struct {
    double *px, *py, *pz, *vx, *vy, *V, *score, *total_score, *path_len, *reward64;
    int32_t *step, *cursor, *scenario;
    uint8_t *done;
} uavrl_env_state_host;
int uavrl_env_get_state(uavrl_env *env, const uavrl_env_state_host *out);
/* Overwrite per-UAV state from host arrays (NULL members are left alone): position, V_vector / V, Step, score,
 * total_score, path_len, done -- the attributes UAV.reset / update_PathPlan maintain (Agents/UAV.py:335-366, 397-513);
 * resume from a saved state, or start a step from a constructed one.  The sub-goal queue, goal and cursor belong to the
 * scenario (uavrl_env_set_pool + uavrl_env_reset): cursor / scenario / reward64 must be NULL.  Synchronises the device. */
int uavrl_env_set_state(uavrl_env *env, const uavrl_env_state_host *in);

/* Optional models of the UAV (off by default; set after uavrl_env_create, before uavrl_env_reset):
 *   energy   UAV.Calc_Fly_Power (Agents/UAV.py:239-245) with the constants of config/UAV.xml <Power_param><Fly_power>
 *            (xi = 0.8 + 0.02 j, UAV.py:58): P(V) = P_i sqrt(sqrt(1 + V^4/(4 v_0^4)) - V^2/(2 v_0^2)) + d_0 rho s A V^3 / 2
 *            + xi P_b (1 + 3 V^2 / F_b^2), accumulated per step into a per-UAV energy column (the reference evaluates the
 *            formula but never accumulates it, UAV.py:59,93 -- the accumulator is the north-star's "energy model" output).
 *   apf      moving-obstacle artificial potential field (UAV.cal_force / Adjust_subgoal, UAV.py:156-210, and the reward
 *            term :448-453): obstacle_v_host [n_buildings][3] = the `v` attribute of each obstacle; obstacles with v = 0
 *            exert no force (UAV.py:180-182).  Every step shifts each remaining sub-goal by the force at its position,
 *            so sub-goal queues become per-UAV state.
 *   track    UAV.path (UAV.py:432) of the first track_envs UAVs, up to track_capacity points per episode, double
 *            buffered: the episode in progress and the last finished one (what path.csv holds, UAV.py:461-464). */
typedef struct {
    int32_t energy_enabled;
    double P_i, v_0, d_0, rho, s, A, P_b, F_b, xi;
    int32_t apf_enabled;
    const double *obstacle_v_host;
    int32_t track_envs, track_capacity;
} uavrl_env_extras;
int uavrl_env_set_extras(uavrl_env *env, const uavrl_env_extras *extras);
/* energy_host [n_envs]: sum of Calc_Fly_Power(V) over the steps of the episode in progress (joules per unit step time) */
int uavrl_env_get_energy(uavrl_env *env, double *energy_host);
/* flight energy of all UAVs over all steps since uavrl_env_create (the sum of UAV.energy_cost_total over the batch) */
int uavrl_env_get_energy_total(uavrl_env *env, double *total_out);
/* which = 0: episode in progress, 1: last finished episode.  xyz_host [capacity][3]; *n_out = points recorded */
int uavrl_env_get_path(uavrl_env *env, int32_t e, int32_t which, int32_t capacity, double *xyz_host, int32_t *n_out);
/* the per-UAV sub-goal queues [n_envs][K][3] (the scenario's queue, shifted by APF when enabled) */
int uavrl_env_get_subgoals(uavrl_env *env, double *sub_host);

/* PathPlan_City.Threaten_rate (Envs/PathPlan_City.py:215-223) on arbitrary points (device kernel):
 * pts_host [n][3] -> out_host [n] u8. */
int uavrl_env_threaten_rate(uavrl_env *env, int32_t n, const double *pts_host, uint8_t *out_host);

/* ------------------------------------------------------------------ learner (Q-net + replay) */
typedef struct uavrl_learner uavrl_learner;

typedef struct {
    int32_t in_dim;                       /* w (config/Trainer.xml <w>) */
    int32_t n_hidden;                     /* trunk layers with ReLU: Qnet2/VAnet2 = 1, QValueNet_SAC/VAnet3 = 2 ... */
    int32_t hidden[UAVRL_MAX_HIDDEN];     /* e.g. {64}, {64,64}, {128,64} (BaseClass/BaseCNN.py) */
    int32_t n_actions;                    /* <output> */
    int32_t dueling;                      /* 1: fc_A + fc_V heads, Q = V + A - mean(A) (BaseCNN.py:131-139) */
    int32_t algo;                         /* uavrl_algo */
    float lr;                             /* LEARNING_RATE (BaseTrainer.py:33) */
    float gamma;                          /* BaseTrainer.py:35 */
    int32_t batch_size;                   /* Batch_Size (BaseTrainer.py:34) */
    int32_t update_loop;                  /* hard target update period (DuelingDQN_Trainer.py:30,183) */
    int64_t replay_capacity;              /* replay_size, in transitions (BaseTrainer.py:32,39) */
    int32_t lockstep_envs;                /* >0: frame-ring replay fed by uavrl_train_* (N envs per frame);
                                             0: generic transition store fed by uavrl_replay_push */
    uint64_t seed;                        /* Philox key for eps-greedy and replay sampling */
    int32_t device;
    int32_t loss_kind;                    /* 0 = MSE, what every reference trainer uses (BaseTrainer.py:40, DQN_Trainer.py:119);
                                             1 = Huber / torch SmoothL1Loss(beta = 1): 0.5 d^2 for |d| < 1, |d| - 0.5 otherwise
                                             (an option the reference does not have; off for parity) */
} uavrl_learner_config;

int uavrl_learner_create(const uavrl_learner_config *cfg, uavrl_learner **out);
int uavrl_learner_destroy(uavrl_learner *l);
int64_t uavrl_learner_param_count(const uavrl_learner *l);

/* state_dict()-ordered flat fp32 parameters (fc1.weight [out][in], fc1.bias, ..., for dueling nets
 * ..., fc_A.weight, fc_A.bias, fc_V.weight, fc_V.bias) -- what torch.save({'model': ...}) holds
 * (DuelingDQN_Trainer.py:79-84).  which: 0 = q_local, 1 = q_target, 2 = Adam exp_avg,
 * 3 = Adam exp_avg_sq, 4 = last gradient. */
int uavrl_learner_set_params(uavrl_learner *l, int32_t which, const float *params_host);
int uavrl_learner_get_params(uavrl_learner *l, int32_t which, float *params_host);
int uavrl_learner_set_counters(uavrl_learner *l, int64_t epoch, int64_t adam_step);
int uavrl_learner_get_counters(uavrl_learner *l, int64_t *epoch, int64_t *adam_step);

/* Trainer.get_action (DuelingDQN_Trainer.py:86-97) for n observations: u > eps (or !is_train)
 * -> argmax_a q_local(obs), else a uniformly random action.  u_tape_dev / rand_tape_dev inject the
 * random draws (parity tests); NULL = Philox.  q_out_dev optional [n][A]. */
int uavrl_learner_act(uavrl_learner *l, const float *obs_dev, int32_t n, float eps, int32_t is_train,
                      const float *u_tape_dev, const int32_t *rand_tape_dev, int32_t *actions_dev,
                      float *q_out_dev, void *stream);

/* ReplayMemory.add (BaseClass/replay_buffer.py:41-42), n transitions, FIFO over replay_capacity. */
int uavrl_replay_push(uavrl_learner *l, int32_t n, const float *obs_dev, const int32_t *actions_dev,
                      const float *reward_dev, const float *next_obs_dev, const uint8_t *done_dev,
                      void *stream);
int64_t uavrl_replay_size(const uavrl_learner *l);
/* Read back n stored transitions by logical index (0 = oldest) into host arrays (checkpointing the
 * replay, tests): s/s2 [n][in], a [n], r [n], d [n]. */
int uavrl_replay_gather(uavrl_learner *l, int32_t n, const int64_t *logical_idx_host, float *s_host,
                        int32_t *a_host, float *r_host, float *s2_host, uint8_t *d_host);

/* Trainer.update (DuelingDQN_Trainer.py:150-190; DQN_Trainer.py:85-136; DDQN_Trainer.py:72-117):
 * epoch += 1; sample Batch_Size distinct transitions uniformly (replay_buffer.py:48-51;
 * idx_tape_dev injects the indices, NULL = Philox), TD target, MSE loss, backward, Adam step, hard
 * target update every update_loop epochs.  Skipped (epoch still counts) while the replay holds
 * <= Batch_Size transitions (PathPlan_City.py:383).  loss_dev optional [1]. */
int uavrl_learner_update(uavrl_learner *l, const int32_t *idx_tape_dev, float *loss_dev, void *stream);

/* The same update on an explicit batch (transition_dict of DuelingDQN_Trainer.update): device arrays
 * s [B][in], a [B], r [B], s2 [B][in], d [B] (float 0/1). */
int uavrl_learner_update_batch(uavrl_learner *l, int32_t B, const float *s_dev, const int32_t *a_dev,
                               const float *r_dev, const float *s2_dev, const float *d_dev,
                               float *loss_dev, void *stream);

/* Split form for data-parallel training: grads only (sum over the local batch of d(loss)/d(theta),
 * loss normalised by global_batch), then -- after the caller all-reduced uavrl_learner_grad_ptr()
 * across ranks -- the optimiser step.  */
int uavrl_learner_compute_grads(uavrl_learner *l, const int32_t *idx_tape_dev, int32_t global_batch,
                                float *loss_dev, void *stream);
float *uavrl_learner_grad_ptr(uavrl_learner *l);          /* device, [param_count] fp32 */
int uavrl_learner_apply_grads(uavrl_learner *l, void *stream);
/* Select the arithmetic path of the Q-network forward passes (get_action, TD target): 1 = tcgen05 tensor
 * cores with the 3xTF32 split (default when the network fits), 0 = fp32 CUDA cores.  Returns the path in use. */
int uavrl_learner_set_tensor_cores(uavrl_learner *l, int32_t enable);
int uavrl_learner_hard_update(uavrl_learner *l, void *stream);   /* DuelingDQN_Trainer.py:199-202 */
/* Trainer.Is_Train (BaseTrainer.py:47) for the lockstep loops below: with 0, get_action is greedy whatever eps is
 * (`sample > eps or not self.Is_Train`, DuelingDQN_Trainer.py:90).  Default 1. */
int uavrl_learner_set_is_train(uavrl_learner *l, int32_t is_train);
/* After an explicit uavrl_env_reset the lockstep ring's current frame no longer matches the env
 * state: call this; the next uavrl_train_run re-observes into a fresh frame and the lockstep replay
 * restarts empty (envs that end episodes restart by themselves with auto_reset and need no call). */
int uavrl_learner_lockstep_restart(uavrl_learner *l);

/* One-shot NVLink all-reduce fused with the optimiser (data-parallel training, one process per GPU):
 *   uavrl_learner_comm_init     allocate this rank's symmetric receive buffer recv[2][world][P+1] + flag words, return their
 *                               CUDA IPC handles (64 bytes each) for exchange (e.g. torch.distributed.all_gather)
 *   uavrl_learner_comm_connect  open every rank's handles (grad_handles / flag_handles: [world][64] bytes)
 *   uavrl_learner_update_dp     epoch += 1; local gradient (loss scaled by 1/global_batch) -> reduced and PUSHED with remote
 *                               stores over NVLink into slot `rank` of every rank's receive buffer -> flags raised on every
 *                               peer -> the optimiser kernel waits for all ranks' flags in local memory, sums the `world`
 *                               vectors of its local receive buffer in rank order (bit-identical replicas) and applies Adam.
 *                               No NCCL call, nothing pulled across NVLink on the critical path.  world = 1 runs the same two
 *                               kernels on the local buffer (self-test on one GPU).
 * loss_dev (optional) receives the GLOBAL batch loss. */
int uavrl_learner_comm_init(uavrl_learner *l, int32_t rank, int32_t world, void *grad_handle_out, void *flag_handle_out);
int uavrl_learner_comm_connect(uavrl_learner *l, const void *grad_handles, const void *flag_handles);
int uavrl_learner_update_dp(uavrl_learner *l, const int32_t *idx_tape_dev, int32_t global_batch, float *loss_dev, void *stream);

/* ------------------------------------------------------------------ fused lockstep training loop
 * PathPlan_City.run_thread_OffPolicy + update (Envs/PathPlan_City.py:364-385,757-776) for all envs:
 *   obs -> get_action -> Move_Agent -> replay add -> [sample -> Trainer.update]
 * n_iters lockstep iterations; observations are written once, straight into the replay frame ring.
 * updates_per_iter optimiser steps follow each env step (reference: 1).  stats_host (optional)
 * receives the counters below. */
typedef struct {
    int64_t env_steps, updates, episodes_ended, collisions;
    int64_t n_success, n_lose;       /* steps whose info was 'success' / 'lose' (PathPlan_City.Run_statistics) */
    double sum_reward;
    float last_loss;
} uavrl_train_stats;
int uavrl_train_run(uavrl_env *env, uavrl_learner *l, int32_t n_iters, float eps,
                    int32_t updates_per_iter, int32_t do_update, uavrl_train_stats *stats_host,
                    void *stream);

/* ------------------------------------------------------------------ SAC, continuous actions (BASELINE config 5)
 * SAC_Trainer (Trainer/SAC_Trainer.py) with PolicyNetContinuous_SAC / QValueNetContinuous_SAC (BaseClass/BaseCNN.py:459-500):
 * actor obs->hidden->{mu, sigma}, twin critics [obs, action]->hidden->hidden->action_dim, their targets, learnable log_alpha
 * (init ln 0.01).  The reference's quirks are kept: [B, action_dim]-shaped critic outputs / TD target / losses and the
 * doubly applied tanh in the log-prob correction. */
typedef struct uavrl_sac uavrl_sac;
typedef struct {
    int32_t obs_dim, hidden, act_dim;     /* <w>, <hiden_dim>, <output>/<action_dim> of config/Trainer.xml: 100, 64, 2 */
    float action_bound;                   /* <action_bound> */
    float actor_lr, critic_lr, alpha_lr;  /* actor.lr, critic.lr, SAC_param.alpha_lr */
    float target_entropy, gamma, tau;     /* SAC_param */
    int32_t batch_size;                   /* Batch_Size */
    int64_t replay_capacity;              /* replay_size */
    int32_t lockstep_envs;                /* > 0: frame-ring replay fed by uavrl_sac_train_run */
    uint64_t seed;
    int32_t device;
} uavrl_sac_config;

int uavrl_sac_create(const uavrl_sac_config *cfg, uavrl_sac **out);
int uavrl_sac_destroy(uavrl_sac *s);
/* role: 0 actor, 1 critic_1, 2 critic_2, 3 target_critic_1, 4 target_critic_2 (flat state_dict order:
 * actor = fc1, fc_mu, fc_std; critic = fc1, fc2, fc_out); 5..7 Adam exp_avg of actor/critic_1/critic_2, 8..10 exp_avg_sq */
int64_t uavrl_sac_param_count(const uavrl_sac *s, int32_t role);
int uavrl_sac_set_params(uavrl_sac *s, int32_t role, const float *params_host);
int uavrl_sac_get_params(uavrl_sac *s, int32_t role, float *params_host);
int uavrl_sac_set_scalars(uavrl_sac *s, float log_alpha, float la_exp_avg, float la_exp_avg_sq, int64_t epoch, int64_t adam_step);
int uavrl_sac_get_scalars(uavrl_sac *s, float *log_alpha, float *la_exp_avg, float *la_exp_avg_sq, int64_t *epoch, int64_t *adam_step);
/* SAC_Trainer.get_action (:444-448): actions_dev [n][2] = tanh(mu + sigma*eps)*bound; eps_dev [n][2] injects the
 * reparameterisation noise (NULL = Philox Box-Muller) */
int uavrl_sac_act(uavrl_sac *s, const float *obs_dev, int32_t n, const float *eps_dev, float *actions_dev, void *stream);
/* SAC_Trainer.update (:317-379, continuous) on an explicit batch: s [B][obs], a [B][2], r [B], s2 [B][obs], d [B];
 * eps_next / eps_cur [B][2] = noise of the two actor evaluations (NULL = Philox); losses_dev (optional) [4] =
 * {actor_loss, critic_1_loss, critic_2_loss, d alpha_loss / d log_alpha}. */
int uavrl_sac_update_batch(uavrl_sac *s, int32_t B, const float *s_dev, const float *a_dev, const float *r_dev, const float *s2_dev,
                           const float *d_dev, const float *eps_next_dev, const float *eps_cur_dev, float *losses_dev, void *stream);
/* PathPlan_City.run_thread_OffPolicy + update with the SAC trainer and the reference's continuous step, N envs in lockstep */
int uavrl_sac_train_run(uavrl_env *env, uavrl_sac *s, int32_t n_iters, int32_t do_update, uavrl_train_stats *stats_host, void *stream);

/* Data-parallel form of uavrl_train_run (after uavrl_learner_comm_connect): every iteration ends with
 * uavrl_learner_update_dp on this rank's replay shard; global_batch = batch_size x world. */
int uavrl_train_run_dp(uavrl_env *env, uavrl_learner *l, int32_t n_iters, float eps, int32_t global_batch, void *stream);

/* The same loop with a CUDA event recorded on `stream` before/after every kernel: ms_out[6] receives the
 * summed device time of {act, env_step, td_target, fwd_bwd, weight_grad, reduce_adam} over the n_iters
 * iterations (bench.py's roofline pass; on the CUDA-core path td_target and weight_grad are 0 because
 * fwd_bwd does everything; event gaps make the loop slower, never use it for throughput). */
int uavrl_train_profile(uavrl_env *env, uavrl_learner *l, int32_t n_iters, float eps, float *ms_out, void *stream);

/* ---- prioritised experience replay ------------------------------------------------------------
 * Replaces SumTree + ReplayTree (BaseClass/replay_buffer.py:57-223): priorities per replay slot, stratified sampling
 * over `batch` equal segments of int(total), importance weights (n p / total)^-beta / max, beta += beta_inc per
 * sampling call (capped at 1), batch_update with (min(|err| + eps, err_upper))^alpha.  Arguments < 0 take the
 * reference's constants (alpha 0.6, beta 0.4, beta_inc 0.001, eps 0.01, err_upper 1).  Enable before the first
 * transition is stored.  Once enabled:
 *   - uavrl_replay_push / the lockstep loops give new transitions the priority of ReplayTree.push with error 0;
 *   - uavrl_learner_update / uavrl_train_run sample through it, minimise mean(w_i (Q - y)^2) and write
 *     |Q - y| back with the batch_update rule.  (The reference multiplies the weights into the already averaged loss,
 *     SAC_Trainer.py:348-352, which cannot be back-propagated; the per-sample form is the documented deviation.)
 * uavrl_per_sample = ReplayTree.sample2 (:186-213): physical slot indices (tree index = slot + capacity - 1) and
 * weights; u_tape_dev (optional, [batch] doubles in [0,1)) replaces the uniform draws.  uavrl_per_set_errors:
 * clip = 0 is ReplayTree.push's rule (:152-154), clip = 1 batch_update's (:216-223).  uavrl_per_set_priorities is
 * SumTree.update with explicit values.  uavrl_per_get copies the leaves [slots] to the host. */
int uavrl_per_enable(uavrl_learner *l, double alpha, double beta0, double beta_inc, double eps, double err_upper);
int uavrl_per_sample(uavrl_learner *l, int32_t batch, const double *u_tape_dev, int32_t *slots_out_dev,
                     float *weights_out_dev, void *stream);
int uavrl_per_set_errors(uavrl_learner *l, int32_t n, const int32_t *slots_dev, const float *abs_err_dev, int32_t clip,
                         void *stream);
int uavrl_per_set_priorities(uavrl_learner *l, int32_t n, const int32_t *slots_dev, const double *priorities_dev,
                             void *stream);
int uavrl_per_get(uavrl_learner *l, double *leaves_host, double *total_out, double *beta_out);
/* Trainer.update(transition_dict) with 'weights' (and |TD error| back for batch_update): uavrl_learner_update_batch
 * with per-sample importance weights in the loss; is_weights_dev / abs_err_out_dev may be NULL. */
int uavrl_learner_update_batch_per(uavrl_learner *l, int32_t batch, const float *obs_dev, const int32_t *act_dev,
                                   const float *rew_dev, const float *next_obs_dev, const float *done_dev,
                                   const float *is_weights_dev, float *abs_err_out_dev, float *loss_dev, void *stream);

const char *uavrl_last_error(void);
const char *uavrl_version(void);
/* number of kernel launches issued by this library in the calling process since load (bench.py) */
int64_t uavrl_launch_count(void);
/* Programmatic dependent launch inside the lockstep loops (each kernel's prologue overlaps its predecessor's
 * tail; results are unchanged).  Process-wide switch, default 1; 0 launches every kernel fully serialised. */
int uavrl_set_pdl(int32_t on);
/* Lockstep loops on the tensor-core path: get_action and Move_Agent as one kernel (each CTA steps the envs whose
 * actions it has just computed; results unchanged).  Process-wide switch, default 0: measured slower than the two
 * kernels chained with programmatic dependent launch (profiles/r01_fused_act_env.txt). */
int uavrl_set_fuse_act_env(int32_t on);
/* Small batches (weight-gradient grid <= number of SMs, one CTA per SM): the weight-gradient kernel writes its partials as 8-byte
 * words {epoch : value}, polls the words of the parameter slice it owns until every slice has delivered (no grid barrier, no fence)
 * and applies the partial reduction + Adam + weight-image refresh itself instead of a separate optimiser launch (results unchanged:
 * the reduction order is the optimiser kernel's).  Process-wide switch, default 0: measured on B200 the loop is 52.8 us per
 * iteration with it and 50.9 us with the PDL-chained pair -- the optimiser's launch latency was already hidden. */
int uavrl_set_fuse_dw_adam(int32_t on);
/* Batches of at most 148 x 32 transitions (148 x 64 on 64-row tiles when the operands fit shared memory) on the tensor-core path: the TD-target forward pass(es) (target network on the next
 * states; double DQN: the local network first) run inside the training kernel, each CTA on the tile it then trains on
 * (weight images restaged in shared memory between the passes, y kept in shared memory) -- one launch instead of two or
 * three per update; the arithmetic is the stand-alone passes'.  Process-wide switch, default 1. */
int uavrl_set_fuse_td(int32_t on);
/* 1 if an update of `batch` transitions on this learner runs the TD-target pass(es) inside the training kernel (see above). */
int uavrl_learner_td_fused(const uavrl_learner *l, int32_t batch);

#ifdef __cplusplus
}
#endif
#endif /* UAVRL_H */
