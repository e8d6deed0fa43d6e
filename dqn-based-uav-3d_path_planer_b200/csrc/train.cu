// train.cu -- the fused lockstep training loop over the env batch and the learner.
//
// Replaces PathPlan_City.run_thread_OffPolicy + PathPlan_City.update (Envs/PathPlan_City.py:364-385,
// 757-776) for N envs: state -> get_action -> Move_Agent -> replay add -> sample -> Trainer.update.
// Observations are produced once by the env kernel, directly into the replay frame ring: the
// frame written as "next_obs" of iteration t is the "obs" the policy reads at t+1 (no copy, no
// separate push kernel; 412 B per stored transition instead of 812 B).
#include "env.cuh"
#include "learner.cuh"

#include <vector>

using namespace uavrl;

extern "C" int uavrl_train_run(uavrl_env *env, uavrl_learner *l, int32_t n_iters, float eps, int32_t updates_per_iter,
                               int32_t do_update, uavrl_train_stats *stats_host, void *stream)
{
    if (!env || !l || n_iters < 0) return fail(UAVRL_ERR_INVALID, "bad argument");
    if (l->mode != kReplayLockstep || l->cfg.lockstep_envs != env->d.n)
        return fail(UAVRL_ERR_INVALID, "learner.lockstep_envs must equal env.n_envs");
    if (l->net.in_dim != kObsDim) return fail(UAVRL_ERR_INVALID, "learner.in_dim must be 100 (the UAV observation)");
    if (!env->reset_done) return fail(UAVRL_ERR_STATE, "uavrl_train_run before uavrl_env_reset");
    if (env->cfg.device != l->cfg.device) return fail(UAVRL_ERR_INVALID, "env and learner live on different devices");
    UAVRL_CUDA(cudaSetDevice(env->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    unsigned long long c0[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    double r0 = 0.0;
    if (stats_host) {
        UAVRL_CUDA(cudaStreamSynchronize(st));
        UAVRL_CUDA(cudaMemcpy(c0, env->d.stat_counts, sizeof(c0), cudaMemcpyDeviceToHost));
        UAVRL_CUDA(cudaMemcpy(&r0, env->d.stat_reward, sizeof(r0), cudaMemcpyDeviceToHost));
    }
    int64_t updates = 0;
    l->pdl_chain = true; l->pdl_prev = kPdlNone;          // the first kernel of the loop is launched plainly
    struct ChainOff { uavrl_learner *l; ~ChainOff() { l->pdl_chain = false; l->pdl_prev = kPdlNone; } } chain_off{ l };
    for (int it = 0; it < n_iters; ++it) {
        float *obs_t, *obs_next, *rew; int32_t *act; uint8_t *done;
        lockstep_begin(l, &obs_t, &obs_next, &act, &rew, &done);
        if (!l->frame0_valid) {                          // very first iteration: materialise obs_0
            if ((rc = launch_env_observe(env->d, obs_t, st))) return rc;
            l->frame0_valid = true;
        }
        // Choose_Action2 -> Trainer.get_action (PathPlan_City.py:338-346), then Move_Agent + replay add (:371-382):
        // reward/done land in the ring slots.  One fused kernel when the tensor-core path is on.
        rc = launch_act_env(l, env->d, obs_t, eps, act, obs_next, rew, done, st);
        if (rc < 0) return rc;
        if (rc == 1) {
            if ((rc = launch_act(l, obs_t, env->d.n, eps, l->is_train, nullptr, nullptr, act, nullptr, st))) return rc;
            if ((rc = launch_env_step(env->d, UAVRL_ACT_DISCRETE27, act, obs_next, rew, done, nullptr, nullptr, nullptr, st,
                                      l->pdl_prev == kPdlAct && g_pdl.load()))) return rc;
            l->pdl_prev = kPdlEnv;
        }
        lockstep_commit(l, st);
        if (do_update) {
            for (int u = 0; u < updates_per_iter; ++u) {  // PathPlan_City.update -> Trainer.update (:757-776)
                l->epoch += 1;
                if (l->count <= l->cfg.batch_size) continue;
                BatchSrc src = replay_source(l, nullptr);
                if ((rc = launch_update(l, src, l->cfg.batch_size, l->cfg.batch_size, l->loss_dev, true, st))) return rc;
                ++updates;
            }
        }
    }
    if (stats_host) {
        UAVRL_CUDA(cudaStreamSynchronize(st));
        unsigned long long c1[8]; double r1; float loss;
        UAVRL_CUDA(cudaMemcpy(c1, env->d.stat_counts, sizeof(c1), cudaMemcpyDeviceToHost));
        UAVRL_CUDA(cudaMemcpy(&r1, env->d.stat_reward, sizeof(r1), cudaMemcpyDeviceToHost));
        UAVRL_CUDA(cudaMemcpy(&loss, l->loss_dev, sizeof(loss), cudaMemcpyDeviceToHost));
        stats_host->env_steps = (int64_t)(c1[0] - c0[0]);
        stats_host->episodes_ended = (int64_t)(c1[1] - c0[1]);
        stats_host->collisions = (int64_t)(c1[2] - c0[2]);
        stats_host->n_success = (int64_t)(c1[3] - c0[3]);
        stats_host->n_lose = (int64_t)(c1[4] - c0[4]);
        stats_host->sum_reward = r1 - r0;
        stats_host->updates = updates;
        stats_host->last_loss = loss;
    }
    return 0;
}

extern "C" int uavrl_train_run_dp(uavrl_env *env, uavrl_learner *l, int32_t n_iters, float eps, int32_t global_batch, void *stream)
{
    if (!env || !l || n_iters < 0 || global_batch <= 0) return fail(UAVRL_ERR_INVALID, "bad argument");
    if (l->mode != kReplayLockstep || l->cfg.lockstep_envs != env->d.n)
        return fail(UAVRL_ERR_INVALID, "learner.lockstep_envs must equal env.n_envs");
    if (!l->comm_ready) return fail(UAVRL_ERR_STATE, "uavrl_train_run_dp before uavrl_learner_comm_connect");
    if (!env->reset_done) return fail(UAVRL_ERR_STATE, "uavrl_train_run_dp before uavrl_env_reset");
    UAVRL_CUDA(cudaSetDevice(env->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    l->pdl_chain = true; l->pdl_prev = kPdlNone;
    struct ChainOff { uavrl_learner *l; ~ChainOff() { l->pdl_chain = false; l->pdl_prev = kPdlNone; } } chain_off{ l };
    for (int it = 0; it < n_iters; ++it) {
        float *obs_t, *obs_next, *rew; int32_t *act; uint8_t *done;
        lockstep_begin(l, &obs_t, &obs_next, &act, &rew, &done);
        if (!l->frame0_valid) {
            if ((rc = launch_env_observe(env->d, obs_t, st))) return rc;
            l->frame0_valid = true;
        }
        rc = launch_act_env(l, env->d, obs_t, eps, act, obs_next, rew, done, st);
        if (rc < 0) return rc;
        if (rc == 1) {
            if ((rc = launch_act(l, obs_t, env->d.n, eps, l->is_train, nullptr, nullptr, act, nullptr, st))) return rc;
            if ((rc = launch_env_step(env->d, UAVRL_ACT_DISCRETE27, act, obs_next, rew, done, nullptr, nullptr, nullptr, st,
                                      l->pdl_prev == kPdlAct && g_pdl.load()))) return rc;
            l->pdl_prev = kPdlEnv;
        }
        lockstep_commit(l, st);
        l->epoch += 1;
        // every rank must take part in every all-reduce: the caller warms the replay up first
        if (l->count <= l->cfg.batch_size) return fail(UAVRL_ERR_STATE, "replay holds <= batch_size transitions (warm up with uavrl_train_run first)");
        BatchSrc src = replay_source(l, nullptr);
        if ((rc = launch_update_dp(l, src, l->cfg.batch_size, global_batch, l->loss_dev, st))) return rc;
    }
    return 0;
}

extern "C" int uavrl_train_profile(uavrl_env *env, uavrl_learner *l, int32_t n_iters, float eps, float *ms_out, void *stream)
{
    if (!env || !l || n_iters <= 0 || !ms_out) return fail(UAVRL_ERR_INVALID, "bad argument");
    if (l->mode != kReplayLockstep || l->cfg.lockstep_envs != env->d.n || !env->reset_done || !l->frame0_valid)
        return fail(UAVRL_ERR_STATE, "uavrl_train_profile needs a warmed-up lockstep env/learner pair");
    UAVRL_CUDA(cudaSetDevice(env->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int NE = 7;                      // events per iteration -> 6 intervals
    std::vector<cudaEvent_t> ev((size_t)n_iters * NE);
    for (auto &e : ev) UAVRL_CUDA(cudaEventCreate(&e));
    int rc;
    for (int it = 0; it < n_iters; ++it) {
        float *obs_t, *obs_next, *rew; int32_t *act; uint8_t *done;
        lockstep_begin(l, &obs_t, &obs_next, &act, &rew, &done);
        cudaEvent_t *e = &ev[(size_t)it * NE];
        UAVRL_CUDA(cudaEventRecord(e[0], st));
        rc = launch_act_env(l, env->d, obs_t, eps, act, obs_next, rew, done, st);      // fused: slot 0 = act + step, slot 1 = 0
        if (rc < 0) return rc;
        const bool fused = (rc == 0);
        if (!fused && (rc = launch_act(l, obs_t, env->d.n, eps, l->is_train, nullptr, nullptr, act, nullptr, st))) return rc;
        UAVRL_CUDA(cudaEventRecord(e[1], st));
        if (!fused && (rc = launch_env_step(env->d, UAVRL_ACT_DISCRETE27, act, obs_next, rew, done, nullptr, nullptr, nullptr, st))) return rc;
        UAVRL_CUDA(cudaEventRecord(e[2], st));
        lockstep_commit(l, st);
        l->epoch += 1;
        if (l->count <= l->cfg.batch_size) return fail(UAVRL_ERR_STATE, "replay not warmed up");
        BatchSrc src = replay_source(l, nullptr);
        if ((rc = launch_update_split(l, src, l->cfg.batch_size, st, &e[3]))) return rc;   // records e[3], e[4], e[5]
        UAVRL_CUDA(cudaEventRecord(e[6], st));
    }
    UAVRL_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < NE - 1; ++k) ms_out[k] = 0.f;
    for (int it = 0; it < n_iters; ++it)
        for (int k = 0; k < NE - 1; ++k) {
            float ms = 0.f;
            UAVRL_CUDA(cudaEventElapsedTime(&ms, ev[(size_t)it * NE + k], ev[(size_t)it * NE + k + 1]));
            ms_out[k] += ms;
        }
    for (auto &e : ev) cudaEventDestroy(e);
    return 0;
}
