// env.cuh -- device-side view of a batch of UAV environments (SoA, fp64 state) + host handle.
#pragma once
#include "common.cuh"
#include "env_core.cuh"

#include <vector>

namespace uavrl {

constexpr int kEnvThreads = 128;       // warp 0 steps the CTA's envs; all 4 warps share their probes
// envs per CTA: 8 while the whole batch then fits ONE wave (112 registers x 128 threads -> 4 CTAs per SM, 4 x 148 CTAs), so
// that the latency-bound fp64 chains of a 4096-env step spread over 512 CTAs instead of 128; 32 above (16 384 envs = 512
// CTAs of 32: again one wave -- with 8 per CTA they took 3.5 waves: 33.5 us against 14.6 us at 4096 envs, measured)
constexpr int kEnvsPerBlockLarge = 32, kEnvsPerBlockSmall = 8;
constexpr int kSmallBatchEnvs = 4 * 148 * kEnvsPerBlockSmall;
constexpr int kMaxCyl = 64;            // candidate sets are 64-bit masks

// Everything a kernel needs, passed by value.
struct EnvDev {
    EnvConst k;
    int32_t n, K, P;
    int32_t auto_reset;
    double cull_w;                      // half-width of the probe window incl. one step of motion
    const Cyl *cyl;
    // per-env state, structure of arrays
    double *px, *py, *pz, *vx, *vy, *V, *score, *total, *path_len, *gx, *gy, *gz, *rew64, *theta;
    int32_t *step, *cursor, *n_sub, *scen;
    uint8_t *done, *alias;
    // scenario pool (read-only during stepping)
    const double *pool_start, *pool_goal, *pool_v0, *pool_sub;
    const int32_t *pool_nsub;
    const uint8_t *pool_alias;
    // running statistics: [0] env steps, [1] episodes ended, [2] collisions ; sum_reward separately
    unsigned long long *stat_counts;
    double *stat_reward;
    // optional models (uavrl_env_set_extras); extras = bit mask kExtra*
    int32_t extras;
    PowerConst pw;                      // energy: Calc_Fly_Power constants
    double *energy;                     // [n] accumulated over the episode in progress
    const ApfObs *apf_obs;              // APF: every obstacle with its velocity (device)
    double *sub_env;                    // APF: per-env sub-goal queues [n][K][3] (shifted every step)
    double *path_buf;                   // track: [2][track_n][track_cap][3]
    int32_t *path_n;                    // [2][track_n] points recorded (buffer 0/1)
    int32_t *path_cur;                  // [track_n] buffer holding the episode in progress
    int32_t track_n, track_cap;
    long long *trace;                   // debug (UAVRL_ENV_TRACE): CTA 0 / thread 0 stage timestamps
};
constexpr int kExtraEnergy = 1, kExtraApf = 2, kExtraTrack = 4;

}  // namespace uavrl

struct uavrl_env {
    uavrl_env_config cfg;
    uavrl::EnvDev d;
    bool pool_set = false, reset_done = false;
    // staging for the host-buffer entry point
    void *h_act_dev = nullptr;
    float *h_obs_dev = nullptr, *h_rew_dev = nullptr;
    uint8_t *h_flags_dev = nullptr;      // done | info | collision | ended, each [n]
    cudaStream_t own_stream = nullptr;
    bool extras_set = false;
    std::vector<double> base_z;          // building base heights (position.z), used by the APF distance only
};

namespace uavrl {
// launched by env.cu and by the fused training loop (train.cu)
int launch_env_step(const EnvDev &d, int action_kind, const void *actions, float *obs, float *reward,
                    uint8_t *done, uint8_t *info, uint8_t *coll, uint8_t *ended, cudaStream_t st, bool pdl = false);
int launch_env_observe(const EnvDev &d, float *obs, cudaStream_t st);
void free_pool(EnvDev &d);            // scenario.cu replaces the pool with a device-generated one
}  // namespace uavrl
