// scenario.cu -- scenario generator: UAV.reset() draws + RRT sub-goal queues (arithmetic in rrt_core.cuh).
//
//   uavrl_make_scenarios      host threads, host arrays out (feeds uavrl_env_set_pool)
//   uavrl_env_generate_pool   device kernel, one thread per scenario, writes the env's device pool directly
//                             (with tens of thousands of auto-resetting envs the host RRT is the
//                             bottleneck); bit-identical to the host generator for the same (seed, index)
//   uavrl_env_get_pool        read a device pool back (tests, checkpoints)
#include <math.h>

#include <atomic>
#include <thread>
#include <vector>

#include "common.cuh"
#include "env.cuh"
#include "rrt_core.cuh"

namespace {

using namespace uavrl;

void fill_city(const uavrl_env_config *cfg, RrtCity &c, std::vector<Cyl> &cyl)
{
    c.k.width = cfg->width; c.k.h = cfg->h; c.k.max_v = cfg->max_v; c.k.min_v = cfg->min_v;
    c.k.steering = cfg->steering_angle; c.k.climb = cfg->climb_rate; c.k.max_step = cfg->max_step;
    c.k.n_cyl = cfg->n_buildings; c.len = cfg->len;
    for (int i = 0; i < cfg->n_buildings; ++i) {
        const double *b = cfg->buildings_host + 5 * i;
        Cyl cy; cy.cx = b[0]; cy.cy = b[1]; cy.R = b[3]; cy.H = b[4];
        const double r2 = cy.R * cy.R; cy.r2lo = r2 * (1.0 - 1e-12); cy.r2hi = r2 * (1.0 + 1e-12);
        cyl.push_back(cy);
    }
    c.cyl = cyl.data();
}

// one thread per scenario; tree scratch nodes[P][kRrtMaxNodes] in global memory
__global__ void rrt_pool_kernel(RrtCity c, uint64_t seed, int P, double step, int K, RrtNode *nodes, double *start,
                                double *goal, double *v0, double *sub, int32_t *n_sub, uint8_t *alias, int *failed)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P) return;
    double heading = 0.0;
    const int len = make_scenario(c, seed, s, step, K, nodes + (size_t)s * kRrtMaxNodes, start + 3 * (size_t)s,
                                  goal + 3 * (size_t)s, &heading, sub + (size_t)s * K * 3);
    if (len == 0) { atomicExch(failed, 1); n_sub[s] = 0; return; }
    n_sub[s] = len;
    alias[s] = 1;                                   // RRT.py:69: queue[0] is the UAV's own position object
    // UAV.py:344-348: V_vector = Max_V*(cos, sin)(heading); V = Calc_V()
    double vx = uavrl::dmul(c.k.max_v, cos(heading)), vy = uavrl::dmul(c.k.max_v, sin(heading));
    const double V = calc_v(c.k, vx, vy);
    v0[3 * (size_t)s] = vx; v0[3 * (size_t)s + 1] = vy; v0[3 * (size_t)s + 2] = V;
}

}  // namespace

extern "C" int uavrl_make_scenarios(const uavrl_env_config *cfg, uint64_t seed, int32_t P, int32_t rrt_step,
                                    double *start, double *goal, double *heading, double *sub, int32_t *n_sub)
{
    if (!cfg || P <= 0 || !start || !goal || !heading || !sub || !n_sub)
        return fail(UAVRL_ERR_INVALID, "uavrl_make_scenarios: null/empty argument");
    if (cfg->n_buildings > 0 && !cfg->buildings_host) return fail(UAVRL_ERR_INVALID, "buildings_host is null");
    RrtCity c;
    std::vector<Cyl> cyl;
    fill_city(cfg, c, cyl);
    const int K = cfg->max_subgoals;
    const double step = rrt_step > 0 ? (double)rrt_step : 30.0;       // config/UAV.xml sub_granularity
    // scenarios are independent: spread them over the host cores
    unsigned nthreads = std::thread::hardware_concurrency();
    if (nthreads == 0) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    if ((int)nthreads > P) nthreads = (unsigned)P;
    std::atomic<int> next{0}, failed{0};
    auto worker = [&]() {
        std::vector<RrtNode> nodes(kRrtMaxNodes);
        for (int s = next.fetch_add(1); s < P; s = next.fetch_add(1)) {
            n_sub[s] = make_scenario(c, seed, s, step, K, nodes.data(), start + 3 * (size_t)s, goal + 3 * (size_t)s,
                                     heading + s, sub + (size_t)s * K * 3);
            if (n_sub[s] == 0) failed.store(1);
        }
    };
    std::vector<std::thread> pool;
    for (unsigned i = 0; i < nthreads; ++i) pool.emplace_back(worker);
    for (auto &th : pool) th.join();
    if (failed.load()) return fail(UAVRL_ERR_INVALID, "RRT found no path within max_subgoals for a scenario");
    return 0;
}

extern "C" int uavrl_env_generate_pool(uavrl_env *env, int32_t P, uint64_t seed, int32_t rrt_step, void *stream)
{
    if (!env || P <= 0) return fail(UAVRL_ERR_INVALID, "uavrl_env_generate_pool: null/empty argument");
    UAVRL_CUDA(cudaSetDevice(env->cfg.device));
    EnvDev &d = env->d;
    cudaStream_t st = (cudaStream_t)stream;
    RrtCity c;
    c.k = d.k; c.len = env->cfg.len; c.cyl = d.cyl;
    const double step = rrt_step > 0 ? (double)rrt_step : 30.0;
    double *ps, *pg, *pv, *pq; int32_t *pn; uint8_t *pa; RrtNode *nodes; int *failed;
    const size_t sub_n = (size_t)P * d.K * 3;
    UAVRL_CUDA(cudaMalloc((void **)&ps, (size_t)P * 3 * sizeof(double)));
    UAVRL_CUDA(cudaMalloc((void **)&pg, (size_t)P * 3 * sizeof(double)));
    UAVRL_CUDA(cudaMalloc((void **)&pv, (size_t)P * 3 * sizeof(double)));
    UAVRL_CUDA(cudaMalloc((void **)&pq, sub_n * sizeof(double)));
    UAVRL_CUDA(cudaMalloc((void **)&pn, (size_t)P * sizeof(int32_t)));
    UAVRL_CUDA(cudaMalloc((void **)&pa, (size_t)P));
    UAVRL_CUDA(cudaMalloc((void **)&nodes, (size_t)P * kRrtMaxNodes * sizeof(RrtNode)));
    UAVRL_CUDA(cudaMalloc((void **)&failed, sizeof(int)));
    UAVRL_CUDA(cudaMemsetAsync(failed, 0, sizeof(int), st));
    const int threads = 32, blocks = (P + threads - 1) / threads;     // divergent single-thread searches: small CTAs spread them over all SMs
    rrt_pool_kernel<<<blocks, threads, 0, st>>>(c, seed, P, step, d.K, nodes, ps, pg, pv, pq, pn, pa, failed);
    UAVRL_LAUNCHED();
    int h_failed = 0;
    UAVRL_CUDA(cudaMemcpyAsync(&h_failed, failed, sizeof(int), cudaMemcpyDeviceToHost, st));
    UAVRL_CUDA(cudaStreamSynchronize(st));
    cudaFree(nodes); cudaFree(failed);
    if (h_failed) {
        cudaFree(ps); cudaFree(pg); cudaFree(pv); cudaFree(pq); cudaFree(pn); cudaFree(pa);
        return fail(UAVRL_ERR_INVALID, "device RRT found no path within max_subgoals for a scenario");
    }
    UAVRL_CUDA(cudaDeviceSynchronize());            // nothing may still read the pool being replaced
    free_pool(d);
    d.pool_start = ps; d.pool_goal = pg; d.pool_v0 = pv; d.pool_sub = pq; d.pool_nsub = pn; d.pool_alias = pa;
    d.P = P;
    env->pool_set = true;
    env->reset_done = false;
    return 0;
}

extern "C" int uavrl_env_get_pool(uavrl_env *env, double *start, double *goal, double *v0, double *sub, int32_t *n_sub)
{
    if (!env) return fail(UAVRL_ERR_INVALID, "null env");
    if (!env->pool_set) return fail(UAVRL_ERR_STATE, "uavrl_env_get_pool before a pool was set");
    UAVRL_CUDA(cudaSetDevice(env->cfg.device));
    UAVRL_CUDA(cudaDeviceSynchronize());
    const EnvDev &d = env->d;
    const size_t P = (size_t)d.P;
    if (start) UAVRL_CUDA(cudaMemcpy(start, d.pool_start, P * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (goal) UAVRL_CUDA(cudaMemcpy(goal, d.pool_goal, P * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (v0) UAVRL_CUDA(cudaMemcpy(v0, d.pool_v0, P * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (sub) UAVRL_CUDA(cudaMemcpy(sub, d.pool_sub, P * d.K * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (n_sub) UAVRL_CUDA(cudaMemcpy(n_sub, d.pool_nsub, P * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return 0;
}
