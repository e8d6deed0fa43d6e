// scenario.cu -- host-side scenario generator: UAV.reset() draws + RRT sub-goal queues.
//
// Restates (statistically -- the reference's Python MT19937 stream is not reproduced; SURVEY.md
// section 7 "RNG"):
//   UAV.reset          Agents/UAV.py:344-360   heading ~U(0,2pi), start ~U(10,210)xU(1,10), z=0,
//                                              goal ~U(330,490)xU(420,490), z=0
//   RRTPlanner.getPath PathPlan/RRT.py:26-105  goal bias 0.5, step = sub_granularity, <=10000
//                                              iterations, collision sampled every 5 m, rewiring
// Runs once per pool on the host (it feeds the lockstep kernel, it is not on the per-step path);
// a device-side RRT is the "next" row 8f-1.
#include <math.h>

#include <atomic>
#include <thread>
#include <vector>

#include "common.cuh"
#include "env_core.cuh"

namespace {

using namespace uavrl;

struct Rng {
    uint64_t key, stream, ctr = 0;
    uint32_t buf[4];
    int have = 0;
    Rng(uint64_t k, uint64_t s) : key(k), stream(s) {}
    uint32_t next32()
    {
        if (have == 0) { Philox::gen(key, ctr++, stream, buf); have = 4; }
        return buf[--have];
    }
    double u01()     // 53-bit uniform in [0,1), like random.random()
    {
        const uint64_t a = next32() >> 5, b = next32() >> 6;
        return (double)(a * 67108864ull + b) * (1.0 / 9007199254740992.0);
    }
    double uniform(double lo, double hi) { return lo + (hi - lo) * u01(); }     // random.uniform
};

struct Node { P3 loc; int parent; double cost; };

struct City {
    EnvConst k;
    std::vector<Cyl> cyl;
    double len;
    int threat(const P3 &p) const
    {
        if (out_of_bounds(k, p.x, p.y, p.z)) return 1;
        for (const Cyl &c : cyl) if (cyl_hit(c, p.x, p.y, p.z)) return 1;
        return 0;
    }
    // RRT.py:48-56
    bool obstacle_free(const P3 &a, const P3 &b, double step_size) const
    {
        const int steps = (int)(dist3(a.x, a.y, a.z, b.x, b.y, b.z) / step_size);
        for (int i = 0; i <= steps; ++i) {
            P3 q;
            q.x = a.x + (b.x - a.x) * i / (steps + 1);
            q.y = a.y + (b.y - a.y) * i / (steps + 1);
            q.z = a.z + (b.z - a.z) * i / (steps + 1);
            if (threat(q) == 1) return false;
        }
        return true;
    }
};

// RRT.py:63-105.  Returns the node chain start..goal (empty if the goal was not connected).
std::vector<P3> rrt(const City &c, Rng &rng, const P3 &start, const P3 &goal, double step_size)
{
    const int max_iter = 10000;
    const double obstacle_step = 5.0;
    std::vector<Node> nodes;
    nodes.push_back({ start, -1, 0.0 });
    int goal_parent = -1;
    for (int it = 0; it < max_iter; ++it) {
        P3 rp;
        if (rng.uniform(0, 1) > 0.5) {                                   // :27-32
            rp.x = rng.uniform(0, c.len); rp.y = rng.uniform(0, c.k.width); rp.z = rng.uniform(0, c.k.h);
        } else {
            rp = goal;
        }
        int nearest = 0;                                                 // :36-37
        double best = 1e300;
        for (size_t i = 0; i < nodes.size(); ++i) {
            const double dd = dist3(nodes[i].loc.x, nodes[i].loc.y, nodes[i].loc.z, rp.x, rp.y, rp.z);
            if (dd < best) { best = dd; nearest = (int)i; }
        }
        const P3 from = nodes[nearest].loc;                              // :39-46
        const double dx = rp.x - from.x, dy = rp.y - from.y, dz = rp.z - from.z;
        const double length = sqrt(dx * dx + dy * dy + dz * dz);
        P3 nl;
        if (length < step_size) nl = rp;
        else { nl.x = from.x + dx / length * step_size; nl.y = from.y + dy / length * step_size; nl.z = from.z + dz / length * step_size; }
        if (!c.obstacle_free(from, nl, obstacle_step)) continue;         // :79-80
        Node nn{ nl, nearest, nodes[nearest].cost + dist3(from.x, from.y, from.z, nl.x, nl.y, nl.z) };
        nodes.push_back(nn);
        const int me = (int)nodes.size() - 1;
        for (int i = 0; i < me; ++i) {                                   // :86-92 rewire the new node
            const double dd = dist3(nodes[i].loc.x, nodes[i].loc.y, nodes[i].loc.z, nl.x, nl.y, nl.z);
            if (dd < step_size && nodes[me].cost > nodes[i].cost + dd) {
                if (c.obstacle_free(nodes[i].loc, nl, obstacle_step)) { nodes[me].parent = i; nodes[me].cost = nodes[i].cost + dd; }
            }
        }
        if (dist3(nl.x, nl.y, nl.z, goal.x, goal.y, goal.z) <= step_size) { goal_parent = me; break; }   // :94-96
    }
    std::vector<P3> path;
    if (goal_parent < 0) return path;
    path.push_back(goal);
    for (int i = goal_parent; i >= 0; i = nodes[i].parent) path.push_back(nodes[i].loc);
    return std::vector<P3>(path.rbegin(), path.rend());
}

}  // namespace

extern "C" int uavrl_make_scenarios(const uavrl_env_config *cfg, uint64_t seed, int32_t P, int32_t rrt_step,
                                    double *start, double *goal, double *heading, double *sub, int32_t *n_sub)
{
    if (!cfg || P <= 0 || !start || !goal || !heading || !sub || !n_sub)
        return fail(UAVRL_ERR_INVALID, "uavrl_make_scenarios: null/empty argument");
    if (cfg->n_buildings > 0 && !cfg->buildings_host) return fail(UAVRL_ERR_INVALID, "buildings_host is null");
    City c;
    c.k.width = cfg->width; c.k.h = cfg->h; c.k.max_v = cfg->max_v; c.k.min_v = cfg->min_v;
    c.k.steering = cfg->steering_angle; c.k.climb = cfg->climb_rate; c.k.max_step = cfg->max_step;
    c.k.n_cyl = cfg->n_buildings; c.len = cfg->len;
    for (int i = 0; i < cfg->n_buildings; ++i) {
        const double *b = cfg->buildings_host + 5 * i;
        Cyl cy; cy.cx = b[0]; cy.cy = b[1]; cy.R = b[3]; cy.H = b[4];
        const double r2 = cy.R * cy.R; cy.r2lo = r2 * (1.0 - 1e-12); cy.r2hi = r2 * (1.0 + 1e-12);
        c.cyl.push_back(cy);
    }
    const int K = cfg->max_subgoals;
    const double step = rrt_step > 0 ? (double)rrt_step : 30.0;       // config/UAV.xml sub_granularity
    // scenarios are independent: spread them over the host cores
    unsigned nthreads = std::thread::hardware_concurrency();
    if (nthreads == 0) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    if ((int)nthreads > P) nthreads = (unsigned)P;
    std::atomic<int> next{0}, failed{0};
    auto worker = [&]() {
    for (int s = next.fetch_add(1); s < P; s = next.fetch_add(1)) {
        bool ok = false;
        for (int attempt = 0; attempt < 64 && !ok; ++attempt) {
            Rng rng(seed, ((uint64_t)s << 8) | (uint64_t)attempt);
            const double seta = rng.uniform(0, 2 * kPi);               // UAV.py:344
            P3 st{ rng.uniform(10, 210), rng.uniform(1, 10), 0.0 };     // :353-355
            P3 gl{ rng.uniform(330, 490), rng.uniform(420, 490), 0.0 }; // :356-358
            std::vector<P3> path = rrt(c, rng, st, gl, step);
            if (path.empty() || (int)path.size() > K) continue;
            heading[s] = seta;
            start[3 * s] = st.x; start[3 * s + 1] = st.y; start[3 * s + 2] = st.z;
            goal[3 * s] = gl.x; goal[3 * s + 1] = gl.y; goal[3 * s + 2] = gl.z;
            double *q = sub + (size_t)s * K * 3;
            for (int i = 0; i < K * 3; ++i) q[i] = 0.0;
            for (size_t i = 0; i < path.size(); ++i) { q[3 * i] = path[i].x; q[3 * i + 1] = path[i].y; q[3 * i + 2] = path[i].z; }
            n_sub[s] = (int32_t)path.size();
            ok = true;
        }
        if (!ok) failed.store(1);
    }
    };
    std::vector<std::thread> pool;
    for (unsigned i = 0; i < nthreads; ++i) pool.emplace_back(worker);
    for (auto &th : pool) th.join();
    if (failed.load()) return fail(UAVRL_ERR_INVALID, "RRT found no path within max_subgoals for a scenario");
    return 0;
}
