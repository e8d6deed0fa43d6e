// rrt_core.cuh -- UAV.reset() draws + RRT sub-goal queue for ONE scenario, host/device.
//
// The same source runs in the host generator (uavrl_make_scenarios) and in the device pool kernel
// (uavrl_env_generate_pool, scenario.cu): every + - * / sqrt goes through the explicitly rounded
// helpers of env_core.cuh, the RNG is counter-based, so both produce the SAME scenario for the same
// (seed, scenario index) -- the device generator is tested bit-for-bit against the host one.
//
// Statistical restatement (the reference's Python MT19937 stream is not reproduced) of
//   UAV.reset          Agents/UAV.py:344-360   heading ~U(0,2pi), start ~U(10,210)xU(1,10), z=0,
//                                              goal ~U(330,490)xU(420,490), z=0
//   RRTPlanner.getPath PathPlan/RRT.py:26-105  goal bias 0.5, step = sub_granularity, <= 10000
//                                              iterations, collision sampled every 5 m, parent choice
//                                              among the nodes within one step of the new node
#pragma once
#include "common.cuh"
#include "env_core.cuh"

namespace uavrl {

constexpr int kRrtMaxNodes = 512;       // tree cap per attempt (the reference's list is unbounded; trees of the
                                        // 500x500x100 city stay below ~200 nodes); a full tree fails the attempt
constexpr int kRrtMaxIter = 10000;      // RRT.py:11
constexpr int kRrtAttempts = 64;

struct RrtRng {
    uint64_t key, stream, ctr;
    uint32_t buf[4];
    int have;
    UAVRL_HD void init(uint64_t k, uint64_t s) { key = k; stream = s; ctr = 0; have = 0; }
    UAVRL_HD uint32_t next32()
    {
        if (have == 0) { Philox::gen(key, ctr++, stream, buf); have = 4; }
        return buf[--have];
    }
    UAVRL_HD double u01()     // 53-bit uniform in [0,1), like random.random()
    {
        const uint64_t a = next32() >> 5, b = next32() >> 6;
        return dmul((double)(a * 67108864ull + b), 1.0 / 9007199254740992.0);
    }
    UAVRL_HD double uniform(double lo, double hi) { return dadd(lo, dmul(dsub(hi, lo), u01())); }     // random.uniform
};

struct RrtNode { double x, y, z, cost; int32_t parent, pad; };

struct RrtCity {
    EnvConst k;
    double len;
    const Cyl *cyl;
};

UAVRL_HD int rrt_threat(const RrtCity &c, double x, double y, double z)      // PathPlan_City.py:215-223
{
    if (out_of_bounds(c.k, x, y, z)) return 1;
    for (int i = 0; i < c.k.n_cyl; ++i) if (cyl_hit(c.cyl[i], x, y, z)) return 1;
    return 0;
}

// RRT.py:48-56: steps = int(dist / 5); points a + (b - a) * i / (steps + 1), i = 0..steps
UAVRL_HD bool rrt_obstacle_free(const RrtCity &c, double ax, double ay, double az, double bx, double by, double bz, double step_size)
{
    const int steps = (int)ddiv(dist3(ax, ay, az, bx, by, bz), step_size);
    const double den = (double)(steps + 1);
    const double dx = dsub(bx, ax), dy = dsub(by, ay), dz = dsub(bz, az);
    for (int i = 0; i <= steps; ++i) {
        const double fi = (double)i;
        const double qx = dadd(ax, ddiv(dmul(dx, fi), den));
        const double qy = dadd(ay, ddiv(dmul(dy, fi), den));
        const double qz = dadd(az, ddiv(dmul(dz, fi), den));
        if (rrt_threat(c, qx, qy, qz) == 1) return false;
    }
    return true;
}

// RRT.py:63-105.  Writes the chain start..goal into path[<= K][3]; returns its length, 0 if the goal was not
// connected, the tree filled up, or the chain is longer than K.
UAVRL_HD int rrt_plan(const RrtCity &c, RrtRng &rng, const P3 &start, const P3 &goal, double step_size, RrtNode *nodes,
                      double *path, int K)
{
    const double obstacle_step = 5.0;
    int n = 1, goal_parent = -1;
    nodes[0].x = start.x; nodes[0].y = start.y; nodes[0].z = start.z; nodes[0].cost = 0.0; nodes[0].parent = -1;
    for (int it = 0; it < kRrtMaxIter; ++it) {
        double rx, ry, rz;
        if (rng.uniform(0.0, 1.0) > 0.5) {                                  // :27-32
            rx = rng.uniform(0.0, c.len); ry = rng.uniform(0.0, c.k.width); rz = rng.uniform(0.0, c.k.h);
        } else {
            rx = goal.x; ry = goal.y; rz = goal.z;
        }
        int nearest = 0;                                                    // :36-37 (first minimum)
        double best = 1e300;
        for (int i = 0; i < n; ++i) {
            const double dd = dist3(nodes[i].x, nodes[i].y, nodes[i].z, rx, ry, rz);
            if (dd < best) { best = dd; nearest = i; }
        }
        const double fx = nodes[nearest].x, fy = nodes[nearest].y, fz = nodes[nearest].z;      // :39-46
        const double dx = dsub(rx, fx), dy = dsub(ry, fy), dz = dsub(rz, fz);
        const double length = dsqrt(dadd(dadd(dmul(dx, dx), dmul(dy, dy)), dmul(dz, dz)));
        double nx, ny, nz;
        if (length < step_size) { nx = rx; ny = ry; nz = rz; }
        else {
            nx = dadd(fx, dmul(ddiv(dx, length), step_size));
            ny = dadd(fy, dmul(ddiv(dy, length), step_size));
            nz = dadd(fz, dmul(ddiv(dz, length), step_size));
        }
        if (!rrt_obstacle_free(c, fx, fy, fz, nx, ny, nz, obstacle_step)) continue;             // :79-80
        if (n >= kRrtMaxNodes) return 0;
        const int me = n++;
        nodes[me].x = nx; nodes[me].y = ny; nodes[me].z = nz; nodes[me].parent = nearest;
        nodes[me].cost = dadd(nodes[nearest].cost, dist3(fx, fy, fz, nx, ny, nz));
        for (int i = 0; i < me; ++i) {                                      // :86-92 cheaper parent for the new node
            const double dd = dist3(nodes[i].x, nodes[i].y, nodes[i].z, nx, ny, nz);
            if (dd < step_size && nodes[me].cost > dadd(nodes[i].cost, dd)) {
                if (rrt_obstacle_free(c, nodes[i].x, nodes[i].y, nodes[i].z, nx, ny, nz, obstacle_step)) {
                    nodes[me].parent = i; nodes[me].cost = dadd(nodes[i].cost, dd);
                }
            }
        }
        if (dist3(nx, ny, nz, goal.x, goal.y, goal.z) <= step_size) { goal_parent = me; break; }   // :94-96
    }
    if (goal_parent < 0) return 0;
    int len = 1;
    for (int i = goal_parent; i >= 0; i = nodes[i].parent) ++len;
    if (len > K) return 0;
    int w = len - 1;
    path[3 * w] = goal.x; path[3 * w + 1] = goal.y; path[3 * w + 2] = goal.z;
    for (int i = goal_parent; i >= 0; i = nodes[i].parent) { --w; path[3 * w] = nodes[i].x; path[3 * w + 1] = nodes[i].y; path[3 * w + 2] = nodes[i].z; }
    return len;
}

// One scenario: up to kRrtAttempts independent (reset draw, RRT) attempts on streams (s << 8 | attempt).
// Outputs: start[3], goal[3], *heading, sub[K][3] (zero padded), returns n_sub (0 = no path found).
UAVRL_HD int make_scenario(const RrtCity &c, uint64_t seed, int s, double step_size, int K, RrtNode *nodes,
                           double *start, double *goal, double *heading, double *sub)
{
    for (int attempt = 0; attempt < kRrtAttempts; ++attempt) {
        RrtRng rng;
        rng.init(seed, ((uint64_t)s << 8) | (uint64_t)attempt);
        const double seta = rng.uniform(0.0, 2 * kPi);                      // UAV.py:344
        P3 st, gl;
        st.x = rng.uniform(10.0, 210.0); st.y = rng.uniform(1.0, 10.0); st.z = 0.0;          // :353-355
        gl.x = rng.uniform(330.0, 490.0); gl.y = rng.uniform(420.0, 490.0); gl.z = 0.0;      // :356-358
        const int len = rrt_plan(c, rng, st, gl, step_size, nodes, sub, K);
        if (len == 0) continue;
        for (int i = 3 * len; i < 3 * K; ++i) sub[i] = 0.0;
        *heading = seta;
        start[0] = st.x; start[1] = st.y; start[2] = st.z;
        goal[0] = gl.x; goal[1] = gl.y; goal[2] = gl.z;
        return len;
    }
    return 0;
}

}  // namespace uavrl
