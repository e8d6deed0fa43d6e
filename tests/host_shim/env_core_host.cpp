// Host compilation of csrc/env_core.cuh for the CPU-side logic test (tests/test_env_core_host.py).
// TEST INFRASTRUCTURE: lets the per-env step/observation source that the CUDA kernel instantiates
// be checked against the oracle and the golden vectors without a GPU.  Not part of the product
// library and never loaded by the package.
#include "../../dqn-based-uav-3d_path_planer_b200/csrc/env_core.cuh"

#include <vector>

using namespace uavrl;

// host twin of env_block.cuh's ApfDev (UAV.cal_force over the obstacle table)
struct HostApf {
    static constexpr bool enabled = true;
    const ApfObs *ob; int n;
    P3 force(double x, double y, double z) const { return apf_force(ob, n, x, y, z); }
};

extern "C" {

struct shim_batch {
    int32_t n, kmax;
    double *px, *py, *pz, *vx, *vy, *V;
    int32_t *step, *cursor, *n_sub;
    uint8_t *done, *alias0;
    double *score, *total_score, *path_len, *goal, *sub;
};

static std::vector<Cyl> make_cyl(int n, const double *b)
{
    std::vector<Cyl> c((size_t)n);
    for (int i = 0; i < n; ++i) {
        c[i].cx = b[5 * i]; c[i].cy = b[5 * i + 1]; c[i].R = b[5 * i + 3]; c[i].H = b[5 * i + 4];
        const double r2 = c[i].R * c[i].R;
        c[i].r2lo = r2 * (1.0 - 1e-12); c[i].r2hi = r2 * (1.0 + 1e-12);
    }
    return c;
}

void shim_step(double width, double h, int n_cyl, const double *buildings, double max_v, double min_v,
               double steering, double climb, int max_step, int act_mode, shim_batch *b,
               const double *actions, double *reward, uint8_t *done_ret, uint8_t *info,
               uint8_t *coll, float *obs)
{
    EnvConst k;
    k.width = width; k.h = h; k.max_v = max_v; k.min_v = min_v; k.steering = steering; k.climb = climb;
    k.max_step = max_step; k.n_cyl = n_cyl;
    const std::vector<Cyl> cyl = make_cyl(n_cyl, buildings);
    auto threat = [&](double x, double y, double z) {
        if (out_of_bounds(k, x, y, z)) return 1;
        for (int c = 0; c < n_cyl; ++c) if (cyl_hit(cyl[c], x, y, z)) return 1;
        return 0;
    };
    for (int e = 0; e < b->n; ++e) {
        EnvRegs s;
        s.px = b->px[e]; s.py = b->py[e]; s.pz = b->pz[e]; s.vx = b->vx[e]; s.vy = b->vy[e]; s.V = b->V[e];
        s.score = b->score[e]; s.total = b->total_score[e]; s.path_len = b->path_len[e];
        s.gx = b->goal[3 * e]; s.gy = b->goal[3 * e + 1]; s.gz = b->goal[3 * e + 2];
        s.step = b->step[e]; s.cursor = b->cursor[e]; s.n_sub = b->n_sub[e];
        s.done = b->done[e]; s.alias = b->alias0[e];
        s.theta = angle_xy(s.vx, s.vy);        // the kernel carries this in its state (same value)
        const double *q = b->sub + (size_t)e * b->kmax * 3;
        auto sub = [q](int i) { P3 p; p.x = q[3 * i]; p.y = q[3 * i + 1]; p.z = q[3 * i + 2]; return p; };
        if (actions) {
            StepOut o;
            step_core(k, s, act_mode, actions[e], sub, threat, o);
            reward[e] = o.reward; done_ret[e] = (uint8_t)o.done_ret; info[e] = (uint8_t)o.info;
            coll[e] = (uint8_t)o.coll;
            b->px[e] = s.px; b->py[e] = s.py; b->pz[e] = s.pz; b->vx[e] = s.vx; b->vy[e] = s.vy; b->V[e] = s.V;
            b->score[e] = s.score; b->total_score[e] = s.total; b->path_len[e] = s.path_len;
            b->step[e] = s.step; b->cursor[e] = s.cursor; b->done[e] = (uint8_t)s.done;
            b->alias0[e] = (uint8_t)s.alias;
        }
        if (obs) {
            float *o = obs + (size_t)e * 100;
            for (int i = 0; i < 100; ++i) o[i] = 0.f;
            obs_scalars(s, sub, o);
            for (int p = 0; p < 80; ++p) {
                double x, y, z; int slot;
                probe_point(p, s.px, s.py, s.pz, x, y, z, slot);
                o[slot] = threat(x, y, z) ? 1.f : 0.f;
            }
        }
    }
}

// The APF (moving-obstacle) variant the kernel instantiates with EXTRAS (env_block.cuh): step_core_apf over the env's own queue,
// then every stored entry moves by the force at its position (phase 1b: UAV.Adjust_subgoal); the observation reads the shifted
// queue.  obstacle_v: [n_cyl][3]; buildings rows are cx, cy, cz, R, H (cz = the obstacle's position.z, used by the APF distance).
void shim_step_apf(double width, double h, int n_cyl, const double *buildings, const double *obstacle_v, double max_v, double min_v,
                   double steering, double climb, int max_step, int act_mode, shim_batch *b, const double *actions, double *reward,
                   uint8_t *done_ret, uint8_t *info, uint8_t *coll, float *obs)
{
    EnvConst k;
    k.width = width; k.h = h; k.max_v = max_v; k.min_v = min_v; k.steering = steering; k.climb = climb;
    k.max_step = max_step; k.n_cyl = n_cyl;
    const std::vector<Cyl> cyl = make_cyl(n_cyl, buildings);
    std::vector<ApfObs> ob((size_t)n_cyl);
    for (int i = 0; i < n_cyl; ++i) {                            // as uavrl_env_set_extras builds the table (env.cu)
        ApfObs o;
        o.x = buildings[5 * i]; o.y = buildings[5 * i + 1]; o.z = buildings[5 * i + 2]; o.R = buildings[5 * i + 3];
        o.vx = obstacle_v[3 * i]; o.vy = obstacle_v[3 * i + 1]; o.vz = obstacle_v[3 * i + 2];
        o.vmag = sqrt(o.vx * o.vx + o.vy * o.vy + o.vz * o.vz);
        const double a = angle_xy(o.vx, o.vy);
        o.cav = cos(a); o.sav = sin(a);
        ob[(size_t)i] = o;
    }
    HostApf apf{ ob.data(), n_cyl };
    auto threat = [&](double x, double y, double z) {
        if (out_of_bounds(k, x, y, z)) return 1;
        for (int c = 0; c < n_cyl; ++c) if (cyl_hit(cyl[c], x, y, z)) return 1;
        return 0;
    };
    for (int e = 0; e < b->n; ++e) {
        EnvRegs s;
        s.px = b->px[e]; s.py = b->py[e]; s.pz = b->pz[e]; s.vx = b->vx[e]; s.vy = b->vy[e]; s.V = b->V[e];
        s.score = b->score[e]; s.total = b->total_score[e]; s.path_len = b->path_len[e];
        s.gx = b->goal[3 * e]; s.gy = b->goal[3 * e + 1]; s.gz = b->goal[3 * e + 2];
        s.step = b->step[e]; s.cursor = b->cursor[e]; s.n_sub = b->n_sub[e];
        s.done = b->done[e]; s.alias = b->alias0[e];
        s.theta = angle_xy(s.vx, s.vy);
        double *q = b->sub + (size_t)e * b->kmax * 3;
        auto sub = [q](int i) { P3 p; p.x = q[3 * i]; p.y = q[3 * i + 1]; p.z = q[3 * i + 2]; return p; };
        StepOut o;
        step_core_apf(k, s, act_mode, actions[e], sub, threat, apf, o);
        reward[e] = o.reward; done_ret[e] = (uint8_t)o.done_ret; info[e] = (uint8_t)o.info; coll[e] = (uint8_t)o.coll;
        b->px[e] = s.px; b->py[e] = s.py; b->pz[e] = s.pz; b->vx[e] = s.vx; b->vy[e] = s.vy; b->V[e] = s.V;
        b->score[e] = s.score; b->total_score[e] = s.total; b->path_len[e] = s.path_len;
        b->step[e] = s.step; b->cursor[e] = s.cursor; b->done[e] = (uint8_t)s.done; b->alias0[e] = (uint8_t)s.alias;
        for (int i = 0; i < s.n_sub && i < b->kmax; ++i) {       // phase 1b: every stored entry moves by the force at its position
            const P3 f = apf_force(ob.data(), n_cyl, q[3 * i], q[3 * i + 1], q[3 * i + 2]);
            q[3 * i] = dadd(q[3 * i], f.x); q[3 * i + 1] = dadd(q[3 * i + 1], f.y); q[3 * i + 2] = dadd(q[3 * i + 2], f.z);
        }
        if (obs) {
            float *od = obs + (size_t)e * 100;
            for (int i = 0; i < 100; ++i) od[i] = 0.f;
            obs_scalars(s, sub, od);
            for (int p = 0; p < 80; ++p) {
                double x, y, z; int slot;
                probe_point(p, s.px, s.py, s.pz, x, y, z, slot);
                od[slot] = threat(x, y, z) ? 1.f : 0.f;
            }
        }
    }
}

// the kernel's Calc_Fly_Power (UAV.py:239-245)
double shim_fly_power(double V, double P_i, double v_0, double d_0, double rho, double s, double A, double P_b, double F_b, double xi)
{
    PowerConst c;
    c.P_i = P_i; c.v_0 = v_0; c.d_0 = d_0; c.rho = rho; c.s = s; c.A = A; c.P_b = P_b; c.F_b = F_b; c.xi = xi;
    return fly_power(c, V);
}

}  // extern "C"
