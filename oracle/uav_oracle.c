/* uav_oracle.c -- CPU restatement of the reference UAV step / observation (TEST INFRASTRUCTURE).
 *
 * ORACLE ONLY (see uav_oracle.h).  Plain C, fp64 like the Python reference, operations in the
 * reference's exact order.  Build with  -O2 -ffp-contract=off -fno-builtin  so that
 * `x**2` stays a libm pow() call exactly as CPython's float_pow does it and no FMA is formed.
 *
 * Pinned against the Python reference via tests/golden/ (tests/test_oracle_golden.py).
 */
#include "uav_oracle.h"
#include <math.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* Python `a ** 2` on floats -> libm pow(a, 2.0) (Objects/floatobject.c float_pow). */
static double py_sq(double a) { return pow(a, 2.0); }

/* Python float `%` (Objects/floatobject.c float_rem): fmod, then sign fix-up. */
static double py_fmod(double v, double w)
{
    double m = fmod(v, w);
    if (m != 0.0) {
        if ((w < 0) != (m < 0)) m += w;
    } else {
        m = copysign(0.0, w);
    }
    return m;
}

/* BaseClass/CalMod.py:64-65  Eu_Loc_distance */
double ora_distance(ora_loc a, ora_loc b)
{
    return sqrt(py_sq(a.x - b.x) + py_sq(a.y - b.y) + py_sq(a.z - b.z));
}

/* BaseClass/CalMod.py:89-102  calculate_angle(p1, p2, mod=1):
 * atan2 -> math.degrees (x * (180/pi)) -> (angle + 360) % 360 / 180 * pi */
double ora_angle(ora_loc p1, ora_loc p2)
{
    double dx = p2.x - p1.x;
    double dy = p2.y - p1.y;
    double angle = atan2(dy, dx);
    angle = angle * (180.0 / M_PI);
    return py_fmod(angle + 360.0, 360.0) / 180.0 * M_PI;
}

/* Obstacles/building.py:20-26  building.check_threaten */
int32_t ora_check_threaten(const double *b, ora_loc p)
{
    const double cx = b[0], cy = b[1], cz = b[2], R = b[3], H = b[4];
    if (p.z > H) return 0;
    ora_loc q = { p.x, p.y, cz };      /* Loc(position.x, position.y, self.position.z) */
    ora_loc c = { cx, cy, cz };
    if (ora_distance(q, c) < R) return 1;
    return 0;
}

/* Envs/PathPlan_City.py:215-223  Threaten_rate -- note y is tested against `width` (:218) */
int32_t ora_threaten_rate(const ora_city *c, ora_loc p)
{
    if (p.x < 0 || p.x > c->width || p.y < 0 || p.y > c->width || p.z < 0 || p.z > c->h) return 1;
    for (int32_t i = 0; i < c->n_buildings; ++i)
        if (ora_check_threaten(c->buildings + 5 * i, p) > 0) return 1;
    return 0;
}

/* Agents/UAV.py:246-253  Calc_V */
double ora_calc_v(ora_uav *u)
{
    ora_loc o = { 0, 0, 0 }, v = { u->vx, u->vy, u->vz };
    double V = ora_distance(o, v);
    if (V > u->max_v) {
        u->vx = u->vx * (u->max_v / V);
        u->vy = u->vy * (u->max_v / V);
        V = u->max_v;
    }
    return V;
}

/* Agents/UAV.py:239-245  Calc_Fly_Power (auxiliary energy model; not part of the step reward) */
double ora_fly_power(double V, double P_i, double v_0, double d_0, double rho, double s,
                     double A, double P_b, double F_b, double xi)
{
    double induced = P_i * sqrt(sqrt(1 + pow(V, 4.0) / (4 * pow(v_0, 4.0))) - py_sq(V) / (2 * py_sq(v_0)));
    double parasite = 0.5 * d_0 * rho * s * A * pow(V, 3.0);
    double blade = xi * P_b * (1 + 3 * py_sq(V) / py_sq(F_b));
    return induced + parasite + blade;
}

/* ---- APF (moving obstacles): Agents/UAV.py:156-210, reward term :448-453.  The shipped obstacles carry no velocity
 * (Obstacles/building.py:6-11) and APF_Enabled = 0 (config/UAV.xml:6); ora_set_apf() supplies the `v` attribute of every
 * obstacle and switches the branch on (test-only global, like the reference's per-UAV flag).  NULL switches it off. */
static const double *g_apf_v = 0;
static int32_t g_apf_n = 0;
void ora_set_apf(const double *obstacle_v, int32_t n) { g_apf_v = obstacle_v; g_apf_n = n; }

/* UAV.cal_force :174-210 */
ora_loc ora_cal_force(const ora_city *c, ora_loc p)
{
    ora_loc total = { 0, 0, 0 };
    const ora_loc origin = { 0, 0, 0 };
    double cum_force = 0;
    for (int32_t i = 0; i < c->n_buildings && i < g_apf_n; ++i) {
        const double *b = c->buildings + 5 * i;
        const ora_loc v = { g_apf_v[3 * i], g_apf_v[3 * i + 1], g_apf_v[3 * i + 2] };
        if (v.x == 0 && v.y == 0 && v.z == 0) continue;                       /* :180-182 */
        const ora_loc centre = { b[0], b[1], b[2] };
        const double dis = ora_distance(p, centre);                            /* :183 */
        const double dis2edge = dis - b[3];                                    /* :185 */
        if (dis2edge > 60) continue;                                           /* :186-187 */
        const double vmag = ora_distance(origin, v);                           /* :189 */
        double f1 = 1 * b[3] / (dis2edge * dis2edge);                          /* :190  min(1, w R / d^2) */
        if (!(f1 < 1)) f1 = 1;
        const double f1_seta = ora_angle(p, centre);                           /* :191 */
        const double v_seta = ora_angle(origin, v);                            /* :193 */
        if (dis2edge < 0) f1 = (-dis2edge > 2) ? -dis2edge : 2;                /* :196-197 max(-d, 2) */
        const double f1x = -f1 * cos(f1_seta), f1y = -f1 * sin(f1_seta);       /* :198 */
        double f2 = vmag * b[3] / (dis2edge * dis2edge);                       /* :200 */
        if (!(f2 < 1)) f2 = 1;
        const double f2x = f2 * cos(v_seta), f2y = f2 * sin(v_seta);           /* :201 */
        cum_force += (f1 + f2);                                                /* :202 */
        total.x = total.x + f1x + f2x;                                         /* :203  Loc.__add__ twice */
        total.y = total.y + f1y + f2y;
        total.z = total.z + 0 + 0;
        /* :205-208: beyond 100 the reference calls Cal_SubTask_Dynamic() without its arguments (TypeError).  Unreachable with
         * fewer than ~50 obstacles in range; the product returns the force accumulated so far, and so does this oracle. */
        if (cum_force > 100) return total;
    }
    return total;
}

/* Agents/UAV.py:397-513  update_PathPlan */
void ora_step(const ora_city *c, ora_uav *u, int32_t act_mode, double action, ora_step_out *out)
{
    double r = 0.0;                                            /* global_r, :399 */
    out->collision = 0;
    if (u->n_sub - u->cursor == 0) {                           /* :400-406 */
        u->done = 1;
        r += (double)(u->max_step - u->step);
        u->score += r;
        out->reward = r; out->done_ret = 1; out->info = ORA_INFO_SUCCESS;
        return;
    }
    /* action decode.  CONTINUOUS = the reference (action[0]); DISCRETE27 = documented extension */
    double a0 = action, dz = 0.0, speed = u->max_v;
    if (act_mode == ORA_ACT_DISCRETE27) {
        int k = (int)action;
        int i = k / 9, j = (k / 3) % 3, l = k % 3;
        a0 = (double)(i - 1);
        dz = (double)(j - 1) * u->climb_rate;
        speed = (l == 0) ? u->min_v : (l == 1) ? (u->min_v + u->max_v) / 2 : u->max_v;
    }
    const ora_loc origin = { 0, 0, 0 };
    ora_loc *sg = u->sub + u->cursor;                          /* sub_goals[0] */
    if (u->alias0 && u->cursor == 0) *sg = u->pos;             /* same object as position (RRT.py:69) */

    u->step += 1;                                              /* :408 */
    ora_loc old_position = u->pos;                             /* :409 */
    ora_loc vvec = { u->vx, u->vy, u->vz };
    double seta_old = ora_angle(origin, vvec);                 /* :411 */
    double dis_old = ora_distance(u->pos, *sg);                /* :412 */
    double dis2goal_old = ora_distance(u->pos, u->goal);       /* :413 */
    double seta_new = seta_old + a0 * u->steering;             /* :414 */
    u->vx = speed * cos(seta_new);                             /* :415 */
    u->vy = speed * sin(seta_new);                             /* :416 */
    u->V = ora_calc_v(u);                                      /* :417 */
    u->pos.x += u->vx;                                         /* :419 */
    u->pos.y += u->vy;                                         /* :420 */
    if (act_mode == ORA_ACT_DISCRETE27) u->pos.z += dz;        /* extension: climb with the same move */
    if (u->alias0 && u->cursor == 0) *sg = u->pos;             /* the aliased sub-goal moved too */
    vvec.x = u->vx; vvec.y = u->vy; vvec.z = u->vz;
    double tri_goal = ora_angle(u->pos, *sg);                  /* :422 */
    double tri_V = ora_angle(origin, vvec);                    /* :423 */
    if (ora_threaten_rate(c, u->pos) == 1) {                   /* :425 */
        r -= 0.3;
        u->pos = old_position;                                 /* :427 rebinds position: alias broken */
        tri_V = ora_angle(u->pos, *sg);                        /* :428 */
        out->collision = 1;
    }
    double dis_new = ora_distance(u->pos, *sg);                /* :429 */
    double dis2goal_new = ora_distance(u->pos, u->goal);       /* :430 */
    r -= 0.13 * fabs(a0);                                      /* :434 */
    r += 0.2 * cos(fabs(tri_goal - tri_V));                    /* :435 */
    r += 0.4 * (dis_old - dis_new);                            /* :436 */
    r += 0.4 * (dis2goal_old - dis2goal_new);                  /* :437 */
    r -= 0.1;                                                  /* :438 */
    r -= 0.01 * fabs(u->pos.z - sg->z);                        /* :439-440 */
    u->path_len += u->V;                                       /* :443 */
    if (g_apf_v) {                                             /* :448-453, APF_Enabled == 1 */
        for (int32_t i = u->cursor; i < u->n_sub; ++i) {       /* Adjust_subgoal :156-166: every remaining sub-goal moves */
            const ora_loc f = ora_cal_force(c, u->sub[i]);
            u->sub[i].x = u->sub[i].x + f.x; u->sub[i].y = u->sub[i].y + f.y; u->sub[i].z = u->sub[i].z + f.z;
        }
        const ora_loc total_force = ora_cal_force(c, u->pos);  /* :450 */
        const double force = ora_distance(origin, total_force);                /* :451 */
        const double tri_force = ora_angle(origin, total_force);               /* :452 */
        r += 0.2 * force * cos(fabs(tri_force - tri_V));       /* :453 */
    }

    if (u->step >= u->max_step) {                              /* :456-465 */
        u->done = 1;
        r += (50 - ora_distance(u->pos, *sg));
        u->score += r; u->total_score += r;
        out->reward = r; out->done_ret = 1; out->info = ORA_INFO_LOSE;
    } else if (ora_distance(u->pos, *sg) < 7 ||
               (ora_distance(u->pos, u->goal) < ora_distance(*sg, u->goal))) {  /* :466 */
        r += (50 - ora_distance(u->pos, *sg));                 /* :468 */
        u->cursor += 1;                                        /* :469 pop(0) */
        if (u->n_sub - u->cursor == 0) {                       /* :470-483 */
            r += 50;
            u->done = 1;
            r += (double)(u->max_step - u->step);
            u->score += r; u->total_score += r;
            out->reward = r; out->done_ret = 1; out->info = ORA_INFO_SUCCESS;
        } else {                                               /* :484-495 */
            u->step = 0; u->score = 0;                         /* reset("local reset") :329-332 */
            u->V = ora_calc_v(u);
            sg = u->sub + u->cursor;
            vvec.x = u->vx; vvec.y = u->vy; vvec.z = u->vz;
            tri_goal = ora_angle(u->pos, *sg);                 /* :488 */
            tri_V = ora_angle(origin, vvec);                   /* :489 */
            r += 0.2 * cos(fabs(tri_goal - tri_V));            /* :490 */
            r += (double)(u->max_step - u->step);              /* :491 (=Max_Step after the reset) */
            u->score += r; u->total_score += r;
            out->reward = r; out->done_ret = 1; out->info = ORA_INFO_SUCCESS;
        }
    } else if (ora_distance(u->pos, u->goal) < 7) {            /* :496-509 */
        u->done = 1;
        r += 50;
        r += (double)(u->max_step - u->step);
        u->score += r; u->total_score += r;
        out->reward = r; out->done_ret = 1; out->info = ORA_INFO_SUCCESS;
    } else {                                                   /* :510-513 */
        u->score += r; u->total_score += r;
        out->reward = r; out->done_ret = 0; out->info = ORA_INFO_NORMAL;
    }
    u->alias0 = 0;   /* after the first step sub_goals[0] was popped (or the episode ended) */
}

/* Agents/UAV.py:515-567  state_PathPlan */
void ora_state(const ora_city *c, const ora_uav *u, double *o)
{
    memset(o, 0, 100 * sizeof(double));
    const int32_t nleft = u->n_sub - u->cursor;
    const ora_loc origin = { 0, 0, 0 };
    const ora_loc p = u->pos;
    o[0] = (double)u->step / 100;                              /* :518 */
    if (nleft >= 1) {                                          /* :519-522 */
        ora_loc sg = u->sub[u->cursor];
        if (u->alias0 && u->cursor == 0) sg = p;
        o[1] = (sg.x - p.x) / 10;
        o[2] = (sg.y - p.y) / 10;
        o[3] = (sg.z - p.z) / 10;
    }
    o[4] = u->V;                                               /* :523 */
    o[5] = u->vx;
    o[6] = u->vy;
    ora_loc vvec = { u->vx, u->vy, u->vz };
    o[7] = ora_angle(origin, vvec);                            /* :526 */
    if (nleft >= 2) {                                          /* :528-531 */
        const ora_loc s1 = u->sub[u->cursor + 1];
        o[8] = (s1.x - p.x) / 10;
        o[9] = (s1.y - p.y) / 10;
        o[10] = (s1.z - p.z) / 10;
    }
    static const int scale[3] = { 1, 5, 10 };
    static const int base[3] = { 11, 36, 61 };
    for (int g = 0; g < 3; ++g)                                /* :533-555 */
        for (int i = 0; i < 5; ++i)
            for (int j = 0; j < 5; ++j) {
                int dx = (i - 2), dy = (j - 2);
                ora_loc t = { p.x + (double)(scale[g] * dx), p.y + (double)(scale[g] * dy), p.z };
                o[base[g] + 5 * i + j] = (double)ora_threaten_rate(c, t);
            }
    o[86] = (u->goal.x - p.x) / 10;                            /* :557-559 */
    o[87] = (u->goal.y - p.y) / 10;
    o[88] = (u->goal.z - p.z) / 10;
    o[89] = p.z / 10;                                          /* :560 */
    for (int k = 0; k < 5; ++k) {                              /* :562-566 */
        ora_loc t = { p.x, p.y, p.z - (double)(k + 1) };
        o[90 + k] = (double)ora_threaten_rate(c, t);
    }
}

/* ---------------------------------------------------------------- batched SoA drivers */
static void load(const ora_batch *b, int32_t e, double max_v, double min_v, double steering,
                 double climb_rate, int32_t max_step, ora_uav *u)
{
    u->max_v = max_v; u->min_v = min_v; u->steering = steering; u->climb_rate = climb_rate;
    u->max_step = max_step;
    u->pos.x = b->px[e]; u->pos.y = b->py[e]; u->pos.z = b->pz[e];
    u->vx = b->vx[e]; u->vy = b->vy[e]; u->vz = 0.0; u->V = b->V[e];
    u->step = b->step[e]; u->done = b->done[e];
    u->score = b->score[e]; u->total_score = b->total_score[e]; u->path_len = b->path_len[e];
    u->goal.x = b->goal[3 * e]; u->goal.y = b->goal[3 * e + 1]; u->goal.z = b->goal[3 * e + 2];
    u->n_sub = b->n_sub[e]; u->cursor = b->cursor[e];
    u->sub = (ora_loc *)(b->sub + (size_t)e * b->kmax * 3);
    u->alias0 = b->alias0[e];
}

static void store(ora_batch *b, int32_t e, const ora_uav *u)
{
    b->px[e] = u->pos.x; b->py[e] = u->pos.y; b->pz[e] = u->pos.z;
    b->vx[e] = u->vx; b->vy[e] = u->vy; b->V[e] = u->V;
    b->step[e] = u->step; b->done[e] = (uint8_t)u->done;
    b->score[e] = u->score; b->total_score[e] = u->total_score; b->path_len[e] = u->path_len;
    b->cursor[e] = u->cursor; b->alias0[e] = (uint8_t)u->alias0;
}

void ora_batch_step(const ora_city *c, ora_batch *b, double max_v, double min_v, double steering,
                    double climb_rate, int32_t max_step, int32_t act_mode, const double *actions,
                    double *reward, uint8_t *done_ret, uint8_t *info, uint8_t *collision, float *obs)
{
    /* envs are independent: the multi-core CPU baseline runs them on all host threads */
    #pragma omp parallel for schedule(static)
    for (int32_t e = 0; e < b->n; ++e) {
        double o[100];
        ora_uav u; ora_step_out out;
        load(b, e, max_v, min_v, steering, climb_rate, max_step, &u);
        ora_step(c, &u, act_mode, actions[e], &out);
        store(b, e, &u);
        reward[e] = out.reward; done_ret[e] = (uint8_t)out.done_ret;
        info[e] = (uint8_t)out.info; collision[e] = (uint8_t)out.collision;
        if (obs) {
            ora_state(c, &u, o);
            for (int k = 0; k < 100; ++k) obs[(size_t)e * 100 + k] = (float)o[k];
        }
    }
}

void ora_batch_state(const ora_city *c, const ora_batch *b, double max_v, double min_v,
                     double steering, double climb_rate, int32_t max_step, float *obs, double *obs64)
{
    #pragma omp parallel for schedule(static)
    for (int32_t e = 0; e < b->n; ++e) {
        double o[100];
        ora_uav u;
        load(b, e, max_v, min_v, steering, climb_rate, max_step, &u);
        ora_state(c, &u, o);
        for (int k = 0; k < 100; ++k) {
            if (obs) obs[(size_t)e * 100 + k] = (float)o[k];
            if (obs64) obs64[(size_t)e * 100 + k] = o[k];
        }
    }
}

/* number of host threads the batched drivers / learner oracle use (cpu_baseline reports it) */
int32_t ora_set_threads(int32_t n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return (int32_t)omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}
