/* sac_oracle.c -- see sac_oracle.h (TEST INFRASTRUCTURE, ORACLE ONLY). */
#include "sac_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXH 256
#define MAXA 8

int64_t ora_sac_actor_params(const ora_sac_cfg *c) { return (int64_t)c->hidden * c->obs_dim + c->hidden + 2 * ((int64_t)c->act_dim * c->hidden + c->act_dim); }
int64_t ora_sac_critic_params(const ora_sac_cfg *c)
{
    const int in = c->obs_dim + c->act_dim;
    return (int64_t)c->hidden * in + c->hidden + (int64_t)c->hidden * c->hidden + c->hidden + (int64_t)c->act_dim * c->hidden + c->act_dim;
}

static float dotf(const float *w, const float *x, int n, float b)
{
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc += (double)w[i] * (double)x[i];
    return (float)(acc + (double)b);
}
static float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }        /* F.softplus, beta 1, threshold 20 */
static float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

typedef struct { float h[MAXH], pm[MAXA], ps[MAXA], mu[MAXA], sd[MAXA], xs[MAXA], a[MAXA], t[MAXA], logp[MAXA]; } actor_cache;

/* BaseCNN.py:471-483 */
static void actor_fwd(const ora_sac_cfg *c, const float *p, const float *s, const float *eps, actor_cache *k)
{
    const int H = c->hidden, D = c->obs_dim, A = c->act_dim;
    const float *W1 = p, *b1 = W1 + (size_t)H * D, *Wm = b1 + H, *bm = Wm + (size_t)A * H, *Ws = bm + A, *bs = Ws + (size_t)A * H;
    for (int o = 0; o < H; ++o) { float v = dotf(W1 + (size_t)o * D, s, D, b1[o]); k->h[o] = v > 0.f ? v : 0.f; }
    for (int j = 0; j < A; ++j) {
        k->pm[j] = dotf(Wm + (size_t)j * H, k->h, H, bm[j]);
        k->ps[j] = dotf(Ws + (size_t)j * H, k->h, H, bs[j]);
        k->mu[j] = tanhf(k->pm[j]);
        k->sd[j] = tanhf(softplus_f(k->ps[j]));
        k->xs[j] = k->mu[j] + k->sd[j] * eps[j];                               /* rsample */
        const float var = k->sd[j] * k->sd[j];
        float lp = -((k->xs[j] - k->mu[j]) * (k->xs[j] - k->mu[j])) / (2.f * var) - logf(k->sd[j]) - 0.91893853320467274178f;
        k->a[j] = tanhf(k->xs[j]);
        k->t[j] = tanhf(k->a[j]);                                              /* tanh applied twice (:481) */
        k->logp[j] = lp - logf(1.f - k->t[j] * k->t[j] + 1e-7f);
    }
}

typedef struct { float x[MAXH + MAXA], h1[MAXH], h2[MAXH], q[MAXA]; } critic_cache;

/* BaseCNN.py:496-500 */
static void critic_fwd(const ora_sac_cfg *c, const float *p, const float *s, const float *a, critic_cache *k)
{
    const int H = c->hidden, A = c->act_dim, I = c->obs_dim + A;
    const float *W1 = p, *b1 = W1 + (size_t)H * I, *W2 = b1 + H, *b2 = W2 + (size_t)H * H, *W3 = b2 + H, *b3 = W3 + (size_t)A * H;
    memcpy(k->x, s, sizeof(float) * c->obs_dim);
    for (int j = 0; j < A; ++j) k->x[c->obs_dim + j] = a[j];
    for (int o = 0; o < H; ++o) { float v = dotf(W1 + (size_t)o * I, k->x, I, b1[o]); k->h1[o] = v > 0.f ? v : 0.f; }
    for (int o = 0; o < H; ++o) { float v = dotf(W2 + (size_t)o * H, k->h1, H, b2[o]); k->h2[o] = v > 0.f ? v : 0.f; }
    for (int j = 0; j < A; ++j) k->q[j] = dotf(W3 + (size_t)j * H, k->h2, H, b3[j]);
}

/* backward of one sample through a critic: dq[A] -> parameter grads (accumulated into g, may be NULL) and the
 * gradient w.r.t. the action inputs da[A] (may be NULL) */
static void critic_bwd(const ora_sac_cfg *c, const float *p, const critic_cache *k, const float *dq, double *g, float *da)
{
    const int H = c->hidden, A = c->act_dim, I = c->obs_dim + A;
    const float *W1 = p, *W2 = W1 + (size_t)H * I + H, *W3 = W2 + (size_t)H * H + H;
    const size_t oW1 = 0, ob1 = (size_t)H * I, oW2 = ob1 + H, ob2 = oW2 + (size_t)H * H, oW3 = ob2 + H, ob3 = oW3 + (size_t)A * H;
    float dh2[MAXH], dh1[MAXH];
    for (int i = 0; i < H; ++i) dh2[i] = 0.f;
    for (int j = 0; j < A; ++j) {
        if (g) { for (int i = 0; i < H; ++i) g[oW3 + (size_t)j * H + i] += (double)dq[j] * k->h2[i]; g[ob3 + j] += dq[j]; }
        for (int i = 0; i < H; ++i) dh2[i] += dq[j] * W3[(size_t)j * H + i];
    }
    for (int i = 0; i < H; ++i) dh1[i] = 0.f;
    for (int o = 0; o < H; ++o) {
        const float go = k->h2[o] > 0.f ? dh2[o] : 0.f;
        if (go == 0.f) continue;
        if (g) { for (int i = 0; i < H; ++i) g[oW2 + (size_t)o * H + i] += (double)go * k->h1[i]; g[ob2 + o] += go; }
        for (int i = 0; i < H; ++i) dh1[i] += go * W2[(size_t)o * H + i];
    }
    if (da) for (int j = 0; j < A; ++j) da[j] = 0.f;
    for (int o = 0; o < H; ++o) {
        const float go = k->h1[o] > 0.f ? dh1[o] : 0.f;
        if (go == 0.f) continue;
        if (g) { for (int i = 0; i < I; ++i) g[oW1 + (size_t)o * I + i] += (double)go * k->x[i]; g[ob1 + o] += go; }
        if (da) for (int j = 0; j < A; ++j) da[j] += go * W1[(size_t)o * I + c->obs_dim + j];
    }
}

static void adam_step(float *p, float *m, float *v, const double *g, int64_t n, int64_t t, float lr)
{
    const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
    const double bc1 = 1.0 - pow(b1, (double)t), bc2 = 1.0 - pow(b2, (double)t);
    const float step_size = (float)((double)lr / bc1), bc2s = (float)sqrt(bc2);
    for (int64_t i = 0; i < n; ++i) {
        const float gi = (float)g[i];
        m[i] = m[i] + (gi - m[i]) * (float)(1.0 - b1);
        v[i] = v[i] * (float)b2 + (float)(1.0 - b2) * gi * gi;
        p[i] = p[i] - step_size * (m[i] / (sqrtf(v[i]) / bc2s + (float)eps));
    }
}

void ora_sac_actor_forward(const ora_sac_cfg *c, const float *actor, const float *s, const float *eps, int32_t B, float *action, float *log_prob)
{
    for (int32_t b = 0; b < B; ++b) {
        actor_cache k;
        actor_fwd(c, actor, s + (size_t)b * c->obs_dim, eps + (size_t)b * c->act_dim, &k);
        for (int j = 0; j < c->act_dim; ++j) { action[(size_t)b * c->act_dim + j] = k.a[j] * c->action_bound; log_prob[(size_t)b * c->act_dim + j] = k.logp[j]; }
    }
}

void ora_sac_critic_forward(const ora_sac_cfg *c, const float *critic, const float *s, const float *a, int32_t B, float *q)
{
    for (int32_t b = 0; b < B; ++b) {
        critic_cache k;
        critic_fwd(c, critic, s + (size_t)b * c->obs_dim, a + (size_t)b * c->act_dim, &k);
        for (int j = 0; j < c->act_dim; ++j) q[(size_t)b * c->act_dim + j] = k.q[j];
    }
}

float ora_sac_update(const ora_sac_cfg *c, ora_sac_state *st, const float *s, const float *a, const float *r, const float *s2,
                     const float *d, const float *eps_next, const float *eps_cur, int32_t B, float *l1_out, float *l2_out, float *la_out)
{
    const int A = c->act_dim, D = c->obs_dim, H = c->hidden;
    const int64_t Pa = ora_sac_actor_params(c), Pc = ora_sac_critic_params(c);
    const float alpha = expf(st->log_alpha);
    const float nel = (float)B * (float)A;
    float *td = (float *)malloc(sizeof(float) * (size_t)B * A);
    /* calc_target (:122-131) */
    for (int32_t b = 0; b < B; ++b) {
        actor_cache ka; critic_cache k1, k2;
        actor_fwd(c, st->actor, s2 + (size_t)b * D, eps_next + (size_t)b * A, &ka);
        float an[MAXA];
        for (int j = 0; j < A; ++j) an[j] = ka.a[j] * c->action_bound;
        critic_fwd(c, st->t1, s2 + (size_t)b * D, an, &k1);
        critic_fwd(c, st->t2, s2 + (size_t)b * D, an, &k2);
        for (int j = 0; j < A; ++j) {
            const float nv = fminf(k1.q[j], k2.q[j]) + alpha * (-ka.logp[j]);
            td[(size_t)b * A + j] = r[b] + c->gamma * nv * (1.f - d[b]);
        }
    }
    /* critics (:343-360): loss = mean over [B,A] of (Q - td)^2 */
    st->step += 1;
    double l1 = 0, l2 = 0;
    for (int which = 0; which < 2; ++which) {
        float *p = which ? st->c2 : st->c1;
        double *g = (double *)calloc((size_t)Pc, sizeof(double));
        double ls = 0;
        for (int32_t b = 0; b < B; ++b) {
            critic_cache k; float dq[MAXA];
            critic_fwd(c, p, s + (size_t)b * D, a + (size_t)b * A, &k);
            for (int j = 0; j < A; ++j) { const float diff = k.q[j] - td[(size_t)b * A + j]; ls += (double)diff * diff; dq[j] = 2.f * diff / nel; }
            critic_bwd(c, p, &k, dq, g, 0);
        }
        if (which) l2 = ls / nel; else l1 = ls / nel;
        adam_step(p, which ? st->c2_m : st->c1_m, which ? st->c2_v : st->c1_v, g, Pc, st->step, c->critic_lr);
        free(g);
    }
    /* actor (:362-369) with the UPDATED critics; alpha (:371-376) */
    double *ga = (double *)calloc((size_t)Pa, sizeof(double));
    double actor_loss = 0, ent_sum = 0;
    const size_t oW1 = 0, ob1 = (size_t)H * D, oWm = ob1 + H, obm = oWm + (size_t)A * H, oWs = obm + A, obs_ = oWs + (size_t)A * H;
    const float *Wm = st->actor + oWm, *Ws = st->actor + oWs;
    for (int32_t b = 0; b < B; ++b) {
        actor_cache ka; critic_cache k1, k2;
        const float *x = s + (size_t)b * D, *e = eps_cur + (size_t)b * A;
        actor_fwd(c, st->actor, x, e, &ka);
        float an[MAXA], dq1[MAXA], dq2[MAXA], da1[MAXA], da2[MAXA];
        for (int j = 0; j < A; ++j) an[j] = ka.a[j] * c->action_bound;
        critic_fwd(c, st->c1, x, an, &k1);
        critic_fwd(c, st->c2, x, an, &k2);
        for (int j = 0; j < A; ++j) {
            actor_loss += (double)(alpha * ka.logp[j]) - (double)fminf(k1.q[j], k2.q[j]);    /* -alpha*entropy - min q */
            ent_sum += (double)(-ka.logp[j]);
            const float gmin = -1.f / nel;                                                   /* d loss / d min(q1,q2) */
            dq1[j] = k1.q[j] < k2.q[j] ? gmin : (k1.q[j] > k2.q[j] ? 0.f : 0.5f * gmin);
            dq2[j] = k2.q[j] < k1.q[j] ? gmin : (k2.q[j] > k1.q[j] ? 0.f : 0.5f * gmin);
        }
        critic_bwd(c, st->c1, &k1, dq1, 0, da1);
        critic_bwd(c, st->c2, &k2, dq2, 0, da2);
        float dpm[MAXA], dps[MAXA];
        for (int j = 0; j < A; ++j) {
            const float glogp = alpha / nel;
            const float t = ka.t[j], aa = ka.a[j];
            const float dc_da = 2.f * t * (1.f - t * t) / (1.f - t * t + 1e-7f);            /* d(-log(1 - tanh(a)^2 + 1e-7)) / da */
            const float dxs = (da1[j] + da2[j]) * c->action_bound * (1.f - aa * aa) + glogp * dc_da * (1.f - aa * aa);
            const float dmu = dxs;
            const float dsd = dxs * e[j] + glogp * (-1.f / ka.sd[j]);
            dpm[j] = dmu * (1.f - ka.mu[j] * ka.mu[j]);
            const float sp_grad = ka.ps[j] > 20.f ? 1.f : sigmoid_f(ka.ps[j]);
            dps[j] = dsd * (1.f - ka.sd[j] * ka.sd[j]) * sp_grad;
        }
        float dh[MAXH];
        for (int i = 0; i < H; ++i) dh[i] = 0.f;
        for (int j = 0; j < A; ++j) {
            for (int i = 0; i < H; ++i) {
                ga[oWm + (size_t)j * H + i] += (double)dpm[j] * ka.h[i];
                ga[oWs + (size_t)j * H + i] += (double)dps[j] * ka.h[i];
                dh[i] += dpm[j] * Wm[(size_t)j * H + i] + dps[j] * Ws[(size_t)j * H + i];
            }
            ga[obm + j] += dpm[j]; ga[obs_ + j] += dps[j];
        }
        for (int o = 0; o < H; ++o) {
            const float go = ka.h[o] > 0.f ? dh[o] : 0.f;
            if (go == 0.f) continue;
            for (int i = 0; i < D; ++i) ga[oW1 + (size_t)o * D + i] += (double)go * x[i];
            ga[ob1 + o] += go;
        }
    }
    adam_step(st->actor, st->actor_m, st->actor_v, ga, Pa, st->step, c->actor_lr);
    free(ga);
    /* alpha_loss = mean((entropy - target_entropy).detach() * exp(log_alpha)) */
    const double mean_term = ent_sum / nel - (double)c->target_entropy;
    const double gla = mean_term * (double)alpha;
    if (la_out) *la_out = (float)gla;
    adam_step(&st->log_alpha, &st->la_m, &st->la_v, &gla, 1, st->step, c->alpha_lr);
    /* soft_update (:145-147) */
    for (int64_t i = 0; i < Pc; ++i) {
        st->t1[i] = st->t1[i] * (1.0f - c->tau) + st->c1[i] * c->tau;
        st->t2[i] = st->t2[i] * (1.0f - c->tau) + st->c2[i] * c->tau;
    }
    if (l1_out) *l1_out = (float)l1;
    if (l2_out) *l2_out = (float)l2;
    free(td);
    return (float)(actor_loss / nel);
}
