/* dqn_oracle.c -- CPU restatement of the DQN-family learner math (TEST INFRASTRUCTURE, ORACLE ONLY).
 * See dqn_oracle.h for the reference citations.  Dot products accumulate in double and round to
 * fp32 once per output (torch rounds per fp32 FMA; both are within ~1e-6 relative of each other,
 * the tolerance the parity tests state).
 */
#include "dqn_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { const float *W, *b; float *gW, *gb; int in, out; } layer;

static int n_layers(const ora_net *n) { return n->n_hidden + (n->dueling ? 2 : 1); }

int64_t ora_net_param_count(const ora_net *n)
{
    int64_t p = 0; int in = n->in_dim;
    for (int l = 0; l < n->n_hidden; ++l) { p += (int64_t)n->hidden[l] * in + n->hidden[l]; in = n->hidden[l]; }
    p += (int64_t)n->n_actions * in + n->n_actions;
    if (n->dueling) p += in + 1;
    return p;
}

/* carve flat params (and optionally a flat grad buffer) into layers */
static void carve(const ora_net *n, const float *params, float *grads, layer *L)
{
    int in = n->in_dim; int64_t off = 0; int k = 0;
    for (int l = 0; l < n->n_hidden; ++l, ++k) {
        int out = n->hidden[l];
        L[k].W = params + off; L[k].gW = grads ? grads + off : 0; off += (int64_t)out * in;
        L[k].b = params + off; L[k].gb = grads ? grads + off : 0; off += out;
        L[k].in = in; L[k].out = out; in = out;
    }
    L[k].W = params + off; L[k].gW = grads ? grads + off : 0; off += (int64_t)n->n_actions * in;
    L[k].b = params + off; L[k].gb = grads ? grads + off : 0; off += n->n_actions;
    L[k].in = in; L[k].out = n->n_actions; ++k;
    if (n->dueling) {
        L[k].W = params + off; L[k].gW = grads ? grads + off : 0; off += in;
        L[k].b = params + off; L[k].gb = grads ? grads + off : 0; off += 1;
        L[k].in = in; L[k].out = 1;
    }
}

static void linear(const layer *l, const float *x, float *y, int relu)
{
    for (int o = 0; o < l->out; ++o) {
        double acc = 0.0;
        const float *w = l->W + (size_t)o * l->in;
        for (int i = 0; i < l->in; ++i) acc += (double)w[i] * (double)x[i];
        float v = (float)(acc + (double)l->b[o]);
        y[o] = (relu && v < 0.f) ? 0.f : v;
    }
}

#define MAXW 1024
/* forward one sample; acts[l] = post-ReLU activation of trunk layer l (acts[-1] = x) */
static void fwd1(const ora_net *n, const layer *L, const float *x, float acts[][MAXW], float *q)
{
    const float *cur = x;
    for (int l = 0; l < n->n_hidden; ++l) { linear(&L[l], cur, acts[l], 1); cur = acts[l]; }
    if (!n->dueling) { linear(&L[n->n_hidden], cur, q, 0); return; }
    float A[MAXW], V;
    linear(&L[n->n_hidden], cur, A, 0);
    linear(&L[n->n_hidden + 1], cur, &V, 0);
    float mean = 0.f;
    { double s = 0; for (int a = 0; a < n->n_actions; ++a) s += A[a]; mean = (float)(s / n->n_actions); }
    for (int a = 0; a < n->n_actions; ++a) q[a] = V + A[a] - mean;   /* BaseCNN.py:138 */
}

void ora_net_forward(const ora_net *n, const float *params, const float *x, int32_t B, float *q)
{
    layer L[ORA_MAX_HIDDEN + 2];
    carve(n, params, 0, L);
    #pragma omp parallel for schedule(static)
    for (int32_t b = 0; b < B; ++b) {
        float acts[ORA_MAX_HIDDEN][MAXW];
        fwd1(n, L, x + (size_t)b * n->in_dim, acts, q + (size_t)b * n->n_actions);
    }
}

static int argmax(const float *q, int n)
{
    int best = 0;
    for (int a = 1; a < n; ++a) if (q[a] > q[best]) best = a;
    return best;
}

void ora_act(const ora_net *n, const float *params, const float *x, int32_t B, float eps,
             int32_t is_train, const float *u, const int32_t *rand_action, int32_t *action, float *q_out)
{
    layer L[ORA_MAX_HIDDEN + 2];
    carve(n, params, 0, L);
    #pragma omp parallel for schedule(static)
    for (int32_t b = 0; b < B; ++b) {
        float acts[ORA_MAX_HIDDEN][MAXW], q[MAXW];
        fwd1(n, L, x + (size_t)b * n->in_dim, acts, q);
        if (q_out) memcpy(q_out + (size_t)b * n->n_actions, q, sizeof(float) * n->n_actions);
        if (u[b] > eps || !is_train) action[b] = argmax(q, n->n_actions);   /* :90-94 */
        else action[b] = rand_action[b];                                      /* :97 */
    }
}

/* is_w / abs_err_out: prioritised-replay form of the update (NULL = the reference's plain MSE): loss =
 * mean(w_i (Q-y)^2), |Q-y| per sample returned for ReplayTree.batch_update.  The reference multiplies the importance
 * weights into the loss the same way (SAC_Trainer.py:348-352) but on the already averaged loss, which is not
 * back-propagatable; the per-sample form is the standard one. */
/* 0 = MSE (every reference trainer, BaseTrainer.py:40); 1 = Huber = torch.nn.SmoothL1Loss(beta = 1), the option the
 * product offers beside it.  Test-only switch. */
static int g_loss_kind = 0;
void ora_set_loss_kind(int32_t k) { g_loss_kind = k; }

static float dqn_update_impl(const ora_net *n, int32_t algo, float *local, const float *target,
                     float *m, float *v, int64_t *t,
                     const float *s, const int32_t *a, const float *r, const float *s2,
                     const float *d, int32_t B, float gamma, float lr, float *grads_out, const float *is_w, float *abs_err_out)
{
    const int64_t P = ora_net_param_count(n);
    const int nl = n_layers(n);
    double *G = (double *)calloc((size_t)P, sizeof(double));
    double loss_sum = 0.0;

    #pragma omp parallel
    {
        double *g = (double *)calloc((size_t)P, sizeof(double));
        float *gdummy = (float *)calloc((size_t)P, sizeof(float));
        layer L[ORA_MAX_HIDDEN + 2], T[ORA_MAX_HIDDEN + 2];
        carve(n, local, gdummy, L);          /* gW/gb pointers used only for offsets */
        carve(n, target, 0, T);
        double lsum = 0.0;
        #pragma omp for schedule(static)
        for (int32_t b = 0; b < B; ++b) {
            float acts[ORA_MAX_HIDDEN][MAXW], acts2[ORA_MAX_HIDDEN][MAXW];
            float q[MAXW], qn_t[MAXW], qn_l[MAXW];
            const float *x = s + (size_t)b * n->in_dim, *x2 = s2 + (size_t)b * n->in_dim;
            fwd1(n, L, x, acts, q);
            fwd1(n, T, x2, acts2, qn_t);
            float next_q;
            if (algo == ORA_ALGO_DQN) {
                next_q = qn_t[argmax(qn_t, n->n_actions)];                /* DQN_Trainer.py:109 */
            } else {
                fwd1(n, L, x2, acts2, qn_l);                              /* DDQN_Trainer.py:94 */
                next_q = qn_t[argmax(qn_l, n->n_actions)];                /* :95 */
            }
            const float y = r[b] + (gamma * next_q * (1.f - d[b]));       /* :99 / :114 / :171 */
            const float diff = q[a[b]] - y;
            const float wb = is_w ? is_w[b] : 1.f;
            if (abs_err_out) abs_err_out[b] = fabsf(diff);
            float gq;
            if (g_loss_kind == 0) {
                lsum += (double)(wb * (diff * diff));
                gq = (2.f * diff * wb) / (float)B;                        /* d mean(w (Q-y)^2) / dQ */
            } else {
                const float ad = fabsf(diff);
                lsum += (double)(wb * (ad < 1.f ? 0.5f * (diff * diff) : ad - 0.5f));
                gq = ((diff > 1.f ? 1.f : diff < -1.f ? -1.f : diff) * wb) / (float)B;
            }
            /* backward through the head(s) */
            float dh[MAXW], dh_prev[MAXW];
            const int top = n->n_hidden;          /* index of the A / plain head */
            const float *hin = acts[n->n_hidden - 1];
            const int hdim = L[top].in;
            for (int i = 0; i < hdim; ++i) dh[i] = 0.f;
            if (!n->dueling) {
                const int o = a[b];
                for (int i = 0; i < hdim; ++i) {
                    g[(L[top].gW - gdummy) + (size_t)o * hdim + i] += (double)gq * hin[i];
                    dh[i] += gq * L[top].W[(size_t)o * hdim + i];
                }
                g[(L[top].gb - gdummy) + o] += gq;
            } else {
                const int nA = n->n_actions;
                for (int o = 0; o < nA; ++o) {
                    const float gA = gq * ((o == a[b] ? 1.f : 0.f) - 1.f / (float)nA);
                    for (int i = 0; i < hdim; ++i) {
                        g[(L[top].gW - gdummy) + (size_t)o * hdim + i] += (double)gA * hin[i];
                        dh[i] += gA * L[top].W[(size_t)o * hdim + i];
                    }
                    g[(L[top].gb - gdummy) + o] += gA;
                }
                for (int i = 0; i < hdim; ++i) {                           /* V head */
                    g[(L[top + 1].gW - gdummy) + i] += (double)gq * hin[i];
                    dh[i] += gq * L[top + 1].W[i];
                }
                g[(L[top + 1].gb - gdummy)] += gq;
            }
            /* trunk, last hidden layer first */
            for (int l = n->n_hidden - 1; l >= 0; --l) {
                const float *in = (l == 0) ? x : acts[l - 1];
                const int din = L[l].in, dout = L[l].out;
                for (int i = 0; i < din; ++i) dh_prev[i] = 0.f;
                for (int o = 0; o < dout; ++o) {
                    const float go = (acts[l][o] > 0.f) ? dh[o] : 0.f;     /* ReLU' */
                    if (go == 0.f) continue;
                    for (int i = 0; i < din; ++i) {
                        g[(L[l].gW - gdummy) + (size_t)o * din + i] += (double)go * in[i];
                        dh_prev[i] += go * L[l].W[(size_t)o * din + i];
                    }
                    g[(L[l].gb - gdummy) + o] += go;
                }
                memcpy(dh, dh_prev, sizeof(float) * din);
            }
        }
        #pragma omp critical
        {
            for (int64_t i = 0; i < P; ++i) G[i] += g[i];
            loss_sum += lsum;
        }
        free(g); free(gdummy);
    }
    (void)nl;

    /* torch.optim.Adam (single-tensor path): lerp m, addcmul v, bias corrections, addcdiv */
    *t += 1;
    const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
    const double bc1 = 1.0 - pow(b1, (double)*t), bc2 = 1.0 - pow(b2, (double)*t);
    const float step_size = (float)((double)lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    for (int64_t i = 0; i < P; ++i) {
        const float gi = (float)G[i];
        if (grads_out) grads_out[i] = gi;
        m[i] = m[i] + (gi - m[i]) * (float)(1.0 - b1);
        v[i] = v[i] * (float)b2 + (float)(1.0 - b2) * gi * gi;
        const float denom = sqrtf(v[i]) / bc2_sqrt + (float)eps;
        local[i] = local[i] - step_size * (m[i] / denom);
    }
    free(G);
    return (float)(loss_sum / (double)B);
}

float ora_dqn_update(const ora_net *n, int32_t algo, float *local, const float *target, float *m, float *v, int64_t *t,
                     const float *s, const int32_t *a, const float *r, const float *s2, const float *d, int32_t B,
                     float gamma, float lr, float *grads_out)
{
    return dqn_update_impl(n, algo, local, target, m, v, t, s, a, r, s2, d, B, gamma, lr, grads_out, 0, 0);
}

float ora_dqn_update_per(const ora_net *n, int32_t algo, float *local, const float *target, float *m, float *v, int64_t *t,
                         const float *s, const int32_t *a, const float *r, const float *s2, const float *d, int32_t B,
                         float gamma, float lr, float *grads_out, const float *is_w, float *abs_err_out)
{
    return dqn_update_impl(n, algo, local, target, m, v, t, s, a, r, s2, d, B, gamma, lr, grads_out, is_w, abs_err_out);
}
