/* per_oracle.c -- see per_oracle.h.  Plain sequential C, fp64 tree like the reference's numpy array. */
#include "per_oracle.h"

#include <math.h>
#include <stdlib.h>

struct ora_per {
    int32_t cap, data_pointer, n_entries;
    double *tree;                 /* 2*cap - 1 nodes, leaves at cap-1 .. 2*cap-2  (replay_buffer.py:58-68) */
    double alpha, beta, beta_inc, eps, err_upper;
};

ora_per *ora_per_create(int32_t capacity)
{
    ora_per *p = (ora_per *)calloc(1, sizeof(ora_per));
    p->cap = capacity;
    p->tree = (double *)calloc((size_t)(2 * capacity - 1), sizeof(double));
    p->alpha = 0.6; p->beta = 0.4; p->beta_inc = 0.001; p->eps = 0.01; p->err_upper = 1.0;   /* :141-148 */
    return p;
}

void ora_per_destroy(ora_per *p) { if (p) { free(p->tree); free(p); } }

static void tree_update(ora_per *p, int64_t idx, double pr)      /* SumTree.update :70-79 */
{
    const double change = pr - p->tree[idx];
    p->tree[idx] = pr;
    while (idx != 0) { idx = (idx - 1) / 2; p->tree[idx] += change; }
}

void ora_per_add(ora_per *p, double pr)                           /* SumTree.add :81-97 */
{
    tree_update(p, (int64_t)p->data_pointer + p->cap - 1, pr);
    p->data_pointer += 1;
    if (p->data_pointer >= p->cap) p->data_pointer = 0;
    if (p->n_entries < p->cap) p->n_entries += 1;
}

void ora_per_push(ora_per *p, float abs_err)
{
    const float pr = powf(fabsf(abs_err) + (float)p->eps, (float)p->alpha);
    ora_per_add(p, (double)pr);
}

static int64_t get_leaf(const ora_per *p, double v)              /* SumTree.get_leaf :99-120 */
{
    const int64_t len = 2 * (int64_t)p->cap - 1;
    int64_t parent = 0;
    for (;;) {
        const int64_t cl = 2 * parent + 1, cr = cl + 1;
        if (cl >= len) return parent;
        if (v <= p->tree[cl]) parent = cl;
        else { v -= p->tree[cl]; parent = cr; }
    }
}

double ora_per_sample(ora_per *p, int32_t B, const double *u, int64_t *tree_idx, double *w)
{
    const double total = (double)(long long)p->tree[0];           /* SumTree.total(): int(tree[0]) :122-123 */
    const double seg = total / (double)B;
    p->beta = fmin(1.0, p->beta + p->beta_inc);                   /* :195 */
    double wmax = 0.0;
    for (int i = 0; i < B; ++i) {
        const double a = seg * (double)i, b = seg * (double)(i + 1);
        const double s = a + (b - a) * u[i];                      /* random.uniform(a, b) */
        const int64_t leaf = get_leaf(p, s);
        tree_idx[i] = leaf;
        const double prob = p->tree[leaf] / total;
        w[i] = pow((double)p->n_entries * prob, -p->beta);        /* :209 */
        if (w[i] > wmax) wmax = w[i];
    }
    for (int i = 0; i < B; ++i) w[i] /= wmax;
    return p->beta;
}

void ora_per_batch_update(ora_per *p, int32_t B, const int64_t *tree_idx, const float *abs_err)
{
    for (int i = 0; i < B; ++i) {
        float e = abs_err[i] + (float)p->eps;
        if (e > (float)p->err_upper) e = (float)p->err_upper;
        tree_update(p, tree_idx[i], (double)powf(e, (float)p->alpha));
    }
}

void ora_per_leaves(const ora_per *p, double *out) { for (int i = 0; i < p->cap; ++i) out[i] = p->tree[p->cap - 1 + i]; }
double ora_per_total(const ora_per *p) { return p->tree[0]; }
int32_t ora_per_n_entries(const ora_per *p) { return p->n_entries; }
