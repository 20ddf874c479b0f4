/* A RECORDING stand-in for the GMG part of libmgb200's C ABI (include/mgb200.h) -- test infrastructure only.
 * It performs no arithmetic: every entry point appends one line to stdout, so that tests/test_facade_queue_cpu.py can
 * check, without a GPU, WHICH library calls the facade headers issue for a given sequence of the reference's operators
 * (the lazy operator queue of mgb200_gmg_facade.hpp).  Vectors are plain host arrays so that uploads / downloads work. */
#include "mgb200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct mgb_gmg { mgb_gmg_config cfg; double *v[4][32]; size_t w[32]; int iter; };

const char *mgb_last_error(void) { return "mock"; }
void mgb_gmg_config_default(mgb_gmg_config *c) { memset(c, 0, sizeof(*c)); c->n_pre = 2; c->nu = 5; c->n_ranks = 1; }
void mgb_gmg_config_fast(mgb_gmg_config *c) { mgb_gmg_config_default(c); c->smoother = MGB_SMOOTH_GS_RB; c->restriction = MGB_RESTRICT_FULL_WEIGHTING; }
int mgb_gmg_create(const mgb_gmg_config *cfg, mgb_gmg_t *out)
{
    struct mgb_gmg *h = (struct mgb_gmg *)calloc(1, sizeof(*h));
    h->cfg = *cfg;
    size_t w = cfg->n;
    for (int l = 0; l < cfg->levels; ++l) {
        h->w[l] = w;
        for (int k = 0; k < 4; ++k) h->v[k][l] = (double *)calloc(w * w, sizeof(double));
        w = (w + 1) / 2;
    }
    printf("create n=%zu levels=%d smoother=%d restriction=%d\n", cfg->n, cfg->levels, cfg->smoother, cfg->restriction);
    *out = h;
    return MGB_OK;
}
void mgb_gmg_destroy(mgb_gmg_t h) { printf("destroy\n"); free(h); }
int mgb_gmg_set_level(mgb_gmg_t h, int level, int which, const double *host)
{
    memcpy(h->v[which][level], host, h->w[level] * h->w[level] * sizeof(double));
    printf("upload level=%d which=%d\n", level, which);
    return MGB_OK;
}
int mgb_gmg_get_level(mgb_gmg_t h, int level, int which, double *host)
{
    memcpy(host, h->v[which][level], h->w[level] * h->w[level] * sizeof(double));
    printf("download level=%d which=%d\n", level, which);
    return MGB_OK;
}
int mgb_gmg_smooth(mgb_gmg_t h, int level, int kind, int sweeps, int sol, int rhs)
{
    h->v[sol][level][0] += 1.0;                       /* a visible effect: counts the sweeps applied */
    printf("smooth level=%d kind=%d sweeps=%d sol=%d rhs=%d\n", level, kind, sweeps, sol, rhs);
    return MGB_OK;
}
int mgb_gmg_residual(mgb_gmg_t h, int level, int sol, int rhs, int store, double *sumsq)
{
    (void)h;
    printf("residual level=%d sol=%d rhs=%d store=%d\n", level, sol, rhs, store);
    *sumsq = 1.0;
    return MGB_OK;
}
int mgb_gmg_sumsq(mgb_gmg_t h, int level, int which, double *sumsq) { (void)h; printf("sumsq level=%d which=%d\n", level, which); *sumsq = 1.0; return MGB_OK; }
int mgb_gmg_prolong(mgb_gmg_t h, int lc) { (void)h; printf("prolong coarse=%d\n", lc); return MGB_OK; }
int mgb_gmg_set_cycle(mgb_gmg_t h, int smoother, int restriction, int nu, double tol, int maxit)
{
    (void)h; (void)tol; (void)maxit;
    printf("set_cycle smoother=%d restriction=%d nu=%d\n", smoother, restriction, nu);
    return MGB_OK;
}
int mgb_gmg_cycle(mgb_gmg_t h, double *coarse_relres, int *coarse_iters)
{
    h->v[MGB_VEC_U][0][0] += 100.0;
    printf("cycle\n");
    if (coarse_relres) *coarse_relres = 0.05;
    if (coarse_iters) *coarse_iters = 3;
    return MGB_OK;
}
int mgb_gmg_iterate(mgb_gmg_t h, double confirm_below, double *sumsq, double *coarse_relres)
{
    (void)confirm_below;
    h->v[MGB_VEC_U][0][0] += 102.0;                   /* = 2 sweeps + 1 cycle */
    h->iter++;
    printf("iterate #%d\n", h->iter);
    if (sumsq) *sumsq = 1.0 / (double)(h->iter * h->iter);
    if (coarse_relres) *coarse_relres = 0.05;
    return MGB_OK;
}
int mgb_gmg_krylov(mgb_gmg_t h, int method, int precond, double tol, int maxit, double *hist, int *n_hist)
{
    (void)h;
    printf("krylov method=%d precond=%d maxit=%d\n", method, precond, maxit);
    hist[0] = 1.0; hist[1] = 0.5; hist[2] = tol * 0.5;      /* "converges" in two steps */
    *n_hist = 3;
    return MGB_OK;
}
