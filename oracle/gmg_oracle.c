/*
 * gmg_oracle.c -- CPU restatement of the reference's geometric-multigrid solve phase.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under multigrid_prj_b200/ links, imports or executes
 * this file; it is the checker used by tests/, by __graft_entry__.smoke() and by the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * Parity status: PINNED.  tests/test_oracle_gmg.py checks this file against
 *   (1) the reference's two committed golden runs (GeometricMultigrid/test/{MGGS4.txt,x.mtx},
 *       WebInterface/{MGGS4.txt,x.mtx}; copies in tests/golden/), and
 *   (2) when oracle/_ref/libgmgref.so exists, the reference's own classes compiled from
 *       /root/reference by oracle/Makefile -- bit for bit, operator by operator.
 *
 * Memory model = the reference's: every level lives in ONE fine-sized N*N row-major array
 * and level l touches the entries mask_l(i) = s*(i/w)*N + s*(i%w), s = 2^l
 * (GeometricMultigrid/include/domain.hpp:78-80).  Arithmetic is written in the reference's
 * source order, compiled with -ffp-contract=off so no FMA is formed (x86-64 g++ -O3 forms
 * none for the reference either), hence results are bit-identical to the reference.
 *
 * All citations are relative to /root/reference/GeometricMultigrid/.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    size_t N;      /* fine points per side                      (domain.hpp:46 m_size) */
    size_t w;      /* points per side on this level             (domain.hpp:49 width)  */
    size_t s;      /* stride 2^level in the fine array          (domain.hpp:47 step)   */
    double diag;   /* 4*alpha/k, k=(h*s)^2                      (linear_system.hpp:27-28) */
    double off;    /* -alpha/k                                  (linear_system.hpp:38)   */
} gmgo_level;

/* src/domain.cpp:4-13 (ctor: width=(width+1)/2, step*=2 per level; m_h=length/(size-1));
 * include/linear_system.hpp:16-17 (k = h()*h(), h() = m_h*step, domain.hpp:92). */
int gmgo_level_init(gmgo_level *lv, size_t N, double length, double alpha, int level)
{
    size_t w = N, s = 1;
    for (int i = 0; i < level; i++) { w = (w + 1) / 2; s *= 2; }
    double m_h = length / (double)(N - 1);
    double h = m_h * (double)s;
    double k = h * h;
    lv->N = N; lv->w = w; lv->s = s;
    lv->diag = 4. * alpha / k;
    lv->off = -alpha / k;
    /* the reference silently requires (N-1) % 2^level == 0 ("Grid constraint") */
    if (N < 3 || w < 2 || (N - 1) % s != 0 || (w - 1) * s != N - 1) return -1;
    return 0;
}

static inline size_t mask_of(const gmgo_level *lv, size_t I, size_t J)
{
    return lv->s * I * lv->N + lv->s * J;          /* domain.hpp:78-80 */
}
static inline int on_boundary(const gmgo_level *lv, size_t I, size_t J)
{
    /* domain.cpp:20-23 evaluated on the fine index mask(i): fine (s*I, s*J) */
    size_t fi = lv->s * I, fj = lv->s * J, e = lv->N - 1;
    return fi == 0 || fj == 0 || fi == e || fj == e;
}

/* include/linear_system.hpp:85-92 + src/utilities.cpp:138-147: b = g on the boundary, f inside,
 * sampled at x = j*h, y = W - i*h (domain.hpp:68). test in {0,1,2}; anything else -> 0. */
static double test_f(int t, double x, double y)
{
    switch (t) {
    case 0: return 1.;
    case 1: return -5.0 * exp(x) * exp(-2.0 * y);
    case 2: { double r = sqrt(x * x + y * y);
              return r != 0.0 ? -30. * (cos(30. * r) / r - 30. * sin(30. * r)) : 0.0; }
    default: return 1.;
    }
}
static double test_g(int t, double x, double y)
{
    switch (t) {
    case 0: return 0.;
    case 1: return exp(x) * exp(-2.0 * y);
    case 2: return sin(30. * sqrt(x * x + y * y));
    default: return 0.;
    }
}
void gmgo_rhs(size_t N, double length, int test, double *b)
{
    if (test < 0 || test > 2) test = 0;            /* utilities.cpp:150-154 default pair */
    double m_h = length / (double)(N - 1);
    for (size_t i = 0; i < N; i++)
        for (size_t j = 0; j < N; j++) {
            double x = (double)j * m_h, y = length - (double)i * m_h;
            int bd = (i == 0 || j == 0 || i == N - 1 || j == N - 1);
            b[i * N + j] = bd ? test_g(test, x, y) : test_f(test, x, y);
        }
}

/* include/solvers.hpp:33-48 -- lexicographic Gauss-Seidel on one level, in place.
 * neighbour order up, left, right, down (domain.cpp:36-38 with the centre skipped). */
void gmgo_gs_sweep(const gmgo_level *lv, double *sol, const double *b)
{
    const size_t w = lv->w;
    for (size_t I = 0; I < w; I++)
        for (size_t J = 0; J < w; J++) {
            size_t idx = mask_of(lv, I, J);
            if (on_boundary(lv, I, J)) {
                sol[idx] = (b[idx] - 0.) / 1.;
            } else {
                double sum = 0;
                sum += lv->off * sol[mask_of(lv, I - 1, J)];
                sum += lv->off * sol[mask_of(lv, I, J - 1)];
                sum += lv->off * sol[mask_of(lv, I, J + 1)];
                sum += lv->off * sol[mask_of(lv, I + 1, J)];
                sol[idx] = (b[idx] - sum) / lv->diag;
            }
        }
}

/* include/solvers.hpp:64-83 -- Jacobi (omega = 1): every level point is computed from `sol`
 * into `temp`, then the two vectors are swapped.  Here the swap is realised by copying the
 * level's points back, which is what an observer restricted to this level's points sees;
 * the off-level entries that the reference's swap exchanges are never read before being
 * overwritten. */
void gmgo_jacobi_sweep(const gmgo_level *lv, double *sol, const double *b, double *temp)
{
    const size_t w = lv->w;
    for (size_t I = 0; I < w; I++)
        for (size_t J = 0; J < w; J++) {
            size_t idx = mask_of(lv, I, J);
            if (on_boundary(lv, I, J)) {
                temp[idx] = (b[idx] - 0.) / 1.;
            } else {
                double sum = 0;
                sum += lv->off * sol[mask_of(lv, I - 1, J)];
                sum += lv->off * sol[mask_of(lv, I, J - 1)];
                sum += lv->off * sol[mask_of(lv, I, J + 1)];
                sum += lv->off * sol[mask_of(lv, I + 1, J)];
                temp[idx] = (b[idx] - sum) / lv->diag;
            }
        }
    for (size_t I = 0; I < w; I++)
        for (size_t J = 0; J < w; J++) {
            size_t idx = mask_of(lv, I, J);
            sol[idx] = temp[idx];
        }
}

/* include/solvers.hpp:257-296 -- r = b - A u on one level; boundary rows r = b - 1*u;
 * interior: sum over up,left,centre,right,down in that order.  res may be NULL (norm only).
 * Returns sum r^2 accumulated in row-major order (the serial build's order). */
double gmgo_residual(const gmgo_level *lv, const double *sol, const double *b, double *res)
{
    const size_t w = lv->w;
    double norm = 0.;
    for (size_t I = 0; I < w; I++)
        for (size_t J = 0; J < w; J++) {
            size_t idx = mask_of(lv, I, J);
            double sum = 0;
            if (on_boundary(lv, I, J)) {
                sum = 1. * sol[idx];
            } else {
                sum += lv->off * sol[mask_of(lv, I - 1, J)];
                sum += lv->off * sol[mask_of(lv, I, J - 1)];
                sum += lv->diag * sol[idx];
                sum += lv->off * sol[mask_of(lv, I, J + 1)];
                sum += lv->off * sol[mask_of(lv, I + 1, J)];
            }
            double r = b[idx] - sum;
            if (res) res[idx] = r;
            norm += r * r;
        }
    return norm;
}

/* include/solvers.hpp:230-254 -- sum of b^2 over the level's points. */
double gmgo_sumsq(const gmgo_level *lv, const double *b)
{
    double k = 0;
    for (size_t I = 0; I < lv->w; I++)
        for (size_t J = 0; J < lv->w; J++) {
            double v = b[mask_of(lv, I, J)];
            k += v * v;
        }
    return k;
}

/* src/multigrid.cpp:3-27 -- in-place bilinear prolongation from level `c` (coarse) to level
 * `f` = c-1: (1) vertical midpoints below every coarse node not in the last coarse row,
 * (2) horizontal midpoints on every row of the finer level. */
void gmgo_prolong(const gmgo_level *c, const gmgo_level *f, double *vec)
{
    for (size_t I = 0; I + 1 < c->w; I++)
        for (size_t J = 0; J < c->w; J++) {
            size_t i1 = mask_of(c, I, J), i2 = mask_of(c, I + 1, J);
            size_t i3 = (i1 + i2) / 2;
            vec[i3] = 0.5 * (vec[i1] + vec[i2]);
        }
    for (size_t I = 0; I < f->w; I++)
        for (size_t J = 0; J + 1 < f->w; J += 2)
            vec[mask_of(f, I, J + 1)] = 0.5 * (vec[mask_of(f, I, J)] + vec[mask_of(f, I, J + 2)]);
}

/* ---- reordered smoother used by the CUDA fast path (NOT in the reference) -------------------
 * Red-black Gauss-Seidel: colour (I+J)%2==0 first, then (I+J)%2==1, same per-point formula as
 * solvers.hpp:33-48.  Boundary points are assigned u=b in the pass of their own colour.
 * It exists so the CUDA red-black kernels can be checked point for point; the comparison
 * against the reference's lexicographic ordering is on converged solutions (tests/). */
void gmgo_rbgs_sweep(const gmgo_level *lv, double *sol, const double *b)
{
    const size_t w = lv->w;
    for (int colour = 0; colour < 2; colour++)
        for (size_t I = 0; I < w; I++)
            for (size_t J = 0; J < w; J++) {
                if (((I + J) & 1) != (size_t)colour) continue;
                size_t idx = mask_of(lv, I, J);
                if (on_boundary(lv, I, J)) {
                    sol[idx] = (b[idx] - 0.) / 1.;
                } else {
                    double sum = 0;
                    sum += lv->off * sol[mask_of(lv, I - 1, J)];
                    sum += lv->off * sol[mask_of(lv, I, J - 1)];
                    sum += lv->off * sol[mask_of(lv, I, J + 1)];
                    sum += lv->off * sol[mask_of(lv, I + 1, J)];
                    sol[idx] = (b[idx] - sum) / lv->diag;
                }
            }
}

/* smoother ids follow include/utilities.hpp:9-14 (0 GS, 1 Jacobi); 3 = red-black GS (ours). */
enum { GMGO_GS = 0, GMGO_JACOBI = 1, GMGO_BICGSTAB = 2, GMGO_RBGS = 3 };

static void sweep(int kind, const gmgo_level *lv, double *sol, const double *b, double *temp)
{
    if (kind == GMGO_JACOBI) gmgo_jacobi_sweep(lv, sol, b, temp);
    else if (kind == GMGO_RBGS) gmgo_rbgs_sweep(lv, sol, b);
    else gmgo_gs_sweep(lv, sol, b);
}

typedef struct {
    int L;
    int kind;             /* smoother used inside the cycle */
    int nu;               /* include/multigrid.hpp:105 (5) */
    size_t coarse_maxit;  /* include/multigrid.hpp:123 (2000) */
    double coarse_tol;    /* include/multigrid.hpp:123 (1e-1) */
    gmgo_level lv[32];
    double *res, *err, *temp;   /* multigrid.hpp:93-94, solvers.hpp:58 */
    size_t n_fine;
    /* restriction of the fine residual to the coarse levels:
     *   0 = the reference's: injection, every level reads res through its mask (solvers.hpp:35,46,69,80)
     *   1 = half injection  (rhs_l = 0.5 * res at the level's points)          -- ours, for red-black GS
     *   2 = full weighting, cascaded level by level ([1 2 1;2 4 2;1 2 1]/16)   -- ours, for red-black GS
     * Modes 1,2 keep one fine-sized strided array per level (rl[l]); rl[0] aliases res. */
    int restrict_mode;
    double *rl[32];
    /* reporting */
    double last_coarse_relres;  /* value printed by multigrid.hpp:131 */
    long last_coarse_iters;
} gmgo_cycle_state;

gmgo_cycle_state *gmgo_cycle_create(size_t N, double length, double alpha, int L, int kind)
{
    if (L < 1 || L > 31) return NULL;
    gmgo_cycle_state *st = (gmgo_cycle_state *)calloc(1, sizeof(*st));
    st->L = L; st->kind = kind; st->nu = 5; st->coarse_maxit = 2000; st->coarse_tol = 1.e-1;
    for (int l = 0; l < L; l++)
        if (gmgo_level_init(&st->lv[l], N, length, alpha, l) != 0) { free(st); return NULL; }
    st->n_fine = N * N;
    st->res = (double *)calloc(st->n_fine, sizeof(double));
    st->err = (double *)calloc(st->n_fine, sizeof(double));
    st->temp = (double *)calloc(st->n_fine, sizeof(double));
    st->restrict_mode = 0;
    for (int l = 0; l < L; l++) st->rl[l] = st->res;
    return st;
}
void gmgo_cycle_destroy(gmgo_cycle_state *st)
{
    if (!st) return;
    for (int l = 1; l < st->L; l++) if (st->rl[l] != st->res) free(st->rl[l]);
    free(st->res); free(st->err); free(st->temp); free(st);
}
void gmgo_cycle_set_restriction(gmgo_cycle_state *st, int mode)
{
    st->restrict_mode = mode;
    for (int l = 1; l < st->L; l++) {
        if (st->rl[l] != st->res) free(st->rl[l]);
        st->rl[l] = mode ? (double *)calloc(st->n_fine, sizeof(double)) : st->res;
    }
}

/* ours (not in the reference): restriction of level f = c-1 rhs to level c, strided layout. */
static void restrict_level(int mode, const gmgo_level *f, const gmgo_level *c,
                           const double *rf, double *rc)
{
    for (size_t I = 0; I < c->w; I++)
        for (size_t J = 0; J < c->w; J++) {
            size_t ic = mask_of(c, I, J);
            size_t i = 2 * I, j = 2 * J;
            if (on_boundary(c, I, J)) { rc[ic] = rf[mask_of(f, i, j)]; continue; }
            if (mode == 1) { rc[ic] = 0.5 * rf[mask_of(f, i, j)]; continue; }
            double edge = rf[mask_of(f, i - 1, j)] + rf[mask_of(f, i, j - 1)]
                        + rf[mask_of(f, i, j + 1)] + rf[mask_of(f, i + 1, j)];
            double corner = rf[mask_of(f, i - 1, j - 1)] + rf[mask_of(f, i - 1, j + 1)]
                          + rf[mask_of(f, i + 1, j - 1)] + rf[mask_of(f, i + 1, j + 1)];
            rc[ic] = 0.25 * rf[mask_of(f, i, j)] + 0.125 * edge + 0.0625 * corner;
        }
}
void gmgo_cycle_set_params(gmgo_cycle_state *st, int nu, long coarse_maxit, double coarse_tol)
{
    st->nu = nu; st->coarse_maxit = (size_t)coarse_maxit; st->coarse_tol = coarse_tol;
}
double gmgo_cycle_last_coarse_relres(const gmgo_cycle_state *st) { return st->last_coarse_relres; }
long gmgo_cycle_last_coarse_iters(const gmgo_cycle_state *st) { return st->last_coarse_iters; }
const double *gmgo_cycle_res(const gmgo_cycle_state *st) { return st->res; }

/* include/multigrid.hpp:126-145 -- one sawtooth cycle applied to sol (rhs b on the fine grid). */
void gmgo_cycle_apply(gmgo_cycle_state *st, double *sol, const double *b)
{
    const gmgo_level *fine = &st->lv[0], *coarse = &st->lv[st->L - 1];
    /* :127  sol * RES  (3-argument Residual: stores into res) */
    gmgo_residual(fine, sol, b, st->res);
    if (st->restrict_mode) {
        if (st->restrict_mode == 1)      /* half injection always samples the FINE residual */
            for (int l = 1; l < st->L; l++) {
                gmgo_level f0 = st->lv[0]; (void)f0;
                const gmgo_level *c = &st->lv[l];
                for (size_t I = 0; I < c->w; I++)
                    for (size_t J = 0; J < c->w; J++) {
                        size_t ic = mask_of(c, I, J);
                        st->rl[l][ic] = on_boundary(c, I, J) ? st->res[ic] : 0.5 * st->res[ic];
                    }
            }
        else
            for (int l = 1; l < st->L; l++)
                restrict_level(2, &st->lv[l - 1], &st->lv[l], st->rl[l - 1], st->rl[l]);
    }
    const double *rc = st->rl[st->L - 1];
    /* :128  COARSE_RES->refresh_normalization_constant()  (solvers.hpp:244-254) */
    double norm_of_b = gmgo_sumsq(coarse, rc);
    /* :130  err * COARSE_SOLVER * COARSE_RES   (Solver::Solve, solvers.hpp:324-342) */
    size_t counter = st->coarse_maxit;
    long its = 0;
    double norm = gmgo_residual(coarse, st->err, rc, NULL);
    while (sqrt(norm / norm_of_b) > st->coarse_tol) {
        if (counter > 0) {
            sweep(st->kind, coarse, st->err, rc, st->temp);
            counter -= 1; its++;
            norm = gmgo_residual(coarse, st->err, rc, NULL);
        } else break;
    }
    norm = gmgo_residual(coarse, st->err, rc, NULL);
    st->last_coarse_relres = sqrt(norm / norm_of_b);
    st->last_coarse_iters = its;
    /* :134-139 upward leg: prolongate level j -> j-1, nu sweeps on level j-1 with rhs res */
    for (int j = st->L - 1; j > 0; --j) {
        gmgo_prolong(&st->lv[j], &st->lv[j - 1], st->err);
        for (int i = 0; i < st->nu; i++)
            sweep(st->kind, &st->lv[j - 1], st->err, st->rl[j - 1], st->temp);
    }
    /* :141-144 */
    for (size_t j = 0; j < st->n_fine; j++) { sol[j] += st->err[j]; st->err[j] = 0; }
}

/* src/main.cpp:73-116 -- the driver loop.  presmoother = lexicographic GS in the reference
 * (main.cpp:62,85,96,107) whatever -smt says; `pre_kind` lets the red-black variant be run
 * through the same loop.  -smt 2 runs the Jacobi cycle (main.cpp:103-106).
 * hist must hold maxiter+1 doubles.  Returns the number of history entries. */
int gmgo_solve_ex(size_t N, double length, double alpha, int L, int smoother, int pre_kind,
                  int restrict_mode, int nu, long coarse_maxit, double coarse_tol,
                  const double *b, double *u, double tol, int maxiter, double *hist,
                  double *coarse_relres_hist, long *coarse_iters_hist);
int gmgo_solve(size_t N, double length, double alpha, int L, int smoother, int pre_kind,
               const double *b, double *u, double tol, int maxiter, double *hist,
               double *coarse_relres_hist, long *coarse_iters_hist)
{
    return gmgo_solve_ex(N, length, alpha, L, smoother, pre_kind, 0, 5, 2000, 1.e-1, b, u, tol,
                         maxiter, hist, coarse_relres_hist, coarse_iters_hist);
}
int gmgo_solve_ex(size_t N, double length, double alpha, int L, int smoother, int pre_kind,
                  int restrict_mode, int nu, long coarse_maxit, double coarse_tol,
                  const double *b, double *u, double tol, int maxiter, double *hist,
                  double *coarse_relres_hist, long *coarse_iters_hist)
{
    int kind = smoother;
    if (kind == GMGO_BICGSTAB) kind = GMGO_JACOBI;
    gmgo_cycle_state *st = gmgo_cycle_create(N, length, alpha, L, kind);
    if (!st) return -1;
    gmgo_cycle_set_params(st, nu, coarse_maxit, coarse_tol);
    gmgo_cycle_set_restriction(st, restrict_mode);
    const gmgo_level *fine = &st->lv[0];
    double norm_of_b = gmgo_sumsq(fine, b);                      /* solvers.hpp:237-242 */
    double *res = (double *)calloc(N * N, sizeof(double));       /* main.cpp:50 */
    int n = 0;
    hist[n++] = sqrt(gmgo_residual(fine, u, b, res) / norm_of_b); /* main.cpp:73-74 */
    for (int i = 0; i < maxiter; i++) {
        sweep(pre_kind, fine, u, b, st->temp);
        sweep(pre_kind, fine, u, b, st->temp);
        gmgo_cycle_apply(st, u, b);
        if (coarse_relres_hist) coarse_relres_hist[i] = st->last_coarse_relres;
        if (coarse_iters_hist) coarse_iters_hist[i] = st->last_coarse_iters;
        hist[n++] = sqrt(gmgo_residual(fine, u, b, res) / norm_of_b);
        if (hist[n - 1] <= tol) break;
    }
    free(res);
    gmgo_cycle_destroy(st);
    return n;
}
