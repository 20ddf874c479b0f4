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

/* =====================================================================================================
 * Textbook cycle options, full multigrid and Krylov solvers.
 *
 * NOT IN THE REFERENCE: its only cycle is the sawtooth above and its BiCGSTAB class is never executed
 * (src/main.cpp:103-106).  What follows is the CPU statement of the algorithms libmgb200 adds
 * (include/mgb200.h: mgb_gmg_config.cycle_type / nu_pre / fmg, mgb_gmg_fmg, mgb_gmg_krylov), written
 * with the reference's per-point formulas (the functions above) so that the CUDA path can be checked
 * bit for bit (sweeps, residuals, transfers) or to rounding (dot products).  Every level keeps its own
 * fine-sized strided arrays e_l, r_l: a correction-scheme cycle needs e_l and e_{l+1} at the same time,
 * which the reference's single shared array cannot hold.
 * ===================================================================================================== */
typedef struct {
    int L, kind, nu1, nu2, restrict_mode, cycle, bottom;
    long coarse_maxit;
    double coarse_tol;
    double omega;              /* weight of the Jacobi smoother (mgb_gmg_config.jacobi_omega); 1 = the reference's */
    gmgo_level lv[32];
    double *e[32], *r[32], *d, *temp;
    size_t n_fine;
} gmgo_tb;

/* weighted Jacobi (north_star; the reference is omega = 1, solvers.hpp:64-83): u <- u + omega (u_jacobi - u) inside */
static void wjacobi_sweep(const gmgo_level *lv, double *sol, const double *b, double *temp, double omega)
{
    const size_t w = lv->w;
    for (size_t I = 0; I < w; I++)
        for (size_t J = 0; J < w; J++) {
            size_t idx = mask_of(lv, I, J);
            if (on_boundary(lv, I, J)) { temp[idx] = b[idx]; continue; }
            double sum = 0;
            sum += lv->off * sol[mask_of(lv, I - 1, J)];
            sum += lv->off * sol[mask_of(lv, I, J - 1)];
            sum += lv->off * sol[mask_of(lv, I, J + 1)];
            sum += lv->off * sol[mask_of(lv, I + 1, J)];
            double o = (b[idx] - sum) / lv->diag;
            temp[idx] = sol[idx] + omega * (o - sol[idx]);
        }
    for (size_t I = 0; I < w; I++)
        for (size_t J = 0; J < w; J++) { size_t idx = mask_of(lv, I, J); sol[idx] = temp[idx]; }
}
static void tb_sweep(const gmgo_tb *st, const gmgo_level *lv, double *sol, const double *b, double *temp)
{
    if (st->kind == GMGO_JACOBI && st->omega != 1.0) wjacobi_sweep(lv, sol, b, temp, st->omega);
    else sweep(st->kind, lv, sol, b, temp);
}
void gmgo_tb_set_omega(gmgo_tb *st, double omega);

enum { GMGO_SAWTOOTH = 0, GMGO_V = 1, GMGO_W = 2, GMGO_F = 3 };

/* bottom = first level of the coarse solver: levels bottom..L-1 are "solved" by one sawtooth pass
 * (multigrid.hpp:128-139 applied to r_bottom), which is what the persistent tail kernel does. */
gmgo_tb *gmgo_tb_create(size_t N, double length, double alpha, int L, int kind, int nu1, int nu2,
                        int restrict_mode, int cycle, int bottom, long coarse_maxit, double coarse_tol)
{
    if (L < 1 || L > 31 || bottom < 0 || bottom >= L) return NULL;
    gmgo_tb *st = (gmgo_tb *)calloc(1, sizeof(*st));
    st->L = L; st->kind = kind; st->nu1 = nu1; st->nu2 = nu2; st->restrict_mode = restrict_mode;
    st->cycle = cycle; st->bottom = bottom; st->coarse_maxit = coarse_maxit; st->coarse_tol = coarse_tol;
    st->n_fine = N * N;
    st->omega = 1.0;
    for (int l = 0; l < L; l++) {
        if (gmgo_level_init(&st->lv[l], N, length, alpha, l) != 0) { free(st); return NULL; }
        st->e[l] = (double *)calloc(st->n_fine, sizeof(double));
        st->r[l] = (double *)calloc(st->n_fine, sizeof(double));
    }
    st->d = (double *)calloc(st->n_fine, sizeof(double));
    st->temp = (double *)calloc(st->n_fine, sizeof(double));
    return st;
}
void gmgo_tb_destroy(gmgo_tb *st)
{
    if (!st) return;
    for (int l = 0; l < st->L; l++) { free(st->e[l]); free(st->r[l]); }
    free(st->d); free(st->temp); free(st);
}
void gmgo_tb_set_omega(gmgo_tb *st, double omega) { st->omega = omega > 0. ? omega : 1.0; }
double *gmgo_tb_e(gmgo_tb *st, int l) { return st->e[l]; }
double *gmgo_tb_r(gmgo_tb *st, int l) { return st->r[l]; }

static void level_zero(const gmgo_level *lv, double *v)
{
    for (size_t I = 0; I < lv->w; I++)
        for (size_t J = 0; J < lv->w; J++) v[mask_of(lv, I, J)] = 0.;
}

/* out (level f = c-1) = bilinear interpolation of ec (level c), the two-stage order of multigrid.cpp:3-27 */
static void prolong_into(const gmgo_level *c, const gmgo_level *f, const double *ec, double *out)
{
    for (size_t I = 0; I < c->w; I++)
        for (size_t J = 0; J < c->w; J++) out[mask_of(c, I, J)] = ec[mask_of(c, I, J)];
    gmgo_prolong(c, f, out);
}

/* restriction used inside the sawtooth pass of the coarse solver and by the cascade of the FMG pass:
 * full weighting level by level, or injection (half injection scales only the step to level 1) */
static void restrict_sawtooth(int mode, int l, const gmgo_level *f, const gmgo_level *c, const double *rf, double *rc)
{
    if (mode == 2) { restrict_level(2, f, c, rf, rc); return; }
    for (size_t I = 0; I < c->w; I++)
        for (size_t J = 0; J < c->w; J++) {
            size_t ic = mask_of(c, I, J);
            double scale = (mode == 1 && l == 1) ? 0.5 : 1.0;
            rc[ic] = on_boundary(c, I, J) ? rf[ic] : scale * rf[ic];
        }
}

/* the coarse solver: one sawtooth pass on levels bottom..L-1 from r[bottom]; leaves e[bottom] */
static void tb_tail(gmgo_tb *st, int restrict_below)
{
    const int L = st->L, b = st->bottom;
    if (restrict_below)
        for (int l = b + 1; l < L; l++)
            restrict_sawtooth(st->restrict_mode, l, &st->lv[l - 1], &st->lv[l], st->r[l - 1], st->r[l]);
    const gmgo_level *C = &st->lv[L - 1];
    double *ec = st->e[L - 1];
    const double *rc = st->r[L - 1];
    level_zero(C, ec);
    double nb = gmgo_sumsq(C, rc);
    long its = 0;
    double norm = gmgo_residual(C, ec, rc, NULL);
    while (sqrt(norm / nb) > st->coarse_tol && its < st->coarse_maxit) {
        tb_sweep(st, C, ec, rc, st->temp);
        its++;
        norm = gmgo_residual(C, ec, rc, NULL);
    }
    for (int l = L - 1; l > b; --l) {
        prolong_into(&st->lv[l], &st->lv[l - 1], st->e[l], st->e[l - 1]);
        for (int i = 0; i < st->nu2; i++) tb_sweep(st, &st->lv[l - 1], st->e[l - 1], st->r[l - 1], st->temp);
    }
}

static void tb_cycle(gmgo_tb *st, int l, int type, int zero)
{
    const gmgo_level *X = &st->lv[l];
    if (l >= st->bottom) { tb_tail(st, 1); return; }
    if (zero) level_zero(X, st->e[l]);
    for (int i = 0; i < st->nu1; i++) tb_sweep(st, X, st->e[l], st->r[l], st->temp);
    gmgo_residual(X, st->e[l], st->r[l], st->d);
    {   /* r_{l+1} = R d: full weighting, or injection (half injection halves on EVERY level here) */
        const gmgo_level *C = &st->lv[l + 1];
        if (st->restrict_mode == 2) restrict_level(2, X, C, st->d, st->r[l + 1]);
        else
            for (size_t I = 0; I < C->w; I++)
                for (size_t J = 0; J < C->w; J++) {
                    size_t ic = mask_of(C, I, J);
                    double scale = st->restrict_mode == 1 ? 0.5 : 1.0;
                    st->r[l + 1][ic] = on_boundary(C, I, J) ? st->d[ic] : scale * st->d[ic];
                }
    }
    if (type == GMGO_F) {
        tb_cycle(st, l + 1, GMGO_F, 1);
        if (l + 1 < st->bottom) tb_cycle(st, l + 1, GMGO_V, 0);
    } else {
        int visits = (type == GMGO_W && l + 1 < st->bottom) ? 2 : 1;
        for (int v = 0; v < visits; v++) tb_cycle(st, l + 1, type, v == 0);
    }
    prolong_into(&st->lv[l + 1], X, st->e[l + 1], st->d);
    for (size_t I = 0; I < X->w; I++)
        for (size_t J = 0; J < X->w; J++) { size_t i = mask_of(X, I, J); st->e[l][i] = st->e[l][i] + st->d[i]; }
    for (int i = 0; i < st->nu2; i++) tb_sweep(st, X, st->e[l], st->r[l], st->temp);
}

/* e_0 ~= A^-1 r_0 with the configured cycle (r_0 = st->r[0] must be set) */
void gmgo_tb_apply(gmgo_tb *st)
{
    if (st->cycle == GMGO_SAWTOOTH) {
        /* the whole hierarchy as one sawtooth pass */
        int b = st->bottom;
        st->bottom = 0;
        tb_tail(st, 1);
        st->bottom = b;
    } else tb_cycle(st, 0, st->cycle, 1);
}

/* one driver iteration with a textbook cycle: n_pre sweeps on u, r_0 = b - A u, cycle, u += e_0 */
void gmgo_tb_iteration(gmgo_tb *st, double *u, const double *b, int pre_kind, int n_pre)
{
    const gmgo_level *F = &st->lv[0];
    for (int i = 0; i < n_pre; i++) sweep(pre_kind, F, u, b, st->temp);
    gmgo_residual(F, u, b, st->r[0]);
    gmgo_tb_apply(st);
    for (size_t j = 0; j < st->n_fine; j++) u[j] += st->e[0][j];
}

/* one full-multigrid pass on the residual equation (mgb_gmg_fmg) */
void gmgo_tb_fmg(gmgo_tb *st, double *u, const double *b)
{
    const gmgo_level *F = &st->lv[0];
    gmgo_residual(F, u, b, st->r[0]);
    for (int l = 1; l <= st->bottom; l++)
        restrict_sawtooth(st->restrict_mode, l, &st->lv[l - 1], &st->lv[l], st->r[l - 1], st->r[l]);
    tb_tail(st, 1);
    for (int l = st->bottom - 1; l >= 0; --l) {
        prolong_into(&st->lv[l + 1], &st->lv[l], st->e[l + 1], st->e[l]);
        tb_cycle(st, l, GMGO_V, 0);
    }
    for (size_t j = 0; j < st->n_fine; j++) u[j] += st->e[0][j];
}

/* ---- Krylov solvers on the fine level, optionally right-preconditioned by one cycle (mgb_gmg_krylov) ---- */
static double vdot(size_t n, const double *a, const double *b) { double s = 0; for (size_t i = 0; i < n; i++) s += a[i] * b[i]; return s; }
/* q = A p on the fine level (identity rows on the boundary), formula of solvers.hpp:278-294 with b = 0, negated */
static void apply_A(const gmgo_level *F, const double *p, double *q)
{
    const size_t w = F->w;
    for (size_t I = 0; I < w; I++)
        for (size_t J = 0; J < w; J++) {
            size_t idx = mask_of(F, I, J);
            if (on_boundary(F, I, J)) { q[idx] = p[idx]; continue; }
            double sum = 0;
            sum += F->off * p[mask_of(F, I - 1, J)];
            sum += F->off * p[mask_of(F, I, J - 1)];
            sum += F->diag * p[idx];
            sum += F->off * p[mask_of(F, I, J + 1)];
            sum += F->off * p[mask_of(F, I + 1, J)];
            q[idx] = -(0. - sum);
        }
}
static const double *tb_precond(gmgo_tb *st, int precond, const double *src)
{
    if (!precond) return src;
    memcpy(st->r[0], src, st->n_fine * sizeof(double));
    if (st->L == 1) tb_tail(st, 0); else gmgo_tb_apply(st);
    return st->e[0];
}

/* method 0 = CG, 1 = BiCGSTAB; precond 0 = none, 1 = one cycle.  hist holds maxit+1 doubles; returns the entries written */
int gmgo_tb_krylov(gmgo_tb *st, int method, int precond, const double *b, double *u, double tol, int maxit, double *hist)
{
    const gmgo_level *F = &st->lv[0];
    const size_t n = st->n_fine, w = F->w;
    double *r = (double *)calloc(n, sizeof(double)), *p = (double *)calloc(n, sizeof(double));
    double *q = (double *)calloc(n, sizeof(double)), *rh = (double *)calloc(n, sizeof(double));
    double *s = (double *)calloc(n, sizeof(double)), *t = (double *)calloc(n, sizeof(double));
    int k = 0;
    for (size_t I = 0; I < w; I++)
        for (size_t J = 0; J < w; J++)
            if (on_boundary(F, I, J)) u[mask_of(F, I, J)] = b[mask_of(F, I, J)];
    const double nf = gmgo_sumsq(F, b);
    double rr = gmgo_residual(F, u, b, r);
    hist[k++] = sqrt(rr / nf);
    if (hist[0] > tol && maxit > 0) {
        if (method == 0) {
            const double *z = tb_precond(st, precond, r);
            double rho = vdot(n, r, z);
            memcpy(p, z, n * sizeof(double));
            for (int it = 0; it < maxit; it++) {
                apply_A(F, p, q);
                double alpha = rho / vdot(n, p, q);
                for (size_t i = 0; i < n; i++) u[i] += alpha * p[i];
                for (size_t i = 0; i < n; i++) r[i] -= alpha * q[i];
                hist[k++] = sqrt(vdot(n, r, r) / nf);
                if (hist[k - 1] <= tol) break;
                z = tb_precond(st, precond, r);
                double rho_new = vdot(n, r, z), beta = rho_new / rho;
                for (size_t i = 0; i < n; i++) p[i] = z[i] + beta * p[i];
                rho = rho_new;
            }
        } else {
            memcpy(rh, r, n * sizeof(double));
            memcpy(p, r, n * sizeof(double));
            double rho = rr, alpha, omega;
            for (int it = 0; it < maxit; it++) {
                const double *y = tb_precond(st, precond, p);
                apply_A(F, y, q);                                   /* v */
                alpha = rho / vdot(n, rh, q);
                for (size_t i = 0; i < n; i++) u[i] += alpha * y[i];
                for (size_t i = 0; i < n; i++) s[i] = r[i] - alpha * q[i];
                double ss = vdot(n, s, s);
                if (sqrt(ss / nf) <= tol) { hist[k++] = sqrt(ss / nf); break; }
                const double *z = tb_precond(st, precond, s);
                apply_A(F, z, t);
                omega = vdot(n, t, s) / vdot(n, t, t);
                for (size_t i = 0; i < n; i++) u[i] += omega * z[i];
                for (size_t i = 0; i < n; i++) r[i] = s[i] - omega * t[i];
                hist[k++] = sqrt(vdot(n, r, r) / nf);
                if (hist[k - 1] <= tol) break;
                double rho_new = vdot(n, rh, r);
                double beta = (rho_new / rho) * (alpha / omega);
                for (size_t i = 0; i < n; i++) p[i] = r[i] + beta * (p[i] - omega * q[i]);
                rho = rho_new;
            }
        }
    }
    free(r); free(p); free(q); free(rh); free(s); free(t);
    return k;
}
