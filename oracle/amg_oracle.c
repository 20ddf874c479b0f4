/*
 * amg_oracle.c -- CPU restatement of the reference's AMG setup and one-pass cycle on CSR.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md): the checker for the CUDA AMG path.
 *
 * Parity status: PINNED against the reference's own classes compiled from /root/reference
 * (oracle/_ref/libamgref.so) with the random start index of the C/F splitting injected:
 * hierarchy (every A_l, P_l, rhs_l) and the solution after AMG::apply_AMG() are compared bit
 * for bit in tests/test_oracle_amg.py, live when _ref is present and against stored outputs
 * (tests/golden/amg_*.npz) otherwise.  The reference holds no AMG golden vectors of its own.
 *
 * The reference's setup is O(N*Nc) through std::map (AMG/include/AMG.hpp:303-369); this file
 * restates the SAME arithmetic, term for term and in the same order, with O(nnz) sparse loops:
 * every term the dense loops add on top is an exact 0.0 and cannot change a sum.
 *
 * All citations are relative to /root/reference/AMG/.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t idx;

typedef struct {
    idx n_rows, n_cols, nnz;
    idx *ptr, *col;
    double *val;
} csr;

static csr csr_alloc(idx n_rows, idx n_cols, idx nnz)
{
    csr m;
    m.n_rows = n_rows; m.n_cols = n_cols; m.nnz = nnz;
    m.ptr = (idx *)calloc((size_t)n_rows + 1, sizeof(idx));
    m.col = (idx *)malloc(sizeof(idx) * (size_t)(nnz > 0 ? nnz : 1));
    m.val = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    return m;
}
static void csr_free(csr *m) { free(m->ptr); free(m->col); free(m->val); m->ptr = m->col = NULL; m->val = NULL; }

/* CSRMatrix::coeff (src/CSRMatrix.cpp:24-40): linear scan of the row, 0.0 when absent */
static double coeff(const csr *A, idx i, idx j)
{
    for (idx k = A->ptr[i]; k < A->ptr[i + 1]; k++)
        if (A->col[k] == j) return A->val[k];
    return 0.0;
}

/* ---- smoother / residual / transfers --------------------------------------------------------- */

/* Gauss_Seidel_iteration::apply_iteration_to_vec_no_mask (include/Utilities.hpp:44-58) */
void amgo_gs_sweep(idx n, const idx *ptr, const idx *col, const double *val, double *x, const double *b)
{
    for (idx i = 0; i < n; i++) {
        double sum = 0, aii = 0.0;
        int have = 0;
        for (idx k = ptr[i]; k < ptr[i + 1]; k++) {
            if (col[k] != i) sum += val[k] * x[col[k]];
            else if (!have) { aii = val[k]; have = 1; }
        }
        x[i] = (b[i] - sum) / aii;
    }
}

/* ..._with_mask (include/Utilities.hpp:60-75): x and b are indexed through component_mask */
void amgo_gs_sweep_masked(idx n, const idx *ptr, const idx *col, const double *val, const idx *mask,
                          double *x, const double *b)
{
    for (idx i = 0; i < n; i++) {
        idx mi = mask[i];
        double sum = 0, aii = 0.0;
        int have = 0;
        for (idx k = ptr[i]; k < ptr[i + 1]; k++) {
            if (col[k] != i) sum += val[k] * x[mask[col[k]]];
            else if (!have) { aii = val[k]; have = 1; }
        }
        x[mi] = (b[mi] - sum) / aii;
    }
}

/* AMG::compute_residual (src/AMG.cpp:256-275) / RestrictionOperator::compute_residual
 * (include/AMG.hpp:397-418): r = b - A x, returns ||r||_2 */
double amgo_residual(idx n, const idx *ptr, const idx *col, const double *val, const double *x,
                     const double *b, double *r)
{
    double norm = 0.0;
    for (idx i = 0; i < n; i++) {
        double Ax = 0.0;
        for (idx k = ptr[i]; k < ptr[i + 1]; k++) Ax += val[k] * x[col[k]];
        double ri = b[i] - Ax;
        if (r) r[i] = ri;
        norm += ri * ri;
    }
    return sqrt(norm);
}

/* AMG::apply_restriction_operator (src/AMG.cpp:50-74), also the rhs restriction of
 * AMG::initialization (src/AMG.cpp:100-109): y = P^T x, accumulated over fine rows ascending */
void amgo_restrict(idx n_f, idx n_c, const idx *ptr, const idx *col, const double *val, const double *xf, double *xc)
{
    for (idx m = 0; m < n_c; m++) xc[m] = 0.0;
    for (idx i = 0; i < n_f; i++)
        for (idx k = ptr[i]; k < ptr[i + 1]; k++) xc[col[k]] += val[k] * xf[i];
}

/* AMG::apply_prolungation_operator (src/AMG.cpp:218-232): x_f += P x_c, term by term */
void amgo_prolong_add(idx n_f, const idx *ptr, const idx *col, const double *val, const double *xc, double *xf)
{
    for (idx i = 0; i < n_f; i++)
        for (idx k = ptr[i]; k < ptr[i + 1]; k++) xf[i] += val[k] * xc[col[k]];
}

/* ---- setup ----------------------------------------------------------------------------------------- */
#define EPSILON 0.2 /* include/AMG.hpp:21 */

/* RestrictionOperator::strong_connections_in_row (include/AMG.hpp:105-130); out holds <= row length */
static idx strong_in_row(const csr *A, idx row, idx *out)
{
    double max_value = 0.0;
    for (idx k = A->ptr[row]; k < A->ptr[row + 1]; k++) {
        if (A->col[k] == row) continue;
        if (max_value < fabs(A->val[k])) max_value = fabs(A->val[k]);
    }
    idx n = 0;
    for (idx k = A->ptr[row]; k < A->ptr[row + 1]; k++) {
        if (A->col[k] == row) continue;
        if (fabs(A->val[k]) >= EPSILON * max_value) out[n++] = A->col[k];
    }
    return n;
}

/* RestrictionOperator::select_coarse_nodes (include/AMG.hpp:150-198).  mask byte: low 6 bits = count
 * of strong connections, top 2 bits set = FINE.  `start` replaces getRandomInit(n) (Utilities.cpp:30-40).
 * The "next index" scan of the reference picks the LARGEST i whose count is non-zero; counts only
 * ever go non-zero -> zero, so a pointer walking down from n-1 finds the same node in O(n) overall. */
static idx select_coarse(const csr *A, idx start, unsigned char *mask)
{
    const idx n = A->n_rows;
    idx *sptr = (idx *)malloc(sizeof(idx) * (size_t)(n + 1));
    idx *scol = (idx *)malloc(sizeof(idx) * (size_t)(A->nnz > 0 ? A->nnz : 1));
    sptr[0] = 0;
    for (idx i = 0; i < n; i++) {
        idx c = strong_in_row(A, i, scol + sptr[i]);
        sptr[i + 1] = sptr[i] + c;
        mask[i] = (unsigned char)c;                                   /* AMG.hpp:145 */
    }
    idx index = start, counter_fine = 0, top = n - 1;
    while (mask[index] & 0x3F) {
        mask[index] = 0;
        for (idx a = sptr[index]; a < sptr[index + 1]; a++) {
            idx c = scol[a];
            if (mask[c] & 0x3F) {
                mask[c] |= 0xC0;
                mask[c] &= 0xC0;
                counter_fine++;
                for (idx b = sptr[c]; b < sptr[c + 1]; b++) {
                    idx s2 = scol[b];
                    if (mask[s2] & 0x3F) mask[s2] += 2;
                }
            }
        }
        while (top >= 0 && !(mask[top] & 0x3F)) top--;                /* AMG.hpp:184-192 */
        if (top >= 0) index = top;
    }
    free(sptr); free(scol);
    return n - counter_fine;
}

/* CSRMatrix::copy_from drops exact zeros (src/CSRMatrix.cpp:13-14) */
static csr csr_from_rows(idx n_rows, idx n_cols, const idx *rptr, const idx *rcol, const double *rval)
{
    idx nnz = 0;
    for (idx k = 0; k < rptr[n_rows]; k++) if (rval[k] != 0) nnz++;
    csr m = csr_alloc(n_rows, n_cols, nnz);
    idx p = 0;
    for (idx i = 0; i < n_rows; i++) {
        m.ptr[i] = p;
        for (idx k = rptr[i]; k < rptr[i + 1]; k++)
            if (rval[k] != 0) { m.col[p] = rcol[k]; m.val[p] = rval[k]; p++; }
    }
    m.ptr[n_rows] = p;
    return m;
}

/* build_component_mask + build_prolongation_matrix (include/AMG.hpp:201-300) */
static csr build_P(const csr *A, const unsigned char *mask, idx nc, idx *rev)
{
    const idx n = A->n_rows;
    idx k = 0;
    for (idx i = 0; i < n; i++) { rev[i] = -1; if (!(mask[i] & 0xC0)) rev[i] = k++; }
    (void)nc;
    idx *rptr = (idx *)malloc(sizeof(idx) * (size_t)(n + 1));
    idx *rcol = (idx *)malloc(sizeof(idx) * (size_t)(A->nnz + n + 1));
    double *rval = (double *)malloc(sizeof(double) * (size_t)(A->nnz + n + 1));
    idx *strong = (idx *)malloc(sizeof(idx) * (size_t)(A->nnz > 0 ? A->nnz : 1));
    idx p = 0;
    for (idx i = 0; i < n; i++) {
        rptr[i] = p;
        if (!(mask[i] & 0xC0)) { rcol[p] = rev[i]; rval[p] = 1.0; p++; continue; }
        double alpha_num = 0.0;
        for (idx a = A->ptr[i]; a < A->ptr[i + 1]; a++) if (A->col[a] != i) alpha_num += A->val[a];
        idx ns = strong_in_row(A, i, strong);
        double alpha_denum = 0.0;
        for (idx a = 0; a < ns; a++) if (!(mask[strong[a]] & 0xC0)) alpha_denum += coeff(A, i, strong[a]);
        double alpha = alpha_num / alpha_denum;
        double sum = 0.0;
        for (idx a = 0; a < ns; a++) if (!(mask[strong[a]] & 0xC0)) sum += alpha * coeff(A, i, strong[a]);
        for (idx a = 0; a < ns; a++)
            if (!(mask[strong[a]] & 0xC0)) { rcol[p] = rev[strong[a]]; rval[p] = alpha * coeff(A, i, strong[a]) / (sum); p++; }
    }
    rptr[n] = p;
    csr P = csr_from_rows(n, k, rptr, rcol, rval);
    free(rptr); free(rcol); free(rval); free(strong);
    return P;
}

/* sparse row accumulator keyed by column, insertion order irrelevant (emitted sorted) */
typedef struct { idx *mark, *cols; double *acc; idx n; } spa;
static int cmp_idx(const void *a, const void *b) { idx x = *(const idx *)a, y = *(const idx *)b; return (x > y) - (x < y); }

/* build_coarse_matrix (include/AMG.hpp:303-369):
 *   PtA(i,j) = sum_{k in row j of A, ascending} A(j,k) * P(k,i)      (uses the symmetry of A, as the reference does)
 *   Ac(i,j)  = sum_{k in row i of PtA, ascending} PtA(i,k) * P(k,j)
 * exact zeros dropped when each product is compressed. */
static csr build_Ac(const csr *A, const csr *P)
{
    const idx n = A->n_rows, nc = P->n_cols;
    /* T(j, i) = PtA(i, j): computed row j by row j, then transposed */
    spa s; s.mark = (idx *)malloc(sizeof(idx) * (size_t)(nc > n ? nc : n)); s.cols = (idx *)malloc(sizeof(idx) * (size_t)(nc > n ? nc : n));
    s.acc = (double *)malloc(sizeof(double) * (size_t)(nc > n ? nc : n));
    for (idx i = 0; i < (nc > n ? nc : n); i++) s.mark[i] = -1;
    idx cap = 16 * (A->nnz + 1), tn = 0;
    idx *tptr = (idx *)malloc(sizeof(idx) * (size_t)(n + 1)), *tcol = (idx *)malloc(sizeof(idx) * (size_t)cap);
    double *tval = (double *)malloc(sizeof(double) * (size_t)cap);
    for (idx j = 0; j < n; j++) {
        tptr[j] = tn; s.n = 0;
        for (idx a = A->ptr[j]; a < A->ptr[j + 1]; a++) {
            idx k = A->col[a];
            for (idx b = P->ptr[k]; b < P->ptr[k + 1]; b++) {
                idx i = P->col[b];
                if (s.mark[i] != j) { s.mark[i] = j; s.cols[s.n++] = i; s.acc[i] = 0.0; }
                s.acc[i] += A->val[a] * P->val[b];
            }
        }
        qsort(s.cols, (size_t)s.n, sizeof(idx), cmp_idx);
        if (tn + s.n > cap) { cap = 2 * (tn + s.n); tcol = (idx *)realloc(tcol, sizeof(idx) * (size_t)cap); tval = (double *)realloc(tval, sizeof(double) * (size_t)cap); }
        for (idx q = 0; q < s.n; q++) if (s.acc[s.cols[q]] != 0) { tcol[tn] = s.cols[q]; tval[tn] = s.acc[s.cols[q]]; tn++; }
    }
    tptr[n] = tn;
    /* transpose T (n x nc) -> PtA (nc x n), columns ascending within each row */
    csr PtA = csr_alloc(nc, n, tn);
    for (idx q = 0; q < tn; q++) PtA.ptr[tcol[q] + 1]++;
    for (idx i = 0; i < nc; i++) PtA.ptr[i + 1] += PtA.ptr[i];
    idx *fill = (idx *)malloc(sizeof(idx) * (size_t)(nc + 1));
    memcpy(fill, PtA.ptr, sizeof(idx) * (size_t)(nc + 1));
    for (idx j = 0; j < n; j++)
        for (idx q = tptr[j]; q < tptr[j + 1]; q++) { idx i = tcol[q]; PtA.col[fill[i]] = j; PtA.val[fill[i]] = tval[q]; fill[i]++; }
    free(tptr); free(tcol); free(tval); free(fill);
    /* Ac = PtA * P */
    for (idx i = 0; i < (nc > n ? nc : n); i++) s.mark[i] = -1;
    cap = 16 * (A->nnz + 1);
    idx an = 0;
    idx *aptr = (idx *)malloc(sizeof(idx) * (size_t)(nc + 1)), *acol = (idx *)malloc(sizeof(idx) * (size_t)cap);
    double *aval = (double *)malloc(sizeof(double) * (size_t)cap);
    for (idx i = 0; i < nc; i++) {
        aptr[i] = an; s.n = 0;
        for (idx a = PtA.ptr[i]; a < PtA.ptr[i + 1]; a++) {
            idx k = PtA.col[a];
            for (idx b = P->ptr[k]; b < P->ptr[k + 1]; b++) {
                idx j = P->col[b];
                if (s.mark[j] != i) { s.mark[j] = i; s.cols[s.n++] = j; s.acc[j] = 0.0; }
                s.acc[j] += PtA.val[a] * P->val[b];
            }
        }
        qsort(s.cols, (size_t)s.n, sizeof(idx), cmp_idx);
        if (an + s.n > cap) { cap = 2 * (an + s.n); acol = (idx *)realloc(acol, sizeof(idx) * (size_t)cap); aval = (double *)realloc(aval, sizeof(double) * (size_t)cap); }
        for (idx q = 0; q < s.n; q++) { acol[an] = s.cols[q]; aval[an] = s.acc[s.cols[q]]; an++; }
    }
    aptr[nc] = an;
    csr Ac = csr_from_rows(nc, nc, aptr, acol, aval);
    free(aptr); free(acol); free(aval); free(s.mark); free(s.cols); free(s.acc);
    csr_free(&PtA);
    return Ac;
}

/* ---- hierarchy object ------------------------------------------------------------------------------ */
#define AMGO_MAX_LEVELS 16
typedef struct {
    int levels;
    csr A[AMGO_MAX_LEVELS];
    csr P[AMGO_MAX_LEVELS];          /* P[l]: level l+1 -> level l */
    double *rhs[AMGO_MAX_LEVELS];
    unsigned char *cf[AMGO_MAX_LEVELS];   /* C/F byte per node of level l (0 = coarse, 0xC0 = fine) */
} amgo_hier;

/* AMG ctor + AMG::initialization (include/AMG.hpp:33-41, src/AMG.cpp:76-120).
 * starts[l-1] is the start index used when coarsening level l-1 -> l (negative: n/2). */
amgo_hier *amgo_build(idx n, const idx *ptr, const idx *col, const double *val, const double *rhs,
                      int levels, const idx *starts)
{
    if (levels < 1 || levels > AMGO_MAX_LEVELS) return NULL;
    amgo_hier *h = (amgo_hier *)calloc(1, sizeof(*h));
    h->levels = levels;
    h->A[0] = csr_from_rows(n, n, ptr, col, val);
    h->rhs[0] = (double *)malloc(sizeof(double) * (size_t)n);
    memcpy(h->rhs[0], rhs, sizeof(double) * (size_t)n);
    for (int l = 1; l < levels; l++) {
        const csr *A = &h->A[l - 1];
        unsigned char *mask = (unsigned char *)calloc((size_t)A->n_rows + 1, 1);
        idx start = (starts && starts[l - 1] >= 0) ? starts[l - 1] : A->n_rows / 2;
        if (start >= A->n_rows) start = A->n_rows - 1;      /* the reference throws on start == n */
        idx nc = select_coarse(A, start, mask);
        idx *rev = (idx *)malloc(sizeof(idx) * (size_t)(A->n_rows + 1));
        h->P[l - 1] = build_P(A, mask, nc, rev);
        free(rev);
        h->cf[l - 1] = mask;
        h->rhs[l] = (double *)malloc(sizeof(double) * (size_t)(nc > 0 ? nc : 1));
        amgo_restrict(A->n_rows, nc, h->P[l - 1].ptr, h->P[l - 1].col, h->P[l - 1].val, h->rhs[l - 1], h->rhs[l]);
        h->A[l] = build_Ac(A, &h->P[l - 1]);
    }
    return h;
}

void amgo_free(amgo_hier *h)
{
    if (!h) return;
    for (int l = 0; l < h->levels; l++) {
        csr_free(&h->A[l]);
        free(h->rhs[l]);
        if (l + 1 < h->levels) { csr_free(&h->P[l]); free(h->cf[l]); }
    }
    free(h);
}

int amgo_levels(const amgo_hier *h) { return h->levels; }
void amgo_level_info(const amgo_hier *h, int l, idx *n, idx *nnzA, idx *nnzP, idx *ncP)
{
    *n = h->A[l].n_rows; *nnzA = h->A[l].nnz;
    *nnzP = (l + 1 < h->levels) ? h->P[l].nnz : 0;
    *ncP = (l + 1 < h->levels) ? h->P[l].n_cols : 0;
}
void amgo_get_A(const amgo_hier *h, int l, idx *ptr, idx *col, double *val)
{
    memcpy(ptr, h->A[l].ptr, sizeof(idx) * (size_t)(h->A[l].n_rows + 1));
    memcpy(col, h->A[l].col, sizeof(idx) * (size_t)h->A[l].nnz);
    memcpy(val, h->A[l].val, sizeof(double) * (size_t)h->A[l].nnz);
}
void amgo_get_P(const amgo_hier *h, int l, idx *ptr, idx *col, double *val)
{
    memcpy(ptr, h->P[l].ptr, sizeof(idx) * (size_t)(h->P[l].n_rows + 1));
    memcpy(col, h->P[l].col, sizeof(idx) * (size_t)h->P[l].nnz);
    memcpy(val, h->P[l].val, sizeof(double) * (size_t)h->P[l].nnz);
}
void amgo_get_rhs(const amgo_hier *h, int l, double *b) { memcpy(b, h->rhs[l], sizeof(double) * (size_t)h->A[l].n_rows); }
void amgo_get_cf(const amgo_hier *h, int l, unsigned char *cf) { memcpy(cf, h->cf[l], (size_t)h->A[l].n_rows); }

/* AMG::apply_AMG after initialization (src/AMG.cpp:277-308): pre 10 / coarse 200 / post 10 sweeps.
 * x holds x_levels[0] on entry and on return; returns ||b - A x||_2 on level 0. */
double amgo_pass(const amgo_hier *h, double *x, int pre, int coarse, int post)
{
    double *xl[AMGO_MAX_LEVELS];
    xl[0] = x;
    int i;
    for (i = 0; i < h->levels - 1; ++i) {
        const csr *A = &h->A[i];
        for (int s = 0; s < pre; s++) amgo_gs_sweep(A->n_rows, A->ptr, A->col, A->val, xl[i], h->rhs[i]);
        xl[i + 1] = (double *)malloc(sizeof(double) * (size_t)(h->A[i + 1].n_rows > 0 ? h->A[i + 1].n_rows : 1));
        amgo_restrict(A->n_rows, h->P[i].n_cols, h->P[i].ptr, h->P[i].col, h->P[i].val, xl[i], xl[i + 1]);
    }
    for (int s = 0; s < coarse; s++) amgo_gs_sweep(h->A[i].n_rows, h->A[i].ptr, h->A[i].col, h->A[i].val, xl[i], h->rhs[i]);
    for (i--; i >= 0; --i) {
        amgo_prolong_add(h->A[i].n_rows, h->P[i].ptr, h->P[i].col, h->P[i].val, xl[i + 1], xl[i]);
        free(xl[i + 1]);
        for (int s = 0; s < post; s++) amgo_gs_sweep(h->A[i].n_rows, h->A[i].ptr, h->A[i].col, h->A[i].val, xl[i], h->rhs[i]);
    }
    return amgo_residual(h->A[0].n_rows, h->A[0].ptr, h->A[0].col, h->A[0].val, x, h->rhs[0], NULL);
}
