// ref_harness_amg.cpp -- C entry points around the REFERENCE's own AMG classes (test infrastructure
// only; see oracle/README.md).  Compiled by oracle/Makefile together with the reference sources where
// they lie under /root/reference/AMG into oracle/_ref/libamgref.so.
//
// The only intervention: getRandomInit (AMG/src/Utilities.cpp:30-40), which seeds the C/F splitting
// from std::random_device, is overridden at link time (this object comes first and the link uses
// --allow-multiple-definition) so that the hierarchy is reproducible; the value it returns is set
// through amgref_set_starts().
#include "CSRMatrix.hpp"
#include "Utilities.hpp"
#include "AMG.hpp"
#include "FEM.hpp"

namespace {
std::vector<long> g_starts;     // start index per coarsening step; negative -> n/2
size_t g_start_pos = 0;
struct Quiet {
    std::streambuf *o, *e; std::ostringstream sink;
    Quiet() : o(std::cout.rdbuf(sink.rdbuf())), e(std::cerr.rdbuf(sink.rdbuf())) {}
    ~Quiet() { std::cout.rdbuf(o); std::cerr.rdbuf(e); }
};
std::unique_ptr<Matrix> g_A;
std::vector<double> g_rhs;
std::unique_ptr<AMG> g_amg;

void csr_out(CSRMatrix &M, long *ptr, long *col, double *val)
{
    long p = 0;
    for (size_t i = 0; i < M.rows(); ++i) {
        ptr[i] = p;
        for (const auto &e : M.nonZerosInRow(i)) { col[p] = (long)e.first; val[p] = e.second; ++p; }
    }
    ptr[M.rows()] = p;
}
long csr_nnz(CSRMatrix &M)
{
    long p = 0;
    for (size_t i = 0; i < M.rows(); ++i) p += (long)M.nonZerosInRow(i).size();
    return p;
}
}

int getRandomInit(int max)      // replaces AMG/src/Utilities.cpp:30-40
{
    long s = g_start_pos < g_starts.size() ? g_starts[g_start_pos] : -1;
    ++g_start_pos;
    if (s < 0) s = max / 2;
    if (s >= max) s = max - 1;
    return (int)s;
}

#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {

// AMG.hpp:314-331 races on std::map rows with more than one thread: the checker always runs it serial
void amgref_set_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void amgref_set_starts(const long *starts, int n)
{
    g_starts.assign(starts, starts + n);
    g_start_pos = 0;
}

// P1 stiffness matrix and load vector on the interior nodes of a Gmsh 4.1 mesh, assembled with the
// reference's mesh reader and element (FEM.cpp:3-316, FEM.hpp:174-258) in the way of AMG/src/main.cpp:34-117:
// vertex quadrature with weights area2/3, Dirichlet values lifted into the right-hand side.
int amgref_assemble(const char *msh, long *n_out, long *nnz_out)
{
    Quiet q;
    LinearFE fe;
    TriangularMesh mesh(fe);
    mesh.import_from_msh(msh);
    const size_t n = mesh.n_nodes() - mesh.n_b_nodes();
    g_A = std::make_unique<Matrix>(n, n);
    g_rhs.assign(n, 0.0);
    for (const auto &elem : mesh.element_iterators()) {
        std::vector<Point> v(fe.get_ndofs());
        for (size_t a = 0; a < v.size(); ++a) v[a] = mesh.get_nodes()[elem.at(a)];
        fe.set_dofs(v);
        auto &qp = fe.get_quadrature_points();
        auto &qw = fe.get_quadrature_weights();
        auto &gr = fe.get_gradients();
        for (size_t a = 0; a < v.size(); ++a) {
            if (v[a].is_on_boundary) continue;
            for (size_t b = 0; b < v.size(); ++b) {
                if (v[b].is_on_boundary) continue;
                for (size_t k = 0; k < qp.size(); ++k)
                    g_A->at(v[a].set_index, v[b].set_index) +=
                        alpha(qp[k].x, qp[k].y) * (gr[a][0] * gr[b][0] + gr[a][1] * gr[b][1]) * qw[k];
            }
            for (size_t k = 0; k < qp.size(); ++k)
                g_rhs[v[a].set_index] += forcing_term(v[a].x, v[a].y) * fe.get_basis_function(a)(qp[k]) * qw[k];
        }
        if (fe.is_on_boundary())
            for (size_t a = 0; a < v.size(); ++a) {
                if (v[a].is_on_boundary) continue;
                for (size_t b = 0; b < v.size(); ++b) {
                    if (!v[b].is_on_boundary) continue;
                    for (size_t k = 0; k < qp.size(); ++k)
                        g_rhs[v[a].set_index] -= boundary_function(v[b].x, v[b].y) * alpha(qp[k].x, qp[k].y) *
                                                 (gr[a][0] * gr[b][0] + gr[a][1] * gr[b][1]) * qw[k];
                }
            }
    }
    g_A->count_non_zeros();
    *n_out = (long)n;
    *nnz_out = (long)g_A->non_zeros();
    return 0;
}

void amgref_get_system(long *ptr, long *col, double *val, double *rhs)
{
    CSRMatrix M(*g_A);
    M.copy_from(*g_A);
    csr_out(M, ptr, col, val);
    std::copy(g_rhs.begin(), g_rhs.end(), rhs);
}

// AMG ctor + initialization() (AMG.hpp:33-41, AMG.cpp:76-120) on a caller-supplied CSR system
int amgref_build(long n, const long *ptr, const long *col, const double *val, const double *rhs, const double *x0, int levels)
{
    Quiet q;
    Matrix A(n, n);
    for (long i = 0; i < n; ++i)
        for (long k = ptr[i]; k < ptr[i + 1]; ++k) A.at(i, col[k]) = val[k];
    A.count_non_zeros();
    std::vector<double> b(rhs, rhs + n), x(x0, x0 + n);
    g_amg = std::make_unique<AMG>(A, x, (size_t)levels, b);
    g_start_pos = 0;
    g_amg->initialization();
    return (int)g_amg->levels_matrix.size();
}

void amgref_level_info(int l, long *n, long *nnzA, long *nnzP, long *ncP)
{
    *n = (long)g_amg->levels_matrix[l]->rows();
    *nnzA = csr_nnz(*g_amg->levels_matrix[l]);
    const bool hasP = (size_t)l < g_amg->P_matrices.size();
    *nnzP = hasP ? csr_nnz(*g_amg->P_matrices[l]) : 0;
    *ncP = hasP ? (long)g_amg->P_matrices[l]->cols() : 0;
}
void amgref_get_A(int l, long *ptr, long *col, double *val) { csr_out(*g_amg->levels_matrix[l], ptr, col, val); }
void amgref_get_P(int l, long *ptr, long *col, double *val) { csr_out(*g_amg->P_matrices[l], ptr, col, val); }
void amgref_get_rhs(int l, double *b) { std::copy(g_amg->rhs[l].begin(), g_amg->rhs[l].end(), b); }

// the body of AMG::apply_AMG() after initialization() (AMG.cpp:282-304); returns the printed residual norm
double amgref_pass(double *x_out)
{
    Quiet q;
    AMG &a = *g_amg;
    int i;
    for (i = 0; i < (int)a.number_of_levels - 1; ++i) {
        a.apply_smoother_operator(i, 10);
        a.apply_restriction_operator(i + 1);
    }
    a.apply_smoother_operator(i, 200);
    for (i--; i >= 0; --i) {
        a.apply_prolungation_operator(i);
        a.apply_smoother_operator(i, 10);
    }
    double r = a.compute_residual(0);
    std::vector<double> s = a.get_solution();
    std::copy(s.begin(), s.end(), x_out);
    return r;
}

// `sweeps` lexicographic Gauss-Seidel sweeps (Utilities.hpp:44-58) on a caller-supplied system
void amgref_gs(long n, const long *ptr, const long *col, const double *val, const double *b, double *x, int sweeps)
{
    Matrix A(n, n);
    for (long i = 0; i < n; ++i)
        for (long k = ptr[i]; k < ptr[i + 1]; ++k) A.at(i, col[k]) = val[k];
    A.count_non_zeros();
    CSRMatrix M(A);
    M.copy_from(A);
    std::vector<double> rhs(b, b + n), sol(x, x + n);
    Gauss_Seidel_iteration<std::vector<double>> GS(M, rhs);
    for (int s = 0; s < sweeps; ++s) sol * GS;
    std::copy(sol.begin(), sol.end(), x);
}

}  // extern "C"
