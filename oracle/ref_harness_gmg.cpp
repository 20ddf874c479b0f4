// ref_harness_gmg.cpp -- C entry points around the REFERENCE's own GMG classes.
//
// TEST INFRASTRUCTURE ONLY (see oracle/README.md).  This file contains no algorithm: it
// #includes the reference headers where they lie under /root/reference/GeometricMultigrid
// and drives the reference's classes exactly as its src/main.cpp does, so that
//   * oracle/gmg_oracle.c can be pinned against the real implementation, bit for bit, and
//   * bench.py can time the reference's CPU solver on the GPU box's host cores
//     (cpu_baseline.kind = "reference").
// Built by oracle/Makefile into oracle/_ref/libgmgref.so (git-ignored, travels with gpurun).
#include "allIncludes.hpp"
#include <sstream>
#include <cstring>

namespace {
using Vec = std::vector<double>;
using namespace MultiGrid;

struct CoutSilencer {   // the reference prints one line per cycle (multigrid.hpp:131)
    std::streambuf *old; std::ostringstream sink;
    CoutSilencer() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~CoutSilencer() { std::cout.rdbuf(old); }
};
}

extern "C" {

// DataVector (linear_system.hpp:85-92) with the test functions of utilities.cpp:138-147
void gmgref_rhs(size_t N, double length, int test, double *b)
{
    CoutSilencer q;
    std::function<double(const double, const double)> f, g;
    Utils::init_test_functions(f, g, test);
    SquareDomain dom(N, length, 0);
    DataVector<double> fvec(dom, f, g);
    for (size_t i = 0; i < N * N; i++) b[i] = fvec[i];
}

// one sweep of the reference's smoother (0 GS solvers.hpp:24-49, 1 Jacobi :53-84) on `level`
void gmgref_sweep(size_t N, double length, double alpha, int level, int kind,
                  double *sol, const double *b)
{
    SquareDomain dom(N, length, level);
    PoissonMatrix<double> A(dom, alpha);
    Vec s(sol, sol + N * N), rhs(b, b + N * N);
    if (kind == 1) { Jacobi_iteration<Vec> it(A, rhs); s * it; }
    else           { Gauss_Seidel_iteration<Vec> it(A, rhs); s * it; }
    // Jacobi swaps sol with its zero-initialised temp: only this level's points are defined.
    for (size_t i = 0; i < A.rows(); i++) sol[A.mask(i)] = s[A.mask(i)];
}

// Residual (solvers.hpp:219-308); returns sum r^2 * 1 (Norm()^2 * sum b^2), writes res if given
double gmgref_residual(size_t N, double length, double alpha, int level,
                       const double *sol, const double *b, double *res, double *relnorm)
{
    SquareDomain dom(N, length, level);
    PoissonMatrix<double> A(dom, alpha);
    Vec s(sol, sol + N * N), rhs(b, b + N * N), r(N * N, 0.);
    Residual<Vec> R(A, rhs, r);
    s * R;
    double nb = 0;   // same accumulation as the 3-argument ctor, solvers.hpp:237-242
    for (size_t i = 0; i < A.rows(); i++) { double v = rhs[A.mask(i)]; nb += v * v; }
    if (res) for (size_t i = 0; i < A.rows(); i++) res[A.mask(i)] = r[A.mask(i)];
    if (relnorm) *relnorm = R.Norm();
    return R.Norm() * R.Norm() * nb;
}

// InterpolationClass (multigrid.cpp:3-27): level_coarse -> level_coarse-1, in place
void gmgref_prolong(size_t N, double length, double alpha, int level_coarse, double *vec)
{
    SquareDomain dc(N, length, level_coarse), df(N, length, level_coarse - 1);
    PoissonMatrix<double> Ac(dc, alpha), Af(df, alpha);
    InterpolationClass P(Ac, Af);
    Vec v(vec, vec + N * N);
    v * P;
    std::memcpy(vec, v.data(), N * N * sizeof(double));
}

// The driver loop of src/main.cpp:73-116 on caller-supplied b and u (u is updated in place).
// smoother follows -smt (0 GS, 1 Jacobi, 2 -> Jacobi as in main.cpp:103-106).
// Runs until hist.back() <= tol or maxiter cycles.  coarse_relres (nullable) receives the value
// the reference prints per cycle.  Returns the number of history entries written.
int gmgref_solve(size_t N, double length, double alpha, int L, int smoother,
                 const double *b, double *u, double tol, int maxiter, double *hist,
                 double *coarse_relres)
{
    CoutSilencer q;
    std::vector<SquareDomain> domains;
    for (int i = 0; i < L; i++) domains.push_back(SquareDomain(N, length, i));
    std::vector<PoissonMatrix<double>> mats;
    for (auto &d : domains) mats.push_back(PoissonMatrix<double>(d, alpha));
    Vec fvec(b, b + N * N), uu(u, u + N * N), res(N * N, 0.);
    Residual<Vec> RES(mats.front(), fvec, res);
    Gauss_Seidel_iteration<Vec> GS(mats.front(), fvec);
    int n = 0;
    uu * RES;
    hist[n++] = RES.Norm();
    auto loop = [&](auto &MG) {
        for (int i = 0; i < maxiter; i++) {
            q.sink.str("");
            uu * GS * GS * MG;
            if (coarse_relres) {
                std::string s = q.sink.str();
                auto p = s.rfind(": ");
                coarse_relres[i] = (p == std::string::npos) ? -1. : std::atof(s.c_str() + p + 2);
            }
            uu * RES;
            hist[n++] = RES.Norm();
            if (hist[n - 1] <= tol) break;
        }
    };
    if (smoother == 0) {
        SawtoothMGIteration<Vec, Gauss_Seidel_iteration<Vec>> MG(mats, fvec);
        loop(MG);
    } else {
        SawtoothMGIteration<Vec, Jacobi_iteration<Vec>> MG(mats, fvec);
        loop(MG);
    }
    std::memcpy(u, uu.data(), N * N * sizeof(double));
    return n;
}

// One SawtoothMGIteration::apply_iteration_to_vec (multigrid.hpp:126-145) without pre-smoothing
void gmgref_cycle(size_t N, double length, double alpha, int L, int smoother,
                  const double *b, double *u)
{
    CoutSilencer q;
    std::vector<SquareDomain> domains;
    for (int i = 0; i < L; i++) domains.push_back(SquareDomain(N, length, i));
    std::vector<PoissonMatrix<double>> mats;
    for (auto &d : domains) mats.push_back(PoissonMatrix<double>(d, alpha));
    Vec fvec(b, b + N * N), uu(u, u + N * N);
    if (smoother == 0) { SawtoothMGIteration<Vec, Gauss_Seidel_iteration<Vec>> MG(mats, fvec); uu * MG; }
    else               { SawtoothMGIteration<Vec, Jacobi_iteration<Vec>> MG(mats, fvec); uu * MG; }
    std::memcpy(u, uu.data(), N * N * sizeof(double));
}

void gmgref_set_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int gmgref_openmp_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
