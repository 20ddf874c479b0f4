// amg_tail.cuh -- the coarse tail of an AMG cycle as ONE persistent CTA (north_star item 3).
//
// Below a few thousand rows a level costs launch latency, not bandwidth: a multicolour sweep is one launch per
// colour (13-16 colours on the Galerkin levels), a V(2,2) visit of one level is ~60 launches of a few microseconds
// of work each, and the reference's pass runs 200 sweeps on the last level (AMG/src/AMG.cpp:295).  Here a single
// CTA of 1024 threads walks every level from `first` down to the coarsest one and back with __syncthreads()
// between the independent sets (colours / wavefronts): the level operators stay in L2, the host sees one launch.
// Levels in the tail are never sharded (multi-GPU: they are replicated, every rank runs the same tail).
//
// Every row update calls the same device functions as the per-level kernels of amg_kernels.cuh, in the same
// order inside a row, so the tail changes no result: exact-order arithmetic stays bit-identical to the reference.
#pragma once
#include "amg_kernels.cuh"

namespace mgb {

constexpr int kAmgTailThreads = 1024;
constexpr int kAmgTailMaxLevels = 12;

struct AmgTailLevel {
    CsrDev A, P, R;                  // P (this level <- next), R = P^T; unused on the last level
    const double *diag, *dl1;        // a_ii; a_ii + sum |a_ij| (l1-Jacobi, may be null otherwise)
    int kind;                        // smoother of this level: 0 lexicographic GS, 1 Jacobi, 3 multicolour GS, 4 l1-Jacobi
    double *x, *b, *tmp;
    const int *grp_ptr, *grp_rows;   // independent sets of the smoother: colours, or wavefronts of the lexicographic sweep
    int n_groups;
};

struct AmgTailParams {
    int nlev;                        // lv[0] = first tail level ... lv[nlev-1] = coarsest level
    int exact;                       // 1: reference term order, unfused IEEE ops; 0: the fast kernels' arithmetic
    int mode;                        // 0: the reference's pass (x_c = P^T x, AMG.cpp:277-308); 1: correction scheme (b_c = P^T r, x_c = 0)
    int pre, coarse, post;           // sweeps before the transfer down, on the last level, after the transfer up
    double omega;                    // Jacobi weight
    AmgTailLevel lv[kAmgTailMaxLevels];
};

// one multicolour / lexicographic Gauss-Seidel sweep: the groups in order, a barrier after each
__device__ __forceinline__ void tail_gs_sweep(const AmgTailLevel &L, bool exact)
{
    for (int g = 0; g < L.n_groups; ++g) {
        const int a = L.grp_ptr[g], b = L.grp_ptr[g + 1];
        for (int t = a + (int)threadIdx.x; t < b; t += kAmgTailThreads) {
            const int i = L.grp_rows[t];
            if (exact) gs_row_exact(L.A, L.diag, L.x, L.b, i);
            else L.x[i] = (L.b[i] - offdiag_dot_fast(L.A, L.x, i)) / L.diag[i];
        }
        __syncthreads();
    }
}

// one Jacobi sweep: all rows into tmp, then back (the per-level kernels swap the two pointers instead)
__device__ __forceinline__ void tail_jacobi_sweep(const AmgTailLevel &L, bool exact, double omega)
{
    const int n = L.A.n_rows;
    if (exact) {                     // k_amg_jacobi_vec: kLanes lanes per row, tree reduction
        const int lane = threadIdx.x & (kLanes - 1), sub = threadIdx.x / kLanes;
        for (int base = 0; base < n; base += kAmgTailThreads / kLanes) {
            const int i = base + sub;
            const bool ok = i < n;
            const double sum = offdiag_dot_vec(L.A, L.x, ok ? i : 0, lane, ok);
            if (ok && lane == 0) L.tmp[i] = relax((L.b[i] - sum) / L.diag[i], L.x[i], omega);
        }
    } else {                         // k_amg_sell<1>: one thread per row, entries in ascending column order
        for (int i = threadIdx.x; i < n; i += kAmgTailThreads)
            L.tmp[i] = relax((L.b[i] - offdiag_dot_fast(L.A, L.x, i)) / L.diag[i], L.x[i], omega);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kAmgTailThreads) L.x[i] = L.tmp[i];
    __syncthreads();
}

// one l1-Jacobi sweep (k_amg_sell<5>)
__device__ __forceinline__ void tail_l1_sweep(const AmgTailLevel &L)
{
    const int n = L.A.n_rows;
    for (int i = threadIdx.x; i < n; i += kAmgTailThreads) {
        const double xi = L.x[i];
        L.tmp[i] = xi + (L.b[i] - (offdiag_dot_fast(L.A, L.x, i) + L.diag[i] * xi)) / L.dl1[i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kAmgTailThreads) L.x[i] = L.tmp[i];
    __syncthreads();
}

__device__ __forceinline__ void tail_smooth(const AmgTailLevel &L, bool exact, double omega, int sweeps)
{
    for (int s = 0; s < sweeps; ++s) {
        if (L.kind == 1) tail_jacobi_sweep(L, exact, omega);
        else if (L.kind == 4) tail_l1_sweep(L);
        else tail_gs_sweep(L, exact || L.kind == 0);
    }
}

// tmp = b - A x
__device__ __forceinline__ void tail_residual(const AmgTailLevel &L, bool exact)
{
    for (int i = threadIdx.x; i < L.A.n_rows; i += kAmgTailThreads)
        L.tmp[i] = exact ? residual_row_exact(L.A, L.x, L.b, i)
                         : L.b[i] - (offdiag_dot_fast(L.A, L.x, i) + L.diag[i] * L.x[i]);
    __syncthreads();
}

// out = R in   (R = P^T of level F; rows = entries of the next level)
__device__ __forceinline__ void tail_restrict(const AmgTailLevel &F, const double *in, double *out, bool exact)
{
    const int n = F.R.n_rows;
    for (int m = threadIdx.x; m < n; m += kAmgTailThreads)
        out[m] = exact ? spmv_row_exact(F.R, in, m) : row_dot_fast(F.R, in, m);
    __syncthreads();
}

__global__ void __launch_bounds__(kAmgTailThreads)
k_amg_tail(AmgTailParams p)
{
    const bool exact = p.exact != 0;
    const int last = p.nlev - 1;
    for (int l = 0; l < last; ++l) {                          // downward
        const AmgTailLevel &F = p.lv[l], &C = p.lv[l + 1];
        tail_smooth(F, exact, p.omega, p.pre);
        if (p.mode == 0) tail_restrict(F, F.x, C.x, exact);   // AMG.cpp:50-74: the SOLUTION is restricted
        else {
            tail_residual(F, exact);
            tail_restrict(F, F.tmp, C.b, exact);
            for (int i = threadIdx.x; i < C.A.n_rows; i += kAmgTailThreads) C.x[i] = 0.;
            __syncthreads();
        }
    }
    tail_smooth(p.lv[last], exact, p.omega, p.coarse);
    for (int l = last - 1; l >= 0; --l) {                     // upward
        const AmgTailLevel &F = p.lv[l], &C = p.lv[l + 1];
        for (int i = threadIdx.x; i < F.P.n_rows; i += kAmgTailThreads) prolong_row_add(F.P, C.x, F.x, i);
        __syncthreads();
        tail_smooth(F, exact, p.omega, p.post);
    }
}

}  // namespace mgb
