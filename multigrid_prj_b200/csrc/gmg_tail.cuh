// gmg_tail.cuh -- the coarse tail of the sawtooth cycle as ONE persistent CTA.
//
// Levels whose side is <= tail_max_width hold a few thousand points at most: as separate launches
// they cost a launch latency each (restriction, the data-dependent coarse-solve loop of
// Solver::Solve, solvers.hpp:324-342, then prolongation + nu sweeps per level).  Here a single CTA
// of 1024 threads walks the whole tail with __syncthreads() between phases: the level arrays stay
// in L1/L2, the coarse-solve loop `while (relres > tol)` runs on the device, and the host sees one
// launch and no synchronisation -- which is what makes the whole cycle capturable in a CUDA graph.
// All formulas are the same device functions the per-level kernels use (bit-identical results).
#pragma once
#include "gmg_kernels.cuh"

namespace mgb {

constexpr int kTailThreads = 1024;
constexpr int kTailMaxLevels = 12;

struct TailLevel {
    LevelGeom g;
    double *e, *r, *t;
};

struct TailParams {
    int nlev;                    // tail levels: lv[0] is the finest of them, lv[nlev-1] the coarsest level
    TailLevel lv[kTailMaxLevels];
    int kind;                    // MGB_SMOOTH_* (0 lexicographic GS, 1 Jacobi, 3 red-black GS)
    int fast;                    // red-black arithmetic: 0 reference formula, 1 b/diag + sum/4 with FMA
    int restriction;             // 0 injection, 1 half injection, 2 full weighting
    int first_is_level1;         // lv[1] is level 1 of the hierarchy (half injection scales only there)
    int nu;
    int coarse_maxit;
    double coarse_tol;
    double *out;                 // out[0] = coarse relative residual (multigrid.hpp:131), out[1] = coarse sweeps
};

__device__ __forceinline__ double tail_block_sum(double v, double *red)
{
    double t = block_sum(v, red);
    __shared__ double bc;
    if (threadIdx.x == 0) bc = t;
    __syncthreads();
    t = bc;
    __syncthreads();
    return t;
}

// A level as the tail sees it: either its arrays in global memory (L2) or, when it fits, a copy in
// shared memory (kTailSmemW^2 points or fewer) -- every sweep phase then costs a shared-memory round
// trip instead of an L2 one.
constexpr int kTailSmemW = 65;
constexpr int kTailSmemBytes = 3 * kTailSmemW * kTailSmemW * (int)sizeof(double);

struct TailView {
    double *e, *r, *t;
    int pitch;
    bool in_smem;
};

__device__ __forceinline__ TailView tail_view(const TailLevel &L, double *smem)
{
    TailView v;
    if (L.g.w <= kTailSmemW) {
        const int n = L.g.w * L.g.w;
        v.e = smem; v.r = smem + n; v.t = smem + 2 * n; v.pitch = L.g.w; v.in_smem = true;
    } else {
        v.e = L.e; v.r = L.r; v.t = L.t; v.pitch = L.g.pitch; v.in_smem = false;
    }
    return v;
}

// bring the level's rhs into the view / write the level's solution back to HBM (no-ops for global views)
__device__ __forceinline__ void tail_load_rhs(const TailLevel &L, const TailView &v)
{
    if (!v.in_smem) return;
    const int w = L.g.w;
    for (int idx = threadIdx.x; idx < w * w; idx += blockDim.x) {
        const int i = idx / w, j = idx - i * w;
        v.r[i * v.pitch + j] = L.r[(size_t)i * L.g.pitch + j];
    }
    __syncthreads();
}
__device__ __forceinline__ void tail_store_sol(const TailLevel &L, const TailView &v)
{
    if (!v.in_smem) return;
    const int w = L.g.w;
    for (int idx = threadIdx.x; idx < w * w; idx += blockDim.x) {
        const int i = idx / w, j = idx - i * w;
        L.e[(size_t)i * L.g.pitch + j] = v.e[i * v.pitch + j];
    }
    __syncthreads();
}

__device__ __forceinline__ double tail_update(const LevelGeom &g, const double *u, const double *b, int P, int i, int j,
                                              bool fast, double inv_diag)
{
    const double *c = u + i * P + j;
    const double bv = b[i * P + j];
    if (on_bdry(g, i, j)) return bv;
    if (fast) return fma(0.25, (c[-P] + c[P]) + (c[-1] + c[1]), __dmul_rn(bv, inv_diag));
    return smooth_point(bv, c[-P], c[-1], c[1], c[P], g.off, g.diag);
}

// one smoothing sweep of `kind` on a whole (replicated) level, in place
__device__ void tail_sweep(const TailParams &p, const LevelGeom &g, const TailView &L)
{
    const int w = g.w, tid = threadIdx.x, nt = blockDim.x, P = L.pitch;
    const double inv_diag = 1.0 / g.diag;
    if (p.kind == 3) {                                   // red-black: colour 0 then colour 1
        const int half = (w + 1) / 2;
        for (int colour = 0; colour < 2; ++colour) {
            for (int idx = tid; idx < w * half; idx += nt) {
                const int i = idx / half, j = 2 * (idx - i * half) + ((i + colour) & 1);
                if (j < w) L.e[i * P + j] = tail_update(g, L.e, L.r, P, i, j, p.fast != 0, inv_diag);
            }
            __syncthreads();
        }
    } else if (p.kind == 1) {                            // Jacobi: into t, then back (solvers.hpp:64-83)
        for (int idx = tid; idx < w * w; idx += nt) {
            const int i = idx / w, j = idx - i * w;
            L.t[i * P + j] = tail_update(g, L.e, L.r, P, i, j, false, inv_diag);
        }
        __syncthreads();
        for (int idx = tid; idx < w * w; idx += nt) {
            const int i = idx / w, j = idx - i * w;
            L.e[i * P + j] = L.t[i * P + j];
        }
        __syncthreads();
    } else {                                             // lexicographic GS as an anti-diagonal wavefront
        for (int d = 0; d <= 2 * (w - 1); ++d) {
            const int ilo = max(0, d - (w - 1)), ihi = min(d, w - 1);
            for (int i = ilo + tid; i <= ihi; i += nt) {
                const int j = d - i;
                L.e[i * P + j] = tail_update(g, L.e, L.r, P, i, j, false, inv_diag);
            }
            __syncthreads();
        }
    }
}

// sum (r - A e)^2 over the level (solvers.hpp:278-294, norm only)
__device__ double tail_residual_sumsq(const LevelGeom &g, const TailView &L, double *red)
{
    const int w = g.w, P = L.pitch;
    double acc = 0.;
    for (int idx = threadIdx.x; idx < w * w; idx += blockDim.x) {
        const int i = idx / w, j = idx - i * w;
        const double *c = L.e + i * P + j;
        const double bv = L.r[i * P + j];
        const double r = on_bdry(g, i, j) ? __dsub_rn(bv, c[0])
                                           : resid_point(bv, c[-P], c[-1], c[0], c[1], c[P], g.off, g.diag);
        acc += r * r;
    }
    return tail_block_sum(acc, red);
}

__global__ void __launch_bounds__(kTailThreads)
k_coarse_tail(TailParams p)
{
    extern __shared__ double tail_smem[];
    __shared__ double red[kTailThreads / 32];
    const int tid = threadIdx.x, nt = blockDim.x;
    // ---- restriction of the residual down the tail (lv[0].r was produced by the caller) -------------
    for (int l = 1; l < p.nlev; ++l) {
        const LevelGeom &gf = p.lv[l - 1].g, &gc = p.lv[l].g;
        const double *rf = p.lv[l - 1].r;
        double *rc = p.lv[l].r;
        const ptrdiff_t P = gf.pitch;
        const double scale = (p.restriction == 1 && l == 1 && p.first_is_level1) ? 0.5 : 1.0;
        for (int idx = tid; idx < gc.w * gc.w; idx += nt) {
            const int I = idx / gc.w, J = idx - I * gc.w;
            const double *c = rf + (size_t)(2 * I) * P + 2 * J;
            double v;
            if (on_bdry(gc, I, J)) v = c[0];
            else if (p.restriction != 2) v = __dmul_rn(scale, c[0]);
            else {
                double edge = __dadd_rn(__dadd_rn(__dadd_rn(c[-P], c[-1]), c[1]), c[P]);
                double corner = __dadd_rn(__dadd_rn(__dadd_rn(c[-P - 1], c[-P + 1]), c[P - 1]), c[P + 1]);
                v = __dadd_rn(__dadd_rn(__dmul_rn(0.25, c[0]), __dmul_rn(0.125, edge)), __dmul_rn(0.0625, corner));
            }
            rc[(size_t)I * gc.pitch + J] = v;
        }
        __syncthreads();
    }
    // ---- coarse solve (multigrid.hpp:128-131, solvers.hpp:324-342) -----------------------------------------
    {
        const TailLevel &C = p.lv[p.nlev - 1];
        const TailView V = tail_view(C, tail_smem);
        tail_load_rhs(C, V);
        double nb = 0.;
        {
            double acc = 0.;
            for (int idx = tid; idx < C.g.w * C.g.w; idx += nt) {
                const int i = idx / C.g.w, j = idx - i * C.g.w;
                const double v = V.r[i * V.pitch + j];
                acc += v * v;
                V.e[i * V.pitch + j] = 0.;                                // err == 0 on entry (multigrid.hpp:143)
            }
            nb = tail_block_sum(acc, red);
        }
        double norm = tail_residual_sumsq(C.g, V, red);
        int its = 0;
        while (sqrt(norm / nb) > p.coarse_tol && its < p.coarse_maxit) {
            tail_sweep(p, C.g, V);
            ++its;
            norm = tail_residual_sumsq(C.g, V, red);
        }
        if (tid == 0) { p.out[0] = sqrt(norm / nb); p.out[1] = (double)its; }
        tail_store_sol(C, V);
    }
    // ---- upward leg inside the tail (multigrid.hpp:134-139) -----------------------------------------------
    for (int l = p.nlev - 1; l > 0; --l) {
        const TailLevel &Lc = p.lv[l], &Lf = p.lv[l - 1];
        const LevelGeom &gc = Lc.g, &gf = Lf.g;
        const TailView V = tail_view(Lf, tail_smem);
        tail_load_rhs(Lf, V);
        for (int idx = tid; idx < gf.w * gf.w; idx += nt) {
            const int i = idx / gf.w, j = idx - i * gf.w;
            const double *cn = Lc.e + (size_t)(i >> 1) * gc.pitch + (j >> 1);      // coarse solution: in HBM/L2
            double a, c2 = 0.;
            const bool oddj = j & 1;
            if ((i & 1) == 0) { a = cn[0]; if (oddj) c2 = cn[1]; }
            else {
                const double *cs = cn + gc.pitch;
                a = __dmul_rn(0.5, __dadd_rn(cn[0], cs[0]));
                if (oddj) c2 = __dmul_rn(0.5, __dadd_rn(cn[1], cs[1]));
            }
            V.e[i * V.pitch + j] = oddj ? __dmul_rn(0.5, __dadd_rn(a, c2)) : a;
        }
        __syncthreads();
        for (int s = 0; s < p.nu; ++s) tail_sweep(p, gf, V);
        tail_store_sol(Lf, V);
    }
}

}  // namespace mgb
