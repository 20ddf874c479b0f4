// gmg_krylov.cuh -- fine-level vector kernels of the Krylov solvers (mgb_gmg_krylov): conjugate gradients and
// BiCGSTAB with one multigrid cycle as the preconditioner.  They replace the reference's never-executed BiCGSTAB
// "smoother" (GeometricMultigrid/include/solvers.hpp:86-216).
//
// All vectors use the level layout of gmg_kernels.cuh (rows x pitch, 128-B aligned rows, halo rows around the
// slab).  Every kernel is one pass over HBM in the marching frame; dot products leave one partial sum per CTA
// (deterministic second stage: k_reduce_partials).  Boundary rows of A are identity rows (linear_system.hpp:24-26).
#pragma once
#include "gmg_kernels.cuh"

namespace mgb {

// q = A p; partial[cta] = sum p_i q_i over the owned points
__global__ void __launch_bounds__(kTPB)
k_apply_dot(LevelGeom g, const double *__restrict__ p, double *__restrict__ q, double *__restrict__ partial)
{
    __shared__ double red[kTPB / 32];
    March m(g);
    double acc = 0.;
    if (m.i0 < g.rows) {
        const size_t P = g.pitch;
        const double *ur = p + (size_t)m.i0 * P;
        double2 up = ld2(ur - P + m.jl), ce = ld2(ur + m.jl);
#pragma unroll 4
        for (int i = m.i0; i < m.i1; ++i, ur += P) {
            double2 dn = ld2(ur + P + m.jl);
            double left, right;
            m.sides(g, ur, ce, left, right);
            const int gi = g.row0 + i;
            double2 o;
            // A p = -(0 - A p): the residual formula of solvers.hpp:278-294 with b = 0, negated exactly
            o.x = on_bdry(g, gi, m.j0) ? ce.x : -resid_point(0., up.x, left, ce.x, ce.y, dn.x, g.off, g.diag);
            o.y = on_bdry(g, gi, m.j0 + 1) ? ce.y : -resid_point(0., up.y, ce.x, ce.y, right, dn.y, g.off, g.diag);
            if (m.has1) st2(q + (size_t)i * P + m.j0, o);
            else if (m.has0) q[(size_t)i * P + m.j0] = o.x;
            if (m.has0) acc += ce.x * o.x;
            if (m.has1) acc += ce.y * o.y;
            up = ce; ce = dn;
        }
    }
    double t = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = t;
}

// partial[cta] = sum a_i b_i over the owned points
__global__ void __launch_bounds__(kTPB)
k_dot(LevelGeom g, const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ partial)
{
    __shared__ double red[kTPB / 32];
    March m(g);
    double acc = 0.;
    for (int i = m.i0; i < m.i1; ++i) {
        const size_t o = (size_t)i * g.pitch + m.jl;
        double2 x = ld2(a + o), y = ld2(b + o);
        if (m.has0) acc += x.x * y.x;
        if (m.has1) acc += x.y * y.y;
    }
    double t = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = t;
}

// The scalars of an iteration stay on the device: coefficients are read from d_scal slots, so the host only reads
// the residual norm back.  coef = sgn * s[num] / s[den] (den < 0: coef = sgn * s[num]).
struct Coef { int num, den; double sgn; };
__device__ __forceinline__ double coef_of(const double *s, Coef c)
{
    return c.den < 0 ? c.sgn * s[c.num] : c.sgn * (s[c.num] / s[c.den]);
}

// y += ca x (+ cb z when z != nullptr); optionally w = v + cw t; partial = sum w_i^2 (or of y when w == nullptr)
//   CG step:        u += alpha p            and r -= alpha q, sum r^2:  two launches of this kernel
//   BiCGSTAB:       s = r - alpha v;  x += alpha y + omega z;  r = s - omega t
__global__ void __launch_bounds__(kTPB)
k_axpy_dot(LevelGeom g, const double *__restrict__ scal, double *__restrict__ y, const double *__restrict__ x, Coef ca,
           const double *__restrict__ z, Coef cb, double *__restrict__ partial)
{
    __shared__ double red[kTPB / 32];
    March m(g);
    const double a = coef_of(scal, ca), b = z ? coef_of(scal, cb) : 0.;
    double acc = 0.;
    for (int i = m.i0; i < m.i1; ++i) {
        const size_t o = (size_t)i * g.pitch + m.jl;
        double2 yy = ld2(y + o), xx = ld2(x + o);
        yy.x = fma(a, xx.x, yy.x); yy.y = fma(a, xx.y, yy.y);
        if (z) { double2 zz = ld2(z + o); yy.x = fma(b, zz.x, yy.x); yy.y = fma(b, zz.y, yy.y); }
        if (m.has1) st2(y + (size_t)i * g.pitch + m.j0, yy);
        else if (m.has0) y[(size_t)i * g.pitch + m.j0] = yy.x;
        if (m.has0) acc += yy.x * yy.x;
        if (m.has1) acc += yy.y * yy.y;
    }
    if (partial) {
        double t = block_sum(acc, red);
        if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}

// out = x + ca (y + cb z)   (z may be nullptr: out = x + ca y).  out may alias y.
//   CG:        p = z + beta p                     -> out = p, x = z, y = p, ca = beta
//   BiCGSTAB:  p = r + beta (p - omega v)         -> out = p, x = r, y = p, z = v, ca = beta, cb = -omega
//              s = r - alpha v                    -> out = s, x = r, y = v, ca = -alpha
__global__ void __launch_bounds__(kTPB)
k_xpay(LevelGeom g, const double *__restrict__ scal, double *out, const double *__restrict__ x, const double *y, Coef ca,
       const double *__restrict__ z, Coef cb, double *__restrict__ partial)
{
    __shared__ double red[kTPB / 32];
    March m(g);
    const double a = coef_of(scal, ca), b = z ? coef_of(scal, cb) : 0.;
    double acc = 0.;
    for (int i = m.i0; i < m.i1; ++i) {
        const size_t o = (size_t)i * g.pitch + m.jl;
        double2 xx = ld2(x + o), yy = ld2(y + o);
        if (z) { double2 zz = ld2(z + o); yy.x = fma(b, zz.x, yy.x); yy.y = fma(b, zz.y, yy.y); }
        double2 r;
        r.x = fma(a, yy.x, xx.x); r.y = fma(a, yy.y, xx.y);
        if (m.has1) st2(out + (size_t)i * g.pitch + m.j0, r);
        else if (m.has0) out[(size_t)i * g.pitch + m.j0] = r.x;
        if (m.has0) acc += r.x * r.x;
        if (m.has1) acc += r.y * r.y;
    }
    if (partial) {
        double t = block_sum(acc, red);
        if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}

// u = f on the boundary points of the slab (the identity rows of A, linear_system.hpp:24-26)
__global__ void __launch_bounds__(256)
k_set_boundary(LevelGeom g, double *__restrict__ u, const double *__restrict__ f)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    // top and bottom rows (if owned), then the two side columns
    if (t < g.w) {
        if (g.row0 == 0) u[t] = f[t];
        if (g.row0 + g.rows == g.w) { const size_t o = (size_t)(g.rows - 1) * g.pitch + t; u[o] = f[o]; }
    }
    if (t < g.rows) {
        const size_t o = (size_t)t * g.pitch;
        u[o] = f[o];
        u[o + g.w - 1] = f[o + g.w - 1];
    }
}

// scalar bookkeeping between the launches of an iteration: s[dst] = s[src]
__global__ void k_scal_copy(double *s, int dst, int src) { if (threadIdx.x == 0) s[dst] = s[src]; }

}  // namespace mgb
