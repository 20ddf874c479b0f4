// gmg_kernels.cuh -- sm_100a kernels of the geometric-multigrid solve phase.
//
// Data layout (HBM): every level l keeps COMPACT fp64 arrays of (rows_l + 2) x pitch_l doubles:
// rows_l owned rows of this rank's slab plus one halo row above and below; pitch_l is a multiple
// of 16 doubles (128 B) so that column 0 of every row is 128-B aligned and every thread can move
// 16-B (double2) words.  There is no shared fine-sized vector and no mask() indirection as in the
// reference (GeometricMultigrid/include/domain.hpp:78-80); level l point (I,J) is the reference's
// fine entry (2^l I, 2^l J).
//
// Arithmetic: every stencil formula is evaluated in the reference's source order with the
// round-to-nearest intrinsics (__dmul_rn, __dadd_rn, __ddiv_rn ...), which nvcc never contracts
// into FMAs.  The x86-64 reference build contains no FMA either, so sweeps, residuals and grid
// transfers are bit-identical to the reference, not merely within 1e-12.  The kernels are
// HBM-bound (24 B per point against ~30 flops), so the unfused arithmetic is free.
#pragma once
#include "gmg_common.cuh"

namespace mgb {

constexpr int kTPB = 128;      // threads per CTA in the marching kernels: 256 columns per CTA
constexpr int kRowsPerCta = 32;

// ---- marching frame ------------------------------------------------------------------------
// A thread owns two adjacent columns (one double2) and marches down kRowsPerCta rows with the
// rows i-1, i, i+1 in registers, so every row of u is requested from L2/HBM once per CTA
// (plus one halo row at each end of the chunk).  Left/right neighbours come from the adjacent
// lanes by shuffle; only the two edge lanes of a warp issue an extra scalar load.
struct March {
    int j0, jl, lane, i0, i1;
    bool has0, has1;
    __device__ __forceinline__ March(const LevelGeom &g)
    {
        j0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
        lane = threadIdx.x & 31;
        jl = min(j0, g.pitch - 2);           // clamped column for loads of idle threads
        has0 = j0 < g.w;
        has1 = j0 + 1 < g.w;
        i0 = blockIdx.y * kRowsPerCta;
        i1 = min(i0 + kRowsPerCta, g.rows);
    }
    // left neighbour of column j0 and right neighbour of column j0+1 on the row held in `c`
    __device__ __forceinline__ void sides(const LevelGeom &g, const double *row, double2 c,
                                          double &left, double &right) const
    {
        left = __shfl_up_sync(0xffffffffu, c.y, 1);
        right = __shfl_down_sync(0xffffffffu, c.x, 1);
        if (lane == 0) left = (j0 > 0 && has0) ? row[j0 - 1] : 0.;
        if (lane == 31) right = (j0 + 2 < g.w) ? row[j0 + 2] : 0.;
    }
};

__device__ __forceinline__ bool on_bdry(const LevelGeom &g, int gi, int j)
{
    return gi == 0 || gi == g.w - 1 || j == 0 || j == g.w - 1;      // domain.cpp:20-23
}

// ---- Jacobi sweep, out of place (solvers.hpp:64-83; the reference writes `temp` then swaps) --
__global__ void __launch_bounds__(kTPB)
k_jacobi(LevelGeom g, const double *__restrict__ u, const double *__restrict__ b, double *__restrict__ out, double omega)
{
    March m(g);
    if (m.i0 >= g.rows) return;
    const size_t P = g.pitch;
    const double *ur = u + (size_t)m.i0 * P;
    double2 up = ld2(ur - P + m.jl), ce = ld2(ur + m.jl);
#pragma unroll 4
    for (int i = m.i0; i < m.i1; ++i, ur += P) {
        double2 dn = ld2(ur + P + m.jl);
        double2 bb = ld2(b + (size_t)i * P + m.jl);
        double left, right;
        m.sides(g, ur, ce, left, right);
        const int gi = g.row0 + i;
        double2 o;
        o.x = on_bdry(g, gi, m.j0) ? bb.x : smooth_point(bb.x, up.x, left, ce.y, dn.x, g.off, g.diag);
        o.y = on_bdry(g, gi, m.j0 + 1) ? bb.y : smooth_point(bb.y, up.y, ce.x, right, dn.y, g.off, g.diag);
        if (omega != 1.) {                  // weighted Jacobi (north_star); the reference is omega = 1 (solvers.hpp:64-83)
            if (!on_bdry(g, gi, m.j0)) o.x = ce.x + omega * (o.x - ce.x);
            if (!on_bdry(g, gi, m.j0 + 1)) o.y = ce.y + omega * (o.y - ce.y);
        }
        if (m.has1) st2(out + (size_t)i * P + m.j0, o);
        else if (m.has0) out[(size_t)i * P + m.j0] = o.x;
        up = ce; ce = dn;
    }
}

// ---- red-black Gauss-Seidel, one colour, in place ---------------------------------------------
// colour of (gi, j) = (gi + j) & 1 with gi the GLOBAL row: slabs agree on the colouring.
// All four neighbours of a point have the other colour and are not written by this launch.
__global__ void __launch_bounds__(kTPB)
k_rbgs_colour(LevelGeom g, double *u, const double *__restrict__ b, int colour)
{
    March m(g);
    if (m.i0 >= g.rows) return;
    const size_t P = g.pitch;
    double *ur = u + (size_t)m.i0 * P;
    double2 up = ld2(ur - P + m.jl), ce = ld2(ur + m.jl);
#pragma unroll 2
    for (int i = m.i0; i < m.i1; ++i, ur += P) {
        double2 dn = ld2(ur + P + m.jl);
        double2 bb = ld2(b + (size_t)i * P + m.jl);
        double left, right;
        m.sides(g, ur, ce, left, right);
        const int gi = g.row0 + i;
        double2 o = ce;
        if (((gi + colour) & 1) == 0) {       // column j0 (even) has this colour
            o.x = on_bdry(g, gi, m.j0) ? bb.x : smooth_point(bb.x, up.x, left, ce.y, dn.x, g.off, g.diag);
            if (m.has0) ur[m.j0] = o.x;
        } else {
            o.y = on_bdry(g, gi, m.j0 + 1) ? bb.y : smooth_point(bb.y, up.y, ce.x, right, dn.y, g.off, g.diag);
            if (m.has1) ur[m.j0 + 1] = o.y;
        }
        up = o; ce = dn;
    }
}

// ---- residual (solvers.hpp:257-296) -------------------------------------------------------------
// r = b - A u on every owned point; optionally stored; sum r^2 per CTA -> partial[cta]
template <bool STORE>
__global__ void __launch_bounds__(kTPB)
k_residual(LevelGeom g, const double *__restrict__ u, const double *__restrict__ b,
           double *__restrict__ r, double *__restrict__ partial)
{
    __shared__ double red[kTPB / 32];
    March m(g);
    double acc = 0.;
    if (m.i0 < g.rows) {
        const size_t P = g.pitch;
        const double *ur = u + (size_t)m.i0 * P;
        double2 up = ld2(ur - P + m.jl), ce = ld2(ur + m.jl);
#pragma unroll 4
        for (int i = m.i0; i < m.i1; ++i, ur += P) {
            double2 dn = ld2(ur + P + m.jl);
            double2 bb = ld2(b + (size_t)i * P + m.jl);
            double left, right;
            m.sides(g, ur, ce, left, right);
            const int gi = g.row0 + i;
            double2 o;
            // boundary rows are identity rows: sum = 1*u (solvers.hpp:265-266)
            o.x = on_bdry(g, gi, m.j0) ? __dsub_rn(bb.x, ce.x)
                                       : resid_point(bb.x, up.x, left, ce.x, ce.y, dn.x, g.off, g.diag);
            o.y = on_bdry(g, gi, m.j0 + 1) ? __dsub_rn(bb.y, ce.y)
                                           : resid_point(bb.y, up.y, ce.x, ce.y, right, dn.y, g.off, g.diag);
            if (STORE) {
                if (m.has1) st2(r + (size_t)i * P + m.j0, o);
                else if (m.has0) r[(size_t)i * P + m.j0] = o.x;
            }
            if (m.has0) acc += o.x * o.x;
            if (m.has1) acc += o.y * o.y;
            up = ce; ce = dn;
        }
    }
    double t = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = t;
}

// sum of squares of a level vector (solvers.hpp:244-254)
__global__ void __launch_bounds__(kTPB)
k_sumsq(LevelGeom g, const double *__restrict__ v, double *__restrict__ partial)
{
    __shared__ double red[kTPB / 32];
    March m(g);
    double acc = 0.;
    for (int i = m.i0; i < m.i1; ++i) {
        double2 x = ld2(v + (size_t)i * g.pitch + m.jl);
        if (m.has0) acc += x.x * x.x;
        if (m.has1) acc += x.y * x.y;
    }
    double t = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = t;
}

// deterministic second stage: one CTA adds the per-CTA partials in a fixed order.
// out[0] = sum (or out[0] += sum when accumulate).
__global__ void __launch_bounds__(1024)
k_reduce_partials(const double *__restrict__ partial, int n, double *__restrict__ out)
{
    __shared__ double red[32];
    double acc = 0.;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
    double t = block_sum(acc, red);
    if (threadIdx.x == 0) out[0] = t;
}

// out = sum over ranks p = 0..n-1 of (p == rank ? *local : parts[p]), added in rank order: every rank forms the same bits
__global__ void k_sum_ranks(const double *__restrict__ parts, const double *__restrict__ local, int n, int rank, double *__restrict__ out)
{
    if (threadIdx.x != 0) return;
    double acc = 0.;
    for (int p = 0; p < n; ++p) acc += (p == rank) ? local[0] : parts[p];
    out[0] = acc;
}

// ---- restriction of the residual to the next coarser level ------------------------------------------
// MODE 0: injection (what the reference's mask() read amounts to), scale = 1 or 0.5 on interior
//         points (half injection); boundary points are copied unscaled.
// MODE 2: full weighting [1 2 1; 2 4 2; 1 2 1]/16 on interior points, boundary points copied.
// One thread per coarse point; gc = coarse level, gf = fine level (gf.w = 2*gc.w - 1).
template <int MODE>
__global__ void __launch_bounds__(256)
k_restrict(LevelGeom gf, LevelGeom gc, const double *__restrict__ rf, double *__restrict__ rc, double scale)
{
    int J = blockIdx.x * blockDim.x + threadIdx.x;
    int I = blockIdx.y;                         // local coarse row
    if (J >= gc.w || I >= gc.rows) return;
    int gI = gc.row0 + I;
    int fi = 2 * gI - gf.row0;                  // local fine row of the coincident point
    const size_t P = gf.pitch;
    const double *c = rf + (size_t)fi * P + 2 * J;
    double v;
    if (on_bdry(gc, gI, J)) v = c[0];
    else if (MODE == 0) v = __dmul_rn(scale, c[0]);
    else {
        double edge = __dadd_rn(__dadd_rn(__dadd_rn(c[-(ptrdiff_t)P], c[-1]), c[1]), c[P]);
        double corner = __dadd_rn(__dadd_rn(__dadd_rn(c[-(ptrdiff_t)P - 1], c[-(ptrdiff_t)P + 1]), c[P - 1]), c[P + 1]);
        v = __dadd_rn(__dadd_rn(__dmul_rn(0.25, c[0]), __dmul_rn(0.125, edge)), __dmul_rn(0.0625, corner));
    }
    rc[(size_t)I * gc.pitch + J] = v;
}

// ---- bilinear prolongation (multigrid.cpp:3-27), coarse level gc -> fine level gf, overwrite ---
// even row, even col: copy; odd row, even col: 0.5*(N+S); even row, odd col: 0.5*(W+E) of the
// copied values; odd row, odd col: 0.5*(vm_W + vm_E) with vm = the vertical midpoints -- the same
// two-stage order as the reference's in-place loops, hence bit-identical.
// ADD: ef += P ec instead (the coarse-grid correction of the textbook cycles; not in the reference).
template <bool ADD>
__global__ void __launch_bounds__(kTPB)
k_prolong(LevelGeom gc, LevelGeom gf, const double *__restrict__ ec, double *__restrict__ ef)
{
    int j0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);   // even fine column
    int i = blockIdx.y * 4;                                   // 4 fine rows per CTA row-group
    if (j0 >= gf.w) return;
    const int J = j0 >> 1;
    const bool has1 = j0 + 1 < gf.w;
#pragma unroll
    for (int k = 0; k < 4; ++k, ++i) {
        if (i >= gf.rows) return;
        int gi = gf.row0 + i;
        int I = (gi >> 1) - gc.row0;                          // local coarse row (north)
        const double *cn = ec + (size_t)I * gc.pitch + J;
        double a, c;                                          // values at fine cols j0 and j0+2
        if ((gi & 1) == 0) { a = cn[0]; c = has1 ? cn[1] : 0.; }
        else {
            const double *cs = cn + gc.pitch;
            a = __dmul_rn(0.5, __dadd_rn(cn[0], cs[0]));
            c = has1 ? __dmul_rn(0.5, __dadd_rn(cn[1], cs[1])) : 0.;
        }
        double *o = ef + (size_t)i * gf.pitch + j0;
        const double m = has1 ? __dmul_rn(0.5, __dadd_rn(a, c)) : 0.;
        if (ADD) {
            if (has1) { const double2 old = ld2(o); st2(o, make_double2(__dadd_rn(old.x, a), __dadd_rn(old.y, m))); }
            else o[0] = __dadd_rn(o[0], a);
        } else {
            if (has1) st2(o, make_double2(a, m));
            else o[0] = a;
        }
    }
}

// ---- u += e (multigrid.hpp:141-144) -----------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_axpy_rows(LevelGeom g, double *__restrict__ u, const double *__restrict__ e)
{
    int j0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (j0 >= g.pitch) return;
    for (int i = blockIdx.y; i < g.rows; i += gridDim.y) {
        size_t o = (size_t)i * g.pitch + j0;
        double2 a = ld2(u + o), b = ld2(e + o);
        a.x = __dadd_rn(a.x, b.x); a.y = __dadd_rn(a.y, b.y);
        st2(u + o, a);
    }
}

// ---- rows of a pitched level vector packed back to back (w doubles per row): a contiguous device-to-host copy follows ---
__global__ void __launch_bounds__(256)
k_pack_rows(LevelGeom g, const double *__restrict__ v, double *__restrict__ packed)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= g.w) return;
    for (int i = blockIdx.y; i < g.rows; i += gridDim.y) packed[(size_t)i * g.w + j] = v[(size_t)i * g.pitch + j];
}

// ---- lexicographic Gauss-Seidel, exact order (solvers.hpp:33-48) -------------------------------------
// Parity mode.  One CTA sweeps a band of <= 1024 rows as a skewed wavefront: thread t owns row
// band0+t and at step s updates column s-t, so (i-1,j) and (i,j-1) are already new and (i,j+1),
// (i+1,j) still old -- exactly the lexicographic dependency pattern.  Bands run as consecutive
// launches.  The result is bit-identical to the reference's serial loop.
__global__ void __launch_bounds__(1024)
k_gs_lex_band(LevelGeom g, double *u, const double *__restrict__ b, int band0, int band_rows)
{
    const int t = threadIdx.x;
    const int i = band0 + t;
    const bool rowok = t < band_rows;
    const int gi = g.row0 + i;
    const bool brow = gi == 0 || gi == g.w - 1;
    const size_t P = g.pitch;
    double *ur = u + (size_t)i * P;
    const double *br = b + (size_t)i * P;
    double left = 0.;
    const int nsteps = g.w + band_rows - 1;
    for (int s = 0; s < nsteps; ++s) {
        const int j = s - t;
        if (rowok && j >= 0 && j < g.w) {
            double bv = br[j], nv;
            if (brow || j == 0 || j == g.w - 1) nv = bv;      // (b - 0) / 1
            else {
                double up = ur[j - (ptrdiff_t)P];
                nv = smooth_point(bv, up, left, ur[j + 1], ur[j + P], g.off, g.diag);
            }
            ur[j] = nv;
            left = nv;
        }
        __syncthreads();
    }
}

// ---- 64-bit checksum of a level vector: wrap-around sum of the bit patterns of the owned points -----------------
// Integer addition is associative, so the value does not depend on the launch geometry, on the slab
// partition or on the order of the atomics: equal checksums <=> (up to collisions) bit-identical vectors.
__global__ void __launch_bounds__(256)
k_checksum(LevelGeom g, const double *__restrict__ v, unsigned long long *out)
{
    unsigned long long acc = 0ull;
    for (int i = blockIdx.y; i < g.rows; i += gridDim.y)
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < g.w; j += gridDim.x * blockDim.x) {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(v[(size_t)i * g.pitch + j]);
            // position-dependent mixing, so that a permutation of the values changes the sum
            acc += bits * (2ull * (unsigned long long)((size_t)(g.row0 + i) * g.w + j) + 1ull);
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

// ---- rhs sampling on device (linear_system.hpp:85-92 with utilities.cpp:138-147) -------------------
__global__ void __launch_bounds__(256)
k_sample_rhs(LevelGeom g, double *__restrict__ b, double length, int test)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int i = blockIdx.y;
    if (j >= g.w || i >= g.rows) return;
    int gi = g.row0 + i;
    double h = length / (double)(g.w - 1);
    double x = (double)j * h, y = length - (double)gi * h;
    bool bd = on_bdry(g, gi, j);
    double v;
    if (test == 1) { double e = exp(x) * exp(-2.0 * y); v = bd ? e : -5.0 * e; }
    else if (test == 2) {
        double rr = sqrt(x * x + y * y);
        v = bd ? sin(30. * rr) : (rr != 0.0 ? -30. * (cos(30. * rr) / rr - 30. * sin(30. * rr)) : 0.0);
    } else v = bd ? 0. : 1.;
    b[(size_t)i * g.pitch + j] = v;
}


// =====================================================================================================
// Streaming red-black Gauss-Seidel with temporal blocking: S half-sweeps (S/2 full sweeps) in ONE pass
// over HBM -- 24 B per point for the whole group of sweeps instead of 24 B per sweep.
//
// A CTA owns a strip of OW = TW - 2S columns (plus S halo columns on each side, recomputed
// redundantly) and marches down a chunk of rows.  Shared memory holds a ring of 2S+3 rows of u
// and of the rhs.  When row i arrives, half-sweep s (s = 1..S; odd s = red, even s = black) is
// applied to row i-2s.  With a lag of TWO rows per half-sweep all S updates of a step read only
// values produced in earlier steps, so one __syncthreads per row suffices and the S updates of a
// thread are independent (ILP).  Row i-2S is final after the step and is written out.
// Red-black GS does not depend on the traversal order inside a colour, so the result is identical
// -- bit for bit with EXACT arithmetic -- to S/2 plain red-then-black sweeps.
//
// Shared-memory rows are stored de-interleaved (even columns | odd columns) so that a half-sweep,
// which touches every other point, reads and writes CONTIGUOUS doubles (no bank conflicts).
// Input and output arrays differ (neighbouring CTAs re-read the halo from the input).
// =====================================================================================================
constexpr int kStreamTW = 256;            // tile columns held in shared memory
constexpr int kStreamNT = kStreamTW / 2;  // threads: one per column pair
constexpr int kStreamPF = 4;              // rows in flight from HBM per thread (register staged)
constexpr int kStreamL2Ahead = 24;        // rows ahead that are prefetched into L2 (one 128-B line per 8 threads)

__device__ __forceinline__ void prefetch_l2(const void *p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

template <int S>
constexpr int stream_smem_bytes() { return (2 * S + 3) * kStreamTW * 2 * (int)sizeof(double); }

// fast arithmetic: off/diag == -1/4 exactly, so u = b/diag + (up+down+left+right)/4
template <bool EXACT>
__device__ __forceinline__ double rb_point(double b, double up, double left, double right, double down,
                                           double off, double diag, double inv_diag)
{
    if (EXACT) return smooth_point(b, up, left, right, down, off, diag);
    return fma(b, inv_diag, 0.25 * ((up + down) + (left + right)));
}

// One step of the stream: row i arrives in `nu`/`nb`; half-sweep s is applied to row i-2s.
// The thread keeps ITS OWN column pair of the whole row window in registers (uw[d] = row i-d), so
// up, down and the neighbour inside the pair never touch shared memory; only the neighbour that
// belongs to the adjacent thread, and the rhs, are read from the ring, and the new value is
// published to the ring for the adjacent threads.  3 shared-memory words per update instead of 6.
// ro[s] = element offset of [ring slot of row i-2s][t] (ro[0]: the arriving row), carried from
// step to step so that no index arithmetic is left in the loop.
// PAR = parity of the global row index of row i (static: the main loop is unrolled by 4 rows).
// GUARD = false is the steady state: every half-sweep row is an interior row inside the
// streamed range, so the only predicate left is the per-thread Dirichlet-column flag.  Threads
// at the tile edge / outside the domain compute harmless garbage there: tile-edge columns are
// halo (invalid after s half-sweeps by construction) and the Dirichlet columns j=0, j=w-1 never
// read their neighbours, so nothing crosses into the domain.
template <int S, bool EXACT, int PAR, bool GUARD, int BATCH>
__device__ __forceinline__ void rb_stream_step(
    const LevelGeom &g, double2 (&uw)[2 * S + 3], int (&ro)[S + 1], double2 nu, double2 nb, double *su,
    double *sb, int i, bool bc0, bool bc1, int ifirst, int ilast, bool first_is_bdry,
    bool last_is_bdry, int glast, double inv_diag)
{
    constexpr int TW = kStreamTW, H = TW / 2, WR = 2 * S + 3;
    // fast arithmetic: the ring holds bq = b/diag (or b itself on Dirichlet points) and the update is
    // u = bq + q * (sum of neighbours) with q = 1/4 (0 on Dirichlet points): 3 DADD + 1 DFMA, no select.
    const double q0 = bc0 ? 0. : 0.25, q1 = bc1 ? 0. : 0.25;
    if (!EXACT) {
        const bool brow = (i + g.row0 == 0) || (i == glast);   // the ARRIVING row may be a Dirichlet row in any step
        nb.x = (bc0 || brow) ? nb.x : nb.x * inv_diag;
        nb.y = (bc1 || brow) ? nb.y : nb.y * inv_diag;
    }
    // rotate the register window by one row and publish the arriving row
#pragma unroll
    for (int d = 2 * S + 2; d > 0; --d) uw[d] = uw[d - 1];
    su[ro[0]] = nu.x; su[ro[0] + H] = nu.y;
    sb[ro[0]] = nb.x; sb[ro[0] + H] = nb.y;
    __syncthreads();
    // the window copy of the arriving row is re-read from the ring rather than kept from the global
    // load: the load registers then die here and the next prefetch can land in them without a
    // dependent register move at the loop back-edge
    uw[0] = make_double2(su[ro[0]], su[ro[0] + H]);
    // the S half-sweep updates of a step are mutually independent: operands of a whole batch are read first
    // (shared-memory latency paid once per batch), then computed, then published
#pragma unroll
    for (int s0 = 1; s0 <= S; s0 += BATCH) {
        double bv[BATCH], ob[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
            const int s = s0 + k;
            if (s > S) break;
            const int which = (PAR + (s - 1)) & 1;        // rows i-2s have the parity of row i
            bv[k] = sb[ro[s] + which * H];
            // neighbour held by the adjacent thread: even column 2t -> odd[t-1]; odd column 2t+1 -> even[t+1]
            ob[k] = su[ro[s] + (which ? 1 : (H - 1))];
        }
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
            const int s = s0 + k;
            if (s > S) break;
            const int which = (PAR + (s - 1)) & 1;
            const int d = 2 * s;
            bool act = true;
            bool isb = which ? bc1 : bc0;
            bool brow = false;
            if (GUARD) {
                const int r = i - d;
                const int vlo = first_is_bdry ? ifirst : ifirst + s;
                const int vhi = last_is_bdry ? ilast : ilast - s;
                act = (r >= vlo) && (r <= vhi);
                brow = (r + g.row0 == 0) || (r == glast);
                isb = isb || brow;
            }
            const double up = which ? uw[d + 1].y : uw[d + 1].x;
            const double dn = which ? uw[d - 1].y : uw[d - 1].x;
            const double left = which ? uw[d].x : ob[k];
            const double right = which ? ob[k] : uw[d].y;
            double nv;
            if (EXACT) {
                nv = smooth_point(bv[k], up, left, right, dn, g.off, g.diag);
                nv = isb ? bv[k] : nv;
            } else {
                double q = which ? q1 : q0;
                if (GUARD) q = brow ? 0. : q;
                nv = fma(q, (up + dn) + (left + right), bv[k]);
            }
            if (act) {
                if (which) uw[d].y = nv; else uw[d].x = nv;
                su[ro[s] + which * H] = nv;
            }
        }
    }
    // advance every carried ring offset by one row
#pragma unroll
    for (int s = 0; s <= S; ++s) {
        ro[s] += TW;
        if (ro[s] >= WR * TW) ro[s] -= WR * TW;
    }
}

// MODE 0: uout = S/2 sweeps applied to uin.
// MODE 1 (last post-smoothing launch on the fine level): the smoothed error is not written out; instead
//   ucorr += e            (the correction u += err of multigrid.hpp:141-144, in place on the owned rows), and
//   partial[cta] = sum (rhs - A e)^2   over the CTA's output rows.
// Since rhs = f - A u_old, rhs - A e = f - A (u_old + e): this is the residual of the NEW iterate (main.cpp:86),
// obtained from data already on chip -- no separate axpy pass and no separate residual pass over HBM.
// It needs the final values of one more row on each side, so the stream starts/ends one row further out.
// MODE 2 (the driver's pre-sweeps): uout = the smoothed u AND ucorr(:= the residual array) = rhs - A uout on the
//   output rows, i.e. the fine residual of multigrid.hpp:127 without a separate pass that re-reads u and f.
// MODE 3 = MODE 2 + the restriction of that residual to the next coarser level (geometry gc, array `partial`) in
//   the same pass (north_star: "residual+restriction fused into a single pass"): full weighting
//   [1 2 1; 2 4 2; 1 2 1]/16 (restr = 2) or (half) injection (restr = 0, scale), evaluated in the term order of
//   k_restrict, so the coarse residual is bit-identical to the separate kernel's.  The 3x3 neighbourhood comes from a
//   4-row register window of the thread's own residual pairs plus, for column j0-1, the adjacent thread's value
//   published through a 4-row shared ring one step earlier.
// PIN (prolongation fused into the input): uin is the COARSER level's solution (geometry gc) and the rows that
//   arrive are its bilinear interpolation, evaluated in the two-stage order of multigrid.cpp:3-27 (bit-identical
//   to k_prolong); the prolonged field is never written to HBM.
template <int S, bool EXACT, int MODE, bool PIN>
__global__ void __launch_bounds__(kStreamNT)
k_rb_stream(LevelGeom g, const double *__restrict__ uin, const double *__restrict__ b,
            double *__restrict__ uout, int rows_per_chunk, double *ucorr, double *__restrict__ partial, LevelGeom gc,
            int restr, double rscale)
{
    constexpr int TW = kStreamTW, PF = kStreamPF, H = TW / 2;
    constexpr int WR = 2 * S + 3;
    constexpr int X = (MODE == 3) ? 2 : ((MODE != 0) ? 1 : 0);
    constexpr int BATCH = (MODE == 1 && S > 5) ? 5 : S;     // MODE 1 is register-bound: operands in two batches
    static_assert(PF == 4, "the main loop is unrolled by 4 rows");
    extern __shared__ double smem[];
    double *su = smem;                 // [WR][TW]: [slot][0..H) even columns, [slot][H..TW) odd columns
    double *sb = smem + WR * TW;
    __shared__ double red[kStreamNT / 32];
    __shared__ double sr[(MODE == 3) ? 8 * kStreamNT : 1];     // MODE 3: residual pairs of the last 4 rows ([row][even|odd][t])

    const int t = threadIdx.x;
    // halo columns per side: S for the S half-sweeps; MODE 1 reads FINAL values of the lateral neighbours, i.e. one
    // more valid column, and a column pair is the unit of ownership
    constexpr int HC = S + 2 * (X > 0);
    const int OW = TW - 2 * HC;
    const int jbase = blockIdx.x * OW - HC;          // global column of tile column 0 (even)
    const int j0 = jbase + 2 * t;                    // this thread's even column
    const int jl = min(max(j0, 0), g.pitch - 2);     // clamped for loads
    const int i0 = blockIdx.y * rows_per_chunk;      // rows_per_chunk is even (host)
    const int i1 = min(i0 + rows_per_chunk, g.rows);
    if (i0 >= g.rows) return;
    const bool top_is_domain = (g.row0 == 0), bot_is_domain = (g.row0 + g.rows == g.w);
    const int lo = top_is_domain ? 0 : -S - X, hi = bot_is_domain ? g.rows - 1 : g.rows - 1 + S + X;
    int ifirst = max(i0 - S - X, lo);
    // the unrolled loop assumes the first streamed row has an even global index; if a slab starts on
    // an odd row, stream one more (halo) row -- it only feeds values that are never written out
    if ((g.row0 + ifirst) & 1) ifirst -= 1;
    const int ilast = min(i1 - 1 + S + X, hi);
    const bool first_is_bdry = (g.row0 + ifirst == 0), last_is_bdry = (g.row0 + ilast == g.w - 1);
    const ptrdiff_t P = g.pitch;
    const double inv_diag = 1.0 / g.diag;
    const bool own = (2 * t >= HC) && (2 * t < TW - HC) && (j0 < g.w) && (j0 >= 0);
    // Dirichlet-column flags of the pair; columns outside the domain are frozen the same way
    const bool bc0 = (j0 <= 0) || (j0 >= g.w - 1);
    const bool bc1 = (j0 + 1 <= 0) || (j0 + 1 >= g.w - 1);
    const int glast = g.w - 1 - g.row0;              // local index of the global last row
    const int i_lo = ifirst + (first_is_bdry ? 2 * S + 1 : 3 * S), i_hi = ilast + 1;   // steady steps

    double2 uw[2 * S + 3];
#pragma unroll
    for (int d = 0; d < 2 * S + 3; ++d) uw[d] = make_double2(0., 0.);
    double2 pu[PF], pb[PF], ps[PF];                  // ps: second coarse row of an odd fine row (PIN only)
    // PIN: the two coarse columns this thread's fine pair interpolates from
    const int Jc = min(max(j0 >> 1, 0), gc.w - 1), Jc1 = min(Jc + 1, gc.w - 1);
    const ptrdiff_t Pc = gc.pitch;
    auto fetch = [&](int r, double2 &cu, double2 &cs, double2 &cb) {
        cb = ld2(b + (ptrdiff_t)r * P + jl);
        if (!PIN) { cu = ld2(uin + (ptrdiff_t)r * P + jl); return; }
        const int gi = g.row0 + r;
        const double *cn = uin + (ptrdiff_t)((gi >> 1) - gc.row0) * Pc;
        cu = make_double2(cn[Jc], cn[Jc1]);
        if (gi & 1) cs = make_double2(cn[Pc + Jc], cn[Pc + Jc1]);
    };
    // the fine pair (columns j0, j0+1) of global row gi from the coarse values (multigrid.cpp:3-27)
    auto interpolate = [&](int gi, double2 cu, double2 cs) -> double2 {
        double a = cu.x, c2 = cu.y;
        if (gi & 1) { a = __dmul_rn(0.5, __dadd_rn(cu.x, cs.x)); c2 = __dmul_rn(0.5, __dadd_rn(cu.y, cs.y)); }
        return make_double2(a, __dmul_rn(0.5, __dadd_rn(a, c2)));
    };
#pragma unroll
    for (int p = 0; p < PF; ++p) fetch(min(ifirst + p, ilast), pu[p], ps[p], pb[p]);
    const int ksteps = (i1 - 1 + 2 * S + X) - ifirst + 1 + (MODE == 3 ? 1 : 0);
    int ro[S + 1];                                   // ring offsets of rows i-2s (s = 0: arriving row)
#pragma unroll
    for (int s = 0; s <= S; ++s) ro[s] = ((WR - 2 * s) % WR) * TW + t;
    int roq = ((WR - (2 * S + 1)) % WR) * TW + t;    // ring offset of row i-2S-1 (MODE 1: the residual row)
    double2 pc[PF];                                  // MODE 1: rows of ucorr, requested PF steps ahead
    double acc = 0.;
    if (MODE == 1) {
#pragma unroll
        for (int p = 0; p < PF; ++p) {
            int r = min(max(ifirst + p - 2 * S, i0), i1 - 1);
            pc[p] = ld2(ucorr + (ptrdiff_t)r * P + jl);
        }
    }
#define MGB_STREAM_STEP(p)                                                                                  \
    {                                                                                                       \
        const int i = ifirst + k0 + (p);             /* row arriving at this step */                         \
        const double2 nu = PIN ? interpolate(g.row0 + min(i, ilast), pu[p], ps[p]) : pu[p], nb = pb[p];     \
        fetch(min(i + PF, ilast), pu[p], ps[p], pb[p]);                                                     \
        const bool steady = (i >= i_lo) && (i <= i_hi);                                                     \
        if (steady)                                                                                         \
            rb_stream_step<S, EXACT, ((p) & 1), false, BATCH>(g, uw, ro, nu, nb, su, sb, i, bc0, bc1, ifirst, \
                                                       ilast, first_is_bdry, last_is_bdry, glast, inv_diag); \
        else                                                                                                \
            rb_stream_step<S, EXACT, ((p) & 1), true, BATCH>(g, uw, ro, nu, nb, su, sb, i, bc0, bc1, ifirst, \
                                                      ilast, first_is_bdry, last_is_bdry, glast, inv_diag); \
        const int r = i - 2 * S;                     /* final after this step */                             \
        if (MODE == 0 || MODE == 2 || MODE == 3) {                                                          \
            if (r >= i0 && r < i1 && own) {                                                                 \
                double *dstp = uout + (ptrdiff_t)r * P + j0;                                                \
                if (j0 + 1 < g.w) st2(dstp, uw[2 * S]); else dstp[0] = uw[2 * S].x;                         \
            }                                                                                               \
        }                                                                                                   \
        if (MODE == 2 || MODE == 3) {                                                                       \
            const int q = r - 1;                     /* rows q-1, q, q+1 are final: residual of row q */     \
            /* MODE 3 needs the residual one row beyond the output rows and in the lateral halo pair as well */ \
            const int fin_lo = first_is_bdry ? ifirst : ifirst + S, fin_hi = last_is_bdry ? ilast : ilast - S; /* final rows */ \
            const bool rows_ok = (MODE == 3) ? (q >= i0 - 1 && q <= i1 && q + g.row0 >= 0 && q <= glast &&   \
                                                (q - 1 >= fin_lo || q + g.row0 == 0) && (q + 1 <= fin_hi || q == glast)) \
                                             : (q >= i0 && q < i1);                                        \
            const bool cols_ok = (MODE == 3) ? (2 * t >= HC - 2 && 2 * t < TW - HC + 2 && j0 >= 0 && j0 < g.w) : own; \
            double2 rv = make_double2(0., 0.);                                                              \
            if (rows_ok && cols_ok) {                                                                       \
                const double2 c = uw[2 * S + 1], up = uw[2 * S + 2], dn = uw[2 * S];                        \
                const double lf = su[roq + H - 1], rt = su[roq + 1];                                        \
                const double b0 = sb[roq], b1 = sb[roq + H];                                                \
                const bool brow = (q + g.row0 == 0) || (q == glast);                                        \
                if (EXACT) {                                                                                \
                    rv.x = (bc0 || brow) ? __dsub_rn(b0, c.x) : resid_point(b0, up.x, lf, c.x, c.y, dn.x, g.off, g.diag); \
                    rv.y = (bc1 || brow) ? __dsub_rn(b1, c.y) : resid_point(b1, up.y, c.x, c.y, rt, dn.y, g.off, g.diag); \
                } else {                                                                                    \
                    /* the ring holds b/diag (b itself on Dirichlet points): r = diag*(b/diag - u + sum/4) */ \
                    const double q0 = (bc0 || brow) ? 0. : 0.25, q1 = (bc1 || brow) ? 0. : 0.25;            \
                    const double w0 = (bc0 || brow) ? 1. : g.diag, w1 = (bc1 || brow) ? 1. : g.diag;        \
                    rv.x = w0 * fma(q0, (up.x + dn.x) + (lf + c.y), b0 - c.x);                              \
                    rv.y = w1 * fma(q1, (up.y + dn.y) + (c.x + rt), b1 - c.y);                              \
                }                                                                                           \
                if (q >= i0 && q < i1 && own) {                                                             \
                    double *dstp = ucorr + (ptrdiff_t)q * P + j0;                                           \
                    if (j0 + 1 < g.w) st2(dstp, rv); else dstp[0] = rv.x;                                   \
                }                                                                                           \
            }                                                                                               \
            if (MODE == 3) {                                                                                \
                constexpr int NT2 = 2 * kStreamNT;                                                          \
                const int k4 = (p) & 3;              /* ring slot of row q: k0 is a multiple of 4 */          \
                sr[k4 * NT2 + t] = rv.x; sr[k4 * NT2 + kStreamNT + t] = rv.y;                               \
                /* coarse row centred on fine row c = q-2 (rows c-1, c, c+1 were published in earlier steps) */ \
                const int cq = q - 2, gq = g.row0 + cq;                                                     \
                if (((gq & 1) == 0) && cq >= i0 && cq < i1 && own) {                                        \
                    const int gI = gq >> 1, J = j0 >> 1;                                                    \
                    const double *rm = sr + ((k4 + 1) & 3) * NT2 + t;      /* row c-1 = q-3 */                \
                    const double *rc = sr + ((k4 + 2) & 3) * NT2 + t;      /* row c   = q-2 */                \
                    const double *rp = sr + ((k4 + 3) & 3) * NT2 + t;      /* row c+1 = q-1 */                \
                    const double c0 = rc[0];                                                                \
                    double v;                                                                               \
                    if (gI == 0 || gI == gc.w - 1 || J == 0 || J == gc.w - 1) v = c0;                       \
                    else if (restr != 2) v = __dmul_rn(rscale, c0);                                         \
                    else {   /* [x][t] = column j0, [y][t] = column j0+1, [y][t-1] = column j0-1 */           \
                        const double edge = __dadd_rn(__dadd_rn(__dadd_rn(rm[0], rc[kStreamNT - 1]), rc[kStreamNT]), rp[0]); \
                        const double corner = __dadd_rn(__dadd_rn(__dadd_rn(rm[kStreamNT - 1], rm[kStreamNT]), rp[kStreamNT - 1]), rp[kStreamNT]); \
                        v = __dadd_rn(__dadd_rn(__dmul_rn(0.25, c0), __dmul_rn(0.125, edge)), __dmul_rn(0.0625, corner)); \
                    }                                                                                       \
                    partial[(ptrdiff_t)(gI - gc.row0) * gc.pitch + J] = v;                                  \
                }                                                                                           \
            }                                                                                               \
            roq += TW; if (roq >= WR * TW) roq -= WR * TW;                                                  \
        }                                                                                                   \
        if (MODE == 1) {                                                                                            \
            const double2 uc = pc[p];                                                                       \
            {                                                                                               \
                int r2 = min(max(r + PF, i0), i1 - 1);                                                      \
                pc[p] = ld2(ucorr + (ptrdiff_t)r2 * P + jl);                                                \
            }                                                                                               \
            if (r >= i0 && r < i1 && own) {                                                                 \
                double *dstp = ucorr + (ptrdiff_t)r * P + j0;                                               \
                const double2 un = make_double2(__dadd_rn(uc.x, uw[2 * S].x), __dadd_rn(uc.y, uw[2 * S].y)); \
                if (j0 + 1 < g.w) st2(dstp, un); else dstp[0] = un.x;                                       \
            }                                                                                               \
            const int q = r - 1;                     /* rows q-1, q, q+1 are final: residual of row q */     \
            /* The black point of the pair was the last one updated and nothing around it changed since, so its   \
               residual is a rounding error of its own update (~1e-16 of the terms) and Dirichlet points have     \
               residual 0 exactly: only the RED interior point of the pair contributes to the sum. */              \
            const bool brow = (q + g.row0 == 0) || (q == glast);                                            \
            if (q >= i0 && q < i1 && own && !brow) {                                                        \
                const double2 c = uw[2 * S + 1], up = uw[2 * S + 2], dn = uw[2 * S];                        \
                constexpr int RED = ((p) & 1) ^ 1;   /* row q has the other parity than row i: red = column of that parity */ \
                double rv;                                                                                  \
                if (RED == 0) {                                                                             \
                    const double lf = su[roq + H - 1], b0 = sb[roq];                                        \
                    if (EXACT) rv = bc0 ? 0. : resid_point(b0, up.x, lf, c.x, c.y, dn.x, g.off, g.diag);     \
                    else rv = bc0 ? 0. : g.diag * fma(0.25, (up.x + dn.x) + (lf + c.y), b0 - c.x);           \
                } else {                                                                                    \
                    const double rt = su[roq + 1], b1 = sb[roq + H];                                        \
                    if (EXACT) rv = bc1 ? 0. : resid_point(b1, up.y, c.x, c.y, rt, dn.y, g.off, g.diag);     \
                    else rv = bc1 ? 0. : g.diag * fma(0.25, (up.y + dn.y) + (c.x + rt), b1 - c.y);           \
                }                                                                                           \
                acc += rv * rv;                                                                             \
            }                                                                                               \
            roq += TW; if (roq >= WR * TW) roq -= WR * TW;                                                  \
        }                                                                                                   \
    }
    for (int k0 = 0; k0 < ksteps; k0 += PF) {
        MGB_STREAM_STEP(0)
        MGB_STREAM_STEP(1)
        MGB_STREAM_STEP(2)
        MGB_STREAM_STEP(3)
    }
#undef MGB_STREAM_STEP
    if (MODE == 1) {
        const double tsum = block_sum(acc, red);
        if (t == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = tsum;
    }
}

}  // namespace mgb
