// amg_kernels.cuh -- sm_100a kernels of the AMG solve phase on device-resident CSR levels.
//
// Storage per level (HBM, SoA): int32 row_ptr[n+1], int32 col[nnz], fp64 val[nnz], fp64 diag[n]
// (the reference keeps an AoS pair<size_t,double>[nnz] and finds a_ii by a linear scan on every row
// visit, AMG/src/CSRMatrix.cpp:24-52); the transfer operator is stored twice, P (n x nc) for
// x_f += P x_c and R = P^T (nc x n, rows sorted by fine index) so that restriction is a gather
// SpMV without atomics.
//
// Two families of kernels:
//  * exact-order kernels (one thread per row, terms added in ascending column order with unfused
//    IEEE operations): bit-identical to the reference's serial loops -- the parity path;
//  * vector kernels (a sub-warp of kLanes lanes per row, __shfl_xor tree reduction): the fast path,
//    equal to the exact ones up to summation order.
// Lexicographic Gauss-Seidel is reproduced EXACTLY by level scheduling (rows grouped into wavefronts
// whose members do not depend on one another); the reordered fast smoother is multicolour
// Gauss-Seidel from an on-device greedy (Jones-Plassmann) colouring.
#pragma once
#include <cooperative_groups.h>
#include "p2p.cuh"
#include <cuda_runtime.h>
#include <stdint.h>

#include "amg_types.h"

namespace mgb {

constexpr int kLanes = 8;        // lanes per row in the vector kernels (P1 rows hold ~7 entries)

__device__ __forceinline__ double subwarp_sum(double v)
{
#pragma unroll
    for (int o = kLanes / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, kLanes);
    return v;
}

// ---- Gauss-Seidel row update, reference order (AMG/include/Utilities.hpp:44-58) --------------------------
// (plain pointers: the tail kernel of amg_tail.cuh calls these row functions on vectors it also writes)
__device__ __forceinline__ void gs_row_exact(const CsrDev &A, const double *diag, double *x, const double *b, int i)
{
    double sum = 0.;
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
        const int j = A.col[k];
        if (j != i) sum = __dadd_rn(sum, __dmul_rn(A.val[k], x[j]));
    }
    x[i] = __ddiv_rn(__dsub_rn(b[i], sum), diag[i]);
}

// ---- row functions shared by the per-level kernels and the persistent tail (amg_tail.cuh) ------------------------------
// fast arithmetic, one thread per row: off-diagonal entries in ascending column order, fused multiply-add -- the
// operation sequence of the SELL kernels below
__device__ __forceinline__ double offdiag_dot_fast(const CsrDev &A, const double *x, int i)
{
    double sum = 0.;
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
        const int j = A.col[k];
        if (j != i) sum += A.val[k] * x[j];
    }
    return sum;
}
// kLanes lanes per row (lane l takes entries l, l + kLanes, ...), tree reduction; every lane of the warp must call
__device__ __forceinline__ double offdiag_dot_vec(const CsrDev &A, const double *x, int i, int lane, bool ok)
{
    double sum = 0.;
    if (ok)
        for (int k = A.ptr[i] + lane; k < A.ptr[i + 1]; k += kLanes) {
            const int j = A.col[k];
            if (j != i) sum += A.val[k] * x[j];
        }
    return subwarp_sum(sum);
}
__device__ __forceinline__ double row_dot_vec(const CsrDev &A, const double *x, int i, int lane, bool ok)
{
    double sum = 0.;
    if (ok)
        for (int k = A.ptr[i] + lane; k < A.ptr[i + 1]; k += kLanes) sum += A.val[k] * x[A.col[k]];
    return subwarp_sum(sum);
}
// one thread per row, every entry in ascending column order, fused multiply-add (k_amg_sell<3>)
__device__ __forceinline__ double row_dot_fast(const CsrDev &A, const double *x, int i)
{
    double sum = 0.;
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) sum += A.val[k] * x[A.col[k]];
    return sum;
}
// reference order, unfused IEEE operations
__device__ __forceinline__ double residual_row_exact(const CsrDev &A, const double *x, const double *b, int i)
{
    double Ax = 0.;
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) Ax = __dadd_rn(Ax, __dmul_rn(A.val[k], x[A.col[k]]));
    return __dsub_rn(b[i], Ax);
}
__device__ __forceinline__ double spmv_row_exact(const CsrDev &R, const double *xin, int m)
{
    double s = 0.;
    for (int k = R.ptr[m]; k < R.ptr[m + 1]; ++k) s = __dadd_rn(s, __dmul_rn(R.val[k], xin[R.col[k]]));
    return s;
}
// x_f[i] += sum_k P(i,k) x_c[k], term by term INTO x_f[i] as the reference does (AMG/src/AMG.cpp:218-232)
__device__ __forceinline__ void prolong_row_add(const CsrDev &P, const double *xc, double *xf, int i)
{
    double v = xf[i];
    for (int k = P.ptr[i]; k < P.ptr[i + 1]; ++k) v = __dadd_rn(v, __dmul_rn(P.val[k], xc[P.col[k]]));
    xf[i] = v;
}
__device__ __forceinline__ double relax(double xhat, double xold, double omega)
{
    return omega == 1. ? xhat : xold + omega * (xhat - xold);
}

// rows[first..last) are mutually independent (one wavefront of the level schedule, or one colour)
__global__ void __launch_bounds__(256)
k_amg_gs_rows_exact(CsrDev A, const double *__restrict__ diag, double *x, const double *__restrict__ b,
                    const int *__restrict__ rows, int first, int last)
{
    int t = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (t < last) gs_row_exact(A, diag, x, b, rows[t]);
}

// whole lexicographic sweeps by ONE CTA: loops over the wavefronts with a barrier in between, so a
// sweep costs no launch per wavefront (parity mode on small and medium levels)
__global__ void __launch_bounds__(1024)
k_amg_gs_lex_cta(CsrDev A, const double *__restrict__ diag, double *x, const double *__restrict__ b,
                 const int *__restrict__ wave_ptr, const int *__restrict__ wave_rows, int n_waves, int sweeps)
{
    for (int s = 0; s < sweeps; ++s)
        for (int w = 0; w < n_waves; ++w) {
            for (int t = wave_ptr[w] + threadIdx.x; t < wave_ptr[w + 1]; t += blockDim.x)
                gs_row_exact(A, diag, x, b, wave_rows[t]);
            __syncthreads();
        }
}

// multicolour GS, vector form: kLanes lanes per row of one colour
__global__ void __launch_bounds__(256)
k_amg_gs_color_vec(CsrDev A, const double *__restrict__ diag, double *x, const double *__restrict__ b,
                   const int *__restrict__ rows, int first, int last)
{
    const int lane = threadIdx.x & (kLanes - 1);
    const int t = first + (blockIdx.x * blockDim.x + threadIdx.x) / kLanes;
    const bool ok = t < last;
    const int i = ok ? rows[t] : 0;
    double sum = 0.;
    if (ok)
        for (int k = A.ptr[i] + lane; k < A.ptr[i + 1]; k += kLanes) {
            const int j = A.col[k];
            if (j != i) sum += A.val[k] * x[j];
        }
    sum = subwarp_sum(sum);
    if (ok && lane == 0) x[i] = (b[i] - sum) / diag[i];
}

// weighted Jacobi: xhat = (b - sum_{j != i} a_ij x_j) / a_ii, xnew = x + omega (xhat - x); omega == 1 stores xhat itself
// (the reference's smoothers are unweighted).  Rows [row0, row1): the block this rank owns.
__global__ void __launch_bounds__(256)
k_amg_jacobi_vec(CsrDev A, const double *__restrict__ diag, const double *__restrict__ x, const double *__restrict__ b,
                 double *__restrict__ xnew, double omega, int row0, int row1)
{
    const int lane = threadIdx.x & (kLanes - 1);
    const int i = row0 + (blockIdx.x * blockDim.x + threadIdx.x) / kLanes;
    const bool ok = i < row1;
    const double sum = offdiag_dot_vec(A, x, ok ? i : 0, lane, ok);
    if (ok && lane == 0) xnew[i] = relax((b[i] - sum) / diag[i], x[i], omega);
}

// ---- fast path: sliced-ELLPACK copy of the level operator ---------------------------------------------------------------
// The vector-CSR kernels above are latency-bound on P1 operators (~7 entries per row: row_ptr -> col/val -> x is
// a chain of three dependent loads with one row in flight per sub-warp).  The fast kernels therefore read a second
// copy of A in SELL-32 form: rows are sorted by colour (every colour starts on a slice boundary, so a colour is a
// contiguous range of slices), a slice stores the OFF-DIAGONAL entries of 32 rows column-major, padded to the
// longest row of the slice.  One thread owns one row: its loads of col/val are coalesced across the warp and
// mutually independent (all entries of the row in flight at once), and there is no row_ptr load on the path.
// sum_k val_k * x[col_k] over the `len` stored entries of one slot, in storage (ascending column) order.  The entries
// are fetched in chunks of kSellChunk: all column indices and values of a chunk are in flight together, then all the
// gathers of x -- two dependent memory round trips per chunk.  P1 rows hold 5-8 off-diagonals, i.e. one chunk: the
// row's lifetime is ~3 memory latencies (slice_ptr, col/val, x) instead of ~5 with a 4-wide unroll, which is what
// bounds these kernels (ncu: long-scoreboard stalls, issue slots 13 % busy, DRAM 70 % -- latency, not bandwidth).
constexpr int kSellChunk = 8;
__device__ __forceinline__ double sell_row_dot(const SellDev &A, const double *x, int base, int len)
{
    double sum = 0.;
    for (int k0 = 0; k0 < len; k0 += kSellChunk) {
        int c[kSellChunk];
        double v[kSellChunk];
#pragma unroll
        for (int u = 0; u < kSellChunk; ++u) {
            const bool ok = k0 + u < len;
            c[u] = ok ? __ldcs(A.col + base + 32 * (k0 + u)) : -1;
            v[u] = ok ? __ldcs(A.val + base + 32 * (k0 + u)) : 0.;
        }
        double xv[kSellChunk];
#pragma unroll
        for (int u = 0; u < kSellChunk; ++u) xv[u] = c[u] >= 0 ? x[c[u]] : 0.;
#pragma unroll
        for (int u = 0; u < kSellChunk; ++u)
            if (c[u] >= 0) sum += v[u] * xv[u];
    }
    return sum;
}

// v + sum_k val_k * x[col_k], added term by term into v with unfused IEEE operations (the reference's prolongation)
__device__ __forceinline__ double sell_row_add_exact(const SellDev &A, const double *x, int base, int len, double v)
{
    for (int k0 = 0; k0 < len; k0 += kSellChunk) {
        int c[kSellChunk];
        double w[kSellChunk], xv[kSellChunk];
#pragma unroll
        for (int u = 0; u < kSellChunk; ++u) {
            const bool ok = k0 + u < len;
            c[u] = ok ? __ldcs(A.col + base + 32 * (k0 + u)) : -1;
            w[u] = ok ? __ldcs(A.val + base + 32 * (k0 + u)) : 0.;
        }
#pragma unroll
        for (int u = 0; u < kSellChunk; ++u) xv[u] = c[u] >= 0 ? x[c[u]] : 0.;
#pragma unroll
        for (int u = 0; u < kSellChunk; ++u)
            if (c[u] >= 0) v = __dadd_rn(v, __dmul_rn(w[u], xv[u]));
    }
    return v;
}

// The operator stream (col, val, diag, rhs, row map) is read exactly once per launch: it is loaded with the streaming
// (evict-first) hint so that it does not push the vector x -- gathered ~7 times per sweep, once from every colour --
// out of the 126 MB L2.  ncu before the hint: 793 MB of DRAM traffic for a 447 MB (algorithmic) Jacobi sweep, L2 hit
// rate 25 % (profiles/r01_ncu_amg_ops_4M.txt).
// MODE 0: r = b - A x (+ sum r^2 per CTA); MODE 1: Jacobi into `out`; MODE 2: Gauss-Seidel on the slots
// [first, last) of one colour, in place on x; MODE 3: plain product out[row] = sum (restriction, R = P^T);
// MODE 4: out[row] += sum with the reference's term order and unfused operations (prolongation x_f += P x_c,
// AMG/src/AMG.cpp:218-232 -- bit-identical to k_amg_prolong_add)
// MODE 5: l1-Jacobi into `out`: x + (b - A x) / dl1 with dl1_i = a_ii + sum_{j != i} |a_ij| (no damping parameter, no
// colouring: the smoother of the Galerkin levels, whose 13-16 colours make multicolour GS launch-bound)
// NAT: the copy keeps the natural row order (only permuted inside 512-row windows) and has no slot-ordered vectors:
// diag_s / b_s are then the level's diag / b, indexed by row.  Jacobi and the residual visit every row once in any
// order; in natural order the gathers of x stay local, while the colour-sorted copy walks the whole index range once
// per colour (7 passes over x -- 54 % of peak instead of 77 % once x outgrows the L2, 16 M DoF).
template <int MODE, bool NAT = false>
__global__ void __launch_bounds__(256)
k_amg_sell(SellDev A, const double *x, const double *__restrict__ b_s, double *out, double *__restrict__ partial,
           int first, int last, double omega, const double *__restrict__ dl1)
{
    __shared__ double red[8];
    const int p = first + blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.;
    if (p < last) {
        const int i = __ldcs(A.row_of_slot + p);
        if (i >= 0) {
            const int s = p >> 5;
            const int b0 = A.slice_ptr[s];
            const int base = b0 + (p & 31);
            const int len = (A.slice_ptr[s + 1] - b0) >> 5;
            if (MODE == 4) { out[i] = sell_row_add_exact(A, x, base, len, out[i]); return; }
            const double sum = sell_row_dot(A, x, base, len);
            if (MODE == 3) out[i] = sum;
            else {
                const double d = __ldcs(A.diag_s + (NAT ? i : p)), bi = __ldcs(b_s + (NAT ? i : p));
                if (MODE == 0) {
                    const double ri = bi - (sum + d * x[i]);
                    if (out) out[i] = ri;
                    acc = ri * ri;
                } else if (MODE == 1)
                    out[i] = relax((bi - sum) / d, x[i], omega);
                else if (MODE == 5) {
                    const double xi = x[i];
                    out[i] = xi + (bi - (sum + d * xi)) / __ldcs(dl1 + i);
                } else
                    out[i] = (bi - sum) / d;
            }
        }
    }
    if (MODE == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x < 32) {
            double t = threadIdx.x < 8 ? red[threadIdx.x] : 0.;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (threadIdx.x == 0) partial[blockIdx.x] = t;
        }
    }
}

// `sweeps` whole multicolour Gauss-Seidel sweeps in ONE cooperative launch: the grid walks the colours in order with a
// grid-wide barrier after each (the next colour reads this colour's new values).  A sweep as separate launches costs
// one launch + ramp-up per colour -- 7 on the fine level, 13-16 on the Galerkin levels, where a colour holds only a
// few thousand rows and the launch gap is longer than the work.  Same row arithmetic as k_amg_sell<2>.
__global__ void __launch_bounds__(256)
k_amg_sell_gs_sweeps(SellDev A, double *x, const double *__restrict__ b_s, const int *__restrict__ colour_slot_ptr,
                     int n_colours, int sweeps)
{
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const int stride = gridDim.x * blockDim.x;
    for (int sw = 0; sw < sweeps; ++sw)
        for (int c = 0; c < n_colours; ++c) {
            const int last = colour_slot_ptr[c + 1];
            for (int p = colour_slot_ptr[c] + blockIdx.x * blockDim.x + threadIdx.x; p < last; p += stride) {
                const int i = __ldcs(A.row_of_slot + p);
                if (i < 0) continue;
                const int s = p >> 5;
                const int b0 = A.slice_ptr[s];
                const int base = b0 + (p & 31);
                const int len = (A.slice_ptr[s + 1] - b0) >> 5;
                const double sum = sell_row_dot(A, x, base, len);
                x[i] = (__ldcs(b_s + p) - sum) / __ldcs(A.diag_s + p);
            }
            grid.sync();
        }
}

// slot-ordered copy of a level vector (keeps the SELL kernels' right-hand side in step with L.b)
__global__ void __launch_bounds__(256)
k_amg_to_slots(const int *__restrict__ row_of_slot, int n_slots, const double *__restrict__ v, double *__restrict__ v_s)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_slots) return;
    const int i = row_of_slot[p];
    v_s[p] = i >= 0 ? v[i] : 0.;
}

// ---- residual r = b - A x and its squared norm (AMG/src/AMG.cpp:256-275) ------------------------------------------
template <bool EXACT>
__global__ void __launch_bounds__(256)
k_amg_residual(CsrDev A, const double *__restrict__ x, const double *__restrict__ b, double *__restrict__ r,
               double *__restrict__ partial, int row0, int row1)
{
    __shared__ double red[8];
    double acc = 0.;
    if (EXACT) {
        const int i = row0 + blockIdx.x * blockDim.x + threadIdx.x;
        if (i < row1) {
            const double ri = residual_row_exact(A, x, b, i);
            if (r) r[i] = ri;
            acc = ri * ri;
        }
    } else {
        const int lane = threadIdx.x & (kLanes - 1);
        const int i = row0 + (blockIdx.x * blockDim.x + threadIdx.x) / kLanes;
        const bool ok = i < row1;
        double Ax = 0.;
        if (ok)
            for (int k = A.ptr[i] + lane; k < A.ptr[i + 1]; k += kLanes) Ax += A.val[k] * x[A.col[k]];
        Ax = subwarp_sum(Ax);
        if (ok && lane == 0) {
            const double ri = b[i] - Ax;
            if (r) r[i] = ri;
            acc = ri * ri;
        }
    }
    // CTA reduction (256 threads)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < 8 ? red[threadIdx.x] : 0.;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) partial[blockIdx.x] = t;
    }
}

// ---- transfers -------------------------------------------------------------------------------------------------
// restriction x_c = P^T x_f as a gather over R = P^T (AMG/src/AMG.cpp:50-74): row m of R lists the fine
// rows i in ascending order, which is the order in which the reference's scatter loop adds them
template <bool EXACT>
__global__ void __launch_bounds__(256)
k_amg_spmv(CsrDev R, const double *__restrict__ xin, double *__restrict__ xout, int row0, int row1)
{
    if (EXACT) {
        const int m = row0 + blockIdx.x * blockDim.x + threadIdx.x;
        if (m >= row1) return;
        xout[m] = spmv_row_exact(R, xin, m);
    } else {
        const int lane = threadIdx.x & (kLanes - 1);
        const int m = row0 + (blockIdx.x * blockDim.x + threadIdx.x) / kLanes;
        const bool ok = m < row1;
        const double s = row_dot_vec(R, xin, ok ? m : 0, lane, ok);
        if (ok && lane == 0) xout[m] = s;
    }
}

// prolongation x_f += P x_c (AMG/src/AMG.cpp:218-232): the reference adds term by term INTO x_f[i]
__global__ void __launch_bounds__(256)
k_amg_prolong_add(CsrDev P, const double *__restrict__ xc, double *__restrict__ xf, int row0, int row1)
{
    const int i = row0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= row1) return;
    prolong_row_add(P, xc, xf, i);
}

// ---- ghost-entry exchange of a row-block sharded level (vectors keep GLOBAL indexing on every rank) -----------------
// pack: buf[t] = v[idx[t]] for the entries peers need from this rank; unpack: v[idx[t]] = buf[t] for the ghost
// entries this rank's rows reference.  The index lists are precomputed on the host (amg_solver.cu: halo_plan).
__global__ void __launch_bounds__(256)
k_amg_pack(const double *__restrict__ v, const int *__restrict__ idx, double *__restrict__ buf, int first, int last)
{
    const int t = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (t < last) buf[t] = v[idx[t]];
}
__global__ void __launch_bounds__(256)
k_amg_unpack(double *__restrict__ v, const int *__restrict__ idx, const double *__restrict__ buf, int first, int last)
{
    const int t = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (t < last) v[idx[t]] = buf[t];
}

// The same exchange over NVLink peer stores (csrc/p2p.cuh), neighbour to neighbour: entry t of the send list goes straight
// into the staging buffer of the rank that needs it -- peer[t] names the rank, off[t] the slot inside that rank's receive
// list.  Two ranks are PEERS of an exchange when either sends the other anything in it; every rank signals its peers (an
// empty signal where it has nothing to send) and waits for them, nobody else.  Each directed pair counts its messages:
// the count is the flag value, and its parity selects the half of the staging REGION THAT BELONGS TO THAT PAIR (every
// receiver keeps one region per sender), so a rank that runs one message ahead of a peer writes the half the peer is not
// unpacking; it cannot run two ahead, because the peer's next signal to it comes after that unpack.
struct AmgPush {
    double *stage[kP2PMaxRanks];                  // the region of rank p's staging buffer that receives from THIS rank (both halves)
    unsigned long long *sig[kP2PMaxRanks];        // flags[this rank][0] in the header of rank p
    unsigned long long half;                      // doubles per staging half
    unsigned long long *pair_push;                // this rank's pair_push[]
    unsigned int *done;                           // this rank's done[0]
    int n_ranks, me;
};
__device__ __forceinline__ void amg_push_tail(const AmgPush &a, unsigned mask)
{
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(a.done, 1u);
        if (prev == gridDim.x - 1) {              // the last CTA: every CTA's stores are fenced before its atomicAdd
            __threadfence_system();
            *a.done = 0u;
            for (int p = 0; p < a.n_ranks; ++p)
                if ((mask >> p) & 1u) {
                    const unsigned long long s = a.pair_push[p] + 1ull;
                    a.pair_push[p] = s;
                    st_release_sys(a.sig[p], s);
                }
        }
    }
}
__global__ void __launch_bounds__(256)
k_amg_push(const __grid_constant__ AmgPush a, const double *__restrict__ v, const int *__restrict__ idx,
           const unsigned char *__restrict__ peer, const int *__restrict__ off, int first, int last, unsigned mask)
{
    for (int t = first + blockIdx.x * blockDim.x + threadIdx.x; t < last; t += gridDim.x * blockDim.x) {
        const int p = peer[t];
        const unsigned long long buf = ((a.pair_push[p] + 1ull) & 1ull) * a.half;     // read before the last CTA bumps it
        a.stage[p][buf + (unsigned long long)off[t]] = v[idx[t]];
    }
    amg_push_tail(a, mask);
}
// contiguous variant: this rank's block [r0, r1) of a vector goes to slot r0.. of every other rank's staging buffer
__global__ void __launch_bounds__(256)
k_amg_push_block(const __grid_constant__ AmgPush a, const double *__restrict__ v, int r0, int r1, unsigned mask)
{
    for (int t = r0 + blockIdx.x * blockDim.x + threadIdx.x; t < r1; t += gridDim.x * blockDim.x) {
        const double x = v[t];
        for (int p = 0; p < a.n_ranks; ++p)
            if (p != a.me) a.stage[p][((a.pair_push[p] + 1ull) & 1ull) * a.half + (unsigned long long)(t - r0)] = x;
    }
    amg_push_tail(a, mask);
}
// waits for one more message from every rank in `mask`
__global__ void k_amg_wait(P2PHeader *h, unsigned mask)
{
    const int src = threadIdx.x;
    if (src < kP2PMaxRanks && ((mask >> src) & 1u)) {
        const unsigned long long expect = h->pair_wait[src] + 1ull;
        unsigned long long t0 = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        unsigned int spins = 0;
        while (ld_acquire_sys(&h->flags[src][0]) < expect) {
            if ((++spins & 1023u) == 0) {
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 20000000000ull) { atomicExch(&h->error, 1u + (unsigned)src); break; }
            }
        }
        h->pair_wait[src] = expect;
    }
}
// after k_amg_wait: entries [first, last) of the receive list leave the region of their sender src[t] (slot pos[t] of the message)
__global__ void __launch_bounds__(256)
k_amg_unpack_stage(double *__restrict__ v, const int *__restrict__ idx, const unsigned char *__restrict__ src,
                   const int *__restrict__ pos, const double *__restrict__ stage, unsigned long long half, const unsigned long long *__restrict__ pair_wait,
                   int first, int last)
{
    const int t = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (t < last) {
        const unsigned long long p = src[t];
        v[idx[t]] = stage[(2ull * p + (pair_wait[p] & 1ull)) * half + (unsigned long long)pos[t]];
    }
}
// the blocks of the other ranks from the staging halves into the vector; start[p] = first entry of rank p's block
struct AmgBlocks { int start[kP2PMaxRanks + 1]; };
__global__ void __launch_bounds__(256)
k_amg_unpack_blocks(double *__restrict__ v, const double *__restrict__ stage, unsigned long long half,
                    const unsigned long long *__restrict__ pair_wait, const __grid_constant__ AmgBlocks b, int n_ranks, int me)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= b.start[n_ranks]) return;
    int p = 0;
    while (t >= b.start[p + 1]) ++p;
    if (p != me) v[t] = stage[(2ull * (unsigned long long)p + (pair_wait[p] & 1ull)) * half + (unsigned long long)(t - b.start[p])];
}

// ---- on-device greedy colouring (Jones-Plassmann rounds) ------------------------------------------------------------
__device__ __forceinline__ unsigned hash32(unsigned x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
// One round: an uncoloured row whose (hash, index) priority beats all its uncoloured neighbours takes the
// smallest colour no neighbour holds.  colour[i] < 0 = uncoloured.  *remaining counts rows left.
// The colours in use are collected 64 at a time (window base, base + 64, ...), so rows of dense Galerkin levels
// whose neighbours already hold colours 0..63 still find one.
__global__ void __launch_bounds__(256)
k_amg_colour_round(CsrDev A, const int *__restrict__ colour_in, int *__restrict__ colour_out, int *remaining)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_rows) return;
    int c = colour_in[i];
    if (c >= 0) { colour_out[i] = c; return; }
    const unsigned pi = hash32((unsigned)i);
    const int k0 = A.ptr[i], k1 = A.ptr[i + 1];
    bool top = true;
    for (int k = k0; k < k1 && top; ++k) {
        const int j = A.col[k];
        if (j == i || colour_in[j] >= 0) continue;
        const unsigned pj = hash32((unsigned)j);
        if (pj > pi || (pj == pi && j > i)) top = false;
    }
    if (top) {
        for (int base = 0; c < 0; base += 64) {
            unsigned long long used = 0ull;
            for (int k = k0; k < k1; ++k) {
                const int j = A.col[k];
                if (j == i) continue;
                const int cj = colour_in[j] - base;
                if (cj >= 0 && cj < 64) used |= 1ull << cj;
            }
            if (~used) c = base + __ffsll((long long)~used) - 1;
        }
    } else
        atomicAdd(remaining, 1);
    colour_out[i] = c;
}

// 64-bit checksum of the entries [r0, r1) of a vector: wrap-around sum of bit pattern x (2 i + 1).  Integer addition is
// associative: the value does not depend on the row-block partition or on the order of the atomics.
__global__ void __launch_bounds__(256)
k_amg_checksum(const double *__restrict__ v, int r0, int r1, unsigned long long *out)
{
    unsigned long long acc = 0ull;
    for (int i = r0 + blockIdx.x * blockDim.x + threadIdx.x; i < r1; i += gridDim.x * blockDim.x)
        acc += (unsigned long long)__double_as_longlong(v[i]) * (2ull * (unsigned long long)i + 1ull);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

__global__ void __launch_bounds__(1024)
k_amg_reduce(const double *__restrict__ partial, int n, double *__restrict__ out)
{
    __shared__ double red[32];
    double acc = 0.;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = red[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) out[0] = t;
    }
}

}  // namespace mgb
