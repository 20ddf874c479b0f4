// amg_setup.cu -- AMG setup ON THE DEVICE (replaces the host restatement of
// AMG/include/AMG.hpp:105-369 on the fast path).
//
// The reference's C/F state machine (AMG.hpp:150-198) is sequential by definition -- the next seed depends on every
// earlier decision -- so the parity path keeps its exact host restatement (amg_solver.cu).  This file is the
// parity-EXEMPT scalable path: same strength measure (|a_ij| >= eps max_k |a_ik|, AMG.hpp:105-130), same direct
// interpolation weights (AMG.hpp:230-300: w_ij = a_ij / sum_{k in S_i cap C} a_ik, unit rows on C points), same
// Galerkin operator Ac = P^T A P (AMG.hpp:303-369), but the splitting is PMIS (parallel modified independent set:
// weight = number of points a node strongly influences + a hashed tie-break; local maxima become C, points that
// strongly depend on a new C point become F) and the triple product is formed by expand-sort-compress.
//   * strength / mirror flags / influence counts : one thread per row
//   * PMIS rounds                                : two kernels per round, one counter read back per round
//   * P, R = P^T                                 : count + scan + fill; transpose by a stable radix sort on the column
//   * Ac                                         : every product R(I,i) A(i,k) P(k,J) expanded once, sorted, compressed
//   * SELL-32 copies, diagonals, colour lists    : built where the matrices live, nothing is staged through the host
// Validation (tests/test_amg_device_setup_gpu.py): Ac equals P^T A P computed by scipy to 1e-12, every F point has a
// strong C neighbour, C points form an independent set of the strength graph, the correction-scheme cycle on this
// hierarchy converges at least as fast as on the reference's.
#include "amg_dev.h"
#include "dev_util.cuh"

namespace mgb {
namespace amg {

using dev::DBuf;

namespace {

__device__ __forceinline__ unsigned mix32(unsigned x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// ---- strength of connection (AMG.hpp:105-130) -----------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_strength(CsrDev A, double eps, unsigned char *__restrict__ sflag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_rows) return;
    const int k0 = A.ptr[i], k1 = A.ptr[i + 1];
    double big = 0.;
    for (int k = k0; k < k1; ++k)
        if (A.col[k] != i) big = fmax(big, fabs(A.val[k]));
    for (int k = k0; k < k1; ++k)
        sflag[k] = (A.col[k] != i && big > 0. && fabs(A.val[k]) >= eps * big) ? 1 : 0;
}

// tflag[k] of entry (i, j): is the mirror entry (j, i) strong, i.e. does j strongly depend on i?  Rows are sorted by
// column, so the mirror entry is found by bisection; a structurally unsymmetric pattern simply has no mirror.
// lambda[i] = number of points that strongly depend on i = number of set tflags in row i.
__global__ void __launch_bounds__(256)
k_mirror(CsrDev A, const unsigned char *__restrict__ sflag, unsigned char *__restrict__ tflag, int *__restrict__ lambda)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_rows) return;
    int lam = 0;
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
        const int j = A.col[k];
        unsigned char t = 0;
        if (j != i) {
            int lo = A.ptr[j], hi = A.ptr[j + 1] - 1;
            while (lo <= hi) {
                const int mid = (lo + hi) >> 1;
                const int c = A.col[mid];
                if (c == i) { t = sflag[mid]; break; }
                if (c < i) lo = mid + 1; else hi = mid - 1;
            }
        }
        tflag[k] = t;
        lam += t;
    }
    lambda[i] = lam;
}

// state: -1 undecided, 1 coarse, 0 fine
__global__ void __launch_bounds__(256)
k_pmis_init(CsrDev A, const unsigned char *__restrict__ sflag, const unsigned char *__restrict__ tflag, int *__restrict__ state)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_rows) return;
    bool coupled = false;
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) coupled = coupled || sflag[k] || tflag[k];
    state[i] = coupled ? -1 : 0;          // a row without strong couplings is solved by the smoother alone: fine, empty row of P
}

__device__ __forceinline__ bool pmis_less(int la, unsigned ha, int a, int lb, unsigned hb, int b)
{
    if (la != lb) return la < lb;
    if (ha != hb) return ha < hb;
    return a < b;
}

// an undecided point whose weight beats every undecided neighbour (in S_i or S_i^T) becomes coarse
__global__ void __launch_bounds__(256)
k_pmis_select(CsrDev A, const unsigned char *__restrict__ sflag, const unsigned char *__restrict__ tflag,
              const int *__restrict__ lambda, unsigned seed, const int *__restrict__ state_in, int *__restrict__ state_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_rows) return;
    int s = state_in[i];
    if (s < 0) {
        const int li = lambda[i];
        const unsigned hi = mix32((unsigned)i ^ seed);
        bool top = true;
        for (int k = A.ptr[i]; k < A.ptr[i + 1] && top; ++k) {
            if (!(sflag[k] | tflag[k])) continue;
            const int j = A.col[k];
            if (state_in[j] >= 0) continue;
            if (pmis_less(li, hi, i, lambda[j], mix32((unsigned)j ^ seed), j)) top = false;
        }
        if (top) s = 1;
    }
    state_out[i] = s;
}

// an undecided point that strongly depends on a coarse point becomes fine; the others are counted
__global__ void __launch_bounds__(256)
k_pmis_fine(CsrDev A, const unsigned char *__restrict__ sflag, int *state, int *remaining)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_rows) return;
    if (state[i] >= 0) return;
    bool fine = false;
    for (int k = A.ptr[i]; k < A.ptr[i + 1] && !fine; ++k)
        if (sflag[k] && state[A.col[k]] == 1) fine = true;
    if (fine) state[i] = 0;
    else atomicAdd(remaining, 1);
}

__global__ void __launch_bounds__(256)
k_is_coarse(const int *__restrict__ state, int n, int *__restrict__ flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) flag[i] = (i < n && state[i] == 1) ? 1 : 0;
}

// ---- direct interpolation (AMG.hpp:230-300) -------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_p_count(CsrDev A, const unsigned char *__restrict__ sflag, const int *__restrict__ state, int *__restrict__ len)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > A.n_rows) return;
    int c = 0;
    if (i < A.n_rows) {
        if (state[i] == 1) c = 1;
        else
            for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) c += (sflag[k] && state[A.col[k]] == 1);
    }
    len[i] = c;
}

__global__ void __launch_bounds__(256)
k_p_fill(CsrDev A, const unsigned char *__restrict__ sflag, const int *__restrict__ state, const int *__restrict__ cidx,
         const int *__restrict__ pptr, int *__restrict__ pcol, double *__restrict__ pval)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_rows) return;
    int o = pptr[i];
    if (state[i] == 1) { pcol[o] = cidx[i]; pval[o] = 1.0; return; }
    double denom = 0.;
    int cnt = 0;
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
        if (sflag[k] && state[A.col[k]] == 1) { denom += A.val[k]; ++cnt; }
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
        if (sflag[k] && state[A.col[k]] == 1) {
            pcol[o] = cidx[A.col[k]];
            pval[o] = denom != 0. ? A.val[k] / denom : 1.0 / cnt;
            ++o;
        }
}

__global__ void __launch_bounds__(256)
k_row_of_entry(const int *__restrict__ ptr, int n_rows, int *__restrict__ row_of)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    for (int k = ptr[i]; k < ptr[i + 1]; ++k) row_of[k] = i;
}

__global__ void __launch_bounds__(256)
k_iota(int *v, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

__global__ void __launch_bounds__(256)
k_col_hist(const int *__restrict__ col, int nnz, int *__restrict__ count)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nnz) atomicAdd(&count[col[k]], 1);
}

__global__ void __launch_bounds__(256)
k_transpose_gather(const int *__restrict__ order, int nnz, const int *__restrict__ row_of, const double *__restrict__ val,
                   int *__restrict__ tcol, double *__restrict__ tval)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nnz) return;
    const int k = order[t];
    tcol[t] = row_of[k];
    tval[t] = val[k];
}

// ---- Galerkin product: every term R(I,i) A(i,k) P(k,J) = P(i,I) A(i,k) P(k,J), grouped by the fine row i ----------------
__global__ void __launch_bounds__(256)
k_rap_count(CsrDev A, const int *__restrict__ pptr, int *__restrict__ cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > A.n_rows) return;
    int c = 0;
    if (i < A.n_rows) {
        int inner = 0;
        for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) { const int j = A.col[k]; inner += pptr[j + 1] - pptr[j]; }
        c = (pptr[i + 1] - pptr[i]) * inner;
    }
    cnt[i] = c;
}

__global__ void __launch_bounds__(256)
k_rap_expand(CsrDev A, const int *__restrict__ pptr, const int *__restrict__ pcol, const double *__restrict__ pval,
             const int *__restrict__ off, uint64_t *__restrict__ keys, double *__restrict__ vals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_rows) return;
    size_t o = (size_t)off[i];
    for (int a = pptr[i]; a < pptr[i + 1]; ++a) {
        const int I = pcol[a];
        const double wi = pval[a];
        for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
            const int j = A.col[k];
            const double wa = wi * A.val[k];
            for (int b = pptr[j]; b < pptr[j + 1]; ++b) {
                keys[o] = dev::esc_key(I, pcol[b]);
                vals[o] = wa * pval[b];
                ++o;
            }
        }
    }
}

// ---- diagonals --------------------------------------------------------------------------------------------------------
// diag = a_ii; dl1 = a_ii + sum_{j != i} |a_ij| (the l1-Jacobi smoother: x += (b - A x) / dl1 converges for every SPD
// operator without a damping parameter)
__global__ void __launch_bounds__(256)
k_diagonals(CsrDev A, double *__restrict__ diag, double *__restrict__ dl1)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_rows) return;
    double d = 0., off = 0.;
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
        if (A.col[k] == i) d = A.val[k];
        else off += fabs(A.val[k]);
    }
    diag[i] = d;
    if (dl1) dl1[i] = d + off;
}

// ---- SELL-32 copies ---------------------------------------------------------------------------------------------------
// Inside windows of 512 slots longer rows come first (SELL-C-sigma): a CTA sorts its window by (length desc, position).
constexpr int kWin = 512;
__global__ void __launch_bounds__(kWin)
k_sell_window_sort(CsrDev M, const int *__restrict__ list, int n_slots, int skip_diag, int *__restrict__ row_of_slot,
                   int *__restrict__ slot_len)
{
    __shared__ unsigned key[kWin];
    __shared__ int rows[kWin], lens[kWin];
    const int t = threadIdx.x, p = blockIdx.x * kWin + t;
    int row = p < n_slots ? list[p] : -1, len = 0;
    if (row >= 0) {
        len = M.ptr[row + 1] - M.ptr[row];
        if (skip_diag)
            for (int k = M.ptr[row]; k < M.ptr[row + 1]; ++k) len -= (M.col[k] == row);
    }
    rows[t] = row; lens[t] = len;
    const unsigned rank_len = row >= 0 ? (unsigned)min(len, 0x7ffe) + 1u : 0u;      // padding slots sort last
    key[t] = ((0x7fffu - rank_len) << 16) | (unsigned)t;
    __syncthreads();
    for (int k = 2; k <= kWin; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int q = t ^ j;
            if (q > t) {
                const bool up = (t & k) == 0;
                const unsigned a = key[t], b = key[q];
                if ((a > b) == up) { key[t] = b; key[q] = a; }
            }
            __syncthreads();
        }
    const int src = (int)(key[t] & 0xffffu);
    if (p < n_slots) { row_of_slot[p] = rows[src]; slot_len[p] = lens[src]; }
}

// width of every slice = its longest row (stored as width * 32 for the scan that yields slice_ptr)
__global__ void __launch_bounds__(256)
k_sell_slice_width(const int *__restrict__ slot_len, int n_slices, int *__restrict__ width32)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_slices) return;
    int w = 0;
    if (s < n_slices)
        for (int q = 0; q < 32; ++q) w = max(w, slot_len[32 * s + q]);
    width32[s] = 32 * w;
}

__global__ void __launch_bounds__(256)
k_sell_fill(CsrDev M, const int *__restrict__ row_of_slot, int n_slots, int skip_diag, const int *__restrict__ slice_ptr,
            int *__restrict__ col, double *__restrict__ val, const double *__restrict__ diag, const double *__restrict__ rhs,
            double *__restrict__ diag_s, double *__restrict__ b_s)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_slots) return;
    const int i = row_of_slot[p];
    if (diag_s) diag_s[p] = i >= 0 ? diag[i] : 1.0;
    if (b_s) b_s[p] = (i >= 0 && rhs) ? rhs[i] : 0.0;
    if (i < 0) return;
    const int base = slice_ptr[p >> 5] + (p & 31);
    int k2 = 0;
    for (int k = M.ptr[i]; k < M.ptr[i + 1]; ++k) {
        const int c = M.col[k];
        if (skip_diag && c == i) continue;
        col[base + 32 * k2] = c;
        val[base + 32 * k2] = M.val[k];
        ++k2;
    }
}

__global__ void __launch_bounds__(256)
k_range_list(int *__restrict__ list, int r0, int n_rows, int n_slots)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n_slots) list[p] = p < n_rows ? r0 + p : -1;
}

__global__ void __launch_bounds__(256)
k_group_hist(const int *__restrict__ group, int r0, int r1, int *__restrict__ count)
{
    const int i = r0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < r1) atomicAdd(&count[group[i]], 1);
}

}  // namespace

int dev_diagonals(const DevCsr &A, double *diag, double *dl1, cudaStream_t st)
{
    if (A.n_rows == 0) return MGB_OK;
    k_diagonals<<<(A.n_rows + 255) / 256, 256, 0, st>>>(A.view(), diag, dl1);
    DCK(cudaGetLastError());
    return MGB_OK;
}

// One coarsening step on the device: A -> P (n x nc), R = P^T, Ac = P^T A P.  nc == n or nc == 0 means "stop here".
int dev_coarsen(const DevCsr &A, double eps, unsigned seed, DevCsr &P, DevCsr &R, DevCsr &Ac, cudaStream_t st, int *rounds_out, int **cf_out)
{
    if (cf_out) *cf_out = nullptr;
    const int n = A.n_rows, nnz = A.nnz;
    const int gb = (n + 255) / 256, gb1 = (n + 1 + 255) / 256;
    DBuf<unsigned char> sflag, tflag;
    DBuf<int> lambda, s0, s1, d_left;
    DCK(sflag.alloc((size_t)nnz)); DCK(tflag.alloc((size_t)nnz));
    DCK(lambda.alloc((size_t)n)); DCK(s0.alloc((size_t)n)); DCK(s1.alloc((size_t)n)); DCK(d_left.alloc(1));
    const CsrDev Av = A.view();
    dev::Trace tr("coarsen");
    k_strength<<<gb, 256, 0, st>>>(Av, eps, sflag.p);
    k_mirror<<<gb, 256, 0, st>>>(Av, sflag.p, tflag.p, lambda.p);
    k_pmis_init<<<gb, 256, 0, st>>>(Av, sflag.p, tflag.p, s0.p);
    DCK(cudaGetLastError());
    int left = n, rounds = 0;
    int *cur = s0.p, *nxt = s1.p;
    while (left > 0 && rounds < 1000) {
        DCK(cudaMemsetAsync(d_left.p, 0, sizeof(int), st));
        k_pmis_select<<<gb, 256, 0, st>>>(Av, sflag.p, tflag.p, lambda.p, seed, cur, nxt);
        k_pmis_fine<<<gb, 256, 0, st>>>(Av, sflag.p, nxt, d_left.p);
        DCK(cudaGetLastError());
        if (int rc = dev::read_int(d_left.p, &left, st)) return rc;
        std::swap(cur, nxt);
        ++rounds;
    }
    if (left > 0) return mgb_set_error(MGB_ERR_STATE, "device C/F splitting did not finish");
    if (rounds_out) *rounds_out = rounds;
    tr.mark("strength + PMIS (n, rounds)", n, rounds);
    tflag.free(); lambda.free();
    const int *state = cur;
    // coarse numbering in ascending fine index (AMG.hpp:201-228)
    DBuf<int> flag, cidx;
    DCK(flag.alloc((size_t)n + 1)); DCK(cidx.alloc((size_t)n + 2));
    k_is_coarse<<<gb1, 256, 0, st>>>(state, n, flag.p);
    if (int rc = dev::exclusive_scan(flag.p, cidx.p, (size_t)n, st)) return rc;
    int nc = 0;
    if (int rc = dev::read_int(cidx.p + n, &nc, st)) return rc;
    P = DevCsr{}; R = DevCsr{}; Ac = DevCsr{};
    P.n_rows = n; P.n_cols = nc;
    if (nc == 0 || nc == n) return MGB_OK;
    // P
    DBuf<int> plen, pptr;
    DCK(plen.alloc((size_t)n + 1)); DCK(pptr.alloc((size_t)n + 2));
    k_p_count<<<gb1, 256, 0, st>>>(Av, sflag.p, state, plen.p);
    if (int rc = dev::exclusive_scan(plen.p, pptr.p, (size_t)n, st)) return rc;
    int pnnz = 0;
    if (int rc = dev::read_int(pptr.p + n, &pnnz, st)) return rc;
    plen.free(); flag.free();
    DBuf<int> pcol;
    DBuf<double> pval;
    DCK(pcol.alloc((size_t)pnnz)); DCK(pval.alloc((size_t)pnnz));
    k_p_fill<<<gb, 256, 0, st>>>(Av, sflag.p, state, cidx.p, pptr.p, pcol.p, pval.p);
    DCK(cudaGetLastError());
    sflag.free(); cidx.free();
    tr.mark("P (nc, nnz)", nc, pnnz);
    // R = P^T: stable sort of the entry numbers by column
    {
        DBuf<int> row_of, order0, order1, keys1, count, rptr;
        DCK(row_of.alloc((size_t)pnnz)); DCK(order0.alloc((size_t)pnnz)); DCK(order1.alloc((size_t)pnnz)); DCK(keys1.alloc((size_t)pnnz));
        DCK(count.alloc((size_t)nc + 1)); DCK(rptr.alloc((size_t)nc + 2));
        k_row_of_entry<<<gb, 256, 0, st>>>(pptr.p, n, row_of.p);
        k_iota<<<(pnnz + 255) / 256, 256, 0, st>>>(order0.p, pnnz);
        DCK(cudaMemsetAsync(count.p, 0, sizeof(int) * ((size_t)nc + 1), st));
        k_col_hist<<<(pnnz + 255) / 256, 256, 0, st>>>(pcol.p, pnnz, count.p);
        DCK(cudaGetLastError());
        if (int rc = dev::sort_pairs(pcol.p, keys1.p, order0.p, order1.p, (size_t)pnnz, dev::bits_for((uint64_t)nc), st)) return rc;
        if (int rc = dev::exclusive_scan(count.p, rptr.p, (size_t)nc, st)) return rc;
        DBuf<int> rcol;
        DBuf<double> rval;
        DCK(rcol.alloc((size_t)pnnz)); DCK(rval.alloc((size_t)pnnz));
        k_transpose_gather<<<(pnnz + 255) / 256, 256, 0, st>>>(order1.p, pnnz, row_of.p, pval.p, rcol.p, rval.p);
        DCK(cudaGetLastError());
        DCK(cudaStreamSynchronize(st));
        R.n_rows = nc; R.n_cols = n; R.nnz = pnnz;
        R.ptr = rptr.release(); R.col = rcol.release(); R.val = rval.release();
    }
    tr.mark("R = P^T");
    // Ac = P^T A P
    {
        DBuf<int> cnt, off;
        DCK(cnt.alloc((size_t)n + 1)); DCK(off.alloc((size_t)n + 2));
        k_rap_count<<<gb1, 256, 0, st>>>(Av, pptr.p, cnt.p);
        DCK(cudaGetLastError());
        // the total can exceed 2^31 before any int scan notices: add it up in 64 bits first
        size_t total = 0;
        if (int rc = dev::sum_int64(cnt.p, n, &total, st)) return rc;
        if (total >= ((size_t)1 << 31)) return mgb_set_error(MGB_ERR_ARG, "device Galerkin product: more than 2^31 terms on one level");
        if (int rc = dev::exclusive_scan(cnt.p, off.p, (size_t)n, st)) return rc;
        cnt.free();
        DBuf<uint64_t> keys;
        DBuf<double> vals;
        DCK(keys.alloc(total)); DCK(vals.alloc(total));
        k_rap_expand<<<gb, 256, 0, st>>>(Av, pptr.p, pcol.p, pval.p, off.p, keys.p, vals.p);
        DCK(cudaGetLastError());
        off.free();
        tr.mark("RAP expand (terms)", (long long)total);
        Ac.n_rows = Ac.n_cols = nc;
        if (int rc = dev::esc_to_csr(keys.p, vals.p, total, nc, nc, &Ac.ptr, &Ac.col, &Ac.val, &Ac.nnz, st)) return rc;
    }
    DCK(cudaStreamSynchronize(st));
    tr.mark("RAP sort + compress (nnz)", Ac.nnz);
    if (cf_out) *cf_out = (cur == s0.p) ? s0.release() : s1.release();
    P.nnz = pnnz;
    P.ptr = pptr.release(); P.col = pcol.release(); P.val = pval.release();
    return MGB_OK;
}

// SELL-32 copy of the rows listed in `list` (device array of n_slots entries, n_slots a multiple of 32, -1 = padding
// slot).  skip_diag drops a_ii; diag / rhs (may be null) are copied in slot order when slot_vectors is set.
int dev_build_sell(const DevCsr &M, const int *list, int n_slots, bool skip_diag, const double *diag, const double *rhs,
                   bool slot_vectors, SellCopy &S, cudaStream_t st)
{
    S = SellCopy{};
    S.n_slots = n_slots;
    const int n_slices = n_slots / 32;
    DCK(cudaMalloc(&S.row_of_slot, sizeof(int) * (size_t)std::max(n_slots, 1)));
    DCK(cudaMalloc(&S.slice_ptr, sizeof(int) * ((size_t)n_slices + 2)));
    if (n_slots == 0) {
        DCK(cudaMemsetAsync(S.slice_ptr, 0, sizeof(int) * 2, st));
        DCK(cudaMalloc(&S.col, sizeof(int))); DCK(cudaMalloc(&S.val, sizeof(double)));
        return MGB_OK;
    }
    DBuf<int> slot_len, width32;
    DCK(slot_len.alloc((size_t)n_slots)); DCK(width32.alloc((size_t)n_slices + 1));
    k_sell_window_sort<<<(n_slots + kWin - 1) / kWin, kWin, 0, st>>>(M.view(), list, n_slots, skip_diag ? 1 : 0, S.row_of_slot, slot_len.p);
    k_sell_slice_width<<<(n_slices + 1 + 255) / 256, 256, 0, st>>>(slot_len.p, n_slices, width32.p);
    DCK(cudaGetLastError());
    if (int rc = dev::exclusive_scan(width32.p, S.slice_ptr, (size_t)n_slices, st)) return rc;
    int stored = 0;
    if (int rc = dev::read_int(S.slice_ptr + n_slices, &stored, st)) return rc;
    if (stored < 0) return mgb_set_error(MGB_ERR_ARG, "SELL copy exceeds 2^31 entries");
    S.stored = (size_t)stored;
    DCK(cudaMalloc(&S.col, sizeof(int) * (size_t)std::max(stored, 1)));
    DCK(cudaMalloc(&S.val, sizeof(double) * (size_t)std::max(stored, 1)));
    DCK(cudaMemsetAsync(S.col, 0, sizeof(int) * (size_t)std::max(stored, 1), st));
    DCK(cudaMemsetAsync(S.val, 0, sizeof(double) * (size_t)std::max(stored, 1), st));
    if (slot_vectors) {
        DCK(cudaMalloc(&S.diag_s, sizeof(double) * (size_t)n_slots));
        DCK(cudaMalloc(&S.b_s, sizeof(double) * (size_t)n_slots));
    }
    k_sell_fill<<<(n_slots + 255) / 256, 256, 0, st>>>(M.view(), S.row_of_slot, n_slots, skip_diag ? 1 : 0, S.slice_ptr, S.col, S.val,
                                                       diag, rhs, S.diag_s, S.b_s);
    DCK(cudaGetLastError());
    DCK(cudaStreamSynchronize(st));
    return MGB_OK;
}

// natural-order copy of the rows [rows.r0, rows.r1)
int dev_build_sell_range(const DevCsr &M, Block rows, bool skip_diag, SellCopy &S, cudaStream_t st)
{
    const int n = rows.size(), n_slots = (n + 31) / 32 * 32;
    DBuf<int> list;
    DCK(list.alloc((size_t)std::max(n_slots, 1)));
    if (n_slots) k_range_list<<<(n_slots + 255) / 256, 256, 0, st>>>(list.p, rows.r0, n, n_slots);
    DCK(cudaGetLastError());
    return dev_build_sell(M, list.p, n_slots, skip_diag, nullptr, nullptr, false, S, st);
}

// Independent sets (colours) of the rows [own.r0, own.r1) from a device array group[i]: the device row lists of the
// schedule (rows of a group ascending) and, when `sell` is given, the colour-sorted SELL copy (every colour starts on
// a window boundary).
int dev_group_schedule(const DevCsr &A, const int *group, int n_groups, Block own, Schedule &Sch, SellCopy *sell,
                       const double *diag, const double *rhs, cudaStream_t st)
{
    const int n = own.size();
    Sch.n_groups = n_groups;
    DBuf<int> count;
    DCK(count.alloc((size_t)n_groups + 1));
    DCK(cudaMemsetAsync(count.p, 0, sizeof(int) * ((size_t)n_groups + 1), st));
    if (n) k_group_hist<<<(n + 255) / 256, 256, 0, st>>>(group, own.r0, own.r1, count.p);
    DCK(cudaGetLastError());
    std::vector<int> h_count((size_t)n_groups + 1, 0);
    DCK(cudaMemcpyAsync(h_count.data(), count.p, sizeof(int) * (size_t)n_groups, cudaMemcpyDeviceToHost, st));
    DCK(cudaStreamSynchronize(st));
    Sch.h_ptr.assign((size_t)n_groups + 1, 0);
    for (int g = 0; g < n_groups; ++g) Sch.h_ptr[g + 1] = Sch.h_ptr[g] + h_count[g];
    DCK(cudaMalloc(&Sch.d_ptr, sizeof(int) * ((size_t)n_groups + 1)));
    DCK(cudaMalloc(&Sch.d_rows, sizeof(int) * (size_t)std::max(n, 1)));
    DCK(cudaMemcpyAsync(Sch.d_ptr, Sch.h_ptr.data(), sizeof(int) * ((size_t)n_groups + 1), cudaMemcpyHostToDevice, st));
    if (n) {
        DBuf<int> rows0, keys1;
        DCK(rows0.alloc((size_t)n)); DCK(keys1.alloc((size_t)n));
        k_range_list<<<(n + 255) / 256, 256, 0, st>>>(rows0.p, own.r0, n, n);
        DCK(cudaGetLastError());
        if (int rc = dev::sort_pairs(group + own.r0, keys1.p, rows0.p, Sch.d_rows, (size_t)n, dev::bits_for((uint64_t)std::max(n_groups, 1)), st)) return rc;
    }
    if (!sell) return MGB_OK;
    std::vector<int> start((size_t)n_groups + 1, 0);
    for (int g = 0; g < n_groups; ++g) start[g + 1] = (start[g] + h_count[g] + kWin - 1) / kWin * kWin;
    const int n_slots = start[n_groups];
    DBuf<int> list;
    DCK(list.alloc((size_t)std::max(n_slots, 1)));
    DCK(cudaMemsetAsync(list.p, 0xFF, sizeof(int) * (size_t)std::max(n_slots, 1), st));
    for (int g = 0; g < n_groups; ++g)
        if (h_count[g])
            DCK(cudaMemcpyAsync(list.p + start[g], Sch.d_rows + Sch.h_ptr[g], sizeof(int) * (size_t)h_count[g], cudaMemcpyDeviceToDevice, st));
    if (int rc = dev_build_sell(A, list.p, n_slots, true, diag, rhs, true, *sell, st)) return rc;
    sell->colour_slot_ptr = start;
    DCK(cudaMalloc(&sell->d_colour_slot_ptr, sizeof(int) * ((size_t)n_groups + 1)));
    DCK(cudaMemcpy(sell->d_colour_slot_ptr, start.data(), sizeof(int) * ((size_t)n_groups + 1), cudaMemcpyHostToDevice));
    return MGB_OK;
}

}  // namespace amg
}  // namespace mgb
