// amg_solver.cu -- host side of the B200 AMG solve phase + its C ABI.
//
// Replaces (reference file:line, relative to AMG/):
//   level-hierarchy storage   Matrix (vector<map>) / CSRMatrix (AoS)          include/CSRMatrix.hpp:19-121
//   setup                     AMG::initialization + RestrictionOperator       src/AMG.cpp:76-120, include/AMG.hpp:105-369
//   smoother / residual       Gauss_Seidel_iteration, AMG::compute_residual   include/Utilities.hpp:37-97, src/AMG.cpp:256-275
//   transfers                 apply_restriction/prolungation_operator         src/AMG.cpp:50-74, 218-232
//   cycle                     AMG::apply_AMG                                  src/AMG.cpp:277-308
// The setup keeps the reference's semantics (strength threshold 0.2, its C/F state machine, direct
// interpolation weights, Galerkin product evaluated in the reference's term order) but is written as
// O(nnz) sparse loops on the host; the hierarchy is then uploaded once and the whole solve phase runs
// on the device.  (Device-side setup is the next item of SURVEY.md section 8f.)
#include "../../include/mgb200.h"
#include "amg_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

int mgb_set_error(int code, const std::string &msg);   // gmg_solver.cu

namespace {

#define ACK(call)                                                                             \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return mgb_set_error(MGB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct HostCsr {
    int n_rows = 0, n_cols = 0;
    std::vector<int> ptr, col;
    std::vector<double> val;
    int nnz() const { return ptr.empty() ? 0 : ptr.back(); }
    double at(int i, int j) const      // first entry of row i in column j, 0.0 when absent (CSRMatrix.cpp:24-40)
    {
        for (int k = ptr[i]; k < ptr[i + 1]; ++k) if (col[k] == j) return val[k];
        return 0.0;
    }
};

// strong couplings of a row: off-diagonal entries with |a_ij| >= eps * max_k |a_ik| (AMG.hpp:105-130)
void strong_of_row(const HostCsr &A, int i, double eps, std::vector<int> &out)
{
    out.clear();
    double big = 0.0;
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
        if (A.col[k] != i) big = std::max(big, std::fabs(A.val[k]));
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
        if (A.col[k] != i && std::fabs(A.val[k]) >= eps * big) out.push_back(A.col[k]);
}

// C/F splitting with the reference's byte state machine (AMG.hpp:150-198): low six bits = number of
// strong couplings still pointing at undecided nodes (+2 per neighbour turned fine), 0xC0 = fine.
// The next seed is the largest index whose counter is still non-zero; counters only ever drop to zero,
// so one pointer walking down from n-1 replaces the reference's full rescan.
int split_coarse_fine(const HostCsr &A, double eps, long start, std::vector<unsigned char> &state)
{
    const int n = A.n_rows;
    std::vector<int> sp(n + 1, 0), sc;
    sc.reserve(A.nnz());
    std::vector<int> tmp;
    state.assign(n, 0);
    for (int i = 0; i < n; ++i) {
        strong_of_row(A, i, eps, tmp);
        sc.insert(sc.end(), tmp.begin(), tmp.end());
        sp[i + 1] = (int)sc.size();
        state[i] = (unsigned char)tmp.size();
    }
    if (n == 0) return 0;
    int seed = (int)std::min<long>(std::max<long>(start, 0), n - 1);
    int fine = 0, walker = n - 1;
    while (state[seed] & 0x3F) {
        state[seed] = 0;                                            // coarse
        for (int a = sp[seed]; a < sp[seed + 1]; ++a) {
            const int c = sc[a];
            if (!(state[c] & 0x3F)) continue;
            state[c] = (unsigned char)((state[c] | 0xC0) & 0xC0);   // fine, counter cleared
            ++fine;
            for (int b = sp[c]; b < sp[c + 1]; ++b)
                if (state[sc[b]] & 0x3F) state[sc[b]] = (unsigned char)(state[sc[b]] + 2);
        }
        while (walker >= 0 && !(state[walker] & 0x3F)) --walker;
        if (walker >= 0) seed = walker;
    }
    return n - fine;
}

// direct interpolation (AMG.hpp:230-300): coarse rows are unit rows; a fine row i gets
// w_ij = alpha a_ij / sum_k(alpha a_ik) over its strong coarse neighbours, alpha = (sum_{j!=i} a_ij) / (sum_k a_ik)
HostCsr interpolation(const HostCsr &A, double eps, const std::vector<unsigned char> &state, int nc)
{
    const int n = A.n_rows;
    std::vector<int> cidx(n, -1);
    for (int i = 0, k = 0; i < n; ++i) if (!(state[i] & 0xC0)) cidx[i] = k++;
    HostCsr P;
    P.n_rows = n; P.n_cols = nc; P.ptr.assign(n + 1, 0);
    std::vector<int> strong;
    for (int i = 0; i < n; ++i) {
        if (!(state[i] & 0xC0)) { P.col.push_back(cidx[i]); P.val.push_back(1.0); P.ptr[i + 1] = (int)P.col.size(); continue; }
        double off_sum = 0.0;
        for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) if (A.col[k] != i) off_sum += A.val[k];
        strong_of_row(A, i, eps, strong);
        double coarse_sum = 0.0;
        for (int j : strong) if (!(state[j] & 0xC0)) coarse_sum += A.at(i, j);
        const double alpha = off_sum / coarse_sum;
        double norm = 0.0;
        for (int j : strong) if (!(state[j] & 0xC0)) norm += alpha * A.at(i, j);
        for (int j : strong)
            if (!(state[j] & 0xC0)) {
                const double w = alpha * A.at(i, j) / norm;
                if (w != 0) { P.col.push_back(cidx[j]); P.val.push_back(w); }     // exact zeros are dropped (CSRMatrix.cpp:13-14)
            }
        P.ptr[i + 1] = (int)P.col.size();
    }
    return P;
}

HostCsr transpose(const HostCsr &M)
{
    HostCsr T;
    T.n_rows = M.n_cols; T.n_cols = M.n_rows;
    T.ptr.assign(T.n_rows + 1, 0);
    for (int c : M.col) T.ptr[c + 1]++;
    for (int i = 0; i < T.n_rows; ++i) T.ptr[i + 1] += T.ptr[i];
    T.col.resize(M.col.size()); T.val.resize(M.val.size());
    std::vector<int> fill(T.ptr.begin(), T.ptr.end() - 1);
    for (int i = 0; i < M.n_rows; ++i)
        for (int k = M.ptr[i]; k < M.ptr[i + 1]; ++k) { int p = fill[M.col[k]]++; T.col[p] = i; T.val[p] = M.val[k]; }
    return T;
}

// one sparse row of a product: out(c) = sum over the entries (k, v) of `row` taken in order, of v * B(k, c);
// exact zeros dropped; columns emitted ascending
struct RowAccumulator {
    std::vector<int> stamp, touched;
    std::vector<double> acc;
    explicit RowAccumulator(int n) : stamp(n, -1), acc(n, 0.0) {}
    void run(int tag, const int *rcol, const double *rval, int len, const HostCsr &B, HostCsr &out)
    {
        touched.clear();
        for (int a = 0; a < len; ++a) {
            const int k = rcol[a];
            for (int b = B.ptr[k]; b < B.ptr[k + 1]; ++b) {
                const int c = B.col[b];
                if (stamp[c] != tag) { stamp[c] = tag; acc[c] = 0.0; touched.push_back(c); }
                acc[c] += rval[a] * B.val[b];
            }
        }
        std::sort(touched.begin(), touched.end());
        for (int c : touched) if (acc[c] != 0) { out.col.push_back(c); out.val.push_back(acc[c]); }
    }
};

// Galerkin operator in the reference's evaluation order (AMG.hpp:303-369):
//   PtA(i,j) = sum_k A(j,k) P(k,i)  (k ascending; A taken as symmetric),   Ac(i,j) = sum_k PtA(i,k) P(k,j)
HostCsr galerkin(const HostCsr &A, const HostCsr &P)
{
    const int n = A.n_rows, nc = P.n_cols;
    HostCsr AP;                                  // AP(j, i) = PtA(i, j)
    AP.n_rows = n; AP.n_cols = nc; AP.ptr.assign(n + 1, 0);
    RowAccumulator ra(std::max(n, nc));
    for (int j = 0; j < n; ++j) {
        ra.run(j, &A.col[A.ptr[j]], &A.val[A.ptr[j]], A.ptr[j + 1] - A.ptr[j], P, AP);
        AP.ptr[j + 1] = (int)AP.col.size();
    }
    HostCsr PtA = transpose(AP);
    HostCsr Ac;
    Ac.n_rows = nc; Ac.n_cols = nc; Ac.ptr.assign(nc + 1, 0);
    RowAccumulator rb(std::max(n, nc));
    for (int i = 0; i < nc; ++i) {
        rb.run(i, &PtA.col[PtA.ptr[i]], &PtA.val[PtA.ptr[i]], PtA.ptr[i + 1] - PtA.ptr[i], P, Ac);
        Ac.ptr[i + 1] = (int)Ac.col.size();
    }
    return Ac;
}

struct DevCsr {
    int n_rows = 0, n_cols = 0, nnz = 0;
    int *ptr = nullptr, *col = nullptr;
    double *val = nullptr;
    mgb::CsrDev view() const { return mgb::CsrDev{n_rows, n_cols, nnz, ptr, col, val}; }
    void release() { cudaFree(ptr); cudaFree(col); cudaFree(val); ptr = col = nullptr; val = nullptr; }
};

struct Schedule {          // rows grouped into independent sets (wavefronts or colours)
    int n_groups = 0;
    std::vector<int> h_ptr, h_group;
    int *d_ptr = nullptr, *d_rows = nullptr;
    void release() { cudaFree(d_ptr); cudaFree(d_rows); d_ptr = d_rows = nullptr; }
};

struct SellCopy {              // colour-sorted SELL-32 copy of A for the fast kernels (amg_kernels.cuh)
    int n_slots = 0;
    std::vector<int> colour_slot_ptr;      // first slot of each colour (+ end)
    int *slice_ptr = nullptr, *col = nullptr, *row_of_slot = nullptr;
    double *val = nullptr, *diag_s = nullptr, *b_s = nullptr;
    size_t stored = 0;                     // entries incl. padding
    mgb::SellDev view() const { return mgb::SellDev{n_slots, slice_ptr, col, val, row_of_slot, diag_s, b_s}; }
    void release() { cudaFree(slice_ptr); cudaFree(col); cudaFree(row_of_slot); cudaFree(val); cudaFree(diag_s); cudaFree(b_s); }
};

struct AmgLevel {
    HostCsr hA, hP;                       // host copies (hierarchy queries, schedules)
    std::vector<double> h_rhs;
    DevCsr A, P, R;
    double *diag = nullptr, *x = nullptr, *b = nullptr, *tmp = nullptr;
    Schedule lex, colour;
    SellCopy sell;
};

}  // namespace

struct mgb_amg {
    mgb_amg_config cfg{};
    std::vector<AmgLevel> lv;
    cudaStream_t st = nullptr;
    double *d_partial = nullptr, *d_scal = nullptr, *h_scal = nullptr;
    mgb_gmg_stats stats{};
};

namespace {

int upload(const HostCsr &H, DevCsr &D, cudaStream_t st)
{
    D.n_rows = H.n_rows; D.n_cols = H.n_cols; D.nnz = H.nnz();
    ACK(cudaMalloc(&D.ptr, sizeof(int) * (size_t)(H.n_rows + 1)));
    ACK(cudaMalloc(&D.col, sizeof(int) * (size_t)std::max(D.nnz, 1)));
    ACK(cudaMalloc(&D.val, sizeof(double) * (size_t)std::max(D.nnz, 1)));
    ACK(cudaMemcpyAsync(D.ptr, H.ptr.data(), sizeof(int) * (size_t)(H.n_rows + 1), cudaMemcpyHostToDevice, st));
    if (D.nnz) {
        ACK(cudaMemcpyAsync(D.col, H.col.data(), sizeof(int) * (size_t)D.nnz, cudaMemcpyHostToDevice, st));
        ACK(cudaMemcpyAsync(D.val, H.val.data(), sizeof(double) * (size_t)D.nnz, cudaMemcpyHostToDevice, st));
    }
    return MGB_OK;
}

int upload_schedule(const std::vector<int> &group_of_row, int n_groups, Schedule &S, cudaStream_t st)
{
    const int n = (int)group_of_row.size();
    S.n_groups = n_groups;
    S.h_group = group_of_row;
    S.h_ptr.assign(n_groups + 1, 0);
    for (int g : group_of_row) S.h_ptr[g + 1]++;
    for (int g = 0; g < n_groups; ++g) S.h_ptr[g + 1] += S.h_ptr[g];
    std::vector<int> rows(std::max(n, 1)), fill(S.h_ptr.begin(), S.h_ptr.end() - 1);
    for (int i = 0; i < n; ++i) rows[fill[group_of_row[i]]++] = i;       // ascending inside a group
    ACK(cudaMalloc(&S.d_ptr, sizeof(int) * (size_t)(n_groups + 1)));
    ACK(cudaMalloc(&S.d_rows, sizeof(int) * (size_t)std::max(n, 1)));
    ACK(cudaMemcpyAsync(S.d_ptr, S.h_ptr.data(), sizeof(int) * (size_t)(n_groups + 1), cudaMemcpyHostToDevice, st));
    ACK(cudaMemcpyAsync(S.d_rows, rows.data(), sizeof(int) * (size_t)std::max(n, 1), cudaMemcpyHostToDevice, st));
    ACK(cudaStreamSynchronize(st));
    return MGB_OK;
}

inline void tally(mgb_amg *h, double bytes) { h->stats.kernel_launches++; h->stats.bytes_algorithmic += bytes; }
inline double sweep_bytes(const AmgLevel &L) { return 12. * L.A.nnz + 28. * L.A.n_rows; }

// level schedule of the lexicographic sweep: wave(i) = 1 + max wave(j) over the couplings j < i
int build_lex_schedule(mgb_amg *h, AmgLevel &L)
{
    const HostCsr &A = L.hA;
    std::vector<int> wave(A.n_rows, 0);
    int n_waves = A.n_rows ? 1 : 0;
    for (int i = 0; i < A.n_rows; ++i) {
        int w = 0;
        for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) if (A.col[k] < i) w = std::max(w, wave[A.col[k]] + 1);
        wave[i] = w;
        n_waves = std::max(n_waves, w + 1);
    }
    return upload_schedule(wave, n_waves, L.lex, h->st);
}

// greedy colouring on the device (Jones-Plassmann rounds); the colour lists are then laid out on the host
int build_colouring(mgb_amg *h, AmgLevel &L)
{
    const int n = L.A.n_rows;
    int *c0 = nullptr, *c1 = nullptr, *d_left = nullptr;
    ACK(cudaMalloc(&c0, sizeof(int) * (size_t)std::max(n, 1)));
    ACK(cudaMalloc(&c1, sizeof(int) * (size_t)std::max(n, 1)));
    ACK(cudaMalloc(&d_left, sizeof(int)));
    ACK(cudaMemsetAsync(c0, 0xFF, sizeof(int) * (size_t)std::max(n, 1), h->st));
    int left = n, rounds = 0;
    while (left > 0 && rounds < 10000) {
        ACK(cudaMemsetAsync(d_left, 0, sizeof(int), h->st));
        mgb::k_amg_colour_round<<<(n + 255) / 256, 256, 0, h->st>>>(L.A.view(), c0, c1, d_left);
        tally(h, 0.);
        ACK(cudaMemcpyAsync(&left, d_left, sizeof(int), cudaMemcpyDeviceToHost, h->st));
        ACK(cudaStreamSynchronize(h->st));
        std::swap(c0, c1);
        ++rounds;
    }
    std::vector<int> colour(std::max(n, 1));
    ACK(cudaMemcpy(colour.data(), c0, sizeof(int) * (size_t)std::max(n, 1), cudaMemcpyDeviceToHost));
    cudaFree(c0); cudaFree(c1); cudaFree(d_left);
    colour.resize(n);
    int nc = 0;
    for (int c : colour) { if (c < 0) return mgb_set_error(MGB_ERR_STATE, "colouring did not finish"); nc = std::max(nc, c + 1); }
    return upload_schedule(colour, nc, L.colour, h->st);
}

// colour-sorted SELL-32 copy (off-diagonal entries) + slot-ordered diagonal and rhs
int build_sell(mgb_amg *h, AmgLevel &L)
{
    const HostCsr &A = L.hA;
    const int n = A.n_rows, ncol = L.colour.n_groups;
    SellCopy &S = L.sell;
    std::vector<int> row_of_slot;
    S.colour_slot_ptr.assign(ncol + 1, 0);
    {
        std::vector<std::vector<int>> by_colour(std::max(ncol, 1));
        for (int i = 0; i < n; ++i) by_colour[L.colour.h_group[i]].push_back(i);
        for (int c = 0; c < ncol; ++c) {
            S.colour_slot_ptr[c] = (int)row_of_slot.size();
            row_of_slot.insert(row_of_slot.end(), by_colour[c].begin(), by_colour[c].end());
            while (row_of_slot.size() % 32) row_of_slot.push_back(-1);        // every colour starts on a slice boundary
        }
        S.colour_slot_ptr[ncol] = (int)row_of_slot.size();
    }
    S.n_slots = (int)row_of_slot.size();
    const int n_slices = S.n_slots / 32;
    std::vector<int> slice_ptr(n_slices + 1, 0);
    for (int s = 0; s < n_slices; ++s) {
        int longest = 0;
        for (int q = 0; q < 32; ++q) {
            const int i = row_of_slot[32 * s + q];
            if (i < 0) continue;
            int len = 0;
            for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) len += (A.col[k] != i);
            longest = std::max(longest, len);
        }
        slice_ptr[s + 1] = slice_ptr[s] + 32 * longest;
    }
    S.stored = (size_t)slice_ptr[n_slices];
    std::vector<int> col(std::max<size_t>(S.stored, 1), 0);
    std::vector<double> val(std::max<size_t>(S.stored, 1), 0.0), diag_s(std::max(S.n_slots, 1), 1.0), b_s(std::max(S.n_slots, 1), 0.0);
    for (int p = 0; p < S.n_slots; ++p) {
        const int i = row_of_slot[p];
        if (i < 0) continue;
        int k2 = 0;
        const int base = slice_ptr[p >> 5] + (p & 31);
        for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
            if (A.col[k] == i) continue;
            col[base + 32 * k2] = A.col[k]; val[base + 32 * k2] = A.val[k]; ++k2;
        }
        diag_s[p] = A.at(i, i);
        b_s[p] = L.h_rhs[i];
    }
    ACK(cudaMalloc(&S.slice_ptr, sizeof(int) * (size_t)(n_slices + 1)));
    ACK(cudaMalloc(&S.col, sizeof(int) * col.size()));
    ACK(cudaMalloc(&S.val, sizeof(double) * val.size()));
    ACK(cudaMalloc(&S.row_of_slot, sizeof(int) * (size_t)std::max(S.n_slots, 1)));
    ACK(cudaMalloc(&S.diag_s, sizeof(double) * diag_s.size()));
    ACK(cudaMalloc(&S.b_s, sizeof(double) * b_s.size()));
    ACK(cudaMemcpy(S.slice_ptr, slice_ptr.data(), sizeof(int) * (size_t)(n_slices + 1), cudaMemcpyHostToDevice));
    ACK(cudaMemcpy(S.col, col.data(), sizeof(int) * col.size(), cudaMemcpyHostToDevice));
    ACK(cudaMemcpy(S.val, val.data(), sizeof(double) * val.size(), cudaMemcpyHostToDevice));
    if (S.n_slots) ACK(cudaMemcpy(S.row_of_slot, row_of_slot.data(), sizeof(int) * (size_t)S.n_slots, cudaMemcpyHostToDevice));
    ACK(cudaMemcpy(S.diag_s, diag_s.data(), sizeof(double) * diag_s.size(), cudaMemcpyHostToDevice));
    ACK(cudaMemcpy(S.b_s, b_s.data(), sizeof(double) * b_s.size(), cudaMemcpyHostToDevice));
    return MGB_OK;
}

int do_smooth(mgb_amg *h, int level, int kind, int sweeps)
{
    AmgLevel &L = h->lv[level];
    const mgb::CsrDev A = L.A.view();
    if (A.n_rows == 0 || sweeps <= 0) return MGB_OK;
    if (kind == MGB_SMOOTH_GS_LEX) {
        if (A.n_rows <= (1 << 18)) {
            mgb::k_amg_gs_lex_cta<<<1, 1024, 0, h->st>>>(A, L.diag, L.x, L.b, L.lex.d_ptr, L.lex.d_rows, L.lex.n_groups, sweeps);
            tally(h, sweep_bytes(L) * sweeps);
        } else {
            for (int s = 0; s < sweeps; ++s)
                for (int w = 0; w < L.lex.n_groups; ++w) {
                    const int a = L.lex.h_ptr[w], b = L.lex.h_ptr[w + 1];
                    mgb::k_amg_gs_rows_exact<<<(b - a + 255) / 256, 256, 0, h->st>>>(A, L.diag, L.x, L.b, L.lex.d_rows, a, b);
                    tally(h, sweep_bytes(L) * (double)(b - a) / A.n_rows);
                }
        }
    } else if (kind == MGB_SMOOTH_GS_RB) {            // multicolour Gauss-Seidel
        for (int s = 0; s < sweeps; ++s)
            for (int c = 0; c < L.colour.n_groups; ++c) {
                const int a = L.colour.h_ptr[c], b = L.colour.h_ptr[c + 1];
                if (h->cfg.exact_order)
                    mgb::k_amg_gs_rows_exact<<<(b - a + 255) / 256, 256, 0, h->st>>>(A, L.diag, L.x, L.b, L.colour.d_rows, a, b);
                else {
                    const int p0 = L.sell.colour_slot_ptr[c], p1 = L.sell.colour_slot_ptr[c + 1];
                    if (p1 > p0)
                        mgb::k_amg_sell<2><<<(p1 - p0 + 255) / 256, 256, 0, h->st>>>(L.sell.view(), L.x, L.sell.b_s, L.x, nullptr, p0, p1);
                }
                tally(h, sweep_bytes(L) * (double)(b - a) / A.n_rows);
            }
    } else if (kind == MGB_SMOOTH_JACOBI) {
        for (int s = 0; s < sweeps; ++s) {
            if (h->cfg.exact_order)
                mgb::k_amg_jacobi_vec<<<(A.n_rows * mgb::kLanes + 255) / 256, 256, 0, h->st>>>(A, L.diag, L.x, L.b, L.tmp);
            else
                mgb::k_amg_sell<1><<<(L.sell.n_slots + 255) / 256, 256, 0, h->st>>>(L.sell.view(), L.x, L.sell.b_s, L.tmp, nullptr, 0, L.sell.n_slots);
            tally(h, sweep_bytes(L));
            std::swap(L.x, L.tmp);
        }
    } else
        return mgb_set_error(MGB_ERR_ARG, "unknown AMG smoother");
    ACK(cudaGetLastError());
    return MGB_OK;
}

int do_residual(mgb_amg *h, int level, double *norm)
{
    AmgLevel &L = h->lv[level];
    const mgb::CsrDev A = L.A.view();
    int blocks;
    if (h->cfg.exact_order) {
        blocks = (A.n_rows + 255) / 256;
        mgb::k_amg_residual<true><<<blocks, 256, 0, h->st>>>(A, L.x, L.b, L.tmp, h->d_partial);
    } else {
        blocks = (L.sell.n_slots + 255) / 256;
        mgb::k_amg_sell<0><<<blocks, 256, 0, h->st>>>(L.sell.view(), L.x, L.sell.b_s, L.tmp, h->d_partial, 0, L.sell.n_slots);
    }
    tally(h, sweep_bytes(L));
    mgb::k_amg_reduce<<<1, 1024, 0, h->st>>>(h->d_partial, blocks, h->d_scal);
    tally(h, 0.);
    ACK(cudaGetLastError());
    ACK(cudaMemcpyAsync(h->h_scal, h->d_scal, sizeof(double), cudaMemcpyDeviceToHost, h->st));
    ACK(cudaStreamSynchronize(h->st));
    *norm = std::sqrt(h->h_scal[0]);
    return MGB_OK;
}

// x_{level} = P^T x_{level-1}  (AMG.cpp:50-74)
int do_restrict(mgb_amg *h, int level)
{
    AmgLevel &F = h->lv[level - 1], &C = h->lv[level];
    const mgb::CsrDev R = F.R.view();
    if (R.n_rows == 0) return MGB_OK;
    if (h->cfg.exact_order) mgb::k_amg_spmv<true><<<(R.n_rows + 255) / 256, 256, 0, h->st>>>(R, F.x, C.x);
    else mgb::k_amg_spmv<false><<<(R.n_rows * mgb::kLanes + 255) / 256, 256, 0, h->st>>>(R, F.x, C.x);
    tally(h, 12. * R.nnz + 12. * R.n_rows + 8. * R.n_cols);
    ACK(cudaGetLastError());
    return MGB_OK;
}

// x_level += P x_{level+1}  (AMG.cpp:218-232)
int do_prolong(mgb_amg *h, int level)
{
    AmgLevel &F = h->lv[level], &C = h->lv[level + 1];
    const mgb::CsrDev P = F.P.view();
    if (P.n_rows == 0) return MGB_OK;
    mgb::k_amg_prolong_add<<<(P.n_rows + 255) / 256, 256, 0, h->st>>>(P, C.x, F.x);
    tally(h, 12. * P.nnz + 20. * P.n_rows + 8. * P.n_cols);
    ACK(cudaGetLastError());
    return MGB_OK;
}

}  // namespace

extern "C" {

void mgb_amg_config_default(mgb_amg_config *c)
{
    std::memset(c, 0, sizeof(*c));
    c->levels = 5;                 // AMG/src/main.cpp:126
    c->eps = 0.2;                  // AMG/include/AMG.hpp:21
    c->smoother = MGB_SMOOTH_GS_LEX;
    c->pre_sweeps = 10; c->coarse_sweeps = 200; c->post_sweeps = 10;      // AMG/src/AMG.cpp:287,295,302
    c->exact_order = 1;
    c->device = 0;
    for (int i = 0; i < 16; ++i) c->start_index[i] = -1;                   // -1: n/2 (the reference draws it at random)
}

void mgb_amg_config_fast(mgb_amg_config *c)
{
    mgb_amg_config_default(c);
    c->smoother = MGB_SMOOTH_GS_RB;      // multicolour Gauss-Seidel
    c->exact_order = 0;
}

int mgb_amg_create_from_csr(const mgb_amg_config *cfg, size_t n, const int64_t *ptr, const int64_t *col,
                            const double *val, const double *rhs, mgb_amg_t *out)
{
    if (!cfg || !ptr || !col || !val || !rhs || !out) return mgb_set_error(MGB_ERR_ARG, "null argument");
    *out = nullptr;
    if (cfg->levels < 1 || cfg->levels > 16) return mgb_set_error(MGB_ERR_ARG, "1 <= levels <= 16");
    if (n == 0 || n > (size_t)1 << 30 || ptr[n] > (int64_t)INT32_MAX) return mgb_set_error(MGB_ERR_ARG, "matrix too large for int32 indices");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return mgb_set_error(MGB_ERR_CUDA, "no CUDA device: libmgb200 has no CPU fallback");
    }
    ACK(cudaSetDevice(cfg->device));
    mgb_amg *h = new mgb_amg();
    h->cfg = *cfg;
    ACK(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
    h->lv.resize(cfg->levels);
    // level 0: CSRMatrix::copy_from drops exact zeros (CSRMatrix.cpp:13-14); rows must be column-sorted (they come from a map)
    {
        HostCsr &A = h->lv[0].hA;
        A.n_rows = A.n_cols = (int)n;
        A.ptr.assign(n + 1, 0);
        for (size_t i = 0; i < n; ++i) {
            for (int64_t k = ptr[i]; k < ptr[i + 1]; ++k) {
                if (col[k] < 0 || col[k] >= (int64_t)n) { delete h; return mgb_set_error(MGB_ERR_ARG, "column index out of range"); }
                if (k > ptr[i] && col[k] <= col[k - 1]) { delete h; return mgb_set_error(MGB_ERR_ARG, "rows must be sorted by column without duplicates"); }
                if (val[k] != 0) { A.col.push_back((int)col[k]); A.val.push_back(val[k]); }
            }
            A.ptr[i + 1] = (int)A.col.size();
        }
        h->lv[0].h_rhs.assign(rhs, rhs + n);
    }
    // AMG::initialization (AMG.cpp:76-120)
    for (int l = 1; l < cfg->levels; ++l) {
        AmgLevel &F = h->lv[l - 1], &C = h->lv[l];
        std::vector<unsigned char> state;
        const long start = cfg->start_index[l - 1] >= 0 ? (long)cfg->start_index[l - 1] : F.hA.n_rows / 2;
        const int nc = split_coarse_fine(F.hA, cfg->eps, start, state);
        F.hP = interpolation(F.hA, cfg->eps, state, nc);
        C.h_rhs.assign(nc, 0.0);                                                      // b_c = P^T b  (AMG.cpp:100-109)
        for (int i = 0; i < F.hP.n_rows; ++i)
            for (int k = F.hP.ptr[i]; k < F.hP.ptr[i + 1]; ++k) C.h_rhs[F.hP.col[k]] += F.hP.val[k] * F.h_rhs[i];
        C.hA = galerkin(F.hA, F.hP);
    }
    // upload
    size_t max_blocks = 1;
    for (int l = 0; l < cfg->levels; ++l) {
        AmgLevel &L = h->lv[l];
        const int nl = L.hA.n_rows;
        int rc;
        if ((rc = upload(L.hA, L.A, h->st))) return rc;
        std::vector<double> dg(std::max(nl, 1), 0.0);
        for (int i = 0; i < nl; ++i) dg[i] = L.hA.at(i, i);
        const size_t bytes = sizeof(double) * (size_t)std::max(nl, 1);
        ACK(cudaMalloc(&L.diag, bytes)); ACK(cudaMalloc(&L.x, bytes)); ACK(cudaMalloc(&L.b, bytes)); ACK(cudaMalloc(&L.tmp, bytes));
        ACK(cudaMemcpyAsync(L.diag, dg.data(), bytes, cudaMemcpyHostToDevice, h->st));
        ACK(cudaMemsetAsync(L.x, 0, bytes, h->st));
        if (nl) ACK(cudaMemcpyAsync(L.b, L.h_rhs.data(), sizeof(double) * (size_t)nl, cudaMemcpyHostToDevice, h->st));
        ACK(cudaStreamSynchronize(h->st));
        if (l + 1 < cfg->levels) {
            if ((rc = upload(L.hP, L.P, h->st))) return rc;
            HostCsr R = transpose(L.hP);
            if ((rc = upload(R, L.R, h->st))) return rc;
            ACK(cudaStreamSynchronize(h->st));
        }
        if ((rc = build_lex_schedule(h, L))) return rc;
        if ((rc = build_colouring(h, L))) return rc;
        if ((rc = build_sell(h, L))) return rc;
        max_blocks = std::max(max_blocks, (size_t)(nl * mgb::kLanes + 255) / 256 + 1);
    }
    ACK(cudaMalloc(&h->d_partial, sizeof(double) * max_blocks));
    ACK(cudaMalloc(&h->d_scal, sizeof(double) * 4));
    ACK(cudaMallocHost(&h->h_scal, sizeof(double) * 4));
    *out = h;
    return MGB_OK;
}

void mgb_amg_destroy(mgb_amg_t h)
{
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->st) cudaStreamSynchronize(h->st);
    for (auto &L : h->lv) {
        L.A.release(); L.P.release(); L.R.release(); L.lex.release(); L.colour.release(); L.sell.release();
        cudaFree(L.diag); cudaFree(L.x); cudaFree(L.b); cudaFree(L.tmp);
    }
    cudaFree(h->d_partial); cudaFree(h->d_scal);
    if (h->h_scal) cudaFreeHost(h->h_scal);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
}

int mgb_amg_level_info(mgb_amg_t h, int level, size_t *n, size_t *nnz_a, size_t *nnz_p, size_t *n_coarse,
                       int *n_waves, int *n_colours)
{
    if (!h || level < 0 || level >= (int)h->lv.size()) return mgb_set_error(MGB_ERR_ARG, "bad level");
    const AmgLevel &L = h->lv[level];
    if (n) *n = (size_t)L.hA.n_rows;
    if (nnz_a) *nnz_a = (size_t)L.hA.nnz();
    if (nnz_p) *nnz_p = (size_t)L.hP.nnz();
    if (n_coarse) *n_coarse = (size_t)L.hP.n_cols;
    if (n_waves) *n_waves = L.lex.n_groups;
    if (n_colours) *n_colours = L.colour.n_groups;
    return MGB_OK;
}

static int copy_csr(const HostCsr &M, int64_t *ptr, int64_t *col, double *val)
{
    for (int i = 0; i <= M.n_rows; ++i) ptr[i] = M.ptr[i];
    for (int k = 0; k < M.nnz(); ++k) { col[k] = M.col[k]; val[k] = M.val[k]; }
    return MGB_OK;
}
int mgb_amg_get_matrix(mgb_amg_t h, int level, int which, int64_t *ptr, int64_t *col, double *val)
{
    if (!h || level < 0 || level >= (int)h->lv.size() || !ptr || !col || !val) return mgb_set_error(MGB_ERR_ARG, "bad level/pointer");
    return copy_csr(which == 0 ? h->lv[level].hA : h->lv[level].hP, ptr, col, val);
}

int mgb_amg_get_schedule(mgb_amg_t h, int level, int which, int *group_of_row)
{
    if (!h || level < 0 || level >= (int)h->lv.size() || !group_of_row) return mgb_set_error(MGB_ERR_ARG, "bad level/pointer");
    const Schedule &S = which == 0 ? h->lv[level].lex : h->lv[level].colour;
    std::copy(S.h_group.begin(), S.h_group.end(), group_of_row);
    return MGB_OK;
}

int mgb_amg_get_vector(mgb_amg_t h, int level, int which, double *host)
{
    if (!h || level < 0 || level >= (int)h->lv.size() || !host) return mgb_set_error(MGB_ERR_ARG, "bad level/pointer");
    AmgLevel &L = h->lv[level];
    ACK(cudaSetDevice(h->cfg.device));
    const double *src = which == 0 ? L.x : (which == 1 ? L.b : L.tmp);
    if (L.A.n_rows) ACK(cudaMemcpyAsync(host, src, sizeof(double) * (size_t)L.A.n_rows, cudaMemcpyDeviceToHost, h->st));
    ACK(cudaStreamSynchronize(h->st));
    return MGB_OK;
}

int mgb_amg_set_vector(mgb_amg_t h, int level, int which, const double *host)
{
    if (!h || level < 0 || level >= (int)h->lv.size() || !host || which < 0 || which > 1) return mgb_set_error(MGB_ERR_ARG, "bad level/pointer");
    AmgLevel &L = h->lv[level];
    ACK(cudaSetDevice(h->cfg.device));
    if (L.A.n_rows) ACK(cudaMemcpyAsync(which == 0 ? L.x : L.b, host, sizeof(double) * (size_t)L.A.n_rows, cudaMemcpyHostToDevice, h->st));
    ACK(cudaStreamSynchronize(h->st));
    if (which == 1 && L.sell.n_slots) {           // keep the slot-ordered copy of the right-hand side in step
        std::vector<int> ros(L.sell.n_slots);
        ACK(cudaMemcpy(ros.data(), L.sell.row_of_slot, sizeof(int) * (size_t)L.sell.n_slots, cudaMemcpyDeviceToHost));
        std::vector<double> bs(L.sell.n_slots, 0.0);
        for (int p = 0; p < L.sell.n_slots; ++p) if (ros[p] >= 0) bs[p] = host[ros[p]];
        ACK(cudaMemcpy(L.sell.b_s, bs.data(), sizeof(double) * bs.size(), cudaMemcpyHostToDevice));
        L.h_rhs.assign(host, host + L.A.n_rows);
    }
    return MGB_OK;
}

int mgb_amg_smooth(mgb_amg_t h, int level, int kind, int sweeps)
{
    if (!h || level < 0 || level >= (int)h->lv.size()) return mgb_set_error(MGB_ERR_ARG, "Invalid level");   // AMG.cpp:237-240
    if (sweeps <= 0) return mgb_set_error(MGB_ERR_ARG, "Invalid number of iterations");                         // AMG.cpp:241-244
    ACK(cudaSetDevice(h->cfg.device));
    return do_smooth(h, level, kind, sweeps);
}

int mgb_amg_restrict(mgb_amg_t h, int level)
{
    if (!h || level < 1 || level >= (int)h->lv.size()) return mgb_set_error(MGB_ERR_ARG, "Level does not exist");   // AMG.cpp:51-57
    ACK(cudaSetDevice(h->cfg.device));
    return do_restrict(h, level);
}

int mgb_amg_prolong(mgb_amg_t h, int level)
{
    if (!h || level < 0 || level + 1 >= (int)h->lv.size()) return mgb_set_error(MGB_ERR_ARG, "Level does not exist");
    ACK(cudaSetDevice(h->cfg.device));
    return do_prolong(h, level);
}

int mgb_amg_residual(mgb_amg_t h, int level, double *norm)
{
    if (!h || level < 0 || level >= (int)h->lv.size() || !norm) return mgb_set_error(MGB_ERR_ARG, "bad level");
    ACK(cudaSetDevice(h->cfg.device));
    return do_residual(h, level, norm);
}

// AMG::apply_AMG after initialization (AMG.cpp:282-304)
int mgb_amg_apply(mgb_amg_t h, double *residual_norm)
{
    if (!h) return mgb_set_error(MGB_ERR_ARG, "null handle");
    ACK(cudaSetDevice(h->cfg.device));
    const int L = (int)h->lv.size();
    int rc, i;
    for (i = 0; i < L - 1; ++i) {
        if ((rc = do_smooth(h, i, h->cfg.smoother, h->cfg.pre_sweeps))) return rc;
        if ((rc = do_restrict(h, i + 1))) return rc;
    }
    if ((rc = do_smooth(h, i, h->cfg.smoother, h->cfg.coarse_sweeps))) return rc;
    for (i--; i >= 0; --i) {
        if ((rc = do_prolong(h, i))) return rc;
        if ((rc = do_smooth(h, i, h->cfg.smoother, h->cfg.post_sweeps))) return rc;
    }
    h->stats.cycles++;
    if (residual_norm) return do_residual(h, 0, residual_norm);
    return MGB_OK;
}

// Convergent correction-scheme V-cycle on the same hierarchy (NOT in the reference, whose
// one-pass scheme restricts the solution and is not an iteration): per level nu1 sweeps, r = b - A x, b_c = P^T r,
// x_c = 0, recurse, x += P x_c, nu2 sweeps; `coarse` sweeps on the last level.  hist[0] = ||b - A x0||_2, then one
// entry per cycle; stops when hist <= tol * hist[0] or after maxit cycles.
int mgb_amg_solve(mgb_amg_t h, double tol, int maxit, int nu1, int nu2, int coarse, double *hist, int *n_hist)
{
    if (!h || !hist || !n_hist || maxit < 0 || nu1 < 0 || nu2 < 0 || coarse < 1) return mgb_set_error(MGB_ERR_ARG, "bad argument");
    ACK(cudaSetDevice(h->cfg.device));
    const int L = (int)h->lv.size();
    const int kind = h->cfg.smoother;
    int rc, n = 0;
    double nrm = 0.;
    if ((rc = do_residual(h, 0, &nrm))) return rc;
    hist[n++] = nrm;
    const double target = tol * nrm;
    for (int it = 0; it < maxit && nrm > target; ++it) {
        for (int l = 0; l < L - 1; ++l) {                                  // downward
            AmgLevel &F = h->lv[l], &C = h->lv[l + 1];
            if (nu1 > 0 && (rc = do_smooth(h, l, kind, nu1))) return rc;
            {   // r = b - A x into F.tmp (no host read-back)
                const mgb::CsrDev A = F.A.view();
                if (h->cfg.exact_order) mgb::k_amg_residual<true><<<(A.n_rows + 255) / 256, 256, 0, h->st>>>(A, F.x, F.b, F.tmp, h->d_partial);
                else mgb::k_amg_sell<0><<<(F.sell.n_slots + 255) / 256, 256, 0, h->st>>>(F.sell.view(), F.x, F.sell.b_s, F.tmp, h->d_partial, 0, F.sell.n_slots);
                tally(h, sweep_bytes(F));
            }
            const mgb::CsrDev R = F.R.view();                              // b_c = P^T r
            if (R.n_rows) {
                if (h->cfg.exact_order) mgb::k_amg_spmv<true><<<(R.n_rows + 255) / 256, 256, 0, h->st>>>(R, F.tmp, C.b);
                else mgb::k_amg_spmv<false><<<(R.n_rows * mgb::kLanes + 255) / 256, 256, 0, h->st>>>(R, F.tmp, C.b);
                tally(h, 12. * R.nnz + 12. * R.n_rows + 8. * R.n_cols);
                if (C.sell.n_slots) {
                    mgb::k_amg_to_slots<<<(C.sell.n_slots + 255) / 256, 256, 0, h->st>>>(C.sell.row_of_slot, C.sell.n_slots, C.b, C.sell.b_s);
                    tally(h, 16. * C.A.n_rows);
                }
                ACK(cudaMemsetAsync(C.x, 0, sizeof(double) * (size_t)C.A.n_rows, h->st));
            }
            ACK(cudaGetLastError());
        }
        if ((rc = do_smooth(h, L - 1, kind, coarse))) return rc;
        for (int l = L - 2; l >= 0; --l) {                                  // upward
            if ((rc = do_prolong(h, l))) return rc;
            if (nu2 > 0 && (rc = do_smooth(h, l, kind, nu2))) return rc;
        }
        h->stats.cycles++;
        if ((rc = do_residual(h, 0, &nrm))) return rc;
        hist[n++] = nrm;
    }
    *n_hist = n;
    // the coarse right-hand sides were overwritten by restricted residuals: put the reference's P^T b back
    for (int l = 1; l < L; ++l) {
        AmgLevel &C = h->lv[l];
        if (!C.A.n_rows) continue;
        ACK(cudaMemcpyAsync(C.b, C.h_rhs.data(), sizeof(double) * (size_t)C.A.n_rows, cudaMemcpyHostToDevice, h->st));
        if (C.sell.n_slots) mgb::k_amg_to_slots<<<(C.sell.n_slots + 255) / 256, 256, 0, h->st>>>(C.sell.row_of_slot, C.sell.n_slots, C.b, C.sell.b_s);
    }
    ACK(cudaStreamSynchronize(h->st));
    return MGB_OK;
}

// ---- the setup stages one by one (RestrictionOperator's public methods, AMG/include/AMG.hpp:150-369) ----------------
// The reference's second driver (AMG/debugtest.cpp) calls them directly.  They run on the host in O(nnz) (the device
// setup is the next item of SURVEY.md section 8f); matrices cross the boundary as opaque host-CSR handles.
struct mgb_csr { HostCsr m; };

int mgb_csr_create(size_t n_rows, size_t n_cols, const int64_t *ptr, const int64_t *col, const double *val, mgb_csr_t *out)
{
    if (!ptr || !out || n_rows > (size_t)INT32_MAX || n_cols > (size_t)INT32_MAX || ptr[n_rows] > (int64_t)INT32_MAX)
        return mgb_set_error(MGB_ERR_ARG, "bad CSR arguments");
    mgb_csr *c = new mgb_csr();
    c->m.n_rows = (int)n_rows; c->m.n_cols = (int)n_cols;
    c->m.ptr.assign(n_rows + 1, 0);
    for (size_t i = 0; i < n_rows; ++i) {
        for (int64_t k = ptr[i]; k < ptr[i + 1]; ++k)
            if (val[k] != 0) { c->m.col.push_back((int)col[k]); c->m.val.push_back(val[k]); }     // CSRMatrix.cpp:13-14
        c->m.ptr[i + 1] = (int)c->m.col.size();
    }
    *out = c;
    return MGB_OK;
}
void mgb_csr_destroy(mgb_csr_t c) { delete c; }
int mgb_csr_info(mgb_csr_t c, size_t *n_rows, size_t *n_cols, size_t *nnz)
{
    if (!c) return mgb_set_error(MGB_ERR_ARG, "null matrix");
    if (n_rows) *n_rows = (size_t)c->m.n_rows;
    if (n_cols) *n_cols = (size_t)c->m.n_cols;
    if (nnz) *nnz = (size_t)c->m.nnz();
    return MGB_OK;
}
int mgb_csr_get(mgb_csr_t c, int64_t *ptr, int64_t *col, double *val)
{
    if (!c || !ptr || !col || !val) return mgb_set_error(MGB_ERR_ARG, "null argument");
    return copy_csr(c->m, ptr, col, val);
}
// RestrictionOperator::select_coarse_nodes (AMG.hpp:150-198); start < 0 selects n/2 (the reference draws it at random)
int mgb_amg_select_coarse_nodes(mgb_csr_t A, double eps, int64_t start, unsigned char *coarse_mask, size_t *n_coarse)
{
    if (!A || !coarse_mask || !n_coarse) return mgb_set_error(MGB_ERR_ARG, "null argument");
    std::vector<unsigned char> state;
    const int nc = split_coarse_fine(A->m, eps, start >= 0 ? (long)start : A->m.n_rows / 2, state);
    std::copy(state.begin(), state.end(), coarse_mask);
    *n_coarse = (size_t)nc;
    return MGB_OK;
}
// RestrictionOperator::build_prolongation_matrix (AMG.hpp:230-300)
int mgb_amg_build_prolongation(mgb_csr_t A, double eps, const unsigned char *coarse_mask, mgb_csr_t *P)
{
    if (!A || !coarse_mask || !P) return mgb_set_error(MGB_ERR_ARG, "null argument");
    std::vector<unsigned char> state(coarse_mask, coarse_mask + A->m.n_rows);
    int nc = 0;
    for (unsigned char b : state) nc += !(b & 0xC0);
    mgb_csr *c = new mgb_csr();
    c->m = interpolation(A->m, eps, state, nc);
    *P = c;
    return MGB_OK;
}
// RestrictionOperator::build_coarse_matrix (AMG.hpp:303-369)
int mgb_amg_build_coarse_matrix(mgb_csr_t A, mgb_csr_t P, mgb_csr_t *Ac)
{
    if (!A || !P || !Ac || P->m.n_rows != A->m.n_rows) return mgb_set_error(MGB_ERR_ARG, "bad Galerkin arguments");
    mgb_csr *c = new mgb_csr();
    c->m = galerkin(A->m, P->m);
    *Ac = c;
    return MGB_OK;
}

int mgb_amg_get_stats(mgb_amg_t h, mgb_gmg_stats *s)
{
    if (!h || !s) return mgb_set_error(MGB_ERR_ARG, "null argument");
    *s = h->stats;
    return MGB_OK;
}
int mgb_amg_reset_stats(mgb_amg_t h)
{
    if (!h) return mgb_set_error(MGB_ERR_ARG, "null handle");
    h->stats = mgb_gmg_stats{};
    return MGB_OK;
}
void *mgb_amg_stream(mgb_amg_t h) { return h ? (void *)h->st : nullptr; }
int mgb_amg_sync(mgb_amg_t h)
{
    if (!h) return mgb_set_error(MGB_ERR_ARG, "null handle");
    ACK(cudaSetDevice(h->cfg.device));
    ACK(cudaStreamSynchronize(h->st));
    return MGB_OK;
}

}  // extern "C"
