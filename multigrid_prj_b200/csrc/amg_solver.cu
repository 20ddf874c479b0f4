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
#include "amg_tail.cuh"
#include "amg_dev.h"
#include "dev_util.cuh"
#include "nccl_dyn.h"

#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

int mgb_set_error(int code, const std::string &msg);   // gmg_solver.cu

using namespace mgb::amg;

namespace {

#define ACK(call)                                                                             \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return mgb_set_error(MGB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)
#define ANK(call)                                                                             \
    do {                                                                                      \
        int e_ = (call);                                                                      \
        if (e_ != mgb::kNcclSuccess)                                                          \
            return mgb_set_error(MGB_ERR_NCCL, std::string(#call) + ": " + mgb::nccl().GetErrorString(e_)); \
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

// The per-row setup loops (interpolation weights, the two products of the Galerkin operator) are independent row by
// row: they run on the host threads in contiguous row chunks, every row with the same operations in the same order
// as the serial loop, so the hierarchy does not depend on the thread count.  (The reference's own OpenMP loop over
// these rows races on std::map, AMG/include/AMG.hpp:314-331.)
inline int setup_threads(int n_rows)
{
#ifdef _OPENMP
    return std::max(1, std::min(omp_get_max_threads(), n_rows / 4096 + 1));
#else
    (void)n_rows;
    return 1;
#endif
}

// out = the row chunks of `part` one after the other; out.ptr holds the row LENGTHS on entry (ptr[i + 1] = length of row i)
void concat_chunks(HostCsr &out, std::vector<HostCsr> &part, const std::vector<int> &first_row)
{
    for (int i = 0; i < out.n_rows; ++i) out.ptr[i + 1] += out.ptr[i];
    out.col.resize((size_t)out.ptr[out.n_rows]);
    out.val.resize((size_t)out.ptr[out.n_rows]);
    for (size_t t = 0; t < part.size(); ++t) {
        std::copy(part[t].col.begin(), part[t].col.end(), out.col.begin() + out.ptr[first_row[t]]);
        std::copy(part[t].val.begin(), part[t].val.end(), out.val.begin() + out.ptr[first_row[t]]);
        HostCsr().col.swap(part[t].col); HostCsr().val.swap(part[t].val);
    }
}

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
    const int T = setup_threads(n);
    std::vector<HostCsr> part(T);
    std::vector<int> first(T);
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int t = 0; t < T; ++t) {
        const int r0 = (int)((long long)n * t / T), r1 = (int)((long long)n * (t + 1) / T);
        first[t] = r0;
        HostCsr &Q = part[t];
        std::vector<int> strong;
        for (int i = r0; i < r1; ++i) {
            const size_t before = Q.col.size();
            if (!(state[i] & 0xC0)) { Q.col.push_back(cidx[i]); Q.val.push_back(1.0); }
            else {
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
                        if (w != 0) { Q.col.push_back(cidx[j]); Q.val.push_back(w); }     // exact zeros are dropped (CSRMatrix.cpp:13-14)
                    }
            }
            P.ptr[i + 1] = (int)(Q.col.size() - before);
        }
    }
    concat_chunks(P, part, first);
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

// out(i, :) = sum over the entries (k, v) of row i of M, in order, of v * B(k, :) -- every row by RowAccumulator::run
HostCsr row_product(const HostCsr &M, const HostCsr &B)
{
    HostCsr out;
    out.n_rows = M.n_rows; out.n_cols = B.n_cols; out.ptr.assign(M.n_rows + 1, 0);
    const int T = setup_threads(M.n_rows);
    std::vector<HostCsr> part(T);
    std::vector<int> first(T);
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int t = 0; t < T; ++t) {
        const int r0 = (int)((long long)M.n_rows * t / T), r1 = (int)((long long)M.n_rows * (t + 1) / T);
        first[t] = r0;
        RowAccumulator ra(B.n_cols);
        HostCsr &Q = part[t];
        for (int i = r0; i < r1; ++i) {
            const size_t before = Q.col.size();
            ra.run(i, &M.col[M.ptr[i]], &M.val[M.ptr[i]], M.ptr[i + 1] - M.ptr[i], B, Q);
            out.ptr[i + 1] = (int)(Q.col.size() - before);
        }
    }
    concat_chunks(out, part, first);
    return out;
}

// Galerkin operator in the reference's evaluation order (AMG.hpp:303-369):
//   PtA(i,j) = sum_k A(j,k) P(k,i)  (k ascending; A taken as symmetric),   Ac(i,j) = sum_k PtA(i,k) P(k,j)
HostCsr galerkin(const HostCsr &A, const HostCsr &P)
{
    HostCsr PtA = transpose(row_product(A, P));      // row j of A P holds PtA(:, j)
    return row_product(PtA, P);
}

inline int owner_of(int n, int n_ranks, int i)
{
    int r = (int)std::min<long long>((long long)i * n_ranks / std::max(n, 1), n_ranks - 1);
    while (r + 1 < n_ranks && block_of(n, n_ranks, r + 1).r0 <= i) ++r;
    while (r > 0 && block_of(n, n_ranks, r).r0 > i) --r;
    return r;
}

// Ghost-exchange plan of operator M for one rank: which vector entries (columns of M, global indices) it must receive
// because its rows reference them, and which of its own entries the other ranks' rows reference.  Ordered by
// (group, peer, index) so that one group -- one colour of a multicolour sweep -- is a contiguous range of the lists.
struct HaloPlan {
    int n_groups = 1, n_ranks = 1;
    std::vector<int> send_ptr, send_idx, recv_ptr, recv_idx;       // ptr: n_groups * n_ranks + 1
    int *d_send_idx = nullptr, *d_recv_idx = nullptr;
    // peer-store transport (csrc/p2p.cuh): for every entry of the send list the rank that needs it and its slot in that
    // rank's receive list
    unsigned char *d_send_peer = nullptr, *d_recv_src = nullptr;
    int *d_send_off = nullptr, *d_recv_pos = nullptr;
    int n_send() const { return send_ptr.empty() ? 0 : send_ptr.back(); }
    int n_recv() const { return recv_ptr.empty() ? 0 : recv_ptr.back(); }
    bool empty() const { return n_send() == 0 && n_recv() == 0; }
    void release()
    {
        cudaFree(d_send_idx); cudaFree(d_recv_idx); cudaFree(d_send_peer); cudaFree(d_send_off); cudaFree(d_recv_src); cudaFree(d_recv_pos);
        d_send_idx = d_recv_idx = d_send_off = d_recv_pos = nullptr; d_send_peer = d_recv_src = nullptr;
    }
};

void order_segments(int n_groups, int n_ranks, int n_cols, const int *group_of_col, std::vector<int> &entries /* (index) */,
                    const std::vector<int> &peer_of_entry, std::vector<int> &ptr, std::vector<int> &idx)
{
    const int n_seg = n_groups * n_ranks;
    ptr.assign(n_seg + 1, 0);
    auto seg_of = [&](size_t e) { return (group_of_col ? group_of_col[entries[e]] : 0) * n_ranks + peer_of_entry[e]; };
    for (size_t e = 0; e < entries.size(); ++e) ptr[seg_of(e) + 1]++;
    for (int q = 0; q < n_seg; ++q) ptr[q + 1] += ptr[q];
    idx.assign(entries.size(), 0);
    std::vector<int> fill(ptr.begin(), ptr.end() - 1);
    for (size_t e = 0; e < entries.size(); ++e) idx[fill[seg_of(e)]++] = entries[e];   // entries arrive ascending per peer
    (void)n_cols;
}

// the ghost entries of one rank, before any grouping: what it receives (with the owner of each entry) and what it sends
// (with the rank that needs each entry).  The scans over the matrix happen here, once; a plan is an ordering of these lists.
struct HaloEntries { std::vector<int> recv, recv_peer, send, send_peer; };

HaloEntries halo_entries(const HostCsr &M, int n_ranks, int rank)
{
    HaloEntries E;
    const Block mine_rows = block_of(M.n_rows, n_ranks, rank), mine_cols = block_of(M.n_cols, n_ranks, rank);
    // receive: columns outside my block that my rows reference (ascending, hence ascending per peer)
    {
        std::vector<unsigned char> need(M.n_cols, 0);
        for (int i = mine_rows.r0; i < mine_rows.r1; ++i)
            for (int k = M.ptr[i]; k < M.ptr[i + 1]; ++k) {
                const int j = M.col[k];
                if (j < mine_cols.r0 || j >= mine_cols.r1) need[j] = 1;
            }
        for (int j = 0; j < M.n_cols; ++j) if (need[j]) { E.recv.push_back(j); E.recv_peer.push_back(owner_of(M.n_cols, n_ranks, j)); }
    }
    // send: my entries that the rows of rank q reference, q by q (the same ascending order q derives for its receive list)
    {
        std::vector<int> stamp(std::max(mine_cols.size(), 1), -1);
        for (int q = 0; q < n_ranks; ++q) {
            if (q == rank) continue;
            const Block rows_q = block_of(M.n_rows, n_ranks, q);
            std::vector<int> hit;
            for (int i = rows_q.r0; i < rows_q.r1; ++i)
                for (int k = M.ptr[i]; k < M.ptr[i + 1]; ++k) {
                    const int j = M.col[k];
                    if (j >= mine_cols.r0 && j < mine_cols.r1 && stamp[j - mine_cols.r0] != q) { stamp[j - mine_cols.r0] = q; hit.push_back(j); }
                }
            std::sort(hit.begin(), hit.end());
            for (int j : hit) { E.send.push_back(j); E.send_peer.push_back(q); }
        }
    }
    return E;
}

HaloPlan plan_of(const HaloEntries &E, int n_cols, int n_ranks, const int *group_of_col, int n_groups)
{
    HaloPlan H;
    H.n_groups = std::max(n_groups, 1); H.n_ranks = n_ranks;
    std::vector<int> entries = E.recv;
    order_segments(H.n_groups, n_ranks, n_cols, group_of_col, entries, E.recv_peer, H.recv_ptr, H.recv_idx);
    entries = E.send;
    order_segments(H.n_groups, n_ranks, n_cols, group_of_col, entries, E.send_peer, H.send_ptr, H.send_idx);
    return H;
}

HaloPlan halo_plan(const HostCsr &M, int n_ranks, int rank, const int *group_of_col, int n_groups)
{
    return plan_of(halo_entries(M, n_ranks, rank), M.n_cols, n_ranks, group_of_col, n_groups);
}

struct AmgLevel {
    HostCsr hA, hP;                       // host copies (hierarchy queries, schedules)
    std::vector<double> h_rhs;
    DevCsr A, P, R;
    double *diag = nullptr, *x = nullptr, *b = nullptr, *tmp = nullptr;
    double *dl1 = nullptr;                     // a_ii + sum_{j != i} |a_ij| (l1-Jacobi)
    double *b0 = nullptr;                      // device-built hierarchy: the level's own right-hand side (P^T b), restored after mgb_amg_solve
    int *d_colour = nullptr;                   // device-built hierarchy: colour of every row (the host copy is made on demand)
    int *d_cf = nullptr;                       // device-built hierarchy: C/F state of every row (1 coarse, 0 fine)
    Schedule lex, colour;
    SellCopy sell, sellN, sellR, sellP;        // fast-path copies of A (colour-sorted / natural order) and of this rank's rows of R and P
    mgb::SellDev natural() const { mgb::SellDev v = sellN.view(); v.diag_s = diag; return v; }
    // sharding: rows [own.r0, own.r1) of this level are smoothed here (the whole level when it is replicated)
    bool sharded = false;
    Block own;                                 // rows this rank works on
    size_t own_nnz = 0;                        // entries of A in those rows
    HaloPlan haloA, haloA_colour;              // ghosts of x for A (all at once / colour by colour)
    HaloPlan haloR, haloP;                     // ghosts of the fine vector for R = P^T, of the coarse vector for P
    Block own_c;                               // rows of R this rank computes (block of the next level)
    bool x_halo_ok = true;                     // the ghost entries of x hold the owners' current values
};

}  // namespace

struct mgb_amg {
    mgb_amg_config cfg{};
    std::vector<AmgLevel> lv;
    cudaStream_t st = nullptr;
    double *d_partial = nullptr, *d_scal = nullptr, *h_scal = nullptr;
    mgb_gmg_stats stats{};
    int rank = 0, n_ranks = 1;
    mgb::NcclComm comm = nullptr;
    double *d_send = nullptr, *d_recv = nullptr;      // packed ghost entries
    // peer-store transport: one exported pool per rank = [header][two staging halves of the receive buffer]
    char *pool = nullptr;
    size_t pool_bytes = 0, stage_half = 0;
    mgb::P2PComm p2p;
    mgb::AmgPush push{};                              // what every push launch needs: the peers' staging buffers and flags
    double omega = 1.0;
    int lt = -1;                                      // first level of the persistent coarse tail (-1: none)
    int coop_max_blocks = 0;                          // co-resident CTAs of k_amg_sell_gs_sweeps (0: no cooperative launch)
    // one correction-scheme cycle captured as a CUDA graph (per sweep counts); void when a buffer swap moved x / tmp
    struct CycleGraph { int nu1, nu2, coarse; uint64_t epoch; cudaGraphExec_t exec; uint64_t launches; double bytes; };
    std::vector<CycleGraph> graphs;
    uint64_t ptr_epoch = 0;
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

// group_of_row covers the whole level (queries, halo plans); the device row lists hold the rows [own.r0, own.r1) only
int upload_schedule(const std::vector<int> &group_of_row, int n_groups, Schedule &S, cudaStream_t st, Block own)
{
    const int n = own.size();
    S.n_groups = n_groups;
    S.h_group = group_of_row;
    S.h_ptr.assign(n_groups + 1, 0);
    for (int i = own.r0; i < own.r1; ++i) S.h_ptr[group_of_row[i] + 1]++;
    for (int g = 0; g < n_groups; ++g) S.h_ptr[g + 1] += S.h_ptr[g];
    std::vector<int> rows(std::max(n, 1)), fill(S.h_ptr.begin(), S.h_ptr.end() - 1);
    for (int i = own.r0; i < own.r1; ++i) rows[fill[group_of_row[i]]++] = i;       // ascending inside a group
    ACK(cudaMalloc(&S.d_ptr, sizeof(int) * (size_t)(n_groups + 1)));
    ACK(cudaMalloc(&S.d_rows, sizeof(int) * (size_t)std::max(n, 1)));
    ACK(cudaMemcpyAsync(S.d_ptr, S.h_ptr.data(), sizeof(int) * (size_t)(n_groups + 1), cudaMemcpyHostToDevice, st));
    ACK(cudaMemcpyAsync(S.d_rows, rows.data(), sizeof(int) * (size_t)std::max(n, 1), cudaMemcpyHostToDevice, st));
    ACK(cudaStreamSynchronize(st));
    return MGB_OK;
}

inline void tally(mgb_amg *h, double bytes) { h->stats.kernel_launches++; h->stats.bytes_algorithmic += bytes; }
inline double sweep_bytes(const AmgLevel &L) { return 12. * (double)L.own_nnz + 28. * L.own.size(); }   // SURVEY.md section 8d, rows of this rank

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
    return upload_schedule(wave, n_waves, L.lex, h->st, L.own);
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
    return upload_schedule(colour, nc, L.colour, h->st, L.own);
}

// SELL-32 copy of the rows `row_of_slot` (already padded to slices, -1 = padding slot) of M; skip_diag drops a_ii.
// diag / rhs (may be null) are copied in slot order.
int upload_sell(const HostCsr &M, const std::vector<int> &row_of_slot, bool skip_diag, const double *rhs, SellCopy &S,
                bool slot_vectors = true)
{
    S.n_slots = (int)row_of_slot.size();
    const int n_slices = S.n_slots / 32;
    std::vector<int> slice_ptr(n_slices + 1, 0);
    auto row_len = [&](int i) {
        int len = M.ptr[i + 1] - M.ptr[i];
        if (skip_diag) for (int k = M.ptr[i]; k < M.ptr[i + 1]; ++k) len -= (M.col[k] == i);
        return len;
    };
    for (int s = 0; s < n_slices; ++s) {
        int longest = 0;
        for (int q = 0; q < 32; ++q) {
            const int i = row_of_slot[32 * s + q];
            if (i >= 0) longest = std::max(longest, row_len(i));
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
        for (int k = M.ptr[i]; k < M.ptr[i + 1]; ++k) {
            if (skip_diag && M.col[k] == i) continue;
            col[base + 32 * k2] = M.col[k]; val[base + 32 * k2] = M.val[k]; ++k2;
        }
        if (skip_diag) diag_s[p] = M.at(i, i);
        if (rhs) b_s[p] = rhs[i];
    }
    ACK(cudaMalloc(&S.slice_ptr, sizeof(int) * (size_t)(n_slices + 1)));
    ACK(cudaMalloc(&S.col, sizeof(int) * col.size()));
    ACK(cudaMalloc(&S.val, sizeof(double) * val.size()));
    ACK(cudaMalloc(&S.row_of_slot, sizeof(int) * (size_t)std::max(S.n_slots, 1)));
    if (slot_vectors) {
        ACK(cudaMalloc(&S.diag_s, sizeof(double) * diag_s.size()));
        ACK(cudaMalloc(&S.b_s, sizeof(double) * b_s.size()));
        ACK(cudaMemcpy(S.diag_s, diag_s.data(), sizeof(double) * diag_s.size(), cudaMemcpyHostToDevice));
        ACK(cudaMemcpy(S.b_s, b_s.data(), sizeof(double) * b_s.size(), cudaMemcpyHostToDevice));
    }
    ACK(cudaMemcpy(S.slice_ptr, slice_ptr.data(), sizeof(int) * (size_t)(n_slices + 1), cudaMemcpyHostToDevice));
    ACK(cudaMemcpy(S.col, col.data(), sizeof(int) * col.size(), cudaMemcpyHostToDevice));
    ACK(cudaMemcpy(S.val, val.data(), sizeof(double) * val.size(), cudaMemcpyHostToDevice));
    if (S.n_slots) ACK(cudaMemcpy(S.row_of_slot, row_of_slot.data(), sizeof(int) * (size_t)S.n_slots, cudaMemcpyHostToDevice));
    return MGB_OK;
}

// Inside windows of kSellWindow consecutive rows, longer rows first (SELL-C-sigma): the 32 rows of a slice then have
// nearly the same length and the padding of the slice almost vanishes, while rows stay close to their neighbours
// (the gathers of x keep their locality).  The rows of a list are mutually independent or order-free, so any order
// gives the same values.
constexpr int kSellWindow = 512;
void sort_windows_by_length(const HostCsr &M, std::vector<int> &rows)
{
    for (size_t w = 0; w < rows.size(); w += kSellWindow) {
        const size_t e = std::min(rows.size(), w + kSellWindow);
        std::stable_sort(rows.begin() + w, rows.begin() + e,
                         [&](int a, int b) { return M.ptr[a + 1] - M.ptr[a] > M.ptr[b + 1] - M.ptr[b]; });
    }
}

// colour-sorted SELL-32 copy of A (off-diagonal entries) + slot-ordered diagonal and rhs: rows of this rank only
int build_sell(mgb_amg *h, AmgLevel &L)
{
    const int ncol = L.colour.n_groups;
    SellCopy &S = L.sell;
    std::vector<int> row_of_slot;
    S.colour_slot_ptr.assign(ncol + 1, 0);
    std::vector<std::vector<int>> by_colour(std::max(ncol, 1));
    for (int i = L.own.r0; i < L.own.r1; ++i) by_colour[L.colour.h_group[i]].push_back(i);
    for (int c = 0; c < ncol; ++c) {
        S.colour_slot_ptr[c] = (int)row_of_slot.size();
        sort_windows_by_length(L.hA, by_colour[c]);
        row_of_slot.insert(row_of_slot.end(), by_colour[c].begin(), by_colour[c].end());
        while (row_of_slot.size() % 32) row_of_slot.push_back(-1);        // every colour starts on a slice boundary
    }
    S.colour_slot_ptr[ncol] = (int)row_of_slot.size();
    ACK(cudaMalloc(&S.d_colour_slot_ptr, sizeof(int) * (size_t)(ncol + 1)));
    ACK(cudaMemcpy(S.d_colour_slot_ptr, S.colour_slot_ptr.data(), sizeof(int) * (size_t)(ncol + 1), cudaMemcpyHostToDevice));
    int widest = 0;
    for (int c = 0; c < ncol; ++c) widest = std::max(widest, S.colour_slot_ptr[c + 1] - S.colour_slot_ptr[c]);
    S.coop_blocks = std::max(1, std::min((widest + 255) / 256, h->coop_max_blocks));
    return upload_sell(L.hA, row_of_slot, true, L.h_rhs.data(), S);
}

// SELL-32 copy of this rank's rows of A in natural order (off-diagonals; diag / rhs are read by row) for the kernels
// that visit every row once: Jacobi and the residual
int build_sell_natural(AmgLevel &L)
{
    std::vector<int> list;
    for (int i = L.own.r0; i < L.own.r1; ++i) list.push_back(i);
    sort_windows_by_length(L.hA, list);
    while (list.size() % 32) list.push_back(-1);
    return upload_sell(L.hA, list, true, nullptr, L.sellN, false);
}

// SELL-32 copy of a block of rows of a transfer operator (R = P^T for the restriction, P for the prolongation)
int build_sell_restriction(const HostCsr &R, Block rows, SellCopy &S)
{
    std::vector<int> list;
    for (int m = rows.r0; m < rows.r1; ++m) list.push_back(m);
    sort_windows_by_length(R, list);
    while (list.size() % 32) list.push_back(-1);
    return upload_sell(R, list, false, nullptr, S, false);
}

// ---- ghost exchange -----------------------------------------------------------------------------------------------
int upload_plan(HaloPlan &H)
{
    if (H.n_send()) {
        ACK(cudaMalloc(&H.d_send_idx, sizeof(int) * (size_t)H.n_send()));
        ACK(cudaMemcpy(H.d_send_idx, H.send_idx.data(), sizeof(int) * (size_t)H.n_send(), cudaMemcpyHostToDevice));
    }
    if (H.n_recv()) {
        ACK(cudaMalloc(&H.d_recv_idx, sizeof(int) * (size_t)H.n_recv()));
        ACK(cudaMemcpy(H.d_recv_idx, H.recv_idx.data(), sizeof(int) * (size_t)H.n_recv(), cudaMemcpyHostToDevice));
    }
    return MGB_OK;
}


// peer-store transport: where the entries of my send list land in the receive lists of the others.  Segment (g, me) of
// rank p's receive list holds exactly my send segment (g, p), in the same ascending order (halo_plan), so only the
// segment offsets of the other ranks are needed: one all-gather of the recv_ptr tables per plan.
int plan_p2p(mgb_amg *h, HaloPlan &H)
{
    const int R = h->n_ranks, me = h->rank;
    if (H.send_ptr.empty()) return MGB_OK;
    const int nseg = H.n_groups * R, len = nseg + 1;
    int *d = nullptr;
    ACK(cudaMalloc(&d, sizeof(int) * (size_t)len * (R + 1)));
    ACK(cudaMemcpyAsync(d + (size_t)len * R, H.recv_ptr.data(), sizeof(int) * (size_t)len, cudaMemcpyHostToDevice, h->st));
    ANK(mgb::nccl().AllGather(d + (size_t)len * R, d, sizeof(int) * (size_t)len, mgb::kNcclUint8, h->comm, h->st));
    std::vector<int> all((size_t)len * R);
    ACK(cudaMemcpyAsync(all.data(), d, sizeof(int) * (size_t)len * R, cudaMemcpyDeviceToHost, h->st));
    ACK(cudaStreamSynchronize(h->st));
    cudaFree(d);
    const int ns = H.n_send();
    std::vector<unsigned char> peer((size_t)std::max(ns, 1));
    std::vector<int> off((size_t)std::max(ns, 1));
    for (int g = 0; g < H.n_groups; ++g)
        for (int p = 0; p < R; ++p) {
            const int q = g * R + p, cnt = H.send_ptr[q + 1] - H.send_ptr[q];
            const int *theirs = all.data() + (size_t)p * len;
            if (cnt != theirs[g * R + me + 1] - theirs[g * R + me])
                return mgb_set_error(MGB_ERR_STATE, "ghost plans of two ranks disagree");
            for (int k = 0; k < cnt; ++k) { peer[H.send_ptr[q] + k] = (unsigned char)p; off[H.send_ptr[q] + k] = k; }   // slot inside my message to p
        }
    const int nr = H.n_recv();
    if (nr) {
        std::vector<unsigned char> src((size_t)nr);
        std::vector<int> pos((size_t)nr);
        for (int g = 0; g < H.n_groups; ++g)
            for (int p = 0; p < R; ++p)
                for (int k = H.recv_ptr[g * R + p]; k < H.recv_ptr[g * R + p + 1]; ++k) { src[k] = (unsigned char)p; pos[k] = k - H.recv_ptr[g * R + p]; }
        ACK(cudaMalloc(&H.d_recv_src, (size_t)nr));
        ACK(cudaMalloc(&H.d_recv_pos, sizeof(int) * (size_t)nr));
        ACK(cudaMemcpy(H.d_recv_src, src.data(), (size_t)nr, cudaMemcpyHostToDevice));
        ACK(cudaMemcpy(H.d_recv_pos, pos.data(), sizeof(int) * (size_t)nr, cudaMemcpyHostToDevice));
    }
    if (ns) {
        ACK(cudaMalloc(&H.d_send_peer, (size_t)ns));
        ACK(cudaMalloc(&H.d_send_off, sizeof(int) * (size_t)ns));
        ACK(cudaMemcpy(H.d_send_peer, peer.data(), (size_t)ns, cudaMemcpyHostToDevice));
        ACK(cudaMemcpy(H.d_send_off, off.data(), sizeof(int) * (size_t)ns, cudaMemcpyHostToDevice));
    }
    return MGB_OK;
}

// allocates and exports the pool, maps the peers', prepares every plan; leaves the NCCL transport in place on any failure
int setup_p2p(mgb_amg *h, size_t max_halo)
{
    // staging: one region per sending rank, two halves each; a half holds the largest message one rank can send another
    // (never more than a whole receive list) or one rank's block of the all-gathered first replicated level
    size_t half = max_halo;
    for (size_t l = 1; l < h->lv.size(); ++l)
        if (h->lv[l - 1].sharded && !h->lv[l].sharded) half = std::max(half, (size_t)h->lv[l].A.n_rows / h->n_ranks + 2);
    {   // the same value on every rank: a sender addresses the receiver's regions with it
        double *d = nullptr, v = (double)half;
        ACK(cudaMalloc(&d, sizeof(double)));
        ACK(cudaMemcpyAsync(d, &v, sizeof(double), cudaMemcpyHostToDevice, h->st));
        ANK(mgb::nccl().AllReduce(d, d, 1, mgb::kNcclFloat64, mgb::kNcclMax, h->comm, h->st));
        ACK(cudaMemcpyAsync(&v, d, sizeof(double), cudaMemcpyDeviceToHost, h->st));
        ACK(cudaStreamSynchronize(h->st));
        cudaFree(d);
        half = (size_t)v;
    }
    half = (half + 63) / 64 * 64;
    const size_t MB2 = (size_t)2 << 20;
    h->stage_half = half;
    h->pool_bytes = std::max((mgb::kP2PHeaderBytes + 2 * half * sizeof(double) * h->n_ranks + MB2 - 1) / MB2 * MB2, 2 * MB2);
    ACK(cudaMalloc(&h->pool, h->pool_bytes));
    ACK(cudaMemsetAsync(h->pool, 0, h->pool_bytes, h->st));
    h->p2p.init(h->pool, h->pool_bytes, h->rank, h->n_ranks, h->comm, h->st);
    if (!h->p2p.on) return MGB_OK;
    mgb::AmgPush &a = h->push;
    a = mgb::AmgPush{};
    for (int p = 0; p < h->n_ranks; ++p) {
        a.stage[p] = h->p2p.at<double>(p, mgb::kP2PHeaderBytes) + 2 * half * (size_t)h->rank;     // my region in rank p's buffer
        a.sig[p] = &h->p2p.hdr(p)->flags[h->rank][0];
    }
    a.half = half;
    a.pair_push = h->p2p.hdr(h->rank)->pair_push;
    a.done = &h->p2p.hdr(h->rank)->done[0];
    a.n_ranks = h->n_ranks; a.me = h->rank;
    int rc;
    for (auto &L : h->lv) {
        if (!L.sharded) continue;
        if ((rc = plan_p2p(h, L.haloA)) || (rc = plan_p2p(h, L.haloA_colour)) || (rc = plan_p2p(h, L.haloR)) || (rc = plan_p2p(h, L.haloP))) return rc;
    }
    return MGB_OK;
}

// refresh the ghost entries of `v` listed in groups [g0, g1) of the plan: pack -> grouped ncclSend/ncclRecv -> unpack,
// all on the compute stream
// the same over NVLink peer stores, neighbour to neighbour (amg_kernels.cuh: k_amg_push / k_amg_wait / k_amg_unpack_stage)
int exchange_p2p(mgb_amg *h, const HaloPlan &H, int g0, int g1, double *v)
{
    const int R = h->n_ranks;
    if (H.send_ptr.empty()) return MGB_OK;                    // no plan: no rank has one
    if (g1 != g0 + 1) return mgb_set_error(MGB_ERR_STATE, "peer-store exchange moves one group of a plan at a time");
    const int s0 = H.send_ptr[g0 * R], s1 = H.send_ptr[g1 * R], r0 = H.recv_ptr[g0 * R], r1 = H.recv_ptr[g1 * R];
    unsigned mask = 0;                                         // peers: either direction carries something in these groups
    for (int g = g0; g < g1; ++g)
        for (int p = 0; p < R; ++p) {
            const int q = g * R + p;
            if (H.send_ptr[q + 1] > H.send_ptr[q] || H.recv_ptr[q + 1] > H.recv_ptr[q]) mask |= 1u << p;
        }
    if (!mask) return MGB_OK;
    const int grid = std::max(1, std::min((s1 - s0 + 255) / 256, 296));
    mgb::k_amg_push<<<grid, 256, 0, h->st>>>(h->push, v, H.d_send_idx, H.d_send_peer, H.d_send_off, s0, s1, mask);
    tally(h, 12. * (s1 - s0));
    mgb::P2PHeader *hd = h->p2p.hdr(h->rank);
    mgb::k_amg_wait<<<1, 32, 0, h->st>>>(hd, mask);
    h->stats.kernel_launches++;
    if (r1 > r0) {
        mgb::k_amg_unpack_stage<<<(r1 - r0 + 255) / 256, 256, 0, h->st>>>(v, H.d_recv_idx, H.d_recv_src, H.d_recv_pos, reinterpret_cast<double *>(h->pool + mgb::kP2PHeaderBytes),
                                                                         (unsigned long long)h->stage_half, hd->pair_wait, r0, r1);
        tally(h, 12. * (r1 - r0));
    }
    ACK(cudaGetLastError());
    return MGB_OK;
}

int exchange(mgb_amg *h, const HaloPlan &H, int g0, int g1, double *v)
{
    if (h->n_ranks > 1 && h->p2p.on) return exchange_p2p(h, H, g0, g1, v);
    if (h->n_ranks == 1 || H.empty()) return MGB_OK;
    const int R = H.n_ranks;
    const int s0 = H.send_ptr[g0 * R], s1 = H.send_ptr[g1 * R], r0 = H.recv_ptr[g0 * R], r1 = H.recv_ptr[g1 * R];
    // (every rank walks the same (group, peer) segments, so an empty range here is empty on the peers' side too
    //  only segment by segment: the NCCL calls below are issued per non-empty segment)
    if (s1 > s0) {
        mgb::k_amg_pack<<<(s1 - s0 + 255) / 256, 256, 0, h->st>>>(v, H.d_send_idx, h->d_send, s0, s1);
        tally(h, 12. * (s1 - s0));
    }
    auto &N = mgb::nccl();
    if (s1 > s0 || r1 > r0) {
        ANK(N.GroupStart());
        for (int g = g0; g < g1; ++g)
            for (int p = 0; p < R; ++p) {
                const int q = g * R + p;
                const int ns = H.send_ptr[q + 1] - H.send_ptr[q], nr = H.recv_ptr[q + 1] - H.recv_ptr[q];
                if (ns) ANK(N.Send(h->d_send + H.send_ptr[q], (size_t)ns, mgb::kNcclFloat64, p, h->comm, h->st));
                if (nr) ANK(N.Recv(h->d_recv + H.recv_ptr[q], (size_t)nr, mgb::kNcclFloat64, p, h->comm, h->st));
            }
        ANK(N.GroupEnd());
    }
    if (r1 > r0) {
        mgb::k_amg_unpack<<<(r1 - r0 + 255) / 256, 256, 0, h->st>>>(v, H.d_recv_idx, h->d_recv, r0, r1);
        tally(h, 12. * (r1 - r0));
    }
    ACK(cudaGetLastError());
    return MGB_OK;
}
inline int exchange_all(mgb_amg *h, const HaloPlan &H, double *v) { return exchange(h, H, 0, H.n_groups, v); }

// every rank contributes its block of a replicated level's vector (in place, grouped send/recv)
int allgather_blocks(mgb_amg *h, int n, double *v)
{
    if (h->n_ranks == 1) return MGB_OK;
    auto &N = mgb::nccl();
    const Block mine = block_of(n, h->n_ranks, h->rank);
    if (h->p2p.on && (size_t)mine.size() <= h->stage_half && (size_t)n / h->n_ranks + 2 <= h->stage_half) {
        unsigned mask = 0;
        mgb::AmgBlocks bl{};
        for (int p = 0; p < h->n_ranks; ++p) { if (p != h->rank) mask |= 1u << p; bl.start[p] = block_of(n, h->n_ranks, p).r0; }
        bl.start[h->n_ranks] = n;
        const int grid = std::max(1, std::min((mine.size() + 255) / 256, 296));
        mgb::k_amg_push_block<<<grid, 256, 0, h->st>>>(h->push, v, mine.r0, mine.r1, mask);
        tally(h, 8. * mine.size() * (h->n_ranks - 1));
        mgb::P2PHeader *hd = h->p2p.hdr(h->rank);
        mgb::k_amg_wait<<<1, 32, 0, h->st>>>(hd, mask);
        h->stats.kernel_launches++;
        mgb::k_amg_unpack_blocks<<<(n + 255) / 256, 256, 0, h->st>>>(v, reinterpret_cast<double *>(h->pool + mgb::kP2PHeaderBytes),
                                                                    (unsigned long long)h->stage_half, hd->pair_wait, bl, h->n_ranks, h->rank);
        tally(h, 16. * (n - mine.size()));
        ACK(cudaGetLastError());
        return MGB_OK;
    }
    ANK(N.GroupStart());
    for (int p = 0; p < h->n_ranks; ++p) {
        if (p == h->rank) continue;
        const Block theirs = block_of(n, h->n_ranks, p);
        if (mine.size()) ANK(N.Send(v + mine.r0, (size_t)mine.size(), mgb::kNcclFloat64, p, h->comm, h->st));
        if (theirs.size()) ANK(N.Recv(v + theirs.r0, (size_t)theirs.size(), mgb::kNcclFloat64, p, h->comm, h->st));
    }
    ANK(N.GroupEnd());
    return MGB_OK;
}

inline int need_x_halo(mgb_amg *h, AmgLevel &L)
{
    if (L.x_halo_ok || !L.sharded) { L.x_halo_ok = true; return MGB_OK; }
    if (int rc = exchange_all(h, L.haloA, L.x)) return rc;
    L.x_halo_ok = true;
    return MGB_OK;
}

// the smoother mgb_amg_apply / mgb_amg_solve use on `level`
inline int kind_of(const mgb_amg *h, int level)
{
    return (level > 0 && h->cfg.coarse_smoother > 0) ? h->cfg.coarse_smoother : h->cfg.smoother;
}

int do_smooth(mgb_amg *h, int level, int kind, int sweeps)
{
    AmgLevel &L = h->lv[level];
    const mgb::CsrDev A = L.A.view();
    if (A.n_rows == 0 || sweeps <= 0) return MGB_OK;
    const int rows = L.own.size();
    int rc;
    if (kind == MGB_SMOOTH_GS_LEX) {
        if (L.sharded) return mgb_set_error(MGB_ERR_ARG, "lexicographic Gauss-Seidel is sequential across row blocks: sharded levels take the multicolour or Jacobi smoother");
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
        if (rows && !L.colour.d_rows) return mgb_set_error(MGB_ERR_STATE, "this level was set up without a colouring (its smoother is Jacobi-type)");
        if ((rc = need_x_halo(h, L))) return rc;
        const bool per_colour = L.sharded && !h->cfg.hybrid_gs;
        if (!h->cfg.exact_order && !per_colour && h->coop_max_blocks > 0 && L.sell.n_slots > 0) {
            // no ghost exchange between the colours: whole sweeps in one cooperative launch (all sweeps of the visit on
            // an unsharded level; one sweep + one exchange at a time for the hybrid smoother of a sharded level)
            mgb::SellDev S = L.sell.view();
            double *xp = L.x;
            const double *bs = L.sell.b_s;
            const int *csp = L.sell.d_colour_slot_ptr;
            int ncol = L.colour.n_groups, nsw = L.sharded ? 1 : sweeps;
            void *args[] = {&S, &xp, &bs, &csp, &ncol, &nsw};
            for (int s = 0; s < sweeps; s += nsw) {
                ACK(cudaLaunchCooperativeKernel((const void *)mgb::k_amg_sell_gs_sweeps, dim3(L.sell.coop_blocks), dim3(256), args, 0, h->st));
                tally(h, sweep_bytes(L) * nsw);
                if (L.sharded && (rc = exchange_all(h, L.haloA, L.x))) return rc;
            }
            return MGB_OK;
        }
        for (int s = 0; s < sweeps; ++s) {
            for (int c = 0; c < L.colour.n_groups; ++c) {
                const int a = L.colour.h_ptr[c], b = L.colour.h_ptr[c + 1];
                if (h->cfg.exact_order) {
                    if (b > a) mgb::k_amg_gs_rows_exact<<<(b - a + 255) / 256, 256, 0, h->st>>>(A, L.diag, L.x, L.b, L.colour.d_rows, a, b);
                } else {
                    const int p0 = L.sell.colour_slot_ptr[c], p1 = L.sell.colour_slot_ptr[c + 1];
                    if (p1 > p0)
                        mgb::k_amg_sell<2><<<(p1 - p0 + 255) / 256, 256, 0, h->st>>>(L.sell.view(), L.x, L.sell.b_s, L.x, nullptr, p0, p1, 1.0, nullptr);
                }
                if (rows) tally(h, sweep_bytes(L) * (double)(b - a) / rows);
                // the rows of the next colours (here and on the peers) read this colour's new values
                if (per_colour && (rc = exchange(h, L.haloA_colour, c, c + 1, L.x))) return rc;
            }
            if (L.sharded && h->cfg.hybrid_gs && (rc = exchange_all(h, L.haloA, L.x))) return rc;
        }
    } else if (kind == MGB_SMOOTH_JACOBI || kind == MGB_SMOOTH_L1_JACOBI) {
        if ((rc = need_x_halo(h, L))) return rc;
        if (kind == MGB_SMOOTH_L1_JACOBI && (h->cfg.exact_order || !L.sellN.n_slots) && rows)
            return mgb_set_error(MGB_ERR_ARG, "l1-Jacobi runs on the SELL copies of the fast path (exact_order = 0)");
        // the sweeps alternate between x and tmp; the pointers themselves stay put (an odd count ends with one copy), so a
        // CUDA graph that holds them stays valid
        double *src = L.x, *dst = L.tmp;
        for (int s = 0; s < sweeps; ++s) {
            if (kind == MGB_SMOOTH_L1_JACOBI) {
                if (rows) mgb::k_amg_sell<5, true><<<(L.sellN.n_slots + 255) / 256, 256, 0, h->st>>>(L.natural(), src, L.b, dst, nullptr, 0, L.sellN.n_slots, 1.0, L.dl1);
            } else if (h->cfg.exact_order) {
                if (rows) mgb::k_amg_jacobi_vec<<<(unsigned)(((size_t)rows * mgb::kLanes + 255) / 256), 256, 0, h->st>>>(A, L.diag, src, L.b, dst, h->omega, L.own.r0, L.own.r1);
            } else if (L.sellN.n_slots)
                mgb::k_amg_sell<1, true><<<(L.sellN.n_slots + 255) / 256, 256, 0, h->st>>>(L.natural(), src, L.b, dst, nullptr, 0, L.sellN.n_slots, h->omega, nullptr);
            tally(h, sweep_bytes(L));
            std::swap(src, dst);
            if (L.sharded && (rc = exchange_all(h, L.haloA, src))) return rc;       // the new iterate has no ghosts yet
        }
        if (src != L.x) {
            ACK(cudaMemcpyAsync(L.x, src, sizeof(double) * (size_t)A.n_rows, cudaMemcpyDeviceToDevice, h->st));
            tally(h, 16. * A.n_rows);
        }
    } else
        return mgb_set_error(MGB_ERR_ARG, "unknown AMG smoother");
    ACK(cudaGetLastError());
    return MGB_OK;
}

// r = b - A x on this rank's rows into L.tmp, sum of squares into d_scal[0] (all ranks: the global sum)
int residual_to_tmp(mgb_amg *h, AmgLevel &L, bool want_norm)
{
    const mgb::CsrDev A = L.A.view();
    int rc, blocks;
    if ((rc = need_x_halo(h, L))) return rc;
    const int rows = L.own.size();
    if (h->cfg.exact_order) {
        blocks = (rows + 255) / 256;
        if (blocks) mgb::k_amg_residual<true><<<blocks, 256, 0, h->st>>>(A, L.x, L.b, L.tmp, h->d_partial, L.own.r0, L.own.r1);
    } else {
        blocks = (L.sellN.n_slots + 255) / 256;
        if (blocks) mgb::k_amg_sell<0, true><<<blocks, 256, 0, h->st>>>(L.natural(), L.x, L.b, L.tmp, h->d_partial, 0, L.sellN.n_slots, 1.0, nullptr);
    }
    tally(h, sweep_bytes(L));
    if (want_norm) {
        mgb::k_amg_reduce<<<1, 1024, 0, h->st>>>(h->d_partial, blocks, h->d_scal);
        tally(h, 0.);
        if (L.sharded) ANK(mgb::nccl().AllReduce(h->d_scal, h->d_scal, 1, mgb::kNcclFloat64, mgb::kNcclSum, h->comm, h->st));
    }
    ACK(cudaGetLastError());
    return MGB_OK;
}

int do_residual(mgb_amg *h, int level, double *norm)
{
    AmgLevel &L = h->lv[level];
    if (int rc = residual_to_tmp(h, L, true)) return rc;
    ACK(cudaMemcpyAsync(h->h_scal, h->d_scal, sizeof(double), cudaMemcpyDeviceToHost, h->st));
    ACK(cudaStreamSynchronize(h->st));
    *norm = std::sqrt(h->h_scal[0]);
    return MGB_OK;
}

// out_{level} = P^T in_{level-1} as a gather over R = P^T (AMG.cpp:50-74); `in` is the fine level's x or tmp
int restrict_vec(mgb_amg *h, int level, double *in, double *out)
{
    AmgLevel &F = h->lv[level - 1], &C = h->lv[level];
    const mgb::CsrDev R = F.R.view();
    if (R.n_rows == 0) return MGB_OK;
    int rc;
    if (F.sharded && (rc = exchange_all(h, F.haloR, in))) return rc;        // fine entries of other blocks my coarse rows gather
    const int rows = F.own_c.size();
    if (rows) {
        if (h->cfg.exact_order) mgb::k_amg_spmv<true><<<(rows + 255) / 256, 256, 0, h->st>>>(R, in, out, F.own_c.r0, F.own_c.r1);
        else mgb::k_amg_sell<3><<<(F.sellR.n_slots + 255) / 256, 256, 0, h->st>>>(F.sellR.view(), in, nullptr, out, nullptr, 0, F.sellR.n_slots, 1.0, nullptr);
    }
    const double share = R.n_rows ? (double)rows / R.n_rows : 0.;
    tally(h, (12. * R.nnz + 12. * R.n_rows + 8. * R.n_cols) * share);
    ACK(cudaGetLastError());
    if (F.sharded && !C.sharded && (rc = allgather_blocks(h, R.n_rows, out))) return rc;   // first replicated level: whole on every rank
    return MGB_OK;
}

// x_{level} = P^T x_{level-1}
int do_restrict(mgb_amg *h, int level)
{
    AmgLevel &F = h->lv[level - 1], &C = h->lv[level];
    if (int rc = restrict_vec(h, level, F.x, C.x)) return rc;
    C.x_halo_ok = !C.sharded;
    return MGB_OK;
}

// x_level += P x_{level+1}  (AMG.cpp:218-232)
int do_prolong(mgb_amg *h, int level)
{
    AmgLevel &F = h->lv[level], &C = h->lv[level + 1];
    const mgb::CsrDev P = F.P.view();
    if (P.n_rows == 0) return MGB_OK;
    int rc;
    if (C.sharded && (rc = exchange_all(h, F.haloP, C.x))) return rc;        // coarse entries of other blocks my fine rows interpolate from
    const int rows = F.own.size();
    if (rows && !h->cfg.exact_order)
        mgb::k_amg_sell<4><<<(F.sellP.n_slots + 255) / 256, 256, 0, h->st>>>(F.sellP.view(), C.x, nullptr, F.x, nullptr, 0, F.sellP.n_slots, 1.0, nullptr);
    else if (rows) mgb::k_amg_prolong_add<<<(rows + 255) / 256, 256, 0, h->st>>>(P, C.x, F.x, F.own.r0, F.own.r1);
    tally(h, (12. * P.nnz + 20. * P.n_rows + 8. * P.n_cols) * (P.n_rows ? (double)rows / P.n_rows : 0.));
    ACK(cudaGetLastError());
    F.x_halo_ok = !F.sharded;
    return MGB_OK;
}

// levels lt .. L-1 in one launch (amg_tail.cuh).  mode 0: the reference's pass, mode 1: correction scheme.
int launch_tail(mgb_amg *h, int mode, int pre, int coarse, int post)
{
    const int L = (int)h->lv.size();
    mgb::AmgTailParams p{};
    p.nlev = L - h->lt;
    p.exact = h->cfg.exact_order; p.mode = mode;
    p.pre = pre; p.coarse = coarse; p.post = post; p.omega = h->omega;
    double bytes = 0.;
    for (int l = h->lt; l < L; ++l) {
        AmgLevel &lv = h->lv[l];
        const int kind = kind_of(h, l);
        const Schedule &S = kind == MGB_SMOOTH_GS_LEX ? lv.lex : lv.colour;
        if ((kind == MGB_SMOOTH_GS_LEX || kind == MGB_SMOOTH_GS_RB) && lv.A.n_rows && !S.d_rows)
            return mgb_set_error(MGB_ERR_STATE, "this level holds no row schedule for the requested Gauss-Seidel smoother");
        mgb::AmgTailLevel &t = p.lv[l - h->lt];
        t.A = lv.A.view(); t.P = lv.P.view(); t.R = lv.R.view();
        t.diag = lv.diag; t.dl1 = lv.dl1; t.kind = kind; t.x = lv.x; t.b = lv.b; t.tmp = lv.tmp;
        t.grp_ptr = S.d_ptr; t.grp_rows = S.d_rows; t.n_groups = S.n_groups;
        const int sweeps = (l == L - 1) ? coarse : pre + post;
        bytes += sweep_bytes(lv) * (sweeps + (mode == 1 && l < L - 1 ? 1 : 0));
        if (l < L - 1) bytes += 2. * (12. * lv.P.nnz + 16. * lv.P.n_rows + 10. * lv.P.n_cols);
        lv.x_halo_ok = true;
    }
    mgb::k_amg_tail<<<1, mgb::kAmgTailThreads, 0, h->st>>>(p);
    tally(h, bytes);
    ACK(cudaGetLastError());
    return MGB_OK;
}

// ---- hierarchy built ON THE DEVICE (csrc/amg_setup.cu; parity-exempt fast path) -----------------------------------------------
int download(const DevCsr &D, HostCsr &H)
{
    H.n_rows = D.n_rows; H.n_cols = D.n_cols;
    H.ptr.assign((size_t)D.n_rows + 1, 0); H.col.assign((size_t)D.nnz, 0); H.val.assign((size_t)D.nnz, 0.0);
    if (D.ptr) ACK(cudaMemcpy(H.ptr.data(), D.ptr, sizeof(int) * ((size_t)D.n_rows + 1), cudaMemcpyDeviceToHost));
    if (D.nnz) {
        ACK(cudaMemcpy(H.col.data(), D.col, sizeof(int) * (size_t)D.nnz, cudaMemcpyDeviceToHost));
        ACK(cudaMemcpy(H.val.data(), D.val, sizeof(double) * (size_t)D.nnz, cudaMemcpyDeviceToHost));
    }
    return MGB_OK;
}

__global__ void __launch_bounds__(256)
k_int_max(const int *__restrict__ v, int n, int *out)
{
    int m = -1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, v[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// Jones-Plassmann colouring on the device; the colours stay there (L.d_colour) and the row lists / the colour-sorted
// SELL copy are built from them without a host pass
int build_colouring_device(mgb_amg *h, AmgLevel &L)
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
    cudaFree(c1);
    L.d_colour = c0;
    if (left > 0) { cudaFree(d_left); return mgb_set_error(MGB_ERR_STATE, "colouring did not finish"); }
    int top = -1;
    ACK(cudaMemsetAsync(d_left, 0xFF, sizeof(int), h->st));
    if (n) k_int_max<<<std::min((n + 255) / 256, 1184), 256, 0, h->st>>>(c0, n, d_left);
    ACK(cudaMemcpyAsync(&top, d_left, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    ACK(cudaStreamSynchronize(h->st));
    cudaFree(d_left);
    return dev_group_schedule(L.A, L.d_colour, top + 1, L.own, L.colour, h->cfg.exact_order ? nullptr : &L.sell, L.diag, L.b, h->st);
}

// lv[0].A and lv[0].b are in place on the device (lv sized cfg.levels); builds everything else
int build_levels_device(mgb_amg *h, size_t &max_blocks, size_t &max_halo)
{
    const mgb_amg_config *cfg = &h->cfg;
    const int n_ranks = h->n_ranks, rank = h->rank;
    const int min_rows = cfg->shard_min_rows > 0 ? cfg->shard_min_rows : 131072;
    if (cfg->exact_order) return mgb_set_error(MGB_ERR_ARG, "device_setup = 1 runs the fast kernels only (exact_order = 0)");
    if (cfg->smoother == MGB_SMOOTH_GS_LEX)
        return mgb_set_error(MGB_ERR_ARG, "device_setup = 1: lexicographic Gauss-Seidel needs the host-built level schedule");
    int rc;
    mgb::dev::Trace tr("levels");
    // 1. coarsening, level by level, on the device
    int depth = 1;
    for (int l = 0; l + 1 < cfg->levels; ++l) {
        AmgLevel &F = h->lv[l];
        if (F.A.n_rows <= 16) break;
        DevCsr P, R, Ac;
        int rounds = 0;
        if ((rc = dev_coarsen(F.A, cfg->eps, 12345u + 7919u * (unsigned)l, P, R, Ac, h->st, &rounds, &F.d_cf))) { P.release(); R.release(); Ac.release(); return rc; }
        if (P.n_cols <= 1 || P.n_cols >= F.A.n_rows || !P.ptr) { P.release(); R.release(); Ac.release(); break; }
        F.P = P; F.R = R;
        AmgLevel &C = h->lv[l + 1];
        C.A = Ac;
        ACK(cudaMalloc(&C.b, sizeof(double) * (size_t)Ac.n_rows));
        // right-hand side of the next level as the reference forms it: b_c = P^T b (AMG.cpp:100-109)
        mgb::k_amg_spmv<false><<<(unsigned)(((size_t)R.n_rows * mgb::kLanes + 255) / 256), 256, 0, h->st>>>(R.view(), F.b, C.b, 0, R.n_rows);
        ACK(cudaGetLastError());
        depth = l + 2;
    }
    h->lv.resize(depth);
    const int L = depth;
    tr.mark("hierarchy (levels)", depth);
    // 2. vectors and diagonals
    for (int l = 0; l < L; ++l) {
        AmgLevel &Lv = h->lv[l];
        const size_t bytes = sizeof(double) * (size_t)std::max(Lv.A.n_rows, 1);
        ACK(cudaMalloc(&Lv.diag, bytes)); ACK(cudaMalloc(&Lv.dl1, bytes)); ACK(cudaMalloc(&Lv.x, bytes)); ACK(cudaMalloc(&Lv.tmp, bytes));
        ACK(cudaMalloc(&Lv.b0, bytes));
        ACK(cudaMemsetAsync(Lv.x, 0, bytes, h->st));
        ACK(cudaMemsetAsync(Lv.tmp, 0, bytes, h->st));
        ACK(cudaMemcpyAsync(Lv.b0, Lv.b, sizeof(double) * (size_t)Lv.A.n_rows, cudaMemcpyDeviceToDevice, h->st));
        if ((rc = dev_diagonals(Lv.A, Lv.diag, Lv.dl1, h->st))) return rc;
    }
    ACK(cudaStreamSynchronize(h->st));
    // 3. which levels are cut into row blocks, and the rows of every level this rank works on
    for (int l = 0; l < L; ++l) {
        AmgLevel &Lv = h->lv[l];
        const int nl = Lv.A.n_rows;
        Lv.sharded = n_ranks > 1 && (long long)nl >= (long long)min_rows * n_ranks && (l == 0 || h->lv[l - 1].sharded);
        Lv.own = Lv.sharded ? block_of(nl, n_ranks, rank) : Block{0, nl};
        int p0 = 0, p1 = Lv.A.nnz;
        if (Lv.sharded) {
            ACK(cudaMemcpy(&p0, Lv.A.ptr + Lv.own.r0, sizeof(int), cudaMemcpyDeviceToHost));
            ACK(cudaMemcpy(&p1, Lv.A.ptr + Lv.own.r1, sizeof(int), cudaMemcpyDeviceToHost));
        }
        Lv.own_nnz = (size_t)(p1 - p0);
        Lv.x_halo_ok = true;
    }
    // 4. persistent coarse tail: the trailing run of small, unsharded levels
    {
        const int cap = cfg->tail_max_rows == 0 ? 4000 : cfg->tail_max_rows;
        h->lt = -1;
        for (int l = L - 1; l >= 0 && cap > 0; --l) {
            if (h->lv[l].A.n_rows > cap || h->lv[l].sharded || L - l > mgb::kAmgTailMaxLevels) break;
            h->lt = l;
        }
    }
    tr.mark("vectors, diagonals, layout");
    // 5. SELL copies, colour lists, ghost plans
    for (int l = 0; l < L; ++l) {
        AmgLevel &Lv = h->lv[l];
        const int nl = Lv.A.n_rows;
        if (l + 1 < L) {
            const AmgLevel &C = h->lv[l + 1];
            Lv.own_c = Lv.sharded ? block_of(Lv.R.n_rows, n_ranks, rank) : Block{0, Lv.R.n_rows};
            if ((rc = dev_build_sell_range(Lv.R, Lv.own_c, false, Lv.sellR, h->st))) return rc;
            if ((rc = dev_build_sell_range(Lv.P, Lv.own, false, Lv.sellP, h->st))) return rc;
            if (Lv.sharded) {
                HostCsr hR, hPm;
                if ((rc = download(Lv.R, hR))) return rc;
                Lv.haloR = halo_plan(hR, n_ranks, rank, nullptr, 1);
                if ((rc = upload_plan(Lv.haloR))) return rc;
                if (C.sharded) {
                    if ((rc = download(Lv.P, hPm))) return rc;
                    Lv.haloP = halo_plan(hPm, n_ranks, rank, nullptr, 1);
                    if ((rc = upload_plan(Lv.haloP))) return rc;
                }
                max_halo = std::max<size_t>(max_halo, std::max({Lv.haloR.n_send(), Lv.haloR.n_recv(), Lv.haloP.n_send(), Lv.haloP.n_recv()}));
            }
        }
        if ((rc = dev_build_sell_range(Lv.A, Lv.own, true, Lv.sellN, h->st))) return rc;
        if (l == 0) tr.mark("L0 SELL copies (A, R, P)");
        const bool coloured = kind_of(h, l) == MGB_SMOOTH_GS_RB;
        if (coloured && (rc = build_colouring_device(h, Lv))) return rc;
        if (l == 0 && coloured) tr.mark("L0 colouring + colour SELL");
        if (Lv.sharded) {
            HostCsr hAm;
            if ((rc = download(Lv.A, hAm))) return rc;
            const HaloEntries EA = halo_entries(hAm, n_ranks, rank);          // one pass over the matrix serves both orderings
            Lv.haloA = plan_of(EA, hAm.n_cols, n_ranks, nullptr, 1);
            if ((rc = upload_plan(Lv.haloA))) return rc;
            if (coloured) {
                Lv.colour.h_group.assign((size_t)nl, 0);
                ACK(cudaMemcpy(Lv.colour.h_group.data(), Lv.d_colour, sizeof(int) * (size_t)nl, cudaMemcpyDeviceToHost));
                Lv.haloA_colour = plan_of(EA, hAm.n_cols, n_ranks, Lv.colour.h_group.data(), Lv.colour.n_groups);
                if ((rc = upload_plan(Lv.haloA_colour))) return rc;
            }
            max_halo = std::max<size_t>(max_halo, std::max(Lv.haloA.n_send(), Lv.haloA.n_recv()));
        }
        max_blocks = std::max(max_blocks, ((size_t)nl * mgb::kLanes + 255) / 256 + 1);
    }
    tr.mark("remaining levels: SELL, plans");
    return MGB_OK;
}

int build_levels_device_from_host(mgb_amg *h, size_t n, const int64_t *ptr, const int64_t *col, const double *val, const double *rhs,
                                  size_t &max_blocks, size_t &max_halo)
{
    h->lv.resize(h->cfg.levels);
    HostCsr &A = h->lv[0].hA;                    // the level-0 operator as given (exact zeros dropped, CSRMatrix.cpp:13-14)
    A.n_rows = A.n_cols = (int)n;
    A.ptr.assign(n + 1, 0);
    A.col.reserve((size_t)ptr[n]); A.val.reserve((size_t)ptr[n]);
    for (size_t i = 0; i < n; ++i) {
        for (int64_t k = ptr[i]; k < ptr[i + 1]; ++k) {
            if (col[k] < 0 || col[k] >= (int64_t)n) return mgb_set_error(MGB_ERR_ARG, "column index out of range");
            if (k > ptr[i] && col[k] <= col[k - 1]) return mgb_set_error(MGB_ERR_ARG, "rows must be sorted by column without duplicates");
            if (val[k] != 0) { A.col.push_back((int)col[k]); A.val.push_back(val[k]); }
        }
        A.ptr[i + 1] = (int)A.col.size();
    }
    h->lv[0].h_rhs.assign(rhs, rhs + n);
    int rc;
    if ((rc = upload(A, h->lv[0].A, h->st))) return rc;
    ACK(cudaMalloc(&h->lv[0].b, sizeof(double) * n));
    ACK(cudaMemcpyAsync(h->lv[0].b, rhs, sizeof(double) * n, cudaMemcpyHostToDevice, h->st));
    ACK(cudaStreamSynchronize(h->st));
    return build_levels_device(h, max_blocks, max_halo);
}

// ---- hierarchy with the reference's semantics, built on the host (parity path) --------------------------------------------
int build_levels_host(mgb_amg *h, size_t n, const int64_t *ptr, const int64_t *col, const double *val, const double *rhs,
                      size_t &max_blocks, size_t &max_halo)
{
    const mgb_amg_config *cfg = &h->cfg;
    const int n_ranks = h->n_ranks, rank = h->rank;
    const int min_rows = cfg->shard_min_rows > 0 ? cfg->shard_min_rows : 131072;
    h->lv.resize(cfg->levels);
    // level 0: CSRMatrix::copy_from drops exact zeros (CSRMatrix.cpp:13-14); rows must be column-sorted (they come from a map)
    {
        HostCsr &A = h->lv[0].hA;
        A.n_rows = A.n_cols = (int)n;
        A.ptr.assign(n + 1, 0);
        for (size_t i = 0; i < n; ++i) {
            for (int64_t k = ptr[i]; k < ptr[i + 1]; ++k) {
                if (col[k] < 0 || col[k] >= (int64_t)n) return mgb_set_error(MGB_ERR_ARG, "column index out of range");
                if (k > ptr[i] && col[k] <= col[k - 1]) return mgb_set_error(MGB_ERR_ARG, "rows must be sorted by column without duplicates");
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
    // which levels are cut into row blocks, and the rows of every level this rank works on
    for (int l = 0; l < cfg->levels; ++l) {
        AmgLevel &L = h->lv[l];
        const int nl = L.hA.n_rows;
        L.sharded = n_ranks > 1 && (long long)nl >= (long long)min_rows * n_ranks && (l == 0 || h->lv[l - 1].sharded);
        L.own = L.sharded ? block_of(nl, n_ranks, rank) : Block{0, nl};
        L.own_nnz = (size_t)(L.hA.ptr[L.own.r1] - L.hA.ptr[L.own.r0]);
        L.x_halo_ok = true;                                   // x = 0 everywhere
    }
    // persistent coarse tail: the trailing run of small, unsharded levels (at most kAmgTailMaxLevels of them)
    {
        const int cap = cfg->tail_max_rows == 0 ? 4000 : cfg->tail_max_rows;
        h->lt = -1;
        for (int l = cfg->levels - 1; l >= 0 && cap > 0; --l) {
            if (h->lv[l].hA.n_rows > cap || h->lv[l].sharded || cfg->levels - l > mgb::kAmgTailMaxLevels) break;
            h->lt = l;
        }
    }
    // upload
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
        ACK(cudaMemsetAsync(L.tmp, 0, bytes, h->st));
        if (nl) ACK(cudaMemcpyAsync(L.b, L.h_rhs.data(), sizeof(double) * (size_t)nl, cudaMemcpyHostToDevice, h->st));
        ACK(cudaMalloc(&L.dl1, bytes));
        if ((rc = dev_diagonals(L.A, L.diag, L.dl1, h->st))) return rc;      // (rewrites diag with the same values)
        ACK(cudaStreamSynchronize(h->st));
        if (l + 1 < cfg->levels) {
            if ((rc = upload(L.hP, L.P, h->st))) return rc;
            HostCsr R = transpose(L.hP);
            if ((rc = upload(R, L.R, h->st))) return rc;
            ACK(cudaStreamSynchronize(h->st));
            const AmgLevel &C = h->lv[l + 1];
            // rows of R this rank gathers: its block of the coarse level whenever the fine level is sharded (a replicated
            // coarse level is then completed by an all-gather), everything otherwise
            L.own_c = L.sharded ? block_of(R.n_rows, n_ranks, rank) : Block{0, R.n_rows};
            if (!cfg->exact_order && (rc = build_sell_restriction(R, L.own_c, L.sellR))) return rc;
            if (!cfg->exact_order && (rc = build_sell_restriction(L.hP, L.own, L.sellP))) return rc;
            if (L.sharded) {
                L.haloR = halo_plan(R, n_ranks, rank, nullptr, 1);
                if ((rc = upload_plan(L.haloR))) return rc;
                if (C.sharded) {
                    L.haloP = halo_plan(L.hP, n_ranks, rank, nullptr, 1);
                    if ((rc = upload_plan(L.haloP))) return rc;
                }
                max_halo = std::max<size_t>(max_halo, std::max({L.haloR.n_send(), L.haloR.n_recv(), L.haloP.n_send(), L.haloP.n_recv()}));
            }
        }
        if ((rc = build_lex_schedule(h, L))) return rc;
        if ((rc = build_colouring(h, L))) return rc;       // on the whole graph: every rank derives the same colours
        if ((rc = build_sell(h, L))) return rc;
        if (!cfg->exact_order && (rc = build_sell_natural(L))) return rc;
        if (L.sharded) {
            const HaloEntries EA = halo_entries(L.hA, n_ranks, rank);         // one pass over the matrix serves both orderings
            L.haloA = plan_of(EA, L.hA.n_cols, n_ranks, nullptr, 1);
            L.haloA_colour = plan_of(EA, L.hA.n_cols, n_ranks, L.colour.h_group.data(), L.colour.n_groups);
            if ((rc = upload_plan(L.haloA))) return rc;
            if ((rc = upload_plan(L.haloA_colour))) return rc;
            max_halo = std::max<size_t>(max_halo, std::max(L.haloA.n_send(), L.haloA.n_recv()));
        }
        max_blocks = std::max(max_blocks, ((size_t)nl * mgb::kLanes + 255) / 256 + 1);
    }
    return MGB_OK;
}

// common part of every constructor: handle, stream, communicator; `build` fills the levels
template <class Build>
int create_common(const mgb_amg_config *cfg, int rank, int n_ranks, const unsigned char nccl_id[128], mgb_amg_t *out, Build build)
{
    *out = nullptr;
    if (cfg->levels < 1 || cfg->levels > 16) return mgb_set_error(MGB_ERR_ARG, "1 <= levels <= 16");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return mgb_set_error(MGB_ERR_ARG, "bad rank / n_ranks");
    if (n_ranks > 1 && !nccl_id) return mgb_set_error(MGB_ERR_ARG, "n_ranks > 1 needs the ncclUniqueId of mgb_nccl_unique_id()");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return mgb_set_error(MGB_ERR_CUDA, "no CUDA device: libmgb200 has no CPU fallback");
    }
    ACK(cudaSetDevice(cfg->device));
    mgb_amg *h = new mgb_amg();
    struct Guard { mgb_amg *h; ~Guard() { if (h) mgb_amg_destroy(h); } } guard{h};     // any early return frees what exists so far
    h->cfg = *cfg;
    h->rank = rank; h->n_ranks = n_ranks;
    h->omega = cfg->jacobi_omega > 0. ? cfg->jacobi_omega : 1.0;
    ACK(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
    {
        int coop = 0, sms = 0, per_sm = 0;
        ACK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, cfg->device));
        ACK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device));
        ACK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mgb::k_amg_sell_gs_sweeps, 256, 0));
        h->coop_max_blocks = (coop && cfg->coop_sweeps > 0) ? sms * per_sm : 0;
    }
    if (n_ranks > 1) {
        auto &Nc = mgb::nccl();
        if (!Nc.load()) return mgb_set_error(MGB_ERR_NCCL, Nc.error);
        mgb::NcclUniqueId id;
        std::memcpy(&id, nccl_id, sizeof(id));
        ANK(Nc.CommInitRank(&h->comm, n_ranks, id, rank));
    }
    size_t max_blocks = 1, max_halo = 1;
    if (int rc = build(h, max_blocks, max_halo)) return rc;
    if (n_ranks > 1) {
        ACK(cudaMalloc(&h->d_send, sizeof(double) * max_halo));
        ACK(cudaMalloc(&h->d_recv, sizeof(double) * max_halo));
        const char *e = std::getenv("MGB_P2P");
        if (cfg->p2p && !(e && std::atoi(e) == 0)) {
            if (int rc = setup_p2p(h, max_halo)) return rc;
            if (!h->p2p.on && rank == 0 && std::getenv("MGB_VERBOSE"))
                std::fprintf(stderr, "[mgb] AMG ghost exchanges use NCCL send/recv (%s)\n", h->p2p.why.c_str());
        }
    }
    ACK(cudaMalloc(&h->d_partial, sizeof(double) * max_blocks));
    ACK(cudaMalloc(&h->d_scal, sizeof(double) * 4));
    ACK(cudaMallocHost(&h->h_scal, sizeof(double) * 4));
    guard.h = nullptr;
    *out = h;
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
    c->hybrid_gs = 0;
    c->p2p = 1;                     // ghost exchanges by NVLink peer stores where the pools can be mapped (else NCCL)
    c->shard_min_rows = 131072;     // an exchange costs ~15 us by peer stores (~60 us by NCCL: then 262144 pays better)
    c->jacobi_omega = 1.0;                                                 // the reference's smoothers are unweighted
    c->tail_max_rows = 4000;
}

void mgb_amg_config_fast(mgb_amg_config *c)
{
    mgb_amg_config_default(c);
    c->smoother = MGB_SMOOTH_GS_RB;      // multicolour Gauss-Seidel
    c->exact_order = 0;
}

void mgb_amg_config_device(mgb_amg_config *c)
{
    mgb_amg_config_fast(c);
    c->device_setup = 1;
    c->coarse_smoother = MGB_SMOOTH_L1_JACOBI;
}

int mgb_amg_n_levels(mgb_amg_t h) { return h ? (int)h->lv.size() : 0; }

int mgb_amg_create_from_csr(const mgb_amg_config *cfg, size_t n, const int64_t *ptr, const int64_t *col,
                            const double *val, const double *rhs, mgb_amg_t *out)
{
    return mgb_amg_create_sharded(cfg, n, ptr, col, val, rhs, 0, 1, nullptr, out);
}

int mgb_amg_partition(size_t n, int n_ranks, int rank, size_t *row0, size_t *rows)
{
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks || n > (size_t)INT32_MAX || !row0 || !rows) return mgb_set_error(MGB_ERR_ARG, "bad partition arguments");
    const Block b = block_of((int)n, n_ranks, rank);
    *row0 = (size_t)b.r0; *rows = (size_t)b.size();
    return MGB_OK;
}

int mgb_amg_create_sharded(const mgb_amg_config *cfg, size_t n, const int64_t *ptr, const int64_t *col,
                           const double *val, const double *rhs, int rank, int n_ranks,
                           const unsigned char nccl_id[128], mgb_amg_t *out)
{
    if (!cfg || !ptr || !col || !val || !rhs || !out) return mgb_set_error(MGB_ERR_ARG, "null argument");
    if (n == 0 || n > (size_t)1 << 30 || ptr[n] > (int64_t)INT32_MAX) return mgb_set_error(MGB_ERR_ARG, "matrix too large for int32 indices");
    return create_common(cfg, rank, n_ranks, nccl_id, out, [&](mgb_amg *h, size_t &max_blocks, size_t &max_halo) {
        return cfg->device_setup ? build_levels_device_from_host(h, n, ptr, col, val, rhs, max_blocks, max_halo)
                                 : build_levels_host(h, n, ptr, col, val, rhs, max_blocks, max_halo);
    });
}

int mgb_system_device_view(mgb_system_t S, int *device, int *n, int *nnz, const int **ptr, const int **col, const double **val, const double **rhs);

// the level-0 system is already on the device (mgb_fem_assemble_p1 / mgb_fem_synthetic): nothing is staged through the host
int mgb_amg_create_from_system(const mgb_amg_config *cfg, mgb_system_t sys, int rank, int n_ranks,
                               const unsigned char nccl_id[128], mgb_amg_t *out)
{
    if (!cfg || !sys || !out) return mgb_set_error(MGB_ERR_ARG, "null argument");
    if (!cfg->device_setup) return mgb_set_error(MGB_ERR_ARG, "mgb_amg_create_from_system builds the hierarchy on the device: set device_setup = 1");
    int sdev = 0, n = 0, nnz = 0;
    const int *sp = nullptr, *sc = nullptr;
    const double *sv = nullptr, *sr = nullptr;
    if (int rc = mgb_system_device_view(sys, &sdev, &n, &nnz, &sp, &sc, &sv, &sr)) return rc;
    if (sdev != cfg->device) return mgb_set_error(MGB_ERR_ARG, "the system lives on another device than cfg->device");
    if (n <= 0) return mgb_set_error(MGB_ERR_ARG, "empty system");
    return create_common(cfg, rank, n_ranks, nccl_id, out, [&](mgb_amg *h, size_t &max_blocks, size_t &max_halo) -> int {
        h->lv.resize(h->cfg.levels);
        DevCsr &A = h->lv[0].A;
        A.n_rows = A.n_cols = n; A.nnz = nnz;
        ACK(cudaMalloc(&A.ptr, sizeof(int) * ((size_t)n + 1)));
        ACK(cudaMalloc(&A.col, sizeof(int) * (size_t)std::max(nnz, 1)));
        ACK(cudaMalloc(&A.val, sizeof(double) * (size_t)std::max(nnz, 1)));
        ACK(cudaMalloc(&h->lv[0].b, sizeof(double) * (size_t)n));
        ACK(cudaMemcpy(A.ptr, sp, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToDevice));
        ACK(cudaMemcpy(A.col, sc, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToDevice));
        ACK(cudaMemcpy(A.val, sv, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToDevice));
        ACK(cudaMemcpy(h->lv[0].b, sr, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice));
        return build_levels_device(h, max_blocks, max_halo);
    });
}

void mgb_amg_destroy(mgb_amg_t h)
{
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->st) cudaStreamSynchronize(h->st);
    for (auto &g : h->graphs) cudaGraphExecDestroy(g.exec);
    for (auto &L : h->lv) {
        L.A.release(); L.P.release(); L.R.release(); L.lex.release(); L.colour.release(); L.sell.release(); L.sellN.release(); L.sellR.release(); L.sellP.release();
        L.haloA.release(); L.haloA_colour.release(); L.haloR.release(); L.haloP.release();
        cudaFree(L.diag); cudaFree(L.x); cudaFree(L.b); cudaFree(L.tmp); cudaFree(L.dl1); cudaFree(L.b0); cudaFree(L.d_colour); cudaFree(L.d_cf);
    }
    cudaFree(h->d_partial); cudaFree(h->d_send); cudaFree(h->d_recv);
    {
        const bool mapped = h->p2p.on;
        h->p2p.close_peers();
        if (mapped && h->comm && h->d_scal) {     // no rank frees its pool while a peer still maps it
            mgb::nccl().AllReduce(h->d_scal + 3, h->d_scal + 3, 1, mgb::kNcclFloat64, mgb::kNcclSum, h->comm, h->st);
            cudaStreamSynchronize(h->st);
        }
        if (h->pool) cudaFree(h->pool);
    }
    cudaFree(h->d_scal);
    if (h->comm) mgb::nccl().CommDestroy(h->comm);
    if (h->h_scal) cudaFreeHost(h->h_scal);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
}

int mgb_amg_level_info(mgb_amg_t h, int level, size_t *n, size_t *nnz_a, size_t *nnz_p, size_t *n_coarse,
                       int *n_waves, int *n_colours)
{
    if (!h || level < 0 || level >= (int)h->lv.size()) return mgb_set_error(MGB_ERR_ARG, "bad level");
    const AmgLevel &L = h->lv[level];
    if (n) *n = (size_t)L.A.n_rows;
    if (nnz_a) *nnz_a = (size_t)L.A.nnz;
    if (nnz_p) *nnz_p = (size_t)L.P.nnz;
    if (n_coarse) *n_coarse = (size_t)L.P.n_cols;
    if (n_waves) *n_waves = L.lex.n_groups;
    if (n_colours) *n_colours = L.colour.n_groups;
    return MGB_OK;
}

int mgb_amg_level_rows(mgb_amg_t h, int level, size_t *row0, size_t *rows, int *sharded)
{
    if (!h || level < 0 || level >= (int)h->lv.size()) return mgb_set_error(MGB_ERR_ARG, "bad level");
    const AmgLevel &L = h->lv[level];
    if (row0) *row0 = (size_t)L.own.r0;
    if (rows) *rows = (size_t)L.own.size();
    if (sharded) *sharded = L.sharded ? 1 : 0;
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
    const AmgLevel &L = h->lv[level];
    if (h->cfg.device_setup) {                  // the hierarchy lives on the device: a host copy is made for the caller
        ACK(cudaSetDevice(h->cfg.device));
        HostCsr tmp;
        if (int rc = download(which == 0 ? L.A : L.P, tmp)) return rc;
        return copy_csr(tmp, ptr, col, val);
    }
    return copy_csr(which == 0 ? L.hA : L.hP, ptr, col, val);
}

int mgb_amg_get_schedule(mgb_amg_t h, int level, int which, int *group_of_row)
{
    if (!h || level < 0 || level >= (int)h->lv.size() || !group_of_row) return mgb_set_error(MGB_ERR_ARG, "bad level/pointer");
    const AmgLevel &L = h->lv[level];
    if (which == 2) {
        if (!L.d_cf) return mgb_set_error(MGB_ERR_STATE, "the C/F state is kept for device-built levels that were coarsened");
        ACK(cudaSetDevice(h->cfg.device));
        ACK(cudaMemcpy(group_of_row, L.d_cf, sizeof(int) * (size_t)L.A.n_rows, cudaMemcpyDeviceToHost));
        return MGB_OK;
    }
    const Schedule &S = which == 0 ? L.lex : L.colour;
    if (which == 1 && S.h_group.empty() && L.d_colour) {
        ACK(cudaSetDevice(h->cfg.device));
        ACK(cudaMemcpy(group_of_row, L.d_colour, sizeof(int) * (size_t)L.A.n_rows, cudaMemcpyDeviceToHost));
        return MGB_OK;
    }
    if (S.h_group.empty() && L.A.n_rows) return mgb_set_error(MGB_ERR_STATE, "this level holds no such schedule");
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
    if (which == 0) L.x_halo_ok = true;            // every rank passes the whole vector: ghosts included
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
    const int T = h->lt >= 0 ? h->lt : L - 1;          // the levels T .. L-1 are one launch when the tail is on
    int rc, i;
    for (i = 0; i < T; ++i) {
        if ((rc = do_smooth(h, i, kind_of(h, i), h->cfg.pre_sweeps))) return rc;
        if ((rc = do_restrict(h, i + 1))) return rc;
    }
    if (h->lt >= 0) {
        if ((rc = launch_tail(h, 0, h->cfg.pre_sweeps, h->cfg.coarse_sweeps, h->cfg.post_sweeps))) return rc;
    } else if ((rc = do_smooth(h, i, kind_of(h, i), h->cfg.coarse_sweeps))) return rc;
    for (i--; i >= 0; --i) {
        if ((rc = do_prolong(h, i))) return rc;
        if ((rc = do_smooth(h, i, kind_of(h, i), h->cfg.post_sweeps))) return rc;
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
    int rc, n = 0;
    double nrm = 0.;
    if ((rc = do_residual(h, 0, &nrm))) return rc;
    hist[n++] = nrm;
    const double target = tol * nrm;
    const int T = h->lt >= 0 ? h->lt : L - 1;                              // levels T .. L-1: one launch when the tail is on
    // one cycle: every launch, ghost exchange and the all-reduced norm of the new iterate, no host synchronisation
    auto cycle = [&]() -> int {
        int rc;
        for (int l = 0; l < T; ++l) {                                      // downward
            AmgLevel &F = h->lv[l], &C = h->lv[l + 1];
            if (nu1 > 0 && (rc = do_smooth(h, l, kind_of(h, l), nu1))) return rc;
            if ((rc = residual_to_tmp(h, F, false))) return rc;               // r = b - A x into F.tmp (no host read-back)
            if (F.R.n_rows) {                                                  // b_c = P^T r, x_c = 0
                if ((rc = restrict_vec(h, l + 1, F.tmp, C.b))) return rc;
                if (C.sell.n_slots) {
                    mgb::k_amg_to_slots<<<(C.sell.n_slots + 255) / 256, 256, 0, h->st>>>(C.sell.row_of_slot, C.sell.n_slots, C.b, C.sell.b_s);
                    tally(h, 16. * C.own.size());
                }
                ACK(cudaMemsetAsync(C.x, 0, sizeof(double) * (size_t)C.A.n_rows, h->st));
                C.x_halo_ok = true;
            }
        }
        if (h->lt >= 0) { if ((rc = launch_tail(h, 1, nu1, coarse, nu2))) return rc; }
        else if ((rc = do_smooth(h, L - 1, kind_of(h, L - 1), coarse))) return rc;
        for (int l = T - 1; l >= 0; --l) {                                  // upward
            if ((rc = do_prolong(h, l))) return rc;
            if (nu2 > 0 && (rc = do_smooth(h, l, kind_of(h, l), nu2))) return rc;
        }
        return residual_to_tmp(h, h->lv[0], true);
    };
    const bool graphed = h->cfg.cycle_graph >= 0;      // (Jacobi-type sweeps keep their pointers in place, see do_smooth)
    mgb_amg::CycleGraph *G = nullptr;
    if (graphed && maxit > 0 && nrm > target) {
        for (auto it = h->graphs.begin(); it != h->graphs.end();) {
            if (it->epoch != h->ptr_epoch) { cudaGraphExecDestroy(it->exec); it = h->graphs.erase(it); }
            else ++it;
        }
        for (auto &g : h->graphs) if (g.nu1 == nu1 && g.nu2 == nu2 && g.coarse == coarse) G = &g;
        if (!G) {
            const mgb_gmg_stats before = h->stats;
            cudaGraph_t graph = nullptr;
            ACK(cudaStreamBeginCapture(h->st, cudaStreamCaptureModeThreadLocal));
            rc = cycle();
            const cudaError_t ce = cudaStreamEndCapture(h->st, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            ACK(ce);
            mgb_amg::CycleGraph g{nu1, nu2, coarse, h->ptr_epoch, nullptr, h->stats.kernel_launches - before.kernel_launches,
                                  h->stats.bytes_algorithmic - before.bytes_algorithmic};
            ACK(cudaGraphInstantiate(&g.exec, graph, 0));
            cudaGraphDestroy(graph);
            h->stats = before;                                                 // nothing ran yet
            h->graphs.push_back(g);
            G = &h->graphs.back();
        }
    }
    for (int it = 0; it < maxit && nrm > target; ++it) {
        if (G) {
            ACK(cudaGraphLaunch(G->exec, h->st));
            h->stats.kernel_launches += G->launches; h->stats.bytes_algorithmic += G->bytes; h->stats.graph_launches++;
        } else if ((rc = cycle())) return rc;
        h->stats.cycles++;
        ACK(cudaMemcpyAsync(h->h_scal, h->d_scal, sizeof(double), cudaMemcpyDeviceToHost, h->st));
        ACK(cudaStreamSynchronize(h->st));
        nrm = std::sqrt(h->h_scal[0]);
        hist[n++] = nrm;
    }
    *n_hist = n;
    // the coarse right-hand sides were overwritten by restricted residuals: put the reference's P^T b back
    for (int l = 1; l < L; ++l) {
        AmgLevel &C = h->lv[l];
        if (!C.A.n_rows) continue;
        if (C.b0) ACK(cudaMemcpyAsync(C.b, C.b0, sizeof(double) * (size_t)C.A.n_rows, cudaMemcpyDeviceToDevice, h->st));
        else ACK(cudaMemcpyAsync(C.b, C.h_rhs.data(), sizeof(double) * (size_t)C.A.n_rows, cudaMemcpyHostToDevice, h->st));
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

// host-only view of the ghost-exchange plan (what mgb_amg_create_sharded builds for A, R and P of every sharded level)
int mgb_amg_halo_plan(mgb_csr_t M, int n_ranks, int rank, const int *group_of_col, int n_groups,
                      int64_t *send_ptr, int64_t *send_idx, int64_t *recv_ptr, int64_t *recv_idx)
{
    if (!M || n_ranks < 1 || rank < 0 || rank >= n_ranks || !send_ptr || !recv_ptr) return mgb_set_error(MGB_ERR_ARG, "bad halo-plan arguments");
    if (!group_of_col) n_groups = 1;
    if (n_groups < 1) return mgb_set_error(MGB_ERR_ARG, "n_groups < 1");
    if (group_of_col)
        for (int j = 0; j < M->m.n_cols; ++j)
            if (group_of_col[j] < 0 || group_of_col[j] >= n_groups) return mgb_set_error(MGB_ERR_ARG, "group_of_col out of range");
    const HaloPlan H = halo_plan(M->m, n_ranks, rank, group_of_col, n_groups);
    for (size_t q = 0; q < H.send_ptr.size(); ++q) { send_ptr[q] = H.send_ptr[q]; recv_ptr[q] = H.recv_ptr[q]; }
    if (send_idx) for (size_t t = 0; t < H.send_idx.size(); ++t) send_idx[t] = H.send_idx[t];
    if (recv_idx) for (size_t t = 0; t < H.recv_idx.size(); ++t) recv_idx[t] = H.recv_idx[t];
    return MGB_OK;
}

int mgb_amg_checksum(mgb_amg_t h, int level, int which, uint64_t *out)
{
    if (!h || level < 0 || level >= (int)h->lv.size() || !out || which < 0 || which > 2) return mgb_set_error(MGB_ERR_ARG, "bad level/pointer");
    ACK(cudaSetDevice(h->cfg.device));
    AmgLevel &L = h->lv[level];
    const double *v = which == 0 ? L.x : (which == 1 ? L.b : L.tmp);
    unsigned long long *d = reinterpret_cast<unsigned long long *>(h->d_scal + 2);
    ACK(cudaMemsetAsync(d, 0, sizeof(*d), h->st));
    // a sharded level: every rank adds its own rows; a replicated one: rank 0's copy counts
    Block rows = L.own;
    if (!L.sharded && h->n_ranks > 1 && h->rank != 0) rows = Block{0, 0};
    if (rows.size() > 0) mgb::k_amg_checksum<<<std::min((rows.size() + 255) / 256, 2368), 256, 0, h->st>>>(v, rows.r0, rows.r1, d);
    ACK(cudaGetLastError());
    if (h->n_ranks > 1) ANK(mgb::nccl().AllReduce(d, d, 1, mgb::kNcclUint64, mgb::kNcclSum, h->comm, h->st));
    unsigned long long hv = 0;
    ACK(cudaMemcpyAsync(&hv, d, sizeof(hv), cudaMemcpyDeviceToHost, h->st));
    ACK(cudaStreamSynchronize(h->st));
    *out = (uint64_t)hv;
    return MGB_OK;
}

int mgb_amg_uses_p2p(mgb_amg_t h) { return (h && h->p2p.on) ? 1 : 0; }

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
    if (h->p2p.on) {                  // a wait kernel that gave up (20 s without a peer's signal) leaves a mark instead of hanging
        unsigned int err = 0;
        ACK(cudaMemcpy(&err, &h->p2p.hdr(h->rank)->error, sizeof(err), cudaMemcpyDeviceToHost));
        if (err) return mgb_set_error(MGB_ERR_NCCL, "peer-store exchange timed out waiting for rank " + std::to_string((int)err - 1));
    }
    return MGB_OK;
}

}  // extern "C"
