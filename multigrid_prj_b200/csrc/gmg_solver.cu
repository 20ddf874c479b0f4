// gmg_solver.cu -- host side of the B200 geometric-multigrid solve phase + its C ABI.
//
// Replaces (reference file:line, relative to GeometricMultigrid/):
//   level hierarchy      L x SquareDomain + L x PoissonMatrix          src/main.cpp:32-41
//   cycle                SawtoothMGIteration::apply_iteration_to_vec   include/multigrid.hpp:126-145
//   coarse solve         Solver::Solve                                 include/solvers.hpp:324-342
//   driver loop          main                                          src/main.cpp:73-116
// The host only enqueues kernels on one stream and reads back one double where the reference
// inspects a norm.  No CPU arithmetic on grid data happens here.
#include "../../include/mgb200.h"
#include "gmg_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;
int fail(int code, const std::string &msg) { g_err = msg; return code; }

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(MGB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
    } while (0)

using mgb::LevelGeom;

struct Level {
    LevelGeom g{};
    size_t elems = 0;                 // (rows + 2) * pitch
    double *base[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    // pointers to local row 0 of: u, f (level 0 only), e, r, t (scratch for out-of-place sweeps)
    double *u = nullptr, *f = nullptr, *e = nullptr, *r = nullptr, *t = nullptr;
};

}  // namespace

struct mgb_gmg {
    mgb_gmg_config cfg{};
    std::vector<Level> lv;
    cudaStream_t st = nullptr;
    double *d_partial = nullptr;      // per-CTA partial sums
    size_t n_partial = 0;
    double *d_scal = nullptr;         // device scalars: [0] last sumsq
    double *h_scal = nullptr;         // pinned mirror
    double norm_f = 0.;               // sum f^2 on the fine grid (Residual ctor, solvers.hpp:237-242)
    bool have_rhs = false;
    int n_sm = 148;
    mgb_gmg_stats stats{};

    double **vec(int level, int which)
    {
        Level &L = lv[level];
        switch (which) {
        case MGB_VEC_U: return level == 0 ? &L.u : nullptr;
        case MGB_VEC_F: return level == 0 ? &L.f : nullptr;
        case MGB_VEC_E: return &L.e;
        case MGB_VEC_R: return &L.r;
        default: return nullptr;
        }
    }
};

namespace {

dim3 march_grid(const LevelGeom &g)
{
    return dim3((g.w + 2 * mgb::kTPB - 1) / (2 * mgb::kTPB), (g.rows + mgb::kRowsPerCta - 1) / mgb::kRowsPerCta);
}

inline void count(mgb_gmg *h, double bytes) { h->stats.kernel_launches++; h->stats.bytes_algorithmic += bytes; }
inline double npts(const LevelGeom &g) { return (double)g.w * (double)g.rows; }

int halo_exchange(mgb_gmg *h, int level, double *v)
{
    (void)level; (void)v;
    if (h->cfg.n_ranks > 1) return fail(MGB_ERR_STATE, "halo exchange not wired");
    return MGB_OK;
}

// reduce d_partial[0..n) into d_scal[slot]
int reduce_partials(mgb_gmg *h, int n, int slot)
{
    mgb::k_reduce_partials<<<1, 1024, 0, h->st>>>(h->d_partial, n, h->d_scal + slot);
    count(h, 0.);
    CK(cudaGetLastError());
    return MGB_OK;
}

int read_scalar(mgb_gmg *h, int slot, double *out)
{
    CK(cudaMemcpyAsync(h->h_scal + slot, h->d_scal + slot, sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    *out = h->h_scal[slot];
    return MGB_OK;
}

template <int S, bool EXACT>
int launch_rb_stream_t(mgb_gmg *h, const LevelGeom &g, const double *in, const double *rhs, double *out)
{
    static bool attr_set = false;
    constexpr int smem = mgb::stream_smem_bytes<S>();
    if (!attr_set) {
        CK(cudaFuncSetAttribute(mgb::k_rb_stream<S, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    const int OW = mgb::kStreamTW - 2 * S;
    const int nx = (g.w + OW - 1) / OW;
    const int occ = std::max(1, std::min(16, (227 * 1024) / (smem + 1024)));
    const int slots = h->n_sm * occ;
    int ny = std::max(1, (slots + nx / 2) / nx);
    ny = std::min(ny, std::max(1, g.rows / (8 * S)));
    int rc = (g.rows + ny - 1) / ny;
    rc += rc & 1;                                  // even chunks keep (row0 + first streamed row) even
    ny = (g.rows + rc - 1) / rc;
    mgb::k_rb_stream<S, EXACT><<<dim3(nx, ny), mgb::kStreamNT, smem, h->st>>>(g, in, rhs, out, rc);
    count(h, 24. * (S / 2) * npts(g));     // 24 B per point per sweep, S/2 sweeps per launch
    CK(cudaGetLastError());
    return MGB_OK;
}

// `sweeps` (1, 2 or 5) full red-black sweeps in one pass: in -> out
int launch_rb_stream(mgb_gmg *h, int level, int sweeps, const double *in, const double *rhs, double *out)
{
    const LevelGeom &g = h->lv[level].g;
    const bool ex = !h->cfg.rb_fast_arith;
    switch (sweeps) {
    case 1: return ex ? launch_rb_stream_t<2, true>(h, g, in, rhs, out) : launch_rb_stream_t<2, false>(h, g, in, rhs, out);
    case 2: return ex ? launch_rb_stream_t<4, true>(h, g, in, rhs, out) : launch_rb_stream_t<4, false>(h, g, in, rhs, out);
    case 5: return ex ? launch_rb_stream_t<10, true>(h, g, in, rhs, out) : launch_rb_stream_t<10, false>(h, g, in, rhs, out);
    default: return fail(MGB_ERR_ARG, "unsupported sweep group");
    }
}

int do_smooth(mgb_gmg *h, int level, int kind, int sweeps, double **sol, const double *rhs)
{
    Level &L = h->lv[level];
    const LevelGeom &g = L.g;
    if (kind == MGB_SMOOTH_BICGSTAB) kind = MGB_SMOOTH_JACOBI;      // main.cpp:103-106
    dim3 grid = march_grid(g);
    for (int s = 0; s < sweeps; ++s) {
        if (kind == MGB_SMOOTH_JACOBI) {
            mgb::k_jacobi<<<grid, mgb::kTPB, 0, h->st>>>(g, *sol, rhs, L.t);
            count(h, 24. * npts(g));
            std::swap(*sol, L.t);                                   // solvers.hpp:83 sol.swap(temp)
            if (int rc = halo_exchange(h, level, *sol)) return rc;
        } else if (kind == MGB_SMOOTH_GS_RB && h->cfg.rb_fused) {
            // group the remaining sweeps: 5, 2 or 1 full sweeps per pass over HBM
            const int left = sweeps - s;
            const int grp = left >= 5 ? 5 : (left >= 2 ? 2 : 1);
            if (int rc = launch_rb_stream(h, level, grp, *sol, rhs, L.t)) return rc;
            std::swap(*sol, L.t);
            s += grp - 1;
            if (int rc = halo_exchange(h, level, *sol)) return rc;
        } else if (kind == MGB_SMOOTH_GS_RB) {
            for (int colour = 0; colour < 2; ++colour) {
                mgb::k_rbgs_colour<<<grid, mgb::kTPB, 0, h->st>>>(g, *sol, rhs, colour);
                count(h, 12. * npts(g));
                if (int rc = halo_exchange(h, level, *sol)) return rc;
            }
        } else if (kind == MGB_SMOOTH_GS_LEX) {
            if (h->cfg.n_ranks > 1)
                return fail(MGB_ERR_ARG, "lexicographic GS is sequential across slabs; use one rank for parity mode");
            for (int b0 = 0; b0 < g.rows; b0 += 1024) {
                int nb = std::min(1024, g.rows - b0);
                mgb::k_gs_lex_band<<<1, 1024, 0, h->st>>>(g, *sol, rhs, b0, nb);
                count(h, 24. * (double)g.w * nb);
            }
        } else
            return fail(MGB_ERR_ARG, "unknown smoother kind");
    }
    CK(cudaGetLastError());
    return MGB_OK;
}

// r = rhs - A sol; leaves sum r^2 (this rank) in d_scal[slot]
int do_residual(mgb_gmg *h, int level, const double *sol, const double *rhs, double *store, int slot)
{
    const LevelGeom &g = h->lv[level].g;
    dim3 grid = march_grid(g);
    if (store) mgb::k_residual<true><<<grid, mgb::kTPB, 0, h->st>>>(g, sol, rhs, store, h->d_partial);
    else mgb::k_residual<false><<<grid, mgb::kTPB, 0, h->st>>>(g, sol, rhs, nullptr, h->d_partial);
    count(h, (store ? 24. : 16.) * npts(g));
    CK(cudaGetLastError());
    return reduce_partials(h, grid.x * grid.y, slot);
}

int do_sumsq(mgb_gmg *h, int level, const double *v, int slot)
{
    const LevelGeom &g = h->lv[level].g;
    dim3 grid = march_grid(g);
    mgb::k_sumsq<<<grid, mgb::kTPB, 0, h->st>>>(g, v, h->d_partial);
    count(h, 8. * npts(g));
    CK(cudaGetLastError());
    return reduce_partials(h, grid.x * grid.y, slot);
}

int do_restrict(mgb_gmg *h)
{
    const int L = (int)h->lv.size();
    for (int l = 1; l < L; ++l) {
        Level &F = h->lv[l - 1], &C = h->lv[l];
        dim3 grid((C.g.w + 255) / 256, C.g.rows);
        if (h->cfg.restriction == MGB_RESTRICT_FULL_WEIGHTING) {
            if (int rc = halo_exchange(h, l - 1, F.r)) return rc;
            mgb::k_restrict<2><<<grid, 256, 0, h->st>>>(F.g, C.g, F.r, C.r, 1.0);
            count(h, 8. * npts(F.g) + 8. * npts(C.g));
        } else {
            double scale = (h->cfg.restriction == MGB_RESTRICT_HALF_INJECTION && l == 1) ? 0.5 : 1.0;
            mgb::k_restrict<0><<<grid, 256, 0, h->st>>>(F.g, C.g, F.r, C.r, scale);
            count(h, 16. * npts(C.g));
        }
    }
    CK(cudaGetLastError());
    return MGB_OK;
}

int do_prolong(mgb_gmg *h, int lc)
{
    Level &C = h->lv[lc], &F = h->lv[lc - 1];
    if (int rc = halo_exchange(h, lc, C.e)) return rc;
    dim3 grid((F.g.w + 2 * mgb::kTPB - 1) / (2 * mgb::kTPB), (F.g.rows + 3) / 4);
    mgb::k_prolong<<<grid, mgb::kTPB, 0, h->st>>>(C.g, F.g, C.e, F.e);
    count(h, 8. * (npts(F.g) + npts(C.g)));
    CK(cudaGetLastError());
    return halo_exchange(h, lc - 1, F.e);
}

// multigrid.hpp:126-145
int do_cycle(mgb_gmg *h, double *coarse_relres, int *coarse_iters)
{
    const int L = (int)h->lv.size();
    Level &F = h->lv[0], &C = h->lv[L - 1];
    const int kind = h->cfg.smoother;
    int rc;
    // :127 sol * RES  -> r0 = f - A u on the fine grid (the norm of this residual is never read)
    if ((rc = do_residual(h, 0, F.u, F.f, F.r, 1))) return rc;
    if ((rc = do_restrict(h))) return rc;
    // :128 COARSE_RES->refresh_normalization_constant()
    if ((rc = do_sumsq(h, L - 1, C.r, 2))) return rc;
    double nb = 0., norm = 0.;
    if ((rc = read_scalar(h, 2, &nb))) return rc;
    // :130 err * COARSE_SOLVER  (err == 0 on entry, multigrid.hpp:143)
    CK(cudaMemsetAsync(C.e - C.g.pitch, 0, C.elems * sizeof(double), h->st));
    int its = 0;
    if ((rc = do_residual(h, L - 1, C.e, C.r, nullptr, 3))) return rc;
    if ((rc = read_scalar(h, 3, &norm))) return rc;
    while (std::sqrt(norm / nb) > h->cfg.coarse_tol && its < h->cfg.coarse_maxit) {
        if ((rc = do_smooth(h, L - 1, kind, 1, &C.e, C.r))) return rc;
        ++its;
        if ((rc = do_residual(h, L - 1, C.e, C.r, nullptr, 3))) return rc;
        if ((rc = read_scalar(h, 3, &norm))) return rc;
    }
    if (coarse_relres) *coarse_relres = std::sqrt(norm / nb);       // :131 (the printed value)
    if (coarse_iters) *coarse_iters = its;
    h->stats.coarse_iters_total += its;
    // :134-139
    for (int j = L - 1; j > 0; --j) {
        if ((rc = do_prolong(h, j))) return rc;
        if ((rc = do_smooth(h, j - 1, kind, h->cfg.nu, &h->lv[j - 1].e, h->lv[j - 1].r))) return rc;
    }
    // :141-144 (err is fully rewritten next cycle, so the err = 0 store is not needed)
    dim3 grid((F.g.pitch / 2 + 255) / 256, std::min(F.g.rows, 1024));
    mgb::k_axpy_rows<<<grid, 256, 0, h->st>>>(F.g, F.u, F.e);
    count(h, 24. * npts(F.g));
    CK(cudaGetLastError());
    if ((rc = halo_exchange(h, 0, F.u))) return rc;
    h->stats.cycles++;
    return MGB_OK;
}

int copy_2d(mgb_gmg *h, const LevelGeom &g, double *dev, const double *host_global, bool to_device)
{
    // host arrays are global w x w row-major; this rank moves its slab rows
    const double *hsrc = host_global + (size_t)g.row0 * g.w;
    if (to_device)
        CK(cudaMemcpy2DAsync(dev, (size_t)g.pitch * 8, hsrc, (size_t)g.w * 8, (size_t)g.w * 8, g.rows,
                             cudaMemcpyHostToDevice, h->st));
    else
        CK(cudaMemcpy2DAsync(const_cast<double *>(hsrc), (size_t)g.w * 8, dev, (size_t)g.pitch * 8,
                             (size_t)g.w * 8, g.rows, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return MGB_OK;
}

}  // namespace

extern "C" {

const char *mgb_last_error(void) { return g_err.c_str(); }
const char *mgb_version(void) { return "mgb200 0.1 (sm_100a)"; }

int mgb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void mgb_gmg_config_default(mgb_gmg_config *c)
{
    std::memset(c, 0, sizeof(*c));
    c->n = 200; c->alpha = 10.0; c->length = 10.0; c->levels = 2;        // utilities.hpp:16-21
    c->smoother = MGB_SMOOTH_GS_LEX;
    c->pre_smoother = MGB_SMOOTH_GS_LEX; c->n_pre = 2;                  // main.cpp:62,85
    c->nu = 5; c->coarse_tol = 1.e-1; c->coarse_maxit = 2000;           // multigrid.hpp:105,123
    c->restriction = MGB_RESTRICT_INJECTION;
    c->device = 0; c->rank = 0; c->n_ranks = 1;
    c->tail_max_width = 0; c->use_graph = 0;
    c->rb_fast_arith = 0; c->rb_fused = 1;
}

void mgb_gmg_config_fast(mgb_gmg_config *c)
{
    mgb_gmg_config_default(c);
    c->smoother = MGB_SMOOTH_GS_RB;
    c->pre_smoother = MGB_SMOOTH_GS_RB;
    c->restriction = MGB_RESTRICT_FULL_WEIGHTING;
    c->rb_fast_arith = 1;
}

int mgb_gmg_create(const mgb_gmg_config *cfg, mgb_gmg_t *out)
{
    if (!cfg || !out) return fail(MGB_ERR_ARG, "null argument");
    *out = nullptr;
    const size_t N = cfg->n;
    const int L = cfg->levels;
    if (N < 3 || L < 1 || L > 30) return fail(MGB_ERR_ARG, "need n >= 3 and 1 <= levels <= 30");
    if (N > (size_t)1 << 20) return fail(MGB_ERR_ARG, "n too large");
    // the reference silently requires this: otherwise coarse "boundary" nodes are
    // not boundary nodes and neighbour indices run out of range
    if ((N - 1) % ((size_t)1 << (L - 1)) != 0 || ((N - 1) >> (L - 1)) < 1)
        return fail(MGB_ERR_ARG, "(n-1) must be divisible by 2^(levels-1)");
    if (cfg->n_ranks != 1 || cfg->rank != 0) return fail(MGB_ERR_ARG, "multi-rank slabs not wired in this build");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(MGB_ERR_CUDA, "no CUDA device: libmgb200 has no CPU fallback");
    }
    CK(cudaSetDevice(cfg->device));
    mgb_gmg *h = new mgb_gmg();
    h->cfg = *cfg;
    CK(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
    CK(cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, cfg->device));
    h->lv.resize(L);
    size_t w = N;
    const double m_h = cfg->length / (double)(N - 1);                     // domain.cpp:5
    for (int l = 0; l < L; ++l) {
        Level &lv = h->lv[l];
        const double hl = m_h * (double)((size_t)1 << l);                 // domain.hpp:92
        const double k = hl * hl;                                         // linear_system.hpp:17
        lv.g.w = (int)w; lv.g.rows = (int)w; lv.g.row0 = 0;
        lv.g.pitch = (int)((w + 2 + 15) / 16 * 16);
        lv.g.diag = 4. * cfg->alpha / k;                                  // linear_system.hpp:28
        lv.g.off = -cfg->alpha / k;                                       // linear_system.hpp:38
        lv.elems = (size_t)(lv.g.rows + 2) * lv.g.pitch;
        const int nvec = 5;
        for (int v = 0; v < nvec; ++v) {
            if (l > 0 && v < 2) continue;                                 // u, f exist on level 0 only
            CK(cudaMalloc(&lv.base[v], lv.elems * sizeof(double)));
            CK(cudaMemsetAsync(lv.base[v], 0, lv.elems * sizeof(double), h->st));
        }
        lv.u = lv.base[0] ? lv.base[0] + lv.g.pitch : nullptr;
        lv.f = lv.base[1] ? lv.base[1] + lv.g.pitch : nullptr;
        lv.e = lv.base[2] + lv.g.pitch;
        lv.r = lv.base[3] + lv.g.pitch;
        lv.t = lv.base[4] + lv.g.pitch;
        w = (w + 1) / 2;                                                  // domain.cpp:10
    }
    dim3 g0 = march_grid(h->lv[0].g);
    h->n_partial = (size_t)g0.x * g0.y;
    CK(cudaMalloc(&h->d_partial, h->n_partial * sizeof(double)));
    CK(cudaMalloc(&h->d_scal, 16 * sizeof(double)));
    CK(cudaMemsetAsync(h->d_scal, 0, 16 * sizeof(double), h->st));
    CK(cudaMallocHost(&h->h_scal, 16 * sizeof(double)));
    CK(cudaStreamSynchronize(h->st));
    *out = h;
    return MGB_OK;
}

void mgb_gmg_destroy(mgb_gmg_t h)
{
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->st) cudaStreamSynchronize(h->st);
    for (auto &lv : h->lv)
        for (double *p : lv.base) if (p) cudaFree(p);
    if (h->d_partial) cudaFree(h->d_partial);
    if (h->d_scal) cudaFree(h->d_scal);
    if (h->h_scal) cudaFreeHost(h->h_scal);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
}

int mgb_gmg_level_width(mgb_gmg_t h, int level, size_t *width)
{
    if (!h || level < 0 || level >= (int)h->lv.size() || !width) return fail(MGB_ERR_ARG, "bad level");
    *width = (size_t)h->lv[level].g.w;
    return MGB_OK;
}

int mgb_gmg_level_rows(mgb_gmg_t h, int level, size_t *row0, size_t *rows)
{
    if (!h || level < 0 || level >= (int)h->lv.size()) return fail(MGB_ERR_ARG, "bad level");
    if (row0) *row0 = (size_t)h->lv[level].g.row0;
    if (rows) *rows = (size_t)h->lv[level].g.rows;
    return MGB_OK;
}

int mgb_gmg_set_level(mgb_gmg_t h, int level, int which, const double *host)
{
    if (!h || level < 0 || level >= (int)h->lv.size() || !host) return fail(MGB_ERR_ARG, "bad level/pointer");
    double **p = h->vec(level, which);
    if (!p) return fail(MGB_ERR_ARG, "vector does not exist on this level");
    CK(cudaSetDevice(h->cfg.device));
    int rc = copy_2d(h, h->lv[level].g, *p, host, true);
    if (rc) return rc;
    if (level == 0 && which == MGB_VEC_F) {
        if ((rc = do_sumsq(h, 0, h->lv[0].f, 0))) return rc;
        if ((rc = read_scalar(h, 0, &h->norm_f))) return rc;
        h->have_rhs = true;
    }
    return halo_exchange(h, level, *p);
}

int mgb_gmg_get_level(mgb_gmg_t h, int level, int which, double *host)
{
    if (!h || level < 0 || level >= (int)h->lv.size() || !host) return fail(MGB_ERR_ARG, "bad level/pointer");
    double **p = h->vec(level, which);
    if (!p) return fail(MGB_ERR_ARG, "vector does not exist on this level");
    CK(cudaSetDevice(h->cfg.device));
    return copy_2d(h, h->lv[level].g, *p, host, false);
}

int mgb_gmg_set_rhs(mgb_gmg_t h, const double *b_host) { return mgb_gmg_set_level(h, 0, MGB_VEC_F, b_host); }

int mgb_gmg_set_rhs_test(mgb_gmg_t h, int test)
{
    if (!h) return fail(MGB_ERR_ARG, "null handle");
    CK(cudaSetDevice(h->cfg.device));
    if (test < 0 || test > 2) test = 0;                                   // utilities.cpp:150-154
    const LevelGeom &g = h->lv[0].g;
    dim3 grid((g.w + 255) / 256, g.rows);
    mgb::k_sample_rhs<<<grid, 256, 0, h->st>>>(g, h->lv[0].f, h->cfg.length, test);
    count(h, 8. * npts(g));
    CK(cudaGetLastError());
    int rc;
    if ((rc = do_sumsq(h, 0, h->lv[0].f, 0))) return rc;
    if ((rc = read_scalar(h, 0, &h->norm_f))) return rc;
    h->have_rhs = true;
    return MGB_OK;
}

int mgb_gmg_set_u(mgb_gmg_t h, const double *u_host)
{
    if (!h) return fail(MGB_ERR_ARG, "null handle");
    if (u_host) return mgb_gmg_set_level(h, 0, MGB_VEC_U, u_host);
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaMemsetAsync(h->lv[0].u - h->lv[0].g.pitch, 0, h->lv[0].elems * sizeof(double), h->st));
    return MGB_OK;
}

int mgb_gmg_get_u(mgb_gmg_t h, double *u_host) { return mgb_gmg_get_level(h, 0, MGB_VEC_U, u_host); }

int mgb_gmg_smooth(mgb_gmg_t h, int level, int kind, int sweeps, int sol, int rhs)
{
    if (!h || level < 0 || level >= (int)h->lv.size() || sweeps < 0) return fail(MGB_ERR_ARG, "bad level");
    double **s = h->vec(level, sol), **r = h->vec(level, rhs);
    if (!s || !r || s == r) return fail(MGB_ERR_ARG, "bad vector selector");
    CK(cudaSetDevice(h->cfg.device));
    return do_smooth(h, level, kind, sweeps, s, *r);
}

int mgb_gmg_residual(mgb_gmg_t h, int level, int sol, int rhs, int store, double *sumsq)
{
    if (!h || level < 0 || level >= (int)h->lv.size()) return fail(MGB_ERR_ARG, "bad level");
    double **s = h->vec(level, sol), **r = h->vec(level, rhs);
    if (!s || !r) return fail(MGB_ERR_ARG, "bad vector selector");
    if (store && rhs == MGB_VEC_R) return fail(MGB_ERR_ARG, "cannot store the residual over its own rhs");
    CK(cudaSetDevice(h->cfg.device));
    int rc = do_residual(h, level, *s, *r, store ? h->lv[level].r : nullptr, 1);
    if (rc) return rc;
    if (sumsq) return read_scalar(h, 1, sumsq);
    return MGB_OK;
}

int mgb_gmg_sumsq(mgb_gmg_t h, int level, int which, double *sumsq)
{
    if (!h || level < 0 || level >= (int)h->lv.size() || !sumsq) return fail(MGB_ERR_ARG, "bad level");
    double **v = h->vec(level, which);
    if (!v) return fail(MGB_ERR_ARG, "bad vector selector");
    CK(cudaSetDevice(h->cfg.device));
    int rc = do_sumsq(h, level, *v, 1);
    if (rc) return rc;
    return read_scalar(h, 1, sumsq);
}

int mgb_gmg_restrict(mgb_gmg_t h)
{
    if (!h) return fail(MGB_ERR_ARG, "null handle");
    CK(cudaSetDevice(h->cfg.device));
    return do_restrict(h);
}

int mgb_gmg_prolong(mgb_gmg_t h, int level_coarse)
{
    if (!h || level_coarse < 1 || level_coarse >= (int)h->lv.size()) return fail(MGB_ERR_ARG, "bad level");
    CK(cudaSetDevice(h->cfg.device));
    return do_prolong(h, level_coarse);
}

int mgb_gmg_cycle(mgb_gmg_t h, double *coarse_relres, int *coarse_iters)
{
    if (!h) return fail(MGB_ERR_ARG, "null handle");
    if (!h->have_rhs) return fail(MGB_ERR_STATE, "set the right-hand side first");
    CK(cudaSetDevice(h->cfg.device));
    return do_cycle(h, coarse_relres, coarse_iters);
}

int mgb_gmg_solve(mgb_gmg_t h, double tol, int maxiter, int check_every, double *hist, int *n_hist)
{
    if (!h || !hist || !n_hist || maxiter < 0) return fail(MGB_ERR_ARG, "bad argument");
    if (!h->have_rhs) return fail(MGB_ERR_STATE, "set the right-hand side first");
    CK(cudaSetDevice(h->cfg.device));
    if (check_every < 1) check_every = 1;
    Level &F = h->lv[0];
    int rc, n = 0;
    double ss = 0.;
    // main.cpp:73-74
    if ((rc = do_residual(h, 0, F.u, F.f, nullptr, 1))) return rc;
    if ((rc = read_scalar(h, 1, &ss))) return rc;
    hist[n++] = std::sqrt(ss / h->norm_f);
    for (int i = 0; i < maxiter; ++i) {                                   // main.cpp:84-90
        if ((rc = do_smooth(h, 0, h->cfg.pre_smoother, h->cfg.n_pre, &F.u, F.f))) return rc;
        if ((rc = do_cycle(h, nullptr, nullptr))) return rc;
        if ((rc = do_residual(h, 0, F.u, F.f, nullptr, 1))) return rc;
        if ((i + 1) % check_every == 0 || i + 1 == maxiter) {
            if ((rc = read_scalar(h, 1, &ss))) return rc;
            hist[n++] = std::sqrt(ss / h->norm_f);
            if (hist[n - 1] <= tol) break;
        }
    }
    *n_hist = n;
    return MGB_OK;
}

int mgb_gmg_run_cycles(mgb_gmg_t h, int cycles, double *final_relres)
{
    if (!h || cycles < 0) return fail(MGB_ERR_ARG, "bad argument");
    if (!h->have_rhs) return fail(MGB_ERR_STATE, "set the right-hand side first");
    CK(cudaSetDevice(h->cfg.device));
    Level &F = h->lv[0];
    int rc;
    for (int i = 0; i < cycles; ++i) {
        if ((rc = do_smooth(h, 0, h->cfg.pre_smoother, h->cfg.n_pre, &F.u, F.f))) return rc;
        if ((rc = do_cycle(h, nullptr, nullptr))) return rc;
        if ((rc = do_residual(h, 0, F.u, F.f, nullptr, 1))) return rc;
    }
    if (final_relres) {
        double ss = 0.;
        if (cycles == 0 && (rc = do_residual(h, 0, F.u, F.f, nullptr, 1))) return rc;
        if ((rc = read_scalar(h, 1, &ss))) return rc;
        *final_relres = std::sqrt(ss / h->norm_f);
    }
    return MGB_OK;
}

int mgb_gmg_get_stats(mgb_gmg_t h, mgb_gmg_stats *s)
{
    if (!h || !s) return fail(MGB_ERR_ARG, "null argument");
    *s = h->stats;
    return MGB_OK;
}
int mgb_gmg_reset_stats(mgb_gmg_t h)
{
    if (!h) return fail(MGB_ERR_ARG, "null handle");
    h->stats = mgb_gmg_stats{};
    return MGB_OK;
}
void *mgb_gmg_stream(mgb_gmg_t h) { return h ? (void *)h->st : nullptr; }
int mgb_gmg_sync(mgb_gmg_t h)
{
    if (!h) return fail(MGB_ERR_ARG, "null handle");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->st));
    return MGB_OK;
}

struct mgb_timer { cudaEvent_t a, b; };
int mgb_timer_create(mgb_timer_t *t)
{
    if (!t) return fail(MGB_ERR_ARG, "null argument");
    mgb_timer *x = new mgb_timer();
    CK(cudaEventCreate(&x->a));
    CK(cudaEventCreate(&x->b));
    *t = x;
    return MGB_OK;
}
void mgb_timer_destroy(mgb_timer_t t)
{
    if (!t) return;
    cudaEventDestroy(t->a); cudaEventDestroy(t->b);
    delete t;
}
int mgb_timer_start(mgb_timer_t t, void *stream) { CK(cudaEventRecord(t->a, (cudaStream_t)stream)); return MGB_OK; }
int mgb_timer_stop(mgb_timer_t t, void *stream) { CK(cudaEventRecord(t->b, (cudaStream_t)stream)); return MGB_OK; }
int mgb_timer_elapsed_ms(mgb_timer_t t, double *ms)
{
    float f = 0.f;
    CK(cudaEventSynchronize(t->b));
    CK(cudaEventElapsedTime(&f, t->a, t->b));
    *ms = (double)f;
    return MGB_OK;
}

int mgb_nccl_unique_id(unsigned char id[128])
{
    std::memset(id, 0, 128);
    return fail(MGB_ERR_NCCL, "NCCL not wired in this build");
}

}  // extern "C"
