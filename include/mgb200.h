/*
 * mgb200.h -- C ABI of the B200-native multigrid solve phase (libmgb200.so).
 *
 * This is the drop-in boundary for the reference's solver entry points
 * (Stefo01/multigrid_prj; citations relative to the reference tree).  The reference has no FFI:
 * its boundary is the C++ class API its two drivers use.  The facade headers
 * under multigrid_prj_b200/dropin/ re-declare those classes on top of the functions below, so the
 * reference's GeometricMultigrid/src/main.cpp and AMG/src/main.cpp compile against them unchanged.
 *
 * Conventions: plain pointers and sizes, opaque handles, int status (0 = ok, nonzero = error,
 * text from mgb_last_error()), no exceptions cross the boundary, one host thread per handle.
 * All host arrays are fp64, row-major.  There is NO CPU fallback: every compute entry point
 * fails with MGB_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef MGB200_H
#define MGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    MGB_OK = 0,
    MGB_ERR_ARG = 1,      /* bad argument (e.g. (N-1) not divisible by 2^(levels-1)) */
    MGB_ERR_CUDA = 2,     /* CUDA runtime error / no device */
    MGB_ERR_NCCL = 3,
    MGB_ERR_STATE = 4     /* call sequence error (e.g. solve before set_rhs) */
};

/* smoother ids: 0..2 follow GeometricMultigrid/include/utilities.hpp:9-14 (enum SMOOTHERS) */
enum {
    MGB_SMOOTH_GS_LEX = 0,   /* lexicographic Gauss-Seidel, solvers.hpp:24-49 (exact wavefront order)   */
    MGB_SMOOTH_JACOBI = 1,   /* Jacobi, omega = 1, solvers.hpp:53-84                                       */
    MGB_SMOOTH_BICGSTAB = 2, /* accepted and routed to Jacobi exactly as main.cpp:103-106 does            */
    MGB_SMOOTH_GS_RB = 3,    /* red-black Gauss-Seidel (reordered; the B200 fast path)                    */
    MGB_SMOOTH_L1_JACOBI = 4 /* AMG only: x += (b - A x) / (a_ii + sum_{j != i} |a_ij|): colouring-free, parameter-free,
                                one SELL launch per sweep (the smoother of the Galerkin levels on the fast path)  */
};

/* restriction of the fine residual to the coarse levels */
enum {
    MGB_RESTRICT_INJECTION = 0,      /* the reference's: res read at stride 2^l (solvers.hpp:35,46,69,80) */
    MGB_RESTRICT_HALF_INJECTION = 1, /* 0.5 * injection (pairs with red-black GS)                           */
    MGB_RESTRICT_FULL_WEIGHTING = 2  /* [1 2 1;2 4 2;1 2 1]/16 cascaded level by level (north_star)         */
};

/* which device-resident vector of a level an operator call refers to */
enum {
    MGB_VEC_U = 0,   /* fine solution u           (main.cpp:49)              level 0 only */
    MGB_VEC_F = 1,   /* fine right-hand side fvec (main.cpp:45)              level 0 only */
    MGB_VEC_E = 2,   /* err restricted to a level (multigrid.hpp:94)                       */
    MGB_VEC_R = 3    /* res restricted to a level (multigrid.hpp:93)                       */
};

typedef struct mgb_gmg_config {
    size_t n;             /* points per side of the fine grid            (-n,  utilities.hpp:16) */
    double length;        /* side of the square                          (-w)                    */
    double alpha;         /* diffusion constant                          (-a)                    */
    int levels;           /* multigrid levels L                          (-ml)                   */
    int smoother;         /* smoother inside the cycle                   (-smt)                  */
    int pre_smoother;     /* driver pre-smoother (main.cpp:62: always GS) */
    int n_pre;            /* driver pre-sweeps per cycle (main.cpp:85: 2) */
    int nu;               /* post-sweeps per level (multigrid.hpp:105: 5) */
    int restriction;      /* MGB_RESTRICT_*                                */
    double coarse_tol;    /* multigrid.hpp:123: 1e-1                       */
    int coarse_maxit;     /* multigrid.hpp:123: 2000                       */
    int device;           /* CUDA device ordinal                           */
    /* slab decomposition over one box (rank r owns a contiguous block of fine rows) */
    int rank, n_ranks;
    unsigned char nccl_id[128];   /* ncclUniqueId from mgb_nccl_unique_id(), same on all ranks */
    int tail_max_width;   /* levels with width <= this run inside the persistent coarse-tail kernel */
    int use_graph;        /* capture the static part of the cycle in a CUDA graph */
    int rb_fast_arith;    /* red-black kernels: 0 = the reference's formula, unfused IEEE ops and a true
                             division (bit-identical to the CPU statement of the same ordering);
                             1 = u = b/diag + (sum of neighbours)/4 with FMA (a few ulp away) */
    int rb_fused;         /* 1 = streaming temporally-blocked red-black kernel (all sweeps of a group in one
                             pass over HBM); 0 = one launch per colour */
    int fuse_correction;  /* red-black fused path only: the last post-smoothing launch on the fine level also applies
                             u += err (multigrid.hpp:141-144) and leaves sum (res - A err)^2 = ||f - A u_new||^2, the norm
                             main.cpp:86 asks for next, from data already on chip (no axpy pass, no residual pass).
                             The norm then carries the rounding of res - A err instead of f - A u (same value to ~1e-13 ||f||). */
    int fuse_residual;    /* red-black fused path only: the last driver pre-sweep launch also writes res = f - A u
                             (multigrid.hpp:127) for the rows it produces (no separate residual pass) */
    int fuse_prolong;     /* red-black fused path only: the first post-smoothing launch of a level interpolates its input
                             from the coarser level on the fly (multigrid.cpp:3-27, same arithmetic); the prolonged field
                             is never written to HBM */
    int defer_norm;       /* slabs (n_ranks > 1), mgb_gmg_run_cycles only: 0 (default) = the residual norm of every iteration is
                             all-reduced inside the loop, as the reference's loop reads it every iteration (main.cpp:86-90);
                             1 = every rank keeps its partial sum and the all-reduce happens once, when the norm is read */
    double jacobi_omega;  /* MGB_SMOOTH_JACOBI: u <- u + omega (u_jacobi - u) on interior points.  The reference is
                             omega = 1 (solvers.hpp:64-83) and that value keeps the bit-identical path; other values
                             (north_star: weighted Jacobi) run level by level without the persistent tail kernel.
                             <= 0 is read as 1 */
    /* --- textbook cycle options (NOT in the reference, whose only cycle is the sawtooth) --- */
    int cycle_type;       /* MGB_CYCLE_*: 0 = the reference's sawtooth (multigrid.hpp:126-145); V / W / F = correction-scheme cycles
                             with nu_pre pre-sweeps, the residual of the level's own equation restricted to the next level,
                             prolongation ADDED to the level's iterate and nu post-sweeps.  Levels inside the persistent coarse
                             tail (tail_max_width) are the coarse solver of these cycles: one sawtooth pass of that kernel */
    int nu_pre;           /* pre-sweeps per level of a V / W / F cycle (nu is the number of post-sweeps) */
    int fmg;              /* 1: mgb_gmg_solve starts with one full-multigrid pass (mgb_gmg_fmg) */
    int reserved_cycle;
} mgb_gmg_config;

enum { MGB_CYCLE_SAWTOOTH = 0, MGB_CYCLE_V = 1, MGB_CYCLE_W = 2, MGB_CYCLE_F = 3 };
/* Krylov methods of mgb_gmg_krylov and their preconditioners */
enum { MGB_KRYLOV_CG = 0, MGB_KRYLOV_BICGSTAB = 1 };
enum { MGB_PRECOND_NONE = 0, MGB_PRECOND_MG = 1 };

typedef struct mgb_gmg *mgb_gmg_t;

/* fills cfg with the reference's defaults (utilities.hpp:16-21, multigrid.hpp:105,123, main.cpp:85) */
void mgb_gmg_config_default(mgb_gmg_config *cfg);
/* same, but the B200 fast path: red-black GS everywhere + full weighting */
void mgb_gmg_config_fast(mgb_gmg_config *cfg);

/* replaces: L x SquareDomain (domain.cpp:4-13) + L x PoissonMatrix (linear_system.hpp:16-17)
 * + SawtoothMGIteration ctor (multigrid.hpp:108-124) + Residual/GS objects of main.cpp:58-62 */
int mgb_gmg_create(const mgb_gmg_config *cfg, mgb_gmg_t *out);
void mgb_gmg_destroy(mgb_gmg_t h);

/* Row-slab partition used for n_ranks > 1 (pure host arithmetic, no device needed): level `level` is
 * either sharded (rank owns rows [row0,row0+rows)) or replicated on every rank (row0 = 0, rows = width). */
int mgb_gmg_partition(size_t n, int levels, int n_ranks, int rank, int level, int *sharded, size_t *row0, size_t *rows);

/* Where rank `rank` keeps vector `which_buffer` (0 u, 1 f, 2 e, 3 r, 4 t, 5 tu: the ping-pong partners included) of `level`
 * inside its device pool: byte offset of the allocation that holds the vector with its halo rows (SIZE_MAX when the level
 * has no such vector) and the pool's total size.  Every rank evaluates this for every OTHER rank to address the peers'
 * halo rows in the peer-store exchanges (csrc/p2p.cuh), so it must be a pure function of its arguments: host-only. */
int mgb_gmg_pool_layout(size_t n, int levels, int n_ranks, int rank, int level, int which_buffer, size_t *offset, size_t *total);

/* geometry queries (domain.hpp:82,90,94) */
int mgb_gmg_level_width(mgb_gmg_t h, int level, size_t *width);
/* local slab of level `level`: first global row and number of rows owned by this rank */
int mgb_gmg_level_rows(mgb_gmg_t h, int level, size_t *row0, size_t *rows);

/* replaces DataVector (linear_system.hpp:85-92): b_host is the n*n row-major fvec (global array;
 * each rank copies its slab).  mgb_gmg_set_rhs_test samples f,g of utilities.cpp:138-147 on device. */
int mgb_gmg_set_rhs(mgb_gmg_t h, const double *b_host);
int mgb_gmg_set_rhs_test(mgb_gmg_t h, int test);
int mgb_gmg_set_u(mgb_gmg_t h, const double *u_host);          /* n*n global, NULL = zeros */
int mgb_gmg_get_u(mgb_gmg_t h, double *u_host);                /* n*n global; each rank writes its slab */

/* compact w_l x w_l level arrays, for the facade and the per-operator parity tests */
int mgb_gmg_set_level(mgb_gmg_t h, int level, int which, const double *host);
int mgb_gmg_get_level(mgb_gmg_t h, int level, int which, double *host);

/* replaces SmootherClass::apply_iteration_to_vec (solvers.hpp:33-48, 64-83): `sweeps` sweeps of
 * `kind` on `level`, solution vector `sol` (MGB_VEC_U or MGB_VEC_E), rhs `rhs` (MGB_VEC_F or MGB_VEC_R) */
int mgb_gmg_smooth(mgb_gmg_t h, int level, int kind, int sweeps, int sol, int rhs);
/* replaces Residual::apply_iteration_to_vec + Norm (solvers.hpp:257-307): r = rhs - A sol on `level`;
 * store != 0 writes r into MGB_VEC_R of that level; sumsq = sum r^2 (all ranks) */
int mgb_gmg_residual(mgb_gmg_t h, int level, int sol, int rhs, int store, double *sumsq);
/* sum of squares of a level vector (Residual ctor / refresh_normalization_constant, solvers.hpp:230-254) */
int mgb_gmg_sumsq(mgb_gmg_t h, int level, int which, double *sumsq);
/* restriction of MGB_VEC_R from level 0 to every coarser level with the configured operator */
int mgb_gmg_restrict(mgb_gmg_t h);
/* replaces InterpolationClass::interpolate (multigrid.cpp:3-27): E(level_coarse) -> E(level_coarse-1) */
int mgb_gmg_prolong(mgb_gmg_t h, int level_coarse);

/* the reference builds one SawtoothMGIteration object per smoother (main.cpp:54-56) over the same grids; the
 * handle keeps one hierarchy and switches the cycle's parameters instead */
int mgb_gmg_set_cycle(mgb_gmg_t h, int smoother, int restriction, int nu, double coarse_tol, int coarse_maxit);

/* switches the cycle shape of a live handle (mgb_gmg_config.cycle_type / nu_pre / fmg) */
int mgb_gmg_set_cycle_type(mgb_gmg_t h, int cycle_type, int nu_pre, int fmg);

/* switches mgb_gmg_config.defer_norm of a live handle (measurement: the same handle timed both ways) */
int mgb_gmg_set_defer_norm(mgb_gmg_t h, int defer);

/* generation of the streaming red-black kernel behind the fused launches: 2 (default) = bulk-copy fed kernel with statically
 * addressed rings (csrc/gmg_stream2.cuh) wherever it is instantiated, 1 = first-generation kernel everywhere.  Results are
 * bit-identical; the switch exists for A/B measurement and for the parity test of one against the other.
 * The environment variable MGB_STREAM_IMPL sets the default of new handles. */
int mgb_gmg_set_stream_impl(mgb_gmg_t h, int impl);

/* replaces SawtoothMGIteration::apply_iteration_to_vec (multigrid.hpp:126-145) applied to u.
 * coarse_relres = the value the reference prints per cycle; coarse_iters = coarse-solve sweeps. */
int mgb_gmg_cycle(mgb_gmg_t h, double *coarse_relres, int *coarse_iters);

/* replaces the driver loop main.cpp:73-116: hist[0] = ||f-Au||/||f||, then per iteration n_pre
 * pre-sweeps, one cycle, one residual norm; stops when hist <= tol or after maxiter cycles.
 * hist must hold maxiter+1 doubles. check_every: read the norm back (one double, one sync) every
 * k-th cycle only (1 = the reference's behaviour). */
int mgb_gmg_solve(mgb_gmg_t h, double tol, int maxiter, int check_every, double *hist, int *n_hist);
/* One full-multigrid pass on the residual equation (nested iteration): r = f - A u restricted to
 * every level, coarse solve, then per level upward: bilinear prolongation of the coarser correction as the initial guess
 * and one V(nu_pre, nu) cycle of that level's equation; finally u += e.  No extra storage: the V cycle of level l only
 * overwrites arrays of levels > l, which the pass has already left behind. */
int mgb_gmg_fmg(mgb_gmg_t h);

/* A real Krylov solver on the fine level, replacing the reference's never-executed BiCGSTAB (solvers.hpp:86-216; SURVEY.md
 * section 8 row a11 and 8f item 4): conjugate gradients or BiCGSTAB on A u = f, optionally right-preconditioned by ONE
 * multigrid cycle of the handle's configuration applied to the residual (zero initial guess).  The boundary rows of A are
 * identity rows: u is first set to f there, after which every residual and search direction vanishes on the boundary and
 * A acts as the symmetric interior operator.  CG needs a symmetric preconditioner: use MGB_SMOOTH_JACOBI cycles with full
 * weighting (or MGB_PRECOND_NONE); BiCGSTAB takes any cycle (red-black GS, sawtooth ...).
 * hist[0] = ||f - A u|| / ||f|| on entry, then one entry per iteration (maxit + 1 doubles); stops at hist <= tol. */
int mgb_gmg_krylov(mgb_gmg_t h, int method, int precond, double tol, int maxit, double *hist, int *n_hist);

/* ONE iteration of the driver loop (main.cpp:84-86: n_pre pre-sweeps, one cycle, the residual norm) as a single cached
 * CUDA-graph launch on the fast path -- what `u * GS * GS * MG; u * RES; RES.Norm()` amounts to.  *sumsq = sum (f - A u)^2
 * of the new iterate; on the fused path it is the norm the last launch leaves, and values whose relative size is below
 * confirm_below are re-evaluated by a true f - A u pass (see mgb_gmg_solve).  *coarse_relres = the value the reference
 * prints per cycle (multigrid.hpp:131).  The facade's lazy operator queue dispatches to this. */
int mgb_gmg_iterate(mgb_gmg_t h, double confirm_below, double *sumsq, double *coarse_relres);

/* runs exactly `cycles` driver iterations without any host readback; writes the final relative
 * residual.  This is the timed region of bench.py. */
int mgb_gmg_run_cycles(mgb_gmg_t h, int cycles, double *final_relres);

/* the fine-level part of the upward leg (multigrid.hpp:134-144 for j = 1): E(1) -> E(0) prolongation, nu sweeps on
 * level 0 against R(0), u += err, and sum (f - A u)^2 of the new u.  One fused launch on the fast path. */
int mgb_gmg_fine_leg(mgb_gmg_t h, double *sumsq);

/* 64-bit checksum of a level vector over ALL ranks (wrap-around sum of bit pattern x (2 * global index + 1) over the
 * points; integer addition is associative, so the value is independent of the slab partition and of the launch
 * geometry).  Equal checksums on 1 and N ranks <=> the slab-decomposed solve is bit-identical to the single-GPU one;
 * bench.py asserts exactly that.  Collective when n_ranks > 1. */
int mgb_gmg_checksum(mgb_gmg_t h, int level, int which, uint64_t *out);

/* 1 when the slab exchanges of this handle move by NVLink peer stores (CUDA IPC), 0 when they use NCCL send/recv (single
 * rank, MGB_P2P=0, or the pools could not be mapped) */
int mgb_gmg_uses_p2p(mgb_gmg_t h);

/* measurement hooks */
typedef struct mgb_gmg_stats {
    uint64_t kernel_launches;      /* kernels launched by this handle since create/reset */
    uint64_t graph_launches;
    uint64_t coarse_iters_total;
    uint64_t cycles;
    double bytes_algorithmic;      /* SURVEY.md section 8d accounting, summed over the launches */
    int reserved[8];               /* [0] = halo / gather exchanges posted */
} mgb_gmg_stats;
int mgb_gmg_get_stats(mgb_gmg_t h, mgb_gmg_stats *s);
int mgb_gmg_reset_stats(mgb_gmg_t h);
/* CUDA stream the handle launches on (cudaStream_t as void*), for event timing by the caller */
void *mgb_gmg_stream(mgb_gmg_t h);
int mgb_gmg_sync(mgb_gmg_t h);

/* =====================================================================================================
 * AMG on a CSR operator (reference: AMG/, citations relative to that directory)
 * ===================================================================================================== */
typedef struct mgb_amg_config {
    int levels;            /* number_of_levels (src/main.cpp:126: 5)                                   */
    double eps;            /* strength threshold EPSILON (include/AMG.hpp:21: 0.2)                     */
    int smoother;          /* MGB_SMOOTH_GS_LEX: lexicographic GS reproduced exactly by level scheduling
                              (include/Utilities.hpp:44-58); MGB_SMOOTH_GS_RB: multicolour GS from an
                              on-device greedy colouring (the fast path); MGB_SMOOTH_JACOBI                */
    int pre_sweeps;        /* src/AMG.cpp:287 (10) */
    int coarse_sweeps;     /* src/AMG.cpp:295 (200) */
    int post_sweeps;       /* src/AMG.cpp:302 (10) */
    int exact_order;       /* 1: one thread per row, terms in ascending column order, unfused IEEE ops (bit-identical
                              to the reference's loops); 0: sub-warp-per-row vector kernels with __shfl reductions */
    int device;
    int64_t start_index[16]; /* node the C/F splitting of level l starts from; the reference draws it from
                              std::random_device (src/Utilities.cpp:30-40); < 0 selects n/2 */
    /* --- beyond the reference --- */
    int hybrid_gs;         /* row-block sharded levels only.  0: ghost entries are refreshed after EVERY colour, the
                              multicolour sweep equals the single-GPU sweep bit for bit; 1: once per sweep ("hybrid"
                              Gauss-Seidel: Jacobi-like across block boundaries, one exchange instead of n_colours) */
    int shard_min_rows;    /* a level is cut into row blocks while it keeps at least this many rows per rank; smaller
                              levels are replicated on every rank (<= 0: 131072).  A ghost exchange costs ~15 us by peer stores
                              (p2p = 1) and 40-60 us by NCCL on 8 B200: sharding pays only where the sweep time saved exceeds that */
    double jacobi_omega;   /* MGB_SMOOTH_JACOBI: x <- x + omega (D^-1 (b - (A - D) x) - x); the reference is omega = 1
                              (<= 0 is read as 1) */
    int tail_max_rows;     /* north_star item 3: the trailing levels whose row count is <= this (and that are not sharded)
                              run inside ONE persistent single-CTA kernel per pass / cycle (mgb_amg_apply, mgb_amg_solve)
                              instead of one launch per colour and transfer; results are unchanged.
                              0: default (4000: measured optimum on B200, profiles/r01_amg_scale_1gpu_4M_tailsweep.json), < 0: off */
    int cycle_graph;       /* mgb_amg_solve: 0 (default) = one cycle (all launches, ghost exchanges and the norm) is captured
                              in a CUDA graph once and replayed -- the coarse levels are launch-latency bound; < 0: off.
                              (Jacobi swaps buffers every sweep and always runs uncaptured.) */
    int coop_sweeps;       /* > 0: on a level that needs no ghost exchange between colours, whole multicolour sweeps run as ONE
                              cooperative launch with a grid-wide barrier per colour.  Measured SLOWER than one launch per
                              colour on B200 (grid.sync of a full grid costs more than a launch gap: level-0 sweep 0.207 vs
                              0.151 ms, V(2,2) cycle 9.5 vs 5.8 ms at 4 M DoF, profiles/r01_amg_scale_1gpu_4M_coop.json), so
                              it is off by default (0) */
    int device_setup;      /* 0: the hierarchy is built on the host with the reference's semantics (AMG.hpp:105-369; exact C/F state
                              machine, bit-identical operators); 1: built ON THE DEVICE, parity-exempt: same strength measure,
                              same direct-interpolation weights and Galerkin operator, PMIS splitting instead of the reference's
                              sequential one, triple product by expand-sort-compress (csrc/amg_setup.cu).  Coarsening stops early
                              when a level has <= 1 coarse point or no fine point: mgb_amg_n_levels() gives the depth built.
                              Multicolour GS / Jacobi-type smoothers only. */
    int coarse_smoother;   /* smoother of the levels >= 1 in mgb_amg_apply / mgb_amg_solve: 0 = the same as `smoother`, else a
                              MGB_SMOOTH_* id (fast path: MGB_SMOOTH_L1_JACOBI -- the Galerkin levels need 13-16 colours) */
    int p2p;               /* row-block sharded runs: 1 (default) = ghost entries move by direct peer stores over NVLink into a
                              double-buffered staging area of the rank that needs them (pools exported through CUDA IPC, one push
                              kernel + one flag wait + one unpack per exchange, csrc/p2p.cuh) instead of pack / ncclSend /
                              ncclRecv / unpack; falls back to NCCL when the pools cannot be mapped.  0 = NCCL.
                              The environment variable MGB_P2P=0 forces NCCL for GMG slabs and AMG row blocks alike */
    int reserved[1];
} mgb_amg_config;

typedef struct mgb_amg *mgb_amg_t;

void mgb_amg_config_default(mgb_amg_config *cfg);   /* the reference's constants, exact lexicographic GS */
void mgb_amg_config_fast(mgb_amg_config *cfg);      /* multicolour GS + vector kernels, the reference's hierarchy */
void mgb_amg_config_device(mgb_amg_config *cfg);    /* fast + device_setup = 1 + l1-Jacobi on the levels >= 1 */
/* number of levels the handle holds (== cfg.levels unless the device setup stopped coarsening earlier) */
int mgb_amg_n_levels(mgb_amg_t h);

/* replaces Matrix/CSRMatrix + the AMG constructor + AMG::initialization() (include/AMG.hpp:33-41,
 * src/AMG.cpp:76-120): takes the level-0 operator as CSR (rows sorted by column, as Matrix's std::map
 * yields them; exact zeros are dropped as CSRMatrix::copy_from does) and the right-hand side, builds
 * the hierarchy (strength, C/F split, direct interpolation, Galerkin operators, restricted right-hand
 * sides) and uploads it.  x starts at zero on every level (src/main.cpp:125). */
int mgb_amg_create_from_csr(const mgb_amg_config *cfg, size_t n, const int64_t *row_ptr, const int64_t *col,
                            const double *val, const double *rhs, mgb_amg_t *out);
void mgb_amg_destroy(mgb_amg_t h);

/* ---- operator construction on the device ----
 * A linear system (CSR with int32 indices + right-hand side) that lives in HBM. */
typedef struct mgb_system *mgb_system_t;
/* replaces the assembly loop of AMG/src/main.cpp:34-117 with LinearFE::set_dofs (include/FEM.hpp:174-258) and the problem
 * functions of src/Utilities.cpp:3-28: P1 stiffness matrix and load vector of -div(grad u) = f, u = g on the boundary nodes,
 * unknowns = interior nodes in order of appearance (src/FEM.cpp:291-303).  tri holds 3 node indices per triangle, ascending
 * (src/FEM.cpp:162-170).  exact_order = 1: every quadrature term is added separately in the reference's order -> the matrix
 * is bit-identical to the reference's; 0: the three equal terms of an entry are summed first. */
int mgb_fem_assemble_p1(size_t n_nodes, const double *x, const double *y, const unsigned char *on_boundary, size_t n_tri,
                        const int64_t *tri, int exact_order, int device, mgb_system_t *out);
/* BASELINE config 5: side x side lattice on [0,2]^2, interior nodes jittered by <= 0.2 h, every cell split along a hashed
 * diagonal (counter-based hash of seed: the same mesh on every rank), generated AND assembled on the device. */
int mgb_fem_synthetic(size_t side, uint64_t seed, int device, mgb_system_t *out);
/* the mesh mgb_fem_synthetic uses, on the host (x, y, on_boundary: side*side entries; tri: 6*(side-1)^2) -- for checks */
int mgb_fem_synthetic_mesh(size_t side, uint64_t seed, double *x, double *y, unsigned char *on_boundary, int64_t *tri);
int mgb_system_info(mgb_system_t s, size_t *n, size_t *nnz);
int mgb_system_get(mgb_system_t s, int64_t *row_ptr, int64_t *col, double *val, double *rhs);   /* any pointer may be NULL */
void mgb_system_destroy(mgb_system_t s);

/* Row-block sharded AMG over the GPUs of one box (the reference is single-process).  One process
 * per GPU; EVERY rank passes the same level-0 system and builds the same hierarchy on its host (the setup is
 * deterministic), then keeps on its device only the rows [row0, row0+rows) of each sharded level's operators
 * (mgb_amg_partition).  Vectors keep global indexing; the ghost entries a rank's rows reference are refreshed by
 * grouped ncclSend/ncclRecv from precomputed index lists (mgb_amg_halo_plan), norms by ncclAllReduce, and levels below
 * cfg->shard_min_rows rows per rank are replicated (their restricted vector is all-gathered once per visit).
 * Lexicographic Gauss-Seidel is sequential across blocks: sharded levels take MGB_SMOOTH_GS_RB or MGB_SMOOTH_JACOBI.
 * nccl_id: from mgb_nccl_unique_id() on rank 0, same on all ranks (ignored when n_ranks == 1). */
int mgb_amg_create_sharded(const mgb_amg_config *cfg, size_t n, const int64_t *row_ptr, const int64_t *col,
                           const double *val, const double *rhs, int rank, int n_ranks,
                           const unsigned char nccl_id[128], mgb_amg_t *out);
/* same, from a system that already lives on the device (cfg->device_setup must be 1): the hierarchy is built where the
 * matrix is, nothing is staged through the host; every rank passes its own copy of the (identical) system */
int mgb_amg_create_from_system(const mgb_amg_config *cfg, mgb_system_t sys, int rank, int n_ranks,
                               const unsigned char nccl_id[128], mgb_amg_t *out);
/* the contiguous block of n rows (or vector entries) that `rank` owns.  Host-only, no device needed. */
int mgb_amg_partition(size_t n, int n_ranks, int rank, size_t *row0, size_t *rows);
/* rows of `level` this handle works on, and whether the level is sharded (0: replicated, whole on every rank) */
int mgb_amg_level_rows(mgb_amg_t h, int level, size_t *row0, size_t *rows, int *sharded);

/* hierarchy queries (the reference prints the level sizes, src/AMG.cpp:85-86,111) */
int mgb_amg_level_info(mgb_amg_t h, int level, size_t *n, size_t *nnz_a, size_t *nnz_p, size_t *n_coarse,
                       int *n_wavefronts, int *n_colours);
/* which: 0 = A_level, 1 = P_level (level+1 -> level); arrays sized from mgb_amg_level_info */
int mgb_amg_get_matrix(mgb_amg_t h, int level, int which, int64_t *row_ptr, int64_t *col, double *val);
/* which: 0 = wavefront of each row in the level schedule of lexicographic GS, 1 = colour of each row,
 * 2 = C/F state of each row (1 coarse, 0 fine; device-built hierarchies) */
int mgb_amg_get_schedule(mgb_amg_t h, int level, int which, int *group_of_row);
/* which: 0 = x_level, 1 = rhs_level, 2 = last residual vector of that level */
int mgb_amg_get_vector(mgb_amg_t h, int level, int which, double *host);
int mgb_amg_set_vector(mgb_amg_t h, int level, int which, const double *host);

/* replaces AMG::apply_smoother_operator (src/AMG.cpp:236-254) */
int mgb_amg_smooth(mgb_amg_t h, int level, int kind, int sweeps);
/* replaces AMG::apply_restriction_operator(level) (src/AMG.cpp:50-74): x_level = P^T x_{level-1} */
int mgb_amg_restrict(mgb_amg_t h, int level);
/* replaces AMG::apply_prolungation_operator(level) (src/AMG.cpp:218-232): x_level += P x_{level+1} */
int mgb_amg_prolong(mgb_amg_t h, int level);
/* replaces AMG::compute_residual(level) (src/AMG.cpp:256-275): ||rhs - A x||_2 */
int mgb_amg_residual(mgb_amg_t h, int level, double *norm);
/* replaces the body of AMG::apply_AMG() after initialization (src/AMG.cpp:282-304) */
int mgb_amg_apply(mgb_amg_t h, double *residual_norm);

/* NOT in the reference: a convergent correction-scheme V(nu1,nu2) cycle on the same
 * hierarchy -- the reference's pass restricts the solution and is not an iteration.  hist[0] = ||b - A x||_2 on entry,
 * then one entry per cycle (maxit+1 doubles); stops at hist <= tol * hist[0]. */
int mgb_amg_solve(mgb_amg_t h, double tol, int maxit, int nu1, int nu2, int coarse_sweeps, double *hist, int *n_hist);

/* The setup stages one by one, as RestrictionOperator exposes them (include/AMG.hpp:150-369) and AMG/debugtest.cpp
 * calls them.  Host side in this round (O(nnz), the reference's semantics); matrices are opaque host-CSR handles. */
typedef struct mgb_csr *mgb_csr_t;
int mgb_csr_create(size_t n_rows, size_t n_cols, const int64_t *row_ptr, const int64_t *col, const double *val, mgb_csr_t *out);
void mgb_csr_destroy(mgb_csr_t m);
int mgb_csr_info(mgb_csr_t m, size_t *n_rows, size_t *n_cols, size_t *nnz);
int mgb_csr_get(mgb_csr_t m, int64_t *row_ptr, int64_t *col, double *val);
/* replaces select_strong_connections + select_coarse_nodes (AMG.hpp:132-198): coarse_mask[i] & 0xC0 != 0 <=> node i is FINE */
int mgb_amg_select_coarse_nodes(mgb_csr_t A, double eps, int64_t start, unsigned char *coarse_mask, size_t *n_coarse);
/* replaces build_prolongation_matrix (AMG.hpp:230-300) */
int mgb_amg_build_prolongation(mgb_csr_t A, double eps, const unsigned char *coarse_mask, mgb_csr_t *P);
/* replaces build_coarse_matrix (AMG.hpp:303-369) */
int mgb_amg_build_coarse_matrix(mgb_csr_t A, mgb_csr_t P, mgb_csr_t *Ac);

/* Ghost-exchange plan of operator M for `rank` (host-only; what mgb_amg_create_sharded builds internally, exported so
 * that the index lists can be checked without a GPU).  The rank owns the rows mgb_amg_partition(n_rows) of M and the
 * vector entries mgb_amg_partition(n_cols).  recv lists the entries of other ranks its rows reference, send the
 * entries of its own block that other ranks' rows reference.  group_of_col (may be NULL, then n_groups = 1) assigns
 * every vector entry to a group (the colour of a multicolour sweep); the lists are ordered by (group, peer, index):
 * segment (g, p) of send_idx is [send_ptr[g*n_ranks+p], send_ptr[g*n_ranks+p+1]).  *_ptr hold n_groups*n_ranks+1
 * entries; pass send_idx = recv_idx = NULL to obtain the sizes (the last entry of each ptr array) first. */
int mgb_amg_halo_plan(mgb_csr_t M, int n_ranks, int rank, const int *group_of_col, int n_groups,
                      int64_t *send_ptr, int64_t *send_idx, int64_t *recv_ptr, int64_t *recv_idx);

/* 64-bit checksum over ALL ranks of vector `which` (0 x, 1 rhs, 2 residual) of `level`: independent of the row-block
 * partition, so equal values on 1 and N ranks <=> the sharded solve is bit-identical.  Collective when n_ranks > 1. */
int mgb_amg_checksum(mgb_amg_t h, int level, int which, uint64_t *out);
int mgb_amg_uses_p2p(mgb_amg_t h);     /* as mgb_gmg_uses_p2p, for the ghost exchanges of the row blocks */
int mgb_amg_get_stats(mgb_amg_t h, mgb_gmg_stats *s);
int mgb_amg_reset_stats(mgb_amg_t h);
void *mgb_amg_stream(mgb_amg_t h);
int mgb_amg_sync(mgb_amg_t h);

/* CUDA-event timer on a stream of this library (stream = mgb_gmg_stream()/mgb_amg_stream()) */
typedef struct mgb_timer *mgb_timer_t;
int mgb_timer_create(mgb_timer_t *t);
void mgb_timer_destroy(mgb_timer_t t);
int mgb_timer_start(mgb_timer_t t, void *stream);
int mgb_timer_stop(mgb_timer_t t, void *stream);
int mgb_timer_elapsed_ms(mgb_timer_t t, double *ms);   /* synchronises on the stop event */

int mgb_nccl_unique_id(unsigned char id[128]);
const char *mgb_last_error(void);
const char *mgb_version(void);
int mgb_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MGB200_H */
