// gmg_stream2.cuh -- second generation of the streaming, temporally blocked red-black Gauss-Seidel kernel
// (same contract, template parameters and results -- bit for bit -- as k_rb_stream in gmg_kernels.cuh; replaces
// GeometricMultigrid/include/solvers.hpp:33-48 reordered, multigrid.cpp:3-27, multigrid.hpp:127,134-144, main.cpp:85-86).
//
// What changed against the first generation, and why (ncu of k_rb_stream<10,0,1,1>: 314 instructions per row step
// of which 63 are fp64 arithmetic, issue-active 39 % at 8 warps per SM, 251 registers):
//   * Rows come from HBM by the bulk-copy engine (cp.async.bulk global -> shared, completion on an mbarrier),
//     requested kS2D-1 row steps ahead by ONE thread.  No load instruction, no address arithmetic and no prefetch
//     register (v1: 4 rows x 4 double2 per thread) is spent on the input any more.
//   * Every shared-memory address of the steady state is a compile-time constant: the main loop is unrolled over a
//     whole period R (24 row steps for 10 half-sweeps) of all rings, so `ring slot of row i-2s` is an immediate
//     offset from one per-thread base register instead of 11 carried offsets with wrap-around tests.
//   * The lateral exchange publishes only what the neighbour thread will read: half-sweep s writes the HALF row it
//     just updated into slot [s][step mod 3]; half-sweep s+1 reads it two steps later.  Three slots per stage make a
//     single __syncthreads per row step sufficient (write at step i, read at i+2, overwritten at i+3).
//   * The right-hand side ring is private to the thread (own column pair only): scaled once on arrival.
// Only the first two steps of a chunk run through the guarded variant of the step function (run-time slot indices).
// Everything else is the unrolled steady loop, INCLUDING the warm-up and the flush of the pipeline: half-sweeps applied
// to rows outside the valid trapezoid compute finite garbage (the window, the rhs ring and the landing slots start at
// zero; rows past the end of the streamed range are re-reads of its last row), and a valid row never reads an invalid
// one, because garbage moves towards the owned rows by one row per half-sweep -- exactly the S halo rows the chunk
// streams on either side.  Periods that touch the first or the last row of the domain run a second unrolled variant
// that knows about Dirichlet rows (BR).  Measured before this change: a chunk paid ~124 steady steps of fixed cost for
// ~50 guarded steps (profiles/r02_chunk_fit.txt) -- 25 % of a slab launch at 8 GPUs.
#pragma once
#include "gmg_common.cuh"

namespace mgb {

constexpr int kS2TW = 256;            // tile columns per CTA
constexpr int kS2NT = kS2TW / 2;      // threads: one per column pair
constexpr int kS2D = 6;               // landing slots: rows are requested in groups of kS2G, kS2G..kS2D-1 steps before they are consumed
constexpr int kS2G = 3;               // rows per request group
constexpr int kS2RB = 24;             // rows of the right-hand side ring (>= 2S+2)
constexpr int kS2NS = 3;              // publication slots per half-sweep stage
constexpr int kS2CR = 6;              // coarse-row slots (PIN)
constexpr int kS2CW = 136;            // doubles per coarse-row slot (129 used, start rounded down to an even column)
constexpr bool kS2TmemBConst = true;  // rhs ring in tensor memory (see kS2TmemB below)
constexpr int kS2HP = kS2NT + 2;      // doubles per published half row: [pad][128 values][pad]

// row steps per unrolled period of the steady loop: every ring period divides it (the rhs ring of 24 rows is reached
// through two base addresses that swap every period)
template <int S>
constexpr int s2_period() { return 12; }

// resident CTAs per SM the kernel is compiled for (register budget): 3 for the 23-row window of 10 half-sweeps
template <int S>
constexpr int s2_min_ctas() { return kS2TmemBConst ? (S <= 4 ? 4 : 3) : 2; }

template <int S, int MODE, bool PIN>
struct S2Layout {
    static_assert(2 * S + 2 <= kS2RB, "the right-hand side ring holds at most 24 rows");
    static constexpr int R = s2_period<S>();
    static_assert(R % kS2D == 0 && R % kS2NS == 0 && R % 4 == 0 && (R / 2) % kS2CR == 0 && kS2D % kS2G == 0 && kS2RB == 2 * R, "ring periods must divide R");
    static constexpr int X = (MODE == 3) ? 2 : ((MODE != 0) ? 1 : 0);
    static constexpr int HC = S + 2 * (X > 0);
    static constexpr int OW = kS2TW - 2 * HC;
    // offsets in doubles from the start of dynamic shared memory
    static constexpr int off_b = 0;                                          // [R][even 128 | odd 128]  b/diag of rows i..i-R+1
    static constexpr int off_su = off_b;                                     // [S+1][NS][HP]  published half rows
    static constexpr int off_lb = off_su + (S + 1) * kS2NS * kS2HP;          // [D][TW]        landing: rhs rows
    static constexpr int n_lu = PIN ? kS2CR * kS2CW : kS2D * kS2TW;
    static constexpr int off_lu = off_lb + kS2D * kS2TW;                     // landing: u rows, or coarse rows (PIN)
    static constexpr int n_lc = (MODE == 1) ? kS2D * OW : 0;
    static constexpr int off_lc = off_lu + n_lu;                             // [D][OW]        landing: rows of ucorr (MODE 1)
    static constexpr int n_sf = (MODE >= 2) ? kS2NS * 2 * kS2HP : 0;
    static constexpr int off_sf = off_lc + n_lc;                             // [NS][2][HP]    final pairs (MODE 2/3)
    static constexpr int n_sr = (MODE == 3) ? 8 * kS2NT : 0;
    static constexpr int off_sr = off_sf + n_sf;                             // [4][2][NT]     residual ring (MODE 3)
    static constexpr int off_bar = off_sr + n_sr;                            // D + CR mbarriers
    static constexpr int n_doubles = off_bar + kS2D + kS2CR;
    static constexpr int bytes = n_doubles * 8;
    static_assert(off_su % 2 == 0 && off_lb % 2 == 0 && off_lu % 2 == 0 && off_lc % 2 == 0 && off_sf % 2 == 0, "16-byte alignment");
};

// ---- bulk asynchronous copy + mbarrier primitives (PTX ISA 8.x, sm_90+) --------------------------------------
__device__ __forceinline__ uint32_t s2_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void s2_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void s2_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void s2_mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "S2_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra S2_DONE_%=;\n"
        "bra S2_WAIT_%=;\n"
        "S2_DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void s2_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

// ---- tensor memory as thread-private storage (tcgen05.ld/st 32x32b: lane = thread, columns = 32-bit words) -----
// The right-hand side of the 2S+2 rows in flight is read S+1 times per row step by its own thread only.  Kept in
// registers it costs 88 registers (2 CTAs per SM, spills); kept in shared memory it costs 48 KB per CTA and a third of
// the shared-memory bandwidth of a step.  TMEM (256 KB per SM, idle in a kernel without tensor-core work) holds it at no
// cost to either: 4 columns per row, ring slot = immediate offset in the unrolled loop.
constexpr bool kS2TmemB = kS2TmemBConst;
constexpr uint32_t kS2TmemCols = 128;             // >= 4 * R, power of two
__device__ __forceinline__ void s2_tm_st4(uint32_t taddr, double a, double b)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__double2loint(a)),
                 "r"(__double2hiint(a)), "r"(__double2loint(b)), "r"(__double2hiint(b))
                 : "memory");
}
__device__ __forceinline__ void s2_tm_ld2(uint32_t taddr, uint32_t &lo, uint32_t &hi)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void s2_tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void s2_tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// pins a register loaded by tcgen05.ld behind the wait (volatile asm statements keep their order)
__device__ __forceinline__ double s2_tm_value(uint32_t lo, uint32_t hi)
{
    asm volatile("" : "+r"(lo), "+r"(hi));
    return __hiloint2double((int)hi, (int)lo);
}

// per-thread / per-CTA constants of one launch (all members live in registers or uniform registers)
struct S2Ctx {
    LevelGeom g, gc;
    int t, warp, j0, i0, i1, ifirst, ilast, glast, koff, ksteps, kend;
    bool first_is_bdry, last_is_bdry, own, bc0, bc1;
    double inv_diag, q0, q1, m0, m1;
    const double *b, *uin;
    double *uout, *ucorr, *partial;
    ptrdiff_t P, Pc;
    int f_lo, f_dst, c_lo, c_dst, o_lo;          // first global column / landing offset (doubles) of the row segments
    uint32_t f_bytes, c_bytes, o_bytes;
    int kc0, kc1;                                // PIN: this thread's two coarse columns inside a coarse-row slot
    int ucx;                                     // MODE 1: this thread's pair inside a landed row of ucorr (0 for threads that own nothing)
    uint32_t sbase;                              // shared-space address of the dynamic shared memory
    uint32_t tb;                                 // TMEM address of this warp's lane quarter, column 0 of the rhs ring
    double *sm;
    int restr;
    double rscale;
};

__device__ __forceinline__ bool s2_elect()
{
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ int s2_wrap(int x, int n) { return x < 0 ? x + n : (x >= n ? x - n : x); }

// requests the rows that row step k consumes (called by one thread): rhs row (+ u row | coarse row) (+ ucorr row)
template <int S, int MODE, bool PIN>
__device__ __forceinline__ void s2_issue(const S2Ctx &c, const int k, const int kb, const bool first)
{
    using L = S2Layout<S, MODE, PIN>;
    if (k >= c.kend) return;                         // no step will consume it
    const int iu = c.ifirst + k;
    const int i = min(iu, c.ilast);                  // flush steps re-read the last row of the streamed range
    const int slot = kb % kS2D;
    const uint32_t bar = c.sbase + 8u * (uint32_t)(L::off_bar + slot);
    const uint32_t bytes = c.f_bytes * (PIN ? 1u : 2u) + (MODE == 1 ? c.o_bytes : 0u);
    s2_mbar_expect_tx(bar, bytes);
    s2_bulk_g2s(c.sbase + 8u * (uint32_t)(L::off_lb + slot * kS2TW + c.f_dst), c.b + (ptrdiff_t)i * c.P + c.f_lo, c.f_bytes, bar);
    if (!PIN)
        s2_bulk_g2s(c.sbase + 8u * (uint32_t)(L::off_lu + slot * kS2TW + c.f_dst), c.uin + (ptrdiff_t)i * c.P + c.f_lo, c.f_bytes, bar);
    if (MODE == 1) {
        const int r = min(max(iu - 2 * S, c.i0), c.i1 - 1);
        s2_bulk_g2s(c.sbase + 8u * (uint32_t)(L::off_lc + slot * L::OW), c.ucorr + (ptrdiff_t)r * c.P + c.o_lo, c.o_bytes, bar);
    }
    if (PIN) {
        // coarse row I is first needed by fine row 2I-1 (as its lower neighbour); the very first row also needs its own
        const int gi = c.g.row0 + i;
        const int m = kb >> 1;                       // slot counter of coarse row gi>>1 (koff and the first row are even)
        if (first) {
            const int sc = m % kS2CR;
            const uint32_t barc = c.sbase + 8u * (uint32_t)(L::off_bar + kS2D + sc);
            s2_mbar_expect_tx(barc, c.c_bytes);
            s2_bulk_g2s(c.sbase + 8u * (uint32_t)(L::off_lu + sc * kS2CW + c.c_dst),
                        c.uin + (ptrdiff_t)((gi >> 1) - c.gc.row0) * c.Pc + c.c_lo, c.c_bytes, barc);
        }
        if (kb & 1) {                                // parity of the (unclamped) row
            const int sc = (m + 1) % kS2CR;
            const uint32_t barc = c.sbase + 8u * (uint32_t)(L::off_bar + kS2D + sc);
            s2_mbar_expect_tx(barc, c.c_bytes);
            s2_bulk_g2s(c.sbase + 8u * (uint32_t)(L::off_lu + sc * kS2CW + c.c_dst),
                        c.uin + (ptrdiff_t)((gi >> 1) + 1 - c.gc.row0) * c.Pc + c.c_lo, c.c_bytes, barc);
        }
    }
}

// the same from a step of the unrolled loop: every slot index is a compile-time constant
template <int S, int MODE, bool PIN>
__device__ __forceinline__ void s2_issue_steady(const S2Ctx &c, const int k, const int kb)
{
    using L = S2Layout<S, MODE, PIN>;
    if (k >= c.kend) return;
    const int iu = c.ifirst + k;
    const int i = min(iu, c.ilast);
    const int slot = kb % kS2D;
    const uint32_t bar = c.sbase + 8u * (uint32_t)(L::off_bar + slot);
    s2_mbar_expect_tx(bar, c.f_bytes * (PIN ? 1u : 2u) + (MODE == 1 ? c.o_bytes : 0u));
    s2_bulk_g2s(c.sbase + 8u * (uint32_t)(L::off_lb + slot * kS2TW + c.f_dst), c.b + (ptrdiff_t)i * c.P + c.f_lo, c.f_bytes, bar);
    if (!PIN) s2_bulk_g2s(c.sbase + 8u * (uint32_t)(L::off_lu + slot * kS2TW + c.f_dst), c.uin + (ptrdiff_t)i * c.P + c.f_lo, c.f_bytes, bar);
    if (MODE == 1) {
        const int r = min(max(iu - 2 * S, c.i0), c.i1 - 1);
        s2_bulk_g2s(c.sbase + 8u * (uint32_t)(L::off_lc + slot * L::OW), c.ucorr + (ptrdiff_t)r * c.P + c.o_lo, c.o_bytes, bar);
    }
    if (PIN && (kb & 1)) {
        const int sc = ((kb >> 1) + 1) % kS2CR;
        const uint32_t barc = c.sbase + 8u * (uint32_t)(L::off_bar + kS2D + sc);
        s2_mbar_expect_tx(barc, c.c_bytes);
        s2_bulk_g2s(c.sbase + 8u * (uint32_t)(L::off_lu + sc * kS2CW + c.c_dst),
                    c.uin + (ptrdiff_t)(((c.g.row0 + i) >> 1) + 1 - c.gc.row0) * c.Pc + c.c_lo, c.c_bytes, barc);
    }
}

// One row step: row i = ifirst + k arrives, half-sweep s (s = 1..S; odd = red, even = black) is applied to row i-2s,
// row i-2S is final and leaves.  kb = (k + koff) mod R selects every ring slot; in the unrolled steady loop it is a
// compile-time constant.  GUARD = false: every half-sweep runs (rows outside the valid trapezoid hold finite garbage that
// no valid row reads); BR = true adds the test for the first / last row of the domain (Dirichlet rows) to that variant.
template <int S, bool EXACT, int MODE, bool PIN, bool GUARD, bool BR>
__device__ __forceinline__ void s2_step(const S2Ctx &c, double2 (&uw)[2 * S + 3], double2 (&bw)[kS2TmemB ? 1 : 2 * S + 2], double &acc, uint32_t &phl, uint32_t &phc,
                                        const int k, const int kb, const int kb24, const uint32_t tA, const uint32_t tB)
{
    using L = S2Layout<S, MODE, PIN>;
    constexpr int R = L::R, NT = kS2NT, TW = kS2TW, HP = kS2HP, NS = kS2NS, HC = L::HC;
    double *const sm = c.sm;
    const int t = c.t;
    const int i = c.ifirst + k;
    const int par = kb & 1;                          // parity of the global index of row i
    const LevelGeom &g = c.g;

    __syncthreads();
    // rhs ring in TMEM: column of row i-d.  Steady steps: the slot modulo 12 is static and the half of the ring comes
    // with one of two base addresses (tA = base + 48 h, tB = base - 48 h, h = period parity)
    auto tmaddr = [&](const int d) -> uint32_t {
        if (GUARD) return c.tb + 4u * (uint32_t)s2_wrap(kb24 - d, kS2RB);
        const int s0 = s2_wrap(kb - d, kS2RB);
        return (s0 < R ? tA : tB) + 4u * (uint32_t)s0;
    };
    // the row kS2D-1 steps ahead is requested by one elected lane; the requesting warp rotates over the four warps and
    // everything the request needs is warp-uniform (no per-lane address arithmetic)
    if (c.warp == (kb & 3)) {
        if (s2_elect()) {
            if (GUARD) s2_issue<S, MODE, PIN>(c, k + kS2D - 1, s2_wrap(kb + kS2D - 1, R), false);
            else s2_issue_steady<S, MODE, PIN>(c, k + kS2D - 1, s2_wrap(kb + kS2D - 1, R));
        }
    }

    // ---- the arriving row ----------------------------------------------------------------------------------
    const int sl = kb % kS2D;
    s2_mbar_wait(c.sbase + 8u * (uint32_t)(L::off_bar + sl), (phl >> sl) & 1u);
    phl ^= 1u << sl;
    double2 nb = make_double2(0., 0.), nu = make_double2(0., 0.);
    {
        nb = ld2(sm + L::off_lb + sl * TW + 2 * t);
        if (!PIN) nu = ld2(sm + L::off_lu + sl * TW + 2 * t);
        else {
            const int m = kb >> 1;
            const int sc0 = m % kS2CR, sc1 = (m + 1) % kS2CR;
            if (GUARD && k == 0) { s2_mbar_wait(c.sbase + 8u * (uint32_t)(L::off_bar + kS2D + sc0), (phc >> sc0) & 1u); phc ^= 1u << sc0; }
            if (par) { s2_mbar_wait(c.sbase + 8u * (uint32_t)(L::off_bar + kS2D + sc1), (phc >> sc1) & 1u); phc ^= 1u << sc1; }
            const double *c0 = sm + L::off_lu + sc0 * kS2CW, *c1 = sm + L::off_lu + sc1 * kS2CW;
            double a = c0[c.kc0], c2 = c0[c.kc1];
            if (par) {   // multigrid.cpp:3-27: vertical midpoints first, then the odd columns of every fine row
                a = __dmul_rn(0.5, __dadd_rn(a, c1[c.kc0]));
                c2 = __dmul_rn(0.5, __dadd_rn(c2, c1[c.kc1]));
            }
            nu = make_double2(a, __dmul_rn(0.5, __dadd_rn(a, c2)));
        }
    }
    if (!EXACT) {
        // the ring holds bq = b/diag (b itself on Dirichlet points): u = bq + q * (sum of neighbours), q = 1/4 or 0
        if (GUARD || BR) {
            const bool brow = (i + g.row0 == 0) || (i == c.glast);
            nb.x = (c.bc0 || brow) ? nb.x : nb.x * c.inv_diag;
            nb.y = (c.bc1 || brow) ? nb.y : nb.y * c.inv_diag;
        } else {
            nb.x *= c.m0; nb.y *= c.m1;
        }
    }
    if (kS2TmemB) s2_tm_st4(tmaddr(0), nb.x, nb.y);
    else {
#pragma unroll
        for (int d = 2 * S + 1; d > 0; --d) bw[d] = bw[d - 1];
        bw[0] = nb;
    }
#pragma unroll
    for (int d = 2 * S + 2; d > 0; --d) uw[d] = uw[d - 1];
    uw[0] = nu;
    const int ss = kb % NS, ss1 = s2_wrap(ss - 1, NS), ss2 = s2_wrap(ss - 2, NS);
    // stage 0: the half row that half-sweep 1 (two steps from now, parity `par`) needs from the neighbour thread
    sm[L::off_su + ss * HP + 1 + t] = par ? nu.x : nu.y;

    // operands of the leaving rows, requested together with those of the half-sweeps (their latency overlaps)
    double2 uc = make_double2(0., 0.);
    double e_lf = 0., e_rt = 0.;
    if (MODE == 1) {
        uc = ld2(sm + L::off_lc + sl * L::OW + c.ucx);
        const double *pub = sm + L::off_su + (S * NS + ss1) * HP + 1 + t;      // half row published by half-sweep S one step ago
        if (par) e_lf = pub[-1]; else e_rt = pub[1];
    }
    if (MODE >= 2) {
        e_lf = sm[L::off_sf + (ss1 * 2 + 1) * HP + t];
        e_rt = sm[L::off_sf + (ss1 * 2 + 0) * HP + 2 + t];
    }

    // ---- S half-sweeps, mutually independent within a step ---------------------------------------------------
    double bv[S], ob[S];
    double eb0 = 0., eb1 = 0.;                       // rhs of the row whose residual is formed (row i-2S-1)
    if (kS2TmemB) {
        uint32_t lo[S + 2], hi[S + 2];
        const uint32_t tq = tmaddr(2 * S + 1);
        s2_tm_wait_st();                             // stores of earlier steps (this step's goes to a slot not read now)
#pragma unroll
        for (int s = 1; s <= S; ++s) {
            const int which = (par + s - 1) & 1;
            s2_tm_ld2(tmaddr(2 * s) + 2u * (uint32_t)which, lo[s - 1], hi[s - 1]);
        }
        if (MODE == 1) s2_tm_ld2(tq + (par ? 0u : 2u), lo[S], hi[S]);
        if (MODE >= 2) { s2_tm_ld2(tq, lo[S], hi[S]); s2_tm_ld2(tq + 2u, lo[S + 1], hi[S + 1]); }
#pragma unroll
        for (int s = 1; s <= S; ++s) {
            const int which = (par + s - 1) & 1;
            ob[s - 1] = sm[L::off_su + ((s - 1) * NS + ss2) * HP + 1 + t + (which ? 1 : -1)];
        }
        s2_tm_wait_ld();
#pragma unroll
        for (int s = 1; s <= S; ++s) bv[s - 1] = s2_tm_value(lo[s - 1], hi[s - 1]);
        if (MODE == 1) { const double v = s2_tm_value(lo[S], hi[S]); if (par) eb0 = v; else eb1 = v; }
        if (MODE >= 2) { eb0 = s2_tm_value(lo[S], hi[S]); eb1 = s2_tm_value(lo[S + 1], hi[S + 1]); }
    } else {
#pragma unroll
        for (int s = 1; s <= S; ++s) {
            const int which = (par + s - 1) & 1;
            bv[s - 1] = which ? bw[2 * s].y : bw[2 * s].x;
            ob[s - 1] = sm[L::off_su + ((s - 1) * NS + ss2) * HP + 1 + t + (which ? 1 : -1)];
        }
        eb0 = bw[2 * S + 1].x; eb1 = bw[2 * S + 1].y;
    }
#pragma unroll
    for (int s = 1; s <= S; ++s) {
        const int which = (par + s - 1) & 1;
        const int d = 2 * s;
        bool act = true, brow = false;
        bool isb = which ? c.bc1 : c.bc0;
        if (GUARD) {
            const int r = i - d;
            const int vlo = c.first_is_bdry ? c.ifirst : c.ifirst + s;
            const int vhi = c.last_is_bdry ? c.ilast : c.ilast - s;
            act = (r >= vlo) && (r <= vhi);
            brow = (r + g.row0 == 0) || (r == c.glast);
            isb = isb || brow;
        } else if (BR) {
            const int r = i - d;
            brow = (r + g.row0 == 0) || (r == c.glast);
            isb = isb || brow;
        }
        const double up = which ? uw[d + 1].y : uw[d + 1].x;
        const double dn = which ? uw[d - 1].y : uw[d - 1].x;
        const double left = which ? uw[d].x : ob[s - 1];
        const double right = which ? ob[s - 1] : uw[d].y;
        double nv;
        if (EXACT) {
            nv = smooth_point(bv[s - 1], up, left, right, dn, g.off, g.diag);
            nv = isb ? bv[s - 1] : nv;
        } else {
            double q = which ? c.q1 : c.q0;
            if (GUARD || BR) q = brow ? 0. : q;
            nv = fma(q, (up + dn) + (left + right), bv[s - 1]);
        }
        if (act) {
            if (which) uw[d].y = nv; else uw[d].x = nv;
            if (s < S || MODE == 1) sm[L::off_su + (s * NS + ss) * HP + 1 + t] = nv;
        }
    }
    if (MODE >= 2) {   // both components of the final row for the residual of the next step
        sm[L::off_sf + (ss * 2 + 0) * HP + 1 + t] = uw[2 * S].x;
        sm[L::off_sf + (ss * 2 + 1) * HP + 1 + t] = uw[2 * S].y;
    }

    // ---- the leaving row ---------------------------------------------------------------------------------------
    const int r = i - 2 * S;
    const bool in_rows = (r >= c.i0) && (r < c.i1);
    if (MODE == 0 || MODE == 2 || MODE == 3) {
        if (in_rows && c.own) {
            double *dstp = c.uout + (ptrdiff_t)r * c.P + c.j0;
            if (c.j0 + 1 < g.w) st2(dstp, uw[2 * S]); else dstp[0] = uw[2 * S].x;
        }
    }
    const int q = r - 1;                              // rows q-1, q, q+1 are final: residual of row q
    if (MODE == 2 || MODE == 3) {
        const int fin_lo = c.first_is_bdry ? c.ifirst : c.ifirst + S, fin_hi = c.last_is_bdry ? c.ilast : c.ilast - S;
        const bool rows_ok = (MODE == 3) ? (q >= c.i0 - 1 && q <= c.i1 && q + g.row0 >= 0 && q <= c.glast &&
                                            (q - 1 >= fin_lo || q + g.row0 == 0) && (q + 1 <= fin_hi || q == c.glast))
                                         : (q >= c.i0 && q < c.i1);
        const bool cols_ok = (MODE == 3) ? (2 * t >= HC - 2 && 2 * t < TW - HC + 2 && c.j0 >= 0 && c.j0 < g.w) : c.own;
        double2 rv = make_double2(0., 0.);
        if (rows_ok && cols_ok) {
            const double2 ce = uw[2 * S + 1], up = uw[2 * S + 2], dn = uw[2 * S];
            const double lf = e_lf, rt = e_rt;
            const double b0 = eb0, b1 = eb1;
            const bool brow = (q + g.row0 == 0) || (q == c.glast);
            if (EXACT) {
                rv.x = (c.bc0 || brow) ? __dsub_rn(b0, ce.x) : resid_point(b0, up.x, lf, ce.x, ce.y, dn.x, g.off, g.diag);
                rv.y = (c.bc1 || brow) ? __dsub_rn(b1, ce.y) : resid_point(b1, up.y, ce.x, ce.y, rt, dn.y, g.off, g.diag);
            } else {
                const double q0 = (c.bc0 || brow) ? 0. : 0.25, q1 = (c.bc1 || brow) ? 0. : 0.25;
                const double w0 = (c.bc0 || brow) ? 1. : g.diag, w1 = (c.bc1 || brow) ? 1. : g.diag;
                rv.x = w0 * fma(q0, (up.x + dn.x) + (lf + ce.y), b0 - ce.x);
                rv.y = w1 * fma(q1, (up.y + dn.y) + (ce.x + rt), b1 - ce.y);
            }
            if (q >= c.i0 && q < c.i1 && c.own) {
                double *dstp = c.ucorr + (ptrdiff_t)q * c.P + c.j0;
                if (c.j0 + 1 < g.w) st2(dstp, rv); else dstp[0] = rv.x;
            }
        }
        if (MODE == 3) {
            constexpr int NT2 = 2 * NT;
            double *sr = sm + L::off_sr;
            const int k4 = kb & 3;                   // ring slot of row q
            sr[k4 * NT2 + t] = rv.x; sr[k4 * NT2 + NT + t] = rv.y;
            // coarse row centred on fine row cq = q-2 (rows cq-1, cq, cq+1 were published in earlier steps)
            const int cq = q - 2, gq = g.row0 + cq;
            if (((gq & 1) == 0) && cq >= c.i0 && cq < c.i1 && c.own) {
                const int gI = gq >> 1, J = c.j0 >> 1;
                const double *rm = sr + ((k4 + 1) & 3) * NT2 + t;      // row cq-1
                const double *rc = sr + ((k4 + 2) & 3) * NT2 + t;      // row cq
                const double *rp = sr + ((k4 + 3) & 3) * NT2 + t;      // row cq+1
                const double c0 = rc[0];
                double v;
                if (gI == 0 || gI == c.gc.w - 1 || J == 0 || J == c.gc.w - 1) v = c0;
                else if (c.restr != 2) v = __dmul_rn(c.rscale, c0);
                else {   // [x][t] = column j0, [y][t] = column j0+1, [y][t-1] = column j0-1
                    const double edge = __dadd_rn(__dadd_rn(__dadd_rn(rm[0], rc[NT - 1]), rc[NT]), rp[0]);
                    const double corner = __dadd_rn(__dadd_rn(__dadd_rn(rm[NT - 1], rm[NT]), rp[NT - 1]), rp[NT]);
                    v = __dadd_rn(__dadd_rn(__dmul_rn(0.25, c0), __dmul_rn(0.125, edge)), __dmul_rn(0.0625, corner));
                }
                c.partial[(ptrdiff_t)(gI - c.gc.row0) * c.gc.pitch + J] = v;
            }
        }
    }
    if (MODE == 1) {
        if (in_rows && c.own) {
            double *dstp = c.ucorr + (ptrdiff_t)r * c.P + c.j0;
            const double2 un = make_double2(__dadd_rn(uc.x, uw[2 * S].x), __dadd_rn(uc.y, uw[2 * S].y));
            if (c.j0 + 1 < g.w) st2(dstp, un); else dstp[0] = un.x;
        }
        // Only the RED interior point of the pair contributes: the black one was the last one updated and nothing
        // around it changed since (its residual is a rounding error of its own update); Dirichlet points have residual 0.
        const bool brow = (q + g.row0 == 0) || (q == c.glast);
        if (q >= c.i0 && q < c.i1 && c.own && !brow) {
            const double2 ce = uw[2 * S + 1], up = uw[2 * S + 2], dn = uw[2 * S];
            double rv;
            if (par) {        // row q has the other parity than row i: red = even column
                const double lf = e_lf, b0 = eb0;
                if (EXACT) rv = c.bc0 ? 0. : resid_point(b0, up.x, lf, ce.x, ce.y, dn.x, g.off, g.diag);
                else rv = c.bc0 ? 0. : g.diag * fma(0.25, (up.x + dn.x) + (lf + ce.y), b0 - ce.x);
            } else {
                const double rt = e_rt, b1 = eb1;
                if (EXACT) rv = c.bc1 ? 0. : resid_point(b1, up.y, ce.x, ce.y, rt, dn.y, g.off, g.diag);
                else rv = c.bc1 ? 0. : g.diag * fma(0.25, (up.y + dn.y) + (ce.x + rt), b1 - ce.y);
            }
            acc += rv * rv;
        }
    }
}

template <int S, bool EXACT, int MODE, bool PIN>
__global__ void __launch_bounds__(kS2NT, s2_min_ctas<S>())
k_rb_stream2(LevelGeom g, const double *__restrict__ uin, const double *__restrict__ b, double *__restrict__ uout,
             int rows_per_chunk, double *ucorr, double *__restrict__ partial, LevelGeom gc, int restr, double rscale)
{
    using L = S2Layout<S, MODE, PIN>;
    constexpr int R = L::R, TW = kS2TW, X = L::X, HC = L::HC, OW = L::OW;
    extern __shared__ __align__(128) double s2_smem[];
    __shared__ double red[kS2NT / 32];
    __shared__ uint32_t tm_base;

    S2Ctx c;
    c.g = g; c.gc = gc; c.b = b; c.uin = uin; c.uout = uout; c.ucorr = ucorr; c.partial = partial;
    c.restr = restr; c.rscale = rscale;
    c.sm = s2_smem; c.sbase = s2_smem_u32(s2_smem);
    const int t = threadIdx.x;
    c.t = t;
    c.warp = __shfl_sync(0xffffffffu, t >> 5, 0);      // warp-uniform copy of the warp index
    const int jbase = blockIdx.x * OW - HC;           // global column of tile column 0 (even)
    c.j0 = jbase + 2 * t;
    c.i0 = blockIdx.y * rows_per_chunk;               // rows_per_chunk is even (host)
    c.i1 = min(c.i0 + rows_per_chunk, g.rows);
    if (c.i0 >= g.rows) return;
    const bool top_is_domain = (g.row0 == 0), bot_is_domain = (g.row0 + g.rows == g.w);
    const int lo = top_is_domain ? 0 : -S - X, hi = bot_is_domain ? g.rows - 1 : g.rows - 1 + S + X;
    int ifirst = max(c.i0 - S - X, lo);
    if ((g.row0 + ifirst) & 1) ifirst -= 1;           // the first streamed row has an even global index
    c.ifirst = ifirst;
    c.ilast = min(c.i1 - 1 + S + X, hi);
    c.first_is_bdry = (g.row0 + ifirst == 0);
    c.last_is_bdry = (g.row0 + c.ilast == g.w - 1);
    c.P = g.pitch; c.Pc = gc.pitch;
    c.inv_diag = 1.0 / g.diag;
    c.own = (2 * t >= HC) && (2 * t < TW - HC) && (c.j0 < g.w) && (c.j0 >= 0);
    c.ucx = c.own ? 2 * t - HC : 0;
    c.bc0 = (c.j0 <= 0) || (c.j0 >= g.w - 1);
    c.bc1 = (c.j0 + 1 <= 0) || (c.j0 + 1 >= g.w - 1);
    c.q0 = c.bc0 ? 0. : 0.25; c.q1 = c.bc1 ? 0. : 0.25;
    c.m0 = c.bc0 ? 1. : c.inv_diag; c.m1 = c.bc1 ? 1. : c.inv_diag;
    c.glast = g.w - 1 - g.row0;
    c.ksteps = (c.i1 - 1 + 2 * S + X) - ifirst + 1 + (MODE == 3 ? 1 : 0);
    // row segments moved by the copy engine (16-byte aligned starts and sizes: jbase, HC, OW and the pitch are even)
    c.f_lo = max(jbase, 0); c.f_dst = c.f_lo - jbase;
    c.f_bytes = 8u * (uint32_t)(min(jbase + TW, g.pitch) - c.f_lo);
    c.o_lo = max(jbase + HC, 0);                      // = jbase + HC: a tile's first owned column is inside the domain
    c.o_bytes = 8u * (uint32_t)(min(jbase + TW - HC, g.pitch) - c.o_lo);
    {
        const int cs = jbase >> 1, cs_al = cs & ~1;   // first coarse column of the tile, rounded down to an even column
        c.c_lo = max(cs_al, 0); c.c_dst = c.c_lo - cs_al;
        c.c_bytes = PIN ? 8u * (uint32_t)(min(cs_al + kS2CW, gc.pitch) - c.c_lo) : 0u;
        const int Jc = min(max(c.j0 >> 1, 0), gc.w - 1), Jc1 = min(Jc + 1, gc.w - 1);
        c.kc0 = Jc - cs_al; c.kc1 = Jc1 - cs_al;
    }
    // two guarded steps (the first coarse row of the interpolating variant has its own wait), then whole unrolled periods
    // until the last row has left; requests stop at kend, so no copy is in flight when the CTA exits
    const int k_s0 = 2;
    c.koff = (kS2RB - k_s0 % kS2RB) % kS2RB;        // (k + koff) mod 24 = slot of the rhs ring, mod 12 = every other slot
    c.kend = k_s0 + ((max(c.ksteps - k_s0, 0) + R - 1) / R) * R;

    // zero the whole buffer once (pads and the parts of edge tiles that no copy covers must hold finite values)
    for (int x = t; x < L::n_doubles; x += kS2NT) s2_smem[x] = 0.;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (t == 0) {
        for (int x = 0; x < kS2D + kS2CR; ++x) s2_mbar_init(c.sbase + 8u * (uint32_t)(L::off_bar + x), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (t == 0) {
        for (int kp = 0; kp < kS2D - 1; ++kp) s2_issue<S, MODE, PIN>(c, kp, (kp + c.koff) % R, kp == 0);
    }
    c.tb = 0;
    if (kS2TmemB) {
        static_assert(4 * kS2RB <= (int)kS2TmemCols, "rhs ring does not fit the TMEM allocation");
        if (t < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2_smem_u32(&tm_base)), "r"(kS2TmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        c.tb = tm_base + ((uint32_t)(t & 96) << 16);      // lanes 32*(warp % 4) .. +31 belong to this warp
        // the ring starts at zero: the warm-up steps read rows that never arrived
        for (int x = 0; x < kS2RB; ++x) s2_tm_st4(c.tb + 4u * (uint32_t)x, 0., 0.);
        s2_tm_wait_st();
    }

    double2 uw[2 * S + 3];
#pragma unroll
    for (int d = 0; d < 2 * S + 3; ++d) uw[d] = make_double2(0., 0.);
    double2 bw[kS2TmemB ? 1 : 2 * S + 2];         // register copy of the rhs rows in flight when TMEM is not used
#pragma unroll
    for (int d = 0; d < (kS2TmemB ? 1 : 2 * S + 2); ++d) bw[d] = make_double2(0., 0.);
    double acc = 0.;
    uint32_t phl = 0, phc = 0;
    int k = 0;
    for (; k < k_s0; ++k) s2_step<S, EXACT, MODE, PIN, true, false>(c, uw, bw, acc, phl, phc, k, (k + c.koff) % R, (k + c.koff) % kS2RB, 0u, 0u);
    uint32_t tA = c.tb, tB = c.tb;                    // the first unrolled step has (k + koff) mod 24 == 0
    for (; k < c.kend; k += R) {
        // rows ifirst + k - 2S - 2 .. ifirst + k + R - 1 are touched in this period: does it see a Dirichlet row?
        const int lo_g = g.row0 + ifirst + k - 2 * S - 2, hi_g = g.row0 + ifirst + k + R - 1;
        if (lo_g <= 0 || hi_g >= g.w - 1) {
#pragma unroll
            for (int cc = 0; cc < R; ++cc) s2_step<S, EXACT, MODE, PIN, false, true>(c, uw, bw, acc, phl, phc, k + cc, cc, 0, tA, tB);
        } else {
#pragma unroll
            for (int cc = 0; cc < R; ++cc) s2_step<S, EXACT, MODE, PIN, false, false>(c, uw, bw, acc, phl, phc, k + cc, cc, 0, tA, tB);
        }
        tA = (tA == c.tb) ? c.tb + 4u * R : c.tb;
        tB = (tB == c.tb) ? c.tb - 4u * R : c.tb;
    }

    if (MODE == 1) {
        const double tsum = block_sum(acc, red);
        if (t == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = tsum;
    }
    if (kS2TmemB) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_base), "r"(kS2TmemCols) : "memory");
    }
}

}  // namespace mgb
