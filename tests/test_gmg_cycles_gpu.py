"""GPU tests of the cycle options beyond the reference's sawtooth:
V / W / F correction-scheme cycles with nu_pre pre-sweeps, the full-multigrid pass and the Krylov solvers
(CG / BiCGSTAB with one multigrid cycle as the preconditioner).

The reference has none of these; the checker is the CPU statement of the same algorithms in oracle/gmg_oracle.c
(built from the reference's own per-point formulas).  Sweeps, residuals and transfers are evaluated in the
reference's order with unfused operations on both sides, so whole cycles must agree BIT FOR BIT; the Krylov solvers
contain dot products (summation order differs) and are compared to 1e-9.
"""
import numpy as np
import pytest

import oracle
from multigrid_prj_b200 import Gmg, GmgConfig
from multigrid_prj_b200 import gmg as G

pytestmark = pytest.mark.gpu
W = 10.0
OK = {G.JACOBI: oracle.JACOBI, G.GS_LEX: oracle.GS, G.GS_RB: oracle.RBGS}


@pytest.fixture(scope="module")
def orc():
    return oracle.gmg()


def relres(orc, N, alpha, u, b):
    ss, _ = orc.residual(N, W, alpha, 0, u, b, store=False)
    return np.sqrt(ss / orc.sumsq(N, W, alpha, 0, b))


def tail_level(N, L, tail_max_width):
    """first level of the persistent coarse tail = the coarse solver of the textbook cycles"""
    w = N
    for l in range(L):
        if w <= tail_max_width:
            return l
        w = (w + 1) // 2
    return L - 1


@pytest.mark.parametrize("cycle", [G.CYCLE_V, G.CYCLE_W, G.CYCLE_F])
@pytest.mark.parametrize("kind,restriction,fused", [(G.GS_RB, G.FULL_WEIGHTING, 1), (G.GS_RB, G.FULL_WEIGHTING, 0),
                                                   (G.GS_RB, G.HALF_INJECTION, 1), (G.JACOBI, G.INJECTION, 0),
                                                   (G.GS_LEX, G.FULL_WEIGHTING, 0)])
def test_textbook_cycles_bit_identical_to_cpu_statement(orc, cycle, kind, restriction, fused):
    N, L, alpha, tmw, nu_pre, nu = 129, 6, 1.5, 17, 2, 3
    b = orc.rhs(N, W, 1)
    cfg = GmgConfig(n=N, levels=L, alpha=alpha, smoother=kind, pre_smoother=kind, restriction=restriction, nu=nu,
                    tail_max_width=tmw, cycle_type=cycle, nu_pre=nu_pre, rb_fast_arith=0, rb_fused=fused,
                    fuse_residual=fused, fuse_correction=fused, fuse_prolong=fused)
    with Gmg(cfg) as g:
        g.set_rhs(b.reshape(N, N)); g.set_u(None)
        hist = g.solve(tol=0.0, maxiter=3)
        u = g.get_u().reshape(-1)
    t = orc.textbook(N, W, alpha, L, OK[kind], cycle, tail_level(N, L, tmw), nu_pre=nu_pre, nu=nu, restrict_mode=restriction)
    u_o = np.zeros(N * N)
    h_o = [relres(orc, N, alpha, u_o, b)]
    for _ in range(3):
        t.iteration(u_o, b, pre_kind=OK[kind], n_pre=2)
        h_o.append(relres(orc, N, alpha, u_o, b))
    assert np.array_equal(u, u_o), np.abs(u - u_o).max()
    assert np.allclose(hist, h_o, rtol=1e-11, atol=0)


def test_graph_replay_of_textbook_cycles_equals_stepwise(orc):
    """mgb_gmg_run_cycles captures the V-cycle iteration in a CUDA graph: same bits as the uncaptured steps"""
    N, L = 257, 7
    b = orc.rhs(N, W, 1)
    sums = []
    for graph in (1, 0):
        cfg = GmgConfig.fast(N, L, cycle_type=G.CYCLE_V, nu_pre=2, nu=2, use_graph=graph, tail_max_width=33)
        with Gmg(cfg) as g:
            g.set_rhs(b.reshape(N, N)); g.set_u(None)
            g.run_cycles(3); g.run_cycles(8)
            st = g.stats()
            sums.append((g.checksum(), st["graph_launches"]))
    assert sums[0][0] == sums[1][0]
    assert sums[0][1] > 0 and sums[1][1] == 0


@pytest.mark.parametrize("N,L", [(257, 7), (1025, 9)])
def test_w_and_f_cycles_converge_faster_than_v_and_all_reach_the_reference_solution(orc, N, L):
    """V(2,2), W(2,2), F(2,2) with red-black GS + full weighting: cycles to 1e-10, and the converged solution against
    the reference algorithm's (lexicographic GS sawtooth, the oracle) to <= 1e-8 relative L2 (north_star's bar)"""
    b = orc.rhs(N, W, 1)
    counts, sols = {}, {}
    for name, c in (("V", G.CYCLE_V), ("W", G.CYCLE_W), ("F", G.CYCLE_F), ("sawtooth", G.CYCLE_SAWTOOTH)):
        with Gmg(GmgConfig.fast(N, L, cycle_type=c, nu_pre=2, nu=2 if c else 5, n_pre=0 if c else 2)) as g:
            g.set_rhs(b.reshape(N, N)); g.set_u(None)
            h = g.solve(tol=1e-10, maxiter=60)
            counts[name] = h.size - 1
            sols[name] = g.get_u().reshape(-1)
            assert h[-1] <= 1e-10, (name, h)
    assert counts["W"] <= counts["V"] and counts["F"] <= counts["V"], counts
    assert counts["V"] <= 14 and counts["W"] <= 10, counts
    if N <= 257:
        u_ref, h_ref, _, _ = orc.solve(N, W, 1.0, L, oracle.GS, b, tol=1e-11)
        for name, u in sols.items():
            assert np.linalg.norm(u - u_ref) <= 1e-8 * np.linalg.norm(u_ref), name


@pytest.mark.parametrize("kind,restriction", [(G.GS_RB, G.FULL_WEIGHTING), (G.JACOBI, G.INJECTION)])
def test_fmg_pass_bit_identical_to_cpu_statement(orc, kind, restriction):
    N, L, tmw = 129, 6, 17
    b = orc.rhs(N, W, 1)
    cfg = GmgConfig(n=N, levels=L, smoother=kind, pre_smoother=kind, restriction=restriction, nu=2, nu_pre=2,
                    tail_max_width=tmw, rb_fast_arith=0)
    with Gmg(cfg) as g:
        g.set_rhs(b.reshape(N, N)); g.set_u(None)
        g.fmg()
        u = g.get_u().reshape(-1)
    t = orc.textbook(N, W, 1.0, L, OK[kind], G.CYCLE_V, tail_level(N, L, tmw), nu_pre=2, nu=2, restrict_mode=restriction)
    u_o = t.fmg(np.zeros(N * N), b)
    assert np.array_equal(u, u_o), np.abs(u - u_o).max()


def test_fmg_start_reaches_discretisation_accuracy_in_one_pass(orc):
    """full multigrid: ONE pass from u = 0 lands within an order of magnitude of the discretisation error of the analytic
    solution u = exp(x) exp(-2y) of test problem 1 (bilinear interpolation of the coarser correction, one V(2,2) per
    level: measured 7.3x at 1025^2), and mgb_gmg_solve with cfg.fmg = 1 starts from it"""
    N, L = 1025, 9
    cfg = GmgConfig.fast(N, L, cycle_type=G.CYCLE_V, nu_pre=2, nu=2, fmg=1)
    with Gmg(cfg) as g:
        g.set_rhs_test(1); g.set_u(None)
        h = g.solve(tol=1e-10, maxiter=30)
        u_conv = g.get_u()
        g.set_u(None)
        g.fmg()
        u1 = g.get_u()
    assert h[1] < 5e-3 * h[0] and h[-1] <= 1e-10
    x = np.arange(N) * (W / (N - 1)); y = W - np.arange(N) * (W / (N - 1))
    exact = np.exp(x)[None, :] * np.exp(-2 * y)[:, None]
    disc = np.abs(u_conv - exact).max()
    assert np.abs(u1 - exact).max() <= 10.0 * disc, (np.abs(u1 - exact).max(), disc)
    with Gmg(GmgConfig.fast(N, L, cycle_type=G.CYCLE_V, nu_pre=2, nu=2)) as g:
        g.set_rhs_test(1); g.set_u(None)
        h0 = g.solve(tol=1e-10, maxiter=30)
    assert h.size <= h0.size


@pytest.mark.parametrize("method,precond,kind,omega,cycle,maxit", [
    (G.KRYLOV_CG, G.PRECOND_NONE, G.JACOBI, 1.0, G.CYCLE_V, 40),
    (G.KRYLOV_CG, G.PRECOND_MG, G.JACOBI, 0.8, G.CYCLE_V, 30),
    (G.KRYLOV_BICGSTAB, G.PRECOND_NONE, G.GS_RB, 1.0, G.CYCLE_V, 40),
    (G.KRYLOV_BICGSTAB, G.PRECOND_MG, G.GS_RB, 1.0, G.CYCLE_V, 30),
    (G.KRYLOV_BICGSTAB, G.PRECOND_MG, G.GS_RB, 1.0, G.CYCLE_SAWTOOTH, 30),
])
def test_krylov_against_cpu_statement(orc, method, precond, kind, omega, cycle, maxit):
    N, L, tol = 129, 7, 1e-10
    b = orc.rhs(N, W, 1)
    tmw = 0 if kind == G.JACOBI else 17          # weighted Jacobi runs level by level (no tail): the coarse solve is level L-1
    cfg = GmgConfig(n=N, levels=L, smoother=kind, pre_smoother=kind, restriction=G.FULL_WEIGHTING, nu=2, nu_pre=2,
                    tail_max_width=tmw, cycle_type=cycle, rb_fast_arith=0, jacobi_omega=omega)
    with Gmg(cfg) as g:
        g.set_rhs(b.reshape(N, N)); g.set_u(None)
        hist = g.krylov(method, precond, tol=tol, maxit=maxit)
        u = g.get_u().reshape(-1)
    bottom = L - 1 if kind == G.JACOBI else tail_level(N, L, tmw)
    t = orc.textbook(N, W, 1.0, L, OK[kind], cycle, bottom, nu_pre=2, nu=2, restrict_mode=2, omega=omega)
    u_o, h_o = t.krylov(method, precond, b, np.zeros(N * N), tol=tol, maxit=maxit)
    assert hist.size == h_o.size, (hist, h_o)
    if precond == G.PRECOND_NONE:
        # unpreconditioned recurrences amplify rounding differences step by step: the first steps pin the algorithm
        assert np.allclose(hist[:12], h_o[:12], rtol=1e-6), (hist[:12], h_o[:12])
        return
    big = h_o > 1e-7                                # entries far above the rounding floor of the recurrences
    assert np.allclose(hist[big], h_o[big], rtol=1e-7), (hist, h_o)
    assert np.linalg.norm(u - u_o) <= 1e-9 * np.linalg.norm(u_o)
    if precond == G.PRECOND_MG:
        assert hist[-1] <= tol and hist.size <= 12, hist      # a multigrid-preconditioned Krylov method converges in ~10 steps


def test_mg_preconditioned_krylov_reaches_the_reference_solution(orc):
    """the converged Krylov solutions agree with the reference algorithm's solution (lexicographic GS sawtooth) to 1e-8"""
    N, L = 257, 7
    b = orc.rhs(N, W, 1)
    u_ref, _, _, _ = orc.solve(N, W, 1.0, L, oracle.GS, b, tol=1e-11)
    for method, cfg in ((G.KRYLOV_BICGSTAB, GmgConfig.fast(N, L)),
                        (G.KRYLOV_BICGSTAB, GmgConfig.fast(N, L, cycle_type=G.CYCLE_W, nu_pre=1, nu=1)),
                        (G.KRYLOV_CG, GmgConfig(n=N, levels=L, smoother=G.JACOBI, pre_smoother=G.JACOBI, jacobi_omega=0.8,
                                                restriction=G.FULL_WEIGHTING, cycle_type=G.CYCLE_V, nu_pre=2, nu=2))):
        with Gmg(cfg) as g:
            g.set_rhs(b.reshape(N, N)); g.set_u(None)
            h = g.krylov(method, G.PRECOND_MG, tol=1e-11, maxit=40)
            u = g.get_u().reshape(-1)
        assert h[-1] <= 1e-11, h
        assert np.linalg.norm(u - u_ref) <= 1e-8 * np.linalg.norm(u_ref)


def test_krylov_at_full_size_converges():
    """8193^2 (BASELINE configs[2]): BiCGSTAB preconditioned by the fast sawtooth cycle against the plain iteration"""
    N, L = 8193, 13
    with Gmg(GmgConfig.fast(N, L)) as g:
        g.set_rhs_test(1); g.set_u(None)
        h_mg = g.solve(tol=1e-9, maxiter=30)
        g.set_u(None)
        h_k = g.krylov(G.KRYLOV_BICGSTAB, G.PRECOND_MG, tol=1e-9, maxit=30)
    assert h_mg[-1] <= 1e-9 and h_k[-1] <= 1e-9
    assert h_k.size <= h_mg.size          # one BiCGSTAB step applies the cycle twice and converges in at most as many steps
