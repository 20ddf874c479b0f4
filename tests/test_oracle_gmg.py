"""Pins oracle/gmg_oracle.c (the CPU restatement) to the reference.

(1) the reference's committed golden runs (6 significant digits in the files),
(2) raw-double outputs of the compiled reference stored in tests/golden/gmg_ref_ops.npz,
(3) live, bit for bit, against oracle/_ref/libgmgref.so when it is present.
"""
import numpy as np
import pytest

import oracle

W, ALPHA = 10.0, 1.0


def sig6(x):
    """what `ostream << double` keeps (utilities.hpp:43-54): 6 significant digits"""
    return np.array([float(f"{v:.6g}") for v in np.asarray(x).ravel()])


def test_golden_n145_jacobi(gmg_oracle, goldens):
    g, _ = goldens
    b = gmg_oracle.rhs(145, W, 1)
    u, hist, _, _ = gmg_oracle.solve(145, W, ALPHA, 5, oracle.JACOBI, b)
    assert hist.size == g["n145_hist"].size == 13
    assert np.array_equal(sig6(hist), g["n145_hist"])
    assert np.array_equal(sig6(u), g["n145_x"])


@pytest.mark.slow
def test_golden_n385_smt2(gmg_oracle, goldens):
    g, _ = goldens
    b = gmg_oracle.rhs(385, W, 0)
    u, hist, _, _ = gmg_oracle.solve(385, W, ALPHA, 5, oracle.BICGSTAB, b)   # -smt 2 runs Jacobi-MG
    assert np.array_equal(sig6(hist), g["n385_hist"])
    assert np.array_equal(sig6(u), g["n385_x"])


def test_stored_reference_operators(gmg_oracle, goldens):
    _, ops = goldens
    N = int(ops["ops_N"][0])
    u0, b0 = ops["ops_u0"], ops["ops_b0"]
    rng = np.random.default_rng(20261018)
    assert np.array_equal(rng.standard_normal(N * N), u0)
    for level in range(4):
        assert np.array_equal(gmg_oracle.sweep(oracle.GS, N, W, ALPHA, level, u0.copy(), b0), ops[f"gs_l{level}"])
        assert np.array_equal(gmg_oracle.sweep(oracle.JACOBI, N, W, ALPHA, level, u0.copy(), b0), ops[f"jacobi_l{level}"])
        ss, res = gmg_oracle.residual(N, W, ALPHA, level, u0, b0)
        assert np.array_equal(res, ops[f"res_l{level}"])
        assert abs(ss - ops[f"res_l{level}_sumsq_rel"][0]) <= 1e-13 * ss
    for lc in range(1, 4):
        assert np.array_equal(gmg_oracle.prolong(N, W, ALPHA, lc, u0.copy()), ops[f"prolong_from_l{lc}"])
    b1 = gmg_oracle.rhs(N, W, 1)
    assert np.array_equal(b1, ops["rhs_test1_N33"])
    for sm in (0, 1):
        u, _ = gmg_oracle.cycle(N, W, ALPHA, 4, sm, b1, u0.copy())
        assert np.array_equal(u, ops[f"cycle_smt{sm}"])


@pytest.mark.parametrize("name,n,L,sm,test", [("c1_257_gs", 257, 8, 0, 1), ("n65_jacobi", 65, 5, 1, 1),
                                              ("n65_gs_test2", 65, 6, 0, 2)])
def test_stored_reference_solves(gmg_oracle, goldens, name, n, L, sm, test):
    _, ops = goldens
    b = gmg_oracle.rhs(n, W, test)
    u, hist, crel, _ = gmg_oracle.solve(n, W, ALPHA, L, sm, b)
    assert np.array_equal(hist, ops[f"{name}_hist"])
    assert np.array_equal(u, ops[f"{name}_u"])
    assert np.allclose(crel, ops[f"{name}_coarse_relres"], rtol=2e-6)   # parsed from the printed line


def test_live_reference_bit_exact(gmg_oracle, gmg_ref):
    rng = np.random.default_rng(7)
    N = 65
    u0, b0 = rng.standard_normal(N * N), rng.standard_normal(N * N)
    for level in (0, 2, 5):
        for kind in (oracle.GS, oracle.JACOBI):
            assert np.array_equal(gmg_oracle.sweep(kind, N, W, 3.5, level, u0.copy(), b0),
                                  gmg_ref.sweep(kind, N, W, 3.5, level, u0.copy(), b0))
        ss, res = gmg_oracle.residual(N, W, 3.5, level, u0, b0)
        ssr, resr, _ = gmg_ref.residual(N, W, 3.5, level, u0, b0)
        assert np.array_equal(res, resr) and abs(ss - ssr) <= 1e-13 * ss
    for t in (0, 1, 2, 7):
        assert np.array_equal(gmg_oracle.rhs(N, 2.5, t), gmg_ref.rhs(N, 2.5, t))
    b = gmg_oracle.rhs(129, W, 2)
    u, h, _, _ = gmg_oracle.solve(129, W, 2.0, 4, oracle.GS, b)
    ur, hr, _ = gmg_ref.solve(129, W, 2.0, 4, 0, b)
    assert np.array_equal(u, ur) and np.array_equal(h, hr)


def test_rejects_bad_levels(gmg_oracle):
    with pytest.raises(ValueError):
        gmg_oracle.level(200, W, ALPHA, 1)      # the reference's default N=200 violates (N-1)%2^l==0
    with pytest.raises(ValueError):
        gmg_oracle.solve(33, W, ALPHA, 7, 0, gmg_oracle.rhs(33, W, 1))


@pytest.mark.parametrize("mode", [1, 2])
def test_rbgs_converges_to_same_solution(gmg_oracle, mode):
    """Reordered (red-black) GS reaches the lexicographic-GS converged solution (north_star: <=1e-8
    relative L2, cycle count within +-1).  Red-black GS leaves a residual that is 2x the smooth
    residual on the red points and 0 on the black ones, so plain injection (the reference's
    restriction) over-corrects by 2x; the reordered smoother is therefore paired with half
    injection (mode 1) or full weighting (mode 2)."""
    n, L = 129, 7
    b = gmg_oracle.rhs(n, W, 1)
    u_lex, h_lex, _, _ = gmg_oracle.solve(n, W, ALPHA, L, oracle.GS, b)
    u_rb, h_rb, _, _ = gmg_oracle.solve(n, W, ALPHA, L, oracle.RBGS, b, pre_kind=oracle.RBGS,
                                        restrict_mode=mode)
    rel = np.linalg.norm(u_rb - u_lex) / np.linalg.norm(u_lex)
    assert rel <= 1e-8, rel
    assert abs(h_rb.size - h_lex.size) <= 1, (h_rb.size, h_lex.size)


def test_rbgs_with_plain_injection_stalls(gmg_oracle):
    """documents WHY the reordered smoother needs a different restriction (see above)"""
    n, L = 65, 6
    b = gmg_oracle.rhs(n, W, 1)
    _, h_lex, _, _ = gmg_oracle.solve(n, W, ALPHA, L, oracle.GS, b)
    _, h_rb, _, _ = gmg_oracle.solve(n, W, ALPHA, L, oracle.RBGS, b, pre_kind=oracle.RBGS, maxiter=60)
    assert h_rb.size > 4 * h_lex.size
