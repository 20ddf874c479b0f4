"""Pins oracle/amg_oracle.c to the reference's AMG: stored outputs of the compiled reference
(tests/golden/amg_*.npz: hierarchy + one-pass solution on the bundled meshes, config C2 = mesh1) and,
when oracle/_ref is present, live comparison bit for bit."""
import os

import numpy as np
import pytest

import oracle
from amg_fixtures import load_case


@pytest.mark.parametrize("name", ["mesh2", "mesh_pipe", "mesh1"])
def test_hierarchy_and_pass_equal_stored_reference(name):
    c = load_case(name)
    o = oracle.amg()
    h = o.build(c["A"][0], c["rhs"][0], c["levels"], [-1] * (c["levels"] - 1))
    for l in range(c["levels"]):
        assert h.A(l).same_as(c["A"][l]), f"A{l}"
        assert np.array_equal(h.rhs(l), c["rhs"][l]), f"rhs{l}"
        if l < c["levels"] - 1:
            assert h.P(l).same_as(c["P"][l]), f"P{l}"
    x = np.zeros(c["A"][0].n_rows)
    res = h.apply(x)
    assert np.array_equal(x, c["x"])
    assert res == c["res"]
    assert res < 0.1 * c["res0"]          # mesh1: 25.47 -> 1.70


def test_live_reference(tmp_path):
    r = oracle.ref_amg()
    if r is None or not os.path.isdir("/root/reference"):
        pytest.skip("no compiled reference here")
    o = oracle.amg()
    A, b = r.assemble("/root/reference/AMG/mesh/mesh-corner.msh")
    for starts in ([-1, -1, -1], [0, 5, 3], [A.n_rows - 1, 1, 0]):
        r.build(A, b, 4, starts)
        h = o.build(A, b, 4, starts)
        for l in range(4):
            assert h.A(l).same_as(r.A(l)) and np.array_equal(h.rhs(l), r.rhs(l))
            if l < 3:
                assert h.P(l).same_as(r.P(l))
        xr, rr = r.apply()
        xo = np.zeros(A.n_rows)
        assert h.apply(xo) == rr and np.array_equal(xo, xr)
    rng = np.random.default_rng(0)
    x0 = rng.standard_normal(A.n_rows)
    assert np.array_equal(o.gs(A, b, x0.copy(), 3), r.gs(A, b, x0.copy(), 3))


def test_operators_against_scipy():
    """independent cross-check of the transfer operators and the residual"""
    c = load_case("mesh_pipe")
    o = oracle.amg()
    A, P, b = c["A"][0], c["P"][0], c["rhs"][0]
    rng = np.random.default_rng(1)
    x = rng.standard_normal(A.n_rows)
    nrm, r = o.residual(A, x, b)
    assert np.allclose(r, b - A.to_scipy() @ x, rtol=1e-13, atol=1e-13)
    assert np.isclose(nrm, np.linalg.norm(r))
    assert np.allclose(o.restrict(P, x), P.to_scipy().T @ x, rtol=1e-13, atol=1e-13)
    xc = rng.standard_normal(P.n_cols)
    assert np.allclose(o.prolong_add(P, xc, x.copy()), x + P.to_scipy() @ xc, rtol=1e-13, atol=1e-13)
    # Galerkin operator of the stored hierarchy: Ac = P^T A P up to rounding
    Ac = (P.to_scipy().T @ A.to_scipy() @ P.to_scipy()).toarray()
    assert np.allclose(c["A"][1].to_scipy().toarray(), Ac, rtol=1e-12, atol=1e-12)
