"""helpers shared by the AMG tests: load the stored reference outputs (tests/golden/amg_*.npz)"""
import os

import numpy as np

import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    d = dict(np.load(os.path.join(GOLDEN, f"amg_{name}.npz")))
    L = int(d["levels"][0])

    def csr(p):
        shp = d[p + "_shape"]
        return oracle.Csr(shp[0], shp[1], d[p + "_ptr"], d[p + "_col"], d[p + "_val"])
    return {"levels": L, "A": [csr(f"A{l}") for l in range(L)], "P": [csr(f"P{l}") for l in range(L - 1)],
            "rhs": [d[f"rhs{l}"] for l in range(L)], "x": d["x_after_pass"], "res": float(d["residual_after_pass"][0]),
            "res0": float(d["residual_before"][0])}
