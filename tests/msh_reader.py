"""Minimal Gmsh 4.1 ASCII reader for the tests, restating what the reference's reader keeps (AMG/src/FEM.cpp:30-303):
node coordinates in file order, nodes of line elements (type 1) flagged as boundary, triangles (type 2) with their
vertices sorted ascending."""
import numpy as np


def read_msh(path):
    tok = open(path).read().split()
    i = tok.index("$Nodes") + 1
    n_blocks, n_nodes = int(tok[i]), int(tok[i + 1])
    i += 4
    xy = np.zeros((n_nodes, 2))
    for _ in range(n_blocks):
        parametric, nb = int(tok[i + 2]), int(tok[i + 3])
        i += 4
        tags = [int(t) for t in tok[i:i + nb]]
        i += nb
        for t in tags:
            xy[t - 1] = (float(tok[i]), float(tok[i + 1]))
            i += 3 + (parametric and 0)
    i = tok.index("$Elements") + 1
    n_blocks = int(tok[i])
    i += 4
    bnd = np.zeros(n_nodes, np.uint8)
    tri = []
    for _ in range(n_blocks):
        etype, nb = int(tok[i + 2]), int(tok[i + 3])
        i += 4
        for _ in range(nb):
            if etype == 1:
                bnd[int(tok[i + 1]) - 1] = 1; bnd[int(tok[i + 2]) - 1] = 1
                i += 3
            elif etype == 2:
                tri.append(sorted(int(t) - 1 for t in tok[i + 1:i + 4]))
                i += 4
            elif etype == 15:
                i += 2
            else:
                raise ValueError(f"element type {etype}")
    return xy[:, 0].copy(), xy[:, 1].copy(), bnd, np.array(tri, np.int64)
