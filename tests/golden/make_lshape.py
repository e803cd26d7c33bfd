#!/usr/bin/env python
"""Builds BASELINE config 5 (the reference's examples/benchmark on data/lshape.msh) as a small fixture:

  A = kappa^2 M + K       P1 finite elements on the triangles of the mesh, natural boundary conditions
                          (src/ms.c:86-164: f0 = kappa^2 u, f1 = grad u, g0 = kappa^2, g3 = I), kappa = 5 (examples/benchmark/benchmarkrc:3)
  B[:, i] = M u_i         u_i = nodal interpolant of 1_{|x - c_i| < r_i} / (pi r_i^2)   (MakeObservationMats, src/obs.c:39-68, :135-180)
  S = 1 / sigma^2         sigma^2 = 1e-5, 17 observations (examples/benchmark/lshape.opts:4-8)
  f = B (S * obs_values)  the right-hand side that gives the posterior mean (src/obs.c:166-168)
  meas = M q              q = nodal interpolant of the indicator of the ball (1, 1), r = 0.8 (lshape.opts:11-13), normalised by its area

The mesh is read HERE from /root/reference/data/lshape.msh (Gmsh 4.1 ASCII); only the assembled arrays are committed
(tests/golden/lshape_config5.npz), in the mesh's node order (PETSc's DMPlex renumbers the vertices; the operator is the same up
to that permutation, so sample-by-sample parity is against the oracle on THESE arrays, and the reference's own acceptance test --
the posterior mean -- is permutation invariant).  Usage: python tests/golden/make_lshape.py [mesh] [refinements] [out.npz]; committed: refinements 0 and 1"""
import sys

import numpy as np
import scipy.sparse as sp


def read_gmsh41(path):
    """Nodes and 3-node triangles of a Gmsh 4.1 ASCII file."""
    lines = open(path).read().split("\n")
    i = lines.index("$Nodes") + 1
    nblocks, nnodes, _, _ = map(int, lines[i].split())
    i += 1
    tags, xyz = [], []
    for _ in range(nblocks):
        _, _, parametric, nb = map(int, lines[i].split())
        assert parametric == 0
        i += 1
        tags += [int(lines[i + q]) for q in range(nb)]
        i += nb
        xyz += [[float(v) for v in lines[i + q].split()] for q in range(nb)]
        i += nb
    assert len(tags) == nnodes
    tag2idx = {t: q for q, t in enumerate(tags)}
    i = lines.index("$Elements") + 1
    nblocks, _, _, _ = map(int, lines[i].split())
    i += 1
    tris = []
    for _ in range(nblocks):
        _, _, etype, nb = map(int, lines[i].split())
        i += 1
        if etype == 2:  # 3-node triangle
            tris += [[tag2idx[int(v)] for v in lines[i + q].split()[1:4]] for q in range(nb)]
        i += nb
    return np.asarray(xyz)[:, :2], np.asarray(tris, dtype=np.int64)


def assemble_p1(xy, tris):
    """Consistent mass matrix and stiffness matrix of P1 elements (exact integration)."""
    n = xy.shape[0]
    rows, cols, mv, kv = [], [], [], []
    mloc = (np.ones((3, 3)) + np.eye(3)) / 12.0
    for t in tris:
        p = xy[t]
        d1, d2 = p[1] - p[0], p[2] - p[0]
        det = d1[0] * d2[1] - d1[1] * d2[0]
        area = 0.5 * abs(det)
        g = np.array([[p[1][1] - p[2][1], p[2][1] - p[0][1], p[0][1] - p[1][1]], [p[2][0] - p[1][0], p[0][0] - p[2][0], p[1][0] - p[0][0]]]) / det  # gradients of the hat functions
        kloc = area * (g.T @ g)
        for a in range(3):
            for b in range(3):
                rows.append(t[a]); cols.append(t[b]); mv.append(area * mloc[a, b]); kv.append(kloc[a, b])
    M = sp.csr_matrix((mv, (rows, cols)), shape=(n, n))
    K = sp.csr_matrix((kv, (rows, cols)), shape=(n, n))
    M.sum_duplicates(); K.sum_duplicates()
    return M, K


def refine(xy, tris):
    """One regular refinement (every triangle into four through its edge midpoints), like -dm_refine 1."""
    mid, pts, out = {}, [p for p in xy], []

    def m(a, b):
        key = (min(a, b), max(a, b))
        if key not in mid:
            mid[key] = len(pts)
            pts.append(0.5 * (xy[a] + xy[b]))
        return mid[key]

    for a, b, c in tris:
        ab, bc, ca = m(a, b), m(b, c), m(c, a)
        out += [[a, ab, ca], [ab, b, bc], [ca, bc, c], [ab, bc, ca]]
    return np.asarray(pts), np.asarray(out, dtype=np.int64)


def main():
    mesh = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data/lshape.msh"
    nref = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    out = sys.argv[3] if len(sys.argv) > 3 else __file__.rsplit("/", 1)[0] + f"/lshape_config5_r{nref}.npz"
    kappa, sigma2 = 5.0, 1e-5
    c = np.array([0.2, 1.8, 0.4, 1.8, 0.6, 1.8, 0.8, 1.8, 0.2, 1.6, 0.4, 1.6, 0.6, 1.6, 0.8, 1.6, 0.2, 0.6, 0.4, 0.6, 0.6, 0.5, 0.8, 0.5, 1.0, 0.4, 1.2, 0.4, 1.4, 0.3, 1.6, 0.3, 1.8, 0.2]).reshape(17, 2)
    r = np.array([0.04] * 8 + [0.08] * 9)
    vals = np.array([0.5, -0.5, 0.5, -0.5, -0.5, 0.5, -0.5, 0.5, -0.5, -0.5, 0.5, 0.5, -0.5, -0.5, 0.5, 0.5, -0.5])
    xy, tris = read_gmsh41(mesh)
    for _ in range(nref):
        xy, tris = refine(xy, tris)
    M, K = assemble_p1(xy, tris)
    A = (kappa * kappa * M + K).tocsr()
    A.sort_indices()
    n = A.shape[0]
    B = np.zeros((n, 17))
    for i in range(17):
        u = (((xy - c[i]) ** 2).sum(1) < r[i] ** 2) / (np.pi * r[i] ** 2)  # src/obs.c:39-51
        B[:, i] = M @ u
    S = np.full(17, 1.0 / sigma2)
    f = B @ (S * vals)
    q = (((xy - np.array([1.0, 1.0])) ** 2).sum(1) < 0.8 ** 2) / (np.pi * 0.8 ** 2)
    meas = M @ q
    np.savez_compressed(out, rowptr=A.indptr.astype(np.int64), col=A.indices.astype(np.int32), val=A.data, B=B, S=S, f=f, meas=meas, xy=xy, kappa=kappa, sigma2=sigma2, obs_values=vals)
    print(f"{out}: n = {n}, nnz = {A.nnz}, triangles = {len(tris)}, observations hit = {(np.abs(B).sum(0) > 0).sum()} of 17, area = {M.sum():.6f}")


if __name__ == "__main__":
    main()
