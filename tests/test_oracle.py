"""CPU tests that pin the oracle (oracle/) before anything is compared against it.

What pins it:
  * an independent numpy/scipy re-implementation of every formula (this file),
  * the reference's own deterministic identity (examples/ex5.c:60-70),
  * the exact stationarity identities for a sweep and for the whole MGMC iteration,
  * known-answer vectors for Philox4x32-10 (Random123 kat_vectors) and libc erand48 for rander48,
  * the reference's statistical acceptance check (examples/ex1.c:131-135) at reduced length.
"""
import ctypes
import ctypes.util

import numpy as np
import pytest
import scipy.sparse as sp

SEED = 20260625  # examples/ex13.py:38


def ref_laplace(dim, nx, ny, nz, kappa):
    """Independent construction of src/problems.c:14-75 semantics with scipy kron."""
    h = 1.0 / ((nx - 1) * (nx - 1))

    def path(n):
        return sp.diags([np.ones(n - 1), np.ones(n - 1)], [-1, 1]) if n > 1 else sp.csr_matrix((1, 1))

    if dim == 2:
        G = sp.kron(sp.identity(ny), path(nx)) + sp.kron(path(ny), sp.identity(nx))
    else:
        G = (sp.kron(sp.identity(nz), sp.kron(sp.identity(ny), path(nx))) + sp.kron(sp.identity(nz), sp.kron(path(ny), sp.identity(nx)))
             + sp.kron(path(nz), sp.identity(nx * ny)))
    G = sp.csr_matrix(G)
    deg = np.asarray(G.sum(axis=1)).ravel()
    return sp.csr_matrix(sp.diags(kappa * kappa + deg * h) - h * G)


@pytest.mark.parametrize("dim,dims", [(2, (9, 9, 1)), (2, (7, 4, 1)), (3, (5, 4, 3)), (2, (129, 129, 1))])
def test_laplace_matches_independent_construction(orc, dim, dims):
    A = orc.laplace(dim, *dims, kappa=10.0)
    R = ref_laplace(dim, *dims, 10.0)
    S = A.to_scipy()
    assert S.shape == R.shape and A.nnz == R.nnz
    assert abs(S - R).max() < 1e-15 * 100
    assert S.has_sorted_indices or np.all(np.diff(S.indices[S.indptr[3]:S.indptr[4]]) > 0)
    # lambda_min = kappa^2 on a small case
    if A.n <= 100:
        assert abs(np.linalg.eigvalsh(S.toarray()).min() - 100.0) < 1e-10


def test_laplace_diag_is_repeated_addition(orc):
    """diag = kappa^2 + hinv2 + ... (one += per neighbour, src/problems.c:31-58), bit-exact."""
    A = orc.laplace(2, 9, 9, kappa=1.0)
    h = 1.0 / 64.0
    d = A.val[A.diag_ptrs()].reshape(9, 9)
    corner, edge, inner = 1.0, 1.0, 1.0
    for _ in range(2):
        corner += h
    for _ in range(3):
        edge += h
    for _ in range(4):
        inner += h
    assert d[0, 0] == corner and d[0, 4] == edge and d[4, 4] == inner


def numpy_sweep(S, omega, order, b, y):
    """Plain-python row update of src/mc_sor.c:260-268 (no fma; 1e-14 agreement expected)."""
    y = y.copy()
    indptr, idx, val = S.indptr, S.indices, S.data
    for r in order:
        s, d = b[r], 0.0
        for k in range(indptr[r], indptr[r + 1]):
            if idx[k] == r:
                d = val[k]
            else:
                s -= val[k] * y[idx[k]]
        y[r] = (1 - omega) * y[r] + omega / d * s
    return y


@pytest.mark.parametrize("omega", [1.0, 1.2, 1.6])
@pytest.mark.parametrize("coloring", ["single", "redblack", "greedy"])
def test_sweep_matches_numpy(orc, omega, coloring):
    rng = np.random.default_rng(SEED)
    A = orc.laplace(2, 9, 7, kappa=1.0)
    col = {"single": orc.Coloring.single(A.n), "redblack": orc.Coloring.parity((9, 7)), "greedy": orc.Coloring.greedy(A)}[coloring]
    assert col.violations(A) == 0 or coloring == "single"
    mc = orc.MCSOR(A, col, omega)
    b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
    S = A.to_scipy()
    fwd_order = np.concatenate([col.rows[col.ptr[c]:col.ptr[c + 1]] for c in range(col.ncolors)])
    y = mc.apply(b, y0.copy(), orc.SOR_FORWARD)
    np.testing.assert_allclose(y, numpy_sweep(S, omega, fwd_order, b, y0), rtol=0, atol=1e-13)
    y = mc.apply(b, y0.copy(), orc.SOR_BACKWARD)
    np.testing.assert_allclose(y, numpy_sweep(S, omega, fwd_order[::-1], b, y0), rtol=0, atol=1e-13)


def test_ex5_identity_symmetric_is_forward_then_backward(orc):
    """examples/ex5.c:60-70: ||(fwd;bwd)(x) - sym(x)|| < 1e-15 on 9x9, kappa = 1."""
    rng = np.random.default_rng(SEED)
    A = orc.laplace(2, 9, 9, kappa=1.0)
    for col in (orc.Coloring.single(A.n), orc.Coloring.parity((9, 9))):
        mc = orc.MCSOR(A, col)
        b, x = rng.random(A.n), rng.random(A.n)
        y = x.copy()
        mc.apply(b, x, orc.SOR_FORWARD)
        mc.apply(b, x, orc.SOR_BACKWARD)
        mc.apply(b, y, orc.SOR_SYMMETRIC)
        assert np.linalg.norm(x - y) < 1e-15


def test_levelset_coloring_reproduces_lexicographic_sweep_bitwise(orc):
    rng = np.random.default_rng(SEED)
    for A, dims in ((orc.laplace(2, 13, 9, kappa=1.0), (13, 9)), (orc.laplace(3, 5, 4, 6, kappa=2.0), (5, 4, 6))):
        ls = orc.Coloring.levelset(A)
        assert ls.violations(A) == 0
        assert ls.ncolors == sum(dims) - len(dims) + 1  # wavefronts i+j(+k) = const
        b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
        for sweep in (orc.SOR_FORWARD, orc.SOR_BACKWARD, orc.SOR_SYMMETRIC):
            y1 = orc.MCSOR(A, None, 1.3).apply(b, y0.copy(), sweep)
            y2 = orc.MCSOR(A, ls, 1.3).apply(b, y0.copy(), sweep)
            assert np.array_equal(y1, y2)


@pytest.mark.parametrize("omega", [1.0, 0.7, 1.5])
@pytest.mark.parametrize("sweep", ["fwd", "bwd", "sym"])
def test_sweep_stationarity_identity(orc, omega, sweep):
    """G Sigma G^T + (noise map)(noise map)^T = Sigma = A^-1 for the sampler of
    src/pc_mcgibbs.c:155-188, extracted column by column from the oracle itself."""
    A = orc.laplace(2, 5, 4, kappa=1.0)
    n = A.n
    col = orc.Coloring.parity((5, 4))
    sw = {"fwd": orc.SOR_FORWARD, "bwd": orc.SOR_BACKWARD, "sym": orc.SOR_SYMMETRIC}[sweep]
    ndraw = 2 * n if sweep == "sym" else n

    def step(y, z, b):
        return orc.gibbs_richardson(A, b, y.copy(), 1, orc.Noise.tape(z), col, omega, sw)

    zero, zz = np.zeros(n), np.zeros(ndraw)
    G = np.column_stack([step(e, zz, zero) for e in np.eye(n)])
    N = np.column_stack([step(zero, e, zero) for e in np.eye(ndraw)])
    Sigma = np.linalg.inv(A.to_scipy().toarray())
    np.testing.assert_allclose(G @ Sigma @ G.T + N @ N.T, Sigma, rtol=0, atol=5e-15 * abs(Sigma).max() * 10)
    # mean is a fixed point
    b = np.arange(1, n + 1, dtype=float)
    mu = Sigma @ b
    np.testing.assert_allclose(step(mu, zz, b), mu, rtol=1e-13)


def test_partitioned_sweep_equals_sequential(orc):
    """MCSORApply_MPIAIJ (src/mc_sor.c:298-381) == MCSORApply_SEQAIJ for a global colouring."""
    rng = np.random.default_rng(SEED)
    A = orc.laplace(2, 17, 12, kappa=1.0)
    col = orc.Coloring.parity((17, 12))
    b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
    for omega in (1.0, 1.3):
        for sweep in (orc.SOR_FORWARD, orc.SOR_BACKWARD, orc.SOR_SYMMETRIC):
            ref = orc.MCSOR(A, col, omega).apply(b, y0.copy(), sweep)
            for starts in ([0, A.n], [0, 17 * 6, A.n], [0, 50, 101, 150, A.n]):
                P = orc.Partitioned(A, starts, col, omega)
                for nt in (1, len(starts) - 1):
                    y = P.sweep(b, y0.copy(), sweep, nthreads=nt)
                    np.testing.assert_allclose(y, ref, rtol=0, atol=2e-15)


def test_ghost_lists_follow_matcreatescatters_order(orc):
    """src/mc_sor.c:191-205: per colour, colour-row order, one slot per off-diagonal-block nonzero."""
    A = orc.laplace(2, 4, 4, kappa=1.0)
    col = orc.Coloring.parity((4, 4))
    P = orc.Partitioned(A, [0, 8, 16], col)
    S = A.to_scipy()
    for rank, (r0, r1) in enumerate(((0, 8), (8, 16))):
        for c in range(2):
            exp = []
            for r in range(r0, r1):
                if col.color[r] != c:
                    continue
                for k in range(S.indptr[r], S.indptr[r + 1]):
                    if not (r0 <= S.indices[k] < r1):
                        exp.append(S.indices[k])
            assert list(P.ghost_index(rank, c)) == exp


# ---- RNG -------------------------------------------------------------------------------------
def test_philox_known_answers(orc):
    """Random123 kat_vectors, philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
           ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
           ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]
    for ctr, key, exp in kat:
        out = np.zeros(4, np.uint32)
        orc.lib().orc_philox4x32_10(np.array(ctr, np.uint32), np.array(key, np.uint32), out)
        assert tuple(int(x) for x in out) == exp


def test_philox_normals_partition_independent_and_standard(orc):
    z = orc.normal_philox(0xCAFE, 3, 0, 200001)
    parts = np.concatenate([orc.normal_philox(0xCAFE, 3, a, b - a) for a, b in ((0, 1), (1, 4), (4, 99999), (99999, 200001))])
    assert np.array_equal(z, parts)
    assert abs(z.mean()) < 4 / np.sqrt(z.size) and abs(z.var() - 1) < 0.02
    assert abs(np.mean(z ** 4) - 3) < 0.1
    assert abs(np.corrcoef(z[0::2][:100000], z[1::2][:100000])[0, 1]) < 0.02
    z2 = orc.normal_philox(0xCAFE, 4, 0, 1000)
    assert not np.allclose(z[:1000], z2)


def test_rander48_matches_libc_erand48_and_box_muller(orc):
    """PETSc rander48 == srand48/erand48; Box-Muller pairs of src/parmgmc.c:100-110."""
    libc = ctypes.CDLL(ctypes.util.find_library("c"))
    libc.drand48.restype = ctypes.c_double
    seed = 0x12345678
    libc.srand48(seed)
    u = [libc.drand48() for _ in range(10)]
    z = orc.noise_fill(orc.Noise.rander48(seed), 9)
    for i in range(0, 9, 2):
        r, th = np.sqrt(-2 * np.log(u[i])), 2 * np.pi * u[i + 1]
        assert z[i] == pytest.approx(r * np.cos(th), abs=1e-15)
        if i + 1 < 9:
            assert z[i + 1] == pytest.approx(r * np.sin(th), abs=1e-15)


def test_noise_tape_runs_dry(orc):
    ns = orc.Noise.tape(np.zeros(5))
    orc.noise_fill(ns, 5)
    with pytest.raises(RuntimeError):
        orc.noise_fill(ns, 1)


# ---- Cholesky sampler ------------------------------------------------------------------------
def test_chol_sampler_matches_numpy(orc):
    rng = np.random.default_rng(SEED)
    A = orc.laplace(2, 5, 5, kappa=1.0).to_scipy().toarray()
    n = A.shape[0]
    Lf = orc.potrf_lower(A)
    L = Lf.reshape(n, n).T  # column-major -> matrix
    np.testing.assert_allclose(np.tril(L), np.linalg.cholesky(A), atol=1e-14)
    b, z = rng.standard_normal(n), rng.standard_normal(n)
    y = orc.chol_sample(Lf, n, orc.Noise.tape(z), b)
    Lc = np.linalg.cholesky(A)
    np.testing.assert_allclose(y, np.linalg.solve(Lc.T, np.linalg.solve(Lc, b) + z), rtol=1e-12)
    with pytest.raises(np.linalg.LinAlgError):
        orc.potrf_lower(-np.eye(3))


# ---- multigrid -------------------------------------------------------------------------------
def q1_dense(nf):
    nc = (nf + 1) // 2
    P = np.zeros((nf, nc))
    for i in range(nf):
        if i % 2 == 0:
            P[i, i // 2] = 1
        else:
            P[i, (i - 1) // 2] = 0.5
            if (i + 1) // 2 < nc:
                P[i, (i + 1) // 2] = 0.5
    return P


@pytest.mark.parametrize("dim,dims", [(2, (9, 9, 1)), (2, (9, 5, 1)), (3, (5, 5, 5)), (2, (8, 6, 1))])
def test_galerkin_hierarchy_matches_dense(orc, dim, dims):
    mg = orc.MG.geometric(dim, *dims, 1.0, 3)
    A = ref_laplace(dim, *dims, 1.0).toarray()
    d = list(dims)
    for l in (2, 1):
        assert mg.level_dims(l) == tuple(d)
        np.testing.assert_allclose(mg.level_csr(l).to_scipy().toarray(), A, atol=1e-15)
        P = q1_dense(d[0])
        if d[1] > 1:
            P = np.kron(q1_dense(d[1]), P)
        if d[2] > 1:
            P = np.kron(q1_dense(d[2]), P)
        A = P.T @ A @ P
        d = [(x + 1) // 2 if x > 1 else 1 for x in d]
    np.testing.assert_allclose(mg.level_csr(0).to_scipy().toarray(), A, atol=1e-15)
    # 5-point -> 9-point on interior coarse nodes
    if dims == (9, 9, 1):
        S = mg.level_csr(1).to_scipy()
        assert S.getrow(12).nnz == 9


def mg_tape_len(mg, its_levels=1, its_coarse=1, sym=False):
    L = mg.nlevels
    ns = [mg.level_csr(l).n for l in range(L)]
    return 2 * its_levels * (2 if sym else 1) * sum(ns[1:]) + its_coarse * ns[0]


@pytest.mark.parametrize("smoother", ["sorgibbs", "mcgibbs_sym", "mcgibbs_coarse"])
def test_mgmc_invariance_identity(orc, smoother):
    """the restated MGMC iteration y' = K y + M b + eps leaves
    N(A^-1 b, A^-1) invariant: K A^-1 K^T + N N^T = A^-1, on the 9 -> 5 -> 3 hierarchy of ex1.c:41."""
    mg = orc.MG.geometric(2, 9, 9, 1, 1.0, 3)
    if smoother == "mcgibbs_sym":
        for l in (1, 2):
            mg.set_smoother(l, orc.KIND_MCGIBBS, 1.3, orc.SOR_SYMMETRIC, 1, orc.Coloring.greedy(mg.level_csr(l)))
    if smoother == "mcgibbs_coarse":  # ex1.c:41: mcgibbs everywhere, 2 its per level
        for l in (0, 1, 2):
            mg.set_smoother(l, orc.KIND_MCGIBBS, 1.0, orc.SOR_FORWARD, 2, None)
    mg.setup()
    n = 81
    ns = [mg.level_csr(l).n for l in range(3)]
    if smoother == "sorgibbs":
        T = 2 * (ns[1] + ns[2]) + ns[0]
    elif smoother == "mcgibbs_sym":
        T = 4 * (ns[1] + ns[2]) + ns[0]
    else:
        T = 4 * (ns[1] + ns[2]) + 2 * ns[0]

    def step(y, z, b):
        noise = orc.Noise.tape(z)
        out = mg.richardson(noise, b, y.copy(), 1, guesszero=False)
        assert noise.tape_pos == T  # the noise-tape contract of SURVEY 8(c)
        return out

    zero, zz = np.zeros(n), np.zeros(T)
    K = np.column_stack([step(e, zz, zero) for e in np.eye(n)])
    N = np.column_stack([step(zero, e, zero) for e in np.eye(T)])
    Sigma = np.linalg.inv(ref_laplace(2, 9, 9, 1, 1.0).toarray())
    np.testing.assert_allclose(K @ Sigma @ K.T + N @ N.T, Sigma, rtol=0, atol=1e-14 * abs(Sigma).max())
    b = np.linspace(1, 2, n)
    mu = Sigma @ b
    np.testing.assert_allclose(step(mu, zz, b), mu, rtol=1e-12)
    # V-cycle contracts much faster than one Gibbs sweep (sanity of the coarse correction)
    assert np.abs(np.linalg.eigvals(K)).max() < 0.2


def test_mg_guesszero_first_iteration_is_plain_cycle(orc):
    """src/pc_gamgmc.c:243-246."""
    mg = orc.MG.geometric(2, 9, 9, 1, 1.0, 3)
    mg.setup()
    rng = np.random.default_rng(SEED)
    b, z = rng.standard_normal(81), rng.standard_normal(2 * 300)
    y1 = mg.richardson(orc.Noise.tape(z), b, np.full(81, 7.0), 1, guesszero=True)
    y2 = mg.apply(orc.Noise.tape(z), b, np.empty(81))
    assert np.array_equal(y1, y2)


# ---- estimators --------------------------------------------------------------------------------
def test_autocorrelation_and_iact(orc):
    rng = np.random.default_rng(SEED)
    rho, n = 0.8, 200000
    x = np.empty(n)
    x[0] = 0
    e = rng.standard_normal(n)
    for i in range(1, n):
        x[i] = rho * x[i - 1] + e[i]
    acf = orc.autocorrelation(x[:4096])
    xc = x[:4096] - x[:4096].mean()
    direct = np.array([np.dot(xc[:4096 - k], xc[k:]) for k in range(50)]) / np.dot(xc, xc)
    np.testing.assert_allclose(acf[:50], direct, atol=1e-10)
    tau, valid = orc.iact(x)
    assert valid and abs(tau - (1 + rho) / (1 - rho)) < 0.6
    tau_w, _ = orc.iact(rng.standard_normal(50000))
    assert abs(tau_w - 1) < 0.1
    with pytest.raises(ValueError):
        orc.iact(np.zeros(1))


def test_cov_errors_and_gelman_rubin(orc):
    rng = np.random.default_rng(SEED)
    A = ref_laplace(2, 3, 3, 1, 1.0).toarray()
    L = np.linalg.cholesky(np.linalg.inv(A))
    chains, S = 4000, 3
    samples = np.einsum("ik,sck->sci", L, rng.standard_normal((S, chains, 9)))
    errs = orc.cov_errors(A, samples)
    Q = np.linalg.inv(A)
    for s in range(S):
        Cs = np.cov(samples[s].T, ddof=1)
        assert errs[s] == pytest.approx(np.linalg.norm(Cs - Q) / np.linalg.norm(Q), rel=1e-10)
    assert errs.max() < 0.1
    vals = rng.standard_normal((8, 5000))
    assert abs(orc.gelman_rubin(vals) - 1) < 0.01
    assert orc.gelman_rubin(vals + np.arange(8)[:, None]) > 2


# ---- the reference's statistical acceptance check ------------------------------------------------
@pytest.mark.parametrize("case", ["mcgibbs", "mcgibbs_sym", "gamgmc"])
def test_ex1_mean_convergence(orc, case):
    """examples/ex1.c:83-135: 9x9, kappa = 10, b = 1; rel. error of the running sample mean <= 0.02."""
    A = orc.laplace(2, 9, 9, kappa=10.0)
    b, y = np.ones(81), np.zeros(81)
    ex_mean = np.linalg.solve(A.to_scipy().toarray(), b)
    acc = {"mean": np.zeros(81)}

    def cb(it, yy):
        acc["mean"] = acc["mean"] * (it / (it + 1.0)) + yy / (it + 1.0)  # VecAXPBY of ex1.c:60

    noise = orc.Noise.rander48()
    if case == "gamgmc":
        mg = orc.MG.geometric(2, 9, 9, 1, 10.0, 3)
        mg.setup()
        mg.richardson(noise, b, y, 1000)
        mg.richardson(noise, b, y, 300000, callback=cb)
    else:
        sw = orc.SOR_SYMMETRIC if case == "mcgibbs_sym" else orc.SOR_FORWARD
        orc.gibbs_richardson(A, b, y, 10000, noise, None, 1.0, sw)
        orc.gibbs_richardson(A, b, y, 600000, noise, None, 1.0, sw, callback=cb)
    rel = np.linalg.norm(acc["mean"] - ex_mean) / np.linalg.norm(ex_mean)
    assert rel <= 0.02, rel


# ---- MATLRC operators: the restatement has no reference binary to be pinned against (the PETSc stub
#      has no MATLRC), so it is validated by the property the construction exists for -----------------------------------
@pytest.mark.parametrize("sweep", [1, 2, 3])
@pytest.mark.parametrize("omega", [1.0])
def test_lrc_sampler_leaves_posterior_invariant(orc, sweep, omega):
    """One LRC-Gibbs sample is an affine-Gaussian map y' = G y + N z; N(0, (A + B S B^T)^-1) must be exactly stationary:
    G Sigma G^T + N N^T = Sigma (src/mc_sor.c:456-479: 'enact the Sherman-Morrison-Woodbury correction')."""
    A = orc.laplace(2, 6, 5, kappa=1.0)
    n, k = A.n, 3
    rng = np.random.default_rng(3)
    B, S = rng.standard_normal((n, k)), rng.uniform(0.5, 2.0, k)
    col = orc.Coloring.parity((6, 5))
    per = (n + k) * (2 if sweep == orc.SOR_SYMMETRIC else 1)

    def sample(y0, z):
        return orc.lrc_gibbs_richardson(A, B, S, None, y0.copy(), 1, orc.Noise.tape(z), col, omega, sweep)

    G = np.column_stack([sample(np.eye(n)[:, i], np.zeros(per)) for i in range(n)])
    N = np.column_stack([sample(np.zeros(n), np.eye(per)[:, i]) for i in range(per)])
    Sigma = np.linalg.inv(A.to_scipy().toarray() + B @ np.diag(S) @ B.T)
    assert np.abs(G @ Sigma @ G.T + N @ N.T - Sigma).max() / np.abs(Sigma).max() < 1e-13


def test_lrc_mean_converges_to_posterior_mean(orc):
    """With a right-hand side the chain mean converges to (A + B S B^T)^-1 b (examples/benchmark: posterior sampling)."""
    A = orc.laplace(2, 9, 9, kappa=2.0)
    n, k = A.n, 4
    rng = np.random.default_rng(11)
    B, S = rng.standard_normal((n, k)) * 0.3, np.full(k, 5.0)
    b = rng.standard_normal(n)
    y = np.zeros(n)
    # deterministic part only (zero noise): the fixed point of the map is the posterior mean
    for _ in range(60):
        y = orc.lrc_gibbs_richardson(A, B, S, b, y, 1, orc.Noise.tape(np.zeros(2 * (n + k))), orc.Coloring.parity((9, 9)), 1.0, orc.SOR_SYMMETRIC)
    mean = np.linalg.solve(A.to_scipy().toarray() + B @ np.diag(S) @ B.T, b)
    assert np.abs(y - mean).max() / np.abs(mean).max() < 1e-10


def test_philox_grid_blocks_use_padded_index(orc):
    """Blocks of a matrix-free grid operator are keyed on (k ny + j) pitch + i, pitch = nx rounded up to 4 (philox.cuh)."""
    nx, ny = 13, 5
    z = orc.noise_fill(orc.Noise.philox(0xCAFE, grid=(nx, ny, 1)), nx * ny)
    pitch = 16
    full = orc.normal_philox(0xCAFE, 0, 0, pitch * ny).reshape(ny, pitch)
    assert np.array_equal(z.reshape(ny, nx), full[:, :nx])
    # a block of another size (a coarse level) keeps plain global rows
    ns = orc.Noise.philox(0xCAFE, grid=(nx, ny, 1))
    assert np.array_equal(orc.noise_fill(ns, 21), orc.normal_philox(0xCAFE, 0, 0, 21))
    # nx a multiple of 4: nothing changes
    assert np.array_equal(orc.noise_fill(orc.Noise.philox(7, grid=(8, 3, 1)), 24), orc.normal_philox(7, 0, 0, 24))


# ---- BASELINE config 5: P1 finite elements on data/lshape.msh + 17 low-rank ball observations (tests/golden/make_lshape.py) ----
def _lshape(orc, nref=0):
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"lshape_config5_r{nref}.npz"))
    A = orc.CSR(int(d["rowptr"].size - 1), d["rowptr"], d["col"], d["val"])
    return A, d


def test_config5_fixture_is_the_fe_operator_of_the_reference(orc):
    """kappa^2 M + K with natural boundary conditions (src/ms.c:86-164): symmetric positive definite, K 1 = 0 so that
    A 1 = kappa^2 M 1 and 1^T A 1 = kappa^2 |Omega| = 25 * 3; B = M u_i and f = B (S * values) as MakeObservationMats builds them."""
    for nref, n in ((0, 408), (1, 1549)):
        A, d = _lshape(orc, nref)
        assert A.n == n
        As = A.to_scipy()
        assert abs(As - As.T).max() < 1e-14
        one = np.ones(n)
        assert abs(one @ (As @ one) - 25.0 * 3.0) < 1e-10
        assert np.linalg.eigvalsh(As.toarray()).min() > 0
        B, S, f = d["B"], d["S"], d["f"]
        assert B.shape == (n, 17) and np.allclose(S, 1.0 / 1e-5, rtol=1e-15) and np.allclose(f, B @ (S * d["obs_values"]), rtol=1e-14)
        assert np.all(B >= 0) and (np.abs(B).sum(0) > 0).sum() == (13 if nref == 0 else 17)
        # an observation functional integrates the indicator / volume: close to 1 where the ball is resolved by the mesh
        if nref == 1:
            assert np.all(np.abs(B.sum(0)[8:] - 1.0) < 0.35)
        col = orc.Coloring.greedy(A)
        assert col.violations(A) == 0 and col.ncolors <= 12


@pytest.mark.parametrize("sweep", [1, 3])
def test_config5_lrc_gibbs_leaves_the_posterior_invariant(orc, sweep):
    """The LRC Gibbs sampler on the FE operator with the 17 observations: G Sigma G^T + N N^T = Sigma for
    Sigma = (A + B S B^T)^-1, and the deterministic part converges to the posterior mean (A + B S B^T)^-1 f."""
    A, d = _lshape(orc, 0)
    n, k = A.n, 17
    B, S, f = d["B"], d["S"], d["f"]
    col = orc.Coloring.greedy(A)
    per = (n + k) * (2 if sweep == orc.SOR_SYMMETRIC else 1)

    def sample(y0, z, b=None):
        return orc.lrc_gibbs_richardson(A, B, S, b, y0.copy(), 1, orc.Noise.tape(z), col, 1.0, sweep)

    G = np.column_stack([sample(np.eye(n)[:, i], np.zeros(per)) for i in range(n)])
    N = np.column_stack([sample(np.zeros(n), np.eye(per)[:, i]) for i in range(per)])
    P = A.to_scipy().toarray() + B @ np.diag(S) @ B.T
    Sigma = np.linalg.inv(P)
    assert np.abs(G @ Sigma @ G.T + N @ N.T - Sigma).max() / np.abs(Sigma).max() < 1e-9  # cond(P) ~ 1e6: sigma^2 = 1e-5
    mean = np.linalg.solve(P, f)
    c = sample(np.zeros(n), np.zeros(per), f)  # affine part: y' = G y + c, fixed point (I - G)^-1 c
    fix = np.linalg.solve(np.eye(n) - G, c)
    assert np.abs(fix - mean).max() / np.abs(mean).max() < 1e-7


def test_config5_woodbury_exact_sampler_identity(orc):
    """PCWOODBURY (src/woodbury.c:21-86, :259-286) with an exact sampler on A: y_s ~ N(A^-1 w, A^-1), w = f + B sqrt(S) eta, then
    y = y_s - G B^T y_s with G = C (S^-1 + B^T C)^-1, C = A^-1 B.  Mean and covariance of y must be those of the posterior
    N((A + B S B^T)^-1 f, (A + B S B^T)^-1): the identity the construction rests on, checked on the config-5 operator."""
    A, d = _lshape(orc, 0)
    n = A.n
    Ad, B, S, f = A.to_scipy().toarray(), d["B"], d["S"], d["f"]
    Ainv = np.linalg.inv(Ad)
    Cm = Ainv @ B
    G = Cm @ np.linalg.inv(np.diag(1.0 / S) + B.T @ Cm)
    T = np.eye(n) - G @ B.T
    cov_s = Ainv + Cm @ np.diag(S) @ Cm.T  # A^-1 (noise of the sampler) + A^-1 B S B^T A^-1 (the k extra draws)
    P = Ad + B @ np.diag(S) @ B.T
    Sigma = np.linalg.inv(P)
    assert np.abs(T @ cov_s @ T.T - Sigma).max() / np.abs(Sigma).max() < 1e-8
    assert np.abs(T @ (Ainv @ f) - Sigma @ f).max() / np.abs(Sigma @ f).max() < 1e-8
