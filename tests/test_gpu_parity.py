"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs.  Integer/index work (colourings) and injected-noise sweeps must be bit-exact; paths
whose operation order legitimately differs are held to a relative error of 1e-12 (the tolerance
BASELINE.json's north_star states for FP64)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 20260625  # examples/ex13.py:38
RTOL = 1e-12


@pytest.fixture(scope="module")
def pmg():
    import parmgmc_b200 as m
    if m.device_count() == 0:
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    return m


@pytest.fixture(scope="module")
def ctx(pmg):
    c = pmg.Context(0, seed=0xCAFE)
    yield c
    c.close()


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def make_mat(pmg, ctx, A, coloring=None, policy=None):
    m = pmg.Mat.from_csr(ctx, A.rowptr, A.col, A.val)
    if coloring is not None:
        m.set_coloring(coloring.color, coloring.ncolors)
    if policy is not None:
        m.set_coloring_auto(policy)
    return m


# ---- a6/a8: MCSORApply --------------------------------------------------------------------------
@pytest.mark.parametrize("kappa", [10.0, 1.0])
@pytest.mark.parametrize("omega", [1.0, 1.2, 1.6])
def test_mcsor_config1_redblack_bitexact(pmg, ctx, orc, kappa, omega):
    """Config 1 (129x129, src/problems.c semantics), same injected colouring on both sides."""
    rng = np.random.default_rng(SEED)
    A = orc.laplace(2, 129, 129, kappa=kappa)
    col = orc.Coloring.parity((129, 129))
    mat = make_mat(pmg, ctx, A, col)
    mc = pmg.MCSOR(mat)
    assert mc.get_num_colors() == 2
    mc.set_omega(omega)
    b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
    for sweep, osweep in ((pmg.SOR_FORWARD_SWEEP, orc.SOR_FORWARD), (pmg.SOR_BACKWARD_SWEEP, orc.SOR_BACKWARD), (pmg.SOR_SYMMETRIC_SWEEP, orc.SOR_SYMMETRIC)):
        mc.set_sweep_type(sweep)
        assert mc.get_sweep_type() == sweep
        y = mc.apply(b, y0.copy())
        ref = orc.MCSOR(A, col, omega).apply(b, y0.copy(), osweep)
        assert np.array_equal(y, ref), relerr(y, ref)


def test_mcsor_lexicographic_reproduces_one_rank_reference_bitexact(pmg, ctx, orc):
    """The 1-rank reference sweeps ONE colour in natural order (src/mc_sor.c:397-410).  The level-set
    colouring makes the device do exactly that order: compare with the oracle's one-colour sweep."""
    rng = np.random.default_rng(SEED)
    A = orc.laplace(2, 129, 129, kappa=10.0)
    mat = make_mat(pmg, ctx, A, policy=pmg.COLORING_LEXICOGRAPHIC)
    k, color = mat.get_coloring()
    assert k == 257
    assert np.array_equal(color, orc.Coloring.levelset(A).color)  # colouring matches bit-exactly
    mc = pmg.MCSOR(mat)
    mc.set_omega(1.2)
    b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
    for sweep, osweep in ((1, orc.SOR_FORWARD), (2, orc.SOR_BACKWARD), (3, orc.SOR_SYMMETRIC)):
        mc.set_sweep_type(sweep)
        y = mc.apply(b, y0.copy())
        ref = orc.MCSOR(A, None, 1.2).apply(b, y0.copy(), osweep)  # one colour, lexicographic
        assert np.array_equal(y, ref)


def test_greedy_coloring_matches_oracle_and_is_valid(pmg, ctx, orc):
    for A in (orc.laplace(2, 33, 17, kappa=1.0), orc.laplace(3, 9, 8, 7, kappa=1.0), orc.MG.geometric(2, 17, 17, 1, 1.0, 2).level_csr(0)):
        mat = make_mat(pmg, ctx, A, policy=pmg.COLORING_GREEDY)
        k, color = mat.get_coloring()
        ref = orc.Coloring.greedy(A)
        assert k == ref.ncolors and np.array_equal(color, ref.color)
        assert orc.Coloring(color, k).violations(A) == 0


def test_ex5_identity_on_device(pmg, ctx, orc):
    """examples/ex5.c:60-70."""
    rng = np.random.default_rng(SEED)
    A = orc.laplace(2, 9, 9, kappa=1.0)
    mat = make_mat(pmg, ctx, A, policy=pmg.COLORING_GREEDY)
    mc = pmg.MCSOR(mat)
    b, x = rng.random(81), rng.random(81)
    y = x.copy()
    mc.set_sweep_type(1); mc.apply(b, x)
    mc.set_sweep_type(2); mc.apply(b, x)
    mc.set_sweep_type(3); mc.apply(b, y)
    assert np.linalg.norm(x - y) < 1e-15


def test_mcsor_ragged_rows_and_3d(pmg, ctx, orc):
    """rows of very different length (SELL padding) and a 3D 7-point operator"""
    import scipy.sparse as sp
    rng = np.random.default_rng(SEED)
    n = 300
    M = sp.random(n, n, density=0.03, random_state=7, format="csr")
    M = M + M.T
    M = sp.csr_matrix(M + sp.diags(np.asarray(abs(M).sum(axis=1)).ravel() + 1.0))
    dense_row = np.zeros(n); dense_row[:] = 0.01
    M = sp.lil_matrix(M); M[5, :] = dense_row; M[:, 5] = dense_row.reshape(-1, 1); M[5, 5] = 10.0
    M = sp.csr_matrix(M); M.sort_indices()
    for A in (orc.CSR(n, M.indptr, M.indices, M.data), orc.laplace(3, 9, 8, 7, kappa=2.0)):
        col = orc.Coloring.greedy(A)
        mat = make_mat(pmg, ctx, A, col)
        mc = pmg.MCSOR(mat)
        mc.set_omega(0.9)
        mc.set_sweep_type(3)
        b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
        y = mc.apply(b, y0.copy())
        ref = orc.MCSOR(A, col, 0.9).apply(b, y0.copy(), orc.SOR_SYMMETRIC)
        assert np.array_equal(y, ref)
        np.testing.assert_allclose(mat.mult(y0), A.to_scipy() @ y0, rtol=1e-13, atol=1e-13)


def test_error_codes(pmg, ctx, orc):
    A = orc.laplace(2, 5, 5, kappa=1.0)
    mat = make_mat(pmg, ctx, A)
    mc = pmg.MCSOR(mat)
    with pytest.raises(pmg.PMGError) as e:
        mc.set_sweep_type(7)  # src/mc_sor.c:427 PETSC_ERR_SUP
    assert e.value.code == 2
    with pytest.raises(pmg.PMGError) as e:
        mat.set_coloring(np.zeros(25, np.int32), 1)  # one colour is not a distance-1 colouring
    assert e.value.code == 8
    with pytest.raises(pmg.PMGError) as e:
        pmg.PC(ctx, "parsor")
    assert e.value.code == 2
    pc = pmg.PC(ctx, "mcgibbs")
    with pytest.raises(pmg.PMGError) as e:
        pc.setup()
    assert e.value.code == 6
    pc.set_operator(mat)
    pc.set_option("-pc_mcgibbs_omega", 2.5)
    with pytest.raises(pmg.PMGError) as e:
        pc.setup()
    assert e.value.code == 1
    pc.set_option("-pc_mcgibbs_omega", 1.0)
    pc.setup()
    pc.set_noise_tape(np.zeros(10))
    with pytest.raises(pmg.PMGError) as e:
        pc.apply_richardson(np.zeros(25), np.zeros(25), its=1)
    assert e.value.code == 7
    bad = orc.laplace(2, 4, 4, kappa=1.0)
    bad.val[bad.diag_ptrs()] = -1.0
    chol = pmg.PC(ctx, "cholsampler")
    chol.set_operator(make_mat(pmg, ctx, bad))
    with pytest.raises(pmg.PMGError) as e:
        chol.setup()
    assert e.value.code == 5  # PETSC_ERR_MAT_CH_ZRPVT, src/pc_chols.c:192


def test_recolouring_and_operator_swap_after_setup_rebuild_the_coefficients(pmg, ctx, orc):
    """The per-row sweep coefficients (omega / a_ii, sqrt a_ii) are stored per padded sweep position of ONE matrix and colouring
    (src/mc_sor.c:114-124, src/pc_mcgibbs.c:142-153 keep them per row): re-colouring a matrix that already has a sampler, or
    handing the sampler another operator, must rebuild them -- PCSetUp called again without PCReset does exactly this."""
    rng = np.random.default_rng(SEED)
    A = orc.laplace(2, 41, 29, kappa=2.0)
    A.val[A.diag_ptrs()] *= 1.0 + rng.random(A.n)  # a non-constant diagonal: stale coefficients would show
    col1 = orc.Coloring.parity((41, 29))
    mat = make_mat(pmg, ctx, A, col1)
    pc = pmg.PC(ctx, "mcgibbs")
    pc.set_operator(mat)
    pc.set_options({"-pc_mcgibbs_omega": 1.3})
    pc.setup()
    mc = pmg.MCSOR(mat)
    mc.set_omega(1.3)
    b = rng.standard_normal(A.n)

    def check(Aref, colref, pcx, mcx):
        z = rng.standard_normal(2 * Aref.n)
        pcx.set_noise_tape(z)
        y = np.zeros(Aref.n)
        bb = b[:Aref.n] if Aref.n <= b.size else np.resize(b, Aref.n)
        pcx.apply_richardson(bb, y, its=2)
        ref = orc.gibbs_richardson(Aref, bb, np.zeros(Aref.n), 2, orc.Noise.tape(z), colref, 1.3, orc.SOR_FORWARD)
        assert np.array_equal(y, ref)
        if mcx is not None:
            ym = np.full(Aref.n, 0.25)
            mcx.apply(bb, ym)
            assert np.array_equal(ym, orc.MCSOR(Aref, colref, 1.3, orc.SOR_FORWARD).apply(bb, np.full(Aref.n, 0.25)))

    check(A, col1, pc, mc)
    # (1) re-colour the matrix under the live sampler and MCSOR (lexicographic level sets: a different, longer layout)
    col2 = orc.Coloring.levelset(A)
    mat.set_coloring(col2.color, col2.ncolors)
    pc.setup()
    check(A, col2, pc, mc)
    # (2) hand the same PC a larger operator
    A2 = orc.laplace(2, 57, 44, kappa=3.0)
    A2.val[A2.diag_ptrs()] *= 1.0 + np.random.default_rng(3).random(A2.n)
    colA2 = orc.Coloring.greedy(A2)
    mat2 = make_mat(pmg, ctx, A2, colA2)
    pc.set_operator(mat2)
    pc.setup()
    check(A2, colA2, pc, None)


# ---- a10-a13: the Gibbs samplers -----------------------------------------------------------------
@pytest.mark.parametrize("pctype,opts,osweep,omega", [
    ("mcgibbs", {}, 1, 1.0),
    ("mcgibbs", {"-pc_mcgibbs_omega": 1.6, "-pc_mcgibbs_backward": ""}, 2, 1.6),
    ("mcgibbs", {"-pc_mcgibbs_omega": 1.2, "-pc_mcgibbs_symmetric": ""}, 3, 1.2),
    ("sorgibbs", {}, 1, 1.0),
])
def test_gibbs_sampler_injected_noise_bitexact(pmg, ctx, orc, pctype, opts, osweep, omega):
    rng = np.random.default_rng(SEED)
    A = orc.laplace(2, 129, 129, kappa=10.0)
    col = orc.Coloring.parity((129, 129))
    mat = make_mat(pmg, ctx, A, col)
    pc = pmg.PC(ctx, pctype)
    pc.set_operator(mat)
    pc.set_options(opts)
    pc.setup()
    its = 4
    per = pc.noise_per_sample()
    assert per == (2 if osweep == 3 else 1) * A.n
    z = rng.standard_normal(its * per)
    pc.set_noise_tape(z)
    b, y = np.ones(A.n), np.zeros(A.n)
    seen = []
    pc.set_sample_callback(lambda it, yy: seen.append((it, yy.copy())))
    outits, reason = pc.apply_richardson(b, y, its=its)
    assert (outits, reason) == (its, pmg.PCRICHARDSON_CONVERGED_ITS)
    ref_seen = []
    ref = orc.gibbs_richardson(A, b, np.zeros(A.n), its, orc.Noise.tape(z), col, omega, osweep, callback=lambda it, yy: ref_seen.append((it, yy.copy())))
    assert np.array_equal(y, ref)
    assert [s[0] for s in seen] == list(range(its))
    for (_, a), (_, r) in zip(seen, ref_seen):
        assert np.array_equal(a, r)


def test_sorgibbs_pcapply_zeroes_guess(pmg, ctx, orc):
    """src/pc_sorgibbs.c:105-113."""
    rng = np.random.default_rng(SEED)
    A = orc.laplace(2, 17, 17, kappa=1.0)
    col = orc.Coloring.parity((17, 17))
    pc = pmg.PC(ctx, "sorgibbs")
    pc.set_operator(make_mat(pmg, ctx, A, col))
    pc.setup()
    z, b = rng.standard_normal(A.n), rng.standard_normal(A.n)
    pc.set_noise_tape(z)
    y = pc.apply(b)
    ref = orc.gibbs_richardson(A, b, np.zeros(A.n), 1, orc.Noise.tape(z), col, 1.0, orc.SOR_FORWARD)
    assert np.array_equal(y, ref)


def test_prior_sampling_null_rhs(pmg, ctx, orc):
    """b = NULL (examples/ex8.c:47-49 samples the prior with f = 0)."""
    rng = np.random.default_rng(SEED)
    A = orc.laplace(2, 17, 9, kappa=1.0)
    col = orc.Coloring.parity((17, 9))
    pc = pmg.PC(ctx, "mcgibbs")
    pc.set_operator(make_mat(pmg, ctx, A, col))
    pc.setup()
    z = rng.standard_normal(2 * A.n)
    pc.set_noise_tape(z)
    y = np.zeros(A.n)
    pc.apply_richardson(None, y, its=2)
    ref = orc.gibbs_richardson(A, np.zeros(A.n), np.zeros(A.n), 2, orc.Noise.tape(z), col)
    assert np.array_equal(y, ref)


def test_callback_deleter_and_setters(pmg, ctx, orc):
    A = orc.laplace(2, 9, 9, kappa=1.0)
    pc = pmg.PC(ctx, "mcgibbs")
    pc.set_operator(make_mat(pmg, ctx, A, orc.Coloring.parity((9, 9))))
    pc.setup()
    assert "Number of colours: 2" in pc.view()  # PCView_MulticolorGibbs
    deleted = []
    pc.set_sample_callback(lambda it, y: None, deleter=lambda: deleted.append(1))
    pc.set_sample_callback(lambda it, y: None)  # replacing runs the old deleter (src/pc_mcgibbs.c:295-298)
    assert deleted == [1]
    # a failing callback surfaces as an error (PetscCall propagation)
    pc.set_sample_callback(lambda it, y: 1)
    with pytest.raises(pmg.PMGError) as e:
        pc.apply_richardson(np.ones(81), np.zeros(81), its=1)
    assert e.value.code == 10
    # omega / sweep setters after set-up take effect lazily (omega_changed)
    pc.set_sample_callback(None)
    pc.mcgibbs_set_omega(1.5)
    pc.mcgibbs_set_sweep_type(pmg.SOR_BACKWARD_SWEEP)
    z = np.random.default_rng(1).standard_normal(81)
    pc.set_noise_tape(z)
    y = np.zeros(81)
    pc.apply_richardson(np.ones(81), y, its=1)
    ref = orc.gibbs_richardson(A, np.ones(81), np.zeros(81), 1, orc.Noise.tape(z), orc.Coloring.parity((9, 9)), 1.5, orc.SOR_BACKWARD)
    assert np.array_equal(y, ref)


# ---- a11: device normals ---------------------------------------------------------------------------
def test_device_philox_normals_match_definition(pmg, ctx, orc):
    for row0, n in ((0, 100001), (12345, 4097), (2 ** 33 + 1, 1000)):
        z = ctx.normal_fill(0xCAFE, 7, row0, n)
        ref = orc.normal_philox(0xCAFE, 7, row0, n)
        assert np.abs(z - ref).max() < 1e-12  # lean device Box-Muller (fastnormal.cuh) vs the libm definition
    z = ctx.normal_fill(1, 0, 0, 1 << 20)
    assert abs(z.mean()) < 5e-3 and abs(z.var() - 1) < 5e-3 and abs(np.mean(z ** 4) - 3) < 0.03


def test_gibbs_philox_mode_matches_oracle_philox(pmg, ctx, orc):
    A = orc.laplace(2, 65, 33, kappa=1.0)
    col = orc.Coloring.parity((65, 33))
    pc = pmg.PC(ctx, "mcgibbs")
    pc.set_operator(make_mat(pmg, ctx, A, col))
    pc.set_options({"-pc_mcgibbs_symmetric": "", "-pc_b200_noise": "philox"})
    pc.setup()
    ctx.set_seed(0xCAFE)
    b, y = np.ones(A.n), np.zeros(A.n)
    pc.apply_richardson(b, y, its=5)
    assert ctx.draw_counter == 10
    ref = orc.gibbs_richardson(A, b, np.zeros(A.n), 5, orc.Noise.philox(0xCAFE), col, 1.0, orc.SOR_SYMMETRIC)
    assert relerr(y, ref) < RTOL
    # checkpoint / resume: (y, seed, draw counter) is the whole chain state
    y2 = np.zeros(A.n)
    ctx.set_seed(0xCAFE)
    pc.apply_richardson(b, y2, its=2)
    ctx.draw_counter = 4
    pc.apply_richardson(b, y2, its=3)
    assert np.array_equal(y, y2)


# ---- a16: Cholesky sampler ---------------------------------------------------------------------------
@pytest.mark.parametrize("solve", ["trsv", "gemv"])
@pytest.mark.parametrize("dims", [(5, 5), (9, 9), (17, 17), (33, 20)])
def test_cholsampler_bitexact(pmg, ctx, orc, dims, solve):
    """trsv: the sequential substitution in dtrsv's order, bit-exact; gemv (default): explicit L^-1, two triangular
    matrix-vector products, equal to rounding."""
    rng = np.random.default_rng(SEED)
    A = orc.laplace(2, dims[0], dims[1], kappa=1.0)
    n = A.n
    pc = pmg.PC(ctx, "cholsampler")
    pc.set_operator(make_mat(pmg, ctx, A))
    pc.set_option("-pc_cholsampler_b200_solve", solve)
    pc.setup()

    def same(a, b):
        return np.array_equal(a, b) if solve == "trsv" else relerr(a, b) < RTOL
    assert f"size {n}" in pc.view()
    Lf = orc.potrf_lower(A.to_scipy().toarray())
    z, b = rng.standard_normal(4 * n), rng.standard_normal(n)
    # its == 1 (PCApply path) and its > 1 (cached forward solve, src/pc_chols.c:306-336)
    pc.set_noise_tape(z)
    y = np.zeros(n)
    pc.apply_richardson(b, y, its=1)
    assert same(y, orc.chol_sample(Lf, n, orc.Noise.tape(z[:n]), b))
    y3 = np.zeros(n)
    its_seen = []
    pc.set_sample_callback(lambda it, yy: its_seen.append(it))
    pc.apply_richardson(b, y3, its=3)
    assert same(y3, orc.chol_sample(Lf, n, orc.Noise.tape(z[3 * n:]), b))
    assert len(its_seen) == 3
    # exactness: mean of many samples -> A^-1 b is covered by the statistical test


# ---- a14/a15: MGMC V-cycle ------------------------------------------------------------------------------
def oracle_mg(orc, dim, dims, kappa, levels, smoother="sorgibbs", its=1, coarse="chol", coarse_its=1, omega=1.0, sweep=1):
    mg = orc.MG.geometric(dim, dims[0], dims[1], dims[2] if dim == 3 else 1, kappa, levels)
    kind = orc.KIND_SORGIBBS if smoother == "sorgibbs" else orc.KIND_MCGIBBS
    for l in range(levels):
        d = mg.level_dims(l)
        dd = d[:dim]
        star = l == levels - 1  # the fine operator is a star stencil (red-black); Galerkin levels are box stencils (2^d colours)
        col = orc.Coloring.parity(dd, 2 if star else 2 ** dim)
        if l == 0:
            if coarse == "chol":
                mg.set_smoother(0, orc.KIND_CHOL, 1.0, 1, coarse_its, None)
            else:
                mg.set_smoother(0, kind, omega, sweep, coarse_its, col)
        else:
            mg.set_smoother(l, kind, omega, sweep, its, col)
    mg.setup()
    return mg


@pytest.mark.parametrize("cycle", ["direct", "literal"])
@pytest.mark.parametrize("dims,levels", [((33, 33), 3), ((65, 33), 4), ((17, 17), 1), ((24, 16), 3), ((257, 131), 4), ((300, 202), 5)])
def test_gamgmc_default_cycle_matches_oracle(pmg, ctx, orc, dims, levels, cycle):
    """Defaults of src/pc_gamgmc.c:305-349: sorgibbs on the levels, cholsampler on the coarsest, V(1,1), Galerkin."""
    rng = np.random.default_rng(SEED)
    lap = pmg.Mat.laplace(ctx, 2, dims[0], dims[1], kappa=1.0)
    pc = pmg.PC(ctx, "gamgmc")
    pc.set_operator(lap)
    pc.set_options({"-gamgmc_pc_mg_levels": levels, "-pc_b200_coloring": "parity", "-pc_b200_cycle": cycle})
    pc.setup()
    assert pc.gamgmc_get_levels() == levels
    omg = oracle_mg(orc, 2, dims + (1,), 1.0, levels)
    # hierarchy: same sizes, Galerkin operators within rounding
    for l in range(levels):
        n, nnz, _ = pc.gamgmc_level_info(l)
        ref = omg.level_csr(l)
        assert n == ref.n
        if nnz >= 0:
            rp, col, val = pc.gamgmc_level_csr(l)
            assert np.array_equal(rp, ref.rowptr) and np.array_equal(col, ref.col)
            np.testing.assert_allclose(val, ref.val, rtol=1e-14, atol=1e-18)
    its = 3
    per = pc.noise_per_sample()
    z = rng.standard_normal(its * per)
    n = dims[0] * dims[1]
    b = rng.standard_normal(n)
    for guesszero in (False, True):
        pc.set_noise_tape(z)
        y = np.full(n, 0.5)
        pc.apply_richardson(b, y, its=its, guesszero=guesszero)
        tape = orc.Noise.tape(z)
        ref = omg.richardson(tape, b, np.full(n, 0.5), its, guesszero=guesszero)
        assert tape.tape_pos == its * per  # same noise-tape contract
        assert relerr(y, ref) < RTOL, relerr(y, ref)


def test_gamgmc_ex1_options_mcgibbs_everywhere(pmg, ctx, orc):
    """examples/ex1.c:41: geometric, 3 levels on 9x9, mcgibbs on levels and coarse, 2 its each."""
    rng = np.random.default_rng(SEED)
    lap = pmg.Mat.laplace(ctx, 2, 9, 9, kappa=10.0)
    pc = pmg.PC(ctx, "gamgmc")
    pc.set_operator(lap)
    pc.set_options({"-pc_gamgmc_mg_type": "mg", "-gamgmc_pc_mg_levels": 3, "-gamgmc_mg_levels_ksp_type": "richardson", "-gamgmc_mg_levels_pc_type": "mcgibbs",
                    "-gamgmc_mg_coarse_ksp_type": "richardson", "-gamgmc_mg_coarse_pc_type": "mcgibbs", "-gamgmc_mg_coarse_ksp_max_it": 2,
                    "-gamgmc_mg_levels_ksp_max_it": 2, "-pc_b200_coloring": "parity"})
    pc.setup()
    omg = oracle_mg(orc, 2, (9, 9, 1), 10.0, 3, smoother="mcgibbs", its=2, coarse="mcgibbs", coarse_its=2)
    per = pc.noise_per_sample()
    assert per == 4 * (81 + 25) + 2 * 9
    z = rng.standard_normal(2 * per)
    pc.set_noise_tape(z)
    b, y = np.ones(81), np.zeros(81)
    pc.apply_richardson(b, y, its=2)
    ref = omg.richardson(orc.Noise.tape(z), b, np.zeros(81), 2)
    assert relerr(y, ref) < RTOL
    assert "levels=3" in pc.view()


def test_gamgmc_user_interpolation_on_plain_csr(pmg, ctx, orc):
    """PCMGSetInterpolation path: a CSR operator with user-supplied P (here Q1 built by the oracle)."""
    rng = np.random.default_rng(SEED)
    A = orc.laplace(2, 17, 17, kappa=1.0)
    omg = oracle_mg(orc, 2, (17, 17, 1), 1.0, 2)
    mat = make_mat(pmg, ctx, A, orc.Coloring.parity((17, 17)))
    pc = pmg.PC(ctx, "gamgmc")
    pc.set_operator(mat)
    pc.gamgmc_set_levels(2)
    nf, nc = np.array([17, 17, 1], np.int64), np.array([9, 9, 1], np.int64)
    nnz = orc.lib().orc_q1_nnz(2, nf, nc)
    rp, col, val = np.empty(290, np.int64), np.empty(nnz, np.int32), np.empty(nnz, np.float64)
    orc.lib().orc_q1_interp(2, nf, nc, rp, col, val)
    pc.gamgmc_set_interpolation(1, 289, 81, rp, col, val)
    pc.setup()
    z = rng.standard_normal(pc.noise_per_sample())
    pc.set_noise_tape(z)
    b, y = rng.standard_normal(289), np.zeros(289)
    pc.apply_richardson(b, y, its=1)
    ref = omg.richardson(orc.Noise.tape(z), b, np.zeros(289), 1)
    assert relerr(y, ref) < RTOL
    # algebraic coarsening is PETSc-internal: refused, not faked
    pc2 = pmg.PC(ctx, "gamgmc")
    pc2.set_operator(mat)
    pc2.set_options({"-pc_gamgmc_mg_type": "gamg", "-gamgmc_pc_mg_levels": 2})
    with pytest.raises(pmg.PMGError) as e:
        pc2.setup()
    assert e.value.code == 2


def test_gamgmc_3d(pmg, ctx, orc):
    rng = np.random.default_rng(SEED)
    lap = pmg.Mat.laplace(ctx, 3, 9, 9, 9, kappa=1.0)
    pc = pmg.PC(ctx, "gamgmc")
    pc.set_operator(lap)
    pc.set_options({"-gamgmc_pc_mg_levels": 2, "-pc_b200_coloring": "parity"})
    pc.setup()
    omg = oracle_mg(orc, 3, (9, 9, 9), 1.0, 2)
    z = rng.standard_normal(2 * pc.noise_per_sample())
    pc.set_noise_tape(z)
    b, y = rng.standard_normal(729), np.zeros(729)
    pc.apply_richardson(b, y, its=2)
    ref = omg.richardson(orc.Noise.tape(z), b, np.zeros(729), 2)
    assert relerr(y, ref) < RTOL


@pytest.mark.parametrize("dims,levels,its", [((17, 17, 17), 3, 1), ((33, 17, 9), 2, 2), ((21, 13, 17), 2, 1)])
def test_gamgmc_3d_fused_top_level_matches_oracle(pmg, ctx, orc, dims, levels, its):
    """3D V-cycle whose finest level runs the TMA-fed fused sweep on pitched vectors (odd nx: pitch != nx), with the
    prolongation / residual / restriction kernels indexing the pitched layout: against the oracle's cycle (tape order F7)."""
    rng = np.random.default_rng(SEED)
    n = dims[0] * dims[1] * dims[2]
    lap = pmg.Mat.laplace(ctx, 3, *dims, kappa=1.0)
    pc = pmg.PC(ctx, "gamgmc")
    pc.set_operator(lap)
    pc.set_options({"-gamgmc_pc_mg_levels": levels, "-pc_b200_coloring": "parity", "-gamgmc_mg_levels_ksp_max_it": its})
    pc.setup()
    omg = oracle_mg(orc, 3, dims, 1.0, levels, its=its)
    z = rng.standard_normal(2 * pc.noise_per_sample())
    pc.set_noise_tape(z)
    b, y = rng.standard_normal(n), np.zeros(n)
    pc.apply_richardson(b, y, its=2)
    ref = omg.richardson(orc.Noise.tape(z), b, np.zeros(n), 2)
    assert relerr(y, ref) < RTOL


@pytest.mark.parametrize("dims,levels", [((65, 65, 65), 4), ((129, 33, 65), 4)])
def test_gamgmc_3d_fused_top_level_equals_unfused(pmg, ctx, dims, levels, monkeypatch):
    rng = np.random.default_rng(SEED)
    n = dims[0] * dims[1] * dims[2]
    b, y0 = rng.standard_normal(n), rng.standard_normal(n)
    out = []
    for fused in (True, False):
        if fused:
            monkeypatch.delenv("PMG_NO_FUSED_MG3", raising=False)
        else:
            monkeypatch.setenv("PMG_NO_FUSED_MG3", "1")
        lap = pmg.Mat.laplace(ctx, 3, *dims, kappa=0.5)
        pc = pmg.PC(ctx, "gamgmc")
        pc.set_operator(lap)
        pc.set_options({"-gamgmc_pc_mg_levels": levels, "-pc_b200_noise": "philox"})
        pc.setup()
        ctx.set_seed(21)
        y = y0.copy()
        pc.apply_richardson(b, y, its=3)
        out.append((y, pc.last_stats()["launches"]))
    assert np.array_equal(out[0][0], out[1][0])
    assert out[0][1] != out[1][1]


# ---- statistics with the device RNG (examples/ex1.c) -------------------------------------------------------
@pytest.mark.parametrize("pctype,opts,nsamp", [
    ("mcgibbs", {}, 400000),
    ("mcgibbs", {"-pc_mcgibbs_symmetric": ""}, 300000),
    ("sorgibbs", {}, 400000),
    ("gamgmc", {"-gamgmc_pc_mg_levels": 3}, 600000),
    ("cholsampler", {}, 600000),
])
def test_ex1_mean_convergence_device_rng(pmg, ctx, orc, pctype, opts, nsamp):
    """examples/ex1.c:83-135 acceptance check (rel. error of the sample mean <= 0.02) with the Philox
    generator.  The running mean is formed on the device side of the callback-free path: the chain is
    advanced in blocks and the block end states are averaged, which keeps the test fast; the estimator
    is the same sample mean."""
    A = orc.laplace(2, 9, 9, kappa=10.0)
    b = np.ones(81)
    ex_mean = np.linalg.solve(A.to_scipy().toarray(), b)
    lap = pmg.Mat.laplace(ctx, 2, 9, 9, kappa=10.0)
    pc = pmg.PC(ctx, pctype)
    pc.set_operator(lap)
    pc.set_options(opts)
    pc.set_option("-pc_b200_noise", "philox")
    pc.setup()
    ctx.set_seed(0xCAFE)
    y = np.zeros(81)
    pc.apply_richardson(b, y, its=1000)  # burn-in
    acc = {"m": np.zeros(81), "k": 0}

    def cb(it, yy):
        acc["k"] += 1
        acc["m"] += (yy - acc["m"]) / acc["k"]

    pc.set_sample_callback(cb)
    pc.apply_richardson(b, y, its=nsamp)
    rel = np.linalg.norm(acc["m"] - ex_mean) / np.linalg.norm(ex_mean)
    assert acc["k"] == nsamp
    assert rel <= 0.02, rel


# ---- the matrix-free structured path (K1/K3/K4/K5 of SURVEY 8(d)) ------------------------------------------
@pytest.mark.parametrize("dim,dims", [(2, (129, 129, 1)), (2, (64, 37, 1)), (2, (301, 197, 1)), (2, (9, 5, 1)), (3, (17, 12, 9)), (3, (16, 16, 16)), (3, (131, 21, 70)), (3, (250, 37, 9))])
@pytest.mark.parametrize("omega,sweep", [(1.0, 1), (1.3, 3), (0.8, 2)])
def test_matrix_free_laplace_gibbs_bitexact(pmg, ctx, orc, dim, dims, omega, sweep):
    """The matrix-free operator must reproduce the assembled one (src/problems.c:14-75) bit for bit."""
    rng = np.random.default_rng(SEED)
    A = orc.laplace(dim, *dims, kappa=1.5)
    col = orc.Coloring.parity(dims[:dim])
    lap = pmg.Mat.laplace(ctx, dim, *dims, kappa=1.5)
    k, color = lap.get_coloring()
    assert k == 2 and np.array_equal(color, col.color)
    pc = pmg.PC(ctx, "mcgibbs")
    pc.set_operator(lap)
    pc.mcgibbs_set_omega(omega)
    pc.mcgibbs_set_sweep_type(sweep)
    pc.setup()
    its = 3
    z = rng.standard_normal(its * pc.noise_per_sample())
    pc.set_noise_tape(z)
    b, y = rng.standard_normal(A.n), rng.standard_normal(A.n)
    y0 = y.copy()
    pc.apply_richardson(b, y, its=its)
    ref = orc.gibbs_richardson(A, b, y0.copy(), its, orc.Noise.tape(z), col, omega, sweep)
    assert np.array_equal(y, ref), relerr(y, ref)
    x = rng.standard_normal(A.n)
    assert np.array_equal(lap.mult(x), A.to_scipy().tocsr() @ x) or relerr(lap.mult(x), A.to_scipy() @ x) < 1e-14
    # philox noise: the grid operator keys the generator on the padded index (philox.cuh); the oracle definition agrees
    pc.set_noise_mode(pmg.NOISE_PHILOX)
    ctx.set_seed(7)
    y1 = y0.copy()
    pc.apply_richardson(b, y1, its=2)
    ref = orc.gibbs_richardson(A, b, y0.copy(), 2, orc.Noise.philox(7, grid=dims), col, omega, sweep)
    assert relerr(y1, ref) < RTOL


@pytest.mark.parametrize("dim,dims,levels", [(2, (65, 65, 1), 4), (2, (40, 28, 1), 3), (3, (17, 17, 17), 3), (3, (12, 10, 8), 2)])
def test_matrix_free_hierarchy_equals_assembled_hierarchy(pmg, ctx, orc, dim, dims, levels):
    """Device Galerkin product + stencil-array levels + matrix-free transfers vs the assembled CSR hierarchy:
    the same operators (bitwise) and the same V-cycle output."""
    rng = np.random.default_rng(SEED)
    A = orc.laplace(dim, *dims, kappa=1.0)
    n = A.n
    opts = {"-gamgmc_pc_mg_levels": levels, "-pc_b200_coloring": "parity", "-gamgmc_mg_levels_ksp_max_it": 2}
    lap = pmg.Mat.laplace(ctx, dim, *dims, kappa=1.0)
    pc = pmg.PC(ctx, "gamgmc"); pc.set_operator(lap); pc.set_options(opts); pc.setup()
    omg = orc.MG.geometric(dim, dims[0], dims[1], dims[2], 1.0, levels)
    for l in range(levels):
        rp, col, val = pc.gamgmc_level_csr(l)
        ref = omg.level_csr(l)
        assert np.array_equal(rp, ref.rowptr) and np.array_equal(col, ref.col)
        assert np.array_equal(val, ref.val), np.abs(val - ref.val).max()
    for l in range(levels):
        d = omg.level_dims(l)
        if l == 0:
            omg.set_smoother(0, orc.KIND_CHOL, 1.0, 1, 1, None)
        else:
            omg.set_smoother(l, orc.KIND_SORGIBBS, 1.0, 1, 2, orc.Coloring.parity(d[:dim], 2 if l == levels - 1 else 2 ** dim))
    omg.setup()
    z = rng.standard_normal(2 * pc.noise_per_sample())
    pc.set_noise_tape(z)
    b, y = rng.standard_normal(n), np.zeros(n)
    pc.apply_richardson(b, y, its=2)
    ref = omg.richardson(orc.Noise.tape(z), b, np.zeros(n), 2)
    assert relerr(y, ref) < RTOL, relerr(y, ref)
    assert "shared" in pc.view() or min(dims[:dim]) < 9


# ---- the coarse tail of the V-cycle in one cluster launch (stencil_op.cu grid_tail_kernel) ---------------------------
@pytest.mark.parametrize("dim,dims,levels,extra", [
    (2, (129, 129, 1), 5, {}),
    (2, (257, 129, 1), 6, {"-gamgmc_mg_levels_ksp_max_it": 2}),
    (2, (161, 97, 1), 4, {"-gamgmc_mg_levels_pc_type": "mcgibbs", "-gamgmc_mg_levels_pc_mcgibbs_symmetric": "", "-gamgmc_mg_levels_pc_mcgibbs_omega": 1.3}),
    (3, (33, 33, 33), 3, {}),
    (3, (41, 25, 17), 3, {"-gamgmc_mg_levels_ksp_max_it": 2}),
])
@pytest.mark.parametrize("noise", ["philox", "tape"])
@pytest.mark.parametrize("smem", [True, False])
def test_coarse_tail_launch_is_bit_identical(pmg, ctx, dim, dims, levels, extra, noise, smem, monkeypatch):
    """One launch for levels 0..lt -- one CTA with the level vectors in shared memory (tail2d.cuh; 2D) or one cluster through
    global memory (grid_tail_kernel) -- must give exactly the launch-per-colour result, with the same noise blocks."""
    if smem and dim == 3:
        pytest.skip("the shared-memory tail is 2D")
    if not smem:
        monkeypatch.setenv("PMG_NO_TAIL_SMEM", "1")
    rng = np.random.default_rng(SEED)
    n = dims[0] * dims[1] * dims[2]
    b, y0 = rng.standard_normal(n), rng.standard_normal(n)
    out = []
    for tail_max in (300000, 0):
        lap = pmg.Mat.laplace(ctx, dim, *dims, kappa=1.0)
        pc = pmg.PC(ctx, "gamgmc")
        pc.set_operator(lap)
        pc.set_options(dict(extra, **{"-gamgmc_pc_mg_levels": levels, "-pc_b200_tail_max_n": tail_max}))
        pc.setup()
        if noise == "tape":
            pc.set_noise_tape(np.random.default_rng(5).standard_normal(3 * pc.noise_per_sample()))
        else:
            pc.set_noise_mode(pmg.NOISE_PHILOX)
            ctx.set_seed(99)
        y = y0.copy()
        d0 = ctx.draw_counter
        pc.apply_richardson(b, y, its=3)
        out.append((y, pc.last_stats()["launches"], ctx.draw_counter - d0))
    assert np.array_equal(out[0][0], out[1][0]), relerr(out[0][0], out[1][0])
    assert out[0][1] < out[1][1] and out[0][2] == out[1][2]


# ---- MATLRC operators A + B diag(S) B^T (SURVEY a1 / a9 / a10-LRC, config 5's low-rank observation update) -------------
def _obs_matrix(rng, n, k):
    """k sparse 'ball observation' columns like MakeObservationMats (src/obs.c:135-180): a few positive weights each."""
    B = np.zeros((n, k))
    for j in range(k):
        idx = rng.choice(n, size=min(n, 12), replace=False)
        B[idx, j] = rng.uniform(0.2, 1.0, idx.size)
    return B


@pytest.mark.parametrize("kind", ["csr", "grid"])
@pytest.mark.parametrize("pctype,omega,sweep", [("mcgibbs", 1.0, 1), ("mcgibbs", 1.0, 3), ("mcgibbs", 1.0, 2), ("sorgibbs", 1.0, 1)])
def test_lrc_gibbs_matches_oracle(pmg, ctx, orc, kind, pctype, omega, sweep):
    rng = np.random.default_rng(SEED)
    dims = (33, 21)
    A = orc.laplace(2, *dims, kappa=1.5)
    n, k = A.n, 5
    B, S = _obs_matrix(rng, n, k), rng.uniform(50.0, 200.0, k)  # S = 1/sigma^2 of the observations
    col = orc.Coloring.parity(dims)
    base = make_mat(pmg, ctx, A, col) if kind == "csr" else pmg.Mat.laplace(ctx, 2, *dims, kappa=1.5)
    mat = pmg.Mat.lrc(base, B, S)
    x = rng.standard_normal(n)
    assert relerr(mat.mult(x), A.to_scipy() @ x + B @ (S * (B.T @ x))) < 1e-13
    pc = pmg.PC(ctx, pctype)
    pc.set_operator(mat)
    if pctype == "mcgibbs":
        pc.mcgibbs_set_omega(omega)
        pc.mcgibbs_set_sweep_type(sweep)
    pc.setup()
    its = 3
    per = pc.noise_per_sample()
    assert per == (n + k) * (2 if sweep == 3 else 1)
    z = rng.standard_normal(its * per)
    pc.set_noise_tape(z)
    b, y = rng.standard_normal(n), rng.standard_normal(n)
    ref = orc.lrc_gibbs_richardson(A, B, S, b, y.copy(), its, orc.Noise.tape(z), col, omega, sweep)
    pc.apply_richardson(b, y, its=its)
    assert relerr(y, ref) < RTOL
    # MCSORApply on the same operator: deterministic sweep + post-correction (src/mc_sor.c:216-239)
    mc = pmg.MCSOR(mat)
    mc.set_sweep_type(sweep)
    y2 = rng.standard_normal(n)
    Bb = {d: orc.lrc_build_correction(A, B, S, col, 1.0, d) for d in (orc.SOR_FORWARD, orc.SOR_BACKWARD)}
    ref2 = orc.lrc_mcsor_apply(A, B, Bb, b, y2.copy(), col, 1.0, sweep)
    mc.apply(b, y2)
    assert relerr(y2, ref2) < RTOL


def test_lrc_posterior_mean_device_rng(pmg, ctx, orc):
    """Statistical acceptance in the style of examples/ex4.c: the sample mean approaches (A + B S B^T)^-1 b."""
    rng = np.random.default_rng(SEED)
    dims = (17, 17)
    A = orc.laplace(2, *dims, kappa=3.0)
    n, k = A.n, 6
    B, S = _obs_matrix(rng, n, k), np.full(k, 100.0)
    b = rng.standard_normal(n)
    mat = pmg.Mat.lrc(pmg.Mat.laplace(ctx, 2, *dims, kappa=3.0), B, S)
    pc = pmg.PC(ctx, "mcgibbs")
    pc.set_operator(mat)
    pc.set_options({"-pc_mcgibbs_symmetric": "", "-pc_b200_noise": "philox"})
    pc.setup()
    ctx.set_seed(0xCAFE)
    y, acc, nsamp = np.zeros(n), np.zeros(n), 6000
    pc.apply_richardson(b, y, its=100)  # burn-in

    def cb(it, ys):
        acc[:] += ys

    pc.set_sample_callback(cb)
    pc.apply_richardson(b, y, its=nsamp)
    P = A.to_scipy().toarray() + B @ np.diag(S) @ B.T
    mean = np.linalg.solve(P, b)
    # Monte Carlo error of the mean ~ sqrt(trace(P^-1) / N) = 0.04 |mean| here; the reference's own acceptance is 5-10 % (ex4.c)
    assert np.linalg.norm(acc / nsamp - mean) < 0.1 * np.linalg.norm(mean)


def _q1_1d(nf):
    nc = (nf + 1) // 2
    P = np.zeros((nf, nc))
    for i in range(nf):
        if i % 2 == 0:
            P[i, i // 2] = 1.0
        else:
            P[i, i // 2] = 0.5
            if i // 2 + 1 < nc:
                P[i, i // 2 + 1] = 0.5
    return P


@pytest.mark.parametrize("its_lv", [1, 2])
def test_gamgmc_on_lrc_operator_matches_numpy_restatement(pmg, ctx, orc, its_lv):
    """PCGAMGMC_SetUpHierarchy's MATLRC branch (src/pc_gamgmc.c:157-196): B_c = P^T B_f, every level samples from and
    takes residuals with A_l + B_l S B_l^T, the coarsest level factors the assembled sum (src/pc_chols.c:119-157).
    Checked against a numpy restatement of the V-cycle built from the oracle's LRC sampler."""
    rng = np.random.default_rng(SEED)
    dims = [(17, 17), (9, 9), (5, 5)]  # fine -> coarse
    L = len(dims)
    A = [orc.laplace(2, *dims[0], kappa=1.0)]
    Ps = []
    for l in range(1, L):
        P = np.kron(_q1_1d(dims[l - 1][1]), _q1_1d(dims[l - 1][0]))  # natural order i + nx j: y index is the slow one
        Ps.append(P)
        Ad = P.T @ A[-1].to_scipy().toarray() @ P
        import scipy.sparse as sp
        m = sp.csr_matrix(Ad)
        m.sort_indices()
        A.append(orc.CSR(m.shape[0], m.indptr.astype(np.int64), m.indices.astype(np.int32), m.data.astype(np.float64)))
    n, k = A[0].n, 4
    B = [_obs_matrix(rng, n, k)]
    S = rng.uniform(20.0, 80.0, k)
    for l in range(1, L):
        B.append(Ps[l - 1].T @ B[-1])
    cols = [orc.Coloring.parity(dims[0])] + [orc.Coloring.parity(d, 2) for d in dims[1:]]
    # the stencil-array levels sweep in 4 colours (i mod 2, j mod 2); the fine level is red-black
    cols = [orc.Coloring.parity(dims[0])] + [orc.Coloring(np.array([(i % 2) + 2 * (j % 2) for j in range(d[1]) for i in range(d[0])], np.int32), 4) for d in dims[1:]]

    lap = pmg.Mat.laplace(ctx, 2, *dims[0], kappa=1.0)
    mat = pmg.Mat.lrc(lap, B[0], S)
    pc = pmg.PC(ctx, "gamgmc")
    pc.set_operator(mat)
    pc.set_options({"-gamgmc_pc_mg_levels": L, "-gamgmc_mg_levels_ksp_max_it": its_lv})
    pc.setup()
    nsamp = 2
    per = pc.noise_per_sample()
    z = rng.standard_normal(nsamp * per)
    pc.set_noise_tape(z)
    b, y = rng.standard_normal(n), rng.standard_normal(n)

    noise = orc.Noise.tape(z)
    Pd = [A[l].to_scipy().toarray() + B[l] @ np.diag(S) @ B[l].T for l in range(L)]

    def smooth(l, rhs, x):
        return orc.lrc_gibbs_richardson(A[l], B[l], S, rhs, x, its_lv, noise, cols[l], 1.0, orc.SOR_FORWARD)

    def cycle(l, rhs, x):
        if l == L - 1:  # coarsest: y = L^-T (L^-1 b + z)
            Lc = np.linalg.cholesky(Pd[l])
            v = np.linalg.solve(Lc, rhs) + orc.noise_fill(noise, rhs.size)
            return np.linalg.solve(Lc.T, v)
        x = smooth(l, rhs, x)
        r = rhs - Pd[l] @ x
        xc = cycle(l + 1, Ps[l].T @ r, np.zeros(A[l + 1].n))
        x = x + Ps[l] @ xc
        return smooth(l, rhs, x)

    ref = y.copy()
    for _ in range(nsamp):  # y += MG(b - A y)  (src/pc_gamgmc.c:253-256)
        ref = ref + cycle(0, b - Pd[0] @ ref, np.zeros(n))
    pc.apply_richardson(b, y, its=nsamp)
    assert relerr(y, ref) < 1e-10


# ---- the fused four-colour sweep of the stencil-array levels (box_stream.cuh) ------------------------------------------
@pytest.mark.parametrize("dims,levels,extra", [
    ((257, 257, 1), 4, {}),
    ((301, 173, 1), 4, {"-gamgmc_mg_levels_ksp_max_it": 2}),
    ((129, 513, 1), 4, {"-gamgmc_mg_levels_pc_type": "mcgibbs", "-gamgmc_mg_levels_pc_mcgibbs_symmetric": "", "-gamgmc_mg_levels_pc_mcgibbs_omega": 1.4}),
])
@pytest.mark.parametrize("noise", ["philox", "tape"])
def test_box_stream_sweep_is_bit_identical(pmg, ctx, dims, levels, extra, noise, monkeypatch):
    """One pass per directional sweep on the 9-point levels must reproduce the launch-per-colour result exactly."""
    rng = np.random.default_rng(SEED)
    n = dims[0] * dims[1]
    b, y0 = rng.standard_normal(n), rng.standard_normal(n)
    out = []
    for stream in (True, False):
        monkeypatch.setenv("PMG_NO_BOX2", "1")  # the natural-layout one-pass sweep, not its pitched TMA successor (box2d.cuh)
        if stream:
            monkeypatch.setenv("PMG_BOX_STREAM_MIN", "0")
            monkeypatch.delenv("PMG_NO_BOX_STREAM", raising=False)
        else:
            monkeypatch.setenv("PMG_NO_BOX_STREAM", "1")
        lap = pmg.Mat.laplace(ctx, 2, *dims, kappa=1.0)
        pc = pmg.PC(ctx, "gamgmc")
        pc.set_operator(lap)
        pc.set_options(dict(extra, **{"-gamgmc_pc_mg_levels": levels, "-pc_b200_tail_max_n": 0}))
        pc.setup()
        if noise == "tape":
            pc.set_noise_tape(np.random.default_rng(5).standard_normal(2 * pc.noise_per_sample()))
        else:
            pc.set_noise_mode(pmg.NOISE_PHILOX)
            ctx.set_seed(99)
        y = y0.copy()
        pc.apply_richardson(b, y, its=2)
        out.append((y, pc.last_stats()["launches"]))
    assert np.array_equal(out[0][0], out[1][0]), relerr(out[0][0], out[1][0])
    assert out[0][1] < out[1][1]


# ---- the TMA-fed one-pass kernels of the 9-point levels (box2d.cuh): sweep + residual + restriction / prolongation + sweep ----
@pytest.mark.parametrize("dims,levels,extra", [
    ((257, 257, 1), 4, {}),
    ((301, 173, 1), 4, {"-gamgmc_mg_levels_ksp_max_it": 2}),
    ((129, 513, 1), 4, {"-gamgmc_mg_levels_pc_type": "mcgibbs", "-gamgmc_mg_levels_pc_mcgibbs_symmetric": "", "-gamgmc_mg_levels_pc_mcgibbs_omega": 1.4}),
    ((260, 140, 1), 3, {"-gamgmc_mg_levels_pc_type": "mcgibbs", "-gamgmc_mg_levels_pc_mcgibbs_backward": "", "-gamgmc_mg_levels_pc_mcgibbs_omega": 0.8}),
    ((1025, 769, 1), 6, {}),
])
@pytest.mark.parametrize("noise", ["philox", "tape"])
def test_box2d_one_pass_levels_are_bit_identical(pmg, ctx, dims, levels, extra, noise, monkeypatch):
    """Pre-sample + residual + restriction in one pass and prolongation + post-sample in another, on pitched level vectors,
    must reproduce the launch-per-colour V-cycle (separate residual / restriction / prolongation kernels) exactly."""
    rng = np.random.default_rng(SEED)
    n = dims[0] * dims[1]
    b, y0 = rng.standard_normal(n), rng.standard_normal(n)
    out = []
    for box2 in (True, False):
        monkeypatch.setenv("PMG_NO_BOX_STREAM", "1")
        if box2:
            monkeypatch.delenv("PMG_NO_BOX2", raising=False)
        else:
            monkeypatch.setenv("PMG_NO_BOX2", "1")
        lap = pmg.Mat.laplace(ctx, 2, *dims, kappa=1.0)
        pc = pmg.PC(ctx, "gamgmc")
        pc.set_operator(lap)
        pc.set_options(dict(extra, **{"-gamgmc_pc_mg_levels": levels, "-pc_b200_tail_max_n": 0}))
        pc.setup()
        if noise == "tape":
            pc.set_noise_tape(np.random.default_rng(5).standard_normal(2 * pc.noise_per_sample()))
        else:
            pc.set_noise_mode(pmg.NOISE_PHILOX)
            ctx.set_seed(99)
        y = y0.copy()
        pc.apply_richardson(b, y, its=2)
        out.append((y, pc.last_stats()["launches"]))
    assert np.array_equal(out[0][0], out[1][0]), relerr(out[0][0], out[1][0])
    assert out[0][1] < out[1][1]


@pytest.mark.parametrize("dims,levels", [((257, 131), 4), ((300, 202), 4), ((513, 129), 5)])
def test_box2d_one_pass_levels_match_oracle(pmg, ctx, orc, dims, levels):
    """The same path against the CPU oracle's V-cycle (injected tape, 1e-12)."""
    rng = np.random.default_rng(SEED)
    pc = pmg.PC(ctx, "gamgmc")
    pc.set_operator(pmg.Mat.laplace(ctx, 2, dims[0], dims[1], kappa=1.0))
    pc.set_options({"-gamgmc_pc_mg_levels": levels, "-pc_b200_tail_max_n": 0})
    pc.setup()
    assert "one-pass" in pc.view()
    omg = oracle_mg(orc, 2, dims + (1,), 1.0, levels)
    its, per, n = 2, pc.noise_per_sample(), dims[0] * dims[1]
    z, b = rng.standard_normal(its * per), rng.standard_normal(n)
    pc.set_noise_tape(z)
    y = np.full(n, 0.5)
    pc.apply_richardson(b, y, its=its)
    ref = omg.richardson(orc.Noise.tape(z), b, np.full(n, 0.5), its)
    assert relerr(y, ref) < RTOL, relerr(y, ref)


@pytest.mark.parametrize("dims,levels,extra", [
    ((257, 257, 1), 4, {}),
    ((300, 202, 1), 4, {"-gamgmc_mg_levels_pc_type": "mcgibbs", "-gamgmc_mg_levels_pc_mcgibbs_backward": "", "-gamgmc_mg_levels_pc_mcgibbs_omega": 1.3}),
    ((1030, 517, 1), 6, {"-gamgmc_mg_levels_ksp_max_it": 2}),
])
@pytest.mark.parametrize("noise", ["philox", "tape"])
@pytest.mark.parametrize("cfg", ["0", "1", "2"])
def test_fine_level_tma_residual_restriction_is_bit_identical(pmg, ctx, dims, levels, extra, noise, cfg, monkeypatch):
    """Fine-level pre-sample + residual + restriction on the TMA structure (sweep2d.cuh RESTRICT) vs the plain-load
    streaming kernel (stream2d.cuh): same arithmetic, bit for bit, for every register / occupancy configuration."""
    rng = np.random.default_rng(SEED)
    n = dims[0] * dims[1]
    b, y0 = rng.standard_normal(n), rng.standard_normal(n)
    out = []
    monkeypatch.setenv("PMG_SW2R_CFG", cfg)
    for tma in (True, False):
        if tma:
            monkeypatch.delenv("PMG_NO_TMA_RESTRICT", raising=False)
        else:
            monkeypatch.setenv("PMG_NO_TMA_RESTRICT", "1")
        pc = pmg.PC(ctx, "gamgmc")
        pc.set_operator(pmg.Mat.laplace(ctx, 2, *dims, kappa=1.0))
        pc.set_options(dict(extra, **{"-gamgmc_pc_mg_levels": levels}))
        pc.setup()
        if noise == "tape":
            pc.set_noise_tape(np.random.default_rng(5).standard_normal(2 * pc.noise_per_sample()))
        else:
            pc.set_noise_mode(pmg.NOISE_PHILOX)
            ctx.set_seed(99)
        y = y0.copy()
        pc.apply_richardson(b, y, its=2)
        out.append(y)
    assert np.array_equal(out[0], out[1]), relerr(out[0], out[1])


@pytest.mark.parametrize("dim,dims,levels", [(2, (129, 97, 1), 4), (3, (33, 25, 17), 3)])
def test_fused_residual_restriction_is_bit_identical(pmg, ctx, dim, dims, levels, monkeypatch):
    """b_c = P^T (b - A x) in one kernel (box_restrict_residual_kernel) vs residual + restriction."""
    rng = np.random.default_rng(SEED)
    n = dims[0] * dims[1] * dims[2]
    b, y0 = rng.standard_normal(n), rng.standard_normal(n)
    out = []
    monkeypatch.setenv("PMG_NO_BOX2", "1")  # the stand-alone fused kernel, not the one-pass levels that superseded it (box2d.cuh)
    for fused in (True, False):
        if fused:
            monkeypatch.setenv("PMG_FUSED_RESIDUAL", "1")
        else:
            monkeypatch.delenv("PMG_FUSED_RESIDUAL", raising=False)
        pc = pmg.PC(ctx, "gamgmc")
        pc.set_operator(pmg.Mat.laplace(ctx, dim, *dims, kappa=1.0))
        pc.set_options({"-gamgmc_pc_mg_levels": levels, "-pc_b200_tail_max_n": 0})
        pc.setup()
        pc.set_noise_mode(pmg.NOISE_PHILOX)
        ctx.set_seed(17)
        y = y0.copy()
        pc.apply_richardson(b, y, its=2)
        out.append((y, pc.last_stats()["launches"]))
    assert np.array_equal(out[0][0], out[1][0]), relerr(out[0][0], out[1][0])
    assert out[0][1] < out[1][1]


# ---- estimators on device-RNG chains vs reference-RNG chains (BASELINE north star: mean, covariance and IACT within
#      Monte Carlo error; estimators of src/stats.c:94-117 and src/iact.c:17-92 as restated in oracle/stats.c) -------------
@pytest.mark.parametrize("pctype,kappa", [("mcgibbs", 0.05), ("gamgmc", 0.05)])
def test_device_rng_covariance_and_iact_match_reference_rng_chain(pmg, ctx, orc, pctype, kappa):
    """The same sampler, once on the device with Philox noise and once in the oracle with the reference's rander48
    Box-Muller stream: the covariance estimate converges to A^-1 at the same rate and the integrated autocorrelation time
    of a QOI agrees.  kappa = 0.05 makes A ill-conditioned (cond ~ 50), so plain Gibbs mixes slowly (IACT >> 1) and the
    V-cycle does not (IACT ~ 1): the comparison would notice a wrong noise scale or a correlated generator."""
    dims, N, burn = (9, 9), 40000, 2000
    A = orc.laplace(2, *dims, kappa=kappa)
    n = A.n
    Ad = A.to_scipy().toarray()
    q = np.zeros(n)
    q[n // 2] = 1.0  # QOI: the centre node
    # --- device chain ---
    lap = pmg.Mat.laplace(ctx, 2, *dims, kappa=kappa)
    pc = pmg.PC(ctx, pctype)
    pc.set_operator(lap)
    pc.set_options({"-pc_mcgibbs_symmetric": ""} if pctype == "mcgibbs" else {"-gamgmc_pc_mg_levels": 3})
    pc.set_option("-pc_b200_noise", "philox")
    pc.setup()
    ctx.set_seed(0xCAFE)
    y = np.zeros(n)
    pc.apply_richardson(np.zeros(n), y, its=burn)
    dev = np.empty((N, n))

    def cb(it, ys):
        dev[it] = ys

    pc.set_sample_callback(cb)
    pc.apply_richardson(np.zeros(n), y, its=N)
    # --- reference-RNG chain (oracle) ---
    ref = np.empty((N, n))

    def ocb(it, ys):
        if it >= burn:
            ref[it - burn] = ys

    if pctype == "mcgibbs":
        orc.gibbs_richardson(A, np.zeros(n), np.zeros(n), burn + N, orc.Noise.rander48(), orc.Coloring.parity(dims), 1.0, orc.SOR_SYMMETRIC, callback=ocb)
    else:
        omg = orc.MG.geometric(2, dims[0], dims[1], 1, kappa, 3)
        omg.set_smoother(0, orc.KIND_CHOL, 1.0, 1, 1, None)
        for l in (1, 2):
            d = omg.level_dims(l)
            omg.set_smoother(l, orc.KIND_SORGIBBS, 1.0, orc.SOR_FORWARD, 1, orc.Coloring.parity(d[:2], 2 if l == 2 else 4))
        omg.setup()
        omg.richardson(orc.Noise.rander48(), np.zeros(n), np.zeros(n), burn + N, callback=ocb)
    # covariance (EstimateCovarianceMatErrors, src/stats.c:94-117): relative Frobenius error against A^-1
    e_dev = orc.cov_errors(Ad, dev[None])[0]
    e_ref = orc.cov_errors(Ad, ref[None])[0]
    tau_dev, ok_dev = orc.iact(dev @ q)
    tau_ref, ok_ref = orc.iact(ref @ q)
    assert ok_dev and ok_ref
    assert abs(tau_dev - tau_ref) < 0.3 * tau_ref + 0.3, (tau_dev, tau_ref)
    if pctype == "gamgmc":
        assert tau_dev < 2.0  # the multigrid sampler decorrelates in about one cycle
    else:
        assert tau_dev > 3.0  # plain Gibbs does not, on this operator
    # Monte Carlo error of a covariance estimate from N / tau effective samples; both chains must sit at that level
    bound = 6.0 * np.sqrt(max(tau_ref, 1.0) / N) * np.sqrt(n) / 3.0
    assert e_dev < bound and e_ref < bound, (e_dev, e_ref, bound)
    assert e_dev < 3.0 * e_ref + 0.02 and e_ref < 3.0 * e_dev + 0.02, (e_dev, e_ref)


# ---- PCWOODBURY (src/woodbury.c): a sampler on A + the Woodbury correction of a MATLRC operator -----------------------------
@pytest.mark.parametrize("sampler,sopts", [("mcgibbs", {"-pc_woodbury_sampler_pc_mcgibbs_omega": 1.2, "-pc_woodbury_sampler_pc_mcgibbs_symmetric": ""}), ("sorgibbs", {})])
def test_woodbury_matches_numpy_restatement(pmg, ctx, orc, sampler, sopts):
    """PCApplyRichardson_Woodbury (src/woodbury.c:259-286) with G = C (S^-1 + B^T C)^-1, C = solver(B) (:21-86), exact solver."""
    rng = np.random.default_rng(SEED)
    dims = (21, 17)
    A = orc.laplace(2, *dims, kappa=1.5)
    n, k = A.n, 4
    B, S = _obs_matrix(rng, n, k), rng.uniform(50.0, 200.0, k)
    col = orc.Coloring.parity(dims)
    mat = pmg.Mat.lrc(make_mat(pmg, ctx, A, col), B, S)
    pc = pmg.PC(ctx, "woodbury")
    pc.set_operator(mat)
    pc.set_options(dict(sopts, **{"-pc_woodbury_sampler": sampler, "-pc_woodbury_solver": "cholesky"}))
    pc.setup()
    its = 3
    per = pc.noise_per_sample()
    sym = sampler == "mcgibbs"
    assert per == k + n * (2 if sym else 1)
    z = rng.standard_normal(its * per)
    pc.set_noise_tape(z)
    b, y = rng.standard_normal(n), rng.standard_normal(n)
    Ad = A.to_scipy().toarray()
    Cm = np.linalg.solve(Ad, B)
    G = Cm @ np.linalg.inv(np.diag(1.0 / S) + B.T @ Cm)
    noise = orc.Noise.tape(z)
    ref = y.copy()
    for _ in range(its):
        w = b + B @ (np.sqrt(np.abs(S)) * orc.noise_fill(noise, k))  # the k draws come first (src/woodbury.c:273)
        ref = orc.gibbs_richardson(A, w, ref, 1, noise, col, 1.2 if sym else 1.0, orc.SOR_SYMMETRIC if sym else orc.SOR_FORWARD)
        ref = ref - G @ (B.T @ ref)
    pc.apply_richardson(b, y, its=its)
    assert relerr(y, ref) < 1e-10


def test_woodbury_exact_sampler_gives_the_posterior(pmg, ctx, orc):
    """examples/ex13.py: -pc_woodbury_sampler cholsampler -pc_woodbury_solver cholesky draws EXACT, independent samples of
    N((A + B S B^T)^-1 b, (A + B S B^T)^-1): mean and covariance estimators at the Monte Carlo level, IACT = 1."""
    rng = np.random.default_rng(SEED)
    dims = (9, 9)
    A = orc.laplace(2, *dims, kappa=2.0)
    n, k, N = A.n, 5, 20000
    B, S = _obs_matrix(rng, n, k), np.full(k, 30.0)
    b = rng.standard_normal(n)
    mat = pmg.Mat.lrc(pmg.Mat.from_csr(ctx, A.rowptr, A.col, A.val), B, S)
    pc = pmg.PC(ctx, "woodbury")
    pc.set_operator(mat)
    pc.set_options({"-pc_woodbury_sampler": "cholsampler", "-pc_woodbury_solver": "cholesky", "-pc_b200_noise": "philox"})
    pc.setup()
    ctx.set_seed(0xCAFE)
    out = np.empty((N, n))

    def cb(it, ys):
        out[it] = ys

    pc.set_sample_callback(cb)
    pc.apply_richardson(b, np.zeros(n), its=N)
    P = A.to_scipy().toarray() + B @ np.diag(S) @ B.T
    mean = np.linalg.solve(P, b)
    assert np.linalg.norm(out.mean(0) - mean) < 4.0 * np.sqrt(np.trace(np.linalg.inv(P)) / N)
    assert orc.cov_errors(P, (out - mean)[None])[0] < 6.0 * np.sqrt(n / N) / 3.0
    tau, ok = orc.iact(out[:, n // 2])
    assert ok and abs(tau - 1.0) < 0.15


# ---- statistics on the device (examples/benchmark/main.cc:151-175, src/iact.c) ---------------------------------------------
def test_device_qoi_trace_welford_and_iact(pmg, ctx, orc):
    rng = np.random.default_rng(SEED)
    dims, N = (33, 21), 300
    lap = pmg.Mat.laplace(ctx, 2, *dims, kappa=0.5)
    n = lap.n
    pc = pmg.PC(ctx, "mcgibbs")
    pc.set_operator(lap)
    pc.set_options({"-pc_mcgibbs_symmetric": "", "-pc_b200_noise": "philox"})
    pc.setup()
    meas = rng.standard_normal(n)
    pc.set_qoi(meas, N, True)
    got = np.empty((N, n))

    def cb(it, ys):
        got[it] = ys

    pc.set_sample_callback(cb)
    ctx.set_seed(3)
    b, y = rng.standard_normal(n), np.zeros(n)
    pc.apply_richardson(b, y, its=N)
    q = pc.get_qoi()
    assert q.size == N and np.abs(q - got @ meas).max() < 1e-11 * np.abs(got @ meas).max()
    mean, var, seen = pc.get_mean_var()
    assert seen == N
    assert np.abs(mean - got.mean(0)).max() < 1e-12 and np.abs(var - got.var(0, ddof=1)).max() < 1e-11
    with pytest.raises(pmg.PMGError):
        pc.apply_richardson(b, y, its=1)  # the trace is full
    assert pc.get_qoi(reset=True).size == N and pc.get_qoi().size == 0
    # Autocorrelation / IACT of src/iact.c on cuFFT against the oracle's restatement
    rho, m = 0.9, 60000
    x = np.empty(m)
    x[0] = 0.0
    e = rng.standard_normal(m)
    for i in range(1, m):
        x[i] = rho * x[i - 1] + e[i]
    acf = pmg.autocorrelation(ctx, x[:5000])
    np.testing.assert_allclose(acf, orc.autocorrelation(x[:5000]), atol=1e-10)
    tau, valid = pmg.iact(ctx, x)
    tau_o, valid_o = orc.iact(x)
    assert valid == valid_o and abs(tau - tau_o) < 1e-8 and abs(tau - (1 + rho) / (1 - rho)) < 2.5
    with pytest.raises(pmg.PMGError):
        pmg.iact(ctx, np.zeros(1))


# ---- the fused 3D sweep on shapes that exercise narrow last strips (16 lanes per grid row), thin z-edge bands, ------------
# ---- z-constant edge tiles and ragged tile edges: bitwise equal to one launch per colour (itself pinned to the oracle) ------
@pytest.mark.parametrize("noise", ["philox", "injected", "none"])
@pytest.mark.parametrize("dims", [(150, 40, 20), (130, 35, 9), (50, 64, 16), (512, 33, 12), (176, 61, 70), (57, 31, 12), (36, 30, 5), (300, 100, 40), (260, 50, 141), (128, 48, 7)])
def test_fused_3d_sweep_shapes_equal_per_colour(pmg, ctx, dims, noise, monkeypatch):
    rng = np.random.default_rng(SEED)
    n = dims[0] * dims[1] * dims[2]
    b, y0 = rng.standard_normal(n), rng.standard_normal(n)
    out = []
    for fused in (True, False):
        if fused:
            monkeypatch.delenv("PMG_NO_FUSED", raising=False)
        else:
            monkeypatch.setenv("PMG_NO_FUSED", "1")
        mat = pmg.Mat.laplace(ctx, 3, *dims, kappa=0.7)
        pc = pmg.PC(ctx, "mcgibbs")
        pc.set_operator(mat)
        pc.set_options({"-pc_mcgibbs_omega": 1.3, "-pc_mcgibbs_symmetric": "", "-pc_b200_noise": noise})
        pc.setup()
        if noise == "injected":
            pc.set_noise_tape(np.random.default_rng(SEED + 1).standard_normal(2 * pc.noise_per_sample()))
        ctx.set_seed(11)
        y = y0.copy()
        pc.apply_richardson(b, y, its=2)
        out.append((y, pc.last_stats()["launches"]))
    assert np.array_equal(out[0][0], out[1][0])
    assert out[0][1] != out[1][1]  # two code paths


# ---- the persistent warp-specialised 3D sweep (sweep3d_ws.cuh; device Philox noise; the default without a right-hand side, ----
# ---- PMG_SW3_CFG=7 forces it with one): interior, edge, z-constant and narrow tiles, several z bands, tiles that continue ----
# ---- a CTA's mbarrier phases from the tile before -- bitwise equal to the one-CTA-per-tile kernel and to one launch per colour
@pytest.mark.parametrize("rhs", [False, True])
@pytest.mark.parametrize("dims", [(300, 100, 40), (260, 50, 141), (128, 48, 7), (50, 64, 16), (176, 61, 70), (512, 70, 9)])
def test_persistent_3d_sweep_equals_per_tile_and_per_colour(pmg, ctx, dims, rhs, monkeypatch):
    rng = np.random.default_rng(SEED)
    n = dims[0] * dims[1] * dims[2]
    b, y0 = (rng.standard_normal(n) if rhs else None), rng.standard_normal(n)
    out = []
    for path in ("persistent", "per_tile", "per_colour"):
        monkeypatch.delenv("PMG_NO_FUSED", raising=False)
        monkeypatch.setenv("PMG_SW3_CFG", "7" if path == "persistent" else "6")
        if path == "per_colour":
            monkeypatch.setenv("PMG_NO_FUSED", "1")
        mat = pmg.Mat.laplace(ctx, 3, *dims, kappa=0.7)
        pc = pmg.PC(ctx, "mcgibbs")
        pc.set_operator(mat)
        pc.set_options({"-pc_mcgibbs_omega": 1.3, "-pc_mcgibbs_symmetric": "", "-pc_b200_noise": "philox"})
        pc.setup()
        ctx.set_seed(11)
        y = y0.copy()
        pc.apply_richardson(b, y, its=2)
        out.append(y)
    assert np.array_equal(out[0], out[2])
    assert np.array_equal(out[1], out[2])


# ---- BASELINE config 5: P1 finite elements on data/lshape.msh, 17 ball observations with sigma^2 = 1e-5 as a MATLRC term ------
# ---- (fixture: tests/golden/make_lshape.py; the oracle on the same arrays is pinned by tests/test_oracle.py) ------------------
def _lshape(orc, nref):
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"lshape_config5_r{nref}.npz"))
    return orc.CSR(int(d["rowptr"].size - 1), d["rowptr"], d["col"], d["val"]), d


@pytest.mark.parametrize("nref", [0, 1])
@pytest.mark.parametrize("pctype,sweep", [("mcgibbs", 1), ("mcgibbs", 3), ("sorgibbs", 1)])
def test_config5_lrc_gibbs_matches_oracle(pmg, ctx, orc, nref, pctype, sweep):
    """Unstructured operator, greedy colouring injected on both sides, LRC right-hand side and Woodbury post-correction
    (src/pc_mcgibbs.c:130-140, src/mc_sor.c:101-112, :480-544) with an injected tape.  S = 1e5 makes the k x k system of the
    correction ill-conditioned (cond ~ 1e6), so the sample agrees to 1e-9 rather than to the 1e-12 of a plain sweep."""
    rng = np.random.default_rng(SEED)
    A, d = _lshape(orc, nref)
    n, k = A.n, 17
    B, S, f = d["B"], d["S"], d["f"]
    col = orc.Coloring.greedy(A)
    mat = pmg.Mat.lrc(make_mat(pmg, ctx, A, col), B, S)
    x = rng.standard_normal(n)
    assert relerr(mat.mult(x), A.to_scipy() @ x + B @ (S * (B.T @ x))) < 1e-12
    pc = pmg.PC(ctx, pctype)
    pc.set_operator(mat)
    if pctype == "mcgibbs":
        pc.mcgibbs_set_sweep_type(sweep)
    pc.setup()
    its = 3
    per = pc.noise_per_sample()
    assert per == (n + k) * (2 if sweep == 3 else 1)
    z = rng.standard_normal(its * per)
    pc.set_noise_tape(z)
    y = rng.standard_normal(n)
    ref = orc.lrc_gibbs_richardson(A, B, S, f, y.copy(), its, orc.Noise.tape(z), col, 1.0, sweep)
    pc.apply_richardson(f, y, its=its)
    assert relerr(y, ref) < 1e-9
    # the sweep itself (MCSORApply on the base matrix, no low-rank term) is bit-exact on this operator
    mc = pmg.MCSOR(make_mat(pmg, ctx, A, col))
    mc.set_sweep_type(sweep)
    y2 = rng.standard_normal(n)
    ref2 = orc.MCSOR(A, col, 1.0, sweep).apply(f, y2.copy())
    mc.apply(f, y2)
    assert np.array_equal(y2, ref2)


def test_config5_posterior_mean_and_qoi_device_rng(pmg, ctx, orc):
    """The reference's acceptance for this configuration (examples/benchmark: posterior sampling with -with_lr): the chain
    mean converges to (A + B S B^T)^-1 f; PCWOODBURY with the exact sampler / solver draws independent posterior samples, so
    the QOI of examples/benchmark/lshape.opts:11-13 has IACT 1 and the mean / variance the posterior prescribes."""
    A, d = _lshape(orc, 0)
    n, N = A.n, 20000
    B, S, f, meas = d["B"], d["S"], d["f"], d["meas"]
    mat = pmg.Mat.lrc(pmg.Mat.from_csr(ctx, A.rowptr, A.col, A.val), B, S)
    pc = pmg.PC(ctx, "woodbury")
    pc.set_operator(mat)
    pc.set_options({"-pc_woodbury_sampler": "cholsampler", "-pc_woodbury_solver": "cholesky", "-pc_b200_noise": "philox"})
    pc.setup()
    pc.set_qoi(meas, N, True)
    ctx.set_seed(0xCAFE)
    pc.apply_richardson(f, np.zeros(n), its=N)
    q = pc.get_qoi()
    mean_dev, var_dev, seen = pc.get_mean_var()
    P = A.to_scipy().toarray() + B @ np.diag(S) @ B.T
    Sigma = np.linalg.inv(P)
    mean = Sigma @ f
    assert seen == N
    assert np.linalg.norm(mean_dev - mean) < 4.0 * np.sqrt(np.trace(Sigma) / N)
    assert np.abs(var_dev / np.diag(Sigma) - 1.0).max() < 6.0 * np.sqrt(2.0 / N)
    qm, qv = meas @ mean, meas @ Sigma @ meas
    assert abs(q.mean() - qm) < 4.0 * np.sqrt(qv / N) and abs(q.var(ddof=1) / qv - 1.0) < 5.0 * np.sqrt(2.0 / N)
    tau, ok = pmg.iact(ctx, q)
    assert ok and abs(tau - 1.0) < 0.15


# ---- 27-point Galerkin levels: two colours per launch (box_pair_sweep3_kernel) against one launch per colour ----------------
@pytest.mark.parametrize("noise", ["philox", "tape"])
@pytest.mark.parametrize("dims,levels,extra", [
    ((65, 65, 65), 4, {}),
    ((33, 41, 25), 3, {"-gamgmc_mg_levels_pc_type": "mcgibbs", "-gamgmc_mg_levels_pc_mcgibbs_symmetric": "", "-gamgmc_mg_levels_pc_mcgibbs_omega": 1.3}),
    ((129, 17, 33), 3, {"-gamgmc_mg_levels_ksp_max_it": 2}),
])
def test_box_colour_pair_sweep_is_bit_identical(pmg, ctx, dims, levels, extra, noise, monkeypatch):
    rng = np.random.default_rng(SEED)
    n = dims[0] * dims[1] * dims[2]
    b, y0 = rng.standard_normal(n), rng.standard_normal(n)
    out = []
    monkeypatch.setenv("PMG_NO_BOX3", "1")  # the plane kernels (box3d.cuh) would take over both runs
    for pair in (True, False):
        if pair:
            monkeypatch.delenv("PMG_NO_BOX_PAIR", raising=False)
        else:
            monkeypatch.setenv("PMG_NO_BOX_PAIR", "1")
        lap = pmg.Mat.laplace(ctx, 3, *dims, kappa=0.8)
        pc = pmg.PC(ctx, "gamgmc")
        pc.set_operator(lap)
        pc.set_options(dict(extra, **{"-gamgmc_pc_mg_levels": levels, "-pc_b200_tail_max_n": 0}))
        pc.setup()
        if noise == "tape":
            pc.set_noise_tape(np.random.default_rng(7).standard_normal(2 * pc.noise_per_sample()))
        else:
            pc.set_noise_mode(pmg.NOISE_PHILOX)
            ctx.set_seed(31)
        y = y0.copy()
        pc.apply_richardson(b, y, its=2)
        out.append((y, pc.last_stats()["launches"]))
    assert np.array_equal(out[0][0], out[1][0]), relerr(out[0][0], out[1][0])
    assert out[0][1] < out[1][1]


# ---- 27-point Galerkin levels: four colours per launch, one CTA per plane (box3d.cuh) against one launch per colour --------------
@pytest.mark.parametrize("noise", ["philox", "tape"])
@pytest.mark.parametrize("dims,levels,extra", [
    ((65, 65, 65), 4, {}),
    ((33, 41, 25), 3, {"-gamgmc_mg_levels_pc_type": "mcgibbs", "-gamgmc_mg_levels_pc_mcgibbs_symmetric": "", "-gamgmc_mg_levels_pc_mcgibbs_omega": 1.3}),
    ((129, 17, 33), 3, {"-gamgmc_mg_levels_ksp_max_it": 2}),
    ((21, 37, 19), 3, {"-gamgmc_mg_levels_pc_type": "mcgibbs", "-gamgmc_mg_levels_pc_mcgibbs_backward": ""}),  # odd and even row counts per block
    ((257, 9, 9), 3, {}),
])
def test_box3_plane_sweep_and_residual_are_bit_identical(pmg, ctx, dims, levels, extra, noise, monkeypatch):
    rng = np.random.default_rng(SEED)
    n = dims[0] * dims[1] * dims[2]
    b, y0 = rng.standard_normal(n), rng.standard_normal(n)
    out = []
    for path in ("staged", "plane", "colour"):  # rows staged in shared memory (default) | plain loads | one launch per colour
        monkeypatch.delenv("PMG_NO_BOX3", raising=False)
        monkeypatch.delenv("PMG_BOX3_NO_SMEM", raising=False)
        if path == "plane":
            monkeypatch.setenv("PMG_BOX3_NO_SMEM", "1")
        elif path == "colour":
            monkeypatch.setenv("PMG_NO_BOX3", "1")
            monkeypatch.setenv("PMG_NO_BOX_PAIR", "1")
        lap = pmg.Mat.laplace(ctx, 3, *dims, kappa=0.8)
        pc = pmg.PC(ctx, "gamgmc")
        pc.set_operator(lap)
        pc.set_options(dict(extra, **{"-gamgmc_pc_mg_levels": levels, "-pc_b200_tail_max_n": 0}))
        pc.setup()
        if noise == "tape":
            pc.set_noise_tape(np.random.default_rng(7).standard_normal(2 * pc.noise_per_sample()))
        else:
            pc.set_noise_mode(pmg.NOISE_PHILOX)
            ctx.set_seed(31)
        y = y0.copy()
        pc.apply_richardson(b, y, its=2)
        out.append((y, pc.last_stats()["launches"]))
    assert np.array_equal(out[0][0], out[2][0]), relerr(out[0][0], out[2][0])
    assert np.array_equal(out[1][0], out[2][0]), relerr(out[1][0], out[2][0])
    assert out[0][1] < out[2][1] and out[1][1] < out[2][1]


@pytest.mark.parametrize("dims", [(1025, 14, 5), (257, 22, 7), (33, 70, 9), (61, 30, 9), (2049, 10, 5)])
def test_box3_plane_sweep_block_height_does_not_change_the_result(pmg, ctx, dims, monkeypatch):
    """The number of grid rows per block follows from the row length (level 1 of these grids: 513 nodes -> 1 even row per
    block, 129 -> 7, 17 -> 16): the result must equal the per-colour path for all of them, symmetric sweeps included."""
    rng = np.random.default_rng(SEED)
    n = dims[0] * dims[1] * dims[2]
    b, y0 = rng.standard_normal(n), rng.standard_normal(n)
    out = []
    for path in ("staged", "plane", "colour"):
        monkeypatch.delenv("PMG_NO_BOX3", raising=False)
        monkeypatch.delenv("PMG_BOX3_NO_SMEM", raising=False)
        if path == "plane":
            monkeypatch.setenv("PMG_BOX3_NO_SMEM", "1")
        elif path == "colour":
            monkeypatch.setenv("PMG_NO_BOX3", "1")
            monkeypatch.setenv("PMG_NO_BOX_PAIR", "1")
        lap = pmg.Mat.laplace(ctx, 3, *dims, kappa=0.8)
        pc = pmg.PC(ctx, "gamgmc")
        pc.set_operator(lap)
        pc.set_options({"-gamgmc_pc_mg_levels": 3, "-pc_b200_tail_max_n": 0, "-gamgmc_mg_levels_pc_type": "mcgibbs", "-gamgmc_mg_levels_pc_mcgibbs_symmetric": ""})
        pc.setup()
        pc.set_noise_mode(pmg.NOISE_PHILOX)
        ctx.set_seed(5)
        y = y0.copy()
        pc.apply_richardson(b, y, its=1)
        out.append(y)
    assert np.array_equal(out[0], out[2]), relerr(out[0], out[2])
    assert np.array_equal(out[1], out[2]), relerr(out[1], out[2])


# ---- noise of the 9-point levels written by ONE batched launch at the start of a sample (pc.cu NoisePrefill) and read by the
# ---- one-pass kernels as a tape (opt-in, PMG_PREFILL=1), against generation on the fly inside the kernels: sample 0 records the blocks,
# ---- samples 1.. use the record; several sweeps per level, symmetric sweeps, odd / even row lengths, two calls in a row -----------
@pytest.mark.parametrize("dims,levels,extra", [
    ((513, 257), 5, {"-pc_b200_tail_max_n": 0}),
    ((257, 257), 6, {}),
    ((300, 140), 4, {"-pc_b200_tail_max_n": 0, "-gamgmc_mg_levels_pc_type": "mcgibbs", "-gamgmc_mg_levels_pc_mcgibbs_symmetric": "", "-gamgmc_mg_levels_pc_mcgibbs_omega": 1.4}),
    ((129, 385), 4, {"-pc_b200_tail_max_n": 0, "-gamgmc_mg_levels_ksp_max_it": 2}),
])
def test_prefilled_noise_equals_noise_on_the_fly(pmg, ctx, dims, levels, extra, monkeypatch):
    rng = np.random.default_rng(SEED)
    n = dims[0] * dims[1]
    b, y0 = rng.standard_normal(n), rng.standard_normal(n)
    out = []
    for prefill in (True, False):
        if prefill:
            monkeypatch.setenv("PMG_PREFILL", "1")
        else:
            monkeypatch.delenv("PMG_PREFILL", raising=False)
        lap = pmg.Mat.laplace(ctx, 2, *dims, kappa=0.9)
        pc = pmg.PC(ctx, "gamgmc")
        pc.set_operator(lap)
        pc.set_options(dict(extra, **{"-gamgmc_pc_mg_levels": levels, "-pc_b200_noise": "philox"}))
        pc.setup()
        ctx.set_seed(77)
        y = y0.copy()
        pc.apply_richardson(b, y, its=4)
        pc.apply_richardson(None, y, its=3)  # the record survives from call to call
        out.append((y, pc.last_stats()["launches"], ctx.draw_counter))
    assert np.array_equal(out[0][0], out[1][0]), relerr(out[0][0], out[1][0])
    assert out[0][2] == out[1][2]
    assert out[0][1] == out[1][1] + 3  # one more launch per sample that uses the record


@pytest.mark.parametrize("dims,levels,extra", [
    ((257, 257), 6, {}),                                    # default sizing: levels 0..3 (65^2 and below) in the shared-memory tail
    ((513, 129), 7, {"-gamgmc_mg_levels_ksp_max_it": 2}),
    ((300, 140), 5, {"-gamgmc_mg_levels_pc_type": "mcgibbs", "-gamgmc_mg_levels_pc_mcgibbs_symmetric": "", "-gamgmc_mg_levels_pc_mcgibbs_omega": 0.8}),  # even sizes
])
def test_shared_memory_tail_default_sizing_is_bit_identical(pmg, ctx, dims, levels, extra, monkeypatch):
    """Without -pc_b200_tail_max_n the V-cycle puts every level that fits into the one-CTA tail; the result must equal the
    level-by-level path on the one-pass kernels (PMG_NO_TAIL_SMEM), and use fewer launches."""
    rng = np.random.default_rng(SEED)
    n = dims[0] * dims[1]
    b, y0 = rng.standard_normal(n), rng.standard_normal(n)
    out = []
    for smem in (True, False):
        if smem:
            monkeypatch.delenv("PMG_NO_TAIL_SMEM", raising=False)
        else:
            monkeypatch.setenv("PMG_NO_TAIL_SMEM", "1")
        lap = pmg.Mat.laplace(ctx, 2, *dims, 1, kappa=1.0)
        pc = pmg.PC(ctx, "gamgmc")
        pc.set_operator(lap)
        pc.set_options(dict(extra, **{"-gamgmc_pc_mg_levels": levels}))
        pc.setup()
        pc.set_noise_mode(pmg.NOISE_PHILOX)
        ctx.set_seed(123)
        y = y0.copy()
        pc.apply_richardson(b, y, its=3)
        out.append((y, pc.last_stats()["launches"]))
    assert np.array_equal(out[0][0], out[1][0]), relerr(out[0][0], out[1][0])
    assert out[0][1] < out[1][1]


# ---- the Matern sampler object (src/ms.c) and the Python module (python/main.cc) over the device samplers ------------
def test_ms_object_and_pymgmc_module(pmg, ctx):
    from parmgmc_b200 import pymgmc
    from parmgmc_b200.ms import MS
    import scipy.sparse as sp
    nx, N = 17, 4000
    ms = MS(ctx)
    ms.set_from_options({"-matern_kappa": 6.0, "-ms_gamgmc_pc_mg_levels": 3})
    ms.set_grid(2, nx, nx)
    ms.setup()
    pymgmc.use(ctx)
    pymgmc.seed(77)
    A = ms.get_precision_matrix()
    n = A.n
    # exact covariance of the target N(0, A^-1) from the operator applied to the unit vectors
    dense = np.stack([A.mult(np.eye(n)[:, j].copy()) for j in range(n)], axis=1)
    Sigma = np.linalg.inv(dense)
    x = np.zeros(n)
    ms.set_num_samples(200)
    ms.sample(x)  # burn-in
    ms.set_num_samples(N)
    meas = np.full(n, 1.0 / n)
    ms.set_qoi(lambda it, y: float(meas @ y))
    for keep in (True, False):
        ms.begin_save_samples(keep=keep)
        if keep:
            assert len(ms.get_samples()) == N
        ms.sample(x)
        ms.end_save_samples()
        mean, var = ms.get_mean_and_var()
        q = ms.get_qoi_values()
        assert np.linalg.norm(mean) < 4.5 * np.sqrt(np.trace(Sigma) / N)                  # E y = 0
        assert np.abs(var / np.diag(Sigma) - 1.0).max() < 6.0 * np.sqrt(2.0 / N) * 1.5    # MGMC samples are nearly independent
        qv = meas @ Sigma @ meas
        assert abs(q.mean()) < 4.5 * np.sqrt(qv / N) and abs(q.var(ddof=1) / qv - 1.0) < 0.15
    try:
        ms.get_samples()
        raise AssertionError("MSGetSamples outside a save window must fail (src/ms.c:201)")
    except RuntimeError:
        pass
    # pymgmc: seed() makes the chain reproducible, PCSetSampleCallback sees every sample
    out = []
    for _ in range(2):
        pymgmc.seed(5)
        seen = []
        pymgmc.PCSetSampleCallback(ms.pc, lambda it, y: seen.append((it, float(y[0]))))
        y = np.zeros(n)
        ms.pc.apply_richardson(None, y, its=5)
        out.append((y.copy(), seen))
    assert np.array_equal(out[0][0], out[1][0]) and out[0][1] == out[1][1] and [s[0] for s in out[0][1]] == [0, 1, 2, 3, 4]
