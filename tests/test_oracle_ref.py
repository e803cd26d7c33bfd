"""Pins the oracle restatement against the reference's OWN code.

oracle/_ref/libparmgmc_ref.so is /root/reference/src/{mc_sor.c, pc_mcgibbs.c, parmgmc.c} compiled unmodified against
the container-only PETSc API stub (oracle/petsc_stub).  It exists only where /root/reference does (this container);
elsewhere the committed golden vectors (tests/golden/, written by tests/golden/make_golden.py from the same library)
stand in for it -- see test_golden.py.

Tolerance: the reference writes `sum -= a*y` / `(1-omega)*y + idiag*sum` in plain C, so the compiler decides about
fused multiply-adds; the oracle (and the CUDA path) fix the contraction explicitly (oracle.h "Floating-point
contract").  The two therefore agree to rounding, not bitwise: 1e-13 relative here, against the 1e-12 of north_star.
"""
import numpy as np
import pytest

import oracle as orc
from oracle import ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
RTOL = 1e-13
SEED = 20260625


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("shape,kappa", [((129, 129), 10.0), ((129, 129), 1.0), ((33, 17), 1.0), ((9, 9), 10.0)])
@pytest.mark.parametrize("omega", [1.0, 1.2, 1.6])
@pytest.mark.parametrize("sweep", [orc.SOR_FORWARD, orc.SOR_BACKWARD, orc.SOR_SYMMETRIC])
def test_seq_sweep_one_colour(shape, kappa, omega, sweep):
    """MCSORApply_SEQAIJ with the reference's own 1-rank colouring (config 1: lexicographic Gauss-Seidel)."""
    rng = np.random.default_rng(SEED)
    A = orc.laplace(2, *shape, kappa=kappa)
    assert ref.mcsor_num_colors(A) == 1
    b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
    y_ref = ref.mcsor_apply(A, b, y0.copy(), None, omega, sweep, nsweeps=3)
    mc = orc.MCSOR(A, None, omega, sweep)
    y = y0.copy()
    for _ in range(3):
        mc.apply(b, y)
    assert rel(y, y_ref) < RTOL


@pytest.mark.parametrize("dim,shape", [(2, (129, 129)), (2, (40, 23)), (3, (12, 9, 7))])
@pytest.mark.parametrize("omega,sweep", [(1.0, orc.SOR_FORWARD), (1.3, orc.SOR_SYMMETRIC), (0.7, orc.SOR_BACKWARD)])
def test_seq_sweep_injected_colouring(dim, shape, omega, sweep):
    """The same colouring injected into both sides: red-black, greedy and level-set colourings."""
    rng = np.random.default_rng(SEED + 1)
    A = orc.laplace(dim, *shape, kappa=1.0)
    for col in (orc.Coloring.parity(shape), orc.Coloring.greedy(A), orc.Coloring.levelset(A)):
        assert col.violations(A) == 0
        b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
        y_ref = ref.mcsor_apply(A, b, y0.copy(), col, omega, sweep, nsweeps=2)
        mc = orc.MCSOR(A, col, omega, sweep)
        y = y0.copy()
        mc.apply(b, y)
        mc.apply(b, y)
        assert rel(y, y_ref) < RTOL


def test_seq_sweep_galerkin_operator():
    """A 9-point Galerkin coarse operator (non-constant near the boundary) with 4 colours."""
    rng = np.random.default_rng(SEED + 2)
    mg = orc.MG.geometric(2, 33, 33, 1, 1.0, 3)
    A = mg.level_csr(1)
    col = orc.Coloring.parity(mg.level_dims(1)[:2], 4)
    b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
    y_ref = ref.mcsor_apply(A, b, y0.copy(), col, 1.0, orc.SOR_FORWARD)
    y = orc.MCSOR(A, col, 1.0, orc.SOR_FORWARD).apply(b, y0.copy())
    assert rel(y, y_ref) < RTOL


@pytest.mark.parametrize("nranks", [2, 3, 4])
@pytest.mark.parametrize("omega,sweep", [(1.0, orc.SOR_FORWARD), (1.4, orc.SOR_BACKWARD), (1.2, orc.SOR_SYMMETRIC)])
def test_partitioned_sweep(nranks, omega, sweep):
    """MCSORApply_MPIAIJ + MatCreateScatters (ghost packing order, backward ghost walk) on emulated ranks."""
    rng = np.random.default_rng(SEED + 3)
    shape = (21, 17)
    A = orc.laplace(2, *shape, kappa=1.0)
    col = orc.Coloring.parity(shape)
    cuts = np.sort(rng.choice(np.arange(1, A.n), nranks - 1, replace=False))  # ragged row blocks, not aligned to grid rows
    rowstart = np.concatenate([[0], cuts, [A.n]])
    b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
    y_ref = ref.mcsor_apply_mpi(A, rowstart, col, b, y0.copy(), omega, sweep, nsweeps=2)
    part = orc.Partitioned(A, rowstart, col, omega)
    y = y0.copy()
    part.sweep(b, y, sweep)
    part.sweep(b, y, sweep)
    assert rel(y, y_ref) < RTOL
    # a distributed multicolour sweep equals the sequential one with the same colouring
    y_seq = ref.mcsor_apply(A, b, y0.copy(), col, omega, sweep, nsweeps=2)
    assert rel(y_ref, y_seq) < RTOL


def test_box_muller_stream():
    """VecSetRandomStandardNormal, Box-Muller branch (src/parmgmc.c:100-110): pairing, odd tail, call order."""
    for n in (1, 2, 7, 1000):
        z_ref = ref.normal_fill(0x12345678, n, ncalls=3)
        ns = orc.Noise.rander48(0x12345678)
        z = np.concatenate([orc.noise_fill(ns, n) for _ in range(3)])
        assert np.abs(z - z_ref).max() < 1e-14 * max(1.0, np.abs(z_ref).max())


@pytest.mark.parametrize("omega,sweep_opt,sweep", [(None, "", orc.SOR_FORWARD), (1.2, "-pc_mcgibbs_backward", orc.SOR_BACKWARD), (1.6, "-pc_mcgibbs_symmetric", orc.SOR_SYMMETRIC)])
@pytest.mark.parametrize("colored", [False, True])
def test_mcgibbs_richardson(omega, sweep_opt, sweep, colored):
    """PCMCGIBBS end to end: options, sqrtdiag, PrepareRHS_Default, the noise tape order (symmetric = two fills per
    sample), the callback after every sample (src/pc_mcgibbs.c:119-188)."""
    rng = np.random.default_rng(SEED + 4)
    shape = (33, 21)
    A = orc.laplace(2, *shape, kappa=10.0)
    col = orc.Coloring.parity(shape) if colored else None
    b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
    seen_ref, seen = [], []
    y_ref = ref.mcgibbs_richardson(A, b, y0.copy(), 4, 777, col, omega, sweep_opt, callback=lambda it, y: seen_ref.append((it, y.copy())))
    y = orc.gibbs_richardson(A, b, y0.copy(), 4, orc.Noise.rander48(777), col, 1.0 if omega is None else omega, sweep, callback=lambda it, y: seen.append((it, y.copy())))
    assert rel(y, y_ref) < RTOL
    assert [i for i, _ in seen] == [i for i, _ in seen_ref] == [0, 1, 2, 3]
    for (_, a), (_, r) in zip(seen, seen_ref):
        assert rel(a, r) < RTOL
    # b = NULL (prior sampling, examples/ex8.c:47-49)
    y_ref = ref.mcgibbs_richardson(A, None, y0.copy(), 2, 5, col, omega, sweep_opt)
    y = orc.gibbs_richardson(A, None, y0.copy(), 2, orc.Noise.rander48(5), col, 1.0 if omega is None else omega, sweep)
    assert rel(y, y_ref) < RTOL


def test_ex5_identity_on_reference():
    """examples/ex5.c:60-70 on the reference itself: (forward; backward) == symmetric."""
    rng = np.random.default_rng(SEED + 5)
    A = orc.laplace(2, 9, 9, kappa=1.0)
    b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
    y1 = ref.mcsor_apply(A, b, y0.copy(), None, 1.0, orc.SOR_FORWARD)
    ref.mcsor_apply(A, b, y1, None, 1.0, orc.SOR_BACKWARD)
    y2 = ref.mcsor_apply(A, b, y0.copy(), None, 1.0, orc.SOR_SYMMETRIC)
    assert np.linalg.norm(y1 - y2) < 1e-15


# ---- round 2: pc_sorgibbs.c, pc_chols.c (dense LAPACK branch), iact.c, stats.c and the MATLRC branches of mc_sor.c /
#      pc_mcgibbs.c / pc_sorgibbs.c / pc_chols.c compiled into oracle/_ref as well ----------------------------------------
def test_sorgibbs_richardson_and_apply():
    """PCSORGIBBS on one rank = w = b + sqrt(a_ii) z, then PETSc's MatSOR forward sweep at omega = 1 (src/pc_sorgibbs.c:76-103);
    the callback counts samples from 0 per call (:125-129); PCApply zeroes y first (:105-113)."""
    rng = np.random.default_rng(SEED + 10)
    A = orc.laplace(2, 33, 21, kappa=3.0)
    b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
    seen_ref, seen = [], []
    y_ref = ref.sampler_run("sorgibbs", A, b, y0.copy(), 4, 4711, callback=lambda it, y: seen_ref.append((it, y.copy())))
    y = orc.gibbs_richardson(A, b, y0.copy(), 4, orc.Noise.rander48(4711), None, 1.0, orc.SOR_FORWARD, callback=lambda it, y: seen.append((it, y.copy())))
    assert rel(y, y_ref) < RTOL
    assert [i for i, _ in seen_ref] == [0, 1, 2, 3]
    for (_, a), (_, r) in zip(seen, seen_ref):
        assert rel(a, r) < RTOL
    y_ref = ref.sampler_run("sorgibbs", A, b, y0.copy(), 0, 99)  # PCApply
    y = orc.gibbs_richardson(A, b, np.zeros(A.n), 1, orc.Noise.rander48(99), None, 1.0, orc.SOR_FORWARD)
    assert rel(y, y_ref) < RTOL


@pytest.mark.parametrize("shape,its", [((7, 9), 1), ((7, 9), 3), ((8, 8), 1)])
def test_cholsampler_dense_branch(shape, its):
    """PCCHOLSAMPLER, dense fast path (n <= 64: LAPACK potrf + two trsv, src/pc_chols.c:173-195, :220-291); its > 1 caches the
    forward solve (:306-336)."""
    rng = np.random.default_rng(SEED + 11)
    A = orc.laplace(2, *shape, kappa=2.0)
    b = rng.standard_normal(A.n)
    y_ref = ref.sampler_run("cholsampler", A, b, np.zeros(A.n), its, 31337)
    lflat = orc.potrf_lower(A.to_scipy().toarray())
    ns = orc.Noise.rander48(31337)
    y = None
    for _ in range(its):  # the cached forward solve does not change the arithmetic
        y = orc.chol_sample(lflat, A.n, ns, b)
    assert rel(y, y_ref) < 1e-12
    # the callback-less PCApply path is the same sample (:262-291)
    y_ref0 = ref.sampler_run("cholsampler", A, b, np.zeros(A.n), 0, 31337)
    assert rel(orc.chol_sample(lflat, A.n, orc.Noise.rander48(31337), b), y_ref0) < 1e-12


def test_cholsampler_not_spd_error():
    A = orc.laplace(2, 5, 5, kappa=1.0)
    A.val[A.diag_ptrs()] = -1.0
    with pytest.raises(RuntimeError, match="not positive definite"):
        ref.sampler_run("cholsampler", A, np.ones(A.n), np.zeros(A.n), 1, 1)


@pytest.mark.parametrize("n", [2, 17, 500, 4096, 5000])
def test_iact_and_autocorrelation(n):
    """src/iact.c:17-92: FFT autocorrelation (zero-padded to 2 nextpow2(n), normalised by lag 0) and Sokal's window (c = 5)."""
    rng = np.random.default_rng(SEED + n)
    x = np.empty(n)
    x[0] = rng.standard_normal()
    for i in range(1, n):  # AR(1), tau = (1 + rho) / (1 - rho)
        x[i] = 0.8 * x[i - 1] + rng.standard_normal()
    acf_ref = ref.autocorrelation(x)
    assert np.abs(orc.autocorrelation(x) - acf_ref).max() < 1e-12
    tau_ref, valid_ref, acf2 = ref.iact(x)
    tau, valid = orc.iact(x)
    assert abs(tau - tau_ref) < 1e-10 * max(1.0, abs(tau_ref)) and valid == valid_ref
    assert np.abs(acf2 - acf_ref).max() == 0.0


def test_cov_errors():
    """src/stats.c:94-117: || C_i - A^-1 ||_F / || A^-1 ||_F across chains for every sample index."""
    rng = np.random.default_rng(SEED + 12)
    A = orc.laplace(2, 4, 5, kappa=1.5)
    samples = rng.standard_normal((6, 9, A.n))
    ref_errs = ref.cov_errors(A, samples)
    errs = orc.cov_errors(A.to_scipy().toarray(), samples)
    assert np.abs(errs - ref_errs).max() < 1e-12 * np.abs(ref_errs).max()


def _lrc_problem(rng, shape=(13, 11), k=4):
    A = orc.laplace(2, *shape, kappa=2.0)
    B = rng.standard_normal((A.n, k)) * (rng.random((A.n, k)) < 0.2)
    S = 1.0 + 10.0 * rng.random(k)
    return A, B, S


@pytest.mark.parametrize("omega,sweep", [(1.0, orc.SOR_FORWARD), (1.0, orc.SOR_SYMMETRIC), (1.0, orc.SOR_BACKWARD)])
def test_mcsor_apply_on_matlrc(omega, sweep):
    """MCSORApply on A + B S B^T: the sweep on A, then y -= Bb_dir (B^T y) with Bb from MCSORBuildLRCCorrection
    (src/mc_sor.c:101-112, :480-544; built with the MCSOR's omega at set-up time, :583-593).  One rank, the reference's own
    one-colour ordering (an injected colouring needs the two-rank branch, whose set-up scatter of S is collective)."""
    rng = np.random.default_rng(SEED + 13)
    A, B, S = _lrc_problem(rng)
    col = None
    b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
    y_ref = ref.mcsor_apply_lrc(A, B, S, b, y0.copy(), col, omega, sweep, nsweeps=2)
    Bb = {d: orc.lrc_build_correction(A, B, S, col, 1.0, d) for d in (orc.SOR_FORWARD, orc.SOR_BACKWARD)}
    y = y0.copy()
    for _ in range(2):
        orc.lrc_mcsor_apply(A, B, Bb, b, y, col, omega, sweep)
    assert rel(y, y_ref) < 1e-11


@pytest.mark.parametrize("pctype,opts,omega,sweep", [
    ("mcgibbs", (), 1.0, orc.SOR_FORWARD),
    ("mcgibbs", (("-pc_mcgibbs_omega", 1.3), ("-pc_mcgibbs_symmetric", "")), 1.3, orc.SOR_SYMMETRIC),
    ("sorgibbs", (), 1.0, orc.SOR_FORWARD),
])
def test_gibbs_samplers_on_matlrc(pctype, opts, omega, sweep):
    """PrepareRHS_LRC (n draws, then k; src/pc_mcgibbs.c:130-140, src/pc_sorgibbs.c:86-90) + sweep + post-correction, end to end."""
    rng = np.random.default_rng(SEED + 14)
    A, B, S = _lrc_problem(rng)
    b, y0 = rng.standard_normal(A.n), rng.standard_normal(A.n)
    y_ref = ref.sampler_run(pctype, A, b, y0.copy(), 3, 2024, opts=opts, lrc=(B, S))
    y = orc.lrc_gibbs_richardson(A, B, S, b, y0.copy(), 3, orc.Noise.rander48(2024), None, omega, sweep, omega_build=1.0)
    assert rel(y, y_ref) < 1e-11


def test_cholsampler_on_matlrc():
    """PCCHOLSAMPLER factors A + B S B^T (src/pc_chols.c:119-157)."""
    rng = np.random.default_rng(SEED + 15)
    A, B, S = _lrc_problem(rng, (6, 7), 3)
    b = rng.standard_normal(A.n)
    y_ref = ref.sampler_run("cholsampler", A, b, np.zeros(A.n), 1, 77, lrc=(B, S))
    P = A.to_scipy().toarray() + B @ np.diag(S) @ B.T
    y = orc.chol_sample(orc.potrf_lower(P), A.n, orc.Noise.rander48(77), b)
    assert rel(y, y_ref) < 1e-11


# ---- PCWOODBURY: the reference's own src/woodbury.c compiled into oracle/_ref (stub additions: PCApply / PCApplyRichardson /
# ---- PCSetFromOptions / prefix functions on member PCs, a built-in exact "cholesky" solver type) ------------------------------
@pytest.mark.parametrize("sampler,its", [("mcgibbs", 1), ("mcgibbs", 3), ("sorgibbs", 2), ("cholsampler", 2)])
def test_woodbury_sampler_matches_the_restatement(sampler, its):
    """PCSetUp_Woodbury + PCWoodburyBuildLRCCorrection (src/woodbury.c:21-86, :141-186: C = solver(B) column by column,
    G = C (S^-1 + B^T C)^-1) and PCApplyRichardson_Woodbury (:259-286): per sample the k normals of the observation noise are
    drawn FIRST, w = b + B (sqrt|S| . eta), one Richardson step of the sampler on A with right-hand side w, then
    y -= G (B^T y).  The restatement below is the one tests/test_gpu_parity.py holds the CUDA path against."""
    rng = np.random.default_rng(SEED + 21)
    A, B, S = _lrc_problem(rng, (9, 8) if sampler == "cholsampler" else (13, 11), 3)
    n, k = A.n, B.shape[1]
    b, y0 = rng.standard_normal(n), rng.standard_normal(n)
    seen_ref = []
    y_ref = ref.sampler_run("woodbury", A, b, y0.copy(), its, 5150, opts=[("-pc_woodbury_solver", "cholesky"), ("-pc_woodbury_sampler", sampler)], lrc=(B, S),
                            callback=lambda it, y: seen_ref.append(y.copy()))
    Ad = A.to_scipy().toarray()
    Cm = np.linalg.solve(Ad, B)
    G = Cm @ np.linalg.inv(np.diag(1.0 / S) + B.T @ Cm)
    noise = orc.Noise.rander48(5150)
    lflat = orc.potrf_lower(Ad) if sampler == "cholsampler" else None
    y, seen = y0.copy(), []
    for _ in range(its):
        w = b + B @ (np.sqrt(np.abs(S)) * orc.noise_fill(noise, k))
        if sampler == "cholsampler":
            y = orc.chol_sample(lflat, n, noise, w)  # an exact sampler ignores the previous state
        else:
            y = orc.gibbs_richardson(A, w, y, 1, noise, None, 1.0, orc.SOR_FORWARD)
        y = y - G @ (B.T @ y)
        seen.append(y.copy())
    assert rel(y, y_ref) < 1e-10
    assert len(seen_ref) == its and all(rel(a, r) < 1e-10 for a, r in zip(seen, seen_ref))


def test_woodbury_needs_a_matlrc_operator_and_both_members():
    """src/woodbury.c:149, :159: without solver / sampler, or on a plain AIJ operator, set-up fails."""
    A = orc.laplace(2, 5, 5, kappa=1.0)
    with pytest.raises(RuntimeError, match="sampler and solver"):
        ref.sampler_run("woodbury", A, np.ones(A.n), np.zeros(A.n), 1, 1, lrc=(np.ones((A.n, 1)), np.ones(1)))
    with pytest.raises(RuntimeError, match="LRC"):
        ref.sampler_run("woodbury", A, np.ones(A.n), np.zeros(A.n), 1, 1, opts=[("-pc_woodbury_solver", "cholesky"), ("-pc_woodbury_sampler", "sorgibbs")])


@pytest.mark.parametrize("mx,my,kappa", [(5, 5, 1.0), (9, 6, 2.5), (17, 33, 0.3), (2, 7, 1.0)])
def test_operator_assembly_matches_problems_c(mx, my, kappa):
    """MatAssembleShiftedLaplaceFD (src/problems.c:14-75): off-diagonals -h with h = 1 / (mx - 1)^2 (the reference's `hinv2`, in both
    directions), diagonal kappa^2 plus h once per existing neighbour, added in the order south, west, north, east.  The oracle's
    operator (and with it every matrix the parity tests hand to both sides) must equal it entry for entry, bit for bit."""
    ref_dense = ref.assemble_laplace2d(mx, my, kappa)
    A = orc.laplace(2, mx, my, kappa=kappa)
    mine = A.to_scipy().toarray()
    assert np.array_equal(mine, ref_dense)
    assert np.array_equal(mine != 0, ref_dense != 0)
